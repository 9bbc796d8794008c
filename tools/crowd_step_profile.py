"""One eager crowd SR-GAN step (BASELINE configs[2], per-GPU batch B) between cudaProfilerStart / Stop, for
`ncu --profile-from-start off` launch lists (every kernel of the step is its own row: the timed bench replays the same
kernels from CUDA graphs).  usage: ncu --profile-from-start off ... python tools/crowd_step_profile.py [B] [workload]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import srgan_b200
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
name = sys.argv[2] if len(sys.argv) > 2 else 'crowd'
s = srgan_b200.Settings()
s.batch_size, s.precision, s.use_cuda_graph = B, 'bf16', False
s.matching_loss_multiplier, s.contrasting_loss_multiplier, s.gradient_penalty_multiplier = bench.WORKLOADS[name]['mult']
s.map_multiplier = 1e-3
exp = srgan_b200.Experiment(s, name)
exp.runner.overlap_dnn = False
exp.runner.engine.wgrad_side_stream = exp.runner.engine.branch_streams = False       # one stream: rows in execution order
x, y, u = bench.make_batches(name, B, 3)
x, u = x.cuda(), u.cuda()
y = tuple(t.cuda() for t in y) if isinstance(y, tuple) else y.cuda()
for i in range(2):
    exp.dnn_training_step(x, y, i); exp.gan_training_step(x, y, u, i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
exp.dnn_training_step(x, y, 2); exp.gan_training_step(x, y, u, 2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('launches in the profiled step: see the ncu list; tcgen05 calls so far', exp.runner.engine.ops.tensor_launches)
