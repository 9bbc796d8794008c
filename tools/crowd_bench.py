"""Crowd SR-GAN step (BASELINE configs[2]: DCGenerator + KnnDenseNetCat/DenseNet-201 at 224x224) on one GPU: ms/step,
launches/step, peak memory.  usage: python tools/crowd_bench.py [B] [precision] [steps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import srgan_b200
from oracle import srgan_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
precision = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
s = srgan_b200.Settings()
s.batch_size, s.precision = B, precision
s.matching_loss_multiplier, s.contrasting_loss_multiplier, s.gradient_penalty_multiplier, s.map_multiplier = 1e3, 1e2, 1e2, 1e-3
s.use_cuda_graph = os.environ.get('SRGAN_NO_GRAPH', '0') != '1'
exp = srgan_b200.Experiment(s, 'crowd')
x, y, u, z, alpha, z2 = O.synthetic_crowd_batch(B, 3)
x, u, y = x.cuda(), u.cuda(), tuple(t.cuda() for t in y)
t0 = time.perf_counter()
for i in range(3):
    exp.dnn_training_step(x, y, i); exp.gan_training_step(x, y, u, i)
torch.cuda.synchronize()
print(f'warm-up (3 steps incl. graph capture): {time.perf_counter() - t0:.2f} s', flush=True)
l0 = exp.runner.engine.ops.launches
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
prof = os.environ.get('SRGAN_PROFILE', '0') == '1'       # ncu --profile-from-start off: only the timed steps are captured
if prof:
    torch.cuda.profiler.start()
e0.record()
for i in range(steps):
    exp.dnn_training_step(x, y, 3 + i); exp.gan_training_step(x, y, u, 3 + i)
e1.record()
torch.cuda.synchronize()
if prof:
    torch.cuda.profiler.stop()
print(f'launches/step {(exp.runner.engine.ops.launches - l0) / steps:.0f}')
ms = e0.elapsed_time(e1) / steps
flops = 193.8e9 * B
print(f'crowd srgan B={B} {precision} graph={exp.runner.use_cuda_graph}: {ms:.2f} ms/step, {1e3 / ms:.2f} steps/s, '
      f'{B * 1e3 / ms:.1f} samples/s, {flops / ms / 1e9:.1f} algorithmic TFLOP/s, '
      f'peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, scalars {exp.runner.scalars()}', flush=True)
