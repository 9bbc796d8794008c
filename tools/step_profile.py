"""Host-vs-device split of one step (GPU box): host enqueue time, device time, per-phase device time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import srgan_b200
import bench

B = 100
s = srgan_b200.Settings()
s.batch_size = B
s.matching_loss_multiplier, s.contrasting_loss_multiplier, s.gradient_penalty_multiplier = 1e2, 1e1, 1e2
s.precision = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
exp = srgan_b200.Experiment(s, 'age', image_size=128, conv_dim=64, z_dim=256)
x, y, u = (t.cuda() for t in bench.make_batches(B, 1, 128))
for i in range(5):
    exp.dnn_training_step(x, y, i); exp.gan_training_step(x, y, u, i)
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
for i in range(N):
    exp.dnn_training_step(x, y, i); exp.gan_training_step(x, y, u, i)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f'host enqueue {1e3 * (t1 - t0) / N:.3f} ms/step, total {1e3 * (t2 - t0) / N:.3f} ms/step, launches/step {exp.runner.engine.ops.launches / (N + 5):.0f}')
# phases with events
ev = lambda: torch.cuda.Event(enable_timing=True)
e = [ev() for _ in range(3)]
torch.cuda.synchronize()
e[0].record(); exp.dnn_training_step(x, y, 0); e[1].record(); exp.gan_training_step(x, y, u, 0); e[2].record()
torch.cuda.synchronize()
print(f'dnn step {e[0].elapsed_time(e[1]):.3f} ms, gan step {e[1].elapsed_time(e[2]):.3f} ms (single shot, includes host gaps)')
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(5):
    exp.dnn_training_step(x, y, i); exp.gan_training_step(x, y, u, i)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
