"""Device time of the persistent coefficient kernel alone (dnn + gan launches), without the host-side step plumbing:
eager back-to-back launches and the same two launches replayed from a CUDA graph."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import srgan_b200

B = 5000
s = srgan_b200.Settings()
s.batch_size, s.gradient_penalty_multiplier = B, 10.0
exp = srgan_b200.Experiment(s, 'coefficient', sys.argv[1] if len(sys.argv) > 1 else 'srgan')
r = exp.runner
gen = torch.Generator().manual_seed(0)
x, u = torch.randn(B, 50, generator=gen).cuda(), torch.randn(B, 50, generator=gen).cuda()
y = (torch.rand(B, generator=gen) * 2 - 1).cuda()
z, alpha, z2 = r.draw_noise(B, r.config())
def both():
    r._coef_step(1, x, y, lr_dnn=1e-4)
    r._coef_step(2, x, y, u, z, alpha.reshape(-1), z2)
for _ in range(5):
    both()
torch.cuda.synchronize()
N = 300
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(N):
    both()
e1.record(); torch.cuda.synchronize()
print(f'eager: {e0.elapsed_time(e1) / N * 1e3:.1f} us/step device, host enqueue+wait {(time.perf_counter() - t0) / N * 1e6:.1f} us/step')
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    both()
torch.cuda.synchronize()
e0.record()
for _ in range(N):
    g.replay()
e1.record(); torch.cuda.synchronize()
print(f'graph replay of the two launches: {e0.elapsed_time(e1) / N * 1e3:.1f} us/step')
for ph, name in ((1, 'dnn'), (2, 'gan')):
    e0.record()
    for _ in range(N):
        if ph == 1: r._coef_step(1, x, y, lr_dnn=1e-4)
        else: r._coef_step(2, x, y, u, z, alpha.reshape(-1), z2)
    e1.record(); torch.cuda.synchronize()
    print(f'{name} launch alone (eager loop): {e0.elapsed_time(e1) / N * 1e3:.1f} us')
