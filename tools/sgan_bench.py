"""SGAN method (row f3) at the age configuration's shapes (BASELINE configs[1]: DCGAN pair, 3x128x128, B=100, 10 bins): ms/step of
the B200 path (bf16 and fp32 modes) and the oracle's step (torch autograd on the host cores, bounded sample) beside it."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import srgan_b200
from oracle import srgan_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
bins = tuple(torch.linspace(10, 95, 10).tolist())
gen = torch.Generator().manual_seed(0)
x, u = torch.rand(B, 3, 128, 128, generator=gen) * 2 - 1, torch.rand(B, 3, 128, 128, generator=gen) * 2 - 1
y = torch.rand(B, generator=gen) * 85 + 10
for precision in ('bf16', 'fp32'):
    s = srgan_b200.Settings()
    s.batch_size, s.matching_loss_multiplier, s.gradient_penalty_multiplier, s.precision, s.bins = B, 1.0, 1e2, precision, bins
    D = srgan_b200.DcganDiscriminator(128, 64, 10).cuda()
    DNN = srgan_b200.DcganDiscriminator(128, 64, 10).cuda()
    G = srgan_b200.DcganGenerator(256, 128, 64).cuda()
    r = srgan_b200.StepRunner(D, G, DNN, s, 'sgan', precision=precision)
    xc, yc, uc = x.cuda(), y.cuda(), u.cuda()
    for i in range(5):
        r.dnn_step(xc, yc)
        r.gan_step(xc, yc, uc, i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    N = 50
    e0.record()
    for i in range(N):
        r.dnn_step(xc, yc)
        r.gan_step(xc, yc, uc, i)
    e1.record()
    torch.cuda.synchronize()
    print(f'age SGAN (10 bins) B={B} {precision}: {e0.elapsed_time(e1) / N:.3f} ms/step, tcgen05 contraction calls '
          f'{r.engine.ops.tensor_launches}, scalars {r.scalars()}', flush=True)
# the oracle's step on the host cores, 10-sample batches (scaled)
torch.set_num_threads(os.cpu_count())
st = O.init_dcgan(seed=0, n_out=10)
cfg = O.StepConfig(method='sgan', batch_size=10, matching_loss_multiplier=1.0, gradient_penalty_multiplier=1e2, bins=bins)
z, a, z2 = torch.randn(10, 256), torch.rand(10, 1, 1, 1), torch.randn(10, 256)
O.training_step(st, cfg, x[:10], y[:10], u[:10], z, a, z2, 0)
t0 = time.perf_counter()
for i in range(3):
    O.training_step(st, cfg, x[:10], y[:10], u[:10], z, a, z2, i)
dt = (time.perf_counter() - t0) / 3
print(f'oracle (torch autograd, {os.cpu_count()} host threads) on 10-sample batches: {dt * 1e3:.0f} ms/step -> {dt * B / 10 * 1e3:.0f} ms per {B}-sample step')
