"""The crowd trunk's 1x1 GEMMs at an in-situ shape (14x14 stage, 4B = 256 samples -> 50176 rows, 128 bottleneck channels,
C concat channels): conv1 forward [rows x C] x [C x 128], its data gradient [rows x 128] x [128 x C] (+ ReLU mask of
n1) and its weight gradient, through the C ABI, CUDA-event timed with algorithmic GB/s.  Also the ncu target for these
HBM-bound tcgen05 launches.  usage: python tools/trunk_gemm_bench.py [rows] [C] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srgan_b200.nets import Geom
from srgan_b200.ops_cuda import CudaOps

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 256 * 196
C = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ops = CudaOps()
ops.begin()
dt = torch.bfloat16
g = Geom(1, 1, 128, 1, 1, C, 1, 1, 1, 0)
NB = 3                                              # rotate buffer sets: every launch reads from HBM
n1 = [(torch.rand(rows * C, device='cuda') * 2 - 1).to(dt) for _ in range(NB)]
db = [(torch.rand(rows * 128, device='cuda') * 2 - 1).to(dt) for _ in range(NB)]
dn1 = [torch.empty(rows * C, device='cuda', dtype=dt) for _ in range(NB)]
b = [torch.empty(rows * 128, device='cuda', dtype=dt) for _ in range(NB)]
Wd = ((torch.rand(128 * C, device='cuda') * 2 - 1) * 0.05).to(dt)
Wu = Wd.view(128, C).t().contiguous().view(-1)
dW = torch.zeros(128 * C, device='cuda')
e = 2
cases = (('conv1 forward   (down)', rows * (C + 128) * e, lambda k: ops.conv_down(n1[k], Wd, b[k], rows, g, None, 0, None, 0, 0, 0.0)),
         ('conv1 data grad (up)  ', rows * (128 + 2 * C) * e, lambda k: ops.conv_up(db[k], Wu, dn1[k], rows, g, None, 0, n1[k], 1, 1, 0.0)),
         ('conv1 weight grad     ', rows * (C + 128) * e, lambda k: ops.conv_wgrad(db[k], n1[k], dW, rows, g)))
for name, nbytes, fn in cases:
    fn(0)
    torch.cuda.synchronize()
    ts = []
    for i in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn((i + 1) % NB); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    t = ts[len(ts) // 2]
    flops = 2.0 * rows * C * 128
    print(f'{name} rows={rows} C={C}: {t * 1e3:7.1f} us  {nbytes / 1e6:7.1f} MB  {nbytes / t / 1e6:6.0f} GB/s  {flops / t / 1e9:6.0f} TFLOP/s  '
          f'tensor={ops.lib.srgan_last_path_tensor()}', flush=True)
