"""Stand-alone timing of the crowd BN-affine streaming kernels at in-situ shapes (14x14 stage, 4B = 256 samples, a
1024-channel slice of the 1792-wide concat buffer), through the C ABI, CUDA events around each launch, buffers rotated so
that every launch reads from HBM.  Also the ncu target for these kernels (one small process instead of the whole crowd step).
usage: python tools/stream_bench.py [rows] [C] [pitch] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srgan_b200.ops_cuda import CudaOps

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 256 * 196
C = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
pitch = int(sys.argv[3]) if len(sys.argv) > 3 else 1792
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dev = torch.device('cuda:0')
ops = CudaOps(dev)
ops.begin()
bf = torch.bfloat16
NB = 4                                               # rotating buffer sets: 4 x (180 + 103 + 180) MB >> 126 MB of L2
cat = [torch.randn(rows * pitch, device=dev).to(bf) for _ in range(NB)]
dcat = [torch.randn(rows * pitch, device=dev).to(bf) for _ in range(NB)]
n1 = [torch.empty(rows * C, device=dev, dtype=bf) for _ in range(NB)]
dn1 = [torch.randn(rows * C, device=dev).to(bf) for _ in range(NB)]
gamma, beta, mean = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev), torch.randn(C, device=dev)
var = torch.rand(C, device=dev) + 0.5
dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
e = 2


def timed(name, nbytes, fn):
    for i in range(2):
        fn(i % NB)
    torch.cuda.synchronize()
    ts = []
    for i in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(i % NB); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    t = ts[len(ts) // 2]
    print(f'{name:28s} rows={rows} C={C} pitch={pitch}: {t * 1e3:8.1f} us  {nbytes / 1e6:8.1f} MB  {nbytes / t / 1e6:7.0f} GB/s', flush=True)


timed('affine fwd (BN+ReLU)', rows * C * e * 2,
      lambda k: ops.affine(cat[k], pitch, 0, n1[k], C, rows, C, gamma, beta, mean, var, 1e-5, None, 0, 1, 0.0))
timed('affine tangent', rows * C * e * 3,
      lambda k: ops.affine(cat[k], pitch, 0, n1[k], C, rows, C, gamma, beta, mean, var, 1e-5, dn1[k], 1, 1, 0.0))
timed('affine_bwd_grad (accumulate)', rows * C * e * 4,
      lambda k: ops.affine_bwd_grad(dn1[k], C, cat[k], dcat[k], pitch, 0, rows, C, gamma, mean, var, 1e-5, dg, db, True))
timed('affine_bwd (accumulate)', rows * C * e * 3,
      lambda k: ops.affine_bwd(dn1[k], C, dcat[k], pitch, 0, rows, C, gamma, var, 1e-5, True))
timed('copy2d slice write', rows * 32 * e * 2,
      lambda k: ops.copy2d(dn1[k], C, 0, cat[k], pitch, 1024 if pitch >= 1056 else 0, rows, 32, False))
