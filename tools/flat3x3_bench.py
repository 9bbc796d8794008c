"""Dense-layer conv2 (3x3, 128 -> 32 channels written into a concat window) forward at the crowd trunk's shapes, CUDA-event timed,
algorithmic GB/s = (input + output bytes) / time.  SRGAN_NO_FLAT3X3=1 selects the tap-per-stage kernel for comparison.
usage: python tools/flat3x3_bench.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srgan_b200.nets import Geom
from srgan_b200.ops_cuda import CudaOps

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ops = CudaOps()
ops.begin()
dt = torch.bfloat16
for hw, n, pitch in ((56, 256, 256), (28, 256, 512), (14, 256, 1792), (7, 256, 1920), (56, 64, 256), (14, 64, 1792), (7, 64, 1920)):
    g = Geom(hw, hw, 64, hw, hw, 128, 3, 3, 1, 1)
    pix = n * hw * hw
    L = [torch.randn(pix * 128, device='cuda').to(dt) for _ in range(3)]
    cat = [torch.zeros(pix * pitch, device='cuda', dtype=dt) for _ in range(3)]
    Wd = (torch.randn(64 * 9 * 128, device='cuda') * 0.05).to(dt)
    vw = (pitch, 32, 0, 0)
    ts = []
    for i in range(iters + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k = i % 3
        e0.record(); ops.conv_down(L[k], Wd, cat[k][pitch - 32:], n, g, None, 0, None, 0, 0, 0.0, views=vw); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts[1:])[len(ts[1:]) // 2]
    nbytes = pix * (128 + 32) * 2
    print(f'{hw:3d}x{hw:<3d} n={n:4d}: {t * 1e3:7.1f} us  {nbytes / 1e6:7.1f} MB algorithmic  {nbytes / t / 1e6:6.0f} GB/s', flush=True)
    del L, cat
