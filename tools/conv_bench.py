"""Micro-benchmark (GPU box): the tcgen05 conv-pair kernels on the age-config layer shapes, CUDA-event timed.
usage: python tools/conv_bench.py [filter-substring] [--iters N]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srgan_b200.nets import Geom
from srgan_b200.ops_cuda import CudaOps

ops = CudaOps()
args = sys.argv[1:]
iters = 10
if '--iters' in args:
    k = args.index('--iters')
    iters = int(args[k + 1])
    del args[k:k + 2]
filt = [a for a in args if not a.startswith('--')]
SHAPES = [
    ('D.l2 64->128 @64->32', Geom(32, 32, 128, 64, 64, 64, 4, 4, 2, 1), 400),
    ('D.l3 128->256 @32->16', Geom(16, 16, 256, 32, 32, 128, 4, 4, 2, 1), 400),
    ('D.l4 256->512 @16->8', Geom(8, 8, 512, 16, 16, 256, 4, 4, 2, 1), 400),
    ('G.l1 512->256 @8->16', Geom(8, 8, 512, 16, 16, 256, 4, 4, 2, 1), 100),
    ('G.l3 128->64 @32->64', Geom(32, 32, 128, 64, 64, 64, 4, 4, 2, 1), 100),
    ('thin GEMM 64x64', Geom(1, 1, 64, 1, 1, 64, 1, 1, 1, 0), 400 * 4096),
    ('fc 256->32768', Geom(1, 1, 32768, 1, 1, 256, 1, 1, 1, 0), 100),
]
dt = torch.bfloat16
for name, g, n in SHAPES:
    if filt and not any(f in name for f in filt):
        continue
    L = (torch.rand(n * g.Hl * g.Wl * g.Cb, device='cuda') * 2 - 1).to(dt)
    S = (torch.rand(n * g.Hs * g.Ws * g.Ca, device='cuda') * 2 - 1).to(dt)
    Wd = ((torch.rand(g.Ca * g.R * g.S * g.Cb, device='cuda') * 2 - 1) * 0.1).to(dt)
    Wu = Wd.view(g.Ca, g.R, g.S, g.Cb).permute(3, 1, 2, 0).contiguous().view(-1)
    ba, bb = torch.rand(g.Ca, device='cuda'), torch.rand(g.Cb, device='cuda')
    dW = torch.zeros(Wd.numel(), device='cuda')
    outS, outL = torch.empty_like(S), torch.empty_like(L)
    flops = 2.0 * n * g.Hs * g.Ws * g.Ca * g.R * g.S * g.Cb
    for what, fn in (('down', lambda: ops.conv_down(L, Wd, outS, n, g, ba, 0, None, 0, 1, 0.05)),
                     ('up', lambda: ops.conv_up(S, Wu, outL, n, g, None, 0, L, 1, 1, 0.05)),
                     ('wgrad', lambda: ops.conv_wgrad(S, L, dW, n, g))):
        if filt and len(filt) > 1 and what not in filt:
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f'{name:24s} {what:5s} n={n:7d} {ms * 1e3:9.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s  tensor={ops.lib.srgan_last_path_tensor()}', flush=True)
