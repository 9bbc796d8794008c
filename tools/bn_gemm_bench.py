"""The BatchNorm-fused dense-layer kernels (csrc/bn_gemm.cu) next to the unfused launches they replace, at in-situ shapes
of the crowd trunk (4B = 256 samples), through the C ABI, CUDA-event timed with algorithmic GB/s.
usage: python tools/bn_gemm_bench.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srgan_b200.nets import Geom
from srgan_b200.ops_cuda import CudaOps

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ops = CudaOps()
ops.begin()
dt = torch.bfloat16
NB = 3


def timeit(fn):
    fn(0)
    torch.cuda.synchronize()
    ts = []
    for i in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn((i + 1) % NB); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for rows, C, pitch in ((256 * 196, 1024, 1792), (256 * 784, 320, 512), (256 * 49, 1408, 1920), (256 * 3136, 160, 256), (64 * 196, 1024, 1792)):
    Cp = (C + 63) // 64 * 64
    g = Geom(1, 1, 128, 1, 1, Cp, 1, 1, 1, 0)
    r = lambda *s: (torch.rand(*s, device='cuda') * 2 - 1)
    cat = [r(rows * pitch).to(dt) for _ in range(NB)]
    dcat = [r(rows * pitch).to(dt) for _ in range(NB)]
    n1 = [r(rows * Cp).to(dt) for _ in range(NB)]
    dn1 = [torch.empty(rows * Cp, device='cuda', dtype=dt) for _ in range(NB)]
    db = [r(rows * 128).to(dt) for _ in range(NB)]
    b = [torch.empty(rows * 128, device='cuda', dtype=dt) for _ in range(NB)]
    Wd = (r(128 * Cp) * 0.05).to(dt)
    Wu = Wd.view(128, Cp).t().contiguous().view(-1)
    gamma, beta, mean, var = r(C) + 1.5, r(C) * 0.3, r(C) * 0.2, torch.rand(C, device='cuda') + 0.5
    dg, dbt = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
    dW = torch.zeros(128 * Cp, device='cuda')
    e = 2

    def unfused_bwd(k):
        ops.conv_up(db[k], Wu, dn1[k], rows, g, None, 0, n1[k], 1, 1, 0.0)
        ops.affine_bwd_grad(dn1[k], Cp, cat[k], dcat[k], pitch, 0, rows, C, gamma, mean, var, 1e-5, dg, dbt, True)

    def fused_bwd(k):
        ops.bn_dgrad(db[k], Wu, dcat[k], cat[k], rows, 128, Cp, C, pitch, gamma, beta, mean, var, 1e-5, dg, dbt, None, 0, True)

    def fused_bwd_nograd(k):
        ops.bn_dgrad(db[k], Wu, dcat[k], cat[k], rows, 128, Cp, C, pitch, gamma, beta, mean, var, 1e-5, None, None, None, 0, True)

    cases = [('unfused dgrad + affine_bwd_grad', rows * (128 + 6 * C) * e, unfused_bwd),
             ('bn_dgrad (with dgamma/dbeta)   ', rows * (128 + 3 * C) * e, fused_bwd),
             ('bn_dgrad (data gradient only)  ', rows * (128 + 3 * C) * e, fused_bwd_nograd)]
    if hasattr(ops, 'bn_conv_down'):
        def unfused_fwd(k):
            ops.affine(cat[k], pitch, 0, n1[k], Cp, rows, C, gamma, beta, mean, var, 1e-5, None, 0, 1, 0.0)
            ops.conv_down(n1[k], Wd, b[k], rows, g, None, 0, None, 0, 0, 0.0)

        def fused_fwd(k):
            ops.bn_conv_down(cat[k], Wd, b[k], rows, Cp, 128, C, pitch, gamma, beta, mean, var, 1e-5)

        def fused_fwd_keep(k):
            ops.bn_conv_down(cat[k], Wd, b[k], rows, Cp, 128, C, pitch, gamma, beta, mean, var, 1e-5, n1[k], Cp)

        def unfused_wg(k):
            ops.conv_wgrad(db[k], n1[k], dW, rows, g)

        def fused_wg(k):
            ops.bn_conv_wgrad(db[k], cat[k], dW, rows, 128, Cp, C, pitch, gamma, beta, mean, var, 1e-5)
        cases += [('unfused affine + conv1 forward ', rows * (128 + 3 * C) * e, unfused_fwd),
                  ('bn_conv_down                   ', rows * (128 + C) * e, fused_fwd),
                  ('bn_conv_down (+ n1 stored)     ', rows * (128 + 2 * C) * e, fused_fwd_keep)]
        if hasattr(ops, 'bn_conv_wgrad'):
            cases += [('conv1 weight gradient (n1)     ', rows * (128 + C) * e, unfused_wg),
                      ('bn_conv_wgrad (from cat)       ', rows * (128 + C) * e, fused_wg)]
    for name, nbytes, fn in cases:
        t = timeit(fn)
        print(f'rows={rows:7d} C={C:5d}  {name}: {t * 1e3:7.1f} us  {nbytes / 1e6:7.1f} MB algorithmic  {nbytes / t / 1e6:6.0f} GB/s', flush=True)
    del cat, dcat, n1, dn1, db, b
    torch.cuda.empty_cache()
