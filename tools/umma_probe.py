"""Diagnostic (GPU box): tcgen05 path vs forced-SIMT path for every eligible conv-pair shape; prints error structure."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srgan_b200.nets import Geom
from srgan_b200.ops_cuda import CudaOps

ops = CudaOps()
lib = ops.lib
torch.manual_seed(0)

GEOMS = [
    ('k4s2 64->128 @8x8', Geom(4, 4, 128, 8, 8, 64, 4, 4, 2, 1), 5),
    ('k4s2 64->128 @32x32', Geom(16, 16, 128, 32, 32, 64, 4, 4, 2, 1), 3),
    ('k4s2 128->256 @16x16', Geom(8, 8, 256, 16, 16, 128, 4, 4, 2, 1), 3),
    ('k4s2 256->512 @8x8', Geom(4, 4, 512, 8, 8, 256, 4, 4, 2, 1), 9),
    ('linear 256->2048', Geom(1, 1, 2048, 1, 1, 256, 1, 1, 1, 0), 130),
    ('k3s1 64->64 @8x8', Geom(8, 8, 64, 8, 8, 64, 3, 3, 1, 1), 4),
    ('k1 128->128 @4x4', Geom(4, 4, 128, 4, 4, 128, 1, 1, 1, 0), 8),
]
ONLY = sys.argv[1:] or ['down', 'up', 'wgrad']


def report(name, a, b, cols):
    a, b = a.float(), b.float()
    d = (a - b).abs()
    ref = b.abs().max().item() + 1e-9
    bad = d > 2e-2 * ref
    msg = f'{name}: max rel err {d.max().item() / ref:.3e}, bad {bad.float().mean().item() * 100:.2f}%'
    if bad.any():
        idx = bad.nonzero().flatten()[:6].tolist()
        msg += ' first bad ' + str([(i // cols, i % cols, round(a[i].item(), 4), round(b[i].item(), 4)) for i in idx])
        rows = bad.view(-1, cols)
        msg += f' | bad rows {rows.any(1).float().mean().item() * 100:.1f}% bad cols {rows.any(0).float().mean().item() * 100:.1f}%'
        colbad = rows.any(0).nonzero().flatten()[:16].tolist()
        rowbad = rows.any(1).nonzero().flatten()[:16].tolist()
        msg += f' cols{colbad} rows{rowbad}'
    print(msg, flush=True)
    return not bad.any()


ok_all = True
for name, g, n in GEOMS:
    dt = torch.bfloat16
    L = (torch.rand(n * g.Hl * g.Wl * g.Cb, device='cuda') * 2 - 1).to(dt)
    S = (torch.rand(n * g.Hs * g.Ws * g.Ca, device='cuda') * 2 - 1).to(dt)
    Wd = ((torch.rand(g.Ca * g.R * g.S * g.Cb, device='cuda') * 2 - 1) * 0.2).to(dt)
    Wu = Wd.view(g.Ca, g.R, g.S, g.Cb).permute(3, 1, 2, 0).contiguous().view(-1)
    bias_a, bias_b = torch.rand(g.Ca, device='cuda'), torch.rand(g.Cb, device='cuda')
    for what in ONLY:
        res = []
        for force in (1, 0):
            lib.srgan_set_force_simt(force)
            try:
                if what == 'down':
                    out = torch.zeros_like(S)
                    ops.conv_down(L, Wd, out, n, g, bias_a, 0, None, 0, 1, 0.05)
                elif what == 'up':
                    out = torch.zeros_like(L)
                    ops.conv_up(S, Wu, out, n, g, bias_b, 0, None, 0, 1, 0.05)
                else:
                    out = torch.zeros(Wd.numel(), device='cuda')
                    ops.conv_wgrad(S, L, out, n, g)
                torch.cuda.synchronize()
            except Exception as e:
                print(name, what, 'force', force, 'EXC', e, flush=True)
                out = None
            res.append((out, lib.srgan_last_path_tensor()))
        lib.srgan_set_force_simt(0)
        (ref, _), (got, tensor) = res
        cols = {'down': g.Ca, 'up': g.Cb, 'wgrad': g.R * g.S * g.Cb}[what]
        if got is None or ref is None:
            ok_all = False
            continue
        ok = report(f'{name:24s} {what:5s} tensor={tensor}', got.view(-1), ref.view(-1), cols)
        ok_all &= ok
print('ALL OK' if ok_all else 'MISMATCHES', flush=True)
