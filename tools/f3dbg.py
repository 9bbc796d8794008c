import os, sys
sys.path.insert(0, '/root/repo')
import torch
from tests.torch_ops import TorchOps
from srgan_b200.nets import Geom
from srgan_b200.ops_cuda import CudaOps
ops = CudaOps(); ref = TorchOps(); BF = torch.bfloat16
hw, n, cin, valid = int(sys.argv[1]), int(sys.argv[2]), 64, 32
g = Geom(hw, hw, 64, hw, hw, cin, 3, 3, 1, 1)
pix = n * hw * hw
gen = torch.Generator().manual_seed(0)
L = (torch.rand(pix * cin, generator=gen) * 2 - 1).to(BF)
for tap in list(range(9)) + [-1]:
    W4 = torch.zeros(g.Ca, 9, cin)
    if tap >= 0:
        W4[:valid, tap] = torch.rand(valid, cin, generator=gen) * 0.2 - 0.1
    else:
        W4[:valid] = torch.rand(valid, 9, cin, generator=gen) * 0.2 - 0.1
    Wd = W4.reshape(-1).to(BF)
    out_ref = torch.zeros(pix * valid, dtype=BF); out = torch.zeros(pix * valid, dtype=BF, device='cuda')
    vw = (valid, valid, 0, 0)
    ref.conv_down(L, Wd, out_ref, n, g, None, 0, None, 0, 0, 0.0, views=vw)
    ops.conv_down(L.cuda(), Wd.cuda(), out, n, g, None, 0, None, 0, 0, 0.0, views=vw)
    torch.cuda.synchronize()
    a, b = out.float().cpu().view(n, hw, hw, valid), out_ref.float().view(n, hw, hw, valid)
    e = (a - b).abs()
    bad = (e.amax(dim=(0, 3)) > 0.02)
    print('tap', tap, 'max err', e.max().item(), 'ref max', b.abs().max().item(), 'bad pixels', int(bad.sum()), 'of', hw * hw,
          'first bad', bad.nonzero()[:4].tolist())
