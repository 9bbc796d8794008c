"""GPU: the BatchNorm-fused dense-layer kernels (csrc/bn_gemm.cu) against their op-level semantics in tests/torch_ops.py,
shapes of the DenseNet trunk (K = 128 bottleneck channels against concat widths that are / are not multiples of the 128-wide
tile, ragged row counts, concat pitch > C, transition-sized K)."""
import pytest
import torch

from tests.torch_ops import TorchOps

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(scope='module')
def ops():
    from srgan_b200.ops_cuda import CudaOps
    return CudaOps()


def rnd(gen, *shape, dt=torch.float32):
    return (torch.rand(*shape, generator=gen) * 2 - 1).to(dt)


def close(a, b, t, what='', outliers=0.0):
    """max error relative to the largest reference magnitude; `outliers` = fraction of elements allowed beyond it (a ReLU mask
    recomputed in fp32 on both sides may flip for the handful of pre-activations within an ulp of zero)."""
    a, b = a.float().cpu(), b.float().cpu()
    err = (a - b).abs()
    ref = b.abs().max().item() + 1e-12
    bad = (err / ref > t).float().mean().item()
    assert bad <= outliers, f'{what}: {bad:.2e} of the elements beyond {t:.0e} (max rel err {err.max().item() / ref:.3e})'


def bn_params(gen, C):
    gamma = rnd(gen, C) + 1.5
    gamma[::7] *= -0.5                                   # negative scales exist in trained nets
    return gamma, rnd(gen, C) * 0.3, rnd(gen, C) * 0.2, torch.rand(C, generator=gen) + 0.5


@pytest.mark.parametrize('rows,K,C,pitch', [(3001, 128, 96, 256), (12544 + 5, 128, 1056, 1920), (777, 192, 384, 384),
                                            (50, 128, 64, 64), (4 * 3136, 128, 160, 256), (20000, 64, 128, 136)])
@pytest.mark.parametrize('acc,grads,keep', [(True, True, False), (False, False, True), (True, False, False), (False, True, True)])
def test_bn_dgrad(ops, rows, K, C, pitch, acc, grads, keep):
    gen = torch.Generator().manual_seed(rows + C + K)
    ref = TorchOps()
    Cout = (C + 63) // 64 * 64
    dy = rnd(gen, rows * K, dt=BF)
    Wu = torch.zeros(Cout, K)
    Wu[:C] = rnd(gen, C, K) * 0.1
    Wu = Wu.reshape(-1).to(BF)
    x = rnd(gen, rows * pitch, dt=BF)
    gamma, beta, mean, var = bn_params(gen, C)
    dx_ref = rnd(gen, rows * pitch, dt=BF)
    dx = dx_ref.clone().cuda()
    dpitch = Cout
    d_ref = rnd(gen, rows * dpitch, dt=BF) if keep else None
    d = d_ref.clone().cuda() if keep else None
    dg_ref, db_ref = rnd(gen, C), rnd(gen, C)
    dg, db = dg_ref.clone().cuda(), db_ref.clone().cuda()
    cu = lambda t: t.cuda()
    ref.bn_dgrad(dy, Wu, dx_ref, x, rows, K, Cout, C, pitch, gamma, beta, mean, var, 1e-5, dg_ref if grads else None,
                 db_ref if grads else None, d_ref, dpitch, acc)
    ops.bn_dgrad(cu(dy), cu(Wu), dx, cu(x), rows, K, Cout, C, pitch, cu(gamma), cu(beta), cu(mean), cu(var), 1e-5,
                 dg if grads else None, db if grads else None, d, dpitch, acc)
    torch.cuda.synchronize()
    close(dx, dx_ref, 2e-2, 'bn_dgrad dx (incl. the untouched columns >= C)', outliers=1e-5)
    assert torch.equal(dx.cpu().view(rows, pitch)[:, C:], dx_ref.view(rows, pitch)[:, C:])
    close(dg, dg_ref, 1e-2, 'bn_dgrad dgamma')
    close(db, db_ref, 1e-2, 'bn_dgrad dbeta')
    if keep:
        close(d, d_ref, 2e-2, 'bn_dgrad d_out', outliers=1e-5)
        assert torch.equal(d.cpu().view(rows, dpitch)[:, C:], d_ref.view(rows, dpitch)[:, C:])


def test_bn_dgrad_refuses_ineligible(ops):
    t = torch.zeros(64 * 100, dtype=BF, device='cuda')
    f = torch.zeros(100, device='cuda')
    with pytest.raises(RuntimeError):                    # K not a multiple of 64
        ops.bn_dgrad(t, t, t, t, 10, 100, 64, 64, 64, f, f, f, f, 1e-5, None, None, None, 0, False)


@pytest.mark.parametrize('rows,C,pitch,Cout', [(3001, 96, 256, 128), (12544 + 5, 1056, 1920, 128), (777, 384, 384, 192),
                                               (50, 64, 64, 128), (4 * 3136, 160, 256, 128), (40000, 256, 256, 128),
                                               (2 * 148 * 128 + 17, 1888, 1920, 128), (5000, 512, 512, 256)])
@pytest.mark.parametrize('keep,bn2', [(True, False), (False, False), (False, True), (True, True)])
def test_bn_conv_down(ops, rows, C, pitch, Cout, keep, bn2):
    """norm1 -> relu1 -> conv1 forward in one launch (the operand tiles are normalised in shared memory on their way into the
    GEMM) against affine + GEMM semantics; optional store of the normalised operand, optional second BatchNorm + ReLU."""
    gen = torch.Generator().manual_seed(rows + C + Cout)
    ref = TorchOps()
    Kpad = (C + 63) // 64 * 64
    x = rnd(gen, rows * pitch, dt=BF)
    Wd = torch.zeros(Cout, Kpad)
    Wd[:, :C] = rnd(gen, Cout, C) * 0.1
    Wd = Wd.reshape(-1).to(BF)
    gamma, beta, mean, var = bn_params(gen, C)
    p2 = bn_params(gen, Cout) if bn2 else None
    out_ref, out = torch.zeros(rows * Cout, dtype=BF), torch.zeros(rows * Cout, dtype=BF, device='cuda')
    o2_ref = torch.zeros(rows * Cout, dtype=BF) if bn2 else None
    o2 = torch.zeros(rows * Cout, dtype=BF, device='cuda') if bn2 else None
    n1_ref = rnd(gen, rows * Kpad, dt=BF) if keep else None
    n1 = n1_ref.clone().cuda() if keep else None
    cu = lambda t: t.cuda()
    ref.bn_conv_down(x, Wd, out_ref, rows, Kpad, Cout, C, pitch, gamma, beta, mean, var, 1e-5, n1_ref, Kpad, p2, o2_ref)
    ops.bn_conv_down(cu(x), cu(Wd), out, rows, Kpad, Cout, C, pitch, cu(gamma), cu(beta), cu(mean), cu(var), 1e-5, n1, Kpad,
                     tuple(cu(t) for t in p2) if bn2 else None, o2)
    torch.cuda.synchronize()
    close(out, out_ref, 1e-2, 'bn_conv_down out')
    if keep:
        close(n1, n1_ref, 8e-3, 'bn_conv_down n1_out')
    if bn2:
        close(o2, o2_ref, 2e-2, 'bn_conv_down out2', outliers=1e-5)


def test_bn_conv_down_first_row(ops):
    """n1_first_row: the normalised operand is stored for the GEMM rows >= n1_first_row only, the rest of n1_out is untouched."""
    gen = torch.Generator().manual_seed(5)
    ref = TorchOps()
    rows, C, pitch, Cout, Kpad, first = 4 * 196 * 3 + 7, 160, 256, 128, 192, 3 * 196 * 3
    x = rnd(gen, rows * pitch, dt=BF)
    Wd = (rnd(gen, Cout * Kpad) * 0.1).to(BF)
    gamma, beta, mean, var = bn_params(gen, C)
    out_ref, out = torch.zeros(rows * Cout, dtype=BF), torch.zeros(rows * Cout, dtype=BF, device='cuda')
    n1_ref = rnd(gen, rows * Kpad, dt=BF)
    n1 = n1_ref.clone().cuda()
    cu = lambda t: t.cuda()
    ref.bn_conv_down(x, Wd, out_ref, rows, Kpad, Cout, C, pitch, gamma, beta, mean, var, 1e-5, n1_ref, Kpad, n1_first_row=first)
    ops.bn_conv_down(cu(x), cu(Wd), out, rows, Kpad, Cout, C, pitch, cu(gamma), cu(beta), cu(mean), cu(var), 1e-5, n1, Kpad,
                     n1_first_row=first)
    torch.cuda.synchronize()
    close(out, out_ref, 1e-2, 'bn_conv_down out')
    close(n1, n1_ref, 8e-3, 'bn_conv_down n1_out (rows >= first)')
    assert torch.equal(n1.cpu()[:first * Kpad], n1_ref[:first * Kpad])


@pytest.mark.parametrize('rows,C,pitch,Ca', [(3001, 96, 256, 128), (12544 + 5, 1056, 1920, 128), (777, 384, 384, 192),
                                             (50, 64, 64, 128), (4 * 3136, 160, 256, 128), (40000, 256, 256, 128),
                                             (2 * 148 * 128 + 17, 1888, 1920, 128), (5000, 512, 512, 256), (9000, 1792, 1792, 896)])
def test_bn_conv_wgrad(ops, rows, C, pitch, Ca):
    """Weight gradient of norm1 -> relu1 -> conv1 from the raw concat buffer (operand normalised in shared memory on load),
    accumulated into dW; several a tiles (transitions), ragged rows, concat pitch > C, C not a multiple of the tile widths."""
    gen = torch.Generator().manual_seed(rows + C + Ca)
    ref = TorchOps()
    Kpad = (C + 63) // 64 * 64
    x = rnd(gen, rows * pitch, dt=BF)
    dy = rnd(gen, rows * Ca, dt=BF)
    gamma, beta, mean, var = bn_params(gen, C)
    dW_ref = rnd(gen, Ca * Kpad)
    dW = dW_ref.clone().cuda()
    cu = lambda t: t.cuda()
    ref.bn_conv_wgrad(dy, x, dW_ref, rows, Ca, Kpad, C, pitch, gamma, beta, mean, var, 1e-5)
    ops.bn_conv_wgrad(cu(dy), cu(x), dW, rows, Ca, Kpad, C, pitch, cu(gamma), cu(beta), cu(mean), cu(var), 1e-5)
    torch.cuda.synchronize()
    close(dW, dW_ref, 2e-3, 'bn_conv_wgrad dW')
    assert torch.equal(dW.cpu().view(Ca, Kpad)[:, C:], dW_ref.view(Ca, Kpad)[:, C:])


@pytest.mark.parametrize('n,H,cat_ch,c0,C', [(5, 14, 256, 160, 128), (130, 7, 1920, 1888 - 32, 128), (3, 56, 256, 64, 128),
                                             (9, 28, 512, 480, 128), (67, 14, 96, 64, 64)])
@pytest.mark.parametrize('acc,grads,keep', [(False, True, False), (False, False, True), (True, True, True)])
def test_bn_conv_dgrad(ops, n, H, cat_ch, c0, C, acc, grads, keep):
    """norm2 -> relu2 -> conv2 (3x3) backward in one launch: dy is the layer's 32-channel window of the concat delta (TMA zero
    fill up to the 64 channels per tap the weight matrix counts), the result lands in the bottleneck delta; pixel-patch row
    tiles at every trunk resolution, ragged sample counts."""
    from srgan_b200.nets import Geom
    gen = torch.Generator().manual_seed(n + H + cat_ch)
    ref = TorchOps()
    g = Geom(H, H, 64, H, H, (C + 63) // 64 * 64, 3, 3, 1, 1)
    rows, pitch, growth = n * H * H, g.Cb, 32
    dcat = rnd(gen, rows * cat_ch, dt=BF)
    Wu = torch.zeros(g.Cb, 3, 3, 64)
    Wu[:C, :, :, :growth] = rnd(gen, C, 3, 3, growth) * 0.1
    Wu = Wu.reshape(-1).to(BF)
    x = rnd(gen, rows * pitch, dt=BF)
    gamma, beta, mean, var = bn_params(gen, C)
    dx_ref = rnd(gen, rows * pitch, dt=BF)
    dx = dx_ref.clone().cuda()
    d_ref = rnd(gen, rows * pitch, dt=BF) if keep else None
    d = d_ref.clone().cuda() if keep else None
    dg_ref, db_ref = rnd(gen, C), rnd(gen, C)
    dg, db = dg_ref.clone().cuda(), db_ref.clone().cuda()
    cu = lambda t: t.cuda()
    ref.bn_conv_dgrad(dcat[c0:], cat_ch, growth, Wu, dx_ref, x, n, g, C, pitch, gamma, beta, mean, var, 1e-5,
                      dg_ref if grads else None, db_ref if grads else None, d_ref, pitch, acc)
    ops.bn_conv_dgrad(cu(dcat)[c0:], cat_ch, growth, cu(Wu), dx, cu(x), n, g, C, pitch, cu(gamma), cu(beta), cu(mean), cu(var), 1e-5,
                      dg if grads else None, db if grads else None, d, pitch, acc)
    torch.cuda.synchronize()
    close(dx, dx_ref, 2e-2, 'bn_conv_dgrad dx', outliers=1e-5)
    close(dg, dg_ref, 1e-2, 'bn_conv_dgrad dgamma')
    close(db, db_ref, 1e-2, 'bn_conv_dgrad dbeta')
    if keep:
        close(d, d_ref, 2e-2, 'bn_conv_dgrad d_out', outliers=1e-5)
