"""GPU: every C-ABI kernel against the op-level semantics in tests/torch_ops.py (plain PyTorch fp32 on CPU), fp32 and
bf16 activations, shapes from the coefficient / DCGAN configs incl. ragged and non-vectorisable ones."""
import pytest
import torch

from srgan_b200.nets import Geom
from tests.torch_ops import TorchOps

pytestmark = pytest.mark.gpu

DT = [torch.float32, torch.bfloat16]


def tol(dt):
    return 2e-5 if dt == torch.float32 else 2e-2


@pytest.fixture(scope='module')
def ops():
    from srgan_b200.ops_cuda import CudaOps
    return CudaOps()


def rnd(gen, *shape, dt=torch.float32):
    return (torch.rand(*shape, generator=gen) * 2 - 1).to(dt)


def close(a, b, t, what=''):
    a, b = a.float().cpu(), b.float().cpu()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item() + 1e-12
    assert err / ref < t, f'{what}: rel err {err / ref:.3e} (abs {err:.3e}, ref {ref:.3e})'


GEOMS = [
    Geom(1, 1, 10, 1, 1, 50, 1, 1, 1, 0),          # coefficient linear 50 -> 10
    Geom(1, 1, 50, 1, 1, 10, 1, 1, 1, 0),          # coefficient generator 10 -> 50
    Geom(16, 16, 8, 32, 32, 3, 4, 4, 2, 1),        # DCGAN layer1 (3 channels)
    Geom(8, 8, 16, 16, 16, 8, 4, 4, 2, 1),         # small DCGAN layer
    Geom(4, 4, 128, 8, 8, 64, 4, 4, 2, 1),         # tensor-core eligible channels
    Geom(1, 1, 256, 1, 1, 64, 1, 1, 1, 0),         # fc as linear
    Geom(5, 7, 12, 11, 15, 5, 3, 3, 2, 0),         # ragged, k3 s2 p0
    Geom(6, 6, 8, 6, 6, 4, 3, 3, 1, 1),            # stride 1 same-pad
    Geom(16, 16, 128, 32, 32, 64, 4, 4, 2, 1),     # tcgen05: DCGAN layer 2 channels
    Geom(4, 4, 512, 8, 8, 256, 4, 4, 2, 1),        # tcgen05: DCGAN layer 4 channels (wgrad BN=256)
    Geom(1, 1, 2048, 1, 1, 256, 1, 1, 1, 0),       # tcgen05: fc as linear
    Geom(8, 8, 64, 8, 8, 64, 3, 3, 1, 1),          # tcgen05: k3 s1; wgrad with Ca = 64 (half tile zero-filled)
    Geom(1, 1, 64, 1, 1, 64, 1, 1, 1, 0),          # tcgen05: the [pixels x 64] GEMM of the thin-layer lowering
    # DenseNet trunk 1x1 GEMMs (Ca = 128 bottleneck channels, Cb = padded concat width): wgrad b-tile groups / partial tiles
    Geom(1, 1, 128, 1, 1, 192, 1, 1, 1, 0),        # Cb = 3 x 64
    Geom(1, 1, 128, 1, 1, 1088, 1, 1, 1, 0),       # Cb = 17 x 64: 256-wide tiles, the last one mostly TMA zero fill
    Geom(1, 1, 128, 1, 1, 320, 1, 1, 1, 0),        # Cb = 5 x 64
]
TENSOR_ELIGIBLE = {4, 8, 9, 10, 11, 12, 13, 14, 15}
WGRAD_TENSOR_ELIGIBLE = {4, 8, 9, 10, 11, 12, 13, 14, 15}
TRUNK_GEMM_ROWS = {13: 64 * 196 + 5, 14: 3001, 15: 20 * 3136}       # small / large row counts flip the (BN, NB) choice


@pytest.mark.parametrize('dt', DT)
@pytest.mark.parametrize('gi', range(len(GEOMS)))
def test_conv_down_up_wgrad(ops, dt, gi):
    g = GEOMS[gi]
    gen = torch.Generator().manual_seed(gi)
    n = (5000 if g.Ca == 64 else 130) if g.Hl == 1 else 5
    n = TRUNK_GEMM_ROWS.get(gi, n)
    ref = TorchOps()
    L = rnd(gen, n * g.Hl * g.Wl * g.Cb, dt=dt)
    S = rnd(gen, n * g.Hs * g.Ws * g.Ca, dt=dt)
    Wd = (rnd(gen, g.Ca * g.R * g.S * g.Cb) * 0.3).to(dt)
    Wu = Wd.view(g.Ca, g.R, g.S, g.Cb).permute(3, 1, 2, 0).contiguous().view(-1)
    bias_a, bias_b = rnd(gen, g.Ca), rnd(gen, g.Cb)
    for epi, act, slope in ((0, 1, 0.05), (0, 2, 0.0), (0, 0, 0.0), (1, 1, 0.01), (1, 1, 0.0), (1, 2, 0.0), (1, 0, 0.0)):
        # down            (epi 1, act 1, slope 0 = the ReLU mask: packed fast path of the tcgen05 epilogue)
        href = rnd(gen, S.numel(), dt=dt)
        out_ref = torch.empty_like(S)
        ref.conv_down(L, Wd, out_ref, n, g, bias_a if epi == 0 else None, 0, href if epi == 1 else None, epi, act, slope)
        out = torch.empty_like(S, device='cuda')
        ops.conv_down(L.cuda(), Wd.cuda(), out, n, g, bias_a.cuda() if epi == 0 else None, 0,
                      href.cuda() if epi == 1 else None, epi, act, slope)
        close(out, out_ref, tol(dt), f'down epi{epi} act{act}')
        if dt == torch.bfloat16 and gi in TENSOR_ELIGIBLE:
            assert ops.lib.srgan_last_path_tensor() == 1, 'expected the tcgen05 path'
        # up
        href = rnd(gen, L.numel(), dt=dt)
        out_ref = torch.empty_like(L)
        ref.conv_up(S, Wu, out_ref, n, g, bias_b if epi == 0 else None, 0, href if epi == 1 else None, epi, act, slope)
        out = torch.empty_like(L, device='cuda')
        ops.conv_up(S.cuda(), Wu.cuda(), out, n, g, bias_b.cuda() if epi == 0 else None, 0,
                    href.cuda() if epi == 1 else None, epi, act, slope)
        close(out, out_ref, tol(dt), f'up epi{epi} act{act}')
    dW_ref = rnd(gen, Wd.numel())
    dW = dW_ref.clone().cuda()
    ref.conv_wgrad(S, L, dW_ref, n, g)
    ops.conv_wgrad(S.cuda(), L.cuda(), dW, n, g)
    close(dW, dW_ref, tol(dt) * 2, 'wgrad')
    if dt == torch.bfloat16 and gi in WGRAD_TENSOR_ELIGIBLE:
        assert ops.lib.srgan_last_path_tensor() == 1, 'expected the tcgen05 wgrad path'


@pytest.mark.parametrize('valid,c0,pitch', [(32, 96, 160), (8, 24, 40), (64, 64, 128)])
@pytest.mark.parametrize('hw', [14, 7, 8])
def test_conv_channel_windows(ops, valid, c0, pitch, hw):
    """srgan_views: a DenseNet dense layer's 3x3 convolution (crowd/models.py:345,353) writes its growth_rate channels in
    place into their window [c0, c0+valid) of the concat buffer (pitch channels wide) and its data / weight gradients read
    that window of the concat delta; bf16 tcgen05 path against the dense op-level semantics.  The rest of the concat buffer
    must stay untouched."""
    dt = torch.bfloat16
    g = Geom(hw, hw, 64, hw, hw, 128, 3, 3, 1, 1)          # Ca padded to 64, `valid` of them real
    n = 37
    gen = torch.Generator().manual_seed(valid + hw)
    ref = TorchOps()
    pix = n * hw * hw
    L = rnd(gen, pix * g.Cb, dt=dt)
    Wd4 = (rnd(gen, g.Ca, g.R * g.S * g.Cb) * 0.1)
    Wd4[valid:] = 0                                          # pad rows of the kernel-layout weights are zero
    Wd = Wd4.reshape(-1).to(dt)
    Wu = Wd.view(g.Ca, g.R, g.S, g.Cb).permute(3, 1, 2, 0).contiguous().view(-1)
    vw = (pitch, valid, 0, 0)
    cat0 = rnd(gen, pix * pitch, dt=dt)
    # forward / tangent: window write
    for epi in (0, 1):
        cat_ref, cat = cat0.clone(), cat0.clone().cuda()
        ref.conv_down(L, Wd, cat_ref[c0:], n, g, None, 0, None, epi, 0, 0.0, views=vw)
        ops.conv_down(L.cuda(), Wd.cuda(), cat[c0:], n, g, None, 0, None, epi, 0, 0.0, views=vw)
        assert ops.lib.srgan_last_path_tensor() == 1
        close(cat.view(pix, pitch)[:, c0:c0 + valid], cat_ref.view(pix, pitch)[:, c0:c0 + valid], tol(dt), 'window write')
        mask = torch.ones(pix, pitch, dtype=torch.bool)
        mask[:, c0:c0 + valid] = False
        assert torch.equal(cat.cpu().view(pix, pitch)[mask], cat0.view(pix, pitch)[mask]), 'wrote outside the window'
    # data gradient: window read (ReLU mask of the dense 128-channel operand)
    dcat = rnd(gen, pix * pitch, dt=dt)
    href = rnd(gen, L.numel(), dt=dt)
    dx_ref, dx = torch.empty_like(L), torch.empty_like(L, device='cuda')
    ref.conv_up(dcat[c0:], Wu, dx_ref, n, g, None, 0, href, 1, 1, 0.0, views=vw)
    ops.conv_up(dcat.cuda()[c0:], Wu.cuda(), dx, n, g, None, 0, href.cuda(), 1, 1, 0.0, views=vw)
    assert ops.lib.srgan_last_path_tensor() == 1
    close(dx, dx_ref, tol(dt), 'window dgrad')
    # weight gradient: window read on the small side
    dW_ref = rnd(gen, Wd.numel())
    dW = dW_ref.clone().cuda()
    ref.conv_wgrad(dcat[c0:], L, dW_ref, n, g, views=vw)
    ops.conv_wgrad(dcat.cuda()[c0:], L.cuda(), dW, n, g, views=vw)
    assert ops.lib.srgan_last_path_tensor() == 1
    close(dW, dW_ref, tol(dt) * 2, 'window wgrad')


def test_conv_channel_windows_need_the_tensor_path(ops):
    """No second implementation behind srgan_views: an fp32 call (or a shape that is not tcgen05-eligible) raises."""
    g = Geom(8, 8, 64, 8, 8, 128, 3, 3, 1, 1)
    n = 2
    L = torch.zeros(n * 64 * 128, device='cuda')
    W = torch.zeros(64 * 9 * 128, device='cuda')
    out = torch.zeros(n * 64 * 160, device='cuda')
    with pytest.raises(RuntimeError):
        ops.conv_down(L, W, out, n, g, None, 0, None, 0, 0, 0.0, views=(160, 32, 0, 0))


@pytest.mark.parametrize('dt', DT)
@pytest.mark.parametrize('ca,cb', [(8, 1), (16, 8), (32, 16)])
def test_patch_convs(ops, dt, ca, cb):
    """The crowd MapModule convs (kernel = stride = 2, crowd/models.py:770-776) and their data gradients at a pixel count that
    takes the one-thread-per-pixel patch kernels (simt_conv.cu), every epilogue."""
    g = Geom(24, 20, ca, 48, 40, cb, 2, 2, 2, 0)
    gen = torch.Generator().manual_seed(ca)
    n = 9
    ref = TorchOps()
    L = rnd(gen, n * g.Hl * g.Wl * g.Cb, dt=dt)
    S = rnd(gen, n * g.Hs * g.Ws * g.Ca, dt=dt)
    Wd = (rnd(gen, g.Ca * g.R * g.S * g.Cb) * 0.3).to(dt)
    Wu = Wd.view(g.Ca, g.R, g.S, g.Cb).permute(3, 1, 2, 0).contiguous().view(-1)
    bias_a, bias_b = rnd(gen, g.Ca), rnd(gen, g.Cb)
    for epi, act, slope in ((0, 1, 0.01), (0, 0, 0.0), (0, 2, 0.0), (1, 1, 0.01), (1, 2, 0.0), (1, 0, 0.0)):
        href = rnd(gen, S.numel(), dt=dt)
        out_ref = torch.empty_like(S)
        ref.conv_down(L, Wd, out_ref, n, g, bias_a if epi == 0 else None, 0, href if epi == 1 else None, epi, act, slope)
        out = torch.empty_like(S, device='cuda')
        ops.conv_down(L.cuda(), Wd.cuda(), out, n, g, bias_a.cuda() if epi == 0 else None, 0,
                      href.cuda() if epi == 1 else None, epi, act, slope)
        close(out, out_ref, tol(dt), f'patch down epi{epi} act{act}')
        href = rnd(gen, L.numel(), dt=dt)
        out_ref = torch.empty_like(L)
        ref.conv_up(S, Wu, out_ref, n, g, bias_b if epi == 0 else None, 0, href if epi == 1 else None, epi, act, slope)
        out = torch.empty_like(L, device='cuda')
        ops.conv_up(S.cuda(), Wu.cuda(), out, n, g, bias_b.cuda() if epi == 0 else None, 0,
                    href.cuda() if epi == 1 else None, epi, act, slope)
        close(out, out_ref, tol(dt), f'patch up epi{epi} act{act}')
    dW_ref = rnd(gen, Wd.numel())
    dW = dW_ref.clone().cuda()
    ref.conv_wgrad(S, L, dW_ref, n, g)
    ops.conv_wgrad(S.cuda(), L.cuda(), dW, n, g)
    close(dW, dW_ref, tol(dt) * 2, 'patch wgrad')


@pytest.mark.parametrize('dt', DT)
def test_bias_mod_epilogue(ops, dt):
    g = Geom(1, 1, 4 * 4 * 8, 1, 1, 16, 1, 1, 1, 0)      # fc_up style: bias per channel, broadcast over 16 taps
    gen = torch.Generator().manual_seed(3)
    n = 6
    L, Wd, bias = rnd(gen, n * 16, dt=dt), rnd(gen, g.Ca * 16, dt=dt), rnd(gen, 8)
    ref_out = torch.empty(n * g.Ca, dtype=dt)
    TorchOps().conv_down(L, Wd, ref_out, n, g, bias, 8, None, 0, 0, 0.0)
    out = torch.empty(n * g.Ca, dtype=dt, device='cuda')
    ops.conv_down(L.cuda(), Wd.cuda(), out, n, g, bias.cuda(), 8, None, 0, 0, 0.0)
    close(out, ref_out, tol(dt), 'bias_mod')


@pytest.mark.parametrize('dt', DT)
@pytest.mark.parametrize('rows,cols,mod', [(100, 32768, 0), (64 * 64 * 7, 64, 0), (5000, 10, 0), (37, 1, 0),
                                           (6, 128, 8), (1000, 50, 0), (20001, 3, 0), (8192, 64, 0), (301, 20, 0),
                                           (5000, 1, 0), (9000, 8, 4), (4097, 6, 3)])
def test_colsum_rowdot_seed(ops, dt, rows, cols, mod):
    gen = torch.Generator().manual_seed(rows + cols)
    ref = TorchOps()
    X = rnd(gen, rows * cols, dt=dt)
    rs = rnd(gen, rows)
    for rowscale in (None, rs):
        o_ref = rnd(gen, mod or cols)
        o = o_ref.clone().cuda()
        ref.colsum(X, rows, cols, o_ref, mod, rowscale)
        ops.colsum(X.cuda(), rows, cols, o, mod, rowscale.cuda() if rowscale is not None else None)
        close(o, o_ref, 1e-4 if dt == torch.float32 else 1e-2, 'colsum')
    w, bias = rnd(gen, cols), rnd(gen, 2)
    o_ref = torch.empty(rows)
    o = torch.empty(rows, device='cuda')
    ref.rowdot(X, rows, cols, w, bias, 1, o_ref)
    ops.rowdot(X.cuda(), rows, cols, w.cuda(), bias.cuda(), 1, o)
    close(o, o_ref, 1e-4 if dt == torch.float32 else 1e-2, 'rowdot')
    gvec, wrow = rnd(gen, cols), rnd(gen, cols)
    for gv, rsc in ((gvec, None), (None, rs), (gvec, rs)):
        for act, slope in ((1, 0.05), (2, 0.0), (0, 0.0)):
            o_ref = torch.empty(rows * cols, dtype=dt)
            o = torch.empty(rows * cols, dtype=dt, device='cuda')
            ref.seed_rows(o_ref, rows, cols, gv, rsc, wrow if rsc is not None else None, X, act, slope)
            ops.seed_rows(o, rows, cols, gv.cuda() if gv is not None else None, rsc.cuda() if rsc is not None else None,
                          wrow.cuda() if rsc is not None else None, X.cuda(), act, slope)
            close(o, o_ref, tol(dt), 'seed_rows')


@pytest.mark.parametrize('dt', DT)
@pytest.mark.parametrize('n,c,h,w', [(4, 3, 32, 32), (7, 50, 1, 1), (3, 8, 5, 9), (2, 1, 6, 6)])
def test_layout_and_interpolate(ops, dt, n, c, h, w):
    gen = torch.Generator().manual_seed(n * c)
    ref = TorchOps()
    src = rnd(gen, n, c, h, w)
    d_ref = torch.empty(n * c * h * w, dtype=dt)
    d = torch.empty(n * c * h * w, dtype=dt, device='cuda')
    ref.nchw_to_nhwc(src, d_ref, n, c, h, w)
    ops.nchw_to_nhwc(src.cuda(), d, n, c, h, w)
    close(d, d_ref, 1e-7, 'nchw_to_nhwc')
    back = torch.empty(n, c, h, w, device='cuda')
    ops.nhwc_to_nchw(d, back, n, c, h, w)
    close(back, src.to(dt), 1e-7, 'nhwc_to_nchw')
    E = c * h * w
    u, f, alpha = rnd(gen, n * E, dt=dt), rnd(gen, n * E, dt=dt), torch.rand(n, generator=gen)
    o_ref = torch.empty(n * E, dtype=dt)
    o = torch.empty(n * E, dtype=dt, device='cuda')
    ref.interpolate(u, f, alpha, o_ref, n, E)
    ops.interpolate(u.cuda(), f.cuda(), alpha.cuda(), o, n, E)
    close(o, o_ref, tol(dt), 'interpolate')


def test_scalar_losses(ops):
    gen = torch.Generator().manual_seed(5)
    ref = TorchOps()
    for n in (4, 100, 5000):
        pred, y = rnd(gen, n) * 3, rnd(gen, n) * 3
        for order in (1, 2, 3):
            l_ref, d_ref = torch.zeros(1), torch.empty(n)
            l, d = torch.zeros(1, device='cuda'), torch.empty(n, device='cuda')
            ref.labeled_loss(pred, y, n, order, 0.37 / n, l_ref, d_ref)
            ops.labeled_loss(pred.cuda(), y.cuda(), n, order, 0.37 / n, l, d)
            close(l, l_ref, 1e-5, 'labeled loss'); close(d, d_ref, 1e-5, 'dpred')
        for target in (0.0, 1.0):
            l_ref, d_ref = torch.zeros(1), torch.empty(n)
            l, d = torch.zeros(1, device='cuda'), torch.empty(n, device='cuda')
            ref.bce_logits(pred * 4, n, target, 10.0 / n, l_ref, d_ref)
            ops.bce_logits((pred * 4).cuda(), n, target, 10.0 / n, l, d)
            close(l, l_ref, 1e-5, 'bce'); close(d, d_ref, 1e-5, 'dscore')
    for F in (10, 80, 32768):
        sb, so = rnd(gen, F) * 50, rnd(gen, F) * 50
        for kind in range(6):
            for acc in (False, True):
                l_ref, gb_ref, go_ref = torch.zeros(1), rnd(gen, F), torch.empty(F)
                l, gb, go = torch.zeros(1, device='cuda'), gb_ref.clone().cuda(), torch.empty(F, device='cuda')
                ref.distance(sb, so, F, 1 / 100, kind, 7.0, l_ref, gb_ref, go_ref, acc)
                ops.distance(sb.cuda(), so.cuda(), F, 1 / 100, kind, 7.0, l, gb, go, acc)
                close(l, l_ref, 2e-5, f'distance {kind}'); close(gb, gb_ref, 2e-5, 'gbase'); close(go, go_ref, 2e-5, 'gother')


@pytest.mark.parametrize('dt', DT)
@pytest.mark.parametrize('rows,cols', [(100, 32768), (64, 10), (4, 256), (3, 150528)])
def test_gradient_penalty_kernels(ops, dt, rows, cols):
    gen = torch.Generator().manual_seed(cols)
    ref = TorchOps()
    h, u = rnd(gen, rows * cols, dt=dt), rnd(gen, rows * cols, dt=dt)
    s_ref, gm_ref = torch.empty(rows), torch.empty(rows * cols, dtype=dt)
    s, gm = torch.empty(rows, device='cuda'), torch.empty(rows * cols, dtype=dt, device='cuda')
    ref.feature_norm_seed(h, rows, cols, s_ref, gm_ref, 1, 0.05)
    ops.feature_norm_seed(h.cuda(), rows, cols, s, gm, 1, 0.05)
    close(s, s_ref, 1e-4, 's'); close(gm, gm_ref, tol(dt), 'gamma')
    o_ref = torch.empty(rows * cols, dtype=dt)
    o = torch.empty(rows * cols, dtype=dt, device='cuda')
    ref.gp_feature_seed(u, h, s_ref, o_ref, rows, cols, 1, 0.05)
    ops.gp_feature_seed(u.cuda(), h.cuda(), s_ref.cuda(), o, rows, cols, 1, 0.05)
    close(o, o_ref, tol(dt) * 2, 'gp_feature_seed')
    g0 = (h.float() * (3.0 / cols ** 0.5)).to(dt)          # norms around 1.7: some rows above, none at 0
    gn_ref, p_ref, m_ref, u0_ref = torch.empty(rows), torch.zeros(1), torch.zeros(1), torch.empty(rows * cols, dtype=dt)
    gn, p, m, u0 = (torch.empty(rows, device='cuda'), torch.zeros(1, device='cuda'), torch.zeros(1, device='cuda'),
                    torch.empty(rows * cols, dtype=dt, device='cuda'))
    ref.gradnorm_penalty(g0, rows, cols, 100.0 / rows, 1.0 / rows, gn_ref, p_ref, m_ref, u0_ref)
    ops.gradnorm_penalty(g0.cuda(), rows, cols, 100.0 / rows, 1.0 / rows, gn, p, m, u0)
    close(gn, gn_ref, 1e-4, 'gnorm'); close(p, p_ref, 1e-3, 'penalty'); close(m, m_ref, 1e-4, 'gnorm mean')
    close(u0, u0_ref, tol(dt) * 2, 'u0')


@pytest.mark.parametrize('dims,kind', [((6, 5, 4, 4), 'conv'), ((72, 40, 4, 4), 'conv'), ((33, 50, 3, 3), 'conv'),
                                       ((128, 96, 1, 1), 'conv'), ((20, 9, 7, 7), 'conv'), ((24, 40, 8, 8), 'fc_up'),
                                       ((64, 3, 4, 4), 'thin'), ((64, 64, 4, 4), 'conv'), ((96, 96, 3, 3), 'conv'),
                                       ((128, 544, 1, 1), 'conv'), ((64, 32, 8, 8), 'fc_up')])
def test_adam_and_repack(ops, dims, kind):
    """Fused Adam + kernel-layout rewrite over the stride patterns engine.py produces (conv, fc_up, thin, the crowd trunk's
    padded 1x1 layouts), small and large tensors."""
    gen = torch.Generator().manual_seed(9)
    ref = TorchOps()
    a, b, r, s = dims                        # master [a][b][r][s]
    if kind == 'conv':
        wd_s, wu_s = (r * s * b, 1, s * b, b), (1, r * s * a, s * a, a)
    elif kind == 'fc_up':                    # master[zb][c][k][k] as a Linear (engine._strides_for)
        zb, c, k = a, b, r
        wd_s, wu_s = (1, zb, k * c * zb, c * zb), (k * k * c, 1, k * c, c)
    else:                                    # thin-layer lowering (engine._thin_strides), KPAD = 64
        wd_s, wu_s = (64, 1, s * b, b), (1, a, s * b * a, b * a)
    n = a * b * r * s if kind != 'thin' else a * 64
    for od in (torch.float32, torch.bfloat16):
        p_ref = rnd(gen, *dims)
        grad, m_ref, v_ref = rnd(gen, n), rnd(gen, n) * 0.1, rnd(gen, n).abs() * 0.01
        o1_ref, o2_ref = torch.zeros(n, dtype=od), torch.zeros(n, dtype=od)
        p, m, v = p_ref.clone().cuda(), m_ref.clone().cuda(), v_ref.clone().cuda()
        o1, o2 = torch.zeros(n, dtype=od, device='cuda'), torch.zeros(n, dtype=od, device='cuda')
        ref.repack(p_ref, dims, o1_ref, wd_s, o2_ref, wu_s)
        ops.repack(p, dims, o1, wd_s, o2, wu_s)
        close(o1, o1_ref, 1e-7, 'repack1'); close(o2, o2_ref, 1e-7, 'repack2')
        st_ref, st = torch.zeros(3), torch.zeros(3, device='cuda')
        for step in (1, 2):
            ref.adam_prepare(st_ref, 1e-3, 0.9, 0.999)
            ops.adam_prepare(st, 1e-3, 0.9, 0.999)
            close(st, st_ref, 1e-6, 'adam state')
            assert float(st_ref[0]) == step
            ref.adam(p_ref, grad, m_ref, v_ref, dims, wd_s, o1_ref, wd_s, o2_ref, wu_s, st_ref, 0.9, 0.999, 1e-8, 1e-2)
            ops.adam(p, grad.cuda(), m, v, dims, wd_s, o1, wd_s, o2, wu_s, st, 0.9, 0.999, 1e-8, 1e-2)
            close(p, p_ref, 5e-6, 'adam p'); close(m, m_ref, 5e-6, 'adam m'); close(v, v_ref, 5e-6, 'adam v')
            close(o1, o1_ref, 5e-6 if od == torch.float32 else 1e-2, 'adam out1')
            close(o2, o2_ref, 5e-6 if od == torch.float32 else 1e-2, 'adam out2')


@pytest.mark.parametrize('od', [torch.float32, torch.bfloat16])
def test_adam_layout_multi_matches_per_tensor_adam(ops, od):
    """srgan_adam_layout_multi (one table-driven launch for every tensor with kernel-layout copies: the crowd discriminator's
    200 convolution weights) against srgan_adam tensor by tensor: conv / padded-trunk / fc_up stride patterns, a tensor
    with only one copy, an fp32 prediction-head copy, tensors that are not a multiple of the 2048-element block."""
    gen = torch.Generator().manual_seed(19)
    cases = []
    for dims, kind in (((32, 128, 3, 3), 'conv'), ((128, 96, 1, 1), 'pad'), ((16, 8, 4, 4), 'fc_up'), ((20, 32, 7, 7), 'conv'),
                       ((5, 3, 1, 1), 'conv'), ((64, 64, 3, 3), 'one'), ((2, 80, 1, 1), 'head'), ((256, 512, 4, 4), 'conv')):
        a, b, r, s_ = dims
        if kind == 'fc_up':
            zb, c, k = a, b, r
            wd_s, wu_s, n_out = (1, zb, k * c * zb, c * zb), (k * k * c, 1, k * c, c), a * b * r * s_
        elif kind == 'pad':                  # channel-padded 1x1 layout: 128 x 128 copies of a 128 x 96 master
            wd_s, wu_s, n_out = (128, 1, 0, 0), (1, 128, 0, 0), 128 * 128
        else:
            wd_s, wu_s, n_out = (r * s_ * b, 1, s_ * b, b), (1, r * s_ * a, s_ * a, a), a * b * r * s_
        cases.append((dims, kind, wd_s, wu_s, n_out))
    total = sum(d[0] * d[1] * d[2] * d[3] for d, *_ in cases) + 64
    gtot = sum(nn for *_, nn in cases) + 64
    grad = rnd(gen, gtot).cuda()
    state = torch.zeros(3, device='cuda')
    runs = []
    for multi in (False, True):
        m, v = (rnd(torch.Generator().manual_seed(3), total) * 0.1).cuda(), (rnd(torch.Generator().manual_seed(4), total).abs() * 0.01).cuda()
        g2 = torch.Generator().manual_seed(5)
        entries, outs, params = [], [], []
        po = go = 0
        for dims, kind, wd_s, wu_s, n_out in cases:
            p = rnd(g2, *dims).cuda()
            n = p.numel()
            odt = torch.float32 if kind == 'head' else od
            o1 = torch.zeros(n_out, dtype=odt, device='cuda')
            o2 = None if kind in ('one', 'head') else torch.zeros(n_out, dtype=odt, device='cuda')
            entries.append((p, go, po, dims, wd_s, o1, wd_s, o2, wu_s if o2 is not None else None))
            params.append(p); outs.append((o1, o2))
            po += (n + 3) // 4 * 4
            go += (n_out + 3) // 4 * 4
        state.zero_()
        for step in range(2):
            ops.adam_prepare(state, 1e-3, 0.9, 0.999)
            if multi:
                ops.adam_layout_multi(entries, grad, m, v, state, 0.9, 0.999, 1e-8, 1e-2, od)
            else:
                for p, go_, po_, dims, gs, o1, s1, o2, s2 in entries:
                    n = p.numel()
                    ops.adam(p, grad[go_:], m[po_:po_ + n], v[po_:po_ + n], dims, gs, o1, s1, o2, s2, state, 0.9, 0.999, 1e-8, 1e-2)
        runs.append((params, outs, m, v))
    (pa, oa, ma, va), (pb, ob, mb, vb) = runs
    assert torch.equal(ma, mb) and torch.equal(va, vb)
    for x, y in zip(pa, pb):
        assert torch.equal(x, y)
    for (a1, a2), (b1, b2) in zip(oa, ob):
        assert torch.equal(a1, b1) and (a2 is None or torch.equal(a2, b2))


@pytest.mark.parametrize('dt', DT)
@pytest.mark.parametrize('g', [Geom(16, 16, 64, 32, 32, 3, 4, 4, 2, 1), Geom(5, 7, 64, 11, 15, 4, 3, 3, 2, 0),
                               Geom(8, 8, 128, 8, 8, 1, 3, 3, 1, 1)])
def test_im2col_col2im(ops, dt, g):
    gen = torch.Generator().manual_seed(g.Hl)
    ref = TorchOps()
    n, kpad = 3, 64
    L = rnd(gen, n * g.Hl * g.Wl * g.Cb, dt=dt)
    c_ref = torch.empty(n * g.Hs * g.Ws * kpad, dtype=dt)
    c = torch.full((n * g.Hs * g.Ws * kpad,), 7.0, dtype=dt, device='cuda')
    ref.im2col(L, c_ref, n, g, kpad)
    ops.im2col(L.cuda(), c, n, g, kpad)
    close(c, c_ref, 1e-7, 'im2col')
    col = rnd(gen, n * g.Hs * g.Ws * kpad, dt=dt)
    bias, href = rnd(gen, g.Cb), rnd(gen, L.numel(), dt=dt)
    for epi, act, slope in ((0, 2, 0.0), (0, 1, 0.05), (1, 2, 0.0), (1, 1, 0.05), (1, 0, 0.0)):
        o_ref = torch.empty_like(L)
        o = torch.empty_like(L, device='cuda')
        ref.col2im(col, o_ref, n, g, kpad, bias if epi == 0 else None, href if epi == 1 else None, epi, act, slope)
        ops.col2im(col.cuda(), o, n, g, kpad, bias.cuda() if epi == 0 else None, href.cuda() if epi == 1 else None,
                   epi, act, slope)
        close(o, o_ref, tol(dt), f'col2im epi{epi} act{act}')


# ---- crowd KnnDenseNetCat: conv shapes of the graph + the graph ops of csrc/graph_ops.cu
CROWD_GEOMS = [
    Geom(16, 16, 16, 32, 32, 3, 7, 7, 2, 3),        # stem conv k7 s2 p3
    Geom(8, 8, 64, 8, 8, 40, 1, 1, 1, 0),           # dense-layer bottleneck 1x1, C not a multiple of 64
    Geom(8, 8, 32, 8, 8, 128, 3, 3, 1, 1),          # dense-layer 3x3, growth 32
    Geom(7, 7, 32, 7, 7, 128, 3, 3, 1, 1),          # ... at the 7x7 stage
    Geom(4, 4, 24, 32, 32, 1, 8, 8, 8, 0),          # MapModule ConvTranspose2d k = stride = 8 (pair: small side 4x4x24)
    Geom(16, 16, 8, 32, 32, 1, 2, 2, 2, 0),         # MapModule conv1 1 -> 8, k2 s2
    Geom(4, 4, 32, 8, 8, 16, 2, 2, 2, 0),           # MapModule conv3
    Geom(1, 1, 20, 4, 4, 32, 4, 4, 1, 0),           # MapModule linear1: full-extent conv
    Geom(1, 1, 20, 1, 1, 1088, 1, 1, 1, 0),         # final_count_feature_layer
]


@pytest.mark.parametrize('dt', DT)
@pytest.mark.parametrize('gi', range(len(CROWD_GEOMS)))
def test_crowd_conv_shapes(ops, dt, gi):
    g = CROWD_GEOMS[gi]
    gen = torch.Generator().manual_seed(100 + gi)
    n = 3
    ref = TorchOps()
    L = rnd(gen, n * g.Hl * g.Wl * g.Cb, dt=dt)
    S = rnd(gen, n * g.Hs * g.Ws * g.Ca, dt=dt)
    Wd = (rnd(gen, g.Ca * g.R * g.S * g.Cb) * 0.3).to(dt)
    Wu = Wd.view(g.Ca, g.R, g.S, g.Cb).permute(3, 1, 2, 0).contiguous().view(-1)
    bias_a, bias_b = rnd(gen, g.Ca), rnd(gen, g.Cb)
    for epi, act, slope in ((0, 1, 0.01), (0, 0, 0.0), (1, 1, 0.0), (1, 0, 0.0)):
        href = rnd(gen, S.numel(), dt=dt)
        out_ref = torch.empty_like(S)
        ref.conv_down(L, Wd, out_ref, n, g, bias_a if epi == 0 else None, 0, href if epi == 1 else None, epi, act, slope)
        out = torch.empty_like(S, device='cuda')
        ops.conv_down(L.cuda(), Wd.cuda(), out, n, g, bias_a.cuda() if epi == 0 else None, 0,
                      href.cuda() if epi == 1 else None, epi, act, slope)
        close(out, out_ref, tol(dt), f'down epi{epi} act{act}')
        href = rnd(gen, L.numel(), dt=dt)
        out_ref = torch.empty_like(L)
        ref.conv_up(S, Wu, out_ref, n, g, bias_b if epi == 0 else None, 0, href if epi == 1 else None, epi, act, slope)
        out = torch.empty_like(L, device='cuda')
        ops.conv_up(S.cuda(), Wu.cuda(), out, n, g, bias_b.cuda() if epi == 0 else None, 0,
                    href.cuda() if epi == 1 else None, epi, act, slope)
        close(out, out_ref, tol(dt), f'up epi{epi} act{act}')
    dW_ref = rnd(gen, Wd.numel())
    dW = dW_ref.clone().cuda()
    ref.conv_wgrad(S, L, dW_ref, n, g)
    ops.conv_wgrad(S.cuda(), L.cuda(), dW, n, g)
    close(dW, dW_ref, tol(dt) * 2, 'wgrad')


@pytest.mark.parametrize('dt', DT)
@pytest.mark.parametrize('rows,pitch,c0,C,yp', [(3 * 56 * 56, 256, 0, 96, 128), (1000, 40, 8, 24, 24), (77, 13, 3, 7, 9),
                                                 (5, 1920, 0, 1920, 1920), (20011, 200, 4, 100, 100),
                                                 (4 * 196 * 16 + 3, 1024, 64, 640, 640)])
def test_affine_ops(ops, dt, rows, pitch, c0, C, yp):
    gen = torch.Generator().manual_seed(rows + C)
    ref = TorchOps()
    x = rnd(gen, rows * pitch, dt=dt)
    gamma, beta, mean = rnd(gen, C) + 1.5, rnd(gen, C) * 0.3, rnd(gen, C) * 0.2
    var = torch.rand(C, generator=gen) + 0.5
    href = rnd(gen, rows * yp, dt=dt)
    cu = lambda t: t.cuda()
    for mode, act, slope in ((0, 1, 0.0), (0, 0, 0.0), (1, 1, 0.0), (1, 1, 0.01)):
        y_ref = rnd(gen, rows * yp, dt=dt)                  # pad columns must survive
        y = y_ref.clone().cuda()
        ref.affine(x, pitch, c0, y_ref, yp, rows, C, gamma, beta, mean, var, 1e-5, href, mode, act, slope)
        ops.affine(cu(x), pitch, c0, y, yp, rows, C, cu(gamma), cu(beta), cu(mean), cu(var), 1e-5, cu(href), mode, act, slope)
        close(y, y_ref, tol(dt), f'affine mode{mode}')
    dy = rnd(gen, rows * yp, dt=dt)
    for acc in (False, True):
        dx_ref = rnd(gen, rows * pitch, dt=dt)
        dx = dx_ref.clone().cuda()
        ref.affine_bwd(dy, yp, dx_ref, pitch, c0, rows, C, gamma, var, 1e-5, acc)
        ops.affine_bwd(cu(dy), yp, dx, pitch, c0, rows, C, cu(gamma), cu(var), 1e-5, acc)
        close(dx, dx_ref, tol(dt), f'affine_bwd acc{acc}')
    for sub in (True, False):
        dg_ref, db_ref = rnd(gen, C), rnd(gen, C)
        dg, db = dg_ref.clone().cuda(), db_ref.clone().cuda()
        ref.affine_grad(dy, yp, x, pitch, c0, rows, C, mean, var, 1e-5, dg_ref, db_ref if sub else None, sub)
        ops.affine_grad(cu(dy), yp, cu(x), pitch, c0, rows, C, cu(mean), cu(var), 1e-5, dg, db if sub else None, sub)
        t = 1e-4 if dt == torch.float32 else 1e-2
        close(dg, dg_ref, t, 'affine_grad dgamma')
        close(db, db_ref, t, 'affine_grad dbeta')
    for acc in (False, True):                            # the fused backward: parameter gradients + data gradient in one pass
        dx_ref = rnd(gen, rows * pitch, dt=dt)
        dx = dx_ref.clone().cuda()
        dg_ref, db_ref = rnd(gen, C), rnd(gen, C)
        dg, db = dg_ref.clone().cuda(), db_ref.clone().cuda()
        ref.affine_bwd_grad(dy, yp, x, dx_ref, pitch, c0, rows, C, gamma, mean, var, 1e-5, dg_ref, db_ref, acc)
        ops.affine_bwd_grad(cu(dy), yp, cu(x), dx, pitch, c0, rows, C, cu(gamma), cu(mean), cu(var), 1e-5, dg, db, acc)
        t = 1e-4 if dt == torch.float32 else 1e-2
        close(dx, dx_ref, tol(dt), f'affine_bwd_grad dx acc{acc}')
        close(dg, dg_ref, t, 'affine_bwd_grad dgamma')
        close(db, db_ref, t, 'affine_bwd_grad dbeta')


@pytest.mark.parametrize('dt', DT)
def test_copy2d_and_pools(ops, dt):
    gen = torch.Generator().manual_seed(9)
    ref = TorchOps()
    cu = lambda t: t.cuda()
    for rows, sp, s0, dp, d0, C in ((500, 32, 0, 96, 64, 32), (300, 48, 8, 20, 0, 20), (41, 7, 2, 9, 3, 5),
                                    (30001, 64, 0, 1792, 256, 32), (9000, 512, 128, 384, 0, 380)):
        src = rnd(gen, rows * sp, dt=dt)
        for acc in (False, True):
            d_ref = rnd(gen, rows * dp, dt=dt)
            d = d_ref.clone().cuda()
            ref.copy2d(src, sp, s0, d_ref, dp, d0, rows, C, acc)
            ops.copy2d(cu(src), sp, s0, d, dp, d0, rows, C, acc)
            close(d, d_ref, tol(dt), 'copy2d')
    # max-pool 3/2/1 (stem) incl. ReLU-style ties at zero, odd extent; tangent routing; backward
    for n, H, W, C, k, s, p in ((2, 16, 16, 8, 3, 2, 1), (3, 9, 11, 5, 3, 2, 1), (2, 8, 8, 4, 2, 2, 0), (2, 14, 14, 64, 3, 2, 1),
                                (1, 6, 6, 2, 3, 2, 1)):
        x = torch.relu(rnd(gen, n * H * W * C)).to(dt)
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        pitch, c0 = (C + 8, 4) if C % 4 == 0 else (C + 6, 2)        # 4-aligned slices take the 4-channel kernels
        y_ref = torch.zeros(n * Ho * Wo * pitch, dtype=dt)
        y = y_ref.clone().cuda()
        ref.maxpool(x, None, y_ref, pitch, c0, n, H, W, C, k, s, p)
        ops.maxpool(cu(x), None, y, pitch, c0, n, H, W, C, k, s, p)
        close(y, y_ref, 1e-6, 'maxpool')
        v = rnd(gen, n * H * W * C, dt=dt)
        ref.maxpool(v, x, y_ref, pitch, c0, n, H, W, C, k, s, p)
        ops.maxpool(cu(v), cu(x), y, pitch, c0, n, H, W, C, k, s, p)
        close(y, y_ref, 1e-6, 'maxpool tangent')
        dy = rnd(gen, n * Ho * Wo * pitch, dt=dt)
        dx_ref = torch.empty(n * H * W * C, dtype=dt)
        dx = torch.empty(n * H * W * C, dtype=dt, device='cuda')
        ref.maxpool_bwd(x, dy, pitch, c0, dx_ref, n, H, W, C, k, s, p, 1, 0.0)
        ops.maxpool_bwd(cu(x), cu(dy), pitch, c0, dx, n, H, W, C, k, s, p, 1, 0.0)
        close(dx, dx_ref, tol(dt), 'maxpool_bwd')
        # index-map variants: the forward writes one byte per pooled element, tangent routing and backward read it
        idx_ref = torch.zeros(n * Ho * Wo * C, dtype=torch.uint8)
        idx = idx_ref.clone().cuda()
        ref.maxpool(x, None, y_ref, pitch, c0, n, H, W, C, k, s, p, idx=idx_ref, idx_mode=1)
        ops.maxpool(cu(x), None, y, pitch, c0, n, H, W, C, k, s, p, idx=idx, idx_mode=1)
        close(y, y_ref, 1e-6, 'maxpool idx fwd')
        assert torch.equal(idx.cpu(), idx_ref), 'index maps differ (first-maximum rule)'
        ref.maxpool(v, None, y_ref, pitch, c0, n, H, W, C, k, s, p, idx=idx_ref, idx_mode=2)
        ops.maxpool(cu(v), None, y, pitch, c0, n, H, W, C, k, s, p, idx=idx, idx_mode=2)
        close(y, y_ref, 1e-6, 'maxpool idx route')
        ref.maxpool_bwd(x, dy, pitch, c0, dx_ref, n, H, W, C, k, s, p, 1, 0.0, idx=idx_ref)
        ops.maxpool_bwd(cu(x), cu(dy), pitch, c0, dx, n, H, W, C, k, s, p, 1, 0.0, idx=idx)
        close(dx, dx_ref, tol(dt), 'maxpool_bwd idx')
    # (.., pad): destination slice at channel offset `pad` of rows C + pad wide; pad = 8 with C % 8 == 0 takes the
    # 8-channel vector kernels in bf16, pad = 4 the 4-channel ones
    for n, H, W, C, k, xp, pad in ((3, 8, 8, 12, 2, 12, 4), (2, 7, 7, 40, 7, 64, 4), (1, 4, 4, 3, 2, 5, 4), (3, 14, 14, 64, 2, 64, 8),
                                   (2, 7, 7, 40, 7, 64, 8), (5, 28, 28, 256, 2, 256, 0)):
        x = rnd(gen, n * H * W * xp, dt=dt)
        Ho, Wo = H // k, W // k
        pitch, c0 = C + pad, pad
        y_ref = torch.zeros(n * Ho * Wo * pitch, dtype=dt)
        y = y_ref.clone().cuda()
        ref.avgpool(x, xp, y_ref, pitch, c0, n, H, W, C, k)
        ops.avgpool(cu(x), xp, y, pitch, c0, n, H, W, C, k)
        close(y, y_ref, tol(dt), 'avgpool')
        dy = rnd(gen, n * Ho * Wo * pitch, dt=dt)
        for act in (0, 1):
            dx_ref = rnd(gen, n * H * W * xp, dt=dt)
            dx = dx_ref.clone().cuda()
            ref.avgpool_bwd(dy, pitch, c0, dx_ref, xp, n, H, W, C, k, x, act, 0.0)
            ops.avgpool_bwd(cu(dy), pitch, c0, dx, xp, n, H, W, C, k, cu(x), act, 0.0)
            close(dx, dx_ref, tol(dt), 'avgpool_bwd')


@pytest.mark.parametrize('dt', DT)
@pytest.mark.parametrize('order', [1, 2, 3])
def test_crowd_loss_ops(ops, dt, order):
    gen = torch.Generator().manual_seed(order)
    ref = TorchOps()
    B, HW = 5, 56 * 56
    pred = rnd(gen, B) * 30
    density = (torch.rand(B, HW, generator=gen) < 0.01).float()
    label = 1 / (1 + torch.rand(B, HW, generator=gen) * 50)
    maps = [(torch.rand(B * HW, generator=gen) * 0.5 - 0.1).to(dt) for _ in range(3)]
    l_ref, dp_ref, dm_ref = torch.tensor([0.25]), torch.empty(B), torch.empty(B)
    ref.crowd_loss(pred, density, maps, label, B, HW, order, 0.7, 1e-3, l_ref, dp_ref, dm_ref)
    l, dp, dm = torch.tensor([0.25]).cuda(), torch.empty(B, device='cuda'), torch.empty(B, device='cuda')
    ops.crowd_loss(pred.cuda(), density.cuda(), [m.cuda() for m in maps], label.cuda(), B, HW, order, 0.7, 1e-3, l, dp, dm)
    t = 1e-4 if dt == torch.float32 else 1e-2
    close(l, l_ref, t, 'crowd loss')
    close(dp, dp_ref, t, 'dpred')
    close(dm, dm_ref, t, 'dm')
    d_ref = rnd(gen, B * HW, dt=dt)
    d = d_ref.clone().cuda()
    ref.crowd_map_grad(maps[1], label, dm_ref, d_ref, B, HW, 3, 1, 0.01)
    ops.crowd_map_grad(maps[1].cuda(), label.cuda(), dm_ref.cuda(), d, B, HW, 3, 1, 0.01)
    close(d, d_ref, tol(dt), 'crowd_map_grad')


@pytest.mark.parametrize('dt', DT)
def test_depth_to_space_and_scalar_bias_gemm(ops, dt):
    """MapModule ConvTranspose2d (kernel = stride, one output channel) = GEMM with ONE scalar bias (bias_mod 1, tcgen05 path
    in bf16) + depth-to-space; and the skinny full-extent pair (K = 25088 -> 20 outputs) at an odd row count."""
    gen = torch.Generator().manual_seed(4)
    ref = TorchOps()
    n, Hs, Ws, k = 3, 7, 5, 4
    blk = rnd(gen, n * Hs * Ws * k * k, dt=dt)
    img_ref, img = torch.empty_like(blk), torch.empty_like(blk, device='cuda')
    ref.depth_to_space(blk, img_ref, n, Hs, Ws, k, False)
    ops.depth_to_space(blk.cuda(), img, n, Hs, Ws, k, False)
    close(img, img_ref, 1e-7, 'depth_to_space')
    back = torch.empty_like(blk, device='cuda')
    ops.depth_to_space(img, back, n, Hs, Ws, k, True)
    close(back, blk, 1e-7, 'space_to_depth')
    g = Geom(1, 1, 64, 1, 1, 128, 1, 1, 1, 0)
    rows = 3 * 28 * 28
    L, Wd, bias = rnd(gen, rows * 128, dt=dt), (rnd(gen, 64 * 128) * 0.2).to(dt), rnd(gen, 1)
    o_ref, o = torch.empty(rows * 64, dtype=dt), torch.empty(rows * 64, dtype=dt, device='cuda')
    ref.conv_down(L, Wd, o_ref, rows, g, bias, 1, None, 0, 1, 0.01)
    ops.conv_down(L.cuda(), Wd.cuda(), o, rows, g, bias.cuda(), 1, None, 0, 1, 0.01)
    close(o, o_ref, tol(dt), 'scalar-bias gemm')
    if dt == torch.bfloat16:
        assert ops.lib.srgan_last_path_tensor() == 1
    g = Geom(1, 1, 20, 28, 28, 32, 28, 28, 1, 0)
    n = 5
    K = 28 * 28 * 32
    L, S = rnd(gen, n * K, dt=dt), rnd(gen, n * 20, dt=dt)
    Wd = (rnd(gen, 20 * K) * 0.05).to(dt)
    Wu = Wd.view(20, K).t().contiguous().view(-1)
    b = rnd(gen, 20)
    o_ref, o = torch.empty(n * 20, dtype=dt), torch.empty(n * 20, dtype=dt, device='cuda')
    ref.conv_down(L, Wd, o_ref, n, g, b, 0, None, 0, 1, 0.01)
    ops.conv_down(L.cuda(), Wd.cuda(), o, n, g, b.cuda(), 0, None, 0, 1, 0.01)
    close(o, o_ref, tol(dt), 'skinny down')
    href = rnd(gen, n * K, dt=dt)
    o_ref, o = torch.empty(n * K, dtype=dt), torch.empty(n * K, dtype=dt, device='cuda')
    ref.conv_up(S, Wu, o_ref, n, g, None, 0, href, 1, 1, 0.01)
    ops.conv_up(S.cuda(), Wu.cuda(), o, n, g, None, 0, href.cuda(), 1, 1, 0.01)
    close(o, o_ref, tol(dt), 'skinny up')
    dW_ref = rnd(gen, 20 * K)
    dW = dW_ref.clone().cuda()
    ref.conv_wgrad(S, L, dW_ref, n, g)
    ops.conv_wgrad(S.cuda(), L.cuda(), dW, n, g)
    close(dW, dW_ref, tol(dt) * 2, 'skinny wgrad')


# ------------------------------------------------------------------------------------------------ SGAN K-logit head (sgan.cu)
@pytest.mark.parametrize('dt', DT)
@pytest.mark.parametrize('rows,cols,K', [(100, 32768, 10), (64, 100, 10), (5, 36, 3), (257, 512, 16), (1, 8, 4)])
def test_sgan_head_kernels(ops, dt, rows, cols, K):
    """srgan_head_logits / srgan_sgan_loss (cross entropy, BCE of logsumexp) / srgan_sgan_gp_second / srgan_seed_rows_multi /
    srgan_head_wgrad against the op-level torch semantics; shapes: the age head (F = 32 768, 10 bins), SganMLP (F = 100), ragged."""
    gen = torch.Generator().manual_seed(rows + cols + K)
    ref = TorchOps()
    t = 1e-4 if dt == torch.float32 else 1e-2
    X = rnd(gen, rows * cols, dt=dt)
    W, bias = rnd(gen, K * cols) * (4.0 / cols ** 0.5), rnd(gen, K)
    for b in (bias, None):
        l_ref, l = torch.empty(K, rows), torch.empty(K, rows, device='cuda')
        ref.head_logits(X, rows, cols, W, b, K, l_ref)
        ws = torch.empty(ops.head_logits_workspace(rows, cols, K), device='cuda')
        ops.head_logits(X.cuda(), rows, cols, W.cuda(), b.cuda() if b is not None else None, K, l, ws)
        close(l, l_ref, t, 'head_logits')
    logits = rnd(gen, K, rows) * 3
    y, bins = rnd(gen, rows) * 50 + 50, torch.linspace(10, 95, K)
    for mode, target, scale in ((0, 0.0, 0.37), (1, 1.0, 1.0 / rows), (1, 0.0, -2.5)):
        loss_ref, d_ref = torch.full((1,), 0.25), torch.empty(K, rows)
        loss, d = loss_ref.clone().cuda(), torch.empty(K, rows, device='cuda')
        ref.sgan_loss(logits, K, rows, mode, y, bins, target, scale, loss_ref, d_ref)
        ops.sgan_loss(logits.cuda(), K, rows, mode, y.cuda(), bins.cuda(), target, scale, loss, d)
        close(loss, loss_ref, 2e-5, f'sgan_loss mode {mode}')
        close(d, d_ref, 2e-5, f'sgan_loss gradient mode {mode}')
    ops.sgan_loss(logits.cuda(), K, rows, 1, None, None, 0.0, 1.0, None, d)          # gradient only (the penalty's seed)
    tang = rnd(gen, K, rows)
    q_ref, q = torch.empty(K, rows), torch.empty(K, rows, device='cuda')
    ref.sgan_gp_second(logits, tang, K, rows, 0.7, q_ref)
    ops.sgan_gp_second(logits.cuda(), tang.cuda(), K, rows, 0.7, q)
    close(q, q_ref, 2e-5, 'sgan_gp_second')
    dT = rnd(gen, K, rows)
    for act, slope in ((1, 0.05), (0, 0.0)):
        o_ref, o = torch.empty(rows * cols, dtype=dt), torch.empty(rows * cols, dtype=dt, device='cuda')
        ref.seed_rows_multi(o_ref, rows, cols, dT, W, K, X, act, slope)
        ops.seed_rows_multi(o, rows, cols, dT.cuda(), W.cuda(), K, X.cuda(), act, slope)
        close(o, o_ref, tol(dt), 'seed_rows_multi')
    if cols % 4 == 0:
        for with_bias in (True, False):
            dW_ref, db_ref = rnd(gen, K * cols), rnd(gen, K)
            dW, db = dW_ref.clone().cuda(), db_ref.clone().cuda()
            ref.head_wgrad(X, rows, cols, dT, K, dW_ref, db_ref if with_bias else None)
            ops.head_wgrad(X.cuda(), rows, cols, dT.cuda(), K, dW, db if with_bias else None)
            close(dW, dW_ref, t, 'head_wgrad')
            close(db, db_ref, 1e-4, 'head_wgrad bias')
