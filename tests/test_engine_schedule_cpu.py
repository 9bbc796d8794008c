"""Host-logic test (CPU): the explicit schedule in sr-gan_b200/engine.py, driven through the TEST-ONLY torch emulation of
the op set (tests/torch_ops.py), must reproduce the oracle step (autograd + double-backward) and the reference golden
vectors.  This isolates schedule/algebra bugs from kernel bugs; the CUDA kernels are checked by the -m gpu tests."""
import copy

import pytest
import torch

from oracle import srgan_oracle as O
from srgan_b200 import nets, engine
from tests.golden_io import Golden, SCALARS
from tests.torch_ops import TorchOps


def build_engine(st: O.OracleState, dtype, image_size=None, conv_dim=None, z_dim=None, comm=None):
    if st.d_spec.family == 'coefficient':
        n_out, hidden = st.D['linear4.weight'].shape
        d_net = nets.coefficient_d(hidden, 50, st.d_spec.dggan, n_out=n_out)
        g_net = nets.coefficient_g(st.G['linear1.weight'].shape[0], 10, 50)
    else:
        d_net = nets.dcgan_d(image_size, conv_dim, n_out=st.D['layer5.0.weight'].shape[0])
        g_net = nets.dcgan_g(image_size, conv_dim, z_dim)
    D = {k: v.clone() for k, v in st.D.items()}
    G = {k: v.clone() for k, v in st.G.items()}
    DNN = {k: v.clone() for k, v in st.DNN.items()}
    return engine.Engine(TorchOps(), d_net, g_net, D, G, DNN, act_dtype=dtype, device='cpu', comm=comm)


def rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def read_scalars(eng):
    s = eng.scalars.tolist()
    return dict(dnn_loss=s[0], labeled_loss=s[1], unlabeled_loss=s[2], fake_loss=s[3], gradient_penalty=s[4],
                gradient_norm_mean=s[5], generator_loss=s[6])


@pytest.mark.parametrize('name', ['coefficient_srgan', 'coefficient_srgan_altdist', 'coefficient_dggan', 'dcgan_mini',
                                  'dcgan_sgan_mini', 'coefficient_sgan'])
def test_schedule_matches_oracle_fp64(name):
    g = Golden(name)
    dt = torch.float64
    st, cfg = g.oracle_state(dt), g.step_config()
    eng = build_engine(st, dt, g.cfg.get('image_size'), g.cfg.get('conv_dim'), g.cfg.get('z_dim'))
    for i in range(g.steps):
        x, y, u, z, alpha, z2 = g.step_inputs(i, dt)
        out = O.training_step(st, cfg, x, y, u, z, alpha, z2, step=i)
        eng.dnn_step(x, y, cfg, O.dnn_lr(cfg, i), cfg.weight_decay)
        eng.gan_step(x, y, u, z, alpha, z2, cfg)
        got = read_scalars(eng)
        for k in SCALARS:
            assert got[k] == pytest.approx(out[k], rel=1e-9, abs=1e-12), (name, i, k)
    for net, params, mine in (('D', st.D, eng.D), ('G', st.G, eng.G), ('DNN', st.DNN, eng.DNN)):
        for k, v in params.items():
            assert rel(mine.params[k], v) < 1e-8, (name, net, k, rel(mine.params[k], v))


@pytest.mark.parametrize('name', ['coefficient_srgan', 'coefficient_dggan', 'dcgan_mini'])
def test_schedule_matches_reference_golden_fp32(name):
    g = Golden(name)
    st, cfg = g.oracle_state(), g.step_config()
    eng = build_engine(st, torch.float32, g.cfg.get('image_size'), g.cfg.get('conv_dim'), g.cfg.get('z_dim'))
    for i in range(g.steps):
        x, y, u, z, alpha, z2 = g.step_inputs(i)
        eng.dnn_step(x, y, cfg, O.dnn_lr(cfg, i), cfg.weight_decay)
        eng.gan_step(x, y, u, z, alpha, z2, cfg)
        got, ref = read_scalars(eng), g.scalars(i)
        for k in SCALARS:
            assert got[k] == pytest.approx(ref[k], rel=1e-4, abs=1e-6), (name, i, k, got[k], ref[k])
    for net, mine in (('D', eng.D), ('G', eng.G), ('DNN', eng.DNN)):
        for k, v in g.group(f'final/{net}').items():
            assert rel(mine.params[k], v) < 1e-4, (name, net, k)


def test_thin_layer_lowering_matches_oracle_fp64():
    """The im2col/col2im + [pixels x 64] GEMM lowering of the 3-channel image layers (bf16 mode default) is algebra-
    exact: run it in fp64 through the torch emulation against the oracle (conv_dim 64 so the layers are eligible)."""
    dt = torch.float64
    st = O.init_dcgan(seed=4, image_size=32, conv_dim=64, z_dim=16, dtype=dt, scale=3.0)
    cfg = O.StepConfig(batch_size=2, matching_loss_multiplier=1e2, contrasting_loss_multiplier=1e1,
                       gradient_penalty_multiplier=1e2, weight_decay=1e-3)
    d_net, g_net = nets.dcgan_d(32, 64), nets.dcgan_g(32, 64, 16)
    eng = engine.Engine(TorchOps(), d_net, g_net, {k: v.clone() for k, v in st.D.items()},
                        {k: v.clone() for k, v in st.G.items()}, {k: v.clone() for k, v in st.DNN.items()},
                        act_dtype=dt, device='cpu', thin_lowering=True)
    assert eng.D.thin == {'layer1.0'} and eng.G.thin == {'layer4.0'}
    gen = torch.Generator().manual_seed(1)
    B = 2
    for i in range(2):
        x = (torch.rand(B, 3, 32, 32, generator=gen, dtype=dt) * 2 - 1)
        u = (torch.rand(B, 3, 32, 32, generator=gen, dtype=dt) * 2 - 1)
        y = torch.rand(B, generator=gen, dtype=dt) * 85 + 10
        z, alpha, z2 = (torch.randn(B, 16, generator=gen, dtype=dt), torch.rand(B, 1, 1, 1, generator=gen, dtype=dt),
                        torch.randn(B, 16, generator=gen, dtype=dt))
        out = O.training_step(st, cfg, x, y, u, z, alpha, z2, step=i)
        eng.dnn_step(x, y, cfg, O.dnn_lr(cfg, i), cfg.weight_decay)
        eng.gan_step(x, y, u, z, alpha, z2, cfg)
        got = read_scalars(eng)
        assert out['gradient_penalty'] > 0
        for k in SCALARS:
            assert got[k] == pytest.approx(out[k], rel=1e-9, abs=1e-12), (i, k)
    for net, params, mine in (('D', st.D, eng.D), ('G', st.G, eng.G), ('DNN', st.DNN, eng.DNN)):
        for k, v in params.items():
            assert rel(mine.params[k], v) < 1e-8, (net, k, rel(mine.params[k], v))


def build_crowd_engine(st, dtype, spec_kwargs, image, z_dim, g_conv_dim, direct_concat=False, fuse_bn=0):
    d_net = nets.knn_densenet_cat(spec_kwargs['block_config'], spec_kwargs['growth_rate'], spec_kwargs['num_init_features'],
                                  spec_kwargs['bn_size'], image, spec_kwargs['label_patch_size'],
                                  n_out=2 if st.d_spec.dggan else 1, direct_concat=direct_concat, fuse_bn=fuse_bn)
    g_net = nets.dcgan_g(image, g_conv_dim, z_dim)
    D = {k: v.clone() for k, v in st.D.items()}
    G = {k: v.clone() for k, v in st.G.items()}
    DNN = {k: v.clone() for k, v in st.DNN.items()}
    return engine.Engine(TorchOps(), d_net, g_net, D, G, DNN, act_dtype=dtype, device='cpu')


@pytest.mark.parametrize('direct,fuse', [(False, 0), (True, 0), (True, 1), (True, 2), (True, 3), (True, 4), (True, 5)])
@pytest.mark.parametrize('method', ['srgan', 'dggan'])
def test_crowd_graph_schedule_matches_oracle_fp64(method, direct, fuse):
    """Crowd SR-GAN (KnnDenseNetCat graph: eval-mode BatchNorm affine with trainable weight/bias, ReLU, max/avg pools,
    in-place concat, three MapModules, summed count heads, crowd labeled loss incl. the map term) through the explicit
    schedule -- forward, g-chain, tangent chain, one backward with the tangent-block gradients -- against the oracle's
    autograd double-backward, fp64, reduced DenseNet (2,2,2,2) at 64x64 (the full DenseNet-201 is pinned against the
    reference in tests/test_oracle_golden.py).  direct: the dense layers' 3x3 convolutions write / read their channel
    window of the concat buffers in place (srgan_views; the bf16 product path) instead of going through slice copies.
    fuse: the BatchNorm + ReLU in front of the trunk's 1x1 convolutions is carried out by those convolutions' kernels
    (1 = the data gradient, 2 = the forward pass too, 3 = norm2 / relu2 in conv1's forward epilogue, 4 = norm2 / relu2 backward in conv2's data gradient, 5 = conv1's weight gradient from the concat buffer; csrc/bn_gemm.cu)."""
    dt = torch.float64
    kw = dict(block_config=(2, 2, 2, 2), growth_rate=8, num_init_features=16, bn_size=2, label_patch_size=64)
    st = O.init_crowd(seed=1, image_size=64, z_dim=16, g_conv_dim=8, dtype=dt, scale=2.0, dggan=(method == 'dggan'), **kw)
    cfg = O.StepConfig(method=method, batch_size=3, matching_loss_multiplier=1e3, contrasting_loss_multiplier=1e2,
                       gradient_penalty_multiplier=1e2, map_multiplier=1e-3, weight_decay=1e-3)
    eng = build_crowd_engine(st, dt, kw, 64, 16, 8, direct_concat=direct, fuse_bn=fuse)
    assert sum(1 for op in eng.d_net.graph if op.fuse) == (11 if fuse else 0)      # 8 dense layers + 3 transitions
    assert any(op.C for op in eng.d_net.graph if op.kind == 'conv') == direct
    assert ('new.1.1' in eng.d_net.bufs) != direct
    B = 3
    for i in range(2):
        x, y, u, z, alpha, z2 = O.synthetic_crowd_batch(B, 10 + i, image=64, label=64, z_dim=16, dtype=dt)
        out = O.training_step(st, cfg, x, y, u, z, alpha, z2, step=i)
        eng.dnn_step(x, y, cfg, O.dnn_lr(cfg, i), cfg.weight_decay)
        eng.gan_step(x, y, u, z, alpha, z2, cfg)
        got = read_scalars(eng)
        assert out['gradient_penalty'] > 0
        for k in SCALARS:
            assert got[k] == pytest.approx(out[k], rel=1e-9, abs=1e-12), (i, k, got[k], out[k])
    for net, params, mine in (('D', st.D, eng.D), ('G', st.G, eng.G), ('DNN', st.DNN, eng.DNN)):
        for k, v in params.items():
            if O.is_buffer_key(k):
                assert torch.equal(mine.params[k], v)
            else:
                assert rel(mine.params[k], v) < 1e-8, (net, k, rel(mine.params[k], v))


def test_crowd_container_matches_reference_keys_and_graph():
    """The product's KnnDenseNetCat parameter container has the reference module's state_dict keys, order and shapes (the
    oracle's table is pinned against the reference by oracle/make_golden.py's strict load), and describe_module maps
    it to the same graph nets.knn_densenet_cat builds."""
    import srgan_b200
    for kw, image in ((dict(), 224), (dict(block_config=(2, 2, 2, 2), growth_rate=8, num_init_features=16, bn_size=2,
                                           label_patch_size=64), 64)):
        spec = O.ModelSpec('crowd', **kw)
        m = srgan_b200.KnnDenseNetCat(image_size=image, **kw)
        shapes = O.crowd_param_shapes(spec, image)
        sd = m.state_dict()
        assert list(sd.keys()) == list(shapes.keys())
        assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in shapes)
        net = nets.describe_module(m)
        assert net == nets.knn_densenet_cat(spec.block_config, spec.growth_rate, spec.num_init_features, spec.bn_size,
                                            image, spec.label_patch_size)
    full = nets.knn_densenet_cat()
    assert abs(full.macs_per_sample() - 4.366e9) < 2e6          # SURVEY App. A.4 [probed]


def test_crowd_stem_thin_lowering_matches_oracle_fp64():
    """The crowd stem (Conv2d(3, 64, k7 s2 p3): receptive field 147 -> im2col rows of 192) through the thin-layer lowering
    (im2col + [pixels x 192] GEMM, GEMM^T + col2im for the data gradient, linear wgrad on the saved col buffer), fp64."""
    dt = torch.float64
    kw = dict(block_config=(1, 1, 1, 1), growth_rate=8, num_init_features=64, bn_size=2, label_patch_size=64)
    st = O.init_crowd(seed=3, image_size=64, z_dim=16, g_conv_dim=8, dtype=dt, scale=2.0, **kw)
    cfg = O.StepConfig(batch_size=2, matching_loss_multiplier=1e3, contrasting_loss_multiplier=1e2,
                       gradient_penalty_multiplier=1e2, map_multiplier=1e-3)
    d_net = nets.knn_densenet_cat(kw['block_config'], kw['growth_rate'], kw['num_init_features'], kw['bn_size'], 64, 64)
    g_net = nets.dcgan_g(64, 8, 16)
    eng = engine.Engine(TorchOps(), d_net, g_net, {k: v.clone() for k, v in st.D.items()},
                        {k: v.clone() for k, v in st.G.items()}, {k: v.clone() for k, v in st.DNN.items()},
                        act_dtype=dt, device='cpu', thin_lowering=True)
    assert eng.D.thin == {'conv_layer1.conv0'} and d_net.layers[0].kpad == 192
    x, y, u, z, alpha, z2 = O.synthetic_crowd_batch(2, 17, image=64, label=64, z_dim=16, dtype=dt)
    out = O.training_step(st, cfg, x, y, u, z, alpha, z2, step=0)
    eng.dnn_step(x, y, cfg, O.dnn_lr(cfg, 0), cfg.weight_decay)
    eng.gan_step(x, y, u, z, alpha, z2, cfg)
    got = read_scalars(eng)
    for k in SCALARS:
        assert got[k] == pytest.approx(out[k], rel=1e-9, abs=1e-12), (k, got[k], out[k])
    for net, params, mine in (('D', st.D, eng.D), ('G', st.G, eng.G), ('DNN', st.DNN, eng.DNN)):
        for k, v in params.items():
            if not O.is_buffer_key(k):
                assert rel(mine.params[k], v) < 1e-8, (net, k, rel(mine.params[k], v))


@pytest.mark.parametrize('family', ['dcgan', 'crowd'])
def test_micro_batched_step_is_exact_fp64(family):
    """Micro-batching (BASELINE configs[4]: per-GPU batches too large for the activation buffers) must reproduce the
    full-batch oracle step: feature sums are accumulated over the micro-batches before the distance losses, the
    per-sample terms are normalised by the full batch, gradients accumulate across micro-batches."""
    dt = torch.float64
    B, mb = 4, 2
    if family == 'dcgan':
        st = O.init_dcgan(seed=4, image_size=32, conv_dim=8, z_dim=16, dtype=dt, scale=3.0)
        cfg = O.StepConfig(batch_size=B, matching_loss_multiplier=1e2, contrasting_loss_multiplier=1e1,
                           gradient_penalty_multiplier=1e2)
        eng = build_engine(st, dt, 32, 8, 16)
        gen = torch.Generator().manual_seed(3)

        def batch():
            return ((torch.rand(B, 3, 32, 32, generator=gen, dtype=dt) * 2 - 1), torch.rand(B, generator=gen, dtype=dt) * 85 + 10,
                    (torch.rand(B, 3, 32, 32, generator=gen, dtype=dt) * 2 - 1), torch.randn(B, 16, generator=gen, dtype=dt),
                    torch.rand(B, 1, 1, 1, generator=gen, dtype=dt), torch.randn(B, 16, generator=gen, dtype=dt))
    else:
        kw = dict(block_config=(2, 2, 2, 2), growth_rate=8, num_init_features=16, bn_size=2, label_patch_size=64)
        st = O.init_crowd(seed=1, image_size=64, z_dim=16, g_conv_dim=8, dtype=dt, scale=2.0, **kw)
        cfg = O.StepConfig(batch_size=B, matching_loss_multiplier=1e3, contrasting_loss_multiplier=1e2,
                           gradient_penalty_multiplier=1e2, map_multiplier=1e-3)
        eng = build_crowd_engine(st, dt, kw, 64, 16, 8)
        seeds = iter(range(50, 60))

        def batch():
            return O.synthetic_crowd_batch(B, next(seeds), image=64, label=64, z_dim=16, dtype=dt)
    for i in range(2):
        x, y, u, z, alpha, z2 = batch()
        out = O.training_step(st, cfg, x, y, u, z, alpha, z2, step=i)
        eng.dnn_step_micro(x, y, cfg, O.dnn_lr(cfg, i), cfg.weight_decay, mb)
        eng.gan_step_micro(x, y, u, z, alpha, z2, cfg, True, mb)
        got = read_scalars(eng)
        assert out['gradient_penalty'] > 0
        for k in SCALARS:
            assert got[k] == pytest.approx(out[k], rel=1e-9, abs=1e-12), (family, i, k, got[k], out[k])
    for net, params, mine in (('D', st.D, eng.D), ('G', st.G, eng.G), ('DNN', st.DNN, eng.DNN)):
        for k, v in params.items():
            if not O.is_buffer_key(k):
                assert rel(mine.params[k], v) < 1e-8, (family, net, k, rel(mine.params[k], v))


def test_dnn_and_gan_steps_use_disjoint_scratch_scopes():
    """The DNN step may be in flight on its own stream while the GAN step runs (StepRunner._on_dnn_stream): the two step
    methods must not share a single workspace buffer, and the DNN step allocates B rows, not the 5B-row D workspace."""
    g = Golden('dcgan_mini')
    dt = torch.float64
    st, cfg = g.oracle_state(dt), g.step_config()
    eng = build_engine(st, dt, g.cfg.get('image_size'), g.cfg.get('conv_dim'), g.cfg.get('z_dim'))
    x, y, u, z, alpha, z2 = g.step_inputs(0, dt)
    eng.dnn_step(x, y, cfg, O.dnn_lr(cfg, 0), cfg.weight_decay)
    dnn_keys = set(eng._buf)
    assert dnn_keys and all(k[0] == 'dnn' for k in dnn_keys)
    eng.gan_step(x, y, u, z, alpha, z2, cfg)
    gan_keys = set(eng._buf) - dnn_keys
    assert gan_keys and all(k[0] == 'gan' for k in gan_keys)
    ptrs = {}
    for k, t in eng._buf.items():
        ptrs.setdefault(t.untyped_storage().data_ptr(), []).append(k)
    assert all(len(v) == 1 for v in ptrs.values()), 'two scratch keys share storage'
    B = x.shape[0]
    a_dnn, a_gan = eng._buf[('dnn', ('D', 'a', 1))], eng._buf[('gan', ('D', 'a', 1))]
    assert a_gan.numel() == 5 * a_dnn.numel() and a_dnn.numel() == B * eng.d_net.layers[0].out_elems


def test_crowd_graph_side_branches():
    """nets.knn_densenet_cat marks the three MapModules as side branches: each is one contiguous run of ops that reads a
    trunk concat buffer first and writes `features` last, and touches no buffer of another branch (Engine runs them on
    their own streams, forked at the tap and joined before the trunk adds into the tapped concat delta)."""
    net = nets.knn_densenet_cat(block_config=(2, 2, 2, 2), growth_rate=8, num_init_features=16, bn_size=2, image_size=64,
                                label_size=64)
    assert net.branch_taps() == {'cat2', 'cat3', 'cat4'}
    runs = [(b, [op for op in net.graph if op.branch == b]) for b in (1, 2, 3)]
    idx = {id(op): i for i, op in enumerate(net.graph)}
    owned = {}
    for b, ops_ in runs:
        pos = [idx[id(op)] for op in ops_]
        assert pos == list(range(pos[0], pos[0] + len(pos))), 'a branch is one contiguous run'
        assert ops_[0].kind == 'read' and ops_[0].src == f'cat{b + 1}' and ops_[0].c0 == 0
        assert ops_[-1].kind == 'copy' and ops_[-1].dst == 'features'
        for op in ops_[1:]:
            assert op.src not in net.branch_taps()
        for op in ops_[:-1]:
            assert owned.setdefault(op.dst, b) == b, 'a buffer written by two branches'
    trunk_dsts = {op.dst for op in net.graph if not op.branch}
    assert not (set(owned) & trunk_dsts), 'a branch writes a trunk buffer'
    # the trunk never reads a branch buffer: the only meeting points are the tap (read) and `features`
    assert not ({op.src for op in net.graph if not op.branch} & set(owned))
