"""Measured relative errors of the bf16 tensor-core mode against the reference / oracle scalars, per step scalar, for the
parity cases of tests/test_gpu_parity.py (the bands those tests assert are stated next to them): the evidence behind the
tolerance table in DESIGN.md section 2.  usage: python tools/bf16_parity_report.py > profiles/r2_bf16_parity_errors.txt"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import srgan_oracle as O
from tests.golden_io import Golden, SCALARS, GOLDEN_DIR
from tests.gpu_common import runner_from_state, to_cuda


def report(name, got, ref, band):
    floor = 1e-3 * max(1e-3, abs(ref['labeled_loss']))
    parts = []
    for k in SCALARS:
        e = abs(got[k] - ref[k]) / max(abs(ref[k]), floor)
        parts.append(f'{k}={e:.2e}')
    print(f'{name:44s} band {band}:  ' + '  '.join(parts), flush=True)


torch.set_num_threads(os.cpu_count())
for precision in ('fp32', 'bf16'):
    print(f'---- precision mode {precision} (relative error of each step scalar; gradient_penalty is a hinge: a relative error e of the '
          f'gradient norm r becomes 2e*r/(r-1) on it)')
    band = '1e-4' if precision == 'fp32' else '2e-2 (penalty 8e-2)'
    # reference golden vectors, first step
    for name in ('coefficient_srgan', 'coefficient_dggan', 'dcgan_mini'):
        g = Golden(name)
        st, cfg = g.oracle_state(), g.step_config()
        r = runner_from_state(st, cfg, precision)
        if precision == 'bf16':
            r.persistent = False
        x, y, u, z, alpha, z2 = to_cuda(*g.step_inputs(0))
        r.dnn_step(x, y, lr=O.dnn_lr(cfg, 0))
        r.gan_step(x, y, u, 0, noise=(z, alpha, z2))
        report(f'golden {name} step 0', r.scalars(), g.scalars(0), band)
    # age at full size, B = 100, D conv weights x3 (hinge active)
    B = 100
    gen = torch.Generator().manual_seed(41)
    st = O.init_dcgan(seed=4, image_size=128, conv_dim=64, z_dim=256, scale=3.0)
    cfg = O.StepConfig(batch_size=B, matching_loss_multiplier=1e2, contrasting_loss_multiplier=1e1, gradient_penalty_multiplier=1e2)
    x, u = torch.rand(B, 3, 128, 128, generator=gen) * 2 - 1, torch.rand(B, 3, 128, 128, generator=gen) * 2 - 1
    y = torch.rand(B, generator=gen) * 85 + 10
    z, alpha, z2 = torch.randn(B, 256, generator=gen), torch.rand(B, 1, 1, 1, generator=gen), torch.randn(B, 256, generator=gen)
    r = runner_from_state(st, cfg, precision)
    ref = O.training_step(st, cfg, x, y, u, z, alpha, z2, step=0)
    xc, yc, uc, zc, ac, z2c = to_cuda(x, y, u, z, alpha, z2)
    r.dnn_step(xc, yc)
    r.gan_step(xc, yc, uc, 0, noise=(zc, ac, z2c))
    report('age full size B=100 vs oracle', r.scalars(), ref, band)
    del r
    torch.cuda.empty_cache()
    # crowd: reduced net vs oracle, full DenseNet-201 vs the reference golden
    kw = dict(block_config=(2, 2, 2, 2), growth_rate=8, num_init_features=16, bn_size=2, label_patch_size=64)
    st = O.init_crowd(seed=1, image_size=64, z_dim=16, g_conv_dim=8, scale=2.0, **kw)
    cfg = O.StepConfig(batch_size=3, matching_loss_multiplier=1e3, contrasting_loss_multiplier=1e2, gradient_penalty_multiplier=1e2,
                       map_multiplier=1e-3)
    r = runner_from_state(st, cfg, precision)
    b = O.synthetic_crowd_batch(3, 10, image=64, label=64, z_dim=16)
    ref = O.training_step(st, cfg, *b, step=0)
    xc, yc, uc, zc, ac, z2c = to_cuda(*b)
    r.dnn_step(xc, yc)
    r.gan_step(xc, yc, uc, 0, noise=(zc, ac, z2c))
    report('crowd reduced (2,2,2,2) vs oracle', r.scalars(), ref, band)
    for method in ('srgan', 'dggan'):
        zf = np.load(os.path.join(GOLDEN_DIR, f'crowd_{method}.npz'))
        cfgj = json.loads(bytes(zf['config_json']).decode())
        cfg = O.StepConfig()
        for k, v in cfgj.items():
            if hasattr(cfg, k):
                setattr(cfg, k, v)
        st = O.init_crowd(seed=cfgj['init_seed'], scale=cfgj['d_scale'], dggan=(method == 'dggan'))
        r = runner_from_state(st, cfg, precision)
        x, y, u, zz, alpha, z2 = to_cuda(*O.synthetic_crowd_batch(2, cfgj['input_seed']))
        r.dnn_step(x, y)
        r.gan_step(x, y, u, 0, noise=(zz, alpha, z2))
        ref = {k: float(zf[f'step0/scalars/{k}']) for k in SCALARS}
        report(f'crowd DenseNet-201 {method} vs reference golden', r.scalars(), ref,
               '1e-4' if precision == 'fp32' else '5e-2 (penalty 2e-1)')
        del r
        torch.cuda.empty_cache()
