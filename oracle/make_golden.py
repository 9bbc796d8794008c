"""
Generates tests/golden/*.npz by running the UNMODIFIED reference step (imported from /root/reference).

Run in the build container only:   python oracle/make_golden.py
The committed .npz files are what travels; tests/test_oracle_golden.py checks oracle/srgan_oracle.py against them
and tests/test_gpu_parity.py checks the CUDA path against them.

Each fixture holds: the reference's initial state_dicts (D, G, DNN), per-step inputs (x, y, u) and injected noise
(z, alpha, z2), the scalars the reference logged per step (srgan.py:268-270, 306-319), first-step D/G gradients as
left in `.grad` by the reference, and the final state_dicts + Adam moments after `steps` steps.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_harness  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

SCALAR_TAGS = {
    'dnn/Discriminator/Labeled Loss': 'dnn_loss',
    'gan/Discriminator/Labeled Loss': 'labeled_loss',
    'gan/Discriminator/Unlabeled Loss': 'unlabeled_loss',
    'gan/Discriminator/Fake Loss': 'fake_loss',
    'gan/Discriminator/Gradient Penalty': 'gradient_penalty',
    'gan/Discriminator/Gradient Norm': 'gradient_norm_mean',
    'gan/Generator/Loss': 'generator_loss',
}


def sd_np(module, prefix):
    return {f'{prefix}/{k}': v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def adam_np(opt, module, prefix):
    out = {}
    names = [k for k, _ in module.named_parameters()]
    for name, p in zip(names, opt.param_groups[0]['params']):
        s = opt.state.get(p, {})
        if 'exp_avg' in s:
            out[f'{prefix}/{name}/exp_avg'] = s['exp_avg'].cpu().numpy().copy()
            out[f'{prefix}/{name}/exp_avg_sq'] = s['exp_avg_sq'].cpu().numpy().copy()
    return out


def run_case(name, exp, cfg_json, batches, steps):
    data = {}
    data.update(sd_np(exp.D, 'init/D'))
    data.update(sd_np(exp.G, 'init/G'))
    data.update(sd_np(exp.DNN, 'init/DNN'))
    for step in range(steps):
        x, y, u, z, alpha, z2 = batches[step]
        for k, v in dict(x=x, y=y, u=u, z=z, alpha=alpha, z2=z2).items():
            data[f'step{step}/{k}'] = v.numpy().copy()
        exp.dnn_training_step(x, y, step)
        with ref_harness.injected_noise(z, alpha, z2):
            exp.gan_training_step(x, y, u, step)
        sc = ref_harness.last_scalars(exp)
        for tag, key in SCALAR_TAGS.items():
            data[f'step{step}/scalars/{key}'] = np.float64(sc[tag])
        data[f'step{step}/gradient_norm'] = exp.gradient_norm.detach().numpy().copy()
        if step == 0:
            # D.grad now holds the D-step gradient PLUS what the G-step backward deposited,
            # so only G grads are recorded from .grad; D-step grads are pinned through the Adam moments below.
            for k, p in exp.G.named_parameters():
                data[f'step0/g_grads/{k}'] = p.grad.detach().numpy().copy()
    data.update(sd_np(exp.D, 'final/D'))
    data.update(sd_np(exp.G, 'final/G'))
    data.update(sd_np(exp.DNN, 'final/DNN'))
    data.update(adam_np(exp.d_optimizer, exp.D, 'final_adam/D'))
    data.update(adam_np(exp.g_optimizer, exp.G, 'final_adam/G'))
    data.update(adam_np(exp.dnn_optimizer, exp.DNN, 'final_adam/DNN'))
    cfg_json = dict(cfg_json, steps=steps, torch=torch.__version__)
    data['config_json'] = np.frombuffer(json.dumps(cfg_json).encode(), dtype=np.uint8)
    path = os.path.join(OUT, f'{name}.npz')
    np.savez_compressed(path, **data)
    print(name, {k: float(v) for k, v in data.items() if '/scalars/' in k and k.startswith(f'step{steps - 1}')},
          os.path.getsize(path), 'bytes')


def settings_for(cfg):
    from settings import Settings
    s = Settings()
    for k, v in cfg.items():
        if hasattr(s, k):
            setattr(s, k, v)
    s.summary_step_period = 1
    import utility
    s.matching_distance_function = getattr(utility, cfg.get('matching_distance_function', 'abs_mean'))
    s.contrasting_distance_function = getattr(utility, cfg.get('contrasting_distance_function',
                                                               'abs_plus_one_sqrt_mean_neg'))
    return s


def coefficient_batches(batch, steps, seed):
    from coefficient.data import generate_polynomial_examples
    from utility import seed_all
    seed_all(seed)
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(steps):
        ex, lab = generate_polynomial_examples(batch, 10)
        uex, _ = generate_polynomial_examples(batch, 10)
        x, y, u = torch.tensor(ex), torch.tensor(lab.astype(np.float32)), torch.tensor(uex)
        # z like srgan.py:286-289 with mean_offset 0 (a two-component mixture of identical N(0,1))
        z = torch.randn(batch, 10, generator=gen)
        alpha = torch.rand(batch, 1, generator=gen)
        z2 = torch.randn(batch, 10, generator=gen)
        out.append((x, y, u, z, alpha, z2))
    return out


def crowd_case(method='srgan', d_scale=1.56):
    """CrowdExperiment (crowd/srgan.py; method 'dggan': CrowdDgganExperiment, crowd/dggan.py, with KnnDenseNetCatDggan) with the UNMODIFIED KnnDenseNetCat (DenseNet-201 trunk, crowd/models.py:1049-1166)
    and DCGenerator at 224x224, B=2, one step.  The 20.6 M-parameter initial state is not stored: both sides build it
    with oracle.init_crowd(seed) (pure torch, deterministic) and the reference modules load it with strict=True, which
    also pins the oracle's key/shape table against the reference state_dict.  Stored: scalars, per-sample gradient
    norms, labeled features, and per-tensor checksums of the parameter updates of D, G and DNN."""
    from oracle import srgan_oracle as O
    from crowd.srgan import CrowdExperiment
    from crowd.dggan import CrowdDgganExperiment
    from crowd.models import KnnDenseNetCat, KnnDenseNetCatDggan, DCGenerator
    dggan = method == 'dggan'
    torch.set_num_threads(os.cpu_count())
    cfg = dict(method=method, family='crowd', batch_size=2, learning_rate=1e-4, weight_decay=0.0,
               matching_loss_multiplier=1e3, contrasting_loss_multiplier=1e2, gradient_penalty_multiplier=1e2,
               map_multiplier=1e-3, init_seed=5, input_seed=6, d_scale=d_scale)
    st = O.init_crowd(seed=cfg['init_seed'], scale=cfg['d_scale'], dggan=dggan)
    cls = KnnDenseNetCatDggan if dggan else KnnDenseNetCat
    D, DNN, G = cls(pretrained=False), cls(pretrained=False), DCGenerator()
    D.load_state_dict(st.D, strict=True)
    DNN.load_state_dict(st.DNN, strict=True)
    G.load_state_dict(st.G, strict=True)
    assert list(D.state_dict().keys()) == list(st.D.keys()), 'oracle key order differs from the reference state_dict'
    exp = ref_harness.make_experiment(CrowdDgganExperiment if dggan else CrowdExperiment, settings_for(cfg), D=D, G=G, DNN=DNN)
    x, y, u, z, alpha, z2 = O.synthetic_crowd_batch(2, cfg['input_seed'])
    exp.dnn_training_step(x, y, 0)
    with ref_harness.injected_noise(z, alpha, z2):
        exp.gan_training_step(x, y, u, 0)
    sc = ref_harness.last_scalars(exp)
    data = {}
    for tag, key in SCALAR_TAGS.items():
        data[f'step0/scalars/{key}'] = np.float64(sc[tag])
    data['step0/gradient_norm'] = exp.gradient_norm.detach().numpy().copy()
    if not dggan:                                  # KnnDenseNetCatDggan does not publish .features
        data['step0/labeled_features'] = exp.labeled_features.detach().reshape(2, -1).numpy().copy()
    for net, mod, init in (('D', exp.D, st.D), ('G', exp.G, st.G), ('DNN', exp.DNN, st.DNN)):
        keys, sums, abss = [], [], []
        for k, v in mod.state_dict().items():
            if O.is_buffer_key(k):
                assert torch.equal(v, init[k]), f'{k}: BatchNorm statistics changed (srgan.py:538-542 says they must not)'
                continue
            d = (v.detach() - init[k]).double()
            keys.append(k); sums.append(d.sum().item()); abss.append(d.abs().sum().item())
        data[f'update/{net}/keys'] = np.frombuffer(json.dumps(keys).encode(), dtype=np.uint8)
        data[f'update/{net}/sum'] = np.array(sums)
        data[f'update/{net}/abs_sum'] = np.array(abss)
    cfg_json = dict(cfg, steps=1, torch=torch.__version__)
    data['config_json'] = np.frombuffer(json.dumps(cfg_json).encode(), dtype=np.uint8)
    path = os.path.join(OUT, f'crowd_{method}.npz')
    np.savez_compressed(path, **data)
    print(f'crowd_{method}', {k: float(v) for k, v in data.items() if '/scalars/' in k}, os.path.getsize(path), 'bytes')


def sgan_cases():
    """SGAN method (sgan.py:10-67): AgeSganExperiment with a reduced DCGAN pair
    (Discriminator(number_of_outputs=10), age/sgan.py:16-20) and CoefficientSganExperiment with its own SganMLP (hidden 100,
    coefficient/models.py:75-93).  `bins` (a plain list in the config) is the experiment's own linspace."""
    torch.set_num_threads(1)
    from age.models import Generator, Discriminator
    from age.sgan import AgeSganExperiment
    from coefficient.sgan import CoefficientSganExperiment
    cfg = dict(method='sgan', family='dcgan', batch_size=4, learning_rate=1e-4, weight_decay=0.0, labeled_loss_multiplier=1.0,
               matching_loss_multiplier=1.0, gradient_penalty_multiplier=1e2, image_size=32, conv_dim=8, z_dim=16, number_of_bins=10)
    D = Discriminator(image_size=32, conv_dim=8, number_of_outputs=10)
    DNN = Discriminator(image_size=32, conv_dim=8, number_of_outputs=10)
    G = Generator(z_dim=16, image_size=32, conv_dim=8)
    with torch.no_grad():
        for k, p in D.named_parameters():
            if k.endswith('weight'):
                p.mul_(3.0)
    exp = ref_harness.make_experiment(AgeSganExperiment, settings_for(cfg), D=D, G=G, DNN=DNN)
    cfg['bins'] = [float(v) for v in exp.bins]
    gen = torch.Generator().manual_seed(12)
    batches = []
    for _ in range(3):
        x = torch.rand(4, 3, 32, 32, generator=gen) * 2 - 1
        u = torch.rand(4, 3, 32, 32, generator=gen) * 2 - 1
        y = torch.rand(4, generator=gen) * 85 + 10
        z = torch.randn(4, 16, generator=gen)
        alpha = torch.rand(4, 1, 1, 1, generator=gen)
        z2 = torch.randn(4, 16, generator=gen)
        batches.append((x, y, u, z, alpha, z2))
    run_case('dcgan_sgan_mini', exp, cfg, batches, steps=3)

    cfg = dict(method='sgan', family='coefficient', batch_size=64, learning_rate=1e-3, weight_decay=1e-3,
               labeled_loss_multiplier=1.0, matching_loss_multiplier=1.0, gradient_penalty_multiplier=1e3, number_of_bins=10)
    exp = ref_harness.make_experiment(CoefficientSganExperiment, settings_for(cfg))
    with torch.no_grad():
        for k, p in exp.D.named_parameters():
            if k.endswith('weight'):
                p.mul_(2.0)
    cfg['bins'] = [float(v) for v in exp.bins]
    run_case('coefficient_sgan', exp, cfg, coefficient_batches(64, 3, seed=10), steps=3)


def main():
    ref_harness.install_shims()
    os.makedirs(OUT, exist_ok=True)
    if 'sgan' in sys.argv[1:]:
        sgan_cases()
        return
    if 'crowd' in sys.argv[1:]:
        crowd_case()
        return
    if 'crowd_dggan' in sys.argv[1:]:
        crowd_case('dggan', d_scale=float(os.environ.get('D_SCALE', '1.6')))
        return
    torch.set_num_threads(1)                     # deterministic reduction order for the fixtures
    from coefficient.srgan import CoefficientExperiment
    from coefficient.dggan import CoefficientDgganExperiment
    from age.models import Generator, Discriminator

    # ---- A: coefficient SR-GAN, D weights x4 so that the gradient penalty hinge is active
    cfg = dict(method='srgan', family='coefficient', batch_size=64, learning_rate=1e-3, weight_decay=1e-3,
               matching_loss_multiplier=1.0, contrasting_loss_multiplier=1.0, gradient_penalty_multiplier=10.0,
               hidden_size=10)
    exp = ref_harness.make_experiment(CoefficientExperiment, settings_for(cfg))
    with torch.no_grad():
        for k, p in exp.D.named_parameters():
            if k.endswith('weight') and 'linear4' not in k:
                p.mul_(4.0)
    run_case('coefficient_srgan', exp, cfg, coefficient_batches(64, 3, seed=7), steps=3)

    # ---- A2: same, alternate distance functions (utility.py:211-243)
    cfg2 = dict(cfg, weight_decay=0.0, matching_distance_function='square_mean',
                contrasting_distance_function='abs_plus_one_log_mean_neg')
    exp = ref_harness.make_experiment(CoefficientExperiment, settings_for(cfg2))
    run_case('coefficient_srgan_altdist', exp, cfg2, coefficient_batches(64, 2, seed=8), steps=2)

    # ---- B: coefficient DG-GAN (coefficient/dggan.py)
    cfg = dict(method='dggan', family='coefficient', batch_size=64, learning_rate=1e-3, weight_decay=0.0,
               matching_loss_multiplier=1.0, contrasting_loss_multiplier=1.0, gradient_penalty_multiplier=10.0,
               dggan_loss_multiplier=10.0, hidden_size=10)
    exp = ref_harness.make_experiment(CoefficientDgganExperiment, settings_for(cfg))
    with torch.no_grad():
        for k, p in exp.D.named_parameters():
            if k.endswith('weight'):
                p.mul_(4.0)
    run_case('coefficient_dggan', exp, cfg, coefficient_batches(64, 3, seed=9), steps=3)

    # ---- C: DCGAN (age/driving models at reduced size), multipliers of run.py:30-35, D weights x3 (GP active)
    cfg = dict(method='srgan', family='dcgan', batch_size=4, learning_rate=1e-4, weight_decay=0.0,
               matching_loss_multiplier=1e2, contrasting_loss_multiplier=1e1, gradient_penalty_multiplier=1e2,
               image_size=32, conv_dim=8, z_dim=16)
    from age.srgan import AgeExperiment
    D = Discriminator(image_size=32, conv_dim=8)
    DNN = Discriminator(image_size=32, conv_dim=8)
    G = Generator(z_dim=16, image_size=32, conv_dim=8)
    with torch.no_grad():
        for k, p in D.named_parameters():
            if k.endswith('weight') and 'layer5' not in k:
                p.mul_(3.0)
    exp = ref_harness.make_experiment(AgeExperiment, settings_for(cfg), D=D, G=G, DNN=DNN)
    gen = torch.Generator().manual_seed(11)
    batches = []
    for _ in range(3):
        x = torch.rand(4, 3, 32, 32, generator=gen) * 2 - 1
        u = torch.rand(4, 3, 32, 32, generator=gen) * 2 - 1
        y = torch.rand(4, generator=gen) * 85 + 10
        z = torch.randn(4, 16, generator=gen)
        alpha = torch.rand(4, 1, 1, 1, generator=gen)
        z2 = torch.randn(4, 16, generator=gen)
        batches.append((x, y, u, z, alpha, z2))
    run_case('dcgan_mini', exp, cfg, batches, steps=3)
    crowd_case()
    crowd_case('dggan', d_scale=1.6)


if __name__ == '__main__':
    main()
