"""Helpers shared by the -m gpu tests, smoke() and bench.py: build the product StepRunner from an oracle state."""
import torch

import srgan_b200
from oracle import srgan_oracle as O


def settings_from_cfg(cfg: O.StepConfig, precision='fp32'):
    s = srgan_b200.Settings()
    for k in ('batch_size', 'learning_rate', 'weight_decay', 'labeled_loss_multiplier', 'matching_loss_multiplier',
              'contrasting_loss_multiplier', 'srgan_loss_multiplier', 'dggan_loss_multiplier',
              'gradient_penalty_multiplier', 'labeled_loss_order', 'generator_training_step_period', 'map_multiplier'):
        setattr(s, k, getattr(cfg, k))
    s.matching_distance_function = getattr(srgan_b200, cfg.matching_distance_function)
    s.contrasting_distance_function = getattr(srgan_b200, cfg.contrasting_distance_function)
    s.precision = precision
    s.bins = tuple(cfg.bins)
    return s


def modules_from_state(st: O.OracleState, **dcgan_kwargs):
    if st.d_spec.family == 'coefficient':
        n_out, hidden = st.D['linear4.weight'].shape                     # 1 | 2 (DgganMLP) | number_of_bins (SganMLP, hidden 100)
        D, DNN = srgan_b200.CoefficientMLP(hidden, n_out), srgan_b200.CoefficientMLP(hidden, n_out)
        G = srgan_b200.CoefficientGenerator(st.G['linear1.weight'].shape[0])
    elif st.d_spec.family == 'crowd':
        sp = st.d_spec
        z_dim, c8, k, _ = st.G['fc.0.weight'].shape
        kw = dict(growth_rate=sp.growth_rate, block_config=sp.block_config, num_init_features=sp.num_init_features,
                  bn_size=sp.bn_size, label_patch_size=sp.label_patch_size, image_size=k * 16)
        kw['number_of_outputs'] = 2 if sp.dggan else 1
        D, DNN = srgan_b200.KnnDenseNetCat(**kw), srgan_b200.KnnDenseNetCat(**kw)
        G = srgan_b200.DcganGenerator(z_dim, k * 16, c8 // 8)
    else:
        z_dim, c8, k, _ = st.G['fc.0.weight'].shape
        n_out = st.D['layer5.0.weight'].shape[0]                          # number_of_bins for the SGAN discriminator
        D = srgan_b200.DcganDiscriminator(k * 16, c8 // 8, n_out)
        DNN = srgan_b200.DcganDiscriminator(k * 16, c8 // 8, n_out)
        G = srgan_b200.DcganGenerator(z_dim, k * 16, c8 // 8)
    D.load_state_dict({k: v.float() for k, v in st.D.items()})
    G.load_state_dict({k: v.float() for k, v in st.G.items()})
    DNN.load_state_dict({k: v.float() for k, v in st.DNN.items()})
    return D.cuda(), G.cuda(), DNN.cuda()


def runner_from_state(st: O.OracleState, cfg: O.StepConfig, precision='fp32', comm=None, micro_batch=0):
    D, G, DNN = modules_from_state(st)
    s = settings_from_cfg(cfg, precision)
    s.micro_batch = micro_batch
    return srgan_b200.StepRunner(D, G, DNN, s, cfg.method, precision=precision, comm=comm)


def to_cuda(*ts):
    return tuple(tuple(e.cuda() for e in t) if isinstance(t, (tuple, list)) else t.cuda() for t in ts)


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
