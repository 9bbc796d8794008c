"""
CPU oracle for the SR-GAN training step  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
module. The product path (sr-gan_b200/) never does: it fails loudly when the CUDA library is missing.

This is a plain PyTorch-on-CPU *restatement* (functional style, fp32 or fp64, autograd incl. double-backward)
of the reference algorithm.  Every function cites the reference file:line it follows.  It is pinned against the
unmodified reference, imported from /root/reference in the build container, by oracle/make_golden.py; the vectors
that script produced are committed under tests/golden/ and tests/test_oracle_golden.py re-checks them on every run
(the reference itself has no tests / golden vectors: SURVEY.md section 8c).

State layout: a model is a `dict[str, Tensor]` whose keys and shapes are exactly the reference module's
`state_dict()` (so fixtures and checkpoints interchange).  Noise (z, alpha, z2) is always passed in explicitly.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------------------------
# Distance functions: utility.py:201-243 (the two Settings defaults are abs_mean :226 and
# abs_plus_one_sqrt_mean_neg :216; the others are the alternates run.py:17-18 imports).
# ----------------------------------------------------------------------------------------------------------------
DISTANCES = {
    'abs_mean': lambda t: t.abs().mean(),                                    # utility.py:226-228
    'abs_mean_neg': lambda t: t.abs().mean().neg(),                          # utility.py:221-223
    'abs_plus_one_sqrt_mean_neg': lambda t: t.abs().add(1).sqrt().mean().neg(),   # utility.py:216-218
    'abs_plus_one_log_mean_neg': lambda t: t.abs().add(1).log().mean().neg(),     # utility.py:211-213
    'square_mean': lambda t: t.pow(2).mean(),                                # utility.py:241-243
    'norm_mean': lambda t: t.pow(2).sum().pow(0.5),                          # utility.py:236-238
}


@dataclass
class StepConfig:
    """The subset of settings.py:12-67 the step reads, with the same defaults."""
    method: str = 'srgan'                    # 'srgan' | 'dggan' | 'sgan'   (settings.py:123-127; sgan.py)
    batch_size: int = 1000                   # srgan.py:363 uses settings.batch_size for the alpha shape
    learning_rate: float = 1e-4
    weight_decay: float = 0.0
    labeled_loss_multiplier: float = 1.0
    matching_loss_multiplier: float = 1.0
    contrasting_loss_multiplier: float = 1.0
    srgan_loss_multiplier: float = 1.0
    dggan_loss_multiplier: float = 10.0
    gradient_penalty_multiplier: float = 10.0
    labeled_loss_order: int = 2
    generator_training_step_period: int = 1
    matching_distance_function: str = 'abs_mean'
    contrasting_distance_function: str = 'abs_plus_one_sqrt_mean_neg'
    map_multiplier: float = 1e-6             # crowd only
    betas: Tuple[float, float] = (0.9, 0.999)   # torch.optim.Adam defaults, srgan.py:136-138
    eps: float = 1e-8
    bins: Tuple[float, ...] = ()             # sgan: the bin centres (age/sgan.py:14 linspace(10, 95, number_of_bins))


# ----------------------------------------------------------------------------------------------------------------
# Model families (forward only; gradients come from autograd).
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class ModelSpec:
    """Which network family a Params dict belongs to, plus the constants its forward needs."""
    family: str                    # 'coefficient' | 'dcgan' | 'crowd'
    dggan: bool = False            # D has a second (fake-score) output: coefficient/models.py:53-72
    leaky: float = 0.01            # coefficient: F.leaky_relu default 0.01; dcgan: 0.05 (age/models.py:46-50,70-73)
    # crowd KnnDenseNetCat constructor arguments (crowd/models.py:1060-1062; defaults = DenseNet-201 at 224)
    block_config: Tuple[int, ...] = (6, 12, 48, 32)
    growth_rate: int = 32
    num_init_features: int = 64
    bn_size: int = 4
    label_patch_size: int = 224


def coefficient_d_forward(p: Params, x: torch.Tensor, dggan: bool = False):
    """coefficient/models.py:43-50 (MLP) and :65-72 (DgganMLP): 3x(Linear+leaky 0.01) -> features -> Linear."""
    h = F.leaky_relu(F.linear(x, p['linear1.weight'], p['linear1.bias']), 0.01)
    h = F.leaky_relu(F.linear(h, p['linear2.weight'], p['linear2.bias']), 0.01)
    h = F.leaky_relu(F.linear(h, p['linear3.weight'], p['linear3.bias']), 0.01)
    out = F.linear(h, p['linear4.weight'], p['linear4.bias'])
    if dggan:
        return (out[:, 0].squeeze(), out[:, 1].squeeze()), h
    return out.squeeze(), h             # SganMLP (coefficient/models.py:75-93): [B, number_of_bins] logits


def coefficient_g_forward(p: Params, z: torch.Tensor):
    """coefficient/models.py:22-28: 3x(Linear+leaky 0.01) -> Linear (no activation)."""
    h = F.leaky_relu(F.linear(z, p['linear1.weight'], p['linear1.bias']), 0.01)
    h = F.leaky_relu(F.linear(h, p['linear2.weight'], p['linear2.bias']), 0.01)
    h = F.leaky_relu(F.linear(h, p['linear3.weight'], p['linear3.bias']), 0.01)
    return F.linear(h, p['linear4.weight'], p['linear4.bias'])


def dcgan_d_forward(p: Params, x: torch.Tensor):
    """age/models.py:68-80 (== driving/models.py): 4x(Conv k4 s2 p1 + leaky 0.05) -> features=flatten -> Conv k=H/16."""
    h = x
    for i in (1, 2, 3, 4):
        h = F.leaky_relu(F.conv2d(h, p[f'layer{i}.0.weight'], p[f'layer{i}.0.bias'], stride=2, padding=1), 0.05)
    features = h.reshape(h.size(0), -1)
    out = F.conv2d(h, p['layer5.0.weight'], p['layer5.0.bias'], stride=1, padding=0)
    if out.size(1) > 1:                 # number_of_outputs = number_of_bins (age/sgan.py:18-19, age/models.py:76-79)
        return out.reshape(-1, out.size(1)), features
    return out.reshape(-1), features


def dcgan_g_forward(p: Params, z: torch.Tensor):
    """age/models.py:44-52 and crowd/models.py:139-147: view(B,z,1,1) -> ConvT k=H/16 (no act) -> 3x(ConvT k4 s2 p1 +
    leaky 0.05) -> ConvT + tanh."""
    h = z.reshape(z.size(0), z.size(1), 1, 1)
    h = F.conv_transpose2d(h, p['fc.0.weight'], p['fc.0.bias'], stride=1, padding=0)
    for i in (1, 2, 3):
        h = F.leaky_relu(F.conv_transpose2d(h, p[f'layer{i}.0.weight'], p[f'layer{i}.0.bias'], stride=2, padding=1), 0.05)
    return torch.tanh(F.conv_transpose2d(h, p['layer4.0.weight'], p['layer4.0.bias'], stride=2, padding=1))


def _bn_eval(p: Params, prefix: str, x: torch.Tensor):
    """nn.BatchNorm2d in eval() mode -- what disable_batch_norm_updates (srgan.py:538-542, applied :261,276) leaves on the
    hot path: a per-channel affine of the running statistics (eps 1e-5), weight and bias still trainable."""
    return F.batch_norm(x, p[prefix + '.running_mean'], p[prefix + '.running_var'], p[prefix + '.weight'],
                        p[prefix + '.bias'], training=False, eps=1e-5)


def crowd_map_module(p: Params, prefix: str, x: torch.Tensor):
    """MapModule.forward, crowd/models.py:778-786: ConvT (k = stride = label/input) + leaky -> map; 3 x (Conv k2 s2 +
    leaky), Conv (full extent) + leaky -> 20 features; Conv 1x1 -> count."""
    w = p[prefix + '.map_transposed_conv_layer.weight']
    map_ = F.leaky_relu(F.conv_transpose2d(x, w, p[prefix + '.map_transposed_conv_layer.bias'], stride=w.shape[-1]), 0.01)
    out = map_
    for name in ('conv1', 'conv2', 'conv3'):
        out = F.leaky_relu(F.conv2d(out, p[f'{prefix}.{name}.weight'], p[f'{prefix}.{name}.bias'], stride=2), 0.01)
    out = F.leaky_relu(F.conv2d(out, p[prefix + '.linear1.weight'], p[prefix + '.linear1.bias']), 0.01)
    count = F.conv2d(out, p[prefix + '.count_layer.weight'], p[prefix + '.count_layer.bias'])
    return map_, count, out


def crowd_d_forward(spec: ModelSpec, p: Params, x: torch.Tensor):
    """KnnDenseNetCat.forward, crowd/models.py:1136-1166 (+ _DenseLayer :335-353, _Transition :364-371):
    DenseNet trunk (BN in eval mode) with taps after every transition feeding three MapModules, plus the count head.
    Returns ((count [B], map [B,3,L,L]), features [B,80])."""
    B = x.shape[0]
    out = F.conv2d(x, p['conv_layer1.conv0.weight'], None, stride=2, padding=3)
    out = F.relu(_bn_eval(p, 'conv_layer1.norm0', out))
    out = F.max_pool2d(out, kernel_size=3, stride=2, padding=1)
    taps = []
    for bi, n_layers in enumerate(spec.block_config, 1):
        for li in range(1, n_layers + 1):
            pre = f'dense_blocks.denseblock{bi}.denselayer{li}'
            h = F.relu(_bn_eval(p, pre + '.norm1', out))
            h = F.conv2d(h, p[pre + '.conv1.weight'])
            h = F.relu(_bn_eval(p, pre + '.norm2', h))
            h = F.conv2d(h, p[pre + '.conv2.weight'], padding=1)
            out = torch.cat([out, h], 1)
        if bi != len(spec.block_config):
            pre = f'transition_layers.transition{bi}'
            h = F.relu(_bn_eval(p, pre + '.norm', out))
            h = F.conv2d(h, p[pre + '.conv.weight'])
            out = F.avg_pool2d(h, kernel_size=2, stride=2)
            taps.append(out)
    out = F.relu(_bn_eval(p, 'norm5', out))
    final_pool = F.avg_pool2d(out, kernel_size=out.shape[-1], stride=1)        # kernel 7 at 224 (:1151)
    fcf = F.leaky_relu(F.conv2d(final_pool, p['final_count_feature_layer.weight'], p['final_count_feature_layer.bias']), 0.01)
    final_count = F.conv2d(fcf, p['count_layer.weight'], p['count_layer.bias'])
    maps, counts, hs = [], [], []
    for i, t in enumerate(taps[:3], 1):
        m, c, h = crowd_map_module(p, f'map_module{i}', t)
        maps.append(m), counts.append(c), hs.append(h)
    features = torch.cat([h.reshape(B, -1) for h in hs] + [fcf.reshape(B, -1)], dim=1)
    count = counts[0] + counts[1] + counts[2] + final_count
    L = spec.label_patch_size
    map_ = torch.cat(maps, dim=1).reshape(B, 3, L, L)
    if spec.dggan:
        # KnnDenseNetCatDggan.forward, crowd/models.py:1024-1046: every count layer has two outputs, the second is
        # `real_label` (the DG-GAN score); that module does not publish `.features` (the DG-GAN losses do not use them)
        count = count.reshape(B, 2)
        return (count[:, 0], map_), count[:, 1], features
    return (count.reshape(B), map_), None, features


def d_forward(spec: ModelSpec, p: Params, x: torch.Tensor):
    """Returns (prediction, fake_score_or_None, features)."""
    if spec.family == 'crowd':
        return crowd_d_forward(spec, p, x)
    if spec.family == 'coefficient':
        out, f = coefficient_d_forward(p, x, spec.dggan)
        if spec.dggan:
            return out[0], out[1], f
        return out, None, f
    if spec.family == 'dcgan':
        out, f = dcgan_d_forward(p, x)
        return out, None, f
    raise ValueError(spec.family)


def g_forward(spec: ModelSpec, p: Params, z: torch.Tensor):
    if spec.family == 'coefficient':
        return coefficient_g_forward(p, z)
    if spec.family == 'dcgan':
        return dcgan_g_forward(p, z)
    raise ValueError(spec.family)


# ----------------------------------------------------------------------------------------------------------------
# Losses.
# ----------------------------------------------------------------------------------------------------------------
def labeled_loss_function(predicted, labels, order=2):
    """srgan.py:414-417."""
    return (predicted - labels).abs().pow(order).mean()


def crowd_labeled_loss_function(predicted_count, predicted_maps, head_labels, map_labels, order, map_multiplier):
    """crowd/srgan.py:247-254: count loss + map_multiplier * map loss (predicted_maps [B,3,H,W], map_labels [B,H,W])."""
    maps = map_labels.unsqueeze(1)
    map_loss = (predicted_maps - maps).abs().mean(1).sum(1).sum(1).pow(order).mean()
    count_loss = (predicted_count - head_labels.sum(1).sum(1)).abs().pow(order).mean()
    return count_loss + map_loss * map_multiplier


def labeled_loss(spec: ModelSpec, cfg: 'StepConfig', pred, labels):
    """labeled_loss_function as the application overrides it: srgan.py:414-417, crowd/srgan.py:247-254
    (crowd: pred = (count, maps), labels = (density, map): the tuple built at srgan.py:112-113)."""
    if spec.family == 'crowd':
        count, maps = pred
        density, map_labels = labels
        return crowd_labeled_loss_function(count, maps, density, map_labels, cfg.labeled_loss_order, cfg.map_multiplier)
    return labeled_loss_function(pred, labels, cfg.labeled_loss_order)


def feature_distance_loss(base_features, other_features, distance: str):
    """srgan.py:438-449 with normalize_feature_norm=False (the True branch is a known bug)."""
    return DISTANCES[distance](base_features.mean(0) - other_features.mean(0))


def bce_with_logits(scores, target_value: float):
    """torch.nn.BCEWithLogitsLoss (mean reduction) against a constant target: coefficient/dggan.py:39-40,48-49,62-63."""
    return F.binary_cross_entropy_with_logits(scores, torch.full_like(scores, target_value))


def real_numbers_to_bin_indexes(real_numbers, bins):
    """utility.py:141-144."""
    return (real_numbers.reshape(-1, 1) - bins.reshape(1, -1)).abs().min(dim=1)[1]


def sgan_labeled_loss(cfg: 'StepConfig', logits, labels):
    """sgan.py:20-31: cross entropy against the label's bin."""
    bins = torch.tensor(cfg.bins, dtype=logits.dtype)
    return F.cross_entropy(logits, real_numbers_to_bin_indexes(labels, bins)) * cfg.labeled_loss_multiplier


def sgan_binary_loss(logits, target_value: float):
    """sgan.py:33-67: BCE-with-logits on logsumexp over the class logits (utility.py:161-185)."""
    return bce_with_logits(torch.logsumexp(logits, dim=1), target_value)


# ----------------------------------------------------------------------------------------------------------------
# Adam exactly as torch.optim.Adam runs it (srgan.py:131-138).
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class AdamState:
    step: int = 0
    exp_avg: Dict[str, torch.Tensor] = field(default_factory=dict)
    exp_avg_sq: Dict[str, torch.Tensor] = field(default_factory=dict)


def adam_update(p: Params, grads: Dict[str, Optional[torch.Tensor]], st: AdamState, lr, weight_decay, betas, eps):
    """L2 (coupled) weight decay, bias-corrected moments; parameters whose grad is None are skipped like torch does."""
    b1, b2 = betas
    st.step += 1
    t = st.step
    for k in list(p):
        g = grads.get(k)
        if g is None:
            continue
        if weight_decay != 0:
            g = g + weight_decay * p[k]
        if k not in st.exp_avg:
            st.exp_avg[k] = torch.zeros_like(p[k])
            st.exp_avg_sq[k] = torch.zeros_like(p[k])
        st.exp_avg[k] = b1 * st.exp_avg[k] + (1 - b1) * g
        st.exp_avg_sq[k] = b2 * st.exp_avg_sq[k] + (1 - b2) * g * g
        step_size = lr / (1 - b1 ** t)
        denom = st.exp_avg_sq[k].sqrt() / math.sqrt(1 - b2 ** t) + eps
        p[k] = p[k] - step_size * st.exp_avg[k] / denom


# ----------------------------------------------------------------------------------------------------------------
# The step.
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class OracleState:
    """Everything Experiment owns that the step mutates: three networks + three Adam states (srgan.py:37-42)."""
    d_spec: ModelSpec
    g_spec: ModelSpec
    D: Params
    G: Params
    DNN: Params
    d_adam: AdamState = field(default_factory=AdamState)
    g_adam: AdamState = field(default_factory=AdamState)
    dnn_adam: AdamState = field(default_factory=AdamState)

    def clone(self):
        import copy
        return copy.deepcopy(self)


def is_buffer_key(k: str) -> bool:
    """state_dict entries that are module buffers, not parameters (BatchNorm statistics)."""
    return k.endswith('running_mean') or k.endswith('running_var') or k.endswith('num_batches_tracked')


def _leaf(p: Params) -> Params:
    return {k: (v.detach().clone() if is_buffer_key(k) else v.detach().clone().requires_grad_(True)) for k, v in p.items()}


def _grads(loss, leaf: Params):
    keys = [k for k, v in leaf.items() if v.requires_grad]
    gs = torch.autograd.grad(loss, [leaf[k] for k in keys], allow_unused=True)
    return dict(zip(keys, gs))


def dnn_lr(cfg: StepConfig, step: int) -> float:
    """srgan.py:432-436: only the DNN optimizer is decayed (x0.1 every 100k steps)."""
    return cfg.learning_rate * (0.1 ** (step // 100000))


def dnn_training_step(st: OracleState, cfg: StepConfig, x, y, step: int = 0):
    """srgan.py:259-271 + dnn_loss_calculation :322-327 (DG-GAN: coefficient/dggan.py:22-27)."""
    leaf = _leaf(st.DNN)
    pred, _, _ = d_forward(st.d_spec, leaf, x)
    if cfg.method == 'sgan':
        loss = sgan_labeled_loss(cfg, pred, y)
    else:
        loss = labeled_loss(st.d_spec, cfg, pred, y) * cfg.labeled_loss_multiplier
    g = _grads(loss, leaf)
    adam_update(st.DNN, g, st.dnn_adam, dnn_lr(cfg, step), cfg.weight_decay, cfg.betas, cfg.eps)
    return {'dnn_loss': float(loss.detach())}


def gradient_penalty(st_spec: ModelSpec, leafD: Params, cfg: StepConfig, fake, u, alpha):
    """srgan.py:360-375 + interpolate_loss_calculation :377-381 (DG-GAN target = raw fake score,
    coefficient/dggan.py:54-57)."""
    interp = (alpha * u.detach() + (1 - alpha) * fake.detach()).requires_grad_(True)
    logits, score, feats = d_forward(st_spec, leafD, interp)
    if cfg.method == 'dggan':
        target = score
    elif cfg.method == 'sgan':              # sgan.py:51-58: a scalar, already times the penalty multiplier
        target = sgan_binary_loss(logits, 0.0) * cfg.gradient_penalty_multiplier
    else:
        target = feats.norm(dim=1)
    grads = torch.autograd.grad(target, interp, torch.ones_like(target), create_graph=True)[0]
    gnorm = grads.reshape(u.size(0), -1).norm(dim=1)
    excess = torch.clamp(gnorm - 1, min=0)
    return (excess ** 2).mean() * cfg.gradient_penalty_multiplier, gnorm, feats


def gan_training_step(st: OracleState, cfg: StepConfig, x, y, u, z, alpha, z2, step: int = 0):
    """srgan.py:273-320.  Gradients of the four discriminator losses accumulate (four .backward() calls before one
    d_optimizer.step()), here as one backward of their sum.  z, alpha, z2 are the three noise draws
    (:286-289, :364, :301) supplied by the caller."""
    out = {}
    leafD = _leaf(st.D)
    spec = st.d_spec
    # -- labeled  (:279, :329-335 | dggan.py:29-34)
    pred, _, f_x = d_forward(spec, leafD, x)
    if cfg.method == 'sgan':
        labeled = sgan_labeled_loss(cfg, pred, y)
    else:
        labeled = labeled_loss(spec, cfg, pred, y) * cfg.labeled_loss_multiplier
    # -- unlabeled (:283, :337-346 | dggan.py:36-43 | sgan.py:33-40)
    logits_u, score_u, f_u = d_forward(spec, leafD, u)
    with torch.no_grad():
        fake = g_forward(st.g_spec, st.G, z)                    # :290  (graph unused: fake is detached / G grads zeroed)
    logits_f, score_f, f_f = d_forward(spec, leafD, fake)
    if cfg.method == 'sgan':                # sgan.py:33-49: both terms use matching_loss_multiplier
        unlabeled = sgan_binary_loss(logits_u, 1.0) * cfg.matching_loss_multiplier
        fake_loss = sgan_binary_loss(logits_f, 0.0) * cfg.matching_loss_multiplier
    elif cfg.method == 'dggan':
        unlabeled = bce_with_logits(score_u, 0.0) * cfg.matching_loss_multiplier * cfg.dggan_loss_multiplier
        fake_loss = bce_with_logits(score_f, 1.0) * cfg.contrasting_loss_multiplier * cfg.dggan_loss_multiplier
    else:
        unlabeled = (feature_distance_loss(f_u, f_x, cfg.matching_distance_function)
                     * cfg.matching_loss_multiplier * cfg.srgan_loss_multiplier)
        # -- fake (:291, :348-358)
        fake_loss = (feature_distance_loss(f_u, f_f, cfg.contrasting_distance_function)
                     * cfg.contrasting_loss_multiplier * cfg.srgan_loss_multiplier)
    # -- gradient penalty (:294, :360-375)
    gp, gnorm, f_i = gradient_penalty(spec, leafD, cfg, fake, u, alpha)
    total = labeled + unlabeled + fake_loss + gp
    gD = _grads(total, leafD)
    adam_update(st.D, gD, st.d_adam, cfg.learning_rate, cfg.weight_decay, cfg.betas, cfg.eps)      # :297
    out.update(labeled_loss=float(labeled.detach()), unlabeled_loss=float(unlabeled.detach()), fake_loss=float(fake_loss.detach()),
               gradient_penalty=float(gp.detach()), gradient_norm_mean=float(gnorm.detach().mean()))
    out['features'] = {'labeled': f_x.detach(), 'unlabeled': f_u.detach(), 'fake': f_f.detach(),
                       'interpolates': f_i.detach()}
    out['gradient_norm'] = gnorm.detach()
    out['d_grads'] = {k: (None if v is None else v.detach()) for k, v in gD.items()}
    # -- generator (:299-305, :383-391 | dggan.py:59-64); D is already updated.
    if step % cfg.generator_training_step_period == 0:
        leafG = _leaf(st.G)
        fake2 = g_forward(st.g_spec, leafG, z2)
        logits_f2, score_f2, f_f2 = d_forward(spec, st.D, fake2)
        if cfg.method == 'sgan':            # sgan.py:60-67
            g_loss = -sgan_binary_loss(logits_f2, 0.0)
        elif cfg.method == 'dggan':
            g_loss = bce_with_logits(score_f2, 0.0)
        else:
            with torch.no_grad():
                _, _, f_u2 = d_forward(spec, st.D, u)
            g_loss = feature_distance_loss(f_u2, f_f2, cfg.matching_distance_function) * cfg.matching_loss_multiplier
        gG = _grads(g_loss, leafG)
        adam_update(st.G, gG, st.g_adam, cfg.learning_rate, 0.0, cfg.betas, cfg.eps)               # :137 no wd on G
        out['generator_loss'] = float(g_loss.detach())
        out['g_grads'] = {k: (None if v is None else v.detach()) for k, v in gG.items()}
    return out


def training_step(st: OracleState, cfg: StepConfig, x, y, u, z, alpha, z2, step: int = 0):
    """One iteration of training_loop (srgan.py:105-118): dnn_training_step then gan_training_step."""
    out = dnn_training_step(st, cfg, x, y, step)
    out.update(gan_training_step(st, cfg, x, y, u, z, alpha, z2, step))
    return out


# ----------------------------------------------------------------------------------------------------------------
# Deterministic initial states for seeded tests / benches (NOT the reference initialiser: the reference uses
# nn.Module default init under seed_all(0); golden fixtures carry the reference's actual initial state).
# ----------------------------------------------------------------------------------------------------------------
def _uniform(gen, shape, bound, dtype):
    return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * bound).to(dtype)


def init_coefficient(seed=0, hidden=10, dggan=False, dtype=torch.float32, n_out=None) -> OracleState:
    """Shapes of coefficient/models.py:12-72 (input 50 = observation_count 10 x irrelevant_data_multiplier 5)."""
    gen = torch.Generator().manual_seed(seed)

    def mlp(sizes):
        p = {}
        for i, (a, b) in enumerate(zip(sizes[:-1], sizes[1:]), 1):
            bound = 1 / math.sqrt(a)
            p[f'linear{i}.weight'] = _uniform(gen, (b, a), bound, dtype)
            p[f'linear{i}.bias'] = _uniform(gen, (b,), bound, dtype)
        return p
    d = mlp([50, hidden, hidden, hidden, n_out or (2 if dggan else 1)])     # n_out = number_of_bins: SganMLP (hidden 100)
    dnn = {k: v.clone() for k, v in d.items()}               # D and DNN start identical
    g = mlp([10, hidden, hidden, hidden, 50])
    return OracleState(ModelSpec('coefficient', dggan=dggan), ModelSpec('coefficient'), d, g, dnn)


def init_dcgan(seed=0, image_size=128, conv_dim=64, z_dim=256, dtype=torch.float32, scale=1.0, n_out=1) -> OracleState:
    """Shapes of age/models.py:32-80 (crowd DCGenerator: image_size=224, crowd/models.py:127-147)."""
    gen = torch.Generator().manual_seed(seed)
    k = image_size // 16
    d, g = {}, {}
    chans = [3, conv_dim, conv_dim * 2, conv_dim * 4, conv_dim * 8]
    for i in range(1, 5):
        bound = 1 / math.sqrt(chans[i - 1] * 16)
        d[f'layer{i}.0.weight'] = _uniform(gen, (chans[i], chans[i - 1], 4, 4), bound, dtype) * scale
        d[f'layer{i}.0.bias'] = _uniform(gen, (chans[i],), bound, dtype)
    bound = 1 / math.sqrt(chans[4] * k * k)
    d['layer5.0.weight'] = _uniform(gen, (n_out, chans[4], k, k), bound, dtype)
    d['layer5.0.bias'] = _uniform(gen, (n_out,), bound, dtype)
    dnn = {kk: v.clone() for kk, v in d.items()}
    bound = 1 / math.sqrt(conv_dim * 8 * k * k)
    g['fc.0.weight'] = _uniform(gen, (z_dim, conv_dim * 8, k, k), bound, dtype)
    g['fc.0.bias'] = _uniform(gen, (conv_dim * 8,), bound, dtype)
    gch = [conv_dim * 8, conv_dim * 4, conv_dim * 2, conv_dim, 3]
    for i in range(1, 5):
        bound = 1 / math.sqrt(gch[i] * 16)
        g[f'layer{i}.0.weight'] = _uniform(gen, (gch[i - 1], gch[i], 4, 4), bound, dtype)
        g[f'layer{i}.0.bias'] = _uniform(gen, (gch[i],), bound, dtype)
    return OracleState(ModelSpec('dcgan', leaky=0.05), ModelSpec('dcgan', leaky=0.05), d, g, dnn)



def crowd_param_shapes(spec: ModelSpec, image_size: int):
    """Key -> shape of KnnDenseNetCat.state_dict() (crowd/models.py:1060-1134) in the module's own order, for any
    constructor arguments; the MapModule input sizes follow the trunk (28/14/7 at image 224, hard-coded at :1129-1131)."""
    g, bs = spec.growth_rate, spec.bn_size
    n_out = 2 if spec.dggan else 1               # KnnDenseNetCatDggan / MapModuleDggan: crowd/models.py:915,1020
    out = {}

    def bn(prefix, c):
        out[prefix + '.weight'] = (c,); out[prefix + '.bias'] = (c,)
        out[prefix + '.running_mean'] = (c,); out[prefix + '.running_var'] = (c,); out[prefix + '.num_batches_tracked'] = ()
    c = spec.num_init_features
    trans, tap_c = {}, []
    for bi, n in enumerate(spec.block_config, 1):
        for li in range(1, n + 1):
            pre = f'dense_blocks.denseblock{bi}.denselayer{li}'
            bn(pre + '.norm1', c)
            out[pre + '.conv1.weight'] = (bs * g, c, 1, 1)
            bn(pre + '.norm2', bs * g)
            out[pre + '.conv2.weight'] = (g, bs * g, 3, 3)
            c += g
        if bi != len(spec.block_config):
            trans[bi] = c
            c //= 2
            tap_c.append(c)
    # attribute order of the module: dense_blocks, transition_layers (both created at :1069-1070), conv_layer1, norm5
    c2 = spec.num_init_features
    for bi, n in enumerate(spec.block_config, 1):
        c2 += n * g
        if bi in trans:
            pre = f'transition_layers.transition{bi}'
            bn(pre + '.norm', c2)
            out[pre + '.conv.weight'] = (c2 // 2, c2, 1, 1)
            c2 //= 2
    out['conv_layer1.conv0.weight'] = (spec.num_init_features, 3, 7, 7)
    bn('conv_layer1.norm0', spec.num_init_features)
    bn('norm5', c)
    L = spec.label_patch_size
    for i, ci in enumerate(tap_c[:3], 1):
        size = image_size // (8 * 2 ** (i - 1))
        k = L // size
        pre = f'map_module{i}'
        out[pre + '.map_transposed_conv_layer.weight'] = (ci, 1, k, k); out[pre + '.map_transposed_conv_layer.bias'] = (1,)
        out[pre + '.conv1.weight'] = (8, 1, 2, 2); out[pre + '.conv1.bias'] = (8,)
        out[pre + '.conv2.weight'] = (16, 8, 2, 2); out[pre + '.conv2.bias'] = (16,)
        out[pre + '.conv3.weight'] = (32, 16, 2, 2); out[pre + '.conv3.bias'] = (32,)
        out[pre + '.linear1.weight'] = (20, 32, L // 8, L // 8); out[pre + '.linear1.bias'] = (20,)
        out[pre + '.count_layer.weight'] = (n_out, 20, 1, 1); out[pre + '.count_layer.bias'] = (n_out,)
    out['final_count_feature_layer.weight'] = (20, c, 1, 1); out['final_count_feature_layer.bias'] = (20,)
    out['count_layer.weight'] = (n_out, 20, 1, 1); out['count_layer.bias'] = (n_out,)
    return out


def init_crowd_d(spec: ModelSpec, image_size: int, seed=0, dtype=torch.float32, scale=1.0) -> Params:
    """Deterministic KnnDenseNetCat state (NOT the reference initialiser, which downloads DenseNet-201 weights,
    crowd/models.py:1103-1127): fan-in-scaled uniform weights, BatchNorm weight/bias/running statistics drawn away from
    their trivial values so the eval-mode affine is exercised.  Pure torch: reproducible on the GPU box."""
    gen = torch.Generator().manual_seed(seed)
    p = {}
    for k, shape in crowd_param_shapes(spec, image_size).items():
        if k.endswith('num_batches_tracked'):
            p[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith('running_var'):
            p[k] = (torch.rand(shape, generator=gen, dtype=torch.float64) + 0.5).to(dtype)
        elif k.endswith('running_mean'):
            p[k] = _uniform(gen, shape, 0.2, dtype)
        elif '.norm' in k or k.startswith('norm5'):
            p[k] = (torch.rand(shape, generator=gen, dtype=torch.float64) + 0.5).to(dtype) if k.endswith('weight') \
                else _uniform(gen, shape, 0.2, dtype)
        elif k.endswith('.bias'):
            p[k] = _uniform(gen, shape, 0.1, dtype)
        else:
            fan_in = shape[1] * shape[2] * shape[3] if 'map_transposed' not in k else shape[0]
            w = _uniform(gen, shape, math.sqrt(3.0 / fan_in), dtype)
            p[k] = w * scale if ('conv' in k and 'map_module' not in k) else w
    return p


def init_crowd(seed=0, image_size=224, z_dim=256, g_conv_dim=64, dtype=torch.float32, scale=1.0, **spec_kwargs) -> OracleState:
    """CrowdExperiment.model_setup (crowd/srgan.py:92-96): DCGenerator + two KnnDenseNetCat."""
    spec = ModelSpec('crowd', **spec_kwargs)
    d = init_crowd_d(spec, image_size, seed, dtype, scale)
    dnn = init_crowd_d(spec, image_size, seed + 1, dtype, 1.0)            # KnnDenseNetCat does not reseed (App. E.6)
    g = init_dcgan(seed + 2, image_size, g_conv_dim, z_dim, dtype).G
    return OracleState(spec, ModelSpec('dcgan', leaky=0.05), d, g, dnn)


def synthetic_crowd_batch(B, seed, image=224, label=224, z_dim=256, dtype=torch.float32):
    """Synthetic crowd batch of SURVEY 8d config 3, regenerated from the seed wherever it is needed (the tensors are too
    large to commit): images ~U(-1,1), density = Bernoulli point map (~32 heads per 224x224 patch), map = 1/(1+U(0,50));
    plus the three noise draws.  Returns (x, (density, map), u, z, alpha, z2)."""
    gen = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, image, image, generator=gen) * 2 - 1
    u = torch.rand(B, 3, image, image, generator=gen) * 2 - 1
    density = (torch.rand(B, label, label, generator=gen) < 6.5e-4).float()
    map_ = 1 / (1 + torch.rand(B, label, label, generator=gen) * 50)
    z = torch.randn(B, z_dim, generator=gen)
    alpha = torch.rand(B, 1, 1, 1, generator=gen)
    z2 = torch.randn(B, z_dim, generator=gen)
    c = lambda t: t.to(dtype)
    return c(x), (c(density), c(map_)), c(u), c(z), c(alpha), c(z2)
