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
    method: str = 'srgan'                    # 'srgan' | 'dggan'   (settings.py:123-127)
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


# ----------------------------------------------------------------------------------------------------------------
# Model families (forward only; gradients come from autograd).
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class ModelSpec:
    """Which network family a Params dict belongs to, plus the constants its forward needs."""
    family: str                    # 'coefficient' | 'dcgan'
    dggan: bool = False            # D has a second (fake-score) output: coefficient/models.py:53-72
    leaky: float = 0.01            # coefficient: F.leaky_relu default 0.01; dcgan: 0.05 (age/models.py:46-50,70-73)


def coefficient_d_forward(p: Params, x: torch.Tensor, dggan: bool = False):
    """coefficient/models.py:43-50 (MLP) and :65-72 (DgganMLP): 3x(Linear+leaky 0.01) -> features -> Linear."""
    h = F.leaky_relu(F.linear(x, p['linear1.weight'], p['linear1.bias']), 0.01)
    h = F.leaky_relu(F.linear(h, p['linear2.weight'], p['linear2.bias']), 0.01)
    h = F.leaky_relu(F.linear(h, p['linear3.weight'], p['linear3.bias']), 0.01)
    out = F.linear(h, p['linear4.weight'], p['linear4.bias'])
    if dggan:
        return (out[:, 0].squeeze(), out[:, 1].squeeze()), h
    return out.squeeze(), h


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
    return out.reshape(-1), features


def dcgan_g_forward(p: Params, z: torch.Tensor):
    """age/models.py:44-52 and crowd/models.py:139-147: view(B,z,1,1) -> ConvT k=H/16 (no act) -> 3x(ConvT k4 s2 p1 +
    leaky 0.05) -> ConvT + tanh."""
    h = z.reshape(z.size(0), z.size(1), 1, 1)
    h = F.conv_transpose2d(h, p['fc.0.weight'], p['fc.0.bias'], stride=1, padding=0)
    for i in (1, 2, 3):
        h = F.leaky_relu(F.conv_transpose2d(h, p[f'layer{i}.0.weight'], p[f'layer{i}.0.bias'], stride=2, padding=1), 0.05)
    return torch.tanh(F.conv_transpose2d(h, p['layer4.0.weight'], p['layer4.0.bias'], stride=2, padding=1))


def d_forward(spec: ModelSpec, p: Params, x: torch.Tensor):
    """Returns (prediction, fake_score_or_None, features)."""
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


def feature_distance_loss(base_features, other_features, distance: str):
    """srgan.py:438-449 with normalize_feature_norm=False (the True branch is a known bug)."""
    return DISTANCES[distance](base_features.mean(0) - other_features.mean(0))


def bce_with_logits(scores, target_value: float):
    """torch.nn.BCEWithLogitsLoss (mean reduction) against a constant target: coefficient/dggan.py:39-40,48-49,62-63."""
    return F.binary_cross_entropy_with_logits(scores, torch.full_like(scores, target_value))


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
    for k in p:
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


def _leaf(p: Params) -> Params:
    return {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}


def _grads(loss, leaf: Params):
    gs = torch.autograd.grad(loss, list(leaf.values()), allow_unused=True)
    return dict(zip(leaf.keys(), gs))


def dnn_lr(cfg: StepConfig, step: int) -> float:
    """srgan.py:432-436: only the DNN optimizer is decayed (x0.1 every 100k steps)."""
    return cfg.learning_rate * (0.1 ** (step // 100000))


def dnn_training_step(st: OracleState, cfg: StepConfig, x, y, step: int = 0):
    """srgan.py:259-271 + dnn_loss_calculation :322-327 (DG-GAN: coefficient/dggan.py:22-27)."""
    leaf = _leaf(st.DNN)
    pred, _, _ = d_forward(st.d_spec, leaf, x)
    loss = labeled_loss_function(pred, y, cfg.labeled_loss_order) * cfg.labeled_loss_multiplier
    g = _grads(loss, leaf)
    adam_update(st.DNN, g, st.dnn_adam, dnn_lr(cfg, step), cfg.weight_decay, cfg.betas, cfg.eps)
    return {'dnn_loss': float(loss.detach())}


def gradient_penalty(st_spec: ModelSpec, leafD: Params, cfg: StepConfig, fake, u, alpha):
    """srgan.py:360-375 + interpolate_loss_calculation :377-381 (DG-GAN target = raw fake score,
    coefficient/dggan.py:54-57)."""
    interp = (alpha * u.detach() + (1 - alpha) * fake.detach()).requires_grad_(True)
    _, score, feats = d_forward(st_spec, leafD, interp)
    if cfg.method == 'dggan':
        target = score
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
    labeled = labeled_loss_function(pred, y, cfg.labeled_loss_order) * cfg.labeled_loss_multiplier
    # -- unlabeled (:283, :337-346 | dggan.py:36-43)
    _, score_u, f_u = d_forward(spec, leafD, u)
    with torch.no_grad():
        fake = g_forward(st.g_spec, st.G, z)                    # :290  (graph unused: fake is detached / G grads zeroed)
    _, score_f, f_f = d_forward(spec, leafD, fake)
    if cfg.method == 'dggan':
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
        _, score_f2, f_f2 = d_forward(spec, st.D, fake2)
        if cfg.method == 'dggan':
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


def init_coefficient(seed=0, hidden=10, dggan=False, dtype=torch.float32) -> OracleState:
    """Shapes of coefficient/models.py:12-72 (input 50 = observation_count 10 x irrelevant_data_multiplier 5)."""
    gen = torch.Generator().manual_seed(seed)

    def mlp(sizes):
        p = {}
        for i, (a, b) in enumerate(zip(sizes[:-1], sizes[1:]), 1):
            bound = 1 / math.sqrt(a)
            p[f'linear{i}.weight'] = _uniform(gen, (b, a), bound, dtype)
            p[f'linear{i}.bias'] = _uniform(gen, (b,), bound, dtype)
        return p
    d = mlp([50, hidden, hidden, hidden, 2 if dggan else 1])
    dnn = {k: v.clone() for k, v in d.items()}               # D and DNN start identical
    g = mlp([10, hidden, hidden, hidden, 50])
    return OracleState(ModelSpec('coefficient', dggan=dggan), ModelSpec('coefficient'), d, g, dnn)


def init_dcgan(seed=0, image_size=128, conv_dim=64, z_dim=256, dtype=torch.float32, scale=1.0) -> OracleState:
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
    d['layer5.0.weight'] = _uniform(gen, (1, chans[4], k, k), bound, dtype)
    d['layer5.0.bias'] = _uniform(gen, (1,), bound, dtype)
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
