"""
Host-side mirror of the reference's operator interface for the hot path.

Two ways in, both with the reference's names, argument meaning and error behaviour:

1. `B200StepMixin` -- mix into any reference `Experiment` subclass (coefficient / age / driving SR-GAN, coefficient
   DG-GAN):  `class Fast(B200StepMixin, AgeExperiment): pass`.  It overrides only
   `dnn_training_step(examples, labels, step)` (srgan.py:259-271) and
   `gan_training_step(labeled_examples, labels, unlabeled_examples, step)` (srgan.py:273-320); `train`,
   `training_loop`, `prepare_optimizers`, `save_models` / `load_models`, datasets and `run.py` stay the reference's.
   The same nn.Parameter objects are updated in place and the torch.optim.Adam `state` is kept loadable.
2. `Experiment` below -- a stand-alone mirror (this package's `Settings` + parameter containers with the reference's
   `state_dict` keys) for machines without the reference checkout (the GPU box, bench.py, tests).

Unsupported configurations raise (no silent fallback): `normalize_feature_norm=True` (srgan.py:447 bug, SURVEY
App. E.1), distance callables not in utility.py:201-243, modules without a B200 path (nets.describe_module).
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
from torch import nn

from . import nets
from .engine import (Engine, DIST_KINDS, SC_DNN, SC_LABELED, SC_UNLABELED, SC_FAKE, SC_GP, SC_GNORM, SC_GEN,
                     partial_scalar_slots)


# ------------------------------------------------------------------------------------------------ utility.py mirror
def abs_mean(tensor):
    """utility.py:226-228."""
    return tensor.abs().mean()


def abs_mean_neg(tensor):
    """utility.py:221-223."""
    return tensor.abs().mean().neg()


def abs_plus_one_sqrt_mean_neg(tensor):
    """utility.py:216-218."""
    return tensor.abs().add(1).sqrt().mean().neg()


def abs_plus_one_log_mean_neg(tensor):
    """utility.py:211-213."""
    return tensor.abs().add(1).log().mean().neg()


def square_mean(tensor):
    """utility.py:241-243."""
    return tensor.pow(2).mean()


def norm_mean(tensor):
    """utility.py:236-238."""
    return tensor.pow(2).sum().pow(0.5)


class Settings:
    """settings.py:12-67 -- same attribute names and defaults (only the attributes the step reads plus the generic
    ones; application/data attributes are carried verbatim when present on a reference Settings object)."""

    def __init__(self):
        self.trial_name = 'base'
        self.steps_to_run = 200000
        self.batch_size = 1000
        self.summary_step_period = 2000
        self.learning_rate = 1e-4
        self.weight_decay = 0
        self.labeled_loss_multiplier = 1e0
        self.matching_loss_multiplier = 1e0
        self.contrasting_loss_multiplier = 1e0
        self.srgan_loss_multiplier = 1e0
        self.dggan_loss_multiplier = 1e1
        self.gradient_penalty_on = True
        self.gradient_penalty_multiplier = 1e1
        self.mean_offset = 0
        self.labeled_loss_order = 2
        self.generator_training_step_period = 1
        self.normalize_fake_loss = False
        self.normalize_feature_norm = False
        self.contrasting_distance_function = abs_plus_one_sqrt_mean_neg
        self.matching_distance_function = abs_mean
        self.hidden_size = 10
        self.map_multiplier = 1e-6
        self.number_of_bins = 10         # settings.py:67 (SGAN)
        self.bins = None                 # SGAN: the bin centres (the reference experiments build them: age/sgan.py:14)
        self.async_checkpoint = False    # save_models through checkpoint.AsyncWriter (off = the reference's blocking torch.save)
        # new knobs (added attributes only, defaults = reference behaviour): SURVEY section 5
        self.precision = 'fp32'          # 'fp32' (SIMT, 1e-4 parity) | 'bf16' (tcgen05 tensor cores, 2e-2 parity)
        self.use_cuda_graph = True       # replay the step methods from CUDA graphs (piecewise around collectives)
        self.use_persistent_kernel = True    # coefficient application, single rank: one cooperative kernel per step method
        self.overlap_dnn_step = True     # single rank: the DNN step runs on its own stream, overlapped with the GAN step
        self.micro_batch = 0             # > 0: run the step in micro-batches of this many samples (exact; bounds activation memory)


class StepConfig:
    """What engine.py needs from a (reference or mirror) settings object, validated."""

    def __init__(self, settings, method: str):
        if getattr(settings, 'normalize_feature_norm', False):
            raise ValueError('normalize_feature_norm=True is not supported (reference bug at srgan.py:447)')
        self.method = method
        for k in ('learning_rate', 'weight_decay', 'labeled_loss_multiplier', 'matching_loss_multiplier',
                  'contrasting_loss_multiplier', 'srgan_loss_multiplier', 'dggan_loss_multiplier',
                  'gradient_penalty_multiplier'):
            setattr(self, k, float(getattr(settings, k)))
        self.labeled_loss_order = int(settings.labeled_loss_order)
        self.generator_training_step_period = int(settings.generator_training_step_period)
        self.mean_offset = float(getattr(settings, 'mean_offset', 0))
        self.map_multiplier = float(getattr(settings, 'map_multiplier', 1e-6))
        self.micro_batch = int(getattr(settings, 'micro_batch', 0) or 0)
        self.batch_size = int(settings.batch_size)
        for k in ('matching_distance_function', 'contrasting_distance_function'):
            fn = getattr(settings, k)
            name = fn if isinstance(fn, str) else getattr(fn, '__name__', None)
            if name not in DIST_KINDS:
                raise ValueError(f'{k}={name!r}: only the scalar distance functions of utility.py:201-243 have a '
                                 f'B200 path ({sorted(DIST_KINDS)})')
            setattr(self, k, name)
        self.betas = (0.9, 0.999)
        self.eps = 1e-8
        # sgan: the bin centres the experiment builds in __init__ (age/sgan.py:14, coefficient/sgan.py:15)
        bins = getattr(settings, 'bins', None)
        self.bins = tuple(float(v) for v in bins) if bins is not None else ()


# ------------------------------------------------------------------------------------------------ parameter containers
def _seed_all(seed):
    """utility.py:110-116 (the model constructors call seed_all(0): SURVEY App. E.6)."""
    import random
    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


class _Container(nn.Module):
    features = None

    def forward(self, *a, **k):
        raise NotImplementedError('parameter container: run the network through StepRunner.predict()/generate()')


class CoefficientMLP(_Container):
    """coefficient/models.py:31-50 (MLP) / :53-72 (DgganMLP when n_out=2): same state_dict keys."""

    def __init__(self, hidden_size=10, n_out=1, n_in=50):
        super().__init__()
        _seed_all(0)
        self.linear1 = nn.Linear(n_in, hidden_size)
        self.linear2 = nn.Linear(hidden_size, hidden_size)
        self.linear3 = nn.Linear(hidden_size, hidden_size)
        self.linear4 = nn.Linear(hidden_size, n_out)


class CoefficientGenerator(_Container):
    """coefficient/models.py:12-28."""

    def __init__(self, hidden_size=10, n_out=50):
        super().__init__()
        self.input_size = 10
        self.linear1 = nn.Linear(self.input_size, hidden_size)
        self.linear2 = nn.Linear(hidden_size, hidden_size)
        self.linear3 = nn.Linear(hidden_size, hidden_size)
        self.linear4 = nn.Linear(hidden_size, n_out)


class DcganDiscriminator(_Container):
    """age/models.py:55-80 == driving/models.py."""

    def __init__(self, image_size=128, conv_dim=64, number_of_outputs=1):
        super().__init__()
        _seed_all(0)
        c = conv_dim
        self.layer1 = nn.Sequential(nn.Conv2d(3, c, 4, 2, 1))
        self.layer2 = nn.Sequential(nn.Conv2d(c, c * 2, 4, 2, 1))
        self.layer3 = nn.Sequential(nn.Conv2d(c * 2, c * 4, 4, 2, 1))
        self.layer4 = nn.Sequential(nn.Conv2d(c * 4, c * 8, 4, 2, 1))
        self.layer5 = nn.Sequential(nn.Conv2d(c * 8, number_of_outputs, image_size // 16, 1, 0))


class DcganGenerator(_Container):
    """age/models.py:32-52 (image_size=128) / crowd/models.py:127-147 DCGenerator (image_size=224)."""

    def __init__(self, z_dim=256, image_size=128, conv_dim=64):
        super().__init__()
        _seed_all(0)
        c = conv_dim
        self.fc = nn.Sequential(nn.ConvTranspose2d(z_dim, c * 8, image_size // 16, 1, 0))
        self.layer1 = nn.Sequential(nn.ConvTranspose2d(c * 8, c * 4, 4, 2, 1))
        self.layer2 = nn.Sequential(nn.ConvTranspose2d(c * 4, c * 2, 4, 2, 1))
        self.layer3 = nn.Sequential(nn.ConvTranspose2d(c * 2, c, 4, 2, 1))
        self.layer4 = nn.Sequential(nn.ConvTranspose2d(c, 3, 4, 2, 1))
        self.input_size = z_dim


class _Named(nn.Module):
    """A bag of named sub-modules (only their parameters / buffers matter here)."""

    def __init__(self, **mods):
        super().__init__()
        for k, m in mods.items():
            self.add_module(k, m)


class KnnDenseNetCat(_Container):
    """crowd/models.py:1049-1166: parameter container with the reference module's state_dict keys and order (dense_blocks,
    transition_layers, conv_layer1, norm5, map_module1..3, final_count_feature_layer, count_layer).  BatchNorm layers are
    real nn.BatchNorm2d modules so the running statistics travel in the state_dict; the step treats them as the frozen
    per-channel affine srgan.py:538-542 makes of them.  No weights are downloaded (the reference loads DenseNet-201 from
    the model zoo, crowd/models.py:1103-1127): load a state_dict to start from pretrained weights."""

    def __init__(self, growth_rate=32, block_config=(6, 12, 48, 32), num_init_features=64, bn_size=4, label_patch_size=224,
                 image_size=224, number_of_outputs=1):
        super().__init__()
        g, bs = growth_rate, bn_size
        self.dense_blocks = nn.ModuleList()
        self.transition_layers = nn.ModuleList()
        self.conv_layer1 = _Named(conv0=nn.Conv2d(3, num_init_features, 7, 2, 3, bias=False),
                                  norm0=nn.BatchNorm2d(num_init_features))
        c, taps = num_init_features, []
        for bi, n in enumerate(block_config, 1):
            block = nn.Module()
            for li in range(1, n + 1):
                block.add_module(f'denselayer{li}', _Named(norm1=nn.BatchNorm2d(c), conv1=nn.Conv2d(c, bs * g, 1, bias=False),
                                                           norm2=nn.BatchNorm2d(bs * g),
                                                           conv2=nn.Conv2d(bs * g, g, 3, padding=1, bias=False)))
                c += g
            self.dense_blocks.add_module(f'denseblock{bi}', block)
            if bi != len(block_config):
                self.transition_layers.add_module(f'transition{bi}', _Named(norm=nn.BatchNorm2d(c),
                                                                             conv=nn.Conv2d(c, c // 2, 1, bias=False)))
                c //= 2
                taps.append(c)
        self.norm5 = nn.BatchNorm2d(c)
        L = label_patch_size
        for i, ci in enumerate(taps[:3], 1):
            k = L // (image_size // (8 * 2 ** (i - 1)))
            self.add_module(f'map_module{i}', _Named(
                map_transposed_conv_layer=nn.ConvTranspose2d(ci, 1, k, k), conv1=nn.Conv2d(1, 8, 2, 2), conv2=nn.Conv2d(8, 16, 2, 2),
                conv3=nn.Conv2d(16, 32, 2, 2), linear1=nn.Conv2d(32, 20, L // 8),
                count_layer=nn.Conv2d(20, number_of_outputs, 1)))
        self.final_count_feature_layer = nn.Conv2d(c, 20, 1)
        self.count_layer = nn.Conv2d(20, number_of_outputs, 1)     # 2: KnnDenseNetCatDggan (crowd/models.py:929-1046)


# ------------------------------------------------------------------------------------------------ the runner
class StepRunner:
    """Owns the Engine for one (D, G, DNN) triple and the bookkeeping that keeps torch-side objects coherent."""

    def __init__(self, D: nn.Module, G: nn.Module, DNN: nn.Module, settings, method='srgan', precision=None,
                 comm=None, device=None):
        from .ops_cuda import CudaOps              # raises without CUDA / without the built library
        precision = precision or getattr(settings, 'precision', 'fp32')
        if precision not in ('fp32', 'bf16'):
            raise ValueError(f'precision={precision!r}')
        self.precision = precision
        self.method = method
        self.settings = settings
        dev = device or next(D.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('the B200 step needs the networks on a CUDA device (call gpu_mode() first)')
        self.device = dev
        self.modules = dict(D=D, G=G, DNN=DNN)
        # bf16 mode: dense layers write / read their channel window of the concat buffers in place (tcgen05 kernels only)
        direct = precision == 'bf16' and os.environ.get('SRGAN_NO_DIRECT_CONCAT', '0') != '1'
        # ... and the BatchNorm + ReLU in front of every trunk 1x1 convolution is fused into that convolution's kernels
        fuse = int(os.environ.get('SRGAN_FUSE_BN', '4')) if precision == 'bf16' else 0
        d_net, g_net = nets.describe_module(D, direct, fuse), nets.describe_module(G)
        if nets.describe_module(DNN, direct, fuse) != d_net:
            raise ValueError('DNN and D must share an architecture (srgan.py model_setup)')
        if method not in ('srgan', 'dggan', 'sgan'):
            raise ValueError(f'method={method!r}: the B200 path covers srgan, dggan and sgan')
        if method == 'sgan':
            if d_net.family not in ('dcgan', 'coefficient') or not 3 <= d_net.head_outputs <= 16:
                raise ValueError('SGAN (sgan.py) is covered for the class-logit discriminators of age/sgan.py, driving and '
                                 'coefficient/sgan.py with 3..16 bins; crowd/sgan.py\'s JointDCDiscriminator is not')
            if len(getattr(settings, 'bins', None) or ()) != d_net.head_outputs:
                raise ValueError(f'SGAN: settings.bins must hold the {d_net.head_outputs} bin centres of the experiment '
                                 '(AgeSganExperiment.bins, age/sgan.py:14); the mix-in copies them from self.bins')
        if method == 'dggan' and d_net.head_outputs != 2:
            raise ValueError('DG-GAN needs the two-output discriminator (coefficient/models.py:53-72, crowd/models.py:929-1046)')
        if method == 'srgan' and d_net.head_outputs != 1:
            raise ValueError('SR-GAN feature matching expects a one-output discriminator; a two-output (DG-GAN) module was '
                             'given: use method=\'dggan\' (coefficient/dggan.py, crowd/dggan.py)')

        def pd(m):
            d = {k: p for k, p in m.named_parameters()}
            d.update({k: b for k, b in m.named_buffers() if b.is_floating_point()})      # BatchNorm running statistics
            return d
        self.engine = Engine(CudaOps(dev), d_net, g_net, pd(D), pd(G), pd(DNN),
                             act_dtype=torch.float32 if precision == 'fp32' else torch.bfloat16, device=dev, comm=comm)
        # device noise stream (draw_noise): every rank draws its OWN shard of the global batch's z / alpha / z2
        self.generator = torch.Generator(device=dev)
        self.generator.manual_seed(int(getattr(settings, 'noise_seed', 0)) + (comm.rank if comm is not None else 0))
        ug = getattr(settings, 'use_cuda_graph', True)
        # multi-rank steps are captured PIECEWISE: one graph segment between every two collectives, the NCCL all-reduces
        # (feature sums, gradients) are issued eagerly between the segments on the same stream (_capture_segments)
        self.use_cuda_graph = bool(ug) and os.environ.get('SRGAN_NO_GRAPH', '0') != '1' 
        self._graphs, self._statics = {}, {}
        self._side, self._pending = None, {}       # deferred (overlapped) gradient all-reduce + Adam groups: _run_deferred
        # single rank: the DNN step (its own network, gradients, moments and -- Engine._scope -- scratch buffers) is
        # enqueued on its own stream and overlaps the GAN step that follows it; gan_step joins the two streams at its end,
        # so everything is ordered on the caller's stream again when a step pair returns.  The crowd step is thousands of
        # launches that under-fill the GPU (7x7 / 14x14 stages): two independent chains in flight fill the gaps.
        # Multi-rank as well: the DNN step's gradient all-reduce (a deferred group on the side stream) is issued before the GAN
        # step's collectives on every rank, and NCCL orders the collectives of one communicator across streams, so the GAN
        # step's first feature-sum all-reduce (after its D forward) waits for the DNN backward at most.
        self.overlap_dnn = (bool(getattr(settings, 'overlap_dnn_step', True)) and os.environ.get('SRGAN_NO_OVERLAP', '0') != '1'
                            and (comm is None or os.environ.get('SRGAN_NO_OVERLAP_MULTI', '0') != '1'))
        self._dnn_stream, self._dnn_done = None, None
        # coefficient application: one persistent cooperative kernel per step method (csrc/coef_step.cu) instead of
        # ~150 generic launches; single rank only (the feature sums are combined inside the kernel)
        self.persistent = (method != 'sgan' and self._persistent_shape_ok(d_net, g_net) and comm is None
                           and bool(getattr(settings, 'use_persistent_kernel', True))
                           and os.environ.get('SRGAN_NO_PERSISTENT', '0') != '1')
        self._coef_tables = None
        self._layouts_stale = False

    @staticmethod
    def _persistent_shape_ok(d_net, g_net):
        def dims(net):
            return [(l.geom.Cb, l.geom.Ca) for l in net.layers]
        return (d_net.family == 'coefficient' and g_net.family == 'coefficient'
                and dims(d_net) == [(50, 10), (10, 10), (10, 10)] and d_net.head_outputs in (1, 2)
                and dims(g_net) == [(10, 10), (10, 10), (10, 10), (10, 50)]
                and all(abs(l.slope - 0.01) < 1e-12 for l in d_net.layers + g_net.layers[:3]))

    def _coef_step(self, phases, x, y, u=None, z=None, alpha=None, z2=None, lr_dnn=0.0, wd=None, train_g=True):
        """srgan_coefficient_step: phases 1 = dnn_training_step, 2 = gan_training_step."""
        eng, cfg = self.engine, self.config()
        ops = eng.ops
        ops.begin()
        if self._coef_tables is None:
            def table(st):
                keys = [l.name for l in st.net.layers] + ([st.net.head] if st.net.head else [])
                ts = []
                for k in keys:
                    w, b = k + '.weight', k + '.bias'
                    ts += [st.params[w].detach(), st.params[b].detach(), st.g(w), st.g(b), st.m(w), st.m(b), st.v(w), st.v(b)]
                return ops.pointer_table(ts)
            ws = torch.empty(ops.coefficient_workspace_bytes() // 4, dtype=torch.float32, device=self.device)
            self._coef_tables = (table(eng.D), table(eng.G), table(eng.DNN), ws)
        td, tg, tdnn, ws = self._coef_tables

        def f32(t):
            return None if t is None else t.detach().to(torch.float32).contiguous()
        x, y, u, z, alpha, z2 = (f32(t) for t in (x, y, u, z, alpha, z2))
        B = x.shape[0]
        if x.numel() != B * 50 or y.numel() != B:
            raise ValueError('coefficient step: examples must be [B, 50] and labels [B]')
        if phases & 2:
            if u.shape[0] != B or z.shape[0] != B or alpha.numel() != B or z2.shape[0] != B:
                raise ValueError('labeled, unlabeled and noise batches must have the same size (SURVEY App. E.2)')
        dggan = cfg.method == 'dggan'
        if dggan:
            unl = cfg.matching_loss_multiplier * cfg.dggan_loss_multiplier
            fake = cfg.contrasting_loss_multiplier * cfg.dggan_loss_multiplier
            gen = 1.0
        else:
            unl = cfg.matching_loss_multiplier * cfg.srgan_loss_multiplier
            fake = cfg.contrasting_loss_multiplier * cfg.srgan_loss_multiplier
            gen = cfg.matching_loss_multiplier
        dummy = x
        ops.coefficient_step(td, tg, tdnn, eng.D.adam_state, eng.G.adam_state, eng.DNN.adam_state, x, y,
                             u if u is not None else dummy, z if z is not None else dummy,
                             alpha if alpha is not None else dummy, z2 if z2 is not None else dummy, B, 1.0 / B, dggan,
                             cfg.labeled_loss_order, cfg.labeled_loss_multiplier, unl, fake, gen,
                             cfg.gradient_penalty_multiplier, DIST_KINDS[cfg.matching_distance_function],
                             DIST_KINDS[cfg.contrasting_distance_function], cfg.learning_rate, lr_dnn,
                             cfg.weight_decay if wd is None else wd, cfg.betas[0], cfg.betas[1], cfg.eps, phases, train_g,
                             ws, eng.scalars, self._coef_publish(B) if (phases & 2) and eng.publish_features else None)
        self._layouts_stale = True

    def _coef_publish(self, B):
        """[4][B][10] features + [B] gradient norms written by the persistent kernel (srgan.py:332-386 side effects)."""
        t = getattr(self, '_coef_pub', None)
        if t is None or t.numel() != 41 * B:
            t = self._coef_pub = torch.zeros(41 * B, dtype=torch.float32, device=self.device)
        return t

    def _fresh_layouts(self):
        """The persistent kernel updates only the master parameters: rebuild the kernel-layout copies before the
        generic kernels (predict / generate) read them."""
        if self._layouts_stale:
            self.refresh_weights()

    # ---- noise (srgan.py:286-289, :364, :301) drawn on the device
    def draw_noise(self, B, cfg):
        zdim = self.engine.g_net.input_chw[0]
        g = self.generator
        z = torch.randn(B, zdim, device=self.device, generator=g)
        if cfg.mean_offset != 0:
            sign = (torch.rand(B, zdim, device=self.device, generator=g) < 0.5).float() * 2 - 1
            z = z + sign * cfg.mean_offset
        alpha = torch.rand(B, device=self.device, generator=g)
        z2 = torch.randn(B, zdim, device=self.device, generator=g)
        return z, alpha, z2

    def config(self):
        return StepConfig(self.settings, self.method)

    # ---- CUDA-graph replay of the two step methods -----------------------------------------------------------------
    # The step is ~140 small-to-medium launches; replaying a captured graph removes the per-launch host cost (3.2 ms of
    # Python per age step) and the inter-kernel gaps.  Everything step-dependent lives in device memory (inputs and
    # noise are copied into static buffers, Adam's bias corrections come from srgan_adam_prepare), so a graph is valid
    # until the batch size, lr / weight decay or the generator flag changes.  First call = eager (allocates the
    # workspaces), second call = capture, later calls = replay.  SRGAN_NO_GRAPH=1 or use_cuda_graph=False -> eager.
    def _graphed(self, key, statics, fn, launch=None):
        """statics: list of (static_buffer, source tensor) copied (on the caller's stream) before replay; fn(): enqueues
        the step on statics; launch(f): runs the enqueueing callable f (default: in place; the DNN step passes
        _on_dnn_stream)."""
        launch = launch or (lambda f: f())
        entry = self._graphs.get(key)
        gen = self.engine.buf_generation
        if entry is not None and entry['gen'] != gen:
            # an engine scratch buffer was reallocated since this graph was captured (a larger batch shape, predict() /
            # generate() on a bigger batch): the addresses baked into the graph may dangle -- capture again
            entry = None
        if entry is None:
            for dst, src in statics:
                dst.copy_(src)
            launch(fn)
            # the static input buffers stay referenced by the entry for as long as its graph may replay
            self._graphs[key] = {'calls': 1, 'graph': None, 'gen': self.engine.buf_generation, 'statics': [d for d, _ in statics]}
            return
        for dst, src in statics:
            dst.copy_(src)
        if entry['graph'] is None:
            torch.cuda.synchronize(self.device)
            entry['graph'] = self._capture_segments(fn)
            if self.engine.buf_generation != entry['gen']:
                raise RuntimeError('engine buffers were reallocated during graph capture')

        def replay():
            for seg in entry['graph']:
                if isinstance(seg, tuple):
                    self._run_deferred(*seg)
                else:
                    seg()
        launch(replay)

    def _on_dnn_stream(self, f):
        """Runs f() on the DNN stream, after everything enqueued on the caller's stream so far."""
        main = torch.cuda.current_stream(self.device)
        if self._dnn_stream is None:
            self._dnn_stream = torch.cuda.Stream(self.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self._dnn_stream.wait_event(ev)
        with torch.cuda.stream(self._dnn_stream):
            f()
            done = torch.cuda.Event()
            done.record(self._dnn_stream)
        self._dnn_done = done

    def _join_dnn(self):
        """The caller's stream waits for the DNN step in flight (if any)."""
        if self._dnn_done is not None:
            torch.cuda.current_stream(self.device).wait_event(self._dnn_done)
            self._dnn_done = None

    def _run_deferred(self, tag, items):
        """A deferred group (gradient all-reduce + the Adam graph of one network) on the side stream, after everything
        enqueued so far; whoever touches that network next waits for it (_wait_pending)."""
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self._side.wait_event(ev)
        with torch.cuda.stream(self._side):
            for it in items:
                it()
            done = torch.cuda.Event()
            done.record(self._side)
        self._pending[tag] = done

    def _wait_pending(self, tags=None):
        """Makes the current stream wait for the deferred updates of the given networks (None: all).  dnn_step touches
        only DNN (parameters, gradient buffer, moments; the activation scratch it shares with D is not used by an
        update), gan_step only D and G: so DNN's all-reduce + Adam overlap the whole GAN step, G's the next DNN step."""
        if tags is None or 'DNN' in tags:
            self._join_dnn()
        if self._pending:
            main = torch.cuda.current_stream(self.device)
            for tag in list(self._pending) if tags is None else tags:
                ev = self._pending.pop(tag, None)
                if ev is not None:
                    main.wait_event(ev)

    def _capture_segments(self, fn):
        """Captures fn() as a list of callables: CUDA-graph replays, separated (multi-rank only) by the collectives the
        step issues through Comm.all_reduce_sum, which are not captured but re-issued eagerly at the same places."""
        comm = self.engine.comm
        segs, state = [], {'g': None}
        pool = torch.cuda.graph_pool_handle()
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))

        def begin():
            g = torch.cuda.CUDAGraph()
            g.capture_begin(pool=pool, capture_error_mode='thread_local')
            state['g'] = g

        def end():
            g, state['g'] = state['g'], None
            g.capture_end()
            cur['list'].append(g.replay)

        cur = {'list': segs, 'closed': False}

        def on_collective(kind, arg):
            if cur['closed']:
                raise RuntimeError('a deferrable update must be the last thing its step method enqueues')
            end()
            if kind == 'all_reduce':
                cur['list'].append(lambda t=arg: comm.all_reduce_sum(t))
                comm.all_reduce_sum_now(arg)       # keeps the ranks' NCCL call sequences aligned during capture as well
                begin()
            elif kind == 'defer_begin':
                items = []
                segs.append((arg, items))          # (tag, items): replayed by _run_deferred on the side stream
                cur['list'] = items
                begin()
            else:                                  # defer_end: nothing may follow
                cur['closed'] = True
        with torch.cuda.stream(side):
            if comm is not None:
                comm.capture_hook = on_collective
            try:
                begin()
                fn()
                if not cur['closed']:
                    end()
            finally:
                if comm is not None:
                    comm.capture_hook = None
        torch.cuda.current_stream(self.device).wait_stream(side)
        return segs

    def _static(self, name, like, dtype=torch.float32):
        """Static graph input, one per (name, shape): a graph captured for one batch shape keeps ITS buffers when another
        shape comes by (shapes A, B, A replay graph A on the buffers it was captured with)."""
        key = (name, tuple(like.shape))
        t = self._statics.get(key)
        if t is None:
            t = torch.empty(like.shape, dtype=dtype, device=self.device)
            self._statics[key] = t
        return t

    def _static_labels(self, name, labels):
        """labels: a tensor, or the crowd (density, map) tuple built at srgan.py:112-113."""
        if isinstance(labels, (tuple, list)):
            ys = tuple(self._static(f'{name}{i}', t) for i, t in enumerate(labels))
            return ys, list(zip(ys, labels))
        ys = self._static(name, labels)
        return ys, [(ys, labels)]

    def dnn_step(self, examples, labels, lr=None, weight_decay=None):
        cfg = self.config()
        lr = cfg.learning_rate if lr is None else lr
        wd = cfg.weight_decay if weight_decay is None else weight_decay
        self._wait_pending(('DNN',))
        if self.persistent:
            if not self.use_cuda_graph or os.environ.get('SRGAN_COEF_GRAPH', '0') != '1':
                self._coef_step(1, examples, labels, lr_dnn=lr, wd=wd)
                return
            # SRGAN_COEF_GRAPH=1: the cooperative launch replayed from a CUDA graph (north_star: "one persistent CUDA-graph
            # kernel").  Opt-in because it measures SLOWER: the step is device-bound (two launches of ~115 us, the host needs
            # ~40 us for both), and a replay adds the copies into the static inputs: 262.5 vs 235.7 us/step at B = 5000
            # (profiles/r2_final_coef_graph.txt)
            xs, ys = self._static('cx', examples), self._static('cy', labels.reshape(-1))
            key = ('coef_dnn', tuple(examples.shape), lr, wd, repr(sorted(vars(cfg).items())))
            self._graphed(key, [(xs, examples), (ys, labels.reshape(-1))], lambda: self._coef_step(1, xs, ys, lr_dnn=lr, wd=wd))
            self._layouts_stale = True              # (a replay does not pass through _coef_step)
            return
        mb = cfg.micro_batch if 0 < cfg.micro_batch < examples.shape[0] else 0

        def run(xx, yy):
            if mb:
                self.engine.dnn_step_micro(xx, yy, cfg, lr, wd, mb)
            else:
                self.engine.dnn_step(xx, yy, cfg, lr, wd)
        launch = self._on_dnn_stream if self.overlap_dnn else None
        if not self.use_cuda_graph:
            if launch is None:
                run(examples, labels)
            else:                                   # eager: the inputs themselves are read on the DNN stream
                launch(lambda: run(examples, labels))
                for t in (examples,) + (tuple(labels) if isinstance(labels, (tuple, list)) else (labels,)):
                    t.record_stream(self._dnn_stream)
            return
        xs = self._static('dnn_x', examples)
        ys, ypairs = self._static_labels('dnn_y', labels)
        key = ('dnn', tuple(examples.shape), lr, wd, cfg.labeled_loss_multiplier, cfg.labeled_loss_order, cfg.map_multiplier, mb)
        self._graphed(key, [(xs, examples)] + ypairs, lambda: run(xs, ys), launch)

    def gan_step(self, labeled_examples, labels, unlabeled_examples, step=0, noise=None):
        cfg = self.config()
        B = labeled_examples.shape[0]
        z, alpha, z2 = noise if noise is not None else self.draw_noise(B, cfg)
        train_g = (step % cfg.generator_training_step_period == 0)
        self._wait_pending(('G',))
        if self.persistent:
            if not self.use_cuda_graph or os.environ.get('SRGAN_COEF_GRAPH', '0') != '1':
                self._coef_step(2, labeled_examples, labels, unlabeled_examples, z, alpha.reshape(-1), z2, train_g=train_g)
                return
            al, yl = alpha.reshape(-1), labels.reshape(-1)
            st = [(self._static('cx', labeled_examples), labeled_examples), (self._static('cy', yl), yl),
                  (self._static('cu', unlabeled_examples), unlabeled_examples), (self._static('cz', z), z),
                  (self._static('calpha', al), al), (self._static('cz2', z2), z2)]
            xs, ys, us, zs, als, z2s = (d for d, _ in st)
            key = ('coef_gan', tuple(labeled_examples.shape), train_g, bool(self.engine.publish_features),
                   repr(sorted(vars(cfg).items())))
            self._graphed(key, st, lambda: self._coef_step(2, xs, ys, us, zs, als, z2s, train_g=train_g))
            self._layouts_stale = True
            return
        mb = cfg.micro_batch if 0 < cfg.micro_batch < B else 0

        def run(xx, yy, uu, zz, aa, zz2):
            if mb:
                self.engine.gan_step_micro(xx, yy, uu, zz, aa, zz2, cfg, train_g, mb)
            else:
                self.engine.gan_step(xx, yy, uu, zz, aa, zz2, cfg, train_generator=train_g)
        if not self.use_cuda_graph:
            run(labeled_examples, labels, unlabeled_examples, z, alpha, z2)
            self._join_dnn()
            return
        alpha = alpha.reshape(-1)
        ys, ypairs = self._static_labels('y', labels)
        st = [(self._static('x', labeled_examples), labeled_examples),
              (self._static('u', unlabeled_examples), unlabeled_examples), (self._static('z', z), z),
              (self._static('alpha', alpha), alpha), (self._static('z2', z2), z2)]
        key = ('gan', tuple(labeled_examples.shape), train_g, repr(sorted(vars(cfg).items())))
        xs, us, zs, als, z2s = (d for d, _ in st)
        self._graphed(key, st + ypairs, lambda: run(xs, ys, us, zs, als, z2s))
        self._join_dnn()

    def scalars(self):
        """One device->host read of the step's scalars (the .item() calls of srgan.py:268-270, 306-319)."""
        self._wait_pending()
        s = self.engine.scalars
        if self.engine.comm is not None:
            s = s.clone()
            self.engine.comm.all_reduce_sum_partial(s, partial_scalar_slots(self.method))
        v = s.tolist()
        return {'dnn_loss': v[SC_DNN], 'labeled_loss': v[SC_LABELED], 'unlabeled_loss': v[SC_UNLABELED],
                'fake_loss': v[SC_FAKE], 'gradient_penalty': v[SC_GP], 'gradient_norm_mean': v[SC_GNORM],
                'generator_loss': v[SC_GEN]}

    # ---- srgan.py side effects of the last GAN step (device tensors in the reference's layout)
    FEATURE_BLOCKS = {'labeled': 0, 'unlabeled': 1, 'fake': 2, 'interpolates': 3}

    def step_features(self, which):
        """`.features` of the last gan_step in the reference's order and shape ([B, C*H*W] for the DCGAN discriminator,
        age/models.py:74; [B, 80, 1, 1] for KnnDenseNetCat, crowd/models.py:1163; [B, 10] for the coefficient MLP):
        'labeled' / 'unlabeled' / 'interpolates' with the discriminator of the D step, 'fake' = the generator step's
        G(z2) under the updated discriminator when the generator was trained (srgan.py:332-386).  Needs
        engine.publish_features (set before the step); a per-rank shard under data parallelism."""
        eng = self.engine
        if not eng.publish_features or self.method != 'srgan':
            return None
        self._wait_pending()
        F = eng.d_net.feature_size
        if self.persistent:
            pub = getattr(self, '_coef_pub', None)
            if pub is None:
                return None
            B = pub.numel() // 41
            j = self.FEATURE_BLOCKS[which]
            return pub[j * B * F:(j + 1) * B * F].view(B, F).clone()
        snap = eng._buf.get(('gan', 'feat_snap'))
        if snap is None:
            return None
        B = eng.last_gan_batch
        j = self.FEATURE_BLOCKS[which]
        eng.ops.begin()
        out = self.features_nchw(snap[j * B * F:(j + 1) * B * F], B)
        return out.view(B, F, 1, 1) if eng.d_net.family == 'crowd' else out

    def dnn_step_features(self, examples):
        """`DNN.features` of the dnn_step just run on `examples` (srgan.py:270-271), in the reference's shape.  The
        persistent coefficient kernel keeps DNN's activations on chip: there the features are recomputed with the
        updated network (one Adam update later)."""
        eng, net = self.engine, self.engine.d_net
        if self.persistent:
            return self.predict(examples, net='DNN')[1]
        self._wait_pending(('DNN',))
        t = eng._buf[('dnn', ('D', 'a', net.feature_buf if net.graph is not None else len(net.layers)))]
        B, F = examples.shape[0], net.feature_size
        eng.ops.begin()
        out = self.features_nchw(t[:B * F], B)
        return out.view(B, F, 1, 1) if net.family == 'crowd' else out

    def gradient_norm(self):
        """Per-sample ||d s / d x_hat||_2 of the last gan_step (srgan.py:371-372), [B] fp32."""
        self._wait_pending()
        if self.persistent:
            pub = getattr(self, '_coef_pub', None)
            if pub is not None:
                return pub[40 * (pub.numel() // 41):].clone()
            raise RuntimeError('gradient_norm(): set engine.publish_features before the step (persistent coefficient kernel)')
        g = self.engine._buf.get(('gan', 'gnorm'))
        return None if g is None else g.clone()

    # ---- forward-only helpers (D(x) / G(z) of the reference modules, on the kernels)
    def predict(self, x, net='D'):
        self._wait_pending()
        self._fresh_layouts()
        st = self.engine.D if net == 'D' else self.engine.DNN
        pred, feats = self.engine.d_features(x, st)
        return pred.clone(), self.features_nchw(feats, x.shape[0])

    def predict_crowd(self, x, net='D'):
        """KnnDenseNetCat.forward's return values (crowd/models.py:1138-1166) on the kernels: (density = None for the zeros
        tensor of :1153, count [B], map [B, 3, L, L]) -- the `network(images)` callable of crowd_data.predict_full_example /
        evaluation_epoch (crowd/srgan.py:256-259,359)."""
        if self.engine.d_net.family != 'crowd':
            raise ValueError('predict_crowd() is the crowd application\'s forward')
        self._wait_pending()
        self._fresh_layouts()
        st = self.engine.D if net == 'D' else self.engine.DNN
        pred, _, maps = self.engine.d_features(x, st, with_maps=True)
        L = self.engine.d_net.label_size
        return None, pred.clone(), torch.stack([m.reshape(x.shape[0], L, L) for m in maps], dim=1).to(torch.float32)

    def features_nchw(self, feats_flat, B):
        """`.features` in the reference's order: out.view(B, -1) of an NCHW tensor (age/models.py:74)."""
        c, h, w = self.engine.d_net.feature_chw
        out = torch.empty(B, c * h * w, device=self.device, dtype=torch.float32)
        self.engine.ops.nhwc_to_nchw(feats_flat, out, B, c, h, w)
        return out

    def generate(self, z):
        self._wait_pending()
        self._fresh_layouts()
        gnet = self.engine.g_net
        out_l = gnet.layers[-1]
        flat = self.engine.g_generate(z)
        B = z.shape[0]
        if gnet.family == 'coefficient':
            out = torch.empty(B, out_l.out_ch, device=self.device, dtype=torch.float32)
            self.engine.ops.nhwc_to_nchw(flat, out, B, out_l.out_ch, 1, 1)
            return out
        g = out_l.geom
        out = torch.empty(B, g.Cb, g.Hl, g.Wl, device=self.device, dtype=torch.float32)
        self.engine.ops.nhwc_to_nchw(flat, out, B, g.Cb, g.Hl, g.Wl)
        return out

    # ---- torch.optim.Adam state compatibility (checkpoints stay loadable)
    def export_optimizer_state(self, optimizer, which: str):
        self._wait_pending()
        st = {'D': self.engine.D, 'G': self.engine.G, 'DNN': self.engine.DNN}[which]
        for name, p in self.modules[which].named_parameters():
            if name not in st.slices:
                continue
            optimizer.state[p] = {'step': torch.tensor(float(st.adam_state[0].item())),
                                  'exp_avg': st.m(name).view_as(p).clone(),
                                  'exp_avg_sq': st.v(name).view_as(p).clone()}

    def import_optimizer_state(self, optimizer, which: str):
        st = {'D': self.engine.D, 'G': self.engine.G, 'DNN': self.engine.DNN}[which]
        for name, p in self.modules[which].named_parameters():
            s = optimizer.state.get(p)
            if s and 'exp_avg' in s and name in st.slices:
                st.m(name).copy_(s['exp_avg'].reshape(-1))
                st.v(name).copy_(s['exp_avg_sq'].reshape(-1))
                st.adam_state[0] = float(s['step'])

    def refresh_weights(self):
        """Call after the nn.Parameters were changed from outside (load_state_dict)."""
        for st in (self.engine.D, self.engine.G, self.engine.DNN):
            self.engine.repack(st)
        self._layouts_stale = False


# ------------------------------------------------------------------------------------------------ drop-in mix-in
class B200StepMixin:
    """Mix into a reference Experiment subclass; see module docstring."""
    b200_precision: Optional[str] = None
    b200_comm = None
    # srgan.py:332-386 leave the step's feature tensors and gradient norms on the Experiment; keep doing so (a 4B x F copy
    # inside the step + one layout kernel per attribute).  Set False to skip the materialisation when nothing reads them.
    b200_publish_features = True

    def _b200_method(self):
        """DG-GAN = the experiment overrides the loss hooks the way coefficient/dggan.py:22-64 / crowd/dggan.py:17-49 do;
        recognised by the discriminator's second head output (DgganMLP, KnnDenseNetCatDggan), cross-checked against the
        class hierarchy so that a mismatched pair raises instead of training the wrong loss."""
        outputs = nets.describe_module(self.D).head_outputs
        names = [c.__name__.lower() for c in type(self).__mro__]
        # SGAN = a subclass of sgan.py's SganExperiment with a K-logit discriminator (age/sgan.py:16-20, coefficient/sgan.py:17-21)
        by_class = 'dggan' if any('dggan' in n for n in names) else 'sgan' if any('sgan' in n for n in names) else 'srgan'
        by_module = 'dggan' if outputs == 2 else 'sgan' if outputs > 2 else 'srgan'
        if by_module != by_class:
            raise ValueError(f'{type(self).__name__}: discriminator has {outputs} head output(s) ({by_module}) but the '
                             f'experiment class is a {by_class} experiment')
        if by_class == 'sgan':
            self.settings.bins = tuple(float(v) for v in self.bins)
        return by_class

    def _b200_runner(self) -> StepRunner:
        r = getattr(self, '_b200', None)
        if r is None:
            r = StepRunner(self.D, self.G, self.DNN, self.settings, self._b200_method(),
                           precision=self.b200_precision, comm=self.b200_comm)
            r.import_optimizer_state(self.d_optimizer, 'D')
            r.import_optimizer_state(self.g_optimizer, 'G')
            r.import_optimizer_state(self.dnn_optimizer, 'DNN')
            r.engine.publish_features = bool(self.b200_publish_features)
            self._b200 = r
        return r

    def dnn_training_step(self, examples, labels, step):
        """srgan.py:259-271."""
        r = self._b200_runner()
        self.dnn_summary_writer.step = step
        group = self.dnn_optimizer.param_groups[0]                  # adjust_learning_rate (srgan.py:432-436) writes here
        r.dnn_step(examples, labels, lr=group['lr'], weight_decay=group['weight_decay'])
        if self.dnn_summary_writer.is_summary_step():
            self.dnn_summary_writer.add_scalar('Discriminator/Labeled Loss', r.scalars()['dnn_loss'])
            # srgan.py:270-271: only when the module sets `.features` (KnnDenseNetCatDggan and SganMLP do not)
            fam = r.engine.d_net.family
            if (r.method == 'srgan') or (r.method == 'dggan' and fam == 'coefficient') or (r.method == 'sgan' and fam == 'dcgan'):
                f = r.dnn_step_features(examples)
                self.DNN.features = f
                self.dnn_summary_writer.add_scalar('Feature Norm/Labeled', f.norm(dim=1).mean().item())

    def gan_training_step(self, labeled_examples, labels, unlabeled_examples, step):
        """srgan.py:273-320."""
        r = self._b200_runner()
        self.gan_summary_writer.step = step
        r.gan_step(labeled_examples, labels, unlabeled_examples, step, noise=getattr(self, '_b200_noise', None))
        group = self.d_optimizer.param_groups[0]
        if (group['lr'], group['weight_decay']) != (self.settings.learning_rate, self.settings.weight_decay):
            raise ValueError('the B200 step reads the D / G learning rate and weight decay from settings (srgan.py:131-138); '
                             'd_optimizer.param_groups was changed away from them')
        # side effects of srgan.py:332-386 (device tensors; per-rank shards under data parallelism)
        self.gradient_norm = r.gradient_norm()
        if r.engine.publish_features and r.method == 'srgan':
            self.labeled_features = r.step_features('labeled')
            self.unlabeled_features = r.step_features('unlabeled')
            self.fake_features = r.step_features('fake')
            self.interpolates_features = r.step_features('interpolates')
        # validation_summaries / save_models of the reference read the nn.Parameters on the caller's stream
        r._wait_pending()
        if self.gan_summary_writer.is_summary_step():
            s = r.scalars()
            w = self.gan_summary_writer
            if step % self.settings.generator_training_step_period == 0:
                w.add_scalar('Generator/Loss', s['generator_loss'])
            w.add_scalar('Discriminator/Labeled Loss', s['labeled_loss'])
            w.add_scalar('Discriminator/Unlabeled Loss', s['unlabeled_loss'])
            w.add_scalar('Discriminator/Fake Loss', s['fake_loss'])
            w.add_scalar('Discriminator/Gradient Penalty', s['gradient_penalty'])
            w.add_scalar('Discriminator/Gradient Norm', s['gradient_norm_mean'])
            if self.labeled_features is not None and self.unlabeled_features is not None:        # srgan.py:315-319
                w.add_scalar('Feature Norm/Labeled', self.labeled_features.mean(0).norm().item())
                w.add_scalar('Feature Norm/Unlabeled', self.unlabeled_features.mean(0).norm().item())

    def save_models(self, step):
        """srgan.py:88-97 with the Adam moments exported into the torch optimizers first.  `settings.async_checkpoint`
        (opt-in) hands the same dict to checkpoint.AsyncWriter: device -> pinned-host copies on a side stream, pickling and
        the disk write on a background thread; under data parallelism each rank then writes its round-robin shard of the
        tensors (checkpoint.shard).  `wait_for_checkpoints()` (also run before every later save) completes them."""
        r = getattr(self, '_b200', None)
        if r is not None:
            r.export_optimizer_state(self.d_optimizer, 'D')
            r.export_optimizer_state(self.g_optimizer, 'G')
            r.export_optimizer_state(self.dnn_optimizer, 'DNN')
        if not getattr(self.settings, 'async_checkpoint', False):
            return super().save_models(step)
        from . import checkpoint
        model = {'DNN': self.DNN.state_dict(), 'dnn_optimizer': self.dnn_optimizer.state_dict(),
                 'D': self.D.state_dict(), 'd_optimizer': self.d_optimizer.state_dict(),
                 'G': self.G.state_dict(), 'g_optimizer': self.g_optimizer.state_dict(), 'step': step}
        path = os.path.join(self.trial_directory, f'model_{step}.pth')
        comm = r.engine.comm if r is not None else None
        if comm is not None and comm.world_size > 1:
            model, path = (checkpoint.shard(model, comm.rank, comm.world_size),
                           checkpoint.shard_path(path, comm.rank, comm.world_size))
        if getattr(self, '_checkpoint_writer', None) is None:
            self._checkpoint_writer = checkpoint.AsyncWriter()
        self._checkpoint_writer.save(model, path)

    def wait_for_checkpoints(self):
        w = getattr(self, '_checkpoint_writer', None)
        if w is not None:
            w.wait()


class Experiment:
    """Stand-alone mirror of the reference Experiment's step API (srgan.py:24-50, 259-320) for the supported model
    families; `application` in {'coefficient', 'age', 'driving', 'crowd'}, `method` in {'srgan', 'dggan', 'sgan'} (sgan: the
    class-logit experiments of age/sgan.py and coefficient/sgan.py, with their bins)."""

    def __init__(self, settings: Settings, application='age', method='srgan', device='cuda:0', comm=None, **model_kwargs):
        self.settings = settings
        dev = torch.device(device)
        bins = int(getattr(settings, 'number_of_bins', 10))                 # settings.py:67
        if method == 'sgan':
            if application not in ('age', 'coefficient'):
                raise NotImplementedError(f'SGAN for {application!r}: the reference has age/sgan.py, coefficient/sgan.py and '
                                          'crowd/sgan.py; the first two have a B200 path')
            if getattr(settings, 'bins', None) is None:                     # age/sgan.py:14, coefficient/sgan.py:15
                lo, hi = (10.0, 95.0) if application == 'age' else (-3.0, 3.0)
                settings.bins = tuple(torch.linspace(lo, hi, bins).tolist())
        if application == 'coefficient':
            n_out = bins if method == 'sgan' else 2 if method == 'dggan' else 1
            hidden = 100 if method == 'sgan' else settings.hidden_size     # SganMLP: coefficient/models.py:80-83
            self.D = CoefficientMLP(hidden, n_out)
            self.DNN = CoefficientMLP(hidden, n_out)
            self.G = CoefficientGenerator(settings.hidden_size)
        elif application in ('age', 'driving'):
            self.G = DcganGenerator(**{k: v for k, v in model_kwargs.items() if k in ('z_dim', 'image_size', 'conv_dim')})
            dk = {k: v for k, v in model_kwargs.items() if k in ('image_size', 'conv_dim')}
            if method == 'sgan':
                dk['number_of_outputs'] = bins                              # age/sgan.py:18-19
            self.D = DcganDiscriminator(**dk)
            self.DNN = DcganDiscriminator(**dk)
        elif application == 'crowd':                                        # crowd/srgan.py:92-96 model_setup
            dk = {k: v for k, v in model_kwargs.items() if k in ('growth_rate', 'block_config', 'num_init_features', 'bn_size',
                                                                 'label_patch_size', 'image_size')}
            gk = {k: v for k, v in model_kwargs.items() if k in ('z_dim', 'conv_dim')}
            self.G = DcganGenerator(image_size=dk.get('image_size', 224), **gk)
            n_out = 2 if method == 'dggan' else 1                          # crowd/dggan.py:11-15
            self.D = KnnDenseNetCat(number_of_outputs=n_out, **dk)
            self.DNN = KnnDenseNetCat(number_of_outputs=n_out, **dk)
        else:
            raise NotImplementedError(f'application {application!r}: no B200 path yet')
        for m in (self.D, self.G, self.DNN):
            m.to(dev)
        self.runner = StepRunner(self.D, self.G, self.DNN, settings, method, comm=comm, device=dev)
        self.gradient_norm = None

    def dnn_training_step(self, examples, labels, step):
        lr = self.settings.learning_rate * (0.1 ** (step // 100000))       # srgan.py:432-436
        self.runner.dnn_step(examples, labels, lr=lr)

    def gan_training_step(self, labeled_examples, labels, unlabeled_examples, step, noise=None):
        self.runner.gan_step(labeled_examples, labels, unlabeled_examples, step, noise)
