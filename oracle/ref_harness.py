"""
Harness that imports the UNMODIFIED reference  --  TEST / BENCH INFRASTRUCTURE, never product code.

The reference tree is read from /root/reference when it is mounted (build container) and otherwise from the copy
oracle/stage_reference.py staged under baseline/_ref/ (git-ignored, travels to the GPU box like the built .so).  Users:
oracle/make_golden.py (fixtures under tests/golden/), bench.py's reference arm / cpu_baseline / gpu_baseline legs (the
reference's own step, timed), tests/test_gpu_mixin_reference.py (B200StepMixin composed with the reference classes).

The reference needs eight non-numeric third-party modules that are not installed; they are
replaced by empty stand-ins in sys.modules so that no reference file has to be edited.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np
import torch

def _find_root():
    env = os.environ.get('SRGAN_REFERENCE_ROOT')
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in ([env] if env else []) + ['/root/reference', os.path.join(here, 'baseline', '_ref')]:
        if os.path.isfile(os.path.join(p, 'srgan.py')):
            return p
    return env or '/root/reference'


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'srgan.py'))


def set_reference_device(device):
    """The reference binds `gpu = cuda:0 if available else cpu` at import (utility.py:18) and copies the name into every
    module that uses it; the CPU arm on a GPU box re-points those module globals (no file is edited)."""
    dev = torch.device(device)
    import utility
    utility.gpu = dev
    for name in ('srgan', 'dnn', 'sgan', 'coefficient.srgan', 'coefficient.dggan', 'coefficient.sgan', 'age.srgan', 'age.sgan',
                 'driving.srgan', 'crowd.srgan'):
        m = sys.modules.get(name)
        if m is not None and hasattr(m, 'gpu'):
            m.gpu = dev


class _RecordingWriter:
    """Stand-in for tensorboardX.SummaryWriter: records scalars so the harness can read the step's losses."""
    def __init__(self, log_dir=None, comment='', **kwargs):
        self.scalars = {}

    def add_scalar(self, tag, value, global_step=None, **kwargs):
        self.scalars.setdefault(tag, []).append((global_step, float(value)))

    def add_histogram(self, *a, **k):
        pass

    def add_image(self, *a, **k):
        pass


def install_shims():
    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m
    mod('imageio')
    mpl = mod('matplotlib')
    mpl.cm = mod('matplotlib.cm')
    mpl.pyplot = mod('matplotlib.pyplot', switch_backend=lambda *a, **k: None)
    mpl.use = lambda *a, **k: None
    mod('tensorboardX', SummaryWriter=_RecordingWriter)
    mod('recordclass', RecordClass=type('RecordClass', (), {'__init__': lambda self, **kw: self.__dict__.update(kw)}))
    mod('seaborn', set=lambda *a, **k: None, set_style=lambda *a, **k: None)
    sk = mod('skimage')
    sk.transform = mod('skimage.transform')
    sk.color = mod('skimage.color')
    mod('patoolib')
    mod('mtcnn')
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


class _FixedMixture:
    """Replaces utility.MixtureModel inside srgan.py so that `z` (srgan.py:286-289) is the injected array."""
    next_z = None

    def __init__(self, submodels, *a, **k):
        pass

    def rvs(self, size):
        z = _FixedMixture.next_z
        assert list(z.shape) == list(size), (z.shape, size)
        return np.asarray(z, dtype=np.float64)


@contextlib.contextmanager
def injected_noise(z, alpha, z2):
    """Makes the three random draws of gan_training_step return the given tensors: z via MixtureModel.rvs
    (srgan.py:286-289), alpha via torch.rand (:364), z2 via torch.randn (:301)."""
    import srgan
    old_mm, old_rand, old_randn = srgan.MixtureModel, torch.rand, torch.randn
    _FixedMixture.next_z = z.detach().cpu().numpy()
    srgan.MixtureModel = _FixedMixture
    torch.rand = lambda *a, **k: alpha.clone()
    torch.randn = lambda *a, **k: z2.clone()
    try:
        yield
    finally:
        srgan.MixtureModel, torch.rand, torch.randn = old_mm, old_rand, old_randn


def make_experiment(cls, settings, D=None, G=None, DNN=None):
    """Builds a reference Experiment the way Experiment.train does (srgan.py:72-80) minus datasets / checkpoint IO."""
    exp = cls(settings)
    exp.trial_directory = '/tmp/srgan_oracle_trial'
    exp.prepare_summary_writers()
    if D is None:
        exp.model_setup()
    else:
        exp.D, exp.G, exp.DNN = D, G, DNN
    exp.prepare_optimizers()
    exp.gpu_mode()
    exp.train_mode()
    return exp


def last_scalars(exp):
    """The scalars the step wrote (tags of srgan.py:268-270, 306-319)."""
    out = {}
    for prefix, w in (('dnn', exp.dnn_summary_writer), ('gan', exp.gan_summary_writer)):
        for tag, vals in w.scalars.items():
            out[f'{prefix}/{tag}'] = vals[-1][1]
    return out


def workload_experiment(name, settings_kwargs, device='cpu', state=None, method='srgan', base=None):
    """A reference Experiment of a bench workload ('coefficient' | 'age' | 'driving' | 'crowd') on `device`, built the way
    Experiment.train does, with the reference's own modules.  state = an oracle OracleState whose parameter dicts are
    loaded (strict) into the reference modules: the crowd trunk cannot download its pretrained weights here
    (crowd/models.py:1103-1127), and tests want identical initial parameters on both sides.  base = extra base classes
    placed in front of the reference Experiment subclass (tests: B200StepMixin)."""
    install_shims()
    import utility                                   # noqa: F401  (first import decides utility.gpu)
    from settings import Settings
    s = Settings()
    for k, v in settings_kwargs.items():
        setattr(s, k, v)
    if name == 'coefficient':
        from coefficient.srgan import CoefficientExperiment
        from coefficient.dggan import CoefficientDgganExperiment
        from coefficient.models import Generator, MLP, DgganMLP
        cls = CoefficientDgganExperiment if method == 'dggan' else CoefficientExperiment
        mk = DgganMLP if method == 'dggan' else MLP
        if method == 'sgan':                                    # coefficient/sgan.py:11-21
            from coefficient.sgan import CoefficientSganExperiment as cls
            from coefficient.models import SganMLP as mk
        D, DNN, G = mk(), mk(), Generator()
    elif name in ('age', 'driving'):
        if name == 'age':
            from age.srgan import AgeExperiment as cls
            from age.models import Generator, Discriminator
        else:
            from driving.srgan import DrivingExperiment as cls
            from driving.models import Generator, Discriminator
        if method == 'sgan':                                    # age/sgan.py:10-20
            from age.sgan import AgeSganExperiment as cls
        if state is not None:
            z_dim, c8, k, _ = state.G['fc.0.weight'].shape
            n_out = state.D['layer5.0.weight'].shape[0]
            D, DNN = (Discriminator(image_size=k * 16, conv_dim=c8 // 8, number_of_outputs=n_out) for _ in range(2))
            G = Generator(z_dim=z_dim, image_size=k * 16, conv_dim=c8 // 8)
        else:
            D, DNN, G = Discriminator(), Discriminator(), Generator()
    elif name == 'crowd':
        from crowd.srgan import CrowdExperiment
        from crowd.dggan import CrowdDgganExperiment
        from crowd.models import KnnDenseNetCat, KnnDenseNetCatDggan, DCGenerator
        cls = CrowdDgganExperiment if method == 'dggan' else CrowdExperiment
        mk = KnnDenseNetCatDggan if method == 'dggan' else KnnDenseNetCat
        D, DNN, G = mk(pretrained=False), mk(pretrained=False), DCGenerator()
    else:
        raise ValueError(name)
    if state is not None:
        D.load_state_dict(state.D, strict=True)
        DNN.load_state_dict(state.DNN, strict=True)
        G.load_state_dict(state.G, strict=True)
    set_reference_device(device)
    if base:
        cls = type('B200' + cls.__name__, tuple(base) + (cls,), {})
    return make_experiment(cls, s, D=D, G=G, DNN=DNN)
