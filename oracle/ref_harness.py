"""
Harness that imports the UNMODIFIED reference from /root/reference  --  TEST INFRASTRUCTURE, build container only.

/root/reference does not exist on the GPU box, so nothing under tests/ (-m gpu), smoke() or bench.py imports this
module; only oracle/make_golden.py (run here, output committed under tests/golden/) and the optional
tests/test_oracle_vs_reference.py (skipped when the reference is absent) do.

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

REFERENCE_ROOT = os.environ.get('SRGAN_REFERENCE_ROOT', '/root/reference')


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'srgan.py'))


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
