"""The drop-in seam against the UNMODIFIED reference classes (build container only: /root/reference is not on the GPU box).
`class Fast(B200StepMixin, <reference Experiment subclass>)` must construct through the reference's own setup path, keep the
reference's modules / optimizers, map them onto the kernel-side net descriptions, and -- with no CUDA device -- refuse to
step (RuntimeError) instead of falling back to the PyTorch path."""
import pytest
import torch

from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason='reference checkout not present')


def _settings(**kw):
    from settings import Settings
    s = Settings()
    s.batch_size = 4
    for k, v in kw.items():
        setattr(s, k, v)
    return s


@pytest.mark.parametrize('app', ['coefficient', 'coefficient_dggan', 'age', 'crowd'])
def test_mixin_composes_with_reference_experiments(app):
    ref_harness.install_shims()
    import srgan_b200
    from srgan_b200 import nets
    if app == 'coefficient':
        from coefficient.srgan import CoefficientExperiment as Base
        exp = ref_harness.make_experiment(type('Fast', (srgan_b200.B200StepMixin, Base), {}), _settings())
        want = ('coefficient', 1)
    elif app == 'coefficient_dggan':
        from coefficient.dggan import CoefficientDgganExperiment as Base
        exp = ref_harness.make_experiment(type('FastDggan', (srgan_b200.B200StepMixin, Base), {}), _settings())
        want = ('coefficient', 2)
    elif app == 'age':
        from age.srgan import AgeExperiment as Base
        exp = ref_harness.make_experiment(type('Fast', (srgan_b200.B200StepMixin, Base), {}), _settings())
        want = ('dcgan', 1)
    else:
        from crowd.srgan import CrowdExperiment as Base
        from crowd.models import KnnDenseNetCat, DCGenerator
        exp = ref_harness.make_experiment(type('Fast', (srgan_b200.B200StepMixin, Base), {}), _settings(),
                                          D=KnnDenseNetCat(pretrained=False), G=DCGenerator(), DNN=KnnDenseNetCat(pretrained=False))
        want = ('crowd', 1)
    d_net, g_net = nets.describe_module(exp.D), nets.describe_module(exp.G)
    assert (d_net.family, d_net.head_outputs) == want
    assert nets.describe_module(exp.DNN) == d_net
    assert g_net.kind == 'G'
    # every trainable parameter of the reference module is owned by the kernel-side description
    keys = {l.name + '.weight' for l in d_net.layers} | {l.name + '.bias' for l in d_net.layers if l.has_bias}
    keys |= {op.name + s for op in d_net.affines for s in ('.weight', '.bias')}
    keys |= {h + s for h, _ in d_net.parts() for s in ('.weight', '.bias')}
    assert keys == {k for k, _ in exp.D.named_parameters()}
    assert exp._b200_method() == ('dggan' if 'dggan' in app else 'srgan')
    if not torch.cuda.is_available():
        x = torch.zeros(4, *( (50,) if 'coefficient' in app else (3, 128, 128) if app == 'age' else (3, 224, 224)))
        y = torch.zeros(4) if app != 'crowd' else (torch.zeros(4, 224, 224), torch.zeros(4, 224, 224))
        with pytest.raises(RuntimeError):
            exp.dnn_training_step(x, y, 0)           # no CUDA device: the B200 step refuses, it never falls back
