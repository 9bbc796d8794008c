"""GPU: the flat-patch 3x3 convolution kernel (csrc/flat3x3.cu: one TMA load of a zero-padded image patch per 64 input channels,
nine shifted UMMA descriptors into it) behind srgan_conv_down, at every trunk resolution of the crowd DenseNet: whole-image
patches with several samples (7x7, 14x14), row bands of one image (28x28, 56x56), ragged sample counts, channel windows on
the output side and on the input side."""
import pytest
import torch

from tests.torch_ops import TorchOps
from srgan_b200.nets import Geom

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(scope='module')
def ops():
    from srgan_b200.ops_cuda import CudaOps
    return CudaOps()


def rnd(gen, *shape):
    return (torch.rand(*shape, generator=gen) * 2 - 1)


@pytest.mark.parametrize('hw,n,cin,valid,c0,pitch', [(56, 3, 128, 32, 64, 256), (28, 5, 128, 32, 480, 512), (14, 67, 128, 32, 1760, 1792),
                                                     (7, 130, 128, 32, 896, 1920), (14, 1, 64, 16, 0, 32), (9, 2, 192, 32, 32, 64),
                                                     (56, 1, 64, 8, 8, 16), (30, 2, 128, 24, 0, 24)])
def test_flat3x3_forward(ops, hw, n, cin, valid, c0, pitch):
    gen = torch.Generator().manual_seed(hw * 1000 + n)
    g = Geom(hw, hw, 64, hw, hw, cin, 3, 3, 1, 1)
    ref = TorchOps()
    pix = n * hw * hw
    L = rnd(gen, pix * cin).to(BF)
    Wd4 = rnd(gen, g.Ca, 9 * cin) * 0.1
    Wd4[valid:] = 0
    Wd = Wd4.reshape(-1).to(BF)
    vw = (pitch, valid, 0, 0)
    cat0 = rnd(gen, pix * pitch).to(BF)
    cat_ref, cat = cat0.clone(), cat0.clone().cuda()
    ref.conv_down(L, Wd, cat_ref[c0:], n, g, None, 0, None, 0, 0, 0.0, views=vw)
    before = ops.launches
    ops.conv_down(L.cuda(), Wd.cuda(), cat[c0:], n, g, None, 0, None, 0, 0, 0.0, views=vw)
    torch.cuda.synchronize()
    assert ops.lib.srgan_last_path_tensor() == 1 and ops.launches == before + 1
    a, b = cat.float().cpu().view(pix, pitch), cat_ref.float().view(pix, pitch)
    err = (a[:, c0:c0 + valid] - b[:, c0:c0 + valid]).abs().max().item() / (b[:, c0:c0 + valid].abs().max().item() + 1e-9)
    assert err < 1e-2, err
    mask = torch.ones(pix, pitch, dtype=torch.bool)
    mask[:, c0:c0 + valid] = False
    assert torch.equal(cat.cpu().view(pix, pitch)[mask], cat0.view(pix, pitch)[mask]), 'wrote outside the window'
