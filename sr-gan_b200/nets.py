"""
Network descriptions for the hot path: which contraction each layer of the reference's G / D modules is, in the
"conv pair" vocabulary the kernels use.

A conv pair relates a SMALL side  S[n, Hs, Ws, a]  and a LARGE side  L[n, Hl, Wl, b]  through taps W[a, r, s, b]:
    down :  S[o]  = sum_{t,b} L[stride*o - pad + t, b] * W[a, t, b]      nn.Conv2d forward, ConvTranspose2d backward-data
    up   :  L[i] += sum_{a}   S[o, a]               * W[a, t, b],  i = stride*o - pad + t
                                                                         nn.ConvTranspose2d forward, Conv2d backward-data
    wgrad:  dW[a,t,b] = sum_{n,o} S[o,a] * L[stride*o - pad + t, b]
Both torch weight layouts, Conv2d [K, C, R, S] and ConvTranspose2d [Cin, Cout, R, S], are [a, b, r, s] in this
vocabulary, so one repack rule serves both.  nn.Linear [out, in] is the pair with R=S=Hs=Ws=Hl=Wl=1, a=out, b=in.

Reference modules covered (file:line): coefficient/models.py:12-72 (Generator, MLP, DgganMLP);
age/models.py:32-80 == driving/models.py (Generator, Discriminator); crowd/models.py:127-147 (DCGenerator, same
shape family as the age Generator with image_size 224).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

ACT_NONE, ACT_LEAKY, ACT_TANH = 0, 1, 2


@dataclass(frozen=True)
class Geom:
    Hs: int
    Ws: int
    Ca: int
    Hl: int
    Wl: int
    Cb: int
    R: int
    S: int
    stride: int
    pad: int

    @property
    def small_elems(self):
        return self.Hs * self.Ws * self.Ca

    @property
    def large_elems(self):
        return self.Hl * self.Wl * self.Cb

    @property
    def macs_per_sample(self):
        return self.Hs * self.Ws * self.Ca * self.R * self.S * self.Cb


@dataclass
class Layer:
    name: str                      # state_dict prefix, e.g. 'layer2.0' or 'linear1'
    fwd: str                       # 'down' | 'up'
    geom: Geom
    act: int
    slope: float
    # how the torch master weight [d0,d1,d2,d3] maps to (a, r, s, b):  Wd index = a*sa + r*sr + s*ss + b*sb etc. is
    # derived in engine.py from `perm`, the positions of (a, b, r, s) in the master dims.
    master_dims: Tuple[int, int, int, int] = (0, 0, 0, 0)
    # geometry of the master -> (a, r, s, b) relation; 'conv' = master[a][b][r][s]; 'fc_up' = master[b'][c][r][s] seen as
    # a Linear with a=(r,s,c), b=b' (ConvTranspose2d on a 1x1 input, age/models.py:37,47)
    master_kind: str = 'conv'
    bias_mod: int = 0              # bias index = column % bias_mod  (fc_up: bias per channel, broadcast over r,s)

    @property
    def thin_ok(self):
        """Image layer with <= 4 channels on the large side whose whole receptive field fits one 64-wide im2col row:
        eligible for the im2col/col2im + [pixels x 64] tensor-core GEMM lowering (engine.KPAD)."""
        g = self.geom
        return (g.Cb <= 4 and g.R * g.S * g.Cb <= 64 and g.Ca % 64 == 0 and g.Hl * g.Wl > 1
                and self.master_kind == 'conv')

    @property
    def in_elems(self):
        return self.geom.large_elems if self.fwd == 'down' else self.geom.small_elems

    @property
    def out_elems(self):
        return self.geom.small_elems if self.fwd == 'down' else self.geom.large_elems

    @property
    def in_rows(self):             # pixels per sample on the input side
        g = self.geom
        return g.Hl * g.Wl if self.fwd == 'down' else g.Hs * g.Ws

    @property
    def out_rows(self):
        g = self.geom
        return g.Hs * g.Ws if self.fwd == 'down' else g.Hl * g.Wl

    @property
    def in_ch(self):
        return self.geom.Cb if self.fwd == 'down' else self.geom.Ca

    @property
    def out_ch(self):
        return self.geom.Ca if self.fwd == 'down' else self.geom.Cb


@dataclass
class Net:
    kind: str                      # 'D' | 'G'
    family: str                    # 'coefficient' | 'dcgan'
    layers: List[Layer]
    head: Optional[str] = None     # state_dict prefix of the prediction head (D only)
    head_outputs: int = 0          # 1 (srgan) or 2 (dggan)
    head_master_kind: str = 'linear'   # 'linear' [out, F] | 'conv_full' [out, C, H, W] -> NHWC feature order
    input_chw: Tuple[int, int, int] = (0, 1, 1)    # (C, H, W) of the reference-side input
    feature_chw: Tuple[int, int, int] = (0, 1, 1)  # (C, H, W) of `.features` before flattening

    @property
    def feature_size(self):
        c, h, w = self.feature_chw
        return c * h * w

    def macs_per_sample(self):
        m = sum(l.geom.macs_per_sample for l in self.layers)
        if self.head:
            m += self.feature_size * self.head_outputs
        return m


def linear_geom(n_out, n_in):
    return Geom(1, 1, n_out, 1, 1, n_in, 1, 1, 1, 0)


def coefficient_d(hidden=10, n_in=50, dggan=False) -> Net:
    """coefficient/models.py:31-72."""
    sizes = [n_in, hidden, hidden, hidden]
    layers = [Layer(f'linear{i}', 'down', linear_geom(b, a), ACT_LEAKY, 0.01, (b, a, 1, 1))
              for i, (a, b) in enumerate(zip(sizes[:-1], sizes[1:]), 1)]
    return Net('D', 'coefficient', layers, head='linear4', head_outputs=2 if dggan else 1,
               input_chw=(n_in, 1, 1), feature_chw=(hidden, 1, 1))


def coefficient_g(hidden=10, z_dim=10, n_out=50) -> Net:
    """coefficient/models.py:12-28."""
    sizes = [z_dim, hidden, hidden, hidden, n_out]
    acts = [ACT_LEAKY, ACT_LEAKY, ACT_LEAKY, ACT_NONE]
    layers = [Layer(f'linear{i}', 'down', linear_geom(b, a), acts[i - 1], 0.01, (b, a, 1, 1))
              for i, (a, b) in enumerate(zip(sizes[:-1], sizes[1:]), 1)]
    return Net('G', 'coefficient', layers, input_chw=(z_dim, 1, 1))


def dcgan_d(image_size=128, conv_dim=64, n_out=1) -> Net:
    """age/models.py:55-80: 4 x (Conv k4 s2 p1 + leaky 0.05), features = flatten, head = Conv k=image/16."""
    ch = [3, conv_dim, conv_dim * 2, conv_dim * 4, conv_dim * 8]
    layers, h = [], image_size
    for i in range(1, 5):
        g = Geom(h // 2, h // 2, ch[i], h, h, ch[i - 1], 4, 4, 2, 1)
        layers.append(Layer(f'layer{i}.0', 'down', g, ACT_LEAKY, 0.05, (ch[i], ch[i - 1], 4, 4)))
        h //= 2
    return Net('D', 'dcgan', layers, head='layer5.0', head_outputs=n_out, head_master_kind='conv_full',
               input_chw=(3, image_size, image_size), feature_chw=(ch[4], h, h))


def dcgan_g(image_size=128, conv_dim=64, z_dim=256) -> Net:
    """age/models.py:32-52 / crowd/models.py:127-147: fc = ConvT k=image/16 on 1x1 (no activation), 3 x (ConvT k4 s2 p1 +
    leaky 0.05), ConvT + tanh."""
    k = image_size // 16
    ch = [conv_dim * 8, conv_dim * 4, conv_dim * 2, conv_dim, 3]
    layers = [Layer('fc.0', 'down', linear_geom(k * k * ch[0], z_dim), ACT_NONE, 0.0, (z_dim, ch[0], k, k),
                    master_kind='fc_up', bias_mod=ch[0])]
    h = k
    for i in range(1, 5):
        g = Geom(h, h, ch[i - 1], h * 2, h * 2, ch[i], 4, 4, 2, 1)
        act, slope = (ACT_LEAKY, 0.05) if i < 4 else (ACT_TANH, 0.0)
        layers.append(Layer(f'layer{i}.0', 'up', g, act, slope, (ch[i - 1], ch[i], 4, 4)))
        h *= 2
    return Net('G', 'dcgan', layers, input_chw=(z_dim, 1, 1))


def describe_module(module) -> Net:
    """Maps a reference nn.Module instance (or this package's mirrors) to its Net, by structure, not by import."""
    sd = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    if 'linear4.weight' in sd and 'linear1.weight' in sd:
        h, n_in = sd['linear1.weight']
        n_out = sd['linear4.weight'][0]
        if hasattr(module, 'input_size'):                       # coefficient Generator
            return coefficient_g(h, n_in, n_out)
        if n_out > 2:
            raise NotImplementedError('SganMLP (sgan.py) is outside the SR-GAN hot path (SURVEY 8f rank 3)')
        return coefficient_d(h, n_in, dggan=(n_out == 2))
    if 'fc.0.weight' in sd and 'layer4.0.weight' in sd:
        z_dim, c8, k, _ = sd['fc.0.weight']
        return dcgan_g(k * 16, c8 // 8, z_dim)
    if 'layer5.0.weight' in sd and 'layer1.0.weight' in sd and len(sd) == 10:
        n_out, c8, k, _ = sd['layer5.0.weight']
        return dcgan_d(k * 16, c8 // 8, n_out)
    raise NotImplementedError(
        f'{type(module).__name__}: no B200 path for this module yet (crowd KnnDenseNetCat is SURVEY 8a16, next round); '
        'refusing to fall back to the PyTorch path')
