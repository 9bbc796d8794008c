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
shape family as the age Generator with image_size 224); crowd/models.py:1049-1166 (KnnDenseNetCat: DenseNet trunk with
eval-mode BatchNorm + three MapModules :763-786), described as a GRAPH of ops over named NHWC buffers (`Net.graph`).
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
    has_bias: bool = True          # DenseNet trunk convolutions are bias-free (crowd/models.py:341-345,1075)
    # 1x1 convolutions of graph nets run as a plain [pixels x Cb] GEMM: geom is then the Linear pair and the kernels see
    # gemm_rows x (number of samples) rows.  geom.Ca / geom.Cb may exceed the master dims (channel padding to the
    # tensor-core tile granularity; the pad rows / columns of the kernel-layout weights and of the buffers stay zero).
    gemm_rows: int = 1

    @property
    def macs_per_sample(self):
        """Algorithmic MACs (master dims; channel padding of the kernel operands is not counted)."""
        d0, d1, d2, d3 = self.master_dims
        if self.master_kind in ('conv', 'ct_gemm'):
            return d0 * d1 * d2 * d3 * self.geom.Hs * self.geom.Ws * self.gemm_rows
        return d0 * d1 * d2 * d3

    @property
    def thin_ok(self):
        """Image layer with <= 4 channels on the large side whose whole receptive field fits an im2col row of at most
        256 elements: eligible for the im2col/col2im + [pixels x kpad] tensor-core GEMM lowering (kpad = the receptive
        field rounded up to 64: 64 for the DCGAN k4 layers, 192 for the crowd stem k7)."""
        g = self.geom
        return (g.Cb <= 4 and g.R * g.S * g.Cb <= 256 and g.Ca % 64 == 0 and g.Hl * g.Wl > 1
                and self.master_kind == 'conv' and self.gemm_rows == 1)

    @property
    def kpad(self):
        g = self.geom
        return (g.R * g.S * g.Cb + 63) // 64 * 64

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
class Buf:
    """A named NHWC activation buffer of a graph net: `rows` pixels per sample, `ch` channels (= row pitch).  `act` is the
    activation its producer applied (the stored values are post-activation; deltas are w.r.t. the pre-activation, the
    consumer's backward applies act'); `accumulate`: several consumers add into its delta (concat buffers)."""
    name: str
    rows: int
    ch: int
    act: int = ACT_NONE
    slope: float = 0.0
    accumulate: bool = False


@dataclass
class Op:
    """One node of a graph net.  kind:
      'conv'    layer (dense src -> dense dst, the conv-pair kernels; C > 0: dst[:, c0:c0+C] is a channel window of a concat
                buffer, written / read in place through srgan_views);
      'affine'  eval-mode BatchNorm + ReLU: dst[:, :C] = relu(gamma*(src[:, c0:c0+C]-mean)/sqrt(var+eps)+beta), `name` = BN prefix;
      'copy'    dst[:, c0:c0+C] = src (dense C)        (concat write / feature slice write);
      'read'    dst (dense C)   = src[:, c0:c0+C]      (tap of a concat buffer);
      'maxpool' / 'avgpool'  src dense [H, W, C] -> dst[:, c0:c0+C] at [H', W'] (k, stride, pad);
      'shuffle' depth-to-space of a one-channel map: dst[(i*k+r), (j*k+s)] = src[(i, j), r*k+s] (H x W blocks of k x k)."""
    kind: str
    src: str
    dst: str
    layer: Optional[Layer] = None
    name: str = ''
    C: int = 0
    c0: int = 0
    H: int = 0
    W: int = 0
    k: int = 0
    stride: int = 0
    pad: int = 0
    branch: int = 0                # 0 = trunk; i > 0 = side branch i (crowd MapModule i): reads a trunk buffer, ends in `features`
    # BatchNorm fusion (bf16 tcgen05 path, csrc/bn_gemm.cu): an 'affine' op with fuse > 0 is carried out by the 1x1 'conv' op
    # that consumes its output (that op's `pre` points back at it).  fuse >= 1: the data gradient (srgan_bn_dgrad); >= 2: the
    # forward pass too (srgan_bn_conv_down: the normalised operand is stored only where a later pass reads it); >= 5: the
    # weight gradient too (srgan_bn_conv_wgrad: only the tangent pass of the interpolate rows still reads a stored operand).
    fuse: int = 0
    pre: Optional['Op'] = None
    # fuse >= 3: the BatchNorm + ReLU that FOLLOWS such a convolution (norm2 / relu2 of a dense layer) is carried out by the
    # convolution's epilogue in the forward pass: conv.post = that 'affine' op, whose `absorbed` flag is set
    post: Optional['Op'] = None
    absorbed: bool = False
    # fuse >= 4: the backward pass of the BatchNorm + ReLU in front of a dense layer's 3x3 convolution (norm2 / relu2) is carried
    # out by that convolution's data-gradient kernel (srgan_bn_conv_dgrad): conv.bwd_pre = the 'affine' op, whose bwd_fused is set
    bwd_pre: Optional['Op'] = None
    bwd_fused: bool = False


@dataclass
class Net:
    kind: str                      # 'D' | 'G'
    family: str                    # 'coefficient' | 'dcgan' | 'crowd'
    layers: List[Layer]
    head: Optional[str] = None     # state_dict prefix of the prediction head (D only)
    head_outputs: int = 0          # 1 (srgan), 2 (dggan) or the number of bins (sgan)
    head_master_kind: str = 'linear'   # 'linear' [out, F] | 'conv_full' [out, C, H, W] -> NHWC feature order
    input_chw: Tuple[int, int, int] = (0, 1, 1)    # (C, H, W) of the reference-side input
    feature_chw: Tuple[int, int, int] = (0, 1, 1)  # (C, H, W) of `.features` before flattening
    # graph nets (crowd): ops in topological order over named buffers; `layers` then lists the conv layers of the graph
    graph: Optional[List[Op]] = None
    bufs: Optional[Dict[str, Buf]] = None
    input_buf: str = ''
    feature_buf: str = ''
    # prediction head as a list of (state_dict prefix, number of feature columns): the crowd count is the sum of four
    # 20-column heads (crowd/models.py:1153-1162); chain nets have the single `head`
    head_parts: Optional[List[Tuple[str, int]]] = None
    map_bufs: Tuple[str, ...] = ()                 # crowd: the three predicted maps (label_size^2 x 1 each)
    label_size: int = 0

    @property
    def feature_size(self):
        c, h, w = self.feature_chw
        return c * h * w

    @property
    def feature_act(self):
        """(activation, slope) of the feature layer (its stored values are post-activation)."""
        if self.graph is not None:
            b = self.bufs[self.feature_buf]
            return b.act, b.slope
        return self.layers[-1].act, self.layers[-1].slope

    @property
    def in_elems(self):
        c, h, w = self.input_chw
        return c * h * w

    @property
    def affines(self):
        return [op for op in (self.graph or []) if op.kind == 'affine']

    def branch_taps(self):
        """Trunk buffers read by side-branch ops (crowd: cat2..cat4, tapped by the MapModules)."""
        trunk_made = {op.dst for op in self.graph if not op.branch} | {self.input_buf}
        return {op.src for op in self.graph if op.branch and op.src in trunk_made}

    def parts(self):
        if self.head_parts is not None:
            return self.head_parts
        return [(self.head, self.feature_size)] if self.head else []

    def macs_per_sample(self):
        m = sum(l.macs_per_sample for l in self.layers)
        if self.head:
            m += self.feature_size * self.head_outputs
        return m


def linear_geom(n_out, n_in):
    return Geom(1, 1, n_out, 1, 1, n_in, 1, 1, 1, 0)


def coefficient_d(hidden=10, n_in=50, dggan=False, n_out=None) -> Net:
    """coefficient/models.py:31-72; n_out = number_of_bins: SganMLP (:75-93, hidden 100)."""
    sizes = [n_in, hidden, hidden, hidden]
    layers = [Layer(f'linear{i}', 'down', linear_geom(b, a), ACT_LEAKY, 0.01, (b, a, 1, 1))
              for i, (a, b) in enumerate(zip(sizes[:-1], sizes[1:]), 1)]
    return Net('D', 'coefficient', layers, head='linear4', head_outputs=n_out or (2 if dggan else 1),
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


def knn_densenet_cat(block_config=(6, 12, 48, 32), growth_rate=32, num_init_features=64, bn_size=4, image_size=224,
                     label_size=224, pad_to=64, n_out=1, direct_concat=False, fuse_bn=0) -> Net:
    """crowd/models.py:1049-1166 KnnDenseNetCat as a graph.  Buffers: 'x' input; 'c0','n0' stem; 'cat{i}' the in-place
    concat buffer of dense block i (the stem pool / transition pool write its first channels, every dense layer appends
    growth_rate channels); per dense layer 'n1','b','n2','new'; per transition 'tn','tc'; per MapModule 't','map','m1'..
    'm3','h'; 'n5','fp','fcf'; 'features' [1 x 80].  Spatial sizes follow torch: stem conv k7 s2 p3, max-pool k3 s2 p1,
    transitions avg-pool 2.  The operand buffers of the trunk convolutions ('n1', 'new', ...) are padded to multiples of
    `pad_to` channels so that every trunk contraction is tcgen05-eligible (K and N multiples of 64); 1x1 convolutions are
    declared as [pixels x C] GEMMs (Layer.gemm_rows).  direct_concat: the 3x3 convolution of a dense layer writes its
    growth_rate channels straight into their window of the concat buffer and its gradients read that window of the concat
    delta (Op.C / Op.c0 on the conv op, srgan_views in the kernels): no 'new' buffer, no slice copies.  fuse_bn > 0: norm1 /
    relu1 of every dense layer and norm / relu of every transition are carried out by the 1x1 convolution that follows them
    (Op.fuse / Op.pre)."""
    g, bs = growth_rate, bn_size

    def pad(c):
        return (c + pad_to - 1) // pad_to * pad_to
    bufs: Dict[str, Buf] = {}
    ops: List[Op] = []
    layers: List[Layer] = []
    RELU = dict(act=ACT_LEAKY, slope=0.0)
    LK = dict(act=ACT_LEAKY, slope=0.01)

    def buf(name, rows, ch, **kw):
        bufs[name] = Buf(name, rows, ch, **kw)
        return name

    def conv(name, src, dst, geom, fwd, master, act=ACT_NONE, slope=0.0, bias=False, gemm_rows=1):
        l = Layer(name, fwd, geom, act, slope, master, has_bias=bias, gemm_rows=gemm_rows)
        layers.append(l)
        ops.append(Op('conv', src, dst, layer=l))

    H0 = image_size
    H1 = (H0 + 2 * 3 - 7) // 2 + 1            # stem conv
    H2 = (H1 + 2 * 1 - 3) // 2 + 1            # max-pool
    buf('x', H0 * H0, 3)
    buf('c0', H1 * H1, num_init_features)
    conv('conv_layer1.conv0', 'x', 'c0', Geom(H1, H1, num_init_features, H0, H0, 3, 7, 7, 2, 3), 'down',
         (num_init_features, 3, 7, 7))
    buf('n0', H1 * H1, num_init_features, **RELU)
    ops.append(Op('affine', 'c0', 'n0', name='conv_layer1.norm0', C=num_init_features))
    c, h = num_init_features, H2
    taps = []
    for bi, n_layers in enumerate(block_config, 1):
        ctot = c + n_layers * g
        cat = buf(f'cat{bi}', h * h, ctot, accumulate=True)
        if bi == 1:
            ops.append(Op('maxpool', 'n0', cat, C=c, c0=0, H=H1, W=H1, k=3, stride=2, pad=1))
        else:
            ops.append(Op('avgpool', f'tc{bi - 1}', cat, C=c, c0=0, H=2 * h, W=2 * h, k=2, stride=2))
            taps.append((cat, c, h))
        for li in range(1, n_layers + 1):
            pre = f'dense_blocks.denseblock{bi}.denselayer{li}'
            tag = f'{bi}.{li}'
            cb = pad(bs * g)
            buf('n1.' + tag, h * h, pad(c), **RELU)
            ops.append(Op('affine', cat, 'n1.' + tag, name=pre + '.norm1', C=c, c0=0))
            buf('b.' + tag, h * h, cb)
            conv(pre + '.conv1', 'n1.' + tag, 'b.' + tag, linear_geom(cb, pad(c)), 'down', (bs * g, c, 1, 1), gemm_rows=h * h)
            if fuse_bn:
                ops[-1].pre, ops[-2].fuse = ops[-2], fuse_bn
            buf('n2.' + tag, h * h, cb, **RELU)
            ops.append(Op('affine', 'b.' + tag, 'n2.' + tag, name=pre + '.norm2', C=bs * g))
            if fuse_bn >= 3:
                ops[-2].post, ops[-1].absorbed = ops[-1], True
            if direct_concat:
                conv(pre + '.conv2', 'n2.' + tag, cat, Geom(h, h, pad(g), h, h, cb, 3, 3, 1, 1), 'down', (g, bs * g, 3, 3))
                ops[-1].C, ops[-1].c0 = g, c
                if fuse_bn >= 4:
                    ops[-1].bwd_pre, ops[-2].bwd_fused = ops[-2], True
            else:
                buf('new.' + tag, h * h, pad(g))
                conv(pre + '.conv2', 'n2.' + tag, 'new.' + tag, Geom(h, h, pad(g), h, h, cb, 3, 3, 1, 1), 'down', (g, bs * g, 3, 3))
                ops.append(Op('copy', 'new.' + tag, cat, C=g, c0=c))
            c += g
        if bi != len(block_config):
            pre = f'transition_layers.transition{bi}'
            buf(f'tn{bi}', h * h, pad(c), **RELU)
            ops.append(Op('affine', cat, f'tn{bi}', name=pre + '.norm', C=c, c0=0))
            buf(f'tc{bi}', h * h, pad(c // 2))
            conv(pre + '.conv', f'tn{bi}', f'tc{bi}', linear_geom(pad(c // 2), pad(c)), 'down', (c // 2, c, 1, 1), gemm_rows=h * h)
            if fuse_bn:
                ops[-1].pre, ops[-2].fuse = ops[-2], fuse_bn
            c //= 2
            if h % 2:
                raise ValueError('transition input must have an even extent')
            h //= 2
    buf('n5', h * h, c, **RELU)
    ops.append(Op('affine', f'cat{len(block_config)}', 'n5', name='norm5', C=c, c0=0))
    buf('fp', 1, c)                                 # global average of the ReLU outputs: a linear op, no activation of its own
    ops.append(Op('avgpool', 'n5', 'fp', C=c, c0=0, H=h, W=h, k=h, stride=h))
    buf('fcf', 1, 20, **LK)
    conv('final_count_feature_layer', 'fp', 'fcf', Geom(1, 1, 20, 1, 1, c, 1, 1, 1, 0), 'down', (20, c, 1, 1), bias=True, **LK)
    F = 80
    buf('features', 1, F, **LK)
    L = label_size
    maps = []
    for i, (cat, ci, hi) in enumerate(taps[:3], 1):
        k = L // hi
        if k * hi != L or L % 8:
            raise ValueError('label_size must be a multiple of the tap extents and of 8')
        pre = f'map_module{i}'
        first_op = len(ops)
        buf(f't{i}', hi * hi, ci)
        ops.append(Op('read', cat, f't{i}', C=ci, c0=0))
        # ConvTranspose2d(ci, 1, k, stride k): a [pixels x ci] x [ci x k*k] GEMM (one scalar bias) + depth-to-space
        buf(f'mapb{i}', hi * hi, k * k, **LK)
        l = Layer(pre + '.map_transposed_conv_layer', 'down', linear_geom(k * k, ci), ACT_LEAKY, 0.01, (ci, 1, k, k),
                  master_kind='ct_gemm', bias_mod=1, has_bias=True, gemm_rows=hi * hi)
        layers.append(l)
        ops.append(Op('conv', f't{i}', f'mapb{i}', layer=l))
        buf(f'map{i}', L * L, 1, **LK)
        ops.append(Op('shuffle', f'mapb{i}', f'map{i}', H=hi, W=hi, k=k))
        maps.append(f'map{i}')
        src, hh, cin = f'map{i}', L, 1
        for j, cout in enumerate((8, 16, 32), 1):
            buf(f'm{j}.{i}', (hh // 2) ** 2, cout, **LK)
            conv(f'{pre}.conv{j}', src, f'm{j}.{i}', Geom(hh // 2, hh // 2, cout, hh, hh, cin, 2, 2, 2, 0), 'down',
                 (cout, cin, 2, 2), bias=True, **LK)
            src, hh, cin = f'm{j}.{i}', hh // 2, cout
        buf(f'h{i}', 1, 20, **LK)
        conv(pre + '.linear1', src, f'h{i}', Geom(1, 1, 20, hh, hh, 32, hh, hh, 1, 0), 'down', (20, 32, hh, hh), bias=True, **LK)
        ops.append(Op('copy', f'h{i}', 'features', C=20, c0=20 * (i - 1)))
        for o in ops[first_op:]:          # an independent chain hanging off cat{i+1}: the engine may run it on its own stream
            o.branch = i
    ops.append(Op('copy', 'fcf', 'features', C=20, c0=60))
    # (ops are in topological order: the trunk, then the count head, then the three MapModules, then the feature copies)
    head_parts = [(f'map_module{i}.count_layer', 20) for i in (1, 2, 3)] + [('count_layer', 20)]
    # n_out = 2: KnnDenseNetCatDggan (crowd/models.py:929-1046), every count layer has a second (DG-GAN score) output
    return Net('D', 'crowd', layers, head='count_layer', head_outputs=n_out, head_master_kind='parts',
               input_chw=(3, image_size, image_size), feature_chw=(F, 1, 1), graph=ops, bufs=bufs, input_buf='x',
               feature_buf='features', head_parts=head_parts, map_bufs=tuple(maps), label_size=L)


def describe_module(module, direct_concat=False, fuse_bn=0) -> Net:
    """Maps a reference nn.Module instance (or this package's mirrors) to its Net, by structure, not by import."""
    sd = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    if 'linear4.weight' in sd and 'linear1.weight' in sd:
        h, n_in = sd['linear1.weight']
        n_out = sd['linear4.weight'][0]
        if hasattr(module, 'input_size'):                       # coefficient Generator
            return coefficient_g(h, n_in, n_out)
        return coefficient_d(h, n_in, dggan=(n_out == 2), n_out=n_out)
    if 'conv_layer1.conv0.weight' in sd and 'map_module3.linear1.weight' in sd and 'final_count_feature_layer.weight' in sd:
        import re
        blocks = {}
        for k in sd:
            m = re.match(r'dense_blocks\.denseblock(\d+)\.denselayer(\d+)\.conv1\.weight', k)
            if m:
                blocks[int(m.group(1))] = max(blocks.get(int(m.group(1)), 0), int(m.group(2)))
        cfg = tuple(blocks[i] for i in sorted(blocks))
        w1, w2 = sd['dense_blocks.denseblock1.denselayer1.conv1.weight'], sd['dense_blocks.denseblock1.denselayer1.conv2.weight']
        growth, init = w2[0], sd['conv_layer1.conv0.weight'][0]
        bn_size = w1[0] // growth
        label = sd['map_module1.linear1.weight'][2] * 8
        k1 = sd['map_module1.map_transposed_conv_layer.weight'][2]
        image = (label // k1) * 8
        return knn_densenet_cat(cfg, growth, init, bn_size, image, label, n_out=sd['count_layer.weight'][0],
                                direct_concat=direct_concat and growth % 8 == 0 and init % 8 == 0,
                                fuse_bn=fuse_bn if (growth % 8 == 0 and init % 8 == 0) else 0)
    if 'fc.0.weight' in sd and 'layer4.0.weight' in sd:
        z_dim, c8, k, _ = sd['fc.0.weight']
        return dcgan_g(k * 16, c8 // 8, z_dim)
    if 'layer5.0.weight' in sd and 'layer1.0.weight' in sd and len(sd) == 10:
        n_out, c8, k, _ = sd['layer5.0.weight']
        return dcgan_d(k * 16, c8 // 8, n_out)
    raise NotImplementedError(
        f'{type(module).__name__}: no B200 path for this module; '
        'refusing to fall back to the PyTorch path')
