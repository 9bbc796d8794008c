"""
Explicit (autograd-free) schedule of the SR-GAN / DG-GAN training step on top of the C-ABI kernels.

Replaces, for the supported model families, the bodies of
    Experiment.dnn_training_step   srgan.py:259-271
    Experiment.gan_training_step   srgan.py:273-320   (+ the *_loss_calculation hooks :322-391, feature_distance_loss
                                                        :438-449, gradient_penalty_calculation :360-375)
    coefficient/dggan.py:22-64     (DG-GAN loss overrides)
with the de-duplicated schedule of SURVEY.md App. B / C:
    D step : G fwd (no grad) -> D fwd on the row-concatenated batch [x; u; fake; x_hat] -> feature column sums
             (-> all-reduce across ranks) -> distance losses + seeds -> gradient-penalty chains on the x_hat rows
             (g-chain = 'up' ops, tangent u-chain = 'down' ops, App. C.3) -> ONE backward over [x; u; fake; x_hat]
             with the weight-gradient launches also covering the (u_{l-1}, gamma_l) tangent block -> Adam.
    G step : G fwd -> D fwd on [fake2; u] with the updated D -> seed -> D data-backward on the fake2 rows only
             -> G backward -> Adam.
    DNN    : fwd, seed, backward, Adam.
All tensors are torch tensors used as device memory handles; all arithmetic is done by `ops` (ops_cuda.CudaOps in the
product; tests inject tests/torch_ops.py to check this schedule against the oracle on CPU).

Activation layout: row-major [rows = sample-major pixels, channels] (NHWC).  Buffer blocks, in units of the local
batch B:  acts[l] rows [0,B)=x  [B,2B)=u  [2B,3B)=fake  [3B,4B)=x_hat  [4B,5B)=tangent u_l ;
          delta[l] rows [0,4B) = dLoss/da_l, [4B,5B) = gamma_l = (ds/da_l) of the gradient-penalty g-chain.
"""
from __future__ import annotations

import itertools
import math
import os
from typing import Dict, List, Optional

import torch

from .nets import Net, Layer, ACT_NONE, ACT_LEAKY, ACT_TANH

EPI_BIAS_ACT, EPI_DACT = 0, 1
DIST_KINDS = {'abs_mean': 0, 'abs_mean_neg': 1, 'abs_plus_one_sqrt_mean_neg': 2, 'abs_plus_one_log_mean_neg': 3,
              'square_mean': 4, 'norm_mean': 5}
# scalar slots (device fp32 buffer, read back only on summary steps: same sync pattern as srgan.py:306-319)
SC_DNN, SC_LABELED, SC_UNLABELED, SC_FAKE, SC_GP, SC_GNORM, SC_GEN = range(7)
N_SCALARS = 8


def partial_scalar_slots(method: str):
    """Scalar slots that hold per-rank PARTIAL sums (per-sample loss terms, already divided by the global batch) and must
    be summed across ranks before they are read; the SR-GAN feature-distance losses are computed from all-reduced
    feature sums and are already global on every rank.  In DG-GAN every loss is a per-sample BCE mean."""
    if method in ('dggan', 'sgan'):            # SGAN: cross entropy / BCE-of-logsumexp means, per sample as well
        return (SC_DNN, SC_LABELED, SC_UNLABELED, SC_FAKE, SC_GP, SC_GNORM, SC_GEN)
    return (SC_DNN, SC_LABELED, SC_GP, SC_GNORM)


def _strides_for(layer_kind: str, dims, Ca, Cb, R, S):
    """Element strides, per master dim, into the Wd[a][r][s][b] and Wu[b][r][s][a] kernel layouts (nets.py)."""
    if layer_kind == 'conv':        # master[a][b][r][s]
        wd = (R * S * Cb, 1, S * Cb, Cb)
        wu = (1, R * S * Ca, S * Ca, Ca)
    elif layer_kind == 'fc_up':     # master[zb][c][r][s]; Linear a=(r,s,c), b=zb ; here R=S=1 in the geom, k from dims
        zb, c, k, _ = dims
        wd = (1, zb, k * c * zb, c * zb)
        wu = (Ca, 1, k * c, c)
    elif layer_kind == 'ct_gemm':   # master[c][1][r][s] (ConvTranspose2d, kernel = stride, one output channel) as a Linear
        _, _, k, _ = dims           # with a = (r, s), b = c:  Wd[a][b], Wu[b][a]
        wd = (1, 0, k * Cb, Cb)
        wu = (Ca, 0, k, 1)
    else:
        raise ValueError(layer_kind)
    return wd, wu


def _thin_strides(Ca, Cb, R, S, kpad):
    """Thin-layer lowering: master[a][b][r][s] -> Wd_pad[a][kpad] (k = (r*S+s)*Cb + b) and Wu_pad[kpad][a]; kpad = the
    row width of the im2col buffer (Layer.kpad: a multiple of the 64-element = 128-byte swizzle row in bf16)."""
    return (kpad, 1, S * Cb, Cb), (1, Ca, S * Cb * Ca, Cb * Ca)


class NetState:
    """Device-side state of one network: fp32 master parameters (the nn.Parameters themselves, updated in place),
    kernel-layout weight copies in the activation dtype, one flat fp32 gradient buffer, Adam moments."""

    def __init__(self, net: Net, params: Dict[str, torch.Tensor], act_dtype, device, tag='D', thin=False):
        self.net = net
        self.params = params
        self.act_dtype = act_dtype
        self.device = device
        self.tag = tag                             # buffer namespace: D and DNN share 'D' (same shapes, sequential use)
        # graph nets: only forward-'down' layers fed by the network input (the crowd stem) are lowered
        self.thin = {l.name for l in net.layers if thin and l.thin_ok and (net.graph is None or l.fwd == 'down')}
        n_total = 0
        g_total = 0
        self.slices = {}
        self.gslices = {}
        order = []
        # elements of a layer's kernel-layout weight copies: geom.Ca / geom.Cb may be padded beyond the master dims
        kl = {l.name: l.geom.Ca * l.geom.R * l.geom.S * l.geom.Cb for l in net.layers}
        thin_l = {l.name: l for l in net.layers}
        part_c0, c0 = {}, 0
        for h, ncol in (net.head_parts or []):
            part_c0[h] = c0
            c0 += ncol
        for l in net.layers:
            order += [l.name + '.weight'] + ([l.name + '.bias'] if l.has_bias else [])
        for op in net.affines:                     # eval-mode BatchNorm: weight / bias are trainable, the statistics are not
            order += [op.name + '.weight', op.name + '.bias']
        # head parts: all weights first (contiguous in the flat buffers, in feature-column order), then the biases
        order += [h + '.weight' for h, _ in net.parts()] + [h + '.bias' for h, _ in net.parts()]
        for k in order:
            p = params[k]
            if p.dtype != torch.float32 and p.dtype != torch.float64:
                raise TypeError(f'{k}: master parameters must be fp32')
            if not p.is_contiguous():
                raise ValueError(f'{k}: master parameter must be contiguous')
            n = p.numel()
            self.slices[k] = (n_total, n)
            n_total += (n + 3) // 4 * 4            # keep every tensor 16-byte aligned inside the flat buffers
            lname = k.rsplit('.', 1)[0]
            gn = n
            if k.endswith('.weight') and lname in self.thin:
                gn = p.shape[0] * thin_l[lname].kpad    # gradient in the padded Wd_pad layout [a][kpad]
            elif k.endswith('.weight') and lname in kl:
                gn = kl[lname]                     # gradient in the (possibly channel-padded) Wd layout
            if k.endswith('.weight') and lname in part_c0:
                continue                           # multi-part head: gradients live in the shared [outputs][F] region below
            self.gslices[k] = (g_total, gn)
            g_total += (gn + 3) // 4 * 4
        if net.head_parts is not None:
            # row o of the region = d/d(head output o) over all feature columns; part p owns columns [c0_p, c0_p + ncol_p)
            F = net.feature_size
            self.head_gbase = g_total
            for h, ncol in net.head_parts:
                self.gslices[h + '.weight'] = (g_total + part_c0[h], (net.head_outputs - 1) * F + ncol)
            g_total += (net.head_outputs * F + 3) // 4 * 4
        self.order = order
        mdt = params[order[0]].dtype
        self.grad = torch.zeros(g_total, dtype=mdt, device=device)
        self.exp_avg = torch.zeros(n_total, dtype=mdt, device=device)
        self.exp_avg_sq = torch.zeros(n_total, dtype=mdt, device=device)
        # [t, lr/(1-beta1^t), 1/sqrt(1-beta2^t)] in device memory (fp64 in the fp64 schedule tests): see srgan_adam_prepare
        self.adam_state = torch.zeros(3, dtype=mdt, device=device)
        self.plain_entries = None                  # adam_multi table rows, built on the first update
        self.layout_entries = None                 # adam_layout_multi table rows (tensors with kernel-layout copies)
        self.wd_, self.wu_ = {}, {}
        for l in net.layers:
            n = kl[l.name]
            if l.name in self.thin:
                n = l.geom.Ca * l.kpad             # padded copies; the pad columns / rows stay zero
            self.wd_[l.name] = torch.zeros(n, dtype=act_dtype, device=device)
            self.wu_[l.name] = torch.zeros(n, dtype=act_dtype, device=device)
        if net.head:
            self.whead = torch.empty(net.head_outputs * net.feature_size, dtype=mdt, device=device)
            self.hbias = torch.zeros(net.head_outputs, dtype=mdt, device=device)     # sum of the parts' biases

    def g(self, key):
        o, n = self.gslices[key]
        return self.grad[o:o + n]

    def strides(self, l: Layer):
        g = l.geom
        if l.name in self.thin:
            return _thin_strides(g.Ca, g.Cb, g.R, g.S, l.kpad)
        return _strides_for(l.master_kind, l.master_dims, g.Ca, g.Cb, g.R, g.S)

    def m(self, key):
        o, n = self.slices[key]
        return self.exp_avg[o:o + n]

    def v(self, key):
        o, n = self.slices[key]
        return self.exp_avg_sq[o:o + n]


class Engine:
    def __init__(self, ops, d_net: Net, g_net: Optional[Net], D: Dict[str, torch.Tensor],
                 G: Optional[Dict[str, torch.Tensor]], DNN: Optional[Dict[str, torch.Tensor]],
                 act_dtype=torch.float32, device='cuda', comm=None, thin_lowering=None):
        self.ops = ops
        self.device = torch.device(device)
        self.act_dtype = act_dtype
        self.comm = comm                           # dist.Comm or None (single rank)
        self.d_net, self.g_net = d_net, g_net
        # thin-layer lowering (3-channel image layers -> im2col/col2im + [pixels x 64] GEMM on the tensor cores):
        # on by default in the bf16 tensor-core mode, off in the fp32 SIMT parity mode
        self.thin = (act_dtype == torch.bfloat16) if thin_lowering is None else bool(thin_lowering)
        self.D = NetState(d_net, D, act_dtype, self.device, 'D', self.thin)
        self.G = NetState(g_net, G, act_dtype, self.device, 'G', self.thin) if g_net is not None else None
        self.DNN = NetState(d_net, DNN, act_dtype, self.device, 'D', self.thin) if DNN is not None else None
        self.mdt = self.D.grad.dtype               # fp32 in the product; tests may run the schedule in fp64
        self.scalars = torch.zeros(N_SCALARS, dtype=self.mdt, device=self.device)
        self._buf = {}
        self.buf_generation = 0                    # bumped when an existing scratch buffer is replaced (StepRunner drops its CUDA graphs)
        # scratch-buffer scope: the DNN step ('dnn') and the GAN step ('gan') never share a workspace buffer, so the two
        # step methods may be in flight at the same time on different streams (StepRunner overlaps them)
        self._scope = 'gan'
        # graph nets: the weight-gradient kernels of a backward pass (conv wgrad, bias sums, the tangent block's BatchNorm
        # scale gradients) are off the critical path -- nothing reads the flat gradient buffer before the optimizer -- so
        # they are enqueued on a side stream (one per scratch scope), forked after the op that produced their delta and
        # joined at the end of the pass.  The crowd data path is a chain of small kernels that under-fill the GPU.
        self.wgrad_side_stream = self.device.type == 'cuda' and os.environ.get('SRGAN_NO_WGRAD_STREAM', '0') != '1'
        self._wg_streams = {}
        # graph nets: side branches (the three crowd MapModules: ~10 launches each, hanging off cat2..cat4 and ending in
        # `features`) run on their own streams, concurrently with the trunk: forward = forked when their tap is written,
        # joined at the end of the pass; backward = forked at the start, joined before the trunk first adds into the
        # tapped concat delta
        self.branch_streams = self.device.type == 'cuda' and os.environ.get('SRGAN_NO_BRANCH_STREAMS', '0') != '1'
        self._br_streams = {}
        self._probe = None
        # srgan.py:332-386 side effects: keep a copy of the feature rows the reference would leave in
        # self.{labeled,unlabeled,fake,interpolates}_features (rows x | u | fake | x_hat of the D step; the fake block is
        # replaced by the generator step's, srgan.py:386).  Off by default: the mirror Experiment has no reader.
        self.publish_features = False
        if d_net.graph is not None and any(op.fuse for op in d_net.graph):
            ok = getattr(ops, 'bn_fusion_supported', None)
            if ok is None or not ok(act_dtype):
                raise ValueError('the net description asks for BatchNorm-fused dense-layer kernels (Op.fuse), which exist on '
                                 'the bf16 tcgen05 path only')
        for st in (self.D, self.G, self.DNN):
            if st is not None:
                self.repack(st)

    # ------------------------------------------------------------------ buffers
    def buf(self, key, shape, dtype=None, zero=False):
        dtype = dtype or self.act_dtype
        key = (self._scope, key)
        t = self._buf.get(key)
        n = 1
        for s in shape:
            n *= s
        if t is None or t.numel() < n or t.dtype != dtype:
            if t is not None:
                self.buf_generation += 1           # the old storage is freed: addresses captured in CUDA graphs dangle
            t = (torch.zeros if zero else torch.empty)(n, dtype=dtype, device=self.device)
            self._buf[key] = t
        return t[:n].view(*shape)

    # ------------------------------------------------------------------ weights
    def repack(self, st: NetState):
        """Master (torch layout, fp32) -> kernel layouts.  Also done by the fused Adam after every update."""
        self.ops.begin()
        for l in st.net.layers:
            w = st.params[l.name + '.weight']
            wd_s, wu_s = st.strides(l)
            wd, wu = self._needed_layouts(st, l)
            self.ops.repack(w, l.master_dims, wd, wd_s if wd is not None else None, wu, wu_s if wu is not None else None)
        if st.net.head:
            self._repack_head(st)

    @staticmethod
    def _needed_layouts(st: NetState, l: Layer):
        """Kernel-layout copies a layer really uses: the forward op's layout always; the transposed one only if a
        data-backward / tangent pass ever goes through the layer (never for the generator's first layer)."""
        wd, wu = st.wd_[l.name], st.wu_[l.name]
        first_of_g = st.net.kind == 'G' and l is st.net.layers[0]
        if first_of_g:
            if l.fwd == 'down':
                wu = None
            else:
                wd = None
        return wd, wu

    def _head_strides(self, net: Net):
        F = net.feature_size
        if net.head_master_kind == 'linear':
            return (net.head_outputs, F, 1, 1), (F, 1, 0, 0)
        c, h, w = net.feature_chw
        return (net.head_outputs, c, h, w), (F, 1, w * c, c)

    def _head_part_layouts(self, st: NetState):
        """[(key prefix, master dims, strides into whead, whead slice)] of every head part."""
        net = st.net
        if net.head_parts is None:
            dims, s = self._head_strides(net)
            return [(net.head, dims, s, st.whead)]
        out, c0, F = [], 0, net.feature_size
        for h, ncol in net.head_parts:             # crowd: four Conv2d(20, n_out, 1) heads over consecutive feature columns
            out.append((h, (net.head_outputs, ncol, 1, 1), (F, 1, 0, 0), st.whead[c0:]))
            c0 += ncol
        return out

    def _repack_head(self, st: NetState):
        for h, dims, s, dst in self._head_part_layouts(st):
            self.ops.repack(st.params[h + '.weight'], dims, dst, s, None, None)
        if st.net.head_parts is not None:
            self._head_bias(st)

    def _head_bias(self, st: NetState):
        """hbias[o] = sum over the head parts of bias[o] (one part for the chain nets)."""
        st.hbias.zero_()
        for h, _ in st.net.parts():
            b = st.params[h + '.bias']
            for o in range(st.net.head_outputs):
                self.ops.colsum(b.detach()[o:o + 1], 1, 1, st.hbias[o:o + 1], 0, None)

    # ------------------------------------------------------------------ layer ops
    def _lin(self, l: Layer):
        """The [pixels x kpad] GEMM a thin layer runs as: small side = [pix, Ca], large side = col [pix, kpad]."""
        from .nets import Geom
        return Geom(1, 1, l.geom.Ca, 1, 1, l.kpad, 1, 1, 1, 0)

    def _col(self, st: NetState, l: Layer, lo, n):
        P = l.geom.Hs * l.geom.Ws
        return self._buf[(self._scope, ('col', st.tag, l.name))][lo * P * l.kpad:(lo + n) * P * l.kpad]

    def _coltmp(self, l: Layer, n):
        return self.buf(('coltmp',), (n * l.geom.Hs * l.geom.Ws * l.kpad,))

    def _fwd_layer(self, st: NetState, l: Layer, x, y, n, bias=True, href=None, epi=EPI_BIAS_ACT, act=None, slope=None,
                   lo=0, views=None):
        act = l.act if act is None else act
        slope = l.slope if slope is None else slope
        b = st.params[l.name + '.bias'] if (bias and l.has_bias) else None
        timed = self._probe_open('tangent' if epi == EPI_DACT else 'forward', st, l, n)
        if l.name in st.thin:
            P = l.geom.Hs * l.geom.Ws
            if l.fwd == 'down':      # im2col (kept for the weight gradient) -> GEMM with the layer's epilogue
                col = self._col(st, l, lo, n)
                self.ops.im2col(x, col, n, l.geom, l.kpad)
                self.ops.conv_down(col, st.wd_[l.name], y, n * P, self._lin(l), b, 0, href, epi, act, slope)
            else:                    # GEMM -> col2im with the layer's epilogue
                ycol = self._coltmp(l, n)
                self.ops.conv_up(x, st.wu_[l.name], ycol, n * P, self._lin(l), None, 0, None, EPI_DACT, ACT_NONE, 0.0)
                self.ops.col2im(ycol, y, n, l.geom, l.kpad, b, href, epi, act, slope)
        elif views is not None:                # channel window on the output side (dense-layer conv2 -> concat buffer)
            if l.fwd != 'down':
                raise ValueError('channel windows are implemented for forward-down layers')
            self.ops.conv_down(x, st.wd_[l.name], y, n, l.geom, b, l.bias_mod, href, epi, act, slope, views=views)
        elif l.fwd == 'down':
            self.ops.conv_down(x, st.wd_[l.name], y, n, l.geom, b, l.bias_mod, href, epi, act, slope)
        else:
            self.ops.conv_up(x, st.wu_[l.name], y, n, l.geom, b, l.bias_mod, href, epi, act, slope)
        self._probe_close(timed)

    # ------------------------------------------------------------------ live kernel probe (bench.py roofline)
    def probe_begin(self, select, label):
        """Times, with CUDA events on the launch stream, every contraction launch for which
        select(role, net_state, layer, rows) is true until probe_end(); role in {'forward', 'tangent', 'dgrad', 'wgrad'},
        rows = samples (x gemm_rows for the [pixels x C] GEMM layers).  Thin-lowered layers are timed with their
        im2col / col2im companions."""
        self._probe = {'select': select, 'label': label, 'events': [], 'flops': 0.0, 'elems': 0.0}

    def _probe_open(self, role, st, l, n):
        pr = self._probe
        if pr is None or not pr['select'](role, st, l, n):
            return None
        g = l.geom
        if l.gemm_rows > 1:                        # [pixels x C] GEMM: algorithmic sizes from the master dims
            d0, d1 = l.master_dims[0], l.master_dims[1] * l.master_dims[2] * l.master_dims[3]
            small, large, wts, macs = n * d0, n * d1, d0 * d1, n * d0 * d1
        else:
            d = l.master_dims
            ca, cb = (d[0], d[1]) if l.master_kind == 'conv' else (g.Ca, g.Cb)
            small, large, wts = n * g.Hs * g.Ws * ca, n * g.Hl * g.Wl * cb, ca * g.R * g.S * cb
            macs = n * g.Hs * g.Ws * ca * g.R * g.S * cb
        out_side = (small if l.fwd == 'down' else large) if role in ('forward', 'tangent') else (large if l.fwd == 'down' else small)
        # operands once each; the tangent / data-gradient epilogues also read the stored activation of the output side;
        # the weight gradient reads both activations and writes fp32 weights (counted as 2 elements each); the BatchNorm-
        # fused data gradient reads the concat buffer and reads + writes its delta (3 output-side streams)
        pr['elems'] += (small + large + (2 * wts if role == 'wgrad' else wts) + (out_side if role in ('tangent', 'dgrad') else 0)
                        + (2 * out_side if role == 'dgrad_bn' else 0))
        pr['flops'] += 2.0 * macs
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(self.device))
        return e0, e1

    def _probe_close(self, timed):
        if timed is not None:
            timed[1].record(torch.cuda.current_stream(self.device))
            self._probe['events'].append(timed)

    def probe_end(self):
        pr, self._probe = self._probe, None
        if not pr or not pr['events']:
            return {'count': 0}
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in pr['events'])
        n = len(pr['events'])
        return {'count': n, 'ms': ms, 'flops_per_launch': pr['flops'] / n,
                'bytes_per_launch': pr['elems'] / n * (2 if self.act_dtype == torch.bfloat16 else 4), 'kernel': pr['label']}

    def _bwd_data_layer(self, st: NetState, l: Layer, dy, dx, n, href, act, slope, lo=0, col_ready=False, views=None):
        """dx = (W_l^T dy) * act'(href)   (act = ACT_NONE: no mask)."""
        timed = self._probe_open('dgrad', st, l, n)
        if views is not None:                  # dy is a channel window of a concat delta (forward-down layers only)
            self.ops.conv_up(dy, st.wu_[l.name], dx, n, l.geom, None, 0, href, EPI_DACT, act, slope, views=views)
        else:
            self._bwd_data_layer_(st, l, dy, dx, n, href, act, slope, lo, col_ready)
        self._probe_close(timed)

    def _bwd_data_layer_(self, st: NetState, l: Layer, dy, dx, n, href, act, slope, lo, col_ready):
        if l.name in st.thin:
            P = l.geom.Hs * l.geom.Ws
            if l.fwd == 'down':      # transpose of (im2col -> GEMM): GEMM^T -> col2im with the mask fused
                colg = self._coltmp(l, n)
                self.ops.conv_up(dy, st.wu_[l.name], colg, n * P, self._lin(l), None, 0, None, EPI_DACT, ACT_NONE, 0.0)
                self.ops.col2im(colg, dx, n, l.geom, l.kpad, None, href, EPI_DACT, act, slope)
            else:                    # transpose of (GEMM -> col2im): im2col -> GEMM^T
                col = self._col(st, l, lo, n)
                if not col_ready:
                    self.ops.im2col(dy, col, n, l.geom, l.kpad)
                self.ops.conv_down(col, st.wd_[l.name], dx, n * P, self._lin(l), None, 0, href, EPI_DACT, act, slope)
        elif l.fwd == 'down':
            self.ops.conv_up(dy, st.wu_[l.name], dx, n, l.geom, None, 0, href, EPI_DACT, act, slope)
        else:
            self.ops.conv_down(dy, st.wd_[l.name], dx, n, l.geom, None, 0, href, EPI_DACT, act, slope)

    def _wgrad_layer(self, st: NetState, l: Layer, x_in, dy, n, lo=0, views=None):
        timed = self._probe_open('wgrad', st, l, n)
        if views is not None:
            self.ops.conv_wgrad(dy, x_in, st.g(l.name + '.weight'), n, l.geom, views=views)
        else:
            self._wgrad_layer_(st, l, x_in, dy, n, lo)
        self._probe_close(timed)

    def _wgrad_layer_(self, st: NetState, l: Layer, x_in, dy, n, lo):
        dW = st.g(l.name + '.weight')
        if l.name in st.thin:
            P = l.geom.Hs * l.geom.Ws
            col = self._col(st, l, lo, n)
            if l.fwd == 'down':      # col = im2col(layer input), written by the forward pass
                self.ops.conv_wgrad(dy, col, dW, n * P, self._lin(l))
            else:                    # large side = dy: im2col it here, the data-backward reuses it (col_ready)
                self.ops.im2col(dy, col, n, l.geom, l.kpad)
                self.ops.conv_wgrad(x_in, col, dW, n * P, self._lin(l))
        elif l.fwd == 'down':
            self.ops.conv_wgrad(dy, x_in, dW, n, l.geom)
        else:
            self.ops.conv_wgrad(x_in, dy, dW, n, l.geom)

    def _bias_grad(self, st: NetState, l: Layer, dy, rows):
        self.ops.colsum(dy, rows, l.out_ch, st.g(l.name + '.bias'), l.bias_mod, None)

    # ------------------------------------------------------------------ passes
    def alloc_acts(self, tag, net: Net, nb_rows):
        """Chain nets: acts[0] = network input, acts[l] = output of layer l (a list).  Graph nets: a dict buffer name ->
        tensor.  Flat tensors, nb_rows samples of NHWC rows each."""
        if net.graph is not None:
            # zero-filled on (re)allocation: the pad channels of the padded operand buffers are never written
            if self.thin:
                for l in net.layers:
                    if l.thin_ok and l.fwd == 'down':
                        self.buf(('col', tag, l.name), (nb_rows * l.geom.Hs * l.geom.Ws * l.kpad,))
            acts = {name: self.buf((tag, 'a', name), (nb_rows * b.rows * b.ch,), zero=True) for name, b in net.bufs.items()}
            for op in net.graph:
                if op.kind == 'maxpool':               # one byte per pooled element: the max-pool index map
                    acts['idx:' + op.dst] = self.buf((tag, 'idx', op.dst), (nb_rows * net.bufs[op.dst].rows * op.C,),
                                                     dtype=torch.uint8)
            return acts
        acts = [self.buf((tag, 'a', 0), (nb_rows * net.layers[0].in_elems,))]
        for i, l in enumerate(net.layers, 1):
            acts.append(self.buf((tag, 'a', i), (nb_rows * l.out_elems,)))
            if self.thin and l.thin_ok:
                self.buf(('col', tag, l.name), (nb_rows * l.geom.Hs * l.geom.Ws * l.kpad,))
        return acts

    def alloc_deltas(self, tag, net: Net, nb_rows):
        if net.graph is not None:
            return {name: self.buf((tag, 'd', name), (nb_rows * b.rows * b.ch,), zero=True) for name, b in net.bufs.items()
                    if name != net.input_buf}
        d = [None]
        for i, l in enumerate(net.layers, 1):
            d.append(self.buf((tag, 'd', i), (nb_rows * l.out_elems,)))
        return d

    @staticmethod
    def rows(t, elems_per_sample, lo, hi):
        return t[lo * elems_per_sample: hi * elems_per_sample]

    @staticmethod
    def _brows(t, b, lo, hi):
        e = b.rows * b.ch
        return t[lo * e: hi * e]

    @staticmethod
    def _window(op, net: Net):
        """srgan_views tuple (S_pitch, S_valid, L_pitch, L_valid) of a conv op whose output is a channel window of its
        destination buffer, else None."""
        return (net.bufs[op.dst].ch, op.C, 0, 0) if op.C else None

    @staticmethod
    def _in(net: Net, acts):
        return acts[net.input_buf] if net.graph is not None else acts[0]

    @staticmethod
    def _feat(net: Net, t):
        return t[net.feature_buf] if net.graph is not None else t[len(net.layers)]

    def forward(self, st: NetState, acts, lo, hi, keep_pre=True):
        """Runs the network on sample rows [lo, hi) of the buffers.  keep_pre: which rows of the operands that fused kernels
        produce on the fly (Op.fuse >= 2: the normalised input of the trunk's 1x1 convolutions) are also stored -- True: all
        (a weight gradient over the stored operand follows), False: none (forward / data gradient only), a sample index k:
        samples >= k (nets whose weight gradients read the concat buffer themselves, Op.fuse >= 5: only the tangent pass
        of the interpolate rows reads the stored operand)."""
        net = st.net
        if net.graph is not None:
            return self.graph_forward(st, acts, lo, hi, keep_pre=keep_pre)
        n = hi - lo
        for i, l in enumerate(net.layers, 1):
            self._fwd_layer(st, l, self.rows(acts[i - 1], l.in_elems, lo, hi), self.rows(acts[i], l.out_elems, lo, hi), n,
                            lo=lo)

    def backward(self, st: NetState, acts, deltas, lo, hi, wlo=None, whi=None, need_input_grad=False, dinput=None,
                 input_href=None, input_act=ACT_NONE, weight_grads=True, hook=None):
        """Ordinary reverse pass over sample rows [lo,hi); weight gradients over rows [wlo,whi) (defaults to the
        same rows; the D step widens it to include the tangent block)."""
        net = st.net
        wlo = lo if wlo is None else wlo
        whi = hi if whi is None else whi
        if net.graph is not None:
            return self.graph_backward(st, acts, deltas, lo, hi, lo, wlo, whi, weight_grads,
                                       (dinput, input_href, input_act) if need_input_grad else None, hook)
        for i in range(len(net.layers), 0, -1):
            l = net.layers[i - 1]
            if weight_grads:
                self._wgrad_layer(st, l, self.rows(acts[i - 1], l.in_elems, wlo, whi),
                                  self.rows(deltas[i], l.out_elems, wlo, whi), whi - wlo, lo=wlo)
                if l.has_bias:
                    self._bias_grad(st, l, self.rows(deltas[i], l.out_elems, lo, hi), (hi - lo) * l.out_rows)
            if i > 1:
                lp = net.layers[i - 2]
                self._bwd_data_layer(st, l, self.rows(deltas[i], l.out_elems, lo, hi),
                                     self.rows(deltas[i - 1], lp.out_elems, lo, hi), hi - lo,
                                     self.rows(acts[i - 1], lp.out_elems, lo, hi), lp.act, lp.slope, lo=lo,
                                     col_ready=(weight_grads and wlo == lo and whi == hi))
            elif need_input_grad:
                self._bwd_data_layer(st, l, self.rows(deltas[i], l.out_elems, lo, hi), dinput, hi - lo,
                                     input_href, input_act, 0.0, lo=lo,
                                     col_ready=(weight_grads and wlo == lo and whi == hi))

    def gchain(self, st: NetState, acts, deltas, B, g0):
        """Gradient-penalty g-chain on the x_hat rows: gamma_{l-1} = W_l^T gamma_l * act'(h_{l-1}),
        delta rows [4B,5B), masks from the activations of rows [3B,4B); g0 = d s / d x_hat (no mask)."""
        net = st.net
        if net.graph is not None:
            return self.graph_backward(st, acts, deltas, 4 * B, 5 * B, 3 * B, 0, 0, False, (g0, None, ACT_NONE), None,
                                       keep_pre=True)
        L = len(net.layers)
        for i in range(L, 1, -1):
            l, lp = net.layers[i - 1], net.layers[i - 2]
            self._bwd_data_layer(st, l, self.rows(deltas[i], l.out_elems, 4 * B, 5 * B),
                                 self.rows(deltas[i - 1], lp.out_elems, 4 * B, 5 * B), B,
                                 self.rows(acts[i - 1], lp.out_elems, 3 * B, 4 * B), lp.act, lp.slope)
        l1 = net.layers[0]
        self._bwd_data_layer(st, l1, self.rows(deltas[1], l1.out_elems, 4 * B, 5 * B), g0, B, None, ACT_NONE, 0.0)

    def tangent(self, st: NetState, acts, B):
        """Tangent u-chain: u_l = (W_l u_{l-1}) * act'(h_l of x_hat), no bias; rows [4B,5B), masks from rows [3B,4B)."""
        net = st.net
        if net.graph is not None:
            return self.graph_forward(st, acts, 4 * B, 5 * B, tangent=True, mlo=3 * B)
        for i, l in enumerate(net.layers, 1):
            self._fwd_layer(st, l, self.rows(acts[i - 1], l.in_elems, 4 * B, 5 * B),
                            self.rows(acts[i], l.out_elems, 4 * B, 5 * B), B, bias=False,
                            href=self.rows(acts[i], l.out_elems, 3 * B, 4 * B), epi=EPI_DACT, lo=4 * B)

    def tangent_block_grads(self, st: NetState, acts, deltas, tlo, thi):
        """Parameter gradients of the tangent block alone (wgrad(u_{l-1}, gamma_l); no bias terms): used
        when the block is not contiguous with the ordinary rows of the backward pass (DG-GAN)."""
        net, R = st.net, self._brows
        n = thi - tlo
        if net.graph is None:
            for i, l in enumerate(net.layers, 1):
                self._wgrad_layer(st, l, self.rows(acts[i - 1], l.in_elems, tlo, thi),
                                  self.rows(deltas[i], l.out_elems, tlo, thi), n, lo=tlo)
            return
        for op in net.graph:
            sb, db = net.bufs[op.src], net.bufs[op.dst]
            if op.kind == 'conv':
                self._wgrad_layer(st, op.layer, R(acts[op.src], sb, tlo, thi), R(deltas[op.dst], db, tlo, thi)[op.c0:],
                                  n * op.layer.gemm_rows, lo=tlo, views=self._window(op, net))
            elif op.kind == 'affine':
                P, nm = st.params, op.name
                self.ops.affine_grad(R(deltas[op.dst], db, tlo, thi), db.ch, R(acts[op.src], sb, tlo, thi), sb.ch, op.c0,
                                     n * sb.rows, op.C, P[nm + '.running_mean'], P[nm + '.running_var'], self.BN_EPS,
                                     st.g(nm + '.weight'), None, False)

    # ------------------------------------------------------------------ graph nets (crowd KnnDenseNetCat)
    BN_EPS = 1e-5          # nn.BatchNorm2d default, crowd/models.py:339,343,367,1077,1092

    def graph_forward(self, st: NetState, acts, lo, hi, tangent=False, mlo=None, keep_pre=True):
        """Forward (tangent=False) or tangent pass (tangent=True: linear parts only, activation derivatives taken from the
        stored activations of rows [mlo, mlo+n)) of a graph net over sample rows [lo,hi)."""
        net, ops, R, P = st.net, self.ops, self._brows, st.params
        n = hi - lo
        streams = self._branch_streams_for(net)
        main = torch.cuda.current_stream(self.device) if streams else None
        tap_ev, done = {}, []
        taps = net.branch_taps() if streams else ()
        for branch, run in itertools.groupby(net.graph, key=lambda o: o.branch):
            run = list(run)
            if not streams or branch == 0:
                for op in run:
                    self._graph_forward_op(st, acts, op, lo, hi, n, tangent, mlo, keep_pre)
                    if streams and op.dst in taps and op.dst not in tap_ev:      # first writer of a tapped buffer (the pool)
                        ev = torch.cuda.Event()
                        ev.record(main)
                        tap_ev[op.dst] = ev
                continue
            s = streams[branch]
            for name in {op.src for op in run} & set(taps):
                if name in tap_ev:
                    s.wait_event(tap_ev[name])
                else:
                    s.wait_stream(main)
            with torch.cuda.stream(s):
                prev = ops.use_stream(s)
                try:
                    for op in run:
                        self._graph_forward_op(st, acts, op, lo, hi, n, tangent, mlo, keep_pre)
                finally:
                    ops.restore_stream(prev)
                ev = torch.cuda.Event()
                ev.record(s)
                done.append(ev)
        for ev in done:
            main.wait_event(ev)

    def _branch_streams_for(self, net: Net):
        """{branch id: stream} of the current scratch scope, or None when side branches run inline."""
        if not (self.branch_streams and hasattr(self.ops, 'use_stream')) or not any(op.branch for op in net.graph):
            return None
        d = self._br_streams.setdefault(self._scope, {})
        for op in net.graph:
            if op.branch and op.branch not in d:
                d[op.branch] = torch.cuda.Stream(self.device)
        return d

    @staticmethod
    def _bn_params(P, nm):
        return P[nm + '.weight'], P[nm + '.bias'], P[nm + '.running_mean'], P[nm + '.running_var']

    def _graph_forward_op(self, st: NetState, acts, op, lo, hi, n, tangent, mlo, keep_pre=True):
        """One op of graph_forward over sample rows [lo,hi) (n = hi - lo), on the current stream."""
        net, ops, R, P = st.net, self.ops, self._brows, st.params
        sb, db = net.bufs[op.src], net.bufs[op.dst]
        x, y = R(acts[op.src], sb, lo, hi), R(acts[op.dst], db, lo, hi)
        href = R(acts[op.dst], db, mlo, mlo + n) if tangent else None
        if op.kind == 'conv':
            ng = n * op.layer.gemm_rows
            vw = self._window(op, net)
            if vw is not None:                 # the layer's channels of the concat buffer, written in place (no activation)
                y, href = y[op.c0:], None
            if tangent:
                self._fwd_layer(st, op.layer, x, y, ng, bias=False, href=href, epi=EPI_DACT, lo=lo, views=vw)
            elif op.pre is not None and op.pre.fuse >= 2:
                # BatchNorm + ReLU applied to the operand tiles on their way into the GEMM: reads the concat buffer directly;
                # the normalised operand is stored only for the rows a weight gradient / tangent pass will read
                a, l = op.pre, op.layer
                cbuf, nm = net.bufs[a.src], a.name
                timed = self._probe_open('forward_bn', st, l, ng)
                if keep_pre is True:
                    first = 0
                elif keep_pre is False or a.fuse < 5:          # (a sample index only means something when the weight
                    first = 0 if keep_pre is not False else None   # gradient reads the concat buffer too: fuse >= 5)
                else:
                    first = max(0, keep_pre - lo) * l.gemm_rows
                ops.bn_conv_down(R(acts[a.src], cbuf, lo, hi), st.wd_[l.name], y, ng, l.geom.Cb, l.geom.Ca, a.C, cbuf.ch,
                                 P[nm + '.weight'], P[nm + '.bias'], P[nm + '.running_mean'], P[nm + '.running_var'], self.BN_EPS,
                                 x if first is not None and first < ng else None, sb.ch, n1_first_row=first or 0,
                                 bn2=self._bn_params(P, op.post.name) if op.post is not None else None,
                                 out2=R(acts[op.post.dst], net.bufs[op.post.dst], lo, hi) if op.post is not None else None)
                self._probe_close(timed)
            else:
                self._fwd_layer(st, op.layer, x, y, ng, lo=lo, views=vw)
        elif op.kind == 'affine' and (op.fuse >= 2 or op.absorbed) and not tangent:
            pass                          # carried out by the convolution that consumes its output / produces its input (above)
        elif op.kind == 'affine':
            nm = op.name
            ops.affine(x, sb.ch, op.c0, y, db.ch, n * sb.rows, op.C, P[nm + '.weight'], P[nm + '.bias'],
                       P[nm + '.running_mean'], P[nm + '.running_var'], self.BN_EPS, href, 1 if tangent else 0,
                       db.act, db.slope)
        elif op.kind == 'copy':
            ops.copy2d(x, sb.ch, 0, y, db.ch, op.c0, n * sb.rows, op.C, False)
        elif op.kind == 'read':
            ops.copy2d(x, sb.ch, op.c0, y, db.ch, 0, n * sb.rows, op.C, False)
        elif op.kind == 'maxpool':
            # the forward records the winning window position per pooled element; the tangent pass routes by the map of
            # the x_hat rows, the backward pass reads the map of its own rows
            im = acts['idx:' + op.dst]
            e = db.rows * op.C                  # index bytes per sample
            if tangent:
                ops.maxpool(x, None, y, db.ch, op.c0, n, op.H, op.W, op.C, op.k, op.stride, op.pad,
                            idx=im[mlo * e:(mlo + n) * e], idx_mode=2)
            else:
                ops.maxpool(x, None, y, db.ch, op.c0, n, op.H, op.W, op.C, op.k, op.stride, op.pad,
                            idx=im[lo * e:hi * e], idx_mode=1)
        elif op.kind == 'avgpool':
            ops.avgpool(x, sb.ch, y, db.ch, op.c0, n, op.H, op.W, op.C, op.k)
        elif op.kind == 'shuffle':
            ops.depth_to_space(x, y, n, op.H, op.W, op.k, False)
        else:
            raise ValueError(op.kind)

    def graph_backward(self, st: NetState, acts, deltas, lo, hi, mlo, wlo, whi, weight_grads, input_grad, hook, keep_pre=False):
        """Reverse pass of a graph net over delta rows [lo,hi); the activations that provide masks and weight-gradient
        inputs are rows [mlo, mlo+n) (== [lo,hi) for an ordinary backward; the x_hat rows for the g-chain).
        Convolution weight gradients cover rows [wlo,whi) of (acts, deltas) in one launch -- ordinary rows plus, when
        whi > hi, the tangent block (u_{l-1}, gamma_l); the BatchNorm-scale gradient of the tangent block is a second call
        (no mean subtraction, no bias term).  input_grad = (dinput, href, act): also d/d(network input).
        hook(buffer name) is called before the producer of a map buffer is processed (crowd labeled map loss).
        keep_pre: fused BatchNorm ops (Op.fuse) also store the delta w.r.t. the BatchNorm output's pre-activation, which
        the unfused pass leaves in the delta buffer of that output (the g-chain: its rows are the tangent block of the
        BatchNorm-scale gradient)."""
        net, ops, R, P = st.net, self.ops, self._brows, st.params
        n = hi - lo
        for name, b in net.bufs.items():
            if b.accumulate:
                R(deltas[name], b, lo, hi).zero_()
        side = None
        if weight_grads and self.wgrad_side_stream and hasattr(ops, 'use_stream'):
            side = self._wg_streams.get(self._scope)
            if side is None:
                side = self._wg_streams[self._scope] = torch.cuda.Stream(self.device)

        def off_path(fn):
            """fn() enqueues weight-gradient kernels: on the side stream, after everything enqueued so far."""
            if side is None:
                fn()
                return
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            side.wait_event(ev)
            prev = ops.use_stream(side)
            try:
                fn()
            finally:
                ops.restore_stream(prev)
        def one(op):
            sb, db = net.bufs[op.src], net.bufs[op.dst]
            dy = R(deltas[op.dst], db, lo, hi)
            is_input = op.src == net.input_buf
            dx = None if is_input else R(deltas[op.src], sb, lo, hi)
            xa = R(acts[op.src], sb, mlo, mlo + n)
            if hook is not None and op.dst in net.map_bufs:
                hook(op.dst)
            if op.kind == 'conv':
                l = op.layer
                vw = self._window(op, net)
                if vw is not None:
                    dy = dy[op.c0:]
                if weight_grads:
                    def wg(l=l, op=op, sb=sb, db=db, dy=dy, vw=vw):
                        w0 = wlo
                        if op.pre is not None and op.pre.fuse >= 5:
                            # the ordinary rows straight from the concat buffer (BatchNorm + ReLU applied on load); the
                            # tangent block's operand is not relu(bn(.)): it stays a plain weight gradient over its rows
                            a = op.pre
                            cbuf, nm = net.bufs[a.src], a.name
                            w0 = min(whi, hi)
                            if w0 > wlo:
                                timed = self._probe_open('wgrad_bn', st, l, (w0 - wlo) * l.gemm_rows)
                                ops.bn_conv_wgrad(R(deltas[op.dst], db, wlo, w0), R(acts[a.src], cbuf, wlo, w0),
                                                  st.g(l.name + '.weight'), (w0 - wlo) * l.gemm_rows, l.geom.Ca, l.geom.Cb, a.C,
                                                  cbuf.ch, P[nm + '.weight'], P[nm + '.bias'], P[nm + '.running_mean'],
                                                  P[nm + '.running_var'], self.BN_EPS)
                                self._probe_close(timed)
                        if whi > w0:
                            self._wgrad_layer(st, l, R(acts[op.src], sb, w0, whi), R(deltas[op.dst], db, w0, whi)[op.c0:],
                                              (whi - w0) * l.gemm_rows, lo=w0, views=vw)
                        if l.has_bias:
                            self._bias_grad(st, l, dy, n * db.rows)
                    if l.name in st.thin and l.fwd != 'down':      # its im2col output is reused by the data gradient below
                        wg()
                    else:
                        off_path(wg)
                if is_input:
                    if input_grad is not None:
                        dinput, ihref, iact = input_grad
                        self._bwd_data_layer(st, l, dy, dinput, n * l.gemm_rows, ihref, iact, 0.0, lo=lo)
                elif op.pre is not None and op.pre.fuse:
                    # BatchNorm + ReLU backward fused into this data-gradient GEMM: straight into the concat delta
                    a = op.pre
                    cbuf, nm = net.bufs[a.src], a.name
                    mean, var = P[nm + '.running_mean'], P[nm + '.running_var']
                    timed = self._probe_open('dgrad_bn', st, l, n * l.gemm_rows)
                    ops.bn_dgrad(dy, st.wu_[l.name], R(deltas[a.src], cbuf, lo, hi), R(acts[a.src], cbuf, mlo, mlo + n),
                                 n * l.gemm_rows, l.geom.Ca, l.geom.Cb, a.C, cbuf.ch, P[nm + '.weight'], P[nm + '.bias'], mean, var,
                                 self.BN_EPS, st.g(nm + '.weight') if weight_grads else None,
                                 st.g(nm + '.bias') if weight_grads else None, dx if keep_pre else None, sb.ch, cbuf.accumulate)
                    self._probe_close(timed)
                    if weight_grads and whi > hi:      # the tangent block's BatchNorm-scale gradient (the g-chain kept its delta)
                        off_path(lambda a=a, sb=sb, cbuf=cbuf, nm=nm, mean=mean, var=var, op=op: ops.affine_grad(
                            R(deltas[op.src], sb, hi, whi), sb.ch, R(acts[a.src], cbuf, hi, whi), cbuf.ch, a.c0,
                            (whi - hi) * cbuf.rows, a.C, mean, var, self.BN_EPS, st.g(nm + '.weight'), None, False))
                elif op.bwd_pre is not None:
                    # norm2 + relu2 backward fused into the 3x3 convolution's data gradient: dy is the layer's channel window
                    # of the concat delta, the result goes straight into the delta of the bottleneck buffer
                    a = op.bwd_pre
                    bbuf, nm = net.bufs[a.src], a.name
                    mean, var = P[nm + '.running_mean'], P[nm + '.running_var']
                    timed = self._probe_open('dgrad_bn3', st, l, n)
                    ops.bn_conv_dgrad(dy, db.ch if vw is not None else 0, op.C if vw is not None else 0, st.wu_[l.name],
                                      R(deltas[a.src], bbuf, lo, hi), R(acts[a.src], bbuf, mlo, mlo + n), n, l.geom, a.C, bbuf.ch,
                                      P[nm + '.weight'], P[nm + '.bias'], mean, var, self.BN_EPS,
                                      st.g(nm + '.weight') if weight_grads else None, st.g(nm + '.bias') if weight_grads else None,
                                      dx if keep_pre else None, sb.ch, bbuf.accumulate)
                    self._probe_close(timed)
                    if weight_grads and whi > hi:      # the tangent block's BatchNorm-scale gradient (the g-chain kept its delta)
                        off_path(lambda a=a, sb=sb, bbuf=bbuf, nm=nm, mean=mean, var=var, op=op: ops.affine_grad(
                            R(deltas[op.src], sb, hi, whi), sb.ch, R(acts[a.src], bbuf, hi, whi), bbuf.ch, a.c0,
                            (whi - hi) * bbuf.rows, a.C, mean, var, self.BN_EPS, st.g(nm + '.weight'), None, False))
                else:
                    self._bwd_data_layer(st, l, dy, dx, n * l.gemm_rows, xa, sb.act, sb.slope, lo=lo, views=vw)
            elif op.kind == 'affine' and (op.fuse or op.bwd_fused):
                pass                      # carried out by the convolution that consumes its output (above)
            elif op.kind == 'affine':
                nm = op.name
                mean, var = P[nm + '.running_mean'], P[nm + '.running_var']
                if weight_grads:          # parameter gradients and data gradient from one pass over dy
                    ops.affine_bwd_grad(dy, db.ch, xa, dx, sb.ch, op.c0, n * sb.rows, op.C, P[nm + '.weight'], mean, var,
                                        self.BN_EPS, st.g(nm + '.weight'), st.g(nm + '.bias'), sb.accumulate)
                    if whi > hi:
                        off_path(lambda op=op, sb=sb, db=db, nm=nm, mean=mean, var=var: ops.affine_grad(
                            R(deltas[op.dst], db, hi, whi), db.ch, R(acts[op.src], sb, hi, whi), sb.ch, op.c0,
                            (whi - hi) * sb.rows, op.C, mean, var, self.BN_EPS, st.g(nm + '.weight'), None, False))
                else:
                    ops.affine_bwd(dy, db.ch, dx, sb.ch, op.c0, n * sb.rows, op.C, P[nm + '.weight'], var, self.BN_EPS,
                                   sb.accumulate)
            elif op.kind == 'copy':       # dst slice <- dense src: both are w.r.t. the same pre-activation
                ops.copy2d(dy, db.ch, op.c0, dx, sb.ch, 0, n * sb.rows, op.C, False)
            elif op.kind == 'read':       # dense dst <- src slice: add into the (accumulating) concat delta
                ops.copy2d(dy, db.ch, 0, dx, sb.ch, op.c0, n * sb.rows, op.C, True)
            elif op.kind == 'maxpool':
                im = acts['idx:' + op.dst]
                e = db.rows * op.C
                ops.maxpool_bwd(xa, dy, db.ch, op.c0, dx, n, op.H, op.W, op.C, op.k, op.stride, op.pad, sb.act, sb.slope,
                                idx=im[mlo * e:(mlo + n) * e])
            elif op.kind == 'avgpool':
                ops.avgpool_bwd(dy, db.ch, op.c0, dx, sb.ch, n, op.H, op.W, op.C, op.k, xa, sb.act, sb.slope)
            elif op.kind == 'shuffle':    # a permutation: both sides are w.r.t. the same pre-activation
                ops.depth_to_space(dy, dx, n, op.H, op.W, op.k, True)
            else:
                raise ValueError(op.kind)

        streams = self._branch_streams_for(net)
        main = torch.cuda.current_stream(self.device) if streams else None
        taps = net.branch_taps() if streams else ()
        pending = {}                          # tapped concat buffer -> event: a branch still adds into its delta
        for branch, run in itertools.groupby(reversed(net.graph), key=lambda o: o.branch):
            run = list(run)
            if not streams or branch == 0:
                for op in run:
                    if pending and op.src in pending:          # the trunk's first read-modify-write of that concat delta
                        main.wait_event(pending.pop(op.src))
                    one(op)
                continue
            s = streams[branch]
            s.wait_stream(main)                                # the feature deltas / seeds enqueued so far
            with torch.cuda.stream(s):
                prev = ops.use_stream(s)
                try:
                    for op in run:
                        one(op)
                finally:
                    ops.restore_stream(prev)
                ev = torch.cuda.Event()
                ev.record(s)
            for name in {op.src for op in run} & set(taps):
                pending[name] = ev
        for ev in pending.values():
            main.wait_event(ev)
        if side is not None:                  # join: the optimizer (or the gradient all-reduce) follows
            ev = torch.cuda.Event()
            ev.record(side)
            torch.cuda.current_stream(self.device).wait_event(ev)

    # ------------------------------------------------------------------ Adam
    def adam(self, st: NetState, lr, weight_decay, betas=(0.9, 0.999), eps=1e-8, deferrable=None):
        """torch.optim.Adam semantics, one fused launch per tensor that also rewrites the kernel-
        layout copies.  Gradients are read from (and then zeroed in) the flat buffer.  deferrable = a tag: this update is
        the LAST thing its step method enqueues, so (multi-rank) its gradient all-reduce and the update may overlap with
        the next step method; StepRunner waits for the tag before the network or its gradient buffer is touched again."""
        if self.comm is not None:
            if deferrable:
                self.comm.begin_deferred(deferrable)
            self.comm.all_reduce_sum(st.grad)
        self.ops.adam_prepare(st.adam_state, lr, betas[0], betas[1])
        plain_keys = []

        def plain(k):
            plain_keys.append(k)
        # tensors with kernel-layout copies: (key, dims, gradient strides, out1, s1, out2, s2)
        laid = []
        for l in st.net.layers:
            wd_s, wu_s = st.strides(l)
            wd, wu = self._needed_layouts(st, l)
            laid.append((l.name + '.weight', l.master_dims, wd_s, wd, wd_s if wd is not None else None, wu,
                         wu_s if wu is not None else None))
            if l.has_bias:
                plain(l.name + '.bias')
        for op in st.net.affines:
            plain(op.name + '.weight')
            plain(op.name + '.bias')
        if st.net.head:
            for h, dims, s, dst in self._head_part_layouts(st):
                # multi-part heads: st.g(k) starts at the part's first column of row 0
                laid.append((h + '.weight', dims, s, dst, s, None, None))
                plain(h + '.bias')
        if hasattr(self.ops, 'adam_layout_multi') and len(laid) > 8 and self.mdt == torch.float32:
            # many tensors (the crowd discriminator: 200 convolution weights): one table-driven launch
            if st.layout_entries is None:
                st.layout_entries = [(st.params[k].detach(), st.gslices[k][0], st.slices[k][0], dims, gs, o1, s1, o2, s2)
                                     for k, dims, gs, o1, s1, o2, s2 in laid]
            self.ops.adam_layout_multi(st.layout_entries, st.grad, st.exp_avg, st.exp_avg_sq, st.adam_state, betas[0], betas[1],
                                       eps, weight_decay, st.act_dtype)
        else:
            for k, dims, gs, o1, s1, o2, s2 in laid:
                self.ops.adam(st.params[k], st.g(k), st.m(k), st.v(k), dims, gs, o1, s1, o2, s2, st.adam_state, betas[0],
                              betas[1], eps, weight_decay)
        # every tensor without kernel-layout copies (biases, BatchNorm weight / bias) in one launch
        if st.plain_entries is None:
            st.plain_entries = [(st.params[k].detach(), st.gslices[k][0], st.slices[k][0], st.params[k].numel())
                                for k in plain_keys]
        self.ops.adam_multi(st.plain_entries, st.grad, st.exp_avg, st.exp_avg_sq, st.adam_state, betas[0], betas[1], eps,
                            weight_decay)
        if st.net.head and st.net.head_parts is not None:
            self._head_bias(st)
        st.grad.zero_()
        if self.comm is not None and deferrable:
            self.comm.end_deferred()

    # ------------------------------------------------------------------ inputs
    def load_input(self, net: Net, src: torch.Tensor, dst, n):
        """Reference-side fp32 NCHW (or [B, features]) -> activation dtype NHWC rows."""
        c, h, w = net.input_chw
        if src.numel() != n * c * h * w:
            raise ValueError(f'input has {src.numel()} elements, expected {n}x{c}x{h}x{w}')
        self.ops.nchw_to_nhwc(src.contiguous(), dst, n, c, h, w)

    def _global_batch(self, B):
        return B * (self.comm.world_size if self.comm is not None else 1)

    # ------------------------------------------------------------------ head
    def _head_forward(self, st: NetState, feats, n, out_index, out):
        F = st.net.feature_size
        bias = st.hbias if st.net.head_parts is not None else st.params[st.net.head + '.bias']
        self.ops.rowdot(feats, n, F, st.whead[out_index * F:(out_index + 1) * F], bias, out_index, out)

    def _head_grad_row(self, st: NetState, out_index):
        """Gradient slice of row `out_index` of the [outputs][F] head weight (one tensor, or the region the parts share)."""
        F = st.net.feature_size
        if st.net.head_parts is not None:
            return st.grad[st.head_gbase + out_index * F:st.head_gbase + (out_index + 1) * F]
        return st.g(st.net.head + '.weight')[out_index * F:(out_index + 1) * F]

    def _head_grads(self, st: NetState, feats, n, out_index, drow):
        F = st.net.feature_size
        parts = st.net.parts()
        self.ops.colsum(feats, n, F, self._head_grad_row(st, out_index), 0, drow)
        for h, _ in parts:
            self.ops.colsum(drow, n, 1, st.g(h + '.bias')[out_index:out_index + 1], 0, None)

    # ------------------------------------------------------------------ SGAN K-logit head (sgan.py:18-67; csrc/sgan.cu)
    def _logits(self, st: NetState, feats, n, name, bias=True):
        """[K, n] fp32 class logits of `n` feature rows (bias=False: the tangent of the logits, W u_L)."""
        net = st.net
        out = self.buf(name, (net.head_outputs, n), self.mdt)
        b = (st.hbias if net.head_parts is not None else st.params[net.head + '.bias']) if bias else None
        ws = self.buf('lg_ws', (self.ops.head_logits_workspace(n, net.feature_size, net.head_outputs),), self.mdt)
        self.ops.head_logits(feats, n, net.feature_size, st.whead, b, net.head_outputs, out, ws)
        return out

    def _sgan_bins(self, cfg):
        bins = tuple(float(v) for v in cfg.bins)
        if len(bins) != self.d_net.head_outputs:
            raise ValueError(f'SGAN: {len(bins)} bins for a discriminator with {self.d_net.head_outputs} class logits')
        t = getattr(self, '_bins_cache', None)
        if t is None or t[0] != bins:
            t = self._bins_cache = (bins, torch.tensor(bins, dtype=torch.float32, device=self.device))
        return t[1]

    def _sgan_head_grads(self, st: NetState, feats, n, dT, bias=True):
        """dW_k += sum_r dT[k][r] * feats[r,:] (and db_k += sum_r dT[k][r]) for the K outputs of the head."""
        F, K = st.net.feature_size, st.net.head_outputs
        if st.net.head_parts is None and F % 4 == 0 and os.environ.get('SRGAN_NO_HEAD_WGRAD', '0') != '1':
            # one [K, F] head tensor: a single pass over the feature rows
            self.ops.head_wgrad(feats, n, F, dT, K, st.g(st.net.head + '.weight'), st.g(st.net.head + '.bias') if bias else None)
            return
        for k in range(K):
            if bias:
                self._head_grads(st, feats, n, k, dT[k])
            else:
                self.ops.colsum(feats, n, F, self._head_grad_row(st, k), 0, dT[k])

    def _sgan_labeled(self, st: NetState, feats, dfeats, n, y, cfg, Bg, loss_slot, tag):
        """sgan.py:20-31: cross entropy of the K logits against the bin of the real label, seeds the backward pass."""
        net, ops = st.net, self.ops
        if isinstance(y, (tuple, list)):
            raise ValueError('SGAN labels are real numbers [B] (age/sgan.py, coefficient/sgan.py)')
        K, F = net.head_outputs, net.feature_size
        fact, fslope = net.feature_act
        lg = self._logits(st, feats, n, 'lg_' + tag)
        dl = self.buf('dlg_' + tag, (K, n), self.mdt)
        ops.sgan_loss(lg, K, n, 0, y, self._sgan_bins(cfg), 0.0, cfg.labeled_loss_multiplier / Bg, loss_slot, dl)
        ops.seed_rows_multi(dfeats, n, F, dl, st.whead, K, feats, fact, fslope)
        self._sgan_head_grads(st, feats, n, dl)

    # ------------------------------------------------------------------ labeled loss (srgan.py:414-417 | crowd/srgan.py:247-254)
    def _labeled(self, st: NetState, acts, deltas, B, y, cfg, Bg, loss_slot, pred, dpred):
        """Labeled loss on rows [0,B): writes the loss scalar and dLoss/dprediction; for the crowd application also
        returns the hook that adds the map-loss gradient into the three map buffers' deltas during the backward pass."""
        net, ops = st.net, self.ops
        scale = cfg.labeled_loss_multiplier / Bg
        self._head_forward(st, self._brows_feat(net, acts, 0, B), B, 0, pred)
        if net.family != 'crowd':
            if isinstance(y, (tuple, list)):
                raise ValueError('tuple labels are the crowd application\'s (density, map) pair')
            ops.labeled_loss(pred, y, B, cfg.labeled_loss_order, scale, loss_slot, dpred)
            return None
        density, map_label = y
        HW = net.label_size * net.label_size
        if density.numel() != B * HW or map_label.numel() != B * HW:
            raise ValueError(f'crowd labels must be (density, map) of [{B}, {net.label_size}, {net.label_size}] each')
        maps = [self._brows(acts[m], net.bufs[m], 0, B) for m in net.map_bufs]
        dm = self.buf('dmap_scale', (B,), self.mdt)
        ops.crowd_loss(pred, density, maps, map_label, B, HW, cfg.labeled_loss_order, scale, cfg.map_multiplier,
                       loss_slot, dpred, dm)

        def hook(name):
            b = net.bufs[name]
            ops.crowd_map_grad(self._brows(acts[name], b, 0, B), map_label, dm, self._brows(deltas[name], b, 0, B), B, HW,
                               len(net.map_bufs), b.act, b.slope)
        return hook

    def _brows_feat(self, net: Net, t, lo, hi):
        return self.rows(self._feat(net, t), net.feature_size, lo, hi)

    # ------------------------------------------------------------------ DNN step (srgan.py:259-271)
    def dnn_step(self, x, y, cfg, lr, weight_decay):
        self.ops.begin()
        self._scope = 'dnn'
        st, net = self.DNN, self.d_net
        B = x.shape[0]
        Bg = self._global_batch(B)
        F = net.feature_size
        fact, fslope = net.feature_act
        acts = self.alloc_acts('D', net, B)
        deltas = self.alloc_deltas('D', net, B)
        self.load_input(net, x, self.rows(self._in(net, acts), net.in_elems, 0, B), B)
        self.forward(st, acts, 0, B, keep_pre=B)
        feats = self._brows_feat(net, acts, 0, B)
        pred = self.buf('pred', (B,), self.mdt)
        dpred = self.buf('dpred', (B,), self.mdt)
        self.scalars[SC_DNN:SC_DNN + 1].zero_()
        if cfg.method == 'sgan':
            self._sgan_labeled(st, feats, self._brows_feat(net, deltas, 0, B), B, y, cfg, Bg, self.scalars[SC_DNN:SC_DNN + 1], 'dnn')
            hook = None
        else:
            hook = self._labeled(st, acts, deltas, B, y, cfg, Bg, self.scalars[SC_DNN:SC_DNN + 1], pred, dpred)
            self.ops.seed_rows(self._brows_feat(net, deltas, 0, B), B, F, None, dpred, st.whead[0:F], feats, fact, fslope)
            self._head_grads(st, feats, B, 0, dpred)
        self.backward(st, acts, deltas, 0, B, hook=hook)
        self.adam(st, lr, weight_decay, cfg.betas, cfg.eps, deferrable='DNN')

    # ------------------------------------------------------------------ GAN step (srgan.py:273-320)
    def gan_step(self, x, y, u, z, alpha, z2, cfg, train_generator=True):
        ops, D, G, net, gnet = self.ops, self.D, self.G, self.d_net, self.g_net
        ops.begin()
        self._scope = 'gan'
        B = x.shape[0]
        if u.shape[0] != B or z.shape[0] != B or alpha.numel() != B:
            # srgan.py:363 draws alpha with settings.batch_size rows: the reference itself requires full batches
            raise ValueError('labeled, unlabeled and noise batches must have the same size (SURVEY App. E.2)')
        Bg = self._global_batch(B)
        F = net.feature_size
        fact, fslope = net.feature_act
        dggan = cfg.method == 'dggan'
        acts = self.alloc_acts('D', net, 5 * B)
        deltas = self.alloc_deltas('D', net, 5 * B)
        a_in = self._in(net, acts)
        E = net.in_elems
        sc = self.scalars
        sc[SC_LABELED:SC_GEN + 1].zero_()

        def fblk(lo, hi):
            return self._brows_feat(net, acts, lo, hi)

        def dblk(lo, hi):
            return self._brows_feat(net, deltas, lo, hi)

        # ---- inputs: x, u, fake = G(z) (no grad, srgan.py:290/352), x_hat (srgan.py:362-366)
        self.load_input(net, x, self.rows(a_in, E, 0, B), B)
        self.load_input(net, u, self.rows(a_in, E, B, 2 * B), B)
        gacts = self.alloc_acts('G', gnet, B)
        gacts[-1] = self.rows(a_in, E, 2 * B, 3 * B)
        self.load_input(gnet, z, gacts[0], B)
        self.forward(G, gacts, 0, B)
        ops.interpolate(self.rows(a_in, E, B, 2 * B), self.rows(a_in, E, 2 * B, 3 * B), alpha,
                        self.rows(a_in, E, 3 * B, 4 * B), B, E)
        self.last_gan_batch = B
        # ---- one D forward over [x; u; fake; x_hat]
        self.forward(D, acts, 0, 4 * B, keep_pre=3 * B)
        if self.publish_features and cfg.method == 'srgan':
            self.buf('feat_snap', (4 * B * F,)).copy_(fblk(0, 4 * B))
        # ---- labeled loss (srgan.py:329-335, :414-417)
        pred = self.buf('pred', (B,), self.mdt)
        dpred = self.buf('dpred', (B,), self.mdt)
        sgan = cfg.method == 'sgan'
        hook = None if sgan else self._labeled(D, acts, deltas, B, y, cfg, Bg, sc[SC_LABELED:SC_LABELED + 1], pred, dpred)
        gamma_L = dblk(4 * B, 5 * B)
        s_norm = self.buf('s_norm', (B,), self.mdt)
        if sgan:
            # ---- SGAN (sgan.py:20-58): cross entropy on x; BCE(logsumexp(logits), 1 | 0) on u | fake, both times the
            # matching multiplier; GP target = BCE(logsumexp(logits(x_hat)), 0) * penalty multiplier, a scalar over the batch
            K = net.head_outputs
            self._sgan_labeled(D, fblk(0, B), dblk(0, B), B, y, cfg, Bg, sc[SC_LABELED:SC_LABELED + 1], 'x')
            gp_c = cfg.gradient_penalty_multiplier / Bg
            lg_h = self._logits(D, fblk(3 * B, 4 * B), B, 'lg_h')
            s_h = self.buf('dlg_h', (K, B), self.mdt)
            for tag, lo, target, slot in (('u', B, 1.0, SC_UNLABELED), ('f', 2 * B, 0.0, SC_FAKE)):
                lg = self._logits(D, fblk(lo, lo + B), B, 'lg_' + tag)
                dl = self.buf('dlg_' + tag, (K, B), self.mdt)
                ops.sgan_loss(lg, K, B, 1, None, None, target, cfg.matching_loss_multiplier / Bg, sc[slot:slot + 1], dl)
                ops.seed_rows_multi(dblk(lo, lo + B), B, F, dl, D.whead, K, fblk(lo, lo + B), fact, fslope)
                self._sgan_head_grads(D, fblk(lo, lo + B), B, dl)
            ops.sgan_loss(lg_h, K, B, 1, None, None, 0.0, gp_c, None, s_h)          # s = d(interpolates_loss)/d(logits)
            ops.seed_rows_multi(gamma_L, B, F, s_h, D.whead, K, fblk(3 * B, 4 * B), fact, fslope)
        elif not dggan:
            # ---- feature sums -> (all-reduce) -> distance losses (srgan.py:337-358, :438-449)
            sums = self.buf('fsums', (3, F), self.mdt)
            sums.zero_()
            for j in range(3):
                ops.colsum(fblk(j * B, (j + 1) * B), B, F, sums[j], 0, None)
            if self.comm is not None:
                self.comm.all_reduce_sum(sums)
            gvec = self.buf('gvec', (3, F), self.mdt)
            inv = 1.0 / Bg
            # unlabeled: base = u, other = x ; writes d/dmean_u and d/dmean_x (already divided by the global batch)
            ops.distance(sums[1], sums[0], F, inv, DIST_KINDS[cfg.matching_distance_function],
                         cfg.matching_loss_multiplier * cfg.srgan_loss_multiplier, sc[SC_UNLABELED:SC_UNLABELED + 1],
                         gvec[1], gvec[0], False)
            # fake: base = u (accumulates into gvec[1]), other = fake
            ops.distance(sums[1], sums[2], F, inv, DIST_KINDS[cfg.contrasting_distance_function],
                         cfg.contrasting_loss_multiplier * cfg.srgan_loss_multiplier, sc[SC_FAKE:SC_FAKE + 1],
                         gvec[1], gvec[2], True)
            ops.seed_rows(dblk(0, B), B, F, gvec[0], dpred, D.whead[0:F], fblk(0, B), fact, fslope)
            ops.seed_rows(dblk(B, 2 * B), B, F, gvec[1], None, None, fblk(B, 2 * B), fact, fslope)
            ops.seed_rows(dblk(2 * B, 3 * B), B, F, gvec[2], None, None, fblk(2 * B, 3 * B), fact, fslope)
            # ---- GP target s = ||f(x_hat)||_2 (srgan.py:377-381): gamma_L = (f/s) * act'
            ops.feature_norm_seed(fblk(3 * B, 4 * B), B, F, s_norm, gamma_L, fact, fslope)
        else:
            # ---- DG-GAN (coefficient/dggan.py:36-57): BCE on the second head output; GP target = raw score
            su = self.buf('score_u', (B,), self.mdt)
            sf = self.buf('score_f', (B,), self.mdt)
            dsu = self.buf('dscore_u', (B,), self.mdt)
            dsf = self.buf('dscore_f', (B,), self.mdt)
            self._head_forward(D, fblk(B, 2 * B), B, 1, su)
            self._head_forward(D, fblk(2 * B, 3 * B), B, 1, sf)
            ops.bce_logits(su, B, 0.0, cfg.matching_loss_multiplier * cfg.dggan_loss_multiplier / Bg,
                           sc[SC_UNLABELED:SC_UNLABELED + 1], dsu)
            ops.bce_logits(sf, B, 1.0, cfg.contrasting_loss_multiplier * cfg.dggan_loss_multiplier / Bg,
                           sc[SC_FAKE:SC_FAKE + 1], dsf)
            w1 = D.whead[F:2 * F]
            ops.seed_rows(dblk(0, B), B, F, None, dpred, D.whead[0:F], fblk(0, B), fact, fslope)
            ops.seed_rows(dblk(B, 2 * B), B, F, None, dsu, w1, fblk(B, 2 * B), fact, fslope)
            ops.seed_rows(dblk(2 * B, 3 * B), B, F, None, dsf, w1, fblk(2 * B, 3 * B), fact, fslope)
            self._head_grads(D, fblk(B, 2 * B), B, 1, dsu)
            self._head_grads(D, fblk(2 * B, 3 * B), B, 1, dsf)
            ops.seed_rows(gamma_L, B, F, w1, None, None, fblk(3 * B, 4 * B), fact, fslope)
        if not sgan:
            self._head_grads(D, fblk(0, B), B, 0, dpred)
        # ---- gradient penalty (srgan.py:360-375) without autograd: SURVEY App. C.3
        g0 = self.buf('g0', (B * E,))
        self.gchain(D, acts, deltas, B, g0)
        gnorm = self.buf('gnorm', (B,), self.mdt)
        u0 = self.rows(a_in, E, 4 * B, 5 * B)
        ops.gradnorm_penalty(g0, B, E, cfg.gradient_penalty_multiplier / Bg, 1.0 / Bg, gnorm, sc[SC_GP:SC_GP + 1],
                             sc[SC_GNORM:SC_GNORM + 1], u0)
        self.tangent(D, acts, B)
        if sgan:
            # <u0, g> = sum_k s_k (W_k . u_L): its parameter gradient has three parts -- through the Jacobian (the tangent
            # block of the weight-gradient launches, like SR-GAN), through W explicitly (s_k-weighted sums of the tangent
            # features), and through s(logits(x_hat)): q = Hessian . (W u_L) is an ordinary backward seed on the x_hat rows
            lg_t = self._logits(D, fblk(4 * B, 5 * B), B, 'lg_t', bias=False)
            q = self.buf('q_h', (K, B), self.mdt)
            ops.sgan_gp_second(lg_h, lg_t, K, B, gp_c, q)
            ops.seed_rows_multi(dblk(3 * B, 4 * B), B, F, q, D.whead, K, fblk(3 * B, 4 * B), fact, fslope)
            self._sgan_head_grads(D, fblk(3 * B, 4 * B), B, q)
            self._sgan_head_grads(D, fblk(4 * B, 5 * B), B, s_h, bias=False)
            self.backward(D, acts, deltas, 0, 4 * B, 0, 5 * B, hook=None)
        elif not dggan:
            ops.gp_feature_seed(fblk(4 * B, 5 * B), fblk(3 * B, 4 * B), s_norm, dblk(3 * B, 4 * B), B, F, fact, fslope)
            # ---- one backward over [x; u; fake; x_hat], weight gradients also over the tangent block
            self.backward(D, acts, deltas, 0, 4 * B, 0, 5 * B, hook=hook)
        else:
            # dP/dW_head[1,:] = sum_n u_L,n ; the Jacobian term vanishes (target is linear in the features)
            ops.colsum(fblk(4 * B, 5 * B), B, F, self._head_grad_row(D, 1), 0, None)
            # DG-GAN: rows [3B,4B) carry no ordinary gradient; weight grads = rows [0,3B) + tangent block [4B,5B)
            self.backward(D, acts, deltas, 0, 3 * B, 0, 3 * B, hook=hook)
            self.tangent_block_grads(D, acts, deltas, 4 * B, 5 * B)
        self.adam(D, cfg.learning_rate, cfg.weight_decay, cfg.betas, cfg.eps)                     # srgan.py:297
        if train_generator:
            self._g_step(acts, deltas, B, Bg, z2, cfg, fblk, dblk)

    def _g_step(self, acts, deltas, B, Bg, z2, cfg, fblk, dblk):
        """The generator step of gan_step (srgan.py:299-305, :383-391) with the UPDATED discriminator."""
        ops, D, G, net, gnet = self.ops, self.D, self.G, self.d_net, self.g_net
        F = net.feature_size
        E = net.in_elems
        fact, fslope = net.feature_act
        dggan = cfg.method == 'dggan'
        sc = self.scalars
        a_in = self._in(net, acts)
        gacts = self.alloc_acts('G', gnet, B)
        fake2 = self.rows(a_in, E, 0, B)
        gacts[-1] = fake2
        self.load_input(gnet, z2, gacts[0], B)
        self.forward(G, gacts, 0, B)
        if cfg.method == 'sgan':
            # sgan.py:60-67: generator loss = -BCE(logsumexp(logits(G(z2))), 0), no multiplier
            K = net.head_outputs
            self.forward(D, acts, 0, B, keep_pre=False)
            lg = self._logits(D, fblk(0, B), B, 'lg_x')
            dl = self.buf('dlg_x', (K, B), self.mdt)
            ops.sgan_loss(lg, K, B, 1, None, None, 0.0, -1.0 / Bg, sc[SC_GEN:SC_GEN + 1], dl)
            ops.seed_rows_multi(dblk(0, B), B, F, dl, D.whead, K, fblk(0, B), fact, fslope)
        elif not dggan:
            self.forward(D, acts, 0, 2 * B, keep_pre=False)     # rows [B,2B) still hold u
            if self.publish_features:
                self.buf('feat_snap', (4 * B * F,))[2 * B * F:3 * B * F].copy_(fblk(0, B))      # srgan.py:386
            sums = self.buf('fsums', (3, F), self.mdt)
            sums.zero_()
            ops.colsum(fblk(0, B), B, F, sums[0], 0, None)
            ops.colsum(fblk(B, 2 * B), B, F, sums[1], 0, None)
            if self.comm is not None:
                self.comm.all_reduce_sum(sums)
            gvec = self.buf('gvec', (3, F), self.mdt)
            ops.distance(sums[1], sums[0], F, 1.0 / Bg, DIST_KINDS[cfg.matching_distance_function],
                         cfg.matching_loss_multiplier, sc[SC_GEN:SC_GEN + 1], gvec[1], gvec[0], False)
            ops.seed_rows(dblk(0, B), B, F, gvec[0], None, None, fblk(0, B), fact, fslope)
        else:
            self.forward(D, acts, 0, B, keep_pre=False)
            sf = self.buf('score_f', (B,), self.mdt)
            dsf = self.buf('dscore_f', (B,), self.mdt)
            self._head_forward(D, fblk(0, B), B, 1, sf)
            ops.bce_logits(sf, B, 0.0, 1.0 / Bg, sc[SC_GEN:SC_GEN + 1], dsf)                      # dggan.py:59-64
            ops.seed_rows(dblk(0, B), B, F, None, dsf, D.whead[F:2 * F], fblk(0, B), fact, fslope)
        gl = gnet.layers
        gdeltas = self.alloc_deltas('G', gnet, B)
        # D data-backward only (the reference's D weight grads here are discarded), ending in
        # dLoss/d(pre-activation of G's last layer) = (W_1^T delta_1) * act_G'(fake2)
        self.backward(D, acts, deltas, 0, B, need_input_grad=True, dinput=gdeltas[len(gl)], input_href=gacts[-1],
                      input_act=gl[-1].act, weight_grads=False)
        self.backward(G, gacts, gdeltas, 0, B)
        self.adam(G, cfg.learning_rate, 0.0, cfg.betas, cfg.eps, deferrable='G')                  # srgan.py:137, :305

    # ------------------------------------------------------------------ micro-batched steps
    # Activation memory is 5 row blocks per buffer: the crowd discriminator needs ~0.7 GB per sample of the local batch, so a
    # per-GPU batch of 512 (global 4096 on 8 GPUs, BASELINE configs[4]) cannot be held at once.  Micro-batching is EXACT
    # for SR-GAN because the feature losses see the activations only through the batch mean (hard parts):
    # pass 1 runs the forwards micro-batch by micro-batch and accumulates the feature sums; the sums are all-reduced and
    # the distance losses give d(loss)/d(mean); pass 2 re-runs each micro-batch (forward, gradient penalty, backward) with
    # that known seed and accumulates the parameter gradients; one Adam update at the end.  Cost: one extra forward.
    @staticmethod
    def _ysl(y, r0, r1):
        return tuple(t[r0:r1] for t in y) if isinstance(y, (tuple, list)) else y[r0:r1]

    def dnn_step_micro(self, x, y, cfg, lr, weight_decay, mb):
        if cfg.method == 'sgan':
            raise NotImplementedError('micro-batched steps cover srgan and dggan')
        self.ops.begin()
        self._scope = 'dnn'
        st, net = self.DNN, self.d_net
        B = x.shape[0]
        if B % mb:
            raise ValueError(f'local batch {B} is not a multiple of the micro-batch {mb}')
        Bg = self._global_batch(B)
        F = net.feature_size
        fact, fslope = net.feature_act
        acts = self.alloc_acts('D', net, mb)
        deltas = self.alloc_deltas('D', net, mb)
        pred = self.buf('pred', (mb,), self.mdt)
        dpred = self.buf('dpred', (mb,), self.mdt)
        self.scalars[SC_DNN:SC_DNN + 1].zero_()
        for m in range(B // mb):
            r0, r1 = m * mb, (m + 1) * mb
            self.load_input(net, x[r0:r1], self.rows(self._in(net, acts), net.in_elems, 0, mb), mb)
            self.forward(st, acts, 0, mb, keep_pre=mb)
            feats = self._brows_feat(net, acts, 0, mb)
            hook = self._labeled(st, acts, deltas, mb, self._ysl(y, r0, r1), cfg, Bg, self.scalars[SC_DNN:SC_DNN + 1], pred, dpred)
            self.ops.seed_rows(self._brows_feat(net, deltas, 0, mb), mb, F, None, dpred, st.whead[0:F], feats, fact, fslope)
            self._head_grads(st, feats, mb, 0, dpred)
            self.backward(st, acts, deltas, 0, mb, hook=hook)
        self.adam(st, lr, weight_decay, cfg.betas, cfg.eps, deferrable='DNN')

    def gan_step_micro(self, x, y, u, z, alpha, z2, cfg, train_generator, mb):
        if cfg.method == 'sgan':
            raise NotImplementedError('micro-batched steps cover srgan and dggan')
        ops, D, G, net, gnet = self.ops, self.D, self.G, self.d_net, self.g_net
        ops.begin()
        self._scope = 'gan'
        B = x.shape[0]
        if u.shape[0] != B or z.shape[0] != B or alpha.numel() != B:
            raise ValueError('labeled, unlabeled and noise batches must have the same size (SURVEY App. E.2)')
        if B % mb:
            raise ValueError(f'local batch {B} is not a multiple of the micro-batch {mb}')
        if cfg.method == 'dggan':
            raise NotImplementedError('micro-batching is implemented for the SR-GAN feature losses only')
        b, nm = mb, B // mb
        Bg = self._global_batch(B)
        F = net.feature_size
        fact, fslope = net.feature_act
        acts = self.alloc_acts('D', net, 5 * b)
        deltas = self.alloc_deltas('D', net, 5 * b)
        a_in = self._in(net, acts)
        E = net.in_elems
        alpha = alpha.reshape(-1)
        sc = self.scalars
        sc[SC_LABELED:SC_GEN + 1].zero_()

        def fblk(lo, hi):
            return self._brows_feat(net, acts, lo, hi)

        def dblk(lo, hi):
            return self._brows_feat(net, deltas, lo, hi)

        def d_inputs(m, with_hat):
            """rows [0,b) = x_m, [b,2b) = u_m, [2b,3b) = G(z_m), [3b,4b) = x_hat_m"""
            r0, r1 = m * b, (m + 1) * b
            self.load_input(net, x[r0:r1], self.rows(a_in, E, 0, b), b)
            self.load_input(net, u[r0:r1], self.rows(a_in, E, b, 2 * b), b)
            gacts = self.alloc_acts('G', gnet, b)
            gacts[-1] = self.rows(a_in, E, 2 * b, 3 * b)
            self.load_input(gnet, z[r0:r1], gacts[0], b)
            self.forward(G, gacts, 0, b)
            if with_hat:
                ops.interpolate(self.rows(a_in, E, b, 2 * b), self.rows(a_in, E, 2 * b, 3 * b), alpha[r0:r1],
                                self.rows(a_in, E, 3 * b, 4 * b), b, E)
        # ---- D step, pass 1: feature sums of x, u, fake over all micro-batches
        sums = self.buf('fsums', (3, F), self.mdt)
        sums.zero_()
        for m in range(nm):
            d_inputs(m, False)
            self.forward(D, acts, 0, 3 * b, keep_pre=False)
            for j in range(3):
                ops.colsum(fblk(j * b, (j + 1) * b), b, F, sums[j], 0, None)
        if self.comm is not None:
            self.comm.all_reduce_sum(sums)
        gvec = self.buf('gvec', (3, F), self.mdt)
        inv = 1.0 / Bg
        ops.distance(sums[1], sums[0], F, inv, DIST_KINDS[cfg.matching_distance_function],
                     cfg.matching_loss_multiplier * cfg.srgan_loss_multiplier, sc[SC_UNLABELED:SC_UNLABELED + 1],
                     gvec[1], gvec[0], False)
        ops.distance(sums[1], sums[2], F, inv, DIST_KINDS[cfg.contrasting_distance_function],
                     cfg.contrasting_loss_multiplier * cfg.srgan_loss_multiplier, sc[SC_FAKE:SC_FAKE + 1],
                     gvec[1], gvec[2], True)
        # ---- D step, pass 2: per micro-batch forward over [x; u; fake; x_hat], gradient penalty, one backward
        pred = self.buf('pred', (b,), self.mdt)
        dpred = self.buf('dpred', (b,), self.mdt)
        s_norm = self.buf('s_norm', (b,), self.mdt)
        g0 = self.buf('g0', (b * E,))
        gnorm = self.buf('gnorm', (b,), self.mdt)
        for m in range(nm):
            r0, r1 = m * b, (m + 1) * b
            d_inputs(m, True)
            self.forward(D, acts, 0, 4 * b, keep_pre=3 * b)
            hook = self._labeled(D, acts, deltas, b, self._ysl(y, r0, r1), cfg, Bg, sc[SC_LABELED:SC_LABELED + 1], pred, dpred)
            ops.seed_rows(dblk(0, b), b, F, gvec[0], dpred, D.whead[0:F], fblk(0, b), fact, fslope)
            ops.seed_rows(dblk(b, 2 * b), b, F, gvec[1], None, None, fblk(b, 2 * b), fact, fslope)
            ops.seed_rows(dblk(2 * b, 3 * b), b, F, gvec[2], None, None, fblk(2 * b, 3 * b), fact, fslope)
            ops.feature_norm_seed(fblk(3 * b, 4 * b), b, F, s_norm, dblk(4 * b, 5 * b), fact, fslope)
            self._head_grads(D, fblk(0, b), b, 0, dpred)
            self.gchain(D, acts, deltas, b, g0)
            ops.gradnorm_penalty(g0, b, E, cfg.gradient_penalty_multiplier / Bg, 1.0 / Bg, gnorm, sc[SC_GP:SC_GP + 1],
                                 sc[SC_GNORM:SC_GNORM + 1], self.rows(a_in, E, 4 * b, 5 * b))
            self.tangent(D, acts, b)
            ops.gp_feature_seed(fblk(4 * b, 5 * b), fblk(3 * b, 4 * b), s_norm, dblk(3 * b, 4 * b), b, F, fact, fslope)
            self.backward(D, acts, deltas, 0, 4 * b, 0, 5 * b, hook=hook)
        self.adam(D, cfg.learning_rate, cfg.weight_decay, cfg.betas, cfg.eps)
        if not train_generator:
            return
        # ---- G step with the updated D, pass 1: feature sums of G(z2) and u
        gl = gnet.layers

        def g_inputs(m):
            r0, r1 = m * b, (m + 1) * b
            gacts = self.alloc_acts('G', gnet, b)
            gacts[-1] = self.rows(a_in, E, 0, b)
            self.load_input(gnet, z2[r0:r1], gacts[0], b)
            self.forward(G, gacts, 0, b)
            return gacts
        sums.zero_()
        for m in range(nm):
            g_inputs(m)
            self.load_input(net, u[m * b:(m + 1) * b], self.rows(a_in, E, b, 2 * b), b)
            self.forward(D, acts, 0, 2 * b, keep_pre=False)
            ops.colsum(fblk(0, b), b, F, sums[0], 0, None)
            ops.colsum(fblk(b, 2 * b), b, F, sums[1], 0, None)
        if self.comm is not None:
            self.comm.all_reduce_sum(sums)
        ops.distance(sums[1], sums[0], F, 1.0 / Bg, DIST_KINDS[cfg.matching_distance_function],
                     cfg.matching_loss_multiplier, sc[SC_GEN:SC_GEN + 1], gvec[1], gvec[0], False)
        # ---- G step, pass 2: per micro-batch G forward, D forward on the fake rows, D data-backward, G backward
        gdeltas = self.alloc_deltas('G', gnet, b)
        for m in range(nm):
            gacts = g_inputs(m)
            self.forward(D, acts, 0, b, keep_pre=False)
            ops.seed_rows(dblk(0, b), b, F, gvec[0], None, None, fblk(0, b), fact, fslope)
            self.backward(D, acts, deltas, 0, b, need_input_grad=True, dinput=gdeltas[len(gl)], input_href=gacts[-1],
                          input_act=gl[-1].act, weight_grads=False)
            self.backward(G, gacts, gdeltas, 0, b)
        self.adam(G, cfg.learning_rate, 0.0, cfg.betas, cfg.eps, deferrable='G')

    # ------------------------------------------------------------------ inference-style helpers
    def d_features(self, x, st: Optional[NetState] = None, with_maps=False):
        """D(x) forward only: returns (prediction [B], features rows [B, F] in NHWC order)."""
        self.ops.begin()
        st = st or self.D
        self._scope = 'dnn' if st is self.DNN else 'gan'       # the forward shares the scratch of its network's step
        net = st.net
        B = x.shape[0]
        acts = self.alloc_acts('D', net, B if st is self.DNN else 5 * B)
        self.load_input(net, x, self.rows(self._in(net, acts), net.in_elems, 0, B), B)
        self.forward(st, acts, 0, B, keep_pre=False)
        feats = self._brows_feat(net, acts, 0, B)
        pred = self.buf('pred', (B,), self.mdt)
        self._head_forward(st, feats, B, 0, pred)
        if with_maps:           # crowd: the three predicted maps, [B, label_size^2] each (crowd/models.py:1155-1165)
            return pred, feats, [self._brows(acts[m], net.bufs[m], 0, B) for m in net.map_bufs]
        return pred, feats

    def g_generate(self, z):
        self.ops.begin()
        self._scope = 'gan'
        gnet = self.g_net
        B = z.shape[0]
        gacts = self.alloc_acts('G', gnet, B)
        self.load_input(gnet, z, gacts[0], B)
        self.forward(self.G, gacts, 0, B)
        return gacts[-1]
