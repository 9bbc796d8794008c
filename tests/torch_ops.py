"""
TEST-ONLY emulation of the C-ABI op set (include/srgan_b200.h) with PyTorch on CPU.

Purpose: check the explicit schedule in sr-gan_b200/engine.py (de-duplicated backward, gradient-penalty chains,
Adam, layout strides) against the oracle on CPU, in fp32 or fp64, without a GPU.  It is never imported by the
product package; sr-gan_b200/ops_cuda.py is the only ops implementation the product has and it raises when the CUDA
library is missing.  On the GPU box the same op-by-op semantics are what tests/test_gpu_kernels.py checks the CUDA
kernels against.
"""
import torch
import torch.nn.functional as F

ACT_NONE, ACT_LEAKY, ACT_TANH = 0, 1, 2
EPI_BIAS_ACT, EPI_DACT = 0, 1


def dact(href, act, slope):
    if act == ACT_LEAKY:
        return torch.where(href > 0, torch.ones_like(href), torch.full_like(href, slope))
    if act == ACT_TANH:
        return 1 - href * href
    return torch.ones_like(href)


def apply_act(x, act, slope):
    if act == ACT_LEAKY:
        return torch.where(x > 0, x, x * slope)
    if act == ACT_TANH:
        return torch.tanh(x)
    return x


class TorchOps:
    def __init__(self):
        self.launches = 0

    def begin(self):
        pass

    # ---- layouts
    @staticmethod
    def _scatter(w, dims, out, strides):
        d0, d1, d2, d3 = dims
        w4 = w.detach().reshape(d0, d1, d2, d3)
        idx = (torch.arange(d0).view(-1, 1, 1, 1) * strides[0] + torch.arange(d1).view(1, -1, 1, 1) * strides[1]
               + torch.arange(d2).view(1, 1, -1, 1) * strides[2] + torch.arange(d3).view(1, 1, 1, -1) * strides[3])
        out[idx.reshape(-1)] = w4.reshape(-1).to(out.dtype)

    @staticmethod
    def _gather(src, dims, strides):
        d0, d1, d2, d3 = dims
        idx = (torch.arange(d0).view(-1, 1, 1, 1) * strides[0] + torch.arange(d1).view(1, -1, 1, 1) * strides[1]
               + torch.arange(d2).view(1, 1, -1, 1) * strides[2] + torch.arange(d3).view(1, 1, 1, -1) * strides[3])
        return src[idx.reshape(-1)].reshape(d0, d1, d2, d3)

    def repack(self, w, dims, out1, s1, out2, s2):
        self.launches += 1
        if out1 is not None:
            self._scatter(w, dims, out1, s1)
        if out2 is not None:
            self._scatter(w, dims, out2, s2)

    # ---- contractions (weights given in kernel layout; converted back to torch layout for F.conv*)
    @staticmethod
    def _w_from_wd(Wd, g):
        return Wd.view(g.Ca, g.R, g.S, g.Cb).permute(0, 3, 1, 2)        # [a, b, r, s]

    @staticmethod
    def _w_from_wu(Wu, g):
        return Wu.view(g.Cb, g.R, g.S, g.Ca).permute(3, 0, 1, 2)        # [a, b, r, s]

    def _epilogue(self, acc_nhwc, out, bias, bias_mod, href, epi, act, slope):
        flat = acc_nhwc.reshape(-1)
        C = acc_nhwc.shape[-1]
        if epi == EPI_BIAS_ACT:
            if bias is not None:
                if bias_mod:
                    b = bias.detach().to(flat.dtype).repeat(C // bias_mod)
                else:
                    b = bias.detach().to(flat.dtype)
                flat = (acc_nhwc + b).reshape(-1)
            out.copy_(apply_act(flat, act, slope).to(out.dtype))
        else:
            if href is not None and act != ACT_NONE:
                flat = flat * dact(href.to(flat.dtype), act, slope)
            out.copy_(flat.to(out.dtype))

    def conv_down(self, L, Wd, S_out, n, g, bias, bias_mod, href, epi, act, slope):
        self.launches += 1
        cd = torch.float64 if L.dtype == torch.float64 else torch.float32
        x = L.view(n, g.Hl, g.Wl, g.Cb).permute(0, 3, 1, 2).to(cd)
        y = F.conv2d(x, self._w_from_wd(Wd, g).to(cd), None, stride=g.stride, padding=g.pad)
        self._epilogue(y.permute(0, 2, 3, 1), S_out, bias, bias_mod, href, epi, act, slope)

    def conv_up(self, S, Wu, L_out, n, g, bias, bias_mod, href, epi, act, slope):
        self.launches += 1
        cd = torch.float64 if S.dtype == torch.float64 else torch.float32
        x = S.view(n, g.Hs, g.Ws, g.Ca).permute(0, 3, 1, 2).to(cd)
        opad = g.Hl - ((g.Hs - 1) * g.stride - 2 * g.pad + g.R)
        y = F.conv_transpose2d(x, self._w_from_wu(Wu, g).to(cd), None, stride=g.stride, padding=g.pad,
                               output_padding=opad)
        self._epilogue(y.permute(0, 2, 3, 1), L_out, bias, bias_mod, href, epi, act, slope)

    def conv_wgrad(self, S, L, dW, n, g):
        self.launches += 1
        cd = dW.dtype
        s = S.view(n, g.Hs, g.Ws, g.Ca).permute(0, 3, 1, 2).to(cd)
        l = L.view(n, g.Hl, g.Wl, g.Cb).permute(0, 3, 1, 2).to(cd)
        gw = torch.nn.grad.conv2d_weight(l, (g.Ca, g.Cb, g.R, g.S), s, stride=g.stride, padding=g.pad)   # [a,b,r,s]
        dW += gw.permute(0, 2, 3, 1).reshape(-1)

    # ---- reductions / element-wise
    def colsum(self, X, rows, cols, out, mod, rowscale):
        self.launches += 1
        x = X[:rows * cols].view(rows, cols).to(out.dtype)
        if rowscale is not None:
            x = x * rowscale[:rows].view(rows, 1).to(out.dtype)
        s = x.sum(0)
        if mod:
            s = s.view(cols // mod, mod).sum(0)
        out += s

    def rowdot(self, X, rows, cols, w, bias, bias_index, out):
        self.launches += 1
        x = X[:rows * cols].view(rows, cols).to(out.dtype)
        out.copy_(x @ w.to(out.dtype) + bias.detach().reshape(-1)[bias_index].to(out.dtype))

    def seed_rows(self, out, rows, cols, gvec, rowscale, wrow, href, act, slope):
        self.launches += 1
        cd = href.dtype if href.dtype == torch.float64 else torch.float32
        v = torch.zeros(rows, cols, dtype=cd)
        if gvec is not None:
            v = v + gvec.to(cd).view(1, cols)
        if rowscale is not None:
            v = v + rowscale.to(cd).view(rows, 1) * wrow.to(cd).view(1, cols)
        v = v * dact(href[:rows * cols].view(rows, cols).to(cd), act, slope)
        out.copy_(v.reshape(-1).to(out.dtype))

    def nchw_to_nhwc(self, src, dst, n, c, h, w):
        self.launches += 1
        dst.copy_(src.detach().reshape(n, c, h, w).permute(0, 2, 3, 1).reshape(-1).to(dst.dtype))

    def interpolate(self, u, fake, alpha, out, n, E):
        self.launches += 1
        a = alpha.reshape(n, 1).to(torch.float64 if u.dtype == torch.float64 else torch.float32)
        r = a * u.view(n, E).to(a.dtype) + (1 - a) * fake.view(n, E).to(a.dtype)
        out.copy_(r.reshape(-1).to(out.dtype))

    def labeled_loss(self, pred, y, n, order, scale, loss_out, dpred):
        self.launches += 1
        d = pred - y.to(pred.dtype)
        loss_out += scale * d.abs().pow(order).sum()
        dpred.copy_(scale * order * d.abs().pow(order - 1) * torch.sign(d))

    def bce_logits(self, scores, n, target, scale, loss_out, dscore):
        self.launches += 1
        x = scores
        loss_out += scale * (torch.clamp(x, min=0) - x * target + torch.log1p(torch.exp(-x.abs()))).sum()
        dscore.copy_(scale * (torch.sigmoid(x) - target))

    def distance(self, sum_base, sum_other, Fdim, inv_B, kind, mult, loss_out, gbase, gother, accumulate_base):
        self.launches += 1
        d = (sum_base - sum_other) * inv_B
        if kind == 0:
            loss, g = d.abs().mean(), torch.sign(d) / Fdim
        elif kind == 1:
            loss, g = -d.abs().mean(), -torch.sign(d) / Fdim
        elif kind == 2:
            r = (d.abs() + 1).sqrt()
            loss, g = -r.mean(), -torch.sign(d) / (2 * r) / Fdim
        elif kind == 3:
            loss, g = -(d.abs() + 1).log().mean(), -torch.sign(d) / (d.abs() + 1) / Fdim
        elif kind == 4:
            loss, g = d.pow(2).mean(), 2 * d / Fdim
        elif kind == 5:
            nrm = d.pow(2).sum().sqrt()
            loss, g = nrm, d / nrm
        else:
            raise ValueError(kind)
        loss_out += mult * loss
        g = g * (mult * inv_B)
        if accumulate_base:
            gbase += g
        else:
            gbase.copy_(g)
        gother.copy_(-g)

    def feature_norm_seed(self, h, rows, cols, s_out, gamma_out, act, slope):
        self.launches += 1
        cd = s_out.dtype
        f = h[:rows * cols].view(rows, cols).to(cd)
        s = f.norm(dim=1)
        s_out.copy_(s)
        gamma_out.copy_(((f / s.view(rows, 1)) * dact(f, act, slope)).reshape(-1).to(gamma_out.dtype))

    def gradnorm_penalty(self, g0, n, E, lam_over_B, inv_B, gnorm_out, pen_out, gnmean_out, u0_out):
        self.launches += 1
        cd = gnorm_out.dtype
        g = g0[:n * E].view(n, E).to(cd)
        r = g.norm(dim=1)
        gnorm_out.copy_(r)
        ex = torch.clamp(r - 1, min=0)
        pen_out += lam_over_B * (ex ** 2).sum()
        gnmean_out += inv_B * r.sum()
        coef = torch.where(r > 0, 2 * lam_over_B * ex / r, torch.zeros_like(r))
        u0_out.copy_((g * coef.view(n, 1)).reshape(-1).to(u0_out.dtype))

    def gp_feature_seed(self, uL, hL, s, out, rows, cols, act, slope):
        self.launches += 1
        cd = s.dtype
        u = uL[:rows * cols].view(rows, cols).to(cd)
        f = hL[:rows * cols].view(rows, cols).to(cd)
        g = f / s.view(rows, 1)
        dot = (g * u).sum(1, keepdim=True)
        df = (u - g * dot) / s.view(rows, 1)
        out.copy_((df * dact(f, act, slope)).reshape(-1).to(out.dtype))

    def im2col(self, L, col, n, g, kpad):
        self.launches += 1
        x = L.view(n, g.Hl, g.Wl, g.Cb).permute(0, 3, 1, 2).to(torch.float64)
        u = F.unfold(x, (g.R, g.S), padding=g.pad, stride=g.stride)            # [n, Cb*R*S, P]  (b, r, s) order
        P = u.shape[-1]
        assert P == g.Hs * g.Ws
        u = u.view(n, g.Cb, g.R * g.S, P).permute(0, 3, 2, 1).reshape(n * P, g.R * g.S * g.Cb)   # (tap, b) order
        out = torch.zeros(n * P, kpad, dtype=torch.float64)
        out[:, :u.shape[1]] = u
        col.copy_(out.reshape(-1).to(col.dtype))

    def col2im(self, col, L_out, n, g, kpad, bias, href, epi, act, slope):
        self.launches += 1
        K = g.R * g.S * g.Cb
        P = g.Hs * g.Ws
        c = col.view(n * P, kpad)[:, :K].to(torch.float64).view(n, P, g.R * g.S, g.Cb).permute(0, 3, 2, 1)
        c = c.reshape(n, g.Cb * g.R * g.S, P)
        y = F.fold(c, (g.Hl, g.Wl), (g.R, g.S), padding=g.pad, stride=g.stride)       # [n, Cb, Hl, Wl]
        cd = torch.float64 if col.dtype == torch.float64 else torch.float32
        self._epilogue(y.permute(0, 2, 3, 1).to(cd), L_out, bias, 0, href, epi, act, slope)

    def adam_prepare(self, state, lr, b1, b2):
        self.launches += 1
        t = float(state[0]) + 1.0
        state[0] = t
        state[1] = lr / (1.0 - b1 ** t)
        state[2] = 1.0 / (1.0 - b2 ** t) ** 0.5

    def adam(self, param, grad, m, v, dims, gstrides, out1, s1, out2, s2, state, b1, b2, eps, wd):
        self.launches += 1
        p = param.detach()
        g = self._gather(grad, dims, gstrides).reshape(p.shape).to(p.dtype)
        if wd != 0:
            g = g + wd * p
        mm = m.view(-1)[:p.numel()].view(p.shape)
        vv = v.view(-1)[:p.numel()].view(p.shape)
        mm.mul_(b1).add_(g, alpha=1 - b1)
        vv.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = vv.sqrt() * float(state[2]) + eps
        p.addcdiv_(mm, denom, value=-float(state[1]))
        self.repack(p, dims, out1, s1, out2, s2)
