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

    # srgan_views emulation: a channel window (pitch, valid) of a wider buffer <-> the dense [pixels, C] operand
    @staticmethod
    def views_supported(g, dtype):
        return True

    @staticmethod
    def _window_read(t, pixels, C, pitch, valid):
        if not pitch and not valid:
            return t[:pixels * C].view(pixels, C)
        pitch, valid = pitch or C, valid or C
        w = torch.as_strided(t, (pixels, valid), (pitch, 1))
        return torch.cat([w, w.new_zeros(pixels, C - valid)], dim=1) if valid < C else w

    @staticmethod
    def _window_write(t, dense, pixels, C, pitch, valid):
        """dense: [pixels * C] result; only the window's channels are stored."""
        pitch, valid = pitch or C, valid or C
        torch.as_strided(t, (pixels, valid), (pitch, 1)).copy_(dense.view(pixels, C)[:, :valid].to(t.dtype))

    def conv_down(self, L, Wd, S_out, n, g, bias, bias_mod, href, epi, act, slope, views=None):
        self.launches += 1
        sp, sv, lp, lv = views or (0, 0, 0, 0)
        cd = torch.float64 if L.dtype == torch.float64 else torch.float32
        x = self._window_read(L, n * g.Hl * g.Wl, g.Cb, lp, lv).reshape(n, g.Hl, g.Wl, g.Cb).permute(0, 3, 1, 2).to(cd)
        y = F.conv2d(x, self._w_from_wd(Wd, g).to(cd), None, stride=g.stride, padding=g.pad)
        if sp or sv:
            dense = torch.empty(n * g.Hs * g.Ws * g.Ca, dtype=S_out.dtype)
            hr = None if href is None else self._window_read(href, n * g.Hs * g.Ws, g.Ca, sp, sv).reshape(-1)
            self._epilogue(y.permute(0, 2, 3, 1), dense, bias, bias_mod, hr, epi, act, slope)
            self._window_write(S_out, dense, n * g.Hs * g.Ws, g.Ca, sp, sv)
        else:
            self._epilogue(y.permute(0, 2, 3, 1), S_out, bias, bias_mod, href, epi, act, slope)

    def conv_up(self, S, Wu, L_out, n, g, bias, bias_mod, href, epi, act, slope, views=None):
        self.launches += 1
        sp, sv, lp, lv = views or (0, 0, 0, 0)
        cd = torch.float64 if S.dtype == torch.float64 else torch.float32
        x = self._window_read(S, n * g.Hs * g.Ws, g.Ca, sp, sv).reshape(n, g.Hs, g.Ws, g.Ca).permute(0, 3, 1, 2).to(cd)
        opad = g.Hl - ((g.Hs - 1) * g.stride - 2 * g.pad + g.R)
        y = F.conv_transpose2d(x, self._w_from_wu(Wu, g).to(cd), None, stride=g.stride, padding=g.pad,
                               output_padding=opad)
        if lp or lv:
            dense = torch.empty(n * g.Hl * g.Wl * g.Cb, dtype=L_out.dtype)
            hr = None if href is None else self._window_read(href, n * g.Hl * g.Wl, g.Cb, lp, lv).reshape(-1)
            self._epilogue(y.permute(0, 2, 3, 1), dense, bias, bias_mod, hr, epi, act, slope)
            self._window_write(L_out, dense, n * g.Hl * g.Wl, g.Cb, lp, lv)
        else:
            self._epilogue(y.permute(0, 2, 3, 1), L_out, bias, bias_mod, href, epi, act, slope)

    def conv_wgrad(self, S, L, dW, n, g, views=None):
        self.launches += 1
        sp, sv, lp, lv = views or (0, 0, 0, 0)
        cd = dW.dtype
        s = self._window_read(S, n * g.Hs * g.Ws, g.Ca, sp, sv).reshape(n, g.Hs, g.Ws, g.Ca).permute(0, 3, 1, 2).to(cd)
        l = self._window_read(L, n * g.Hl * g.Wl, g.Cb, lp, lv).reshape(n, g.Hl, g.Wl, g.Cb).permute(0, 3, 1, 2).to(cd)
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

    # ---- SGAN K-logit head (csrc/sgan.cu); logit-shaped tensors are [K, rows]
    @staticmethod
    def head_logits_workspace(rows, cols, K):
        return 1

    def head_logits(self, X, rows, cols, W, bias, K, out, ws=None):
        self.launches += 1
        x = X[:rows * cols].view(rows, cols).to(out.dtype)
        l = x @ W[:K * cols].view(K, cols).to(out.dtype).t()
        if bias is not None:
            l = l + bias.detach().reshape(-1)[:K].to(out.dtype).view(1, K)
        out.copy_(l.t())

    def sgan_loss(self, logitsT, K, n, mode, y, bins, target, scale, loss_out, dlogitsT):
        self.launches += 1
        l = logitsT.t()
        p = torch.softmax(l, dim=1)
        z = torch.logsumexp(l, dim=1)
        if mode == 0:
            idx = (y.to(l.dtype).reshape(-1, 1) - bins.to(l.dtype).reshape(1, -1)).abs().min(dim=1)[1]
            loss = (z - l.gather(1, idx.view(-1, 1)).squeeze(1)).sum()
            d = p - torch.nn.functional.one_hot(idx, K).to(l.dtype)
        else:
            loss = (torch.clamp(z, min=0) - z * target + torch.log1p(torch.exp(-z.abs()))).sum()
            d = (torch.sigmoid(z) - target).view(-1, 1) * p
        if loss_out is not None:
            loss_out += scale * loss
        if dlogitsT is not None:
            dlogitsT.copy_((scale * d).t())

    def sgan_gp_second(self, logitsT, tangentT, K, n, c, qT):
        self.launches += 1
        l, t = logitsT.t(), tangentT.t()
        p = torch.softmax(l, dim=1)
        sg = torch.sigmoid(torch.logsumexp(l, dim=1)).view(-1, 1)
        a = (p * t).sum(1, keepdim=True)
        qT.copy_((c * (sg * (1 - sg) * a * p + sg * (p * t - a * p))).t())

    def head_wgrad(self, X, rows, cols, dT, K, dW, db):
        self.launches += 1
        x = X[:rows * cols].view(rows, cols).to(dW.dtype)
        dW[:K * cols].view(K, cols).add_(dT.to(dW.dtype) @ x)
        if db is not None:
            db[:K].add_(dT.to(db.dtype).sum(1))

    def seed_rows_multi(self, out, rows, cols, dT, W, K, href, act, slope):
        self.launches += 1
        cd = href.dtype if href.dtype == torch.float64 else torch.float32
        v = dT.to(cd).t() @ W[:K * cols].view(K, cols).to(cd)
        v = v * dact(href[:rows * cols].view(rows, cols).to(cd), act, slope)
        out.copy_(v.reshape(-1).to(out.dtype))

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

    # ---- graph-net ops (crowd KnnDenseNetCat): eval-mode BatchNorm affine, slice copies, pools, crowd labeled loss
    @staticmethod
    def _cd(t):
        return torch.float64 if t.dtype == torch.float64 else torch.float32

    def affine(self, x, x_pitch, x_c0, y, y_pitch, rows, C, gamma, beta, mean, var, eps, href, mode, act, slope):
        """mode 0: y = act(gamma*(x-mean)/sqrt(var+eps)+beta); mode 1 (tangent): y = gamma/sqrt(var+eps)*x * act'(href).
        x is the channel slice [x_c0, x_c0+C) of rows with pitch x_pitch; y / href are columns [0, C) of rows with pitch
        y_pitch (columns >= C are left untouched)."""
        self.launches += 1
        cd = self._cd(x)
        xs = x.view(rows, x_pitch)[:, x_c0:x_c0 + C].to(cd)
        s = gamma.detach().to(cd) / torch.sqrt(var.detach().to(cd) + eps)
        if mode == 0:
            v = apply_act((xs - mean.detach().to(cd)) * s + beta.detach().to(cd), act, slope)
        else:
            v = xs * s * dact(href.view(rows, y_pitch)[:, :C].to(cd), act, slope)
        y.view(rows, y_pitch)[:, :C] = v.to(y.dtype)

    def affine_bwd(self, dy, dy_pitch, dx, dx_pitch, dx_c0, rows, C, gamma, var, eps, accumulate):
        """dx[:, c0:c0+C] (+)= dy[:, :C] * gamma/sqrt(var+eps)   (dy w.r.t. the affine's pre-activation, pitch dy_pitch)."""
        self.launches += 1
        cd = self._cd(dy)
        s = gamma.detach().to(cd) / torch.sqrt(var.detach().to(cd) + eps)
        v = dy.view(rows, dy_pitch)[:, :C].to(cd) * s
        d = dx.view(rows, dx_pitch)
        if accumulate:
            d[:, dx_c0:dx_c0 + C] += v.to(dx.dtype)
        else:
            d[:, dx_c0:dx_c0 + C] = v.to(dx.dtype)

    def affine_grad(self, dy, dy_pitch, x, x_pitch, x_c0, rows, C, mean, var, eps, dgamma, dbeta, subtract_mean):
        """dgamma[c] += sum_r dy[r,c] * (x[r,c0+c] - mean[c]*subtract_mean) / sqrt(var[c]+eps); dbeta[c] += sum_r dy[r,c]."""
        self.launches += 1
        cd = dgamma.dtype
        d = dy.view(rows, dy_pitch)[:, :C].to(cd)
        xs = x.view(rows, x_pitch)[:, x_c0:x_c0 + C].to(cd)
        if subtract_mean:
            xs = xs - mean.detach().to(cd)
        dgamma += (d * xs).sum(0) / torch.sqrt(var.detach().to(cd) + eps)
        if dbeta is not None:
            dbeta += d.sum(0)

    def copy2d(self, src, src_pitch, src_c0, dst, dst_pitch, dst_c0, rows, C, accumulate):
        self.launches += 1
        s = src.view(rows, src_pitch)[:, src_c0:src_c0 + C]
        d = dst.view(rows, dst_pitch)
        if accumulate:
            d[:, dst_c0:dst_c0 + C] += s.to(dst.dtype)
        else:
            d[:, dst_c0:dst_c0 + C] = s.to(dst.dtype)

    @staticmethod
    def _pool_out(H, k, s, p):
        return (H + 2 * p - k) // s + 1

    def _maxpool_idx(self, xref, n, H, W, C, k, s, p):
        """argmax (flat h*W+w index, first maximum in scan order like torch) of every window: [n, C, Ho, Wo]."""
        xr = xref.view(n, H, W, C).permute(0, 3, 1, 2).to(torch.float64)
        _, idx = F.max_pool2d(xr, k, s, p, return_indices=True)
        return idx

    @staticmethod
    def _window_origin(Ho, Wo, s, p):
        h0 = (torch.arange(Ho) * s - p).view(1, 1, Ho, 1)
        w0 = (torch.arange(Wo) * s - p).view(1, 1, 1, Wo)
        return h0, w0

    def _idx_encode(self, flat, W, k, s, p):
        """torch's flat argmax (h*W+w) [n,C,Ho,Wo] -> the CUDA op's byte map [n,Ho,Wo,C] of window positions dh*k+dw."""
        h0, w0 = self._window_origin(flat.shape[2], flat.shape[3], s, p)
        code = (flat // W - h0) * k + (flat % W - w0)
        return code.permute(0, 2, 3, 1).contiguous().to(torch.uint8)

    def _idx_decode(self, idx, n, Ho, Wo, C, W, k, s, p):
        code = idx.view(n, Ho, Wo, C).permute(0, 3, 1, 2).to(torch.int64)
        h0, w0 = self._window_origin(Ho, Wo, s, p)
        return (h0 + code // k) * W + (w0 + code % k)

    def maxpool(self, x, xref, y, y_pitch, y_c0, n, H, W, C, k, s, p, idx=None, idx_mode=0):
        """y[:, c0:c0+C] = x at the argmax of xref (xref None: of x itself) over each k x k window (stride s, pad p).
        idx (uint8 [n,Ho,Wo,C], window position dh*k+dw of the winner): idx_mode 1 = written here, 2 = routed by it."""
        self.launches += 1
        Ho, Wo = self._pool_out(H, k, s, p), self._pool_out(W, k, s, p)
        if idx_mode == 2:
            flat = self._idx_decode(idx, n, Ho, Wo, C, W, k, s, p)
        else:
            flat = self._maxpool_idx(x if xref is None else xref, n, H, W, C, k, s, p)
            if idx_mode == 1:
                idx.view(n, Ho, Wo, C).copy_(self._idx_encode(flat, W, k, s, p))
        xv = x.view(n, H * W, C).permute(0, 2, 1)                                   # [n, C, HW]
        out = torch.gather(xv, 2, flat.reshape(n, C, Ho * Wo))                      # [n, C, HoWo]
        y.view(n * Ho * Wo, y_pitch)[:, y_c0:y_c0 + C] = out.permute(0, 2, 1).reshape(n * Ho * Wo, C).to(y.dtype)

    def maxpool_bwd(self, xref, dy, dy_pitch, dy_c0, dx, n, H, W, C, k, s, p, act, slope, idx=None):
        """dx[i] = act'(xref[i]) * sum over the windows whose argmax is i of dy[window]   (dx dense [n,H,W,C], overwritten);
        with idx the argmax comes from the forward's index map."""
        self.launches += 1
        cd = self._cd(dy)
        Ho, Wo = self._pool_out(H, k, s, p), self._pool_out(W, k, s, p)
        flat = self._idx_decode(idx, n, Ho, Wo, C, W, k, s, p) if idx is not None else self._maxpool_idx(xref, n, H, W, C, k, s, p)
        d = dy.view(n * Ho * Wo, dy_pitch)[:, dy_c0:dy_c0 + C].to(cd).reshape(n, Ho * Wo, C).permute(0, 2, 1)
        g = torch.zeros(n, C, H * W, dtype=cd)
        g.scatter_add_(2, flat.reshape(n, C, Ho * Wo), d)
        g = g.permute(0, 2, 1).reshape(-1) * dact(xref.to(cd), act, slope)
        dx.copy_(g.to(dx.dtype))

    def avgpool(self, x, x_pitch, y, y_pitch, y_c0, n, H, W, C, k):
        """y[:, c0:c0+C] = mean over non-overlapping k x k windows of x[..., :C] (x: [n,H,W,x_pitch])."""
        self.launches += 1
        cd = self._cd(x)
        Ho, Wo = H // k, W // k
        v = x.view(n, Ho, k, Wo, k, x_pitch)[..., :C].to(cd).mean(dim=(2, 4))
        y.view(n * Ho * Wo, y_pitch)[:, y_c0:y_c0 + C] = v.reshape(n * Ho * Wo, C).to(y.dtype)

    def avgpool_bwd(self, dy, dy_pitch, dy_c0, dx, x_pitch, n, H, W, C, k, href, act, slope):
        """dx[n,h,w,c] = dy[n,h/k,w/k,c0+c] / k^2 * act'(href[n,h,w,c]) for c < C (dx, href: [n,H,W,x_pitch])."""
        self.launches += 1
        cd = self._cd(dy)
        Ho, Wo = H // k, W // k
        d = dy.view(n * Ho * Wo, dy_pitch)[:, dy_c0:dy_c0 + C].to(cd).view(n, Ho, 1, Wo, 1, C) / (k * k)
        g = d.expand(n, Ho, k, Wo, k, C).reshape(n * H * W, C)
        if href is not None and act != ACT_NONE:
            g = g * dact(href.view(n * H * W, x_pitch)[:, :C].to(cd), act, slope)
        dx.view(n * H * W, x_pitch)[:, :C] = g.to(dx.dtype)

    def crowd_loss(self, pred, density, maps, map_label, B, HW, order, scale, map_mult, loss_out, dpred, dm):
        """crowd/srgan.py:247-254 on rows [0,B): loss += scale * sum_b (|pred_b - sum(density_b)|^o + map_mult * m_b^o),
        m_b = sum_hw mean_c |map_c - label|; dpred_b = dLoss/dpred_b; dm_b = dLoss/dm_b."""
        self.launches += 1
        cd = pred.dtype
        target = density.detach().reshape(B, HW).to(cd).sum(1)
        lab = map_label.detach().reshape(B, HW).to(cd)
        m = sum((mp.view(B, HW).to(cd) - lab).abs() for mp in maps).sum(1) / len(maps)
        d = pred - target
        loss_out += scale * (d.abs().pow(order) + map_mult * m.pow(order)).sum()
        dpred.copy_(scale * order * d.abs().pow(order - 1) * torch.sign(d))
        dm.copy_(scale * map_mult * order * m.pow(order - 1))

    def crowd_map_grad(self, mp, map_label, dm, delta, B, HW, nmaps, act, slope):
        """delta[b,hw] += dm_b / nmaps * sign(map - label) * act'(map)   (delta w.r.t. the map layer's pre-activation)."""
        self.launches += 1
        cd = dm.dtype
        v = mp.view(B, HW).to(cd)
        g = dm.view(B, 1) / nmaps * torch.sign(v - map_label.detach().reshape(B, HW).to(cd)) * dact(v, act, slope)
        delta.view(B, HW).add_(g.to(delta.dtype))

    def depth_to_space(self, src, dst, n, Hs, Ws, k, inverse):
        """img[n, i*k+r, j*k+s] = blk[n, i, j, r*k+s]; inverse: blk from img."""
        self.launches += 1
        if not inverse:
            v = src.view(n, Hs, Ws, k, k).permute(0, 1, 3, 2, 4).reshape(-1)
        else:
            v = src.view(n, Hs, k, Ws, k).permute(0, 1, 3, 2, 4).reshape(-1)
        dst.copy_(v)

    def adam_multi(self, entries, grad, m, v, state, b1, b2, eps, wd):
        self.launches += 1
        for p, go, mo, n in entries:
            self.adam(p, grad[go:go + n], m[mo:mo + n], v[mo:mo + n], (n, 1, 1, 1), (1, 0, 0, 0), None, None, None, None,
                      state, b1, b2, eps, wd)
            self.launches -= 1

    def affine_bwd_grad(self, dy, dy_pitch, x, dx, x_pitch, x_c0, rows, C, gamma, mean, var, eps, dgamma, dbeta, accumulate):
        self.affine_grad(dy, dy_pitch, x, x_pitch, x_c0, rows, C, mean, var, eps, dgamma, dbeta, True)
        self.affine_bwd(dy, dy_pitch, dx, x_pitch, x_c0, rows, C, gamma, var, eps, accumulate)
        self.launches -= 1

    # ---- fused dense-layer kernels (csrc/bn_gemm.cu)
    @staticmethod
    def bn_fusion_supported(dtype):
        return True

    def bn_dgrad(self, dy, Wu, dx, x, rows, K, Cout, C, pitch, gamma, beta, mean, var, eps, dgamma, dbeta, d_out, d_pitch,
                 accumulate):
        """d = (dy . Wu^T)[:, :C] * [bn(x) > 0]; dx[:, :C] (+)= d * gamma/sigma; dgamma += sum d*(x-mean)/sigma; dbeta += sum d;
        d_out[:, :C] = d (optional).  x / dx: rows of `pitch` elements; dy dense [rows, K]; Wu [Cout, K]."""
        self.launches += 1
        cd = self._cd(dy)
        g, b = gamma.detach().to(cd), beta.detach().to(cd)
        mu, sig = mean.detach().to(cd), torch.sqrt(var.detach().to(cd) + eps)
        prod = dy.view(rows, K).to(cd) @ Wu.view(Cout, K)[:C].to(cd).t()
        prod = prod.to(dy.dtype).to(cd)                       # the accumulator is rounded to the activation dtype first
        self._bn_bwd_tail(prod, dx, x, rows, C, pitch, g, b, mu, sig, dgamma, dbeta, d_out, d_pitch, accumulate, cd)

    @staticmethod
    def _bn_bwd_tail(prod, dx, x, rows, C, pitch, g, b, mu, sig, dgamma, dbeta, d_out, d_pitch, accumulate, cd):
        xs = x.view(rows, pitch)[:, :C].to(cd)
        d = prod * (((xs - mu) * (g / sig) + b) > 0).to(cd)
        v = d * (g / sig)
        dxv = dx.view(rows, pitch)
        if accumulate:
            dxv[:, :C] = (dxv[:, :C].to(cd) + v).to(dx.dtype)
        else:
            dxv[:, :C] = v.to(dx.dtype)
        if dgamma is not None:
            dgamma += ((d * (xs - mu)).sum(0) / sig).to(dgamma.dtype)
            if dbeta is not None:
                dbeta += d.sum(0).to(dbeta.dtype)
        if d_out is not None:
            d_out.view(rows, d_pitch)[:, :C] = d.to(d_out.dtype)

    def bn_conv_down(self, x, Wd, out, rows, Kpad, Cout, C, pitch, gamma, beta, mean, var, eps, n1_out=None, n1_pitch=0, bn2=None,
                     out2=None, n1_first_row=0):
        """n1 = relu(bn(x[:, :C])) rounded to the activation dtype; out = n1 @ Wd[:, :C]^T; n1_out (optional) = n1, zeros in
        the padding columns [C, min(Kpad, n1_pitch)); out2 (optional) = relu(bn2(out))."""
        self.launches += 1
        cd = self._cd(x)
        s = gamma.detach().to(cd) / torch.sqrt(var.detach().to(cd) + eps)
        t = beta.detach().to(cd) - mean.detach().to(cd) * s
        n1 = torch.relu(x.view(rows, pitch)[:, :C].to(cd) * s + t).to(x.dtype)
        o = (n1.to(cd) @ Wd.view(Cout, Kpad)[:, :C].to(cd).t()).to(out.dtype)
        out.view(rows, Cout)[:] = o
        if n1_out is not None:
            v = n1_out.view(rows, n1_pitch)
            v[n1_first_row:, :C] = n1[n1_first_row:]
            v[n1_first_row:, C:min(Kpad, n1_pitch)] = 0
        if out2 is not None:
            g2, b2, m2, v2 = (q.detach().to(cd) for q in bn2)
            s2, C2 = g2 / torch.sqrt(v2 + eps), g2.numel()
            out2.view(rows, Cout)[:, :C2] = torch.relu(o[:, :C2].to(cd) * s2 + (b2 - m2 * s2)).to(out2.dtype)
            out2.view(rows, Cout)[:, C2:] = 0

    def bn_conv_wgrad(self, dy, x, dW, rows, Ca, Kpad, C, pitch, gamma, beta, mean, var, eps):
        """dW[Ca, Kpad][:, :C] += dy^T @ relu(bn(x[:, :C])) (the operand rounded to the activation dtype first)."""
        self.launches += 1
        cd = self._cd(dy)
        s = gamma.detach().to(cd) / torch.sqrt(var.detach().to(cd) + eps)
        t = beta.detach().to(cd) - mean.detach().to(cd) * s
        n1 = torch.relu(x.view(rows, pitch)[:, :C].to(cd) * s + t).to(x.dtype).to(cd)
        dW.view(Ca, Kpad)[:, :C] += (dy.view(rows, Ca).to(cd).t() @ n1).to(dW.dtype)

    def bn_conv_dgrad(self, dy, dy_pitch, dy_valid, Wu, dx, x, n, g, C, pitch, gamma, beta, mean, var, eps, dgamma, dbeta, d_out,
                      d_pitch, accumulate):
        """bn_dgrad with the product a stride-1 same-size transposed convolution (geometry g: dy is the small side, a channel
        window dy_pitch / dy_valid; x / dx the large side with `pitch` channels per pixel)."""
        cd = self._cd(dy)
        rows = n * g.Hl * g.Wl
        prod = torch.empty(rows * g.Cb, dtype=dy.dtype)
        self.launches -= 1                                    # one launch on the device
        self.conv_up(dy, Wu, prod, n, g, None, 0, None, 1, 0, 0.0, views=(dy_pitch, dy_valid, 0, 0) if (dy_pitch or dy_valid) else None)
        prod = prod.view(rows, g.Cb)[:, :C].to(cd)
        gm, b = gamma.detach().to(cd), beta.detach().to(cd)
        mu, sig = mean.detach().to(cd), torch.sqrt(var.detach().to(cd) + eps)
        self._bn_bwd_tail(prod, dx, x, rows, C, pitch, gm, b, mu, sig, dgamma, dbeta, d_out, d_pitch, accumulate, cd)
