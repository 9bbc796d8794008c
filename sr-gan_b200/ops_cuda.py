"""
ctypes binding of libsrgan_b200.so (include/srgan_b200.h): the ONLY ops implementation of the product.

There is no CPU or PyTorch fallback: if the library is missing, cannot be loaded, or a tensor is not on a CUDA
device, this raises.  torch is used only for device memory (data_ptr) and the current stream handle.
"""
from __future__ import annotations

import ctypes
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libsrgan_b200.so')

SRGAN_F32, SRGAN_BF16 = 0, 1


class Geom(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int) for k in ('Hs', 'Ws', 'Ca', 'Hl', 'Wl', 'Cb', 'R', 'S', 'stride', 'pad')]


class Views(ctypes.Structure):
    """srgan_views: channel windows of the small / large side operands (0 = dense)."""
    _fields_ = [(k, ctypes.c_int) for k in ('S_pitch', 'S_valid', 'L_pitch', 'L_valid')]


_EXPORTS = (
    'srgan_version', 'srgan_last_error', 'srgan_launch_count', 'srgan_last_path_tensor', 'srgan_set_force_simt',
    'srgan_tensor_launch_count', 'srgan_simt_fallback_count',
    'srgan_conv_down', 'srgan_conv_up', 'srgan_conv_wgrad', 'srgan_colsum', 'srgan_rowdot', 'srgan_seed_rows',
    'srgan_nchw_to_nhwc', 'srgan_nhwc_to_nchw', 'srgan_interpolate', 'srgan_labeled_loss', 'srgan_bce_logits',
    'srgan_distance', 'srgan_feature_norm_seed', 'srgan_gradnorm_penalty', 'srgan_gp_feature_seed', 'srgan_adam',
    'srgan_repack', 'srgan_im2col', 'srgan_col2im', 'srgan_adam_prepare', 'srgan_coefficient_step',
    'srgan_coefficient_step_workspace_bytes', 'srgan_affine', 'srgan_affine_bwd', 'srgan_affine_grad', 'srgan_copy2d',
    'srgan_maxpool', 'srgan_maxpool_bwd', 'srgan_avgpool', 'srgan_avgpool_bwd', 'srgan_crowd_loss', 'srgan_crowd_map_grad', 'srgan_depth_to_space', 'srgan_adam_multi', 'srgan_affine_bwd_grad',
    'srgan_adam_layout_multi', 'srgan_bn_dgrad', 'srgan_bn_conv_down', 'srgan_bn_conv_wgrad', 'srgan_bn_conv_dgrad',
    'srgan_crowd_extract_patches', 'srgan_sliding_window_merge', 'srgan_crowd_eval_sums', 'srgan_image_batch',
    'srgan_knn_maps', 'srgan_point_density_map', 'srgan_density_label', 'srgan_density_label_workspace_bytes',
    'srgan_head_logits', 'srgan_head_logits_workspace_bytes', 'srgan_sgan_loss', 'srgan_sgan_gp_second', 'srgan_seed_rows_multi', 'srgan_head_wgrad',
    'srgan_sliding_window_workspace_bytes', 'srgan_crowd_eval_workspace_bytes',
)

_lib = None


def load_library(path: str = LIB_PATH):
    """Loads the C-ABI library and declares the prototypes.  Raises if it is absent: build it with
    `python sr-gan_b200/build.py` (or __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError(f'{path} not found: the CUDA library is required (no CPU fallback); run '
                           '`python sr-gan_b200/build.py`')
    lib = ctypes.CDLL(path)
    for name in _EXPORTS:
        if not hasattr(lib, name):
            raise RuntimeError(f'{path} does not export {name}')
    c_int, c_ll, c_f, vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p
    gp = ctypes.POINTER(Geom)
    lib.srgan_version.restype = c_int
    lib.srgan_last_error.restype = ctypes.c_char_p
    lib.srgan_launch_count.restype = c_ll
    lib.srgan_last_path_tensor.restype = c_int
    lib.srgan_set_force_simt.argtypes = [c_int]
    lib.srgan_set_force_simt.restype = None
    vwp = ctypes.POINTER(Views)
    conv_args = [vp, vp, vp, c_int, gp, vp, c_int, vp, c_int, c_int, c_f, c_int, vwp, vp]
    lib.srgan_conv_down.argtypes = conv_args
    lib.srgan_conv_up.argtypes = conv_args
    lib.srgan_conv_wgrad.argtypes = [vp, vp, vp, c_int, gp, c_int, vwp, vp]
    lib.srgan_colsum.argtypes = [vp, c_ll, c_int, vp, c_int, vp, c_int, vp]
    lib.srgan_rowdot.argtypes = [vp, c_int, c_int, vp, vp, c_int, vp, c_int, vp]
    lib.srgan_seed_rows.argtypes = [vp, c_int, c_int, vp, vp, vp, vp, c_int, c_f, c_int, vp]
    lib.srgan_nchw_to_nhwc.argtypes = [vp, vp, c_int, c_int, c_int, c_int, c_int, vp]
    lib.srgan_nhwc_to_nchw.argtypes = [vp, vp, c_int, c_int, c_int, c_int, c_int, vp]
    lib.srgan_interpolate.argtypes = [vp, vp, vp, vp, c_int, c_ll, c_int, vp]
    lib.srgan_labeled_loss.argtypes = [vp, vp, c_int, c_int, c_f, vp, vp, vp]
    lib.srgan_bce_logits.argtypes = [vp, c_int, c_f, c_f, vp, vp, vp]
    lib.srgan_distance.argtypes = [vp, vp, c_int, c_f, c_int, c_f, vp, vp, vp, c_int, vp]
    lib.srgan_feature_norm_seed.argtypes = [vp, c_int, c_int, vp, vp, c_int, c_f, c_int, vp]
    lib.srgan_gradnorm_penalty.argtypes = [vp, c_int, c_ll, c_f, c_f, vp, vp, vp, vp, c_int, vp]
    lib.srgan_gp_feature_seed.argtypes = [vp, vp, vp, vp, c_int, c_int, c_int, c_f, c_int, vp]
    i4, l4 = ctypes.POINTER(c_int * 4), ctypes.POINTER(c_ll * 4)
    lib.srgan_adam.argtypes = [vp, vp, vp, vp, i4, l4, vp, l4, vp, l4, c_int, vp] + [c_f] * 4 + [vp]
    lib.srgan_adam_prepare.argtypes = [vp, ctypes.c_double, ctypes.c_double, ctypes.c_double, vp]
    lib.srgan_repack.argtypes = [vp, i4, vp, l4, vp, l4, c_int, vp]
    lib.srgan_im2col.argtypes = [vp, vp, c_int, gp, c_int, c_int, vp]
    lib.srgan_col2im.argtypes = [vp, vp, c_int, gp, c_int, vp, vp, c_int, c_int, c_f, c_int, vp]
    pp = ctypes.POINTER(vp)
    lib.srgan_coefficient_step.argtypes = ([pp, pp, pp, vp, vp, vp] + [vp] * 6 + [c_int, c_f, c_int, c_int] + [c_f] * 5 +
                                           [c_int, c_int] + [c_f] * 6 + [c_int, c_int, vp, ctypes.c_size_t, vp, vp, vp])
    lib.srgan_affine.argtypes = [vp, c_int, c_int, vp, c_int, c_ll, c_int, vp, vp, vp, vp, c_f, vp, c_int, c_int, c_f, c_int, vp]
    lib.srgan_affine_bwd.argtypes = [vp, c_int, vp, c_int, c_int, c_ll, c_int, vp, vp, c_f, c_int, c_int, vp]
    lib.srgan_affine_grad.argtypes = [vp, c_int, vp, c_int, c_int, c_ll, c_int, vp, vp, c_f, vp, vp, c_int, c_int, vp]
    lib.srgan_affine_bwd_grad.argtypes = [vp, c_int, vp, vp, c_int, c_int, c_ll, c_int, vp, vp, vp, c_f, vp, vp, c_int, c_int, vp]
    lib.srgan_copy2d.argtypes = [vp, c_int, c_int, vp, c_int, c_int, c_ll, c_int, c_int, c_int, vp]
    lib.srgan_maxpool.argtypes = [vp, vp, vp, c_int, c_int, vp, c_int] + [c_int] * 7 + [c_int, vp]
    lib.srgan_maxpool_bwd.argtypes = [vp, vp, vp, c_int, c_int, vp] + [c_int] * 7 + [c_int, c_f, c_int, vp]
    lib.srgan_avgpool.argtypes = [vp, c_int, vp, c_int, c_int] + [c_int] * 5 + [c_int, vp]
    lib.srgan_avgpool_bwd.argtypes = [vp, c_int, c_int, vp, c_int] + [c_int] * 5 + [vp, c_int, c_f, c_int, vp]
    lib.srgan_crowd_loss.argtypes = [vp, vp, pp, c_int, vp, c_int, c_ll, c_int, c_f, c_f, vp, vp, vp, c_int, vp]
    lib.srgan_adam_multi.argtypes = [vp, c_int, vp, vp, vp, vp, c_f, c_f, c_f, c_f, vp]
    lib.srgan_adam_layout_multi.argtypes = [vp, c_int, c_ll, vp, vp, vp, vp, c_f, c_f, c_f, c_f, c_int, vp]
    lib.srgan_depth_to_space.argtypes = [vp, vp, c_int, c_int, c_int, c_int, c_int, c_int, vp]
    lib.srgan_crowd_map_grad.argtypes = [vp, vp, vp, vp, c_int, c_ll, c_int, c_int, c_f, c_int, vp]
    lib.srgan_bn_dgrad.argtypes = [vp, vp, vp, vp, c_ll, c_int, c_int, c_int, c_int, vp, vp, vp, vp, c_f, vp, vp, vp, c_int, c_int, c_int, vp]
    lib.srgan_bn_conv_down.argtypes = [vp, vp, vp, c_ll, c_int, c_int, c_int, c_int, vp, vp, vp, vp, c_f, vp, c_int, c_ll, vp, vp, vp, vp, vp, c_int, c_int, vp]
    lib.srgan_bn_conv_dgrad.argtypes = [vp, c_int, c_int, vp, vp, vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                        vp, vp, vp, vp, c_f, vp, vp, vp, c_int, c_int, c_int, vp]
    lib.srgan_bn_conv_wgrad.argtypes = [vp, vp, vp, c_ll, c_int, c_int, c_int, c_int, vp, vp, vp, vp, c_f, c_int, vp]
    lib.srgan_crowd_extract_patches.argtypes = [vp, vp, vp, vp, vp, vp, c_int, vp, c_int, c_int, vp, vp, vp, vp]
    lib.srgan_sliding_window_merge.argtypes = [vp, vp, vp, c_int, vp, c_int, c_int, c_int, c_int, vp, vp, vp, ctypes.c_size_t, vp]
    lib.srgan_crowd_eval_sums.argtypes = [vp, vp, c_int, vp, c_int, c_ll, vp, vp, ctypes.c_size_t, vp]
    lib.srgan_crowd_eval_workspace_bytes.argtypes = [c_int]
    lib.srgan_image_batch.argtypes = [vp, c_int, vp, c_int, c_int, c_int, c_int, vp, vp, vp, vp]
    c_d = ctypes.c_double
    lib.srgan_knn_maps.argtypes = [vp, c_int, c_int, c_int, c_int, c_d, c_d, vp, vp, vp]
    lib.srgan_point_density_map.argtypes = [vp, c_int, c_int, c_int, vp, vp, vp]
    lib.srgan_density_label.argtypes = [vp, c_int, c_int, c_int, c_d, vp, vp, vp, ctypes.c_size_t, vp]
    lib.srgan_density_label_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.srgan_head_logits.argtypes = [vp, c_int, c_int, vp, vp, c_int, vp, vp, ctypes.c_size_t, c_int, vp]
    lib.srgan_head_logits_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.srgan_sgan_loss.argtypes = [vp, c_int, c_int, c_int, vp, vp, c_f, c_f, vp, vp, vp]
    lib.srgan_sgan_gp_second.argtypes = [vp, vp, c_int, c_int, c_f, vp, vp]
    lib.srgan_seed_rows_multi.argtypes = [vp, c_int, c_int, vp, vp, c_int, vp, c_int, c_f, c_int, vp]
    lib.srgan_head_wgrad.argtypes = [vp, c_int, c_int, vp, c_int, vp, vp, c_int, vp]
    lib.srgan_tensor_launch_count.restype = c_ll
    lib.srgan_simt_fallback_count.restype = c_ll
    for name in _EXPORTS[7:]:
        getattr(lib, name).restype = c_int
    lib.srgan_coefficient_step_workspace_bytes.restype = ctypes.c_size_t
    lib.srgan_sliding_window_workspace_bytes.restype = ctypes.c_size_t
    lib.srgan_crowd_eval_workspace_bytes.restype = ctypes.c_size_t
    lib.srgan_density_label_workspace_bytes.restype = ctypes.c_size_t
    lib.srgan_head_logits_workspace_bytes.restype = ctypes.c_size_t
    _lib = lib
    return lib


def _dt(t: torch.dtype) -> int:
    if t == torch.float32:
        return SRGAN_F32
    if t == torch.bfloat16:
        return SRGAN_BF16
    raise TypeError(f'unsupported activation dtype {t}')


class CudaOps:
    """The op set engine.py schedules, bound to the CUDA library.  Every tensor must live on the CUDA device."""

    def __init__(self, device='cuda:0'):
        if not torch.cuda.is_available():
            raise RuntimeError('CudaOps needs a CUDA device; there is no CPU fallback for the SR-GAN hot path')
        self.lib = load_library()
        self.device = torch.device(device)
        self._geoms = {}
        self._views_cache = {}
        self._stream_handle = None
        self._tables = {}                          # adam_multi device tables, one per NetState

    # -------------------------------------------------------------- helpers
    def _p(self, t, dtype=None):
        if t is None:
            return None
        if not t.is_cuda:
            raise RuntimeError('tensor is not on the CUDA device (no CPU fallback)')
        if dtype is not None and t.dtype != dtype:
            raise TypeError(f'expected {dtype}, got {t.dtype}')
        return t.data_ptr()

    def begin(self):
        """Called by the engine at the start of every step: caches the current stream handle (the torch lookup costs
        ~10 us, once per kernel launch it was 40 % of the host time of a step)."""
        self._stream_handle = torch.cuda.current_stream(self.device).cuda_stream

    def use_stream(self, stream):
        """Directs the following launches to `stream` (a torch.cuda.Stream); returns the previous handle for restore_stream."""
        prev, self._stream_handle = self._stream_handle, stream.cuda_stream
        return prev

    def restore_stream(self, handle):
        self._stream_handle = handle

    def _stream(self):
        h = self._stream_handle
        if h is None:
            h = torch.cuda.current_stream(self.device).cuda_stream
        return h

    def _geom(self, g):
        c = self._geoms.get(g)
        if c is None:
            c = Geom(g.Hs, g.Ws, g.Ca, g.Hl, g.Wl, g.Cb, g.R, g.S, g.stride, g.pad)
            self._geoms[g] = c
        return ctypes.byref(c)

    def _ck(self, rc, name):
        if rc != 0:
            raise RuntimeError(f'{name} failed ({rc}): {self.lib.srgan_last_error().decode()}')

    @property
    def launches(self):
        return self.lib.srgan_launch_count()

    @property
    def tensor_launches(self):
        """Contraction calls that ran on the tcgen05 kernels."""
        return self.lib.srgan_tensor_launch_count()

    @property
    def simt_fallbacks(self):
        """bf16 contraction calls of tensor-core size that were not tcgen05-eligible (also reported once on stderr)."""
        return self.lib.srgan_simt_fallback_count()

    # -------------------------------------------------------------- ops (signatures == tests/torch_ops.TorchOps)
    def repack(self, w, dims, out1, s1, out2, s2):
        f32 = torch.float32
        od = out1.dtype if out1 is not None else out2.dtype
        self._ck(self.lib.srgan_repack(self._p(w.detach(), f32), (ctypes.c_int * 4)(*dims), self._p(out1),
                                       (ctypes.c_longlong * 4)(*s1) if s1 else None, self._p(out2),
                                       (ctypes.c_longlong * 4)(*s2) if s2 else None, _dt(od), self._stream()),
                 'srgan_repack')

    def _views(self, views):
        """views = (S_pitch, S_valid, L_pitch, L_valid) or None -> cached ctypes struct pointer."""
        if views is None:
            return None
        c = self._views_cache.get(views)
        if c is None:
            c = self._views_cache[views] = Views(*views)
        return ctypes.byref(c)

    @staticmethod
    def views_supported(g, dtype):
        """Channel windows are a feature of the TMA-fed tcgen05 kernels (csrc/umma_conv.cu)."""
        return dtype == torch.bfloat16 and g.Ca % 64 == 0 and g.Cb % 64 == 0

    def conv_down(self, L, Wd, S_out, n, g, bias, bias_mod, href, epi, act, slope, views=None):
        self._ck(self.lib.srgan_conv_down(self._p(L), self._p(Wd, L.dtype), self._p(S_out, L.dtype), n, self._geom(g),
                                          self._p(bias.detach(), torch.float32) if bias is not None else None, bias_mod,
                                          self._p(href, L.dtype) if href is not None else None, epi, act, slope,
                                          _dt(L.dtype), self._views(views), self._stream()), 'srgan_conv_down')

    def conv_up(self, S, Wu, L_out, n, g, bias, bias_mod, href, epi, act, slope, views=None):
        self._ck(self.lib.srgan_conv_up(self._p(S), self._p(Wu, S.dtype), self._p(L_out, S.dtype), n, self._geom(g),
                                        self._p(bias.detach(), torch.float32) if bias is not None else None, bias_mod,
                                        self._p(href, S.dtype) if href is not None else None, epi, act, slope,
                                        _dt(S.dtype), self._views(views), self._stream()), 'srgan_conv_up')

    def conv_wgrad(self, S, L, dW, n, g, views=None):
        self._ck(self.lib.srgan_conv_wgrad(self._p(S), self._p(L, S.dtype), self._p(dW, torch.float32), n,
                                           self._geom(g), _dt(S.dtype), self._views(views), self._stream()), 'srgan_conv_wgrad')

    def colsum(self, X, rows, cols, out, mod, rowscale):
        self._ck(self.lib.srgan_colsum(self._p(X), rows, cols, self._p(out, torch.float32), mod,
                                       self._p(rowscale, torch.float32), _dt(X.dtype), self._stream()), 'srgan_colsum')

    def rowdot(self, X, rows, cols, w, bias, bias_index, out):
        self._ck(self.lib.srgan_rowdot(self._p(X), rows, cols, self._p(w, torch.float32),
                                       self._p(bias.detach(), torch.float32), bias_index, self._p(out, torch.float32),
                                       _dt(X.dtype), self._stream()), 'srgan_rowdot')

    def seed_rows(self, out, rows, cols, gvec, rowscale, wrow, href, act, slope):
        f32 = torch.float32
        self._ck(self.lib.srgan_seed_rows(self._p(out), rows, cols, self._p(gvec, f32), self._p(rowscale, f32),
                                          self._p(wrow, f32), self._p(href, out.dtype), act, slope, _dt(out.dtype),
                                          self._stream()), 'srgan_seed_rows')

    def nchw_to_nhwc(self, src, dst, n, c, h, w):
        src = src.detach()
        if src.dtype != torch.float32:
            src = src.float()
        self._ck(self.lib.srgan_nchw_to_nhwc(self._p(src), self._p(dst), n, c, h, w, _dt(dst.dtype), self._stream()),
                 'srgan_nchw_to_nhwc')

    def nhwc_to_nchw(self, src, dst, n, c, h, w):
        self._ck(self.lib.srgan_nhwc_to_nchw(self._p(src), self._p(dst, torch.float32), n, c, h, w, _dt(src.dtype),
                                             self._stream()), 'srgan_nhwc_to_nchw')

    def interpolate(self, u, fake, alpha, out, n, E):
        alpha = alpha.detach().reshape(-1)
        self._ck(self.lib.srgan_interpolate(self._p(u), self._p(fake, u.dtype), self._p(alpha, torch.float32),
                                            self._p(out, u.dtype), n, E, _dt(u.dtype), self._stream()),
                 'srgan_interpolate')

    def labeled_loss(self, pred, y, n, order, scale, loss_out, dpred):
        f32 = torch.float32
        y = y.detach()
        if y.dtype != f32:
            y = y.float()
        self._ck(self.lib.srgan_labeled_loss(self._p(pred, f32), self._p(y.contiguous()), n, int(order), scale,
                                             self._p(loss_out, f32), self._p(dpred, f32), self._stream()),
                 'srgan_labeled_loss')

    def bce_logits(self, scores, n, target, scale, loss_out, dscore):
        f32 = torch.float32
        self._ck(self.lib.srgan_bce_logits(self._p(scores, f32), n, target, scale, self._p(loss_out, f32),
                                           self._p(dscore, f32), self._stream()), 'srgan_bce_logits')

    def distance(self, sum_base, sum_other, F, inv_B, kind, mult, loss_out, gbase, gother, accumulate_base):
        f32 = torch.float32
        self._ck(self.lib.srgan_distance(self._p(sum_base, f32), self._p(sum_other, f32), F, inv_B, kind, mult,
                                         self._p(loss_out, f32), self._p(gbase, f32), self._p(gother, f32),
                                         int(bool(accumulate_base)), self._stream()), 'srgan_distance')

    def feature_norm_seed(self, h, rows, cols, s_out, gamma_out, act, slope):
        self._ck(self.lib.srgan_feature_norm_seed(self._p(h), rows, cols, self._p(s_out, torch.float32),
                                                  self._p(gamma_out, h.dtype), act, slope, _dt(h.dtype),
                                                  self._stream()), 'srgan_feature_norm_seed')

    def gradnorm_penalty(self, g0, n, E, lam_over_B, inv_B, gnorm_out, pen_out, gnmean_out, u0_out):
        f32 = torch.float32
        self._ck(self.lib.srgan_gradnorm_penalty(self._p(g0), n, E, lam_over_B, inv_B, self._p(gnorm_out, f32),
                                                 self._p(pen_out, f32), self._p(gnmean_out, f32),
                                                 self._p(u0_out, g0.dtype), _dt(g0.dtype), self._stream()),
                 'srgan_gradnorm_penalty')

    def gp_feature_seed(self, uL, hL, s, out, rows, cols, act, slope):
        self._ck(self.lib.srgan_gp_feature_seed(self._p(uL), self._p(hL, uL.dtype), self._p(s, torch.float32),
                                                self._p(out, uL.dtype), rows, cols, act, slope, _dt(uL.dtype),
                                                self._stream()), 'srgan_gp_feature_seed')

    def im2col(self, L, col, n, g, kpad):
        self._ck(self.lib.srgan_im2col(self._p(L), self._p(col, L.dtype), n, self._geom(g), kpad, _dt(L.dtype),
                                       self._stream()), 'srgan_im2col')

    def col2im(self, col, L_out, n, g, kpad, bias, href, epi, act, slope):
        self._ck(self.lib.srgan_col2im(self._p(col), self._p(L_out, col.dtype), n, self._geom(g), kpad,
                                       self._p(bias.detach(), torch.float32) if bias is not None else None,
                                       self._p(href, col.dtype) if href is not None else None, epi, act, slope,
                                       _dt(col.dtype), self._stream()), 'srgan_col2im')

    def adam_prepare(self, state, lr, b1, b2):
        self._ck(self.lib.srgan_adam_prepare(self._p(state, torch.float32), float(lr), float(b1), float(b2),
                                             self._stream()), 'srgan_adam_prepare')

    def adam(self, param, grad, m, v, dims, gstrides, out1, s1, out2, s2, state, b1, b2, eps, wd):
        f32 = torch.float32
        od = out1.dtype if out1 is not None else (out2.dtype if out2 is not None else f32)
        self._ck(self.lib.srgan_adam(self._p(param.detach(), f32), self._p(grad, f32), self._p(m, f32), self._p(v, f32),
                                     (ctypes.c_int * 4)(*dims), (ctypes.c_longlong * 4)(*gstrides), self._p(out1),
                                     (ctypes.c_longlong * 4)(*s1) if s1 else None, self._p(out2),
                                     (ctypes.c_longlong * 4)(*s2) if s2 else None, _dt(od), self._p(state, f32), b1, b2,
                                     eps, wd, self._stream()), 'srgan_adam')

    def coefficient_workspace_bytes(self):
        return int(self.lib.srgan_coefficient_step_workspace_bytes())

    def coefficient_step(self, d_ptrs, g_ptrs, dnn_ptrs, d_state, g_state, dnn_state, x, y, u, z, alpha, z2, B, inv_Bg,
                         dggan, order, labeled_mult, unl_mult, fake_mult, gen_mult, gp_lambda, kind_match, kind_contrast,
                         lr, lr_dnn, wd, b1, b2, eps, phases, train_g, workspace, scalars, publish=None):
        """d_ptrs / g_ptrs / dnn_ptrs: ctypes (c_void_p * 32) tables built by pointer_table()."""
        f32 = torch.float32
        self._ck(self.lib.srgan_coefficient_step(
            d_ptrs, g_ptrs, dnn_ptrs, self._p(d_state, f32), self._p(g_state, f32), self._p(dnn_state, f32),
            self._p(x, f32), self._p(y, f32), self._p(u, f32), self._p(z, f32), self._p(alpha, f32), self._p(z2, f32),
            int(B), inv_Bg, int(dggan), int(order), labeled_mult, unl_mult, fake_mult, gen_mult, gp_lambda,
            int(kind_match), int(kind_contrast), lr, lr_dnn, wd, b1, b2, eps, int(phases), int(train_g),
            self._p(workspace), workspace.numel() * workspace.element_size(), self._p(scalars, f32),
            self._p(publish, f32) if publish is not None else None, self._stream()),
            'srgan_coefficient_step')

    def pointer_table(self, tensors):
        for t in tensors:
            if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise TypeError('pointer_table: fp32 contiguous CUDA tensors only')
        return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])

    # -------------------------------------------------------------- graph-net ops (crowd KnnDenseNetCat)
    def _pf(self, t):
        return self._p(t.detach() if t is not None else None, torch.float32)

    def affine(self, x, x_pitch, x_c0, y, y_pitch, rows, C, gamma, beta, mean, var, eps, href, mode, act, slope):
        self._ck(self.lib.srgan_affine(self._p(x), x_pitch, x_c0, self._p(y, x.dtype), y_pitch, rows, C, self._pf(gamma), self._pf(beta),
                                       self._pf(mean), self._pf(var), eps, self._p(href, x.dtype) if href is not None else None,
                                       mode, act, slope, _dt(x.dtype), self._stream()), 'srgan_affine')

    def affine_bwd(self, dy, dy_pitch, dx, dx_pitch, dx_c0, rows, C, gamma, var, eps, accumulate):
        self._ck(self.lib.srgan_affine_bwd(self._p(dy), dy_pitch, self._p(dx, dy.dtype), dx_pitch, dx_c0, rows, C, self._pf(gamma),
                                           self._pf(var), eps, int(bool(accumulate)), _dt(dy.dtype), self._stream()),
                 'srgan_affine_bwd')

    def affine_grad(self, dy, dy_pitch, x, x_pitch, x_c0, rows, C, mean, var, eps, dgamma, dbeta, subtract_mean):
        self._ck(self.lib.srgan_affine_grad(self._p(dy), dy_pitch, self._p(x, dy.dtype), x_pitch, x_c0, rows, C, self._pf(mean),
                                            self._pf(var), eps, self._pf(dgamma), self._pf(dbeta), int(bool(subtract_mean)),
                                            _dt(dy.dtype), self._stream()), 'srgan_affine_grad')

    def copy2d(self, src, src_pitch, src_c0, dst, dst_pitch, dst_c0, rows, C, accumulate):
        self._ck(self.lib.srgan_copy2d(self._p(src), src_pitch, src_c0, self._p(dst, src.dtype), dst_pitch, dst_c0, rows, C,
                                       int(bool(accumulate)), _dt(src.dtype), self._stream()), 'srgan_copy2d')

    def maxpool(self, x, xref, y, y_pitch, y_c0, n, H, W, C, k, s, p, idx=None, idx_mode=0):
        """idx (uint8 [n, Ho, Wo, C]): idx_mode 1 = written by this (forward) call, 2 = read (tangent routing)."""
        self._ck(self.lib.srgan_maxpool(self._p(x), self._p(xref, x.dtype) if xref is not None else None, self._p(y, x.dtype),
                                        y_pitch, y_c0, self._p(idx, torch.uint8) if idx is not None else None, idx_mode, n, H, W,
                                        C, k, s, p, _dt(x.dtype), self._stream()), 'srgan_maxpool')

    def maxpool_bwd(self, xref, dy, dy_pitch, dy_c0, dx, n, H, W, C, k, s, p, act, slope, idx=None):
        self._ck(self.lib.srgan_maxpool_bwd(self._p(xref), self._p(idx, torch.uint8) if idx is not None else None,
                                            self._p(dy, xref.dtype), dy_pitch, dy_c0, self._p(dx, xref.dtype), n, H, W, C, k, s, p,
                                            act, slope, _dt(xref.dtype), self._stream()), 'srgan_maxpool_bwd')

    def avgpool(self, x, x_pitch, y, y_pitch, y_c0, n, H, W, C, k):
        self._ck(self.lib.srgan_avgpool(self._p(x), x_pitch, self._p(y, x.dtype), y_pitch, y_c0, n, H, W, C, k, _dt(x.dtype),
                                        self._stream()), 'srgan_avgpool')

    def avgpool_bwd(self, dy, dy_pitch, dy_c0, dx, x_pitch, n, H, W, C, k, href, act, slope):
        self._ck(self.lib.srgan_avgpool_bwd(self._p(dy), dy_pitch, dy_c0, self._p(dx, dy.dtype), x_pitch, n, H, W, C, k,
                                            self._p(href, dy.dtype) if href is not None else None, act, slope, _dt(dy.dtype),
                                            self._stream()), 'srgan_avgpool_bwd')

    # ---- SGAN K-logit head (csrc/sgan.cu); logit-shaped tensors are [K, rows] fp32
    def head_logits_workspace(self, rows, cols, K):
        """fp32 elements of the partial-sum workspace srgan_head_logits needs."""
        return self.lib.srgan_head_logits_workspace_bytes(rows, cols, K) // 4

    def head_logits(self, X, rows, cols, W, bias, K, out, ws):
        f32 = torch.float32
        self._ck(self.lib.srgan_head_logits(self._p(X), rows, cols, self._p(W, f32), self._p(bias.detach() if bias is not None else None, f32),
                                            K, self._p(out, f32), self._p(ws, f32), ws.numel() * 4, _dt(X.dtype), self._stream()),
                 'srgan_head_logits')

    def sgan_loss(self, logitsT, K, n, mode, y, bins, target, scale, loss_out, dlogitsT):
        f32 = torch.float32
        self._ck(self.lib.srgan_sgan_loss(self._p(logitsT, f32), K, n, mode, self._p(y.detach() if y is not None else None, f32),
                                          self._p(bins, f32), target, scale, self._p(loss_out, f32), self._p(dlogitsT, f32),
                                          self._stream()), 'srgan_sgan_loss')

    def sgan_gp_second(self, logitsT, tangentT, K, n, c, qT):
        f32 = torch.float32
        self._ck(self.lib.srgan_sgan_gp_second(self._p(logitsT, f32), self._p(tangentT, f32), K, n, c, self._p(qT, f32),
                                               self._stream()), 'srgan_sgan_gp_second')

    def head_wgrad(self, X, rows, cols, dT, K, dW, db):
        f32 = torch.float32
        self._ck(self.lib.srgan_head_wgrad(self._p(X), rows, cols, self._p(dT, f32), K, self._p(dW, f32), self._p(db, f32),
                                           _dt(X.dtype), self._stream()), 'srgan_head_wgrad')

    def seed_rows_multi(self, out, rows, cols, dT, W, K, href, act, slope):
        f32 = torch.float32
        self._ck(self.lib.srgan_seed_rows_multi(self._p(out), rows, cols, self._p(dT, f32), self._p(W, f32), K,
                                                self._p(href, out.dtype), act, slope, _dt(out.dtype), self._stream()),
                 'srgan_seed_rows_multi')

    def crowd_loss(self, pred, density, maps, map_label, B, HW, order, scale, map_mult, loss_out, dpred, dm):
        f32 = torch.float32
        density, map_label = density.detach(), map_label.detach()
        if density.dtype != f32 or map_label.dtype != f32 or not density.is_contiguous() or not map_label.is_contiguous():
            raise TypeError('crowd labels (density, map) must be contiguous fp32 tensors')
        tbl = (ctypes.c_void_p * len(maps))(*[self._p(m, maps[0].dtype) for m in maps])
        self._ck(self.lib.srgan_crowd_loss(self._p(pred, f32), self._p(density), tbl, len(maps), self._p(map_label), B, HW,
                                           int(order), scale, map_mult, self._p(loss_out, f32), self._p(dpred, f32),
                                           self._p(dm, f32), _dt(maps[0].dtype), self._stream()), 'srgan_crowd_loss')

    def crowd_map_grad(self, mp, map_label, dm, delta, B, HW, nmaps, act, slope):
        self._ck(self.lib.srgan_crowd_map_grad(self._p(mp), self._p(map_label.detach(), torch.float32), self._p(dm, torch.float32),
                                               self._p(delta, mp.dtype), B, HW, nmaps, act, slope, _dt(mp.dtype),
                                               self._stream()), 'srgan_crowd_map_grad')

    def depth_to_space(self, src, dst, n, Hs, Ws, k, inverse):
        self._ck(self.lib.srgan_depth_to_space(self._p(src), self._p(dst, src.dtype), n, Hs, Ws, k, int(bool(inverse)),
                                               _dt(src.dtype), self._stream()), 'srgan_depth_to_space')

    def adam_multi(self, entries, grad, m, v, state, b1, b2, eps, wd):
        """entries: list of (param, grad offset, moment offset, numel); the device table is built once per list."""
        key = id(entries)
        tbl = self._tables.get(key)
        if tbl is None:
            rows = []
            for p, go, mo, n in entries:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError('adam_multi: fp32 contiguous CUDA parameters only')
                rows.append([p.data_ptr(), go, mo, n])
            tbl = (torch.tensor(rows, dtype=torch.int64).to(self.device), entries)      # keep `entries` alive with its id
            self._tables[key] = tbl
        f32 = torch.float32
        self._ck(self.lib.srgan_adam_multi(self._p(tbl[0]), len(entries), self._p(grad, f32), self._p(m, f32), self._p(v, f32),
                                           self._p(state, f32), b1, b2, eps, wd, self._stream()), 'srgan_adam_multi')

    ADAM_LM_CHUNK = 2048

    def adam_layout_multi(self, entries, grad, m, v, state, b1, b2, eps, wd, out_dtype):
        """entries: list of (param, grad offset, moment offset, dims, grad strides, out1, s1, out2, s2) with out1 / out2 the
        kernel-layout copies (tensors or None); one launch for all of them.  The device table is built once per list."""
        key = id(entries)
        tbl = self._tables.get(key)
        if tbl is None:
            rows, blk = [], 0
            for p, go, mo, dims, gs, o1, s1, o2, s2 in entries:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError('adam_layout_multi: fp32 contiguous CUDA parameters only')
                outs = [o for o in (o1, o2) if o is not None]
                f32 = 1 if outs and all(o.dtype == torch.float32 for o in outs) else 0
                if any((o.dtype == torch.float32) != bool(f32) for o in outs) or any(o.dtype not in (torch.float32, out_dtype) for o in outs):
                    raise TypeError('adam_layout_multi: the layout copies of a tensor must share one dtype')
                n = p.numel()
                rows.append([p.data_ptr(), go, mo] + list(dims) + list(gs) + [o1.data_ptr() if o1 is not None else 0] +
                            list(s1 or (0, 0, 0, 0)) + [o2.data_ptr() if o2 is not None else 0] + list(s2 or (0, 0, 0, 0)) +
                            [f32, blk, n, 0, 0])
                blk += (n + self.ADAM_LM_CHUNK - 1) // self.ADAM_LM_CHUNK
            tbl = (torch.tensor(rows, dtype=torch.int64).to(self.device), blk, entries)
            self._tables[key] = tbl
        f32 = torch.float32
        self._ck(self.lib.srgan_adam_layout_multi(self._p(tbl[0]), len(entries), tbl[1], self._p(grad, f32), self._p(m, f32),
                                                  self._p(v, f32), self._p(state, f32), b1, b2, eps, wd, _dt(out_dtype),
                                                  self._stream()), 'srgan_adam_layout_multi')

    def affine_bwd_grad(self, dy, dy_pitch, x, dx, x_pitch, x_c0, rows, C, gamma, mean, var, eps, dgamma, dbeta, accumulate):
        self._ck(self.lib.srgan_affine_bwd_grad(self._p(dy), dy_pitch, self._p(x, dy.dtype), self._p(dx, dy.dtype), x_pitch, x_c0,
                                                rows, C, self._pf(gamma), self._pf(mean), self._pf(var), eps, self._pf(dgamma),
                                                self._pf(dbeta), int(bool(accumulate)), _dt(dy.dtype), self._stream()),
                 'srgan_affine_bwd_grad')

    # -------------------------------------------------------------- fused dense-layer kernels (csrc/bn_gemm.cu)
    @staticmethod
    def bn_fusion_supported(dtype):
        """The BatchNorm-fused dense-layer GEMMs exist on the bf16 tcgen05 path only."""
        return dtype == torch.bfloat16

    def bn_dgrad(self, dy, Wu, dx, x, rows, K, Cout, C, pitch, gamma, beta, mean, var, eps, dgamma, dbeta, d_out, d_pitch,
                 accumulate):
        self._ck(self.lib.srgan_bn_dgrad(self._p(dy), self._p(Wu, dy.dtype), self._p(dx, dy.dtype), self._p(x, dy.dtype), rows, K, Cout,
                                         C, pitch, self._pf(gamma), self._pf(beta), self._pf(mean), self._pf(var), eps,
                                         self._pf(dgamma), self._pf(dbeta), self._p(d_out, dy.dtype) if d_out is not None else None,
                                         d_pitch, int(bool(accumulate)), _dt(dy.dtype), self._stream()), 'srgan_bn_dgrad')

    def bn_conv_down(self, x, Wd, out, rows, Kpad, Cout, C, pitch, gamma, beta, mean, var, eps, n1_out=None, n1_pitch=0, bn2=None,
                     out2=None, n1_first_row=0):
        """out = relu(bn(x[:, :C])) @ Wd^T in one launch; n1_out: also store the normalised operand (rows >= n1_first_row);
        bn2 = (gamma2, beta2, mean2, var2) + out2: relu(bn2(out)) from the epilogue (include/srgan_b200.h)."""
        g2 = [self._pf(t) for t in bn2] if bn2 is not None else [None] * 4
        self._ck(self.lib.srgan_bn_conv_down(self._p(x), self._p(Wd, x.dtype), self._p(out, x.dtype), rows, Kpad, Cout, C, pitch,
                                             self._pf(gamma), self._pf(beta), self._pf(mean), self._pf(var), eps,
                                             self._p(n1_out, x.dtype) if n1_out is not None else None, n1_pitch, n1_first_row, *g2,
                                             self._p(out2, x.dtype) if out2 is not None else None,
                                             bn2[0].numel() if bn2 is not None else 0, _dt(x.dtype), self._stream()),
                 'srgan_bn_conv_down')

    def bn_conv_wgrad(self, dy, x, dW, rows, Ca, Kpad, C, pitch, gamma, beta, mean, var, eps):
        """dW[Ca, Kpad] += dy^T @ relu(bn(x[:, :C])) from the raw concat buffer (include/srgan_b200.h)."""
        self._ck(self.lib.srgan_bn_conv_wgrad(self._p(dy), self._p(x, dy.dtype), self._pf(dW), rows, Ca, Kpad, C, pitch,
                                              self._pf(gamma), self._pf(beta), self._pf(mean), self._pf(var), eps, _dt(dy.dtype),
                                              self._stream()), 'srgan_bn_conv_wgrad')

    def bn_conv_dgrad(self, dy, dy_pitch, dy_valid, Wu, dx, x, n, g, C, pitch, gamma, beta, mean, var, eps, dgamma, dbeta, d_out,
                      d_pitch, accumulate):
        """bn_dgrad with the product a stride-1 same-size transposed convolution of geometry g (conv2 of a dense layer)."""
        self._ck(self.lib.srgan_bn_conv_dgrad(self._p(dy), dy_pitch, dy_valid, self._p(Wu, dy.dtype), self._p(dx, dy.dtype),
                                              self._p(x, dy.dtype), n, g.Hl, g.Wl, g.R, g.S, g.pad, g.Ca, g.Cb, C, pitch,
                                              self._pf(gamma), self._pf(beta), self._pf(mean), self._pf(var), eps, self._pf(dgamma),
                                              self._pf(dbeta), self._p(d_out, dy.dtype) if d_out is not None else None, d_pitch,
                                              int(bool(accumulate)), _dt(dy.dtype), self._stream()), 'srgan_bn_conv_dgrad')
