/*
 * srgan_b200.h -- C ABI of the B200 (sm_100a) kernels behind the SR-GAN training step.
 *
 * This is the drop-in boundary.  The reference (golmschenk/sr-gan) is pure Python/PyTorch and
 * has no FFI of its own; the seam it offers is method override on `Experiment` (srgan.py:259-391).  Each entry point
 * below replaces the ATen/cuDNN/cuBLAS call(s) the reference issues at the cited lines; the Python host code in
 * sr-gan_b200/ (engine.py, mixin.py) binds them with ctypes (see INTEGRATION.md for the stub).
 *
 * Conventions
 *  - plain C types only; every pointer is a DEVICE pointer owned by the caller (torch tensors on the Python side);
 *  - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*); it never allocates, frees or
 *    synchronises; returns 0 on success, <0 on error (message: srgan_last_error(), thread-local);
 *  - `dtype` selects the activation / kernel-layout-weight element type: SRGAN_F32 (parity mode, SIMT FMA, fp32
 *    accumulate) or SRGAN_BF16 (tensor-core mode: tcgen05 implicit GEMM where the shape allows, fp32 accumulate);
 *    master parameters, gradients, Adam moments, losses and reductions are always fp32;
 *  - activations are NHWC rows: [sample][y][x][channel], contiguous;
 *  - a "conv pair" relates a SMALL side S[n,Hs,Ws,Ca] and a LARGE side L[n,Hl,Wl,Cb] through taps W[a][r][s][b]:
 *      down : S[o]  = sum_{t,b} L[stride*o-pad+t, b] * W[a,t,b]     (nn.Conv2d fwd / nn.ConvTranspose2d bwd-data)
 *      up   : L[i] += sum_a S[o,a] * W[a,t,b],  i = stride*o-pad+t  (nn.ConvTranspose2d fwd / nn.Conv2d bwd-data)
 *      wgrad: dW[a,t,b] += sum_{n,o} S[o,a] * L[stride*o-pad+t, b]
 *    nn.Linear is the pair with every spatial extent 1.  Kernel weight layouts: Wd[a][r][s][b], Wu[b][r][s][a].
 */
#ifndef SRGAN_B200_H
#define SRGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRGAN_F32 0
#define SRGAN_BF16 1

#define SRGAN_ACT_NONE 0
#define SRGAN_ACT_LEAKY 1
#define SRGAN_ACT_TANH 2

#define SRGAN_EPI_BIAS_ACT 0 /* out = act(acc + bias)                                  */
#define SRGAN_EPI_DACT 1     /* out = acc * act'(href)   (backward / tangent passes)   */

#define SRGAN_OK 0
#define SRGAN_ERR_ARG -1
#define SRGAN_ERR_CUDA -2
#define SRGAN_ERR_UNSUPPORTED -3

typedef struct srgan_geom {
    int Hs, Ws, Ca; /* small side */
    int Hl, Wl, Cb; /* large side */
    int R, S, stride, pad;
} srgan_geom;

/* Channel windows of the two activation operands of a contraction (all zero / NULL = dense tensors).  A DenseNet dense
 * layer appends its growth_rate new channels to the block's concat buffer (torch.cat at crowd/models.py:353); with a window
 * the 3x3 convolution writes them in place (pointer = first channel of the window, pitch = channels of the concat buffer,
 * valid = growth_rate) and its data / weight gradients read that window of the concat delta, instead of copying slices.
 * Pitches and windows are in elements, multiples of 8; bf16 tcgen05-eligible shapes only (else SRGAN_ERR_ARG). */
typedef struct srgan_views {
    int S_pitch, S_valid; /* small side: elements between pixels (0 = Ca), channels that exist in memory (0 = Ca) */
    int L_pitch, L_valid; /* large side (0 = Cb) */
} srgan_views;

/* library */
int srgan_version(void);
const char* srgan_last_error(void);
/* number of kernels this library has launched since load (the bench reports it as gpu_launches) */
long long srgan_launch_count(void);
/* 1 if the last conv call on this thread ran on the tcgen05 path, 0 if SIMT */
int srgan_last_path_tensor(void);
/* number of contraction calls that ran on the tcgen05 kernels / that were bf16 calls of tensor-core size (both channel
 * counts >= 64) but not eligible and went to the fp32-FMA kernels instead (each distinct such shape is also reported once
 * on stderr unless SRGAN_QUIET_FALLBACK=1: a 30x slower path must not be silent) */
long long srgan_tensor_launch_count(void);
long long srgan_simt_fallback_count(void);
/* force the SIMT path even for bf16 (debug / A-B checks): 0 = auto, 1 = SIMT only */
void srgan_set_force_simt(int on);

/* ---- dense contractions -------------------------------------------------------------------------------------
 * Replace F.conv2d / F.conv_transpose2d / F.linear forward and their autograd backward (cuDNN fprop/dgrad/wgrad,
 * cuBLAS addmm) issued from age/models.py:44-52,68-80, coefficient/models.py:22-28,43-50,65-72,
 * crowd/models.py:139-147 and from `.backward()` / `autograd.grad` at srgan.py:265,280,284,292,295,304,368. */
int srgan_conv_down(const void* L, const void* Wd, void* S_out, int n, const srgan_geom* g, const float* bias,
                    int bias_mod, const void* href, int epi, int act, float slope, int dtype, const srgan_views* views,
                    void* stream);
int srgan_conv_up(const void* S, const void* Wu, void* L_out, int n, const srgan_geom* g, const float* bias,
                  int bias_mod, const void* href, int epi, int act, float slope, int dtype, const srgan_views* views,
                  void* stream);
/* dW (fp32, Wd layout) += wgrad(S, L).  `views` (may be NULL): see srgan_views; href shares the output's window. */
int srgan_conv_wgrad(const void* S, const void* L, float* dW, int n, const srgan_geom* g, int dtype, const srgan_views* views,
                     void* stream);

/* Thin-layer lowering (image layers with <= 4 channels on the large side, age/models.py:48-51 layer4 of G and :61 layer1
 * of D): col[p][k] = L[n, oh*stride-pad+r, ow*stride-pad+s, b], k = (r*S+s)*Cb + b, zero for k >= R*S*Cb (Kpad columns)
 * and for the padding halo; p = (n, oh, ow) on the small side.  The contraction then runs as a [pixels x Kpad] GEMM on
 * the tensor cores (srgan_conv_down / _up / _wgrad with a 1x1 geometry). */
int srgan_im2col(const void* L, void* col, int n, const srgan_geom* g, int kpad, int dtype, void* stream);
/* Inverse gather with the conv epilogue fused: L_out[n,ih,iw,b] = epilogue(sum over the taps (r,s) with
 * ih = oh*stride-pad+r, iw = ow*stride-pad+s of col[(n,oh,ow)][(r*S+s)*Cb+b]); epilogue as in srgan_conv_up. */
int srgan_col2im(const void* col, void* L_out, int n, const srgan_geom* g, int kpad, const float* bias, const void* href,
                 int epi, int act, float slope, int dtype, void* stream);

/* ---- reductions ---------------------------------------------------------------------------------------------
 * out[c % mod] += sum_r rowscale[r] * X[r,c]  (mod = 0: no folding; rowscale may be NULL).  Replaces
 * `features.mean(0)` (srgan.py:442-443, the division by the GLOBAL batch happens in srgan_distance) and the bias /
 * head-weight gradient reductions of autograd. */
int srgan_colsum(const void* X, long long rows, int cols, float* out, int mod, const float* rowscale, int dtype,
                 void* stream);
/* out[r] = sum_c X[r,c] * w[c] + bias[bias_index]: the prediction head (age/models.py:75 layer5 as a full-extent
 * conv == a dot product per sample; coefficient/models.py:49 linear4). */
int srgan_rowdot(const void* X, int rows, int cols, const float* w, const float* bias, int bias_index, float* out,
                 int dtype, void* stream);

/* ---- element-wise / loss kernels ----------------------------------------------------------------------------*/
/* out[r,c] = (gvec[c] + rowscale[r]*wrow[c]) * act'(href[r,c]); gvec or (rowscale,wrow) may be NULL.
 * Seeds the backward pass with d(loss)/d(pre-activation of the feature layer). */
int srgan_seed_rows(void* out, int rows, int cols, const float* gvec, const float* rowscale, const float* wrow,
                    const void* href, int act, float slope, int dtype, void* stream);
/* fp32 NCHW -> activation-dtype NHWC (the `.to(gpu)` tensors of srgan.py:110-117 enter here) */
int srgan_nchw_to_nhwc(const float* src, void* dst, int n, int c, int h, int w, int dtype, void* stream);
/* activation-dtype NHWC -> fp32 NCHW (results handed back to reference-side code, e.g. Experiment.fake_features) */
int srgan_nhwc_to_nchw(const void* src, float* dst, int n, int c, int h, int w, int dtype, void* stream);
/* x_hat = alpha*u + (1-alpha)*fake, alpha per sample: srgan.py:362-366 */
int srgan_interpolate(const void* u, const void* fake, const float* alpha, void* out, int n, long long per_sample,
                      int dtype, void* stream);
/* loss += scale * sum |pred-y|^order ; dpred = scale*order*|d|^(order-1)*sign(d): srgan.py:414-417 */
int srgan_labeled_loss(const float* pred, const float* y, int n, int order, float scale, float* loss, float* dpred,
                       void* stream);
/* BCE-with-logits against a constant target: coefficient/dggan.py:39-40,48-49,62-63 */
int srgan_bce_logits(const float* scores, int n, float target, float scale, float* loss, float* dscore, void* stream);
/* feature_distance_loss on global feature sums: srgan.py:438-449 + utility.py:201-243.
 * d = (sum_base - sum_other)*inv_B ; loss += mult*fn(d) ; gbase (+)= mult*inv_B*dfn/dd ; gother = -that.
 * kind: 0 abs_mean 1 abs_mean_neg 2 abs_plus_one_sqrt_mean_neg 3 abs_plus_one_log_mean_neg 4 square_mean 5 norm_mean */
int srgan_distance(const float* sum_base, const float* sum_other, int F, float inv_B, int kind, float mult,
                   float* loss, float* gbase, float* gother, int accumulate_base, void* stream);
/* s[r] = ||h[r,:]||_2 ; gamma[r,c] = h[r,c]/s[r] * act'(h[r,c]): seed of the gradient-penalty g-chain,
 * interpolate_loss_calculation srgan.py:377-381 differentiated by hand */
int srgan_feature_norm_seed(const void* h, int rows, int cols, float* s_out, void* gamma_out, int act, float slope,
                            int dtype, void* stream);
/* r = ||g0[n,:]||_2 ; penalty += lam_over_B*max(r-1,0)^2 ; gnorm_mean += inv_B*r ;
 * u0 = 2*lam_over_B*max(r-1,0)/r * g0 : srgan.py:371-374 and the seed of its double-backward */
int srgan_gradnorm_penalty(const void* g0, int n, long long per_sample, float lam_over_B, float inv_B, float* gnorm,
                           float* penalty, float* gnorm_mean, void* u0, int dtype, void* stream);
/* out = ((u - g<g,u>)/s) * act'(h), g = h/s : Jacobian of f/||f|| applied to the tangent */
int srgan_gp_feature_seed(const void* uL, const void* hL, const float* s, void* out, int rows, int cols, int act,
                          float slope, int dtype, void* stream);

/* ---- optimizer ----------------------------------------------------------------------------------------------
 * torch.optim.Adam.step (srgan.py:136-138,266,297,305).  The step-dependent scalars live in DEVICE memory so that a
 * captured CUDA graph of the training step can be replayed: srgan_adam_prepare advances state[0] = t and writes
 * state[1] = lr / (1 - beta1^t), state[2] = 1 / sqrt(1 - beta2^t) (computed in double like torch does on the host);
 * srgan_adam then updates one tensor and rewrites its kernel-layout copies.
 * param/m/v: fp32, torch layout [d0,d1,d2,d3]; grad: fp32, element (i0..i3) at sum i_k*gstride[k];
 * out1/out2 (may be NULL): activation-dtype (or fp32 when out_dtype == SRGAN_F32) copies at sum i_k*ostride[k].
 * L2 (coupled) weight decay like torch. */
int srgan_adam_prepare(float* state3, double lr, double beta1, double beta2, void* stream);
int srgan_adam(float* param, const float* grad, float* m, float* v, const int* dims4, const long long* gstrides4,
               void* out1, const long long* o1strides4, void* out2, const long long* o2strides4, int out_dtype,
               const float* state3, float beta1, float beta2, float eps, float weight_decay, void* stream);
/* the same update for many small tensors that have no kernel-layout copies (biases, BatchNorm weight / bias) in one launch.
 * table: DEVICE array [n_tensors][4] of int64 = {param pointer, offset into grad, offset into m and v, element count};
 * one CTA per tensor.  Moments and gradients are slices of the flat buffers grad / m / v. */
int srgan_adam_multi(const long long* table, int n_tensors, const float* grad, float* m, float* v, const float* state3,
                     float beta1, float beta2, float eps, float weight_decay, void* stream);
/* The same update for every tensor that has kernel-layout copies (convolution / linear weights, prediction heads) in ONE
 * launch.  table: device array of n_tensors rows of 26 int64: [0] param ptr, [1] gradient offset (elements, into grad),
 * [2] moment offset (into m and v), [3..6] master dims, [7..10] gradient strides, [11] out1 ptr (or 0), [12..15] out1 strides,
 * [16] out2 ptr (or 0), [17..20] out2 strides, [21] 1 = this tensor's layout copies are fp32 (else out_dtype), [22] first block of
 * this tensor (blocks of 2048 elements, ascending), [23] number of elements; total_blocks = blocks of all tensors. */
int srgan_adam_layout_multi(const long long* table, int n_tensors, long long total_blocks, const float* grad, float* m, float* v,
                            const float* state3, float beta1, float beta2, float eps, float weight_decay, int out_dtype,
                            void* stream);
/* layout copies only (initial weights / after load_models, srgan.py:221-251) */
int srgan_repack(const float* param, const int* dims4, void* out1, const long long* o1strides4, void* out2,
                 const long long* o2strides4, int out_dtype, void* stream);

/* ---- graph discriminator (crowd KnnDenseNetCat, crowd/models.py:1049-1166) -------------------------------------------
 * A "slice" is the channel range [c0, c0+C) of NHWC rows with pitch `pitch` elements (the in-place concat buffers of the
 * dense blocks, crowd/models.py:347-353 torch.cat).  Dense operands have pitch C.
 *
 * Eval-mode BatchNorm2d (+ReLU): disable_batch_norm_updates srgan.py:538-542 applied at :261,276 makes every
 * _BatchNorm a per-channel affine of its running statistics; weight and bias stay trainable (crowd/models.py:339-343,
 * 367, 1077, 1092).  mode 0: y = act(gamma*(x-mean)/sqrt(var+eps)+beta); mode 1 (tangent pass of the gradient penalty):
 * y = gamma/sqrt(var+eps) * x * act'(href).
 * y, href, dy: columns [0, C) of rows with pitch y_pitch / dy_pitch >= C (operand buffers padded to the tensor-core tile
 * granularity; the pad columns are never touched). */
int srgan_affine(const void* x, int x_pitch, int x_c0, void* y, int y_pitch, long long rows, int C, const float* gamma,
                 const float* beta, const float* mean, const float* var, float eps, const void* href, int mode, int act,
                 float slope, int dtype, void* stream);
/* dx[:, c0:c0+C] (+)= dy[:, :C] * gamma/sqrt(var+eps)   (dy w.r.t. the affine's pre-activation) */
int srgan_affine_bwd(const void* dy, int dy_pitch, void* dx, int dx_pitch, int dx_c0, long long rows, int C,
                     const float* gamma, const float* var, float eps, int accumulate, int dtype, void* stream);
/* dgamma[c] += sum_r dy[r,c]*(x[r,c0+c] - mean[c]*subtract_mean)/sqrt(var[c]+eps); dbeta[c] += sum_r dy[r,c] (may be NULL) */
int srgan_affine_grad(const void* dy, int dy_pitch, const void* x, int x_pitch, int x_c0, long long rows, int C,
                      const float* mean, const float* var, float eps, float* dgamma, float* dbeta, int subtract_mean,
                      int dtype, void* stream);
/* srgan_affine_grad (subtract_mean = 1) and srgan_affine_bwd fused: one pass over dy; x and dx are the same slice of the
 * activation / delta buffers */
int srgan_affine_bwd_grad(const void* dy, int dy_pitch, const void* x, void* dx, int x_pitch, int x_c0, long long rows, int C,
                          const float* gamma, const float* mean, const float* var, float eps, float* dgamma, float* dbeta,
                          int accumulate, int dtype, void* stream);
/* ---- dense-layer GEMMs with the preceding BatchNorm + ReLU fused in (bf16 / tcgen05 only; SRGAN_ERR_UNSUPPORTED otherwise) ----
 * _DenseLayer norm1 -> relu1 -> conv1 (crowd/models.py:339-341) and _Transition norm -> relu -> conv (:367-369): the 1x1
 * convolution is a [rows x C] GEMM over the first C channels of the block's concat buffer x (pitch elements per row).
 *
 * srgan_bn_dgrad: backward of that pair in ONE launch, given dy [rows x K] (delta of the convolution output) and Wu [Cout][K]
 * (Cout = C rounded up to 64, rows >= C zero):
 *   d[r,c]  = (sum_k dy[r,k] * Wu[c,k]) * [ (x[r,c]-mean[c])*gamma[c]/sqrt(var[c]+eps) + beta[c] > 0 ]
 *   dx[r,c] (+)= d[r,c] * gamma[c]/sqrt(var[c]+eps)                              (same pitch as x)
 *   dgamma[c] += sum_r d[r,c]*(x[r,c]-mean[c])/sqrt(var[c]+eps) ; dbeta[c] += sum_r d[r,c]      (dgamma NULL: skipped)
 *   d_out[r*d_pitch + c] = d[r,c]                                                (d_out NULL: not stored)
 * It replaces srgan_conv_up (EPI_DACT against the stored ReLU output) + srgan_affine_bwd_grad: the ReLU mask is recomputed
 * from x, so the normalised activation is not read, and the C-wide intermediate delta is never written. */
int srgan_bn_dgrad(const void* dy, const void* Wu, void* dx, const void* x, long long rows, int K, int Cout, int C, int pitch,
                   const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* dgamma,
                   float* dbeta, void* d_out, int d_pitch, int accumulate, int dtype, void* stream);

/* srgan_bn_conv_dgrad: the same backward with the product a stride-1, same-size R x S transposed convolution instead of a
 * GEMM -- norm2 -> relu2 -> conv2 (3x3) of _DenseLayer (crowd/models.py:343-346): dy is an NHWC activation [n, H, W, .] (a
 * channel window: dy_pitch elements between pixels, dy_valid channels exist, Cin channels per tap counted by Wu), Wu
 * [Cout][R*S*Cin], x / dx [n*H*W rows x pitch].  Replaces srgan_conv_up (EPI_DACT) + srgan_affine_bwd_grad / srgan_affine_bwd. */
int srgan_bn_conv_dgrad(const void* dy, int dy_pitch, int dy_valid, const void* Wu, void* dx, const void* x, int n, int H, int W,
                        int R, int S, int pad, int Cin, int Cout, int C, int pitch, const float* gamma, const float* beta,
                        const float* mean, const float* var, float eps, float* dgamma, float* dbeta, void* d_out, int d_pitch,
                        int accumulate, int dtype, void* stream);

/* srgan_bn_conv_down: forward of the pair in ONE launch.  The GEMM reads the RAW concat buffer x; every operand tile is
 * normalised + rectified in shared memory between its TMA arrival and the tcgen05.mma that consumes it, so the activation
 * n1 = relu(bn(x)) need not exist in memory:
 *   n1[r,c]  = max(0, fma(x[r,c], s[c], t[c])) rounded to bf16,  s = gamma/sqrt(var+eps), t = beta - mean*s   (c < C)
 *   out[r,o] = sum_c n1[r,c] * Wd[o*Kpad + c]                      (o < Cout; Wd [Cout][Kpad], columns >= C zero)
 *   n1_out[r*n1_pitch + c] = n1[r,c] for c < min(Kpad, n1_pitch), zeros for c >= C, rows r >= n1_first_row only
 *                            (n1_out NULL: not stored; the gradient-penalty tangent pass reads it for the interpolate rows)
 *   out2[r,o] = max(0, fma(out[r,o], s2[o], t2[o])) for o < C2, 0 for C2 <= o < Cout   (gamma2 .. out2 NULL: skipped) -- the
 *                            BatchNorm + ReLU that FOLLOWS the convolution (norm2 -> relu2, crowd/models.py:343-344; C2 of its
 *                            channels exist), applied to the rounded out.
 * Replaces srgan_affine (norm1) + srgan_conv_down (+ srgan_affine (norm2)). */
int srgan_bn_conv_down(const void* x, const void* Wd, void* out, long long rows, int Kpad, int Cout, int C, int pitch,
                       const float* gamma, const float* beta, const float* mean, const float* var, float eps, void* n1_out,
                       int n1_pitch, long long n1_first_row, const float* gamma2, const float* beta2, const float* mean2,
                       const float* var2, void* out2, int C2, int dtype, void* stream);
/* srgan_bn_conv_wgrad: weight gradient of the pair from the RAW concat buffer (same transform-on-load):
 *   dW[a*Kpad + c] += sum_r dy[r*Ca + a] * n1[r,c]            (a < Ca, c < Kpad; fp32 reductions, dW is accumulated into)
 * Replaces srgan_conv_wgrad on the stored n1.  The rows of a gradient-penalty tangent block (whose operand is not
 * relu(bn(x))) still go through srgan_conv_wgrad. */
int srgan_bn_conv_wgrad(const void* dy, const void* x, float* dW, long long rows, int Ca, int Kpad, int C, int pitch,
                        const float* gamma, const float* beta, const float* mean, const float* var, float eps, int dtype,
                        void* stream);

/* dst[:, d0:d0+C] (+)= src[:, s0:s0+C] : torch.cat writes / their backward reads, MapModule taps */
int srgan_copy2d(const void* src, int src_pitch, int src_c0, void* dst, int dst_pitch, int dst_c0, long long rows, int C,
                 int accumulate, int dtype, void* stream);
/* nn.MaxPool2d(k, stride, pad) (crowd/models.py:1078): y slice = x at the argmax of xref's windows (xref NULL: x itself;
 * the tangent pass routes the tangent through the forward's argmax); first maximum in scan order like torch.
 * Index map (idx: one byte per pooled element [n, Ho, Wo, C], the winning window position dh*k+dw): idx_mode 0 = none,
 * 1 = the forward pass WRITES it, 2 = y = x at the recorded positions (tangent pass; xref is ignored). */
int srgan_maxpool(const void* x, const void* xref, void* y, int y_pitch, int y_c0, unsigned char* idx, int idx_mode, int n,
                  int H, int W, int C, int k, int stride, int pad, int dtype, void* stream);
/* dx[i] = act'(xref[i]) * sum of dy over the windows whose argmax is i (dx dense, overwritten); with idx (may be NULL) the
 * argmax is read from the forward's index map instead of being recomputed from xref */
int srgan_maxpool_bwd(const void* xref, const unsigned char* idx, const void* dy, int dy_pitch, int dy_c0, void* dx, int n,
                      int H, int W, int C, int k, int stride, int pad, int act, float slope, int dtype, void* stream);
/* nn.AvgPool2d(k, k) (crowd/models.py:370) and avg_pool2d(kernel 7) (:1151): non-overlapping k x k means */
int srgan_avgpool(const void* x, int x_pitch, void* y, int y_pitch, int y_c0, int n, int H, int W, int C, int k, int dtype,
                  void* stream);
int srgan_avgpool_bwd(const void* dy, int dy_pitch, int dy_c0, void* dx, int x_pitch, int n, int H, int W, int C, int k,
                      const void* href, int act, float slope, int dtype, void* stream);
/* MapModule.map_transposed_conv_layer (crowd/models.py:768-769,780): a ConvTranspose2d whose kernel equals its stride and
 * whose output has one channel is a [pixels x C] x [C x k*k] GEMM (srgan_conv_down with a 1x1 geometry) followed by this
 * rearrangement: img[n, i*k+r, j*k+s] = blk[n, i, j, r*k+s]  (inverse != 0: the other direction, for the backward pass). */
int srgan_depth_to_space(const void* src, void* dst, int n, int Hs, int Ws, int k, int inverse, int dtype, void* stream);
/* CrowdExperiment.labeled_loss_function, crowd/srgan.py:247-254, on B samples:
 * loss += scale * sum_b (|pred_b - sum(density_b)|^order + map_mult * m_b^order), m_b = sum_hw mean_c |map_c - map_label|;
 * dpred_b = dLoss/dpred_b, dm_b = dLoss/dm_b.  maps: HOST array of nmaps (<= 4) device pointers, [B, HW] each. */
int srgan_crowd_loss(const float* pred, const float* density, const void* const* maps, int nmaps, const float* map_label,
                     int B, long long HW, int order, float scale, float map_mult, float* loss, float* dpred, float* dm,
                     int dtype, void* stream);
/* delta[b,i] += dm_b/nmaps * sign(map - map_label) * act'(map): the map-loss gradient w.r.t. one map layer's pre-activation */
int srgan_crowd_map_grad(const void* map, const float* map_label, const float* dm, void* delta, int B, long long HW, int nmaps,
                         int act, float slope, int dtype, void* stream);

/* ---- coefficient application: the whole step in one persistent kernel ------------------------------------------
 * BASELINE configs[0] (run.py:46-54): coefficient/models.py:12-72 MLPs (50->10->10->10->{1,2}, generator 10->10->10->10->50,
 * leaky 0.01) are 2.4 k parameters; the reference spends ~850 ATen launches per step on them.  One cooperative launch
 * runs Experiment.dnn_training_step (srgan.py:259-271; phases bit 0) and/or Experiment.gan_training_step
 * (srgan.py:273-320 incl. the DG-GAN overrides coefficient/dggan.py:22-64; phases bit 1): forwards, losses, gradient
 * penalty with its double backward, generator step and the three torch.optim.Adam updates (moments, step counters and
 * master parameters updated in place; the flat gradient buffers must be zero on entry and are zero again on exit).
 * {d,g,dnn}_ptrs: 32 device pointers each = 4 layers x {W, b, gradW, gradb, mW, mb, vW, vb}, fp32, torch layouts.
 * {d,g,dnn}_state: the srgan_adam_prepare state (only state[0] = step counter is read and advanced).
 * x [B,50], y [B], u [B,50], z [B,10], alpha [B], z2 [B,10]; the *_mult arguments already include srgan_loss_multiplier /
 * dggan_loss_multiplier; inv_Bg = 1 / global batch.  workspace: srgan_coefficient_step_workspace_bytes() bytes, needs no
 * initialisation.  scalars: the 7 engine slots (dnn, labeled, unlabeled, fake, penalty, grad-norm mean, generator).
 * publish (may be NULL): [4][B][10] fp32 features of x | u | fake | x_hat as the D step saw them (the fake block is replaced
 * by the generator step's) followed by [B] gradient norms -- the tensors srgan.py:332-386 leaves in
 * Experiment.{labeled,unlabeled,fake,interpolates}_features and Experiment.gradient_norm.
 * Single-rank only: the feature sums are combined inside the kernel. */
size_t srgan_coefficient_step_workspace_bytes(void);
int srgan_coefficient_step(const float* const* d_ptrs, const float* const* g_ptrs, const float* const* dnn_ptrs,
                           float* d_state, float* g_state, float* dnn_state, const float* x, const float* y,
                           const float* u, const float* z, const float* alpha, const float* z2, int B, float inv_Bg,
                           int dggan, int order, float labeled_mult, float unl_mult, float fake_mult, float gen_mult,
                           float gp_lambda, int kind_match, int kind_contrast, float lr, float lr_dnn, float wd,
                           float beta1, float beta2, float eps, int phases, int train_g, void* workspace,
                           size_t workspace_bytes, float* scalars, float* publish, void* stream);

/* ---- crowd input pipeline and sliding-window inference on the device ------------------
 * The full images (uint8 HWC), density labels and kNN maps (fp32 HW) of a dataset split stay resident in device memory,
 * concatenated image after image: image i starts at pixel pixel_offset[i] (byte 3*pixel_offset[i] of `images`, element
 * pixel_offset[i] of `labels` / `maps`) and is heights[i] x widths[i].
 *
 * srgan_crowd_extract_patches: one launch = one batch of the reference's per-sample host transform chain
 *   ExtractPatchForPosition(patch, patch, allow_padded=True)  crowd/data.py:370-492 (window centred at (y, x); outside the
 *                                                             image: constant 0 for image, label and map, :426-452)
 *   -> RandomHorizontalFlip                                   crowd/data.py:92-112 (np.flip(axis=1) of image, label, map)
 *   -> NegativeOneToOneNormalizeImage                         crowd/data.py:115-128 ((u8 -> fp32) / 127.5 - 1)
 *   -> NumpyArraysToTorchTensors                              crowd/data.py:41-63 (HWC -> CHW, fp32)
 * as ShanghaiTechTransformedDataset.__getitem__ (crowd/shanghai_tech_data.py:73-104) and ImageSlidingWindowDataset.__getitem__
 * (crowd/data.py:541-557) compose it.  pos: [B][4] int32 = {image index, y, x, flip}; outputs img_out [B,3,patch,patch],
 * label_out / map_out [B,patch,patch] fp32 (NULL together with their source: the sliding-window dataset has neither).
 * Bit-exact with the reference's numpy arithmetic.  label_patch_size != image_patch_size (scipy.misc.imresize, removed
 * from SciPy) is not supported: BASELINE's crowd configuration uses 224 / 224. */
int srgan_crowd_extract_patches(const uint8_t* images, const float* labels, const float* maps, const long long* pixel_offset,
                                const int* heights, const int* widths, int n_images, const int* pos, int B, int patch,
                                float* img_out, float* label_out, float* map_out, void* stream);
/* Age / driving datasets: one batch of AgeDataset.__getitem__ (age/data.py:52-60: imageio HWC uint8 -> transpose -> fp32 ->
 * utility.to_normalized_range, utility.py:129-132) or SteeringAngleDataset.__getitem__ (driving/data.py:44-51: CHW uint8 .npy)
 * from the decoded dataset resident in device memory: out[b] = images[index[b]] as [C,H,W] fp32 in [-1, 1] (IEEE division by
 * 127.5, then - 1: bit-exact), labels_out[b] = labels[index[b]] (both may be NULL).  hwc != 0: stored samples are [H,W,C]. */
int srgan_image_batch(const uint8_t* images, int hwc, const long long* index, int B, int C, int H, int W, float* out,
                      const float* labels, float* labels_out, void* stream);
/* CrowdExperiment.predict_full_example, crowd/srgan.py:345-395: merges the per-patch predictions of a sliding window over
 * one H x W image.  Window (yi, xi) is patch yi*nx + xi and covers rows [ys[yi]-patch/2, ys[yi]+patch/2) (clipped to the
 * image), columns likewise.  full_label[Y,X] = mean over the covering patches of their predicted density at that pixel
 * (patch_labels [ny*nx, patch, patch] fp32; NULL = all zeros, what KnnDenseNetCat returns, crowd/models.py:1153);
 * *full_count = sum over pixels of the mean of count_p / patch^2 (double).  Per-pixel sums run in patch order in fp32 like
 * the reference's `+=`; pixels no window covers divide by 1 (:391).  workspace: srgan_sliding_window_workspace_bytes(). */
size_t srgan_sliding_window_workspace_bytes(void);
int srgan_sliding_window_merge(const float* patch_labels, const float* patch_counts, const int* ys, int ny, const int* xs,
                               int nx, int H, int W, int patch, float* full_label, double* full_count, void* workspace,
                               size_t workspace_bytes, void* stream);
/* CrowdExperiment.evaluation_epoch, crowd/srgan.py:149-191: the float64 reductions behind Validation ME / MAE / MSE and
 * kNN MAE / MSE over n samples: out[0*n+b] = sum(densities[b]) ([n, HW] fp32; NULL: skipped), out[1*n+b] = sum over the
 * nmaps predicted maps and HW of |pred_maps[b,c] - maps[b]|, out[2*n+b] = the same with squares ([n,nmaps,HW] vs [n,HW];
 * both NULL: skipped).  workspace: srgan_crowd_eval_workspace_bytes(n). */
size_t srgan_crowd_eval_workspace_bytes(int n);
int srgan_crowd_eval_sums(const float* densities, const float* pred_maps, int nmaps, const float* maps, int n, long long HW,
                          double* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- crowd label preprocessing on the device ------------------------------------------------
 * head_yx: [n_heads][2] float64 (y, x) annotations, the array crowd/shanghai_tech_data.py:132-137 hands to
 * DatabasePreprocessor.generate_labels_for_example (crowd/database_preprocessor.py:64-99).
 *
 * srgan_knn_maps = generate_knn_map (crowd/database_preprocessor.py:258-290) for k = 1..kmax (<= 8) in one sweep: a brute-force
 * float64 search instead of the scikit-learn ball tree.  knn[k-1][y][x] = mean of the min(k, n_heads) smallest Euclidean
 * distances from pixel (y, x) to a head, each clipped to upper_bound first when upper_bound > 0 (:282-283); distances are
 * sqrt(dy*dy + dx*dx) without FMA contraction and the mean adds them in ascending order like numpy's mean(axis=1), so the
 * maps carry the reference's bits.  iknn_f16 (may be NULL; so may knn): the [kmax][H][W] float16 maps the preprocessor saves,
 * 1 / (knn + epsilon) (:92-99; epsilon = 1). */
int srgan_knn_maps(const double* head_yx, int n_heads, int H, int W, int kmax, double upper_bound, double epsilon, double* knn,
                   void* iknn_f16, void* stream);
/* generate_density_label (crowd/database_preprocessor.py:113-225) as generate_labels_for_example calls it (:81-89: perspective
 * None, perspective_resizing, yx_order, no body, count-normalised; run.py:67 trains crowd on the beta = 0.3 maps): every head adds
 * a unit-sum square Gaussian with sigma = beta x mean distance to its min(11, n) nearest heads (itself included), half-width
 * int(2 sigma), clipped at the borders; label [H][W] fp32 = n_heads x label / sum(label); label_f16 (may be NULL) = the float16
 * copy the preprocessor saves.  Per pixel the heads are accumulated in annotation order in fp32 like the reference's
 * `label += person_label`; device exp() and the float64 sums differ from numpy's in the last bits, so the result matches the
 * reference to ~1e-6 relative, not bit for bit.  workspace: srgan_density_label_workspace_bytes(n_heads, H, W). */
size_t srgan_density_label_workspace_bytes(int n_heads, int H, int W);
int srgan_density_label(const double* head_yx, int n_heads, int H, int W, double beta, float* label, void* label_f16,
                        void* workspace, size_t workspace_bytes, void* stream);
/* generate_point_density_map (crowd/database_preprocessor.py:246-256): density[round(y)][round(x)] += 1 per head (round half to
 * even, negative indexes wrap once like Python's); *out_of_bounds = heads that fall outside.  density [H][W] fp32 and the
 * counter are cleared by the call. */
int srgan_point_density_map(const double* head_yx, int n_heads, int H, int W, float* density, int* out_of_bounds, void* stream);

/* ---- SGAN method: K-logit head (sgan.py:18-67, age/sgan.py, coefficient/sgan.py) ----------
 * Logit-shaped arrays are TRANSPOSED: [K][rows] fp32, K <= 16 (settings.py:67 number_of_bins = 10).
 * srgan_head_logits: logitsT[k][r] = X[r,:] . W[k,:] + bias[k]  (the K-output head: age/models.py:65,79 layer5 with
 *   number_of_outputs = K as a full-extent conv, coefficient/models.py:83,92 linear4); X is read once for all K; per-column-range
 *   partial sums go through `workspace` (srgan_head_logits_workspace_bytes) and are added in a fixed order: reproducible. */
size_t srgan_head_logits_workspace_bytes(int rows, int cols, int K);
int srgan_head_logits(const void* X, int rows, int cols, const float* W, const float* bias, int K, float* logitsT, void* workspace,
                      size_t workspace_bytes, int dtype, void* stream);
/* mode 0: nn.CrossEntropyLoss(logits, real_numbers_to_bin_indexes(y, bins)) (sgan.py:20-31, utility.py:141-144; mean over the
 *   batch folded into `scale`): loss += scale * sum_r (logsumexp(l_r) - l_r[bin_r]), dlogitsT = scale * (softmax - onehot);
 * mode 1: nn.BCEWithLogitsLoss(logsumexp(logits), target) (sgan.py:33-67): loss += scale * sum_r (softplus(z_r) - target z_r),
 *   dlogitsT = scale * (sigmoid(z_r) - target) * softmax(l_r).   loss / dlogitsT may be NULL. */
int srgan_sgan_loss(const float* logitsT, int K, int n, int mode, const float* y, const float* bins, float target, float scale,
                    float* loss, float* dlogitsT, void* stream);
/* Double backward of the SGAN gradient penalty (srgan.py:360-375 over sgan.py:51-58): qT = H tangentT per sample, H the Hessian
 * of c * softplus(logsumexp(l)) w.r.t. the K logits -- the ordinary-backward seed the penalty leaves on the x_hat rows. */
int srgan_sgan_gp_second(const float* logitsT, const float* tangentT, int K, int n, float c, float* qT, void* stream);
/* dW[k][c] += sum_r dT[k][r] * X[r,c], db[k] += sum_r dT[k][r] (db may be NULL): the head's weight / bias gradients for all K
 * outputs in one pass over the feature rows (autograd of the K-output layer5 / linear4); cols must be a multiple of 4. */
int srgan_head_wgrad(const void* X, int rows, int cols, const float* dT, int K, float* dW, float* db, int dtype, void* stream);
/* out[r,c] = (sum_k dT[k][r] * W[k][c]) * act'(href[r,c]): srgan_seed_rows for a K-output head */
int srgan_seed_rows_multi(void* out, int rows, int cols, const float* dT, const float* W, int K, const void* href, int act,
                          float slope, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SRGAN_B200_H */
