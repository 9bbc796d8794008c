// Bandwidth-bound kernels of the graph discriminator (crowd KnnDenseNetCat, crowd/models.py:1049-1166):
//   eval-mode BatchNorm affine (+ReLU) forward / tangent / backward / parameter gradients (srgan.py:538-542 freezes the
//   statistics, weight and bias stay trainable), channel-slice copies (the in-place concat of the dense blocks), max / average
//   pooling with their transposes, and the crowd labeled loss (crowd/srgan.py:247-254).
// All activations are NHWC rows; a "slice" is the channel range [c0, c0+C) of rows with pitch `pitch` elements.
// Vectorised (4 elements per thread) whenever C, the pitches and the offsets are multiples of 4.
#include "common.cuh"

namespace {

__device__ __forceinline__ float4 ld4f(const float* p) { return *reinterpret_cast<const float4*>(p); }
// gamma / sqrt(var + eps), branch-free (MUFU.RSQ + one Newton step, ~1 ulp): sqrtf / division compile to slow-path branches
// that serialise the per-thread parameter loads of the streaming kernels
__device__ __forceinline__ float bn_scale(float gamma, float var, float eps) {
    const float v = var + eps;
    float y = rsqrtf(v);
    y = y * (1.5f - 0.5f * v * y * y);
    return gamma * y;
}

inline int ew_grid(long long work_items) {
    long long b = (work_items + 255) / 256;
    const long long cap = (long long)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------------------------------------
// affine: mode 0  y = act(gamma*(x-mean)/sqrt(var+eps) + beta) ; mode 1 (tangent)  y = gamma/sqrt(var+eps) * x * act'(href)
// ------------------------------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) affine_kernel(const T* __restrict__ x, int x_pitch, int x_c0, T* __restrict__ y,
                                                     int y_pitch, long long rows, int C, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, const float* __restrict__ mean,
                                                     const float* __restrict__ var, float eps, const T* __restrict__ href,
                                                     int mode, int act, float slope) {
    constexpr int V = VEC ? 4 : 1;
    const int cv = C / V;
    const long long total = rows * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cv;
        const int c = (int)(i - r * cv) * V;
        const T* xp = x + r * x_pitch + x_c0 + c;
        T* yp = y + r * y_pitch + c;
        if (VEC) {
            const float4 xv = ld4(xp), g = ld4f(gamma + c), vr = ld4f(var + c);
            float4 s = make_float4(bn_scale(g.x, vr.x, eps), bn_scale(g.y, vr.y, eps), bn_scale(g.z, vr.z, eps), bn_scale(g.w, vr.w, eps));
            float4 o;
            if (mode == 0) {
                const float4 m = ld4f(mean + c), b = ld4f(beta + c);
                o.x = act_fwd(bn_apply(xv.x, s.x, bn_shift(b.x, m.x, s.x)), act, slope); o.y = act_fwd(bn_apply(xv.y, s.y, bn_shift(b.y, m.y, s.y)), act, slope);
                o.z = act_fwd(bn_apply(xv.z, s.z, bn_shift(b.z, m.z, s.z)), act, slope); o.w = act_fwd(bn_apply(xv.w, s.w, bn_shift(b.w, m.w, s.w)), act, slope);
            } else {
                const float4 h = ld4(href + r * y_pitch + c);
                o.x = xv.x * s.x * act_bwd(h.x, act, slope); o.y = xv.y * s.y * act_bwd(h.y, act, slope);
                o.z = xv.z * s.z * act_bwd(h.z, act, slope); o.w = xv.w * s.w * act_bwd(h.w, act, slope);
            }
            st4(yp, o);
        } else {
            const float s = bn_scale(gamma[c], var[c], eps);
            const float xv = to_f(*xp);
            float o;
            if (mode == 0) o = act_fwd(bn_apply(xv, s, bn_shift(beta[c], mean[c], s)), act, slope);
            else o = xv * s * act_bwd(to_f(href[r * y_pitch + c]), act, slope);
            *yp = from_f<T>(o);
        }
    }
}

// dx[:, c0:c0+C] (+)= dy * gamma/sqrt(var+eps)
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) affine_bwd_kernel(const T* __restrict__ dy, int dy_pitch, T* __restrict__ dx, int dx_pitch, int dx_c0,
                                                         long long rows, int C, const float* __restrict__ gamma,
                                                         const float* __restrict__ var, float eps, int accumulate) {
    constexpr int V = VEC ? 4 : 1;
    const int cv = C / V;
    const long long total = rows * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cv;
        const int c = (int)(i - r * cv) * V;
        T* xp = dx + r * dx_pitch + dx_c0 + c;
        if (VEC) {
            const float4 d = ld4(dy + r * dy_pitch + c), g = ld4f(gamma + c), vr = ld4f(var + c);
            float4 o = make_float4(d.x * bn_scale(g.x, vr.x, eps), d.y * bn_scale(g.y, vr.y, eps), d.z * bn_scale(g.z, vr.z, eps),
                                   d.w * bn_scale(g.w, vr.w, eps));
            if (accumulate) { const float4 p = ld4(xp); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
            st4(xp, o);
        } else {
            float o = to_f(dy[r * dy_pitch + c]) * bn_scale(gamma[c], var[c], eps);
            if (accumulate) o += to_f(*xp);
            *xp = from_f<T>(o);
        }
    }
}

// dgamma[c] += sum_r dy[r,c]*(x[r,c0+c]-mean[c]*sub)/sqrt(var[c]+eps) ; dbeta[c] += sum_r dy[r,c]
// block = 8 warps over a 32*V-column strip; blockIdx.y = row chunk; lanes own V consecutive columns
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) affine_grad_kernel(const T* __restrict__ dy, int dy_pitch, const T* __restrict__ x, int x_pitch, int x_c0,
                                                          long long rows, int C, const float* __restrict__ mean,
                                                          const float* __restrict__ var, float eps, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta, int subtract_mean, long long rows_per_block) {
    constexpr int V = VEC ? 4 : 1;
    __shared__ float sg[8][32 * V + 1], sb[8][32 * V + 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + lane) * V;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(rows, r0 + rows_per_block);
    float ag[V], ab[V], mu[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { ag[v] = 0.f; ab[v] = 0.f; mu[v] = (subtract_mean && c + v < C) ? mean[c + v] : 0.f; }
    if (c < C) {
        for (long long r = r0 + w; r < r1; r += 8) {
            if (VEC) {
                const float4 d = ld4(dy + r * dy_pitch + c), xv = ld4(x + r * x_pitch + x_c0 + c);
                ag[0] = fmaf(d.x, xv.x - mu[0], ag[0]); ab[0] += d.x;
                if (V > 1) {
                    ag[1 % V] = fmaf(d.y, xv.y - mu[1 % V], ag[1 % V]); ab[1 % V] += d.y;
                    ag[2 % V] = fmaf(d.z, xv.z - mu[2 % V], ag[2 % V]); ab[2 % V] += d.z;
                    ag[3 % V] = fmaf(d.w, xv.w - mu[3 % V], ag[3 % V]); ab[3 % V] += d.w;
                }
            } else {
                const float d = to_f(dy[r * dy_pitch + c]);
                ag[0] = fmaf(d, to_f(x[r * x_pitch + x_c0 + c]) - mu[0], ag[0]); ab[0] += d;
            }
        }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) { sg[w][lane * V + v] = ag[v]; sb[w][lane * V + v] = ab[v]; }
    __syncthreads();
    for (int j = threadIdx.x; j < 32 * V; j += 256) {
        const int cc = blockIdx.x * 32 * V + j;
        if (cc < C) {
            float g = 0.f, b = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) { g += sg[k][j]; b += sb[k][j]; }
            atomicAdd(dgamma + cc, g / sqrtf(var[cc] + eps));
            if (dbeta) atomicAdd(dbeta + cc, b);
        }
    }
}

// dst[:, d0:d0+C] (+)= src[:, s0:s0+C]
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) copy2d_kernel(const T* __restrict__ src, int src_pitch, int src_c0, T* __restrict__ dst,
                                                     int dst_pitch, int dst_c0, long long rows, int C, int accumulate) {
    constexpr int V = VEC ? 4 : 1;
    const int cv = C / V;
    const long long total = rows * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cv;
        const int c = (int)(i - r * cv) * V;
        const T* sp = src + r * src_pitch + src_c0 + c;
        T* dp = dst + r * dst_pitch + dst_c0 + c;
        if (VEC) {
            float4 v = ld4(sp);
            if (accumulate) { const float4 p = ld4(dp); v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w; }
            st4(dp, v);
        } else {
            float v = to_f(*sp);
            if (accumulate) v += to_f(*dp);
            *dp = from_f<T>(v);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// pooling.  Window argmax = first maximum in (h, w) scan order (torch's max_pool2d rule); NaN never occurs here.
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ int window_argmax(const T* __restrict__ xr, int H, int W, int C, int ho, int wo, int k, int s, int p) {
    // xr points at (sample, channel) of an NHWC tensor; returns h*W+w of the first maximum (-1: empty window)
    float best = -INFINITY;
    int arg = -1;
    const int h0 = ho * s - p, w0 = wo * s - p;
    for (int dh = 0; dh < k; ++dh) {
        const int h = h0 + dh;
        if (h < 0 || h >= H) continue;
        for (int dw = 0; dw < k; ++dw) {
            const int w = w0 + dw;
            if (w < 0 || w >= W) continue;
            const float v = to_f(xr[((long long)h * W + w) * C]);
            if (v > best || arg < 0) { best = v; arg = h * W + w; }
        }
    }
    return arg;
}

template <typename T>
__global__ void __launch_bounds__(256) maxpool_kernel(const T* __restrict__ x, const T* __restrict__ xref, T* __restrict__ y,
                                                      int y_pitch, int y_c0, int n, int H, int W, int C, int Ho, int Wo, int k,
                                                      int s, int p) {
    const long long total = (long long)n * Ho * Wo * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const long long b = t / Ho;
        const long long base = b * H * W * C + c;
        const int arg = window_argmax(xref + base, H, W, C, ho, wo, k, s, p);
        y[((b * Ho + ho) * Wo + wo) * y_pitch + y_c0 + c] = arg >= 0 ? x[base + (long long)arg * C] : from_f<T>(0.f);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const T* __restrict__ xref, const T* __restrict__ dy, int dy_pitch,
                                                          int dy_c0, T* __restrict__ dx, int n, int H, int W, int C, int Ho,
                                                          int Wo, int k, int s, int p, int act, float slope) {
    const long long total = (long long)n * H * W * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const long long b = t / H;
        const long long base = b * H * W * C + c;
        // windows (ho, wo) that contain (h, w): ho*s - p <= h <= ho*s - p + k - 1
        int ho_lo = (h + p - k + 1 + s - 1) / s, ho_hi = (h + p) / s;
        int wo_lo = (w + p - k + 1 + s - 1) / s, wo_hi = (w + p) / s;
        if (h + p - k + 1 < 0) ho_lo = 0;
        if (w + p - k + 1 < 0) wo_lo = 0;
        ho_hi = min(ho_hi, Ho - 1); wo_hi = min(wo_hi, Wo - 1);
        float acc = 0.f;
        for (int ho = ho_lo; ho <= ho_hi; ++ho)
            for (int wo = wo_lo; wo <= wo_hi; ++wo)
                if (window_argmax(xref + base, H, W, C, ho, wo, k, s, p) == h * W + w)
                    acc += to_f(dy[((b * Ho + ho) * Wo + wo) * dy_pitch + dy_c0 + c]);
        dx[i] = from_f<T>(acc * act_bwd(to_f(xref[i]), act, slope));
    }
}

// 4-channel-per-thread versions (C % 4 == 0 and 4-aligned pitches): the window scan is done once for four channels
template <typename T>
__device__ __forceinline__ void window_argmax4(const T* __restrict__ xr, int H, int W, int C, int ho, int wo, int k, int s, int p,
                                               int (&arg)[4]) {
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    arg[0] = arg[1] = arg[2] = arg[3] = -1;
    const int h0 = ho * s - p, w0 = wo * s - p;
    for (int dh = 0; dh < k; ++dh) {
        const int h = h0 + dh;
        if (h < 0 || h >= H) continue;
        for (int dw = 0; dw < k; ++dw) {
            const int w = w0 + dw;
            if (w < 0 || w >= W) continue;
            const float4 v = ld4(xr + ((long long)h * W + w) * C);
            const int pos = h * W + w;
            if (v.x > best[0] || arg[0] < 0) { best[0] = v.x; arg[0] = pos; }
            if (v.y > best[1] || arg[1] < 0) { best[1] = v.y; arg[1] = pos; }
            if (v.z > best[2] || arg[2] < 0) { best[2] = v.z; arg[2] = pos; }
            if (v.w > best[3] || arg[3] < 0) { best[3] = v.w; arg[3] = pos; }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) maxpool4_kernel(const T* __restrict__ x, const T* __restrict__ xref, T* __restrict__ y,
                                                       int y_pitch, int y_c0, int n, int H, int W, int C, int Ho, int Wo, int k,
                                                       int s, int p) {
    const int C4 = C / 4;
    const long long total = (long long)n * Ho * Wo * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        long long t = i / C4;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const long long b = t / Ho;
        const long long base = b * H * W * C + c;
        int arg[4];
        window_argmax4(xref + base, H, W, C, ho, wo, k, s, p, arg);
        float4 o;
        o.x = arg[0] >= 0 ? to_f(x[base + (long long)arg[0] * C]) : 0.f;
        o.y = arg[1] >= 0 ? to_f(x[base + (long long)arg[1] * C + 1]) : 0.f;
        o.z = arg[2] >= 0 ? to_f(x[base + (long long)arg[2] * C + 2]) : 0.f;
        o.w = arg[3] >= 0 ? to_f(x[base + (long long)arg[3] * C + 3]) : 0.f;
        st4(y + ((b * Ho + ho) * Wo + wo) * y_pitch + y_c0 + c, o);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) maxpool_bwd4_kernel(const T* __restrict__ xref, const T* __restrict__ dy, int dy_pitch,
                                                           int dy_c0, T* __restrict__ dx, int n, int H, int W, int C, int Ho,
                                                           int Wo, int k, int s, int p, int act, float slope) {
    const int C4 = C / 4;
    const long long total = (long long)n * H * W * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        long long t = i / C4;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const long long b = t / H;
        const long long base = b * H * W * C + c;
        int ho_lo = (h + p - k + 1 + s - 1) / s, ho_hi = (h + p) / s;
        int wo_lo = (w + p - k + 1 + s - 1) / s, wo_hi = (w + p) / s;
        if (h + p - k + 1 < 0) ho_lo = 0;
        if (w + p - k + 1 < 0) wo_lo = 0;
        ho_hi = min(ho_hi, Ho - 1); wo_hi = min(wo_hi, Wo - 1);
        const int me = h * W + w;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ho = ho_lo; ho <= ho_hi; ++ho)
            for (int wo = wo_lo; wo <= wo_hi; ++wo) {
                int arg[4];
                window_argmax4(xref + base, H, W, C, ho, wo, k, s, p, arg);
                const float4 d = ld4(dy + ((b * Ho + ho) * Wo + wo) * dy_pitch + dy_c0 + c);
                if (arg[0] == me) acc.x += d.x;
                if (arg[1] == me) acc.y += d.y;
                if (arg[2] == me) acc.z += d.z;
                if (arg[3] == me) acc.w += d.w;
            }
        const long long xi = base + (long long)me * C;
        const float4 hx = ld4(xref + xi);
        acc.x *= act_bwd(hx.x, act, slope); acc.y *= act_bwd(hx.y, act, slope);
        acc.z *= act_bwd(hx.z, act, slope); acc.w *= act_bwd(hx.w, act, slope);
        st4(dx + xi, acc);
    }
}

// ---- index-map variants: the forward pass records, per pooled element, which window position won (one byte, dh*k+dw); the
// tangent pass and the backward pass then read one byte per window instead of re-scanning k*k inputs of the forward tensor
template <typename T, int V>
__global__ void __launch_bounds__(256) maxpool_idx_kernel(const T* __restrict__ x, T* __restrict__ y, int y_pitch, int y_c0,
                                                          unsigned char* __restrict__ idx, int write_idx, int n, int H, int W, int C,
                                                          int Ho, int Wo, int k, int s, int p) {
    const int CV = C / V;
    const long long total = (long long)n * Ho * Wo * CV;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % CV) * V;
        long long t = i / CV;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const long long b = t / Ho;
        const long long base = b * H * W * C + c;
        const long long oi = ((b * Ho + ho) * Wo + wo);
        const int h0 = ho * s - p, w0 = wo * s - p;
        float out[V];
        unsigned char win[V];
        if (write_idx) {
            float best[V];
#pragma unroll
            for (int q = 0; q < V; ++q) { best[q] = -INFINITY; win[q] = 255; out[q] = 0.f; }
            for (int dh = 0; dh < k; ++dh) {
                const int h = h0 + dh;
                if (h < 0 || h >= H) continue;
                for (int dw = 0; dw < k; ++dw) {
                    const int w = w0 + dw;
                    if (w < 0 || w >= W) continue;
                    float v[V];
                    if (V == 4) { const float4 a = ld4(x + base + ((long long)h * W + w) * C); v[0] = a.x; v[1 % V] = a.y; v[2 % V] = a.z; v[3 % V] = a.w; }
                    else v[0] = to_f(x[base + ((long long)h * W + w) * C]);
#pragma unroll
                    for (int q = 0; q < V; ++q)
                        if (v[q] > best[q] || win[q] == 255) { best[q] = v[q]; win[q] = (unsigned char)(dh * k + dw); out[q] = v[q]; }
                }
            }
#pragma unroll
            for (int q = 0; q < V; ++q) idx[oi * C + c + q] = win[q];
        } else {
#pragma unroll
            for (int q = 0; q < V; ++q) {
                const int wq = idx[oi * C + c + q];
                out[q] = wq == 255 ? 0.f : to_f(x[base + q + ((long long)(h0 + wq / k) * W + (w0 + wq % k)) * C]);
            }
        }
        T* yp = y + oi * y_pitch + y_c0 + c;
        if (V == 4) st4(yp, make_float4(out[0], out[1 % V], out[2 % V], out[3 % V]));
        else *yp = from_f<T>(out[0]);
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) maxpool_bwd_idx_kernel(const T* __restrict__ xref, const unsigned char* __restrict__ idx,
                                                              const T* __restrict__ dy, int dy_pitch, int dy_c0, T* __restrict__ dx,
                                                              int n, int H, int W, int C, int Ho, int Wo, int k, int s, int p, int act,
                                                              float slope) {
    const int CV = C / V;
    const long long total = (long long)n * H * W * CV;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % CV) * V;
        long long t = i / CV;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const long long b = t / H;
        int ho_lo = (h + p - k + 1 + s - 1) / s, ho_hi = (h + p) / s;
        int wo_lo = (w + p - k + 1 + s - 1) / s, wo_hi = (w + p) / s;
        if (h + p - k + 1 < 0) ho_lo = 0;
        if (w + p - k + 1 < 0) wo_lo = 0;
        ho_hi = min(ho_hi, Ho - 1); wo_hi = min(wo_hi, Wo - 1);
        float acc[V];
#pragma unroll
        for (int q = 0; q < V; ++q) acc[q] = 0.f;
        for (int ho = ho_lo; ho <= ho_hi; ++ho)
            for (int wo = wo_lo; wo <= wo_hi; ++wo) {
                const int me = (h - (ho * s - p)) * k + (w - (wo * s - p));       // this input's position inside that window
                const long long oi = ((b * Ho + ho) * Wo + wo);
                float d[V];
                if (V == 4) { const float4 a = ld4(dy + oi * dy_pitch + dy_c0 + c); d[0] = a.x; d[1 % V] = a.y; d[2 % V] = a.z; d[3 % V] = a.w; }
                else d[0] = to_f(dy[oi * dy_pitch + dy_c0 + c]);
                if (V == 4) {
                    const uchar4 wv = *reinterpret_cast<const uchar4*>(idx + oi * C + c);
                    if (wv.x == me) acc[0] += d[0];
                    if (wv.y == me) acc[1 % V] += d[1 % V];
                    if (wv.z == me) acc[2 % V] += d[2 % V];
                    if (wv.w == me) acc[3 % V] += d[3 % V];
                } else if (idx[oi * C + c] == me) acc[0] += d[0];
            }
        const long long xi = ((b * H + h) * W + w) * C + c;
        if (V == 4) {
            const float4 hx = ld4(xref + xi);
            st4(dx + xi, make_float4(acc[0] * act_bwd(hx.x, act, slope), acc[1 % V] * act_bwd(hx.y, act, slope),
                                     acc[2 % V] * act_bwd(hx.z, act, slope), acc[3 % V] * act_bwd(hx.w, act, slope)));
        } else {
            dx[xi] = from_f<T>(acc[0] * act_bwd(to_f(xref[xi]), act, slope));
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) avgpool_kernel(const T* __restrict__ x, int x_pitch, T* __restrict__ y, int y_pitch, int y_c0,
                                                      int n, int H, int W, int C, int k) {
    const int Ho = H / k, Wo = W / k;
    const long long total = (long long)n * Ho * Wo * C;
    const float inv = 1.f / (float)(k * k);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const long long b = t / Ho;
        float acc = 0.f;
        for (int dh = 0; dh < k; ++dh)
            for (int dw = 0; dw < k; ++dw)
                acc += to_f(x[((b * H + ho * k + dh) * W + wo * k + dw) * x_pitch + c]);
        y[((b * Ho + ho) * Wo + wo) * y_pitch + y_c0 + c] = from_f<T>(acc * inv);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const T* __restrict__ dy, int dy_pitch, int dy_c0, T* __restrict__ dx,
                                                          int x_pitch, int n, int H, int W, int C, int k,
                                                          const T* __restrict__ href, int act, float slope) {
    const int Ho = H / k, Wo = W / k;
    const long long total = (long long)n * H * W * C;
    const float inv = 1.f / (float)(k * k);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const long long b = t / H;
        float v = to_f(dy[((b * Ho + h / k) * Wo + w / k) * dy_pitch + dy_c0 + c]) * inv;
        const long long xi = ((b * H + h) * W + w) * x_pitch + c;
        if (href) v *= act_bwd(to_f(href[xi]), act, slope);
        dx[xi] = from_f<T>(v);
    }
}

// ------------------------------------------------------------------------------------------------------------
// crowd labeled loss, crowd/srgan.py:247-254.  One CTA per sample.
// ------------------------------------------------------------------------------------------------------------
struct MapPtrs { const void* p[4]; };

template <typename T>
__global__ void __launch_bounds__(1024) crowd_loss_kernel(const float* __restrict__ pred, const float* __restrict__ density,
                                                          MapPtrs maps, int nmaps, const float* __restrict__ map_label,
                                                          long long HW, int order, float scale, float map_mult,
                                                          float* __restrict__ loss, float* __restrict__ dpred,
                                                          float* __restrict__ dm) {
    __shared__ float red[32];
    const int b = blockIdx.x;
    float cnt = 0.f, m = 0.f;
    for (long long i = threadIdx.x; i < HW; i += blockDim.x) {
        cnt += density[b * HW + i];
        const float lab = map_label[b * HW + i];
        float a = 0.f;
        for (int j = 0; j < nmaps; ++j) a += fabsf(to_f(static_cast<const T*>(maps.p[j])[b * HW + i]) - lab);
        m += a;
    }
    cnt = block_sum(cnt, red);
    m = block_sum(m, red) / (float)nmaps;
    if (threadIdx.x == 0) {
        const float d = pred[b] - cnt, ad = fabsf(d), sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        float pw, dpw, mp, dmp;
        if (order == 2) { pw = ad * ad; dpw = 2.f * ad; mp = m * m; dmp = 2.f * m; }
        else if (order == 1) { pw = ad; dpw = 1.f; mp = m; dmp = 1.f; }
        else { pw = powf(ad, (float)order); dpw = order * powf(ad, (float)(order - 1)); mp = powf(m, (float)order); dmp = order * powf(m, (float)(order - 1)); }
        atomicAdd(loss, scale * (pw + map_mult * mp));
        dpred[b] = scale * dpw * sg;
        dm[b] = scale * map_mult * dmp;
    }
}

// delta[b,i] += dm[b]/nmaps * sign(map - label) * act'(map)
template <typename T>
__global__ void __launch_bounds__(256) crowd_map_grad_kernel(const T* __restrict__ mp, const float* __restrict__ map_label,
                                                             const float* __restrict__ dm, T* __restrict__ delta, int B,
                                                             long long HW, float inv_nmaps, int act, float slope) {
    const long long total = (long long)B * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / HW);
        const float v = to_f(mp[i]), d = v - map_label[i];
        const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        delta[i] = from_f<T>(to_f(delta[i]) + dm[b] * inv_nmaps * sg * act_bwd(v, act, slope));
    }
}

// ---- 2-D mapped vector versions of the streaming kernels: a thread owns ONE group of W = 4 or 8 channels (its BatchNorm
// scale / shift are loaded with vector loads and computed once, branch-free) and walks down the rows of its block's row
// range; a warp covers 32 consecutive channel groups of a row (512 contiguous bytes in bf16 with W = 8: one 16-byte access
// per thread).  grid.x = channel-group tiles, grid.y = row ranges sized so that the grid is ~EW_WAVES waves of resident
// blocks.  When a row has fewer than 32 groups the spare lanes take further rows: thread t -> (group t & (cvp-1), row
// t >> lg).  Every row loop is batched: the loads of EW_U rows are issued back to back before the first use, so a thread
// keeps EW_U x 16 bytes (x the number of input streams) in flight instead of one dependent load -> use -> store chain
// per row (the first version reached 2.4 TB/s = 0.37 of the measured HBM peak; ncu: profiles/r1_ncu_crowd_streaming.txt).
constexpr int EW_U = 4;                       // rows in flight per thread
constexpr int EW_WAVES = 2;
struct Ew2d { int lg; long long rpb; };       // log2(cvp), cvp = min(32, next power of two >= C/W); rows per block
template <int W>
__device__ __forceinline__ void ew2d_map(const Ew2d e, long long rows, int& c, long long& r0, long long& r1, int& rstep) {
    const int t = threadIdx.x;
    c = (blockIdx.x * 32 + (t & ((1 << e.lg) - 1))) * W;
    rstep = 256 >> e.lg;
    r0 = (long long)blockIdx.y * e.rpb + (t >> e.lg);
    r1 = min(rows, (long long)(blockIdx.y + 1) * e.rpb);
}
// W consecutive elements kept packed while in flight (16 bytes for 8 x bf16 and 4 x fp32)
template <typename T, int W> struct Raw;
template <> struct Raw<bf16, 8> { uint4 d; };
template <> struct Raw<bf16, 4> { uint2 d; };
template <> struct Raw<float, 4> { float4 d; };
template <> struct Raw<float, 8> { float4 d[2]; };
__device__ __forceinline__ Raw<bf16, 8> ld_raw8(const bf16* p) { Raw<bf16, 8> r; r.d = *reinterpret_cast<const uint4*>(p); return r; }
__device__ __forceinline__ Raw<bf16, 4> ld_raw4(const bf16* p) { Raw<bf16, 4> r; r.d = *reinterpret_cast<const uint2*>(p); return r; }
__device__ __forceinline__ Raw<float, 4> ld_raw4(const float* p) { Raw<float, 4> r; r.d = *reinterpret_cast<const float4*>(p); return r; }
__device__ __forceinline__ Raw<float, 8> ld_raw8(const float* p) {
    Raw<float, 8> r; r.d[0] = *reinterpret_cast<const float4*>(p); r.d[1] = *reinterpret_cast<const float4*>(p + 4); return r;
}
template <typename T, int W> __device__ __forceinline__ Raw<T, W> ld_raw(const T* p) {
    if constexpr (W == 8) return ld_raw8(p); else return ld_raw4(p);
}
__device__ __forceinline__ void unpack(const Raw<bf16, 8>& r, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.d);
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(h[q]); v[2 * q] = f.x; v[2 * q + 1] = f.y; }
}
__device__ __forceinline__ void unpack(const Raw<bf16, 4>& r, float (&v)[4]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.d);
#pragma unroll
    for (int q = 0; q < 2; ++q) { const float2 f = __bfloat1622float2(h[q]); v[2 * q] = f.x; v[2 * q + 1] = f.y; }
}
__device__ __forceinline__ void unpack(const Raw<float, 4>& r, float (&v)[4]) { v[0] = r.d.x; v[1] = r.d.y; v[2] = r.d.z; v[3] = r.d.w; }
__device__ __forceinline__ void unpack(const Raw<float, 8>& r, float (&v)[8]) {
    v[0] = r.d[0].x; v[1] = r.d[0].y; v[2] = r.d[0].z; v[3] = r.d[0].w; v[4] = r.d[1].x; v[5] = r.d[1].y; v[6] = r.d[1].z; v[7] = r.d[1].w;
}
template <int W> __device__ __forceinline__ void stw(float* p, const float (&v)[W]) {
#pragma unroll
    for (int q = 0; q < W / 4; ++q) *reinterpret_cast<float4*>(p + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
template <int W> __device__ __forceinline__ void stw(bf16* p, const float (&v)[W]) {
    if constexpr (W == 8) {
        uint4 r;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
        for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
        *reinterpret_cast<uint4*>(p) = r;
    } else {
        st4(p, make_float4(v[0], v[1], v[2], v[3]));
    }
}
// W per-channel parameters starting at channel c (a multiple of W): float4 loads when the table is 16-byte aligned
template <int W> __device__ __forceinline__ void ldp(const float* __restrict__ p, int c, float (&v)[W]) {
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
        for (int q = 0; q < W / 4; ++q) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p + c) + q);
            v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < W; ++q) v[q] = __ldg(p + c + q);
    }
}

// Software-pipelined row loop of the streaming kernels: the loads of batch k+1 (U rows per thread) are issued before batch
// k is processed, so every thread keeps U..2U rows in flight all the time (two register batches, ping-pong).
template <typename Batch, int U, typename LoadF, typename ProcF, typename TailF>
__device__ __forceinline__ void ew_rows(long long r, const long long r1, const int rstep, LoadF load, ProcF proc, TailF tail) {
    const long long span = (long long)U * rstep;
    Batch a, b;
    bool have = r + span - rstep < r1;
    if (have) load(a, r);
    while (have) {
        long long rn = r + span;
        bool have_n = rn + span - rstep < r1;
        if (have_n) load(b, rn);
        proc(a, r);
        r = rn; have = have_n;
        if (!have) break;
        rn = r + span; have_n = rn + span - rstep < r1;
        if (have_n) load(a, rn);
        proc(b, r);
        r = rn; have = have_n;
    }
    for (; r < r1; r += rstep) tail(r);
}
template <typename T, int W, int U> struct Batch1 { Raw<T, W> a[U]; };
template <typename T, int W, int U> struct Batch2 { Raw<T, W> a[U], b[U]; };
template <typename T, int W, int U> struct Batch3 { Raw<T, W> a[U], b[U], c[U]; };

// MODE 0: y = act(gamma*(x-mean)/sqrt(var+eps)+beta); MODE 1 (tangent): y = gamma/sqrt(var+eps) * x * act'(href)
template <typename T, int W, int MODE>
__global__ void __launch_bounds__(256, 2) affine2d_kernel(const T* __restrict__ x, int x_pitch, int x_c0, T* __restrict__ y, int y_pitch,
                                                          long long rows, int C, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ mean,
                                                          const float* __restrict__ var, float eps, const T* __restrict__ href,
                                                          int act, float slope, Ew2d e) {
    int c, rstep; long long r0, r1;
    ew2d_map<W>(e, rows, c, r0, r1, rstep);
    if (c >= C) return;
    float s[W], m[W], b[W];
    ldp<W>(gamma, c, s);
    ldp<W>(var, c, m);
#pragma unroll
    for (int q = 0; q < W; ++q) s[q] = bn_scale(s[q], m[q], eps);
    if (MODE == 0) {
        ldp<W>(mean, c, m); ldp<W>(beta, c, b);
#pragma unroll
        for (int q = 0; q < W; ++q) b[q] = bn_shift(b[q], m[q], s[q]);
    }
    const T* xp = x + x_c0 + c;
    const T* hp = href + c;
    T* yp = y + c;
    auto one = [&](const Raw<T, W>& xr, const Raw<T, W>& hr, long long r) {
        float xv[W], o[W];
        unpack(xr, xv);
        if (MODE == 0) {
#pragma unroll
            for (int q = 0; q < W; ++q) o[q] = act_fwd(bn_apply(xv[q], s[q], b[q]), act, slope);
        } else {
            float h[W];
            unpack(hr, h);
#pragma unroll
            for (int q = 0; q < W; ++q) o[q] = xv[q] * s[q] * act_bwd(h[q], act, slope);
        }
        stw<W>(yp + r * y_pitch, o);
    };
    typedef Batch2<T, W, EW_U> B;
    ew_rows<B, EW_U>(r0, r1, rstep,
        [&](B& t, long long r) {
#pragma unroll
            for (int u = 0; u < EW_U; ++u) t.a[u] = ld_raw<T, W>(xp + (r + (long long)u * rstep) * x_pitch);
            if (MODE == 1) {
#pragma unroll
                for (int u = 0; u < EW_U; ++u) t.b[u] = ld_raw<T, W>(hp + (r + (long long)u * rstep) * y_pitch);
            }
        },
        [&](const B& t, long long r) {
#pragma unroll
            for (int u = 0; u < EW_U; ++u) one(t.a[u], MODE == 1 ? t.b[u] : t.a[u], r + (long long)u * rstep);
        },
        [&](long long r) {
            Raw<T, W> xr = ld_raw<T, W>(xp + r * x_pitch), hr = xr;
            if (MODE == 1) hr = ld_raw<T, W>(hp + r * y_pitch);
            one(xr, hr, r);
        });
}

// dx[:, c0:c0+C] (+)= dy * gamma/sqrt(var+eps)
template <typename T, int W, bool ACC>
__global__ void __launch_bounds__(256, 2) affine_bwd2d_kernel(const T* __restrict__ dy, int dy_pitch, T* __restrict__ dx, int dx_pitch,
                                                              int dx_c0, long long rows, int C, const float* __restrict__ gamma,
                                                              const float* __restrict__ var, float eps, Ew2d e) {
    int c, rstep; long long r0, r1;
    ew2d_map<W>(e, rows, c, r0, r1, rstep);
    if (c >= C) return;
    float s[W], vr[W];
    ldp<W>(gamma, c, s);
    ldp<W>(var, c, vr);
#pragma unroll
    for (int q = 0; q < W; ++q) s[q] = bn_scale(s[q], vr[q], eps);
    const T* dp = dy + c;
    T* xp = dx + dx_c0 + c;
    auto one = [&](const Raw<T, W>& dr, const Raw<T, W>& pr, long long r) {
        float d[W], o[W];
        unpack(dr, d);
#pragma unroll
        for (int q = 0; q < W; ++q) o[q] = d[q] * s[q];
        if (ACC) {
            float pv[W];
            unpack(pr, pv);
#pragma unroll
            for (int q = 0; q < W; ++q) o[q] += pv[q];
        }
        stw<W>(xp + r * dx_pitch, o);
    };
    typedef Batch2<T, W, EW_U> B;
    ew_rows<B, EW_U>(r0, r1, rstep,
        [&](B& t, long long r) {
#pragma unroll
            for (int u = 0; u < EW_U; ++u) t.a[u] = ld_raw<T, W>(dp + (r + (long long)u * rstep) * dy_pitch);
            if (ACC) {
#pragma unroll
                for (int u = 0; u < EW_U; ++u) t.b[u] = ld_raw<T, W>(xp + (r + (long long)u * rstep) * dx_pitch);
            }
        },
        [&](const B& t, long long r) {
#pragma unroll
            for (int u = 0; u < EW_U; ++u) one(t.a[u], ACC ? t.b[u] : t.a[u], r + (long long)u * rstep);
        },
        [&](long long r) {
            Raw<T, W> dr = ld_raw<T, W>(dp + r * dy_pitch), pr = dr;
            if (ACC) pr = ld_raw<T, W>(xp + r * dx_pitch);
            one(dr, pr, r);
        });
}

template <typename T, int W, bool ACC>
__global__ void __launch_bounds__(256, 2) copy2d2d_kernel(const T* __restrict__ src, int src_pitch, int src_c0, T* __restrict__ dst,
                                                          int dst_pitch, int dst_c0, long long rows, int C, Ew2d e) {
    int c, rstep; long long r0, r1;
    ew2d_map<W>(e, rows, c, r0, r1, rstep);
    if (c >= C) return;
    const T* sp = src + src_c0 + c;
    T* dp = dst + dst_c0 + c;
    auto one = [&](const Raw<T, W>& sr, const Raw<T, W>& pr, long long r) {
        if (ACC) {
            float v[W], pv[W];
            unpack(sr, v);
            unpack(pr, pv);
#pragma unroll
            for (int q = 0; q < W; ++q) v[q] += pv[q];
            stw<W>(dp + r * dst_pitch, v);
        } else {
            *reinterpret_cast<Raw<T, W>*>(dp + r * dst_pitch) = sr;
        }
    };
    typedef Batch2<T, W, EW_U> B;
    ew_rows<B, EW_U>(r0, r1, rstep,
        [&](B& t, long long r) {
#pragma unroll
            for (int u = 0; u < EW_U; ++u) t.a[u] = ld_raw<T, W>(sp + (r + (long long)u * rstep) * src_pitch);
            if (ACC) {
#pragma unroll
                for (int u = 0; u < EW_U; ++u) t.b[u] = ld_raw<T, W>(dp + (r + (long long)u * rstep) * dst_pitch);
            }
        },
        [&](const B& t, long long r) {
#pragma unroll
            for (int u = 0; u < EW_U; ++u) one(t.a[u], ACC ? t.b[u] : t.a[u], r + (long long)u * rstep);
        },
        [&](long long r) {
            Raw<T, W> sr = ld_raw<T, W>(sp + r * src_pitch), pr = sr;
            if (ACC) pr = ld_raw<T, W>(dp + r * dst_pitch);
            one(sr, pr, r);
        });
}

// BatchNorm parameter gradients, optionally fused with the data gradient (DX), from ONE pass over dy:
//   dgamma[c] += sum_r dy[r,c]*(x[r,c0+c]-mean[c]*sub)/sqrt(var[c]+eps) ; dbeta[c] += sum_r dy[r,c] ;
//   DX: dx[r,c0+c] (+)= dy[r,c]*gamma[c]/sqrt(var[c]+eps)     (x and dx are the same slice of the activation / delta buffers)
// Same thread mapping as the kernels above; the row lanes of a block are combined in shared memory, one atomicAdd per
// channel and block.
template <typename T, int W, bool DX>
__global__ void __launch_bounds__(256, 2) affine_grad2d_kernel(const T* __restrict__ dy, int dy_pitch, const T* __restrict__ x,
                                                               T* __restrict__ dx, int x_pitch, int x_c0, long long rows, int C,
                                                               const float* __restrict__ gamma, const float* __restrict__ mean,
                                                               const float* __restrict__ var, float eps, float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta, int subtract_mean, int accumulate, Ew2d e) {
    constexpr int U = 2;
    __shared__ float sg[256 * W], sb[256 * W];
    int c, rstep; long long r0, r1;
    ew2d_map<W>(e, rows, c, r0, r1, rstep);
    float ag[W], ab[W];
#pragma unroll
    for (int q = 0; q < W; ++q) { ag[q] = 0.f; ab[q] = 0.f; }
    if (c < C) {
        float s[W], mu[W];
#pragma unroll
        for (int q = 0; q < W; ++q) { s[q] = 0.f; mu[q] = 0.f; }
        if (subtract_mean) ldp<W>(mean, c, mu);
        if (DX) {
            float vr[W];
            ldp<W>(gamma, c, s);
            ldp<W>(var, c, vr);
#pragma unroll
            for (int q = 0; q < W; ++q) s[q] = bn_scale(s[q], vr[q], eps);
        }
        const T* dp = dy + c;
        const T* xp = x + x_c0 + c;
        T* gp = DX ? dx + x_c0 + c : nullptr;
        auto one = [&](const Raw<T, W>& dr, const Raw<T, W>& xr, const Raw<T, W>& pr, long long r) {
            float d[W], xv[W];
            unpack(dr, d);
            unpack(xr, xv);
#pragma unroll
            for (int q = 0; q < W; ++q) { ag[q] = fmaf(d[q], xv[q] - mu[q], ag[q]); ab[q] += d[q]; }
            if (DX) {
                float o[W];
#pragma unroll
                for (int q = 0; q < W; ++q) o[q] = d[q] * s[q];
                if (accumulate) {
                    float pv[W];
                    unpack(pr, pv);
#pragma unroll
                    for (int q = 0; q < W; ++q) o[q] += pv[q];
                }
                stw<W>(gp + r * x_pitch, o);
            }
        };
        typedef Batch3<T, W, U> B;
        ew_rows<B, U>(r0, r1, rstep,
            [&](B& t, long long r) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    t.a[u] = ld_raw<T, W>(dp + (r + (long long)u * rstep) * dy_pitch);
                    t.b[u] = ld_raw<T, W>(xp + (r + (long long)u * rstep) * x_pitch);
                }
                if (DX && accumulate) {
#pragma unroll
                    for (int u = 0; u < U; ++u) t.c[u] = ld_raw<T, W>(gp + (r + (long long)u * rstep) * x_pitch);
                }
            },
            [&](const B& t, long long r) {
#pragma unroll
                for (int u = 0; u < U; ++u) one(t.a[u], t.b[u], t.c[u], r + (long long)u * rstep);     // t.c is read only when accumulating
            },
            [&](long long r) {
                Raw<T, W> dr = ld_raw<T, W>(dp + r * dy_pitch), xr = ld_raw<T, W>(xp + r * x_pitch), pr = dr;
                if (DX && accumulate) pr = ld_raw<T, W>(gp + r * x_pitch);
                one(dr, xr, pr, r);
            });
    }
    // combine the row lanes: thread t holds columns (t & (cvp-1))*W.. of row lane t >> lg
    const int cvp = 1 << e.lg, ncol = cvp * W, nrl = 256 >> e.lg;
    {
        const int base = (threadIdx.x >> e.lg) * ncol + (threadIdx.x & (cvp - 1)) * W;
#pragma unroll
        for (int q = 0; q < W; ++q) { sg[base + q] = ag[q]; sb[base + q] = ab[q]; }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < ncol; j += 256) {
        const int cc = blockIdx.x * 32 * W + j;
        if (cc < C) {
            float g = 0.f, b = 0.f;
            for (int k = 0; k < nrl; ++k) { g += sg[k * ncol + j]; b += sb[k * ncol + j]; }
            atomicAdd(dgamma + cc, g / sqrtf(var[cc] + eps));
            if (dbeta) atomicAdd(dbeta + cc, b);
        }
    }
}

inline Ew2d ew2d_plan(long long rows, int C, int W, dim3& grid, int unroll = EW_U, int blocks_per_sm = 2) {
    int lg = 0;
    while (lg < 5 && (1 << lg) < C / W) ++lg;
    const int rstep = 256 >> lg;
    const int gx = (C / W + 31) / 32;
    const long long unit = (long long)rstep * unroll;              // one batch of rows per row lane
    long long want = (long long)kNumSMs * blocks_per_sm * EW_WAVES / gx;
    if (want < 1) want = 1;
    long long rpb = (rows + want - 1) / want;
    // whole batches per row lane when the launch is big enough to fill the GPU that way; small launches (narrow concat
    // slices, the 7x7 stage) keep all their parallelism instead: down to one row per thread
    if (rpb >= unit) rpb = (rpb + unit - 1) / unit * unit;
    else rpb = (rpb + rstep - 1) / rstep * rstep;
    grid = dim3((unsigned)gx, (unsigned)((rows + rpb - 1) / rpb));
    return Ew2d{lg, rpb};
}
// 8 elements per thread need 16-byte alignment in bf16: everything a multiple of 8 (fp32 keeps 4 = 16 bytes per access)
inline bool vec8_ok(int dtype, int C, int p0, int o0, int p1 = 0, int o1 = 0) {
    return dtype == SRGAN_BF16 && ((C | p0 | o0 | p1 | o1) & 7) == 0;
}

// vector average pooling: a thread owns W consecutive channels of one output (forward) / input (backward) pixel
template <typename T, int W>
__global__ void __launch_bounds__(256) avgpool_vec_kernel(const T* __restrict__ x, int x_pitch, T* __restrict__ y, int y_pitch, int y_c0,
                                                          int n, int H, int Wd, int C, int k) {
    const int Ho = H / k, Wo = Wd / k, cg = C / W;
    const long long total = (long long)n * Ho * Wo * cg;
    const float inv = 1.f / (float)(k * k);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cg) * W;
        long long t = i / cg;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const long long b = t / Ho;
        float acc[W];
#pragma unroll
        for (int q = 0; q < W; ++q) acc[q] = 0.f;
        const T* xb = x + ((b * H + (long long)ho * k) * Wd + (long long)wo * k) * x_pitch + c;
        if (k == 2) {                                         // the DenseNet transitions: four independent loads
            Raw<T, W> r[4];
#pragma unroll
            for (int d = 0; d < 4; ++d) r[d] = ld_raw<T, W>(xb + ((long long)(d >> 1) * Wd + (d & 1)) * x_pitch);
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                float v[W];
                unpack(r[d], v);
#pragma unroll
                for (int q = 0; q < W; ++q) acc[q] += v[q];
            }
        } else {
            for (int dh = 0; dh < k; ++dh)
                for (int dw = 0; dw < k; ++dw) {
                    float v[W];
                    unpack(ld_raw<T, W>(xb + ((long long)dh * Wd + dw) * x_pitch), v);
#pragma unroll
                    for (int q = 0; q < W; ++q) acc[q] += v[q];
                }
        }
#pragma unroll
        for (int q = 0; q < W; ++q) acc[q] *= inv;
        stw<W>(y + ((b * Ho + ho) * Wo + wo) * y_pitch + y_c0 + c, acc);
    }
}
template <typename T, int W>
__global__ void __launch_bounds__(256) avgpool_bwd_vec_kernel(const T* __restrict__ dy, int dy_pitch, int dy_c0, T* __restrict__ dx,
                                                              int x_pitch, int n, int H, int Wd, int C, int k,
                                                              const T* __restrict__ href, int act, float slope) {
    const int Ho = H / k, Wo = Wd / k, cg = C / W;
    const long long total = (long long)n * H * Wd * cg;
    const float inv = 1.f / (float)(k * k);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cg) * W;
        long long t = i / cg;
        const int w = (int)(t % Wd); t /= Wd;
        const int h = (int)(t % H);
        const long long b = t / H;
        const long long xi = ((b * H + h) * Wd + w) * x_pitch + c;
        const Raw<T, W> dr = ld_raw<T, W>(dy + ((b * Ho + h / k) * Wo + w / k) * dy_pitch + dy_c0 + c);
        float v[W];
        unpack(dr, v);
        if (href) {
            float hv[W];
            unpack(ld_raw<T, W>(href + xi), hv);
#pragma unroll
            for (int q = 0; q < W; ++q) v[q] *= inv * act_bwd(hv[q], act, slope);
        } else {
#pragma unroll
            for (int q = 0; q < W; ++q) v[q] *= inv;
        }
        stw<W>(dx + xi, v);
    }
}

// depth-to-space of a one-channel map: img[n, i*k+r, j*k+s] <-> blk[n, i, j, r*k+s]   (one thread per image pixel)
template <typename T>
__global__ void __launch_bounds__(256) depth_to_space_kernel(const T* __restrict__ src, T* __restrict__ dst, int n, int Hs, int Ws,
                                                             int k, int inverse) {
    const int Wl = Ws * k, Hl = Hs * k;
    const long long total = (long long)n * Hl * Wl;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % Wl);
        long long t = i / Wl;
        const int y = (int)(t % Hl);
        const long long b = t / Hl;
        const long long blk = ((b * Hs + y / k) * Ws + x / k) * (k * k) + (y % k) * k + (x % k);
        if (inverse) dst[blk] = src[i];
        else dst[i] = src[blk];
    }
}

inline bool vec_ok(int C, int p0, int o0, int p1 = 0, int o1 = 0) { return ((C | p0 | o0 | p1 | o1) & 3) == 0; }

}  // namespace

#define DISPATCH_T(dtype, ...)                                                          \
    do {                                                                                \
        if (dtype == SRGAN_F32) { typedef float T; __VA_ARGS__; }                       \
        else if (dtype == SRGAN_BF16) { typedef bf16 T; __VA_ARGS__; }                  \
        else { srgan_set_error("unknown dtype %d", dtype); return SRGAN_ERR_ARG; }      \
    } while (0)

extern "C" {

int srgan_affine(const void* x, int x_pitch, int x_c0, void* y, int y_pitch, long long rows, int C, const float* gamma,
                 const float* beta, const float* mean, const float* var, float eps, const void* href, int mode, int act,
                 float slope, int dtype, void* stream) {
    SRGAN_REQUIRE(x && y && gamma && var && rows >= 0 && C > 0 && x_c0 >= 0 && x_c0 + C <= x_pitch && C <= y_pitch,
                  "srgan_affine: bad arguments");
    SRGAN_REQUIRE(mode == 0 ? (beta && mean) : (mode == 1 && href), "srgan_affine: mode 0 needs beta/mean, mode 1 needs href");
    if (rows == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = vec_ok(C, x_pitch, x_c0, y_pitch);
    dim3 grid;
#define AFFINE2D(W_)                                                                                                       \
    do {                                                                                                                   \
        const Ew2d e = ew2d_plan(rows, C, W_, grid);                                                                       \
        if (mode == 0) affine2d_kernel<T, W_, 0><<<grid, 256, 0, st>>>((const T*)x, x_pitch, x_c0, (T*)y, y_pitch, rows, C, gamma, beta, mean, var, eps, (const T*)href, act, slope, e); \
        else affine2d_kernel<T, W_, 1><<<grid, 256, 0, st>>>((const T*)x, x_pitch, x_c0, (T*)y, y_pitch, rows, C, gamma, beta, mean, var, eps, (const T*)href, act, slope, e); \
    } while (0)
    DISPATCH_T(dtype,
               if (vec8_ok(dtype, C, x_pitch, x_c0, y_pitch)) AFFINE2D(8);
               else if (vec) AFFINE2D(4);
               else affine_kernel<T, false><<<ew_grid(rows * C), 256, 0, st>>>((const T*)x, x_pitch, x_c0, (T*)y, y_pitch, rows, C, gamma, beta, mean, var, eps, (const T*)href, mode, act, slope));
#undef AFFINE2D
    SRGAN_CHECK_LAUNCH("affine_kernel");
    return SRGAN_OK;
}

int srgan_affine_bwd(const void* dy, int dy_pitch, void* dx, int dx_pitch, int dx_c0, long long rows, int C, const float* gamma,
                     const float* var, float eps, int accumulate, int dtype, void* stream) {
    SRGAN_REQUIRE(dy && dx && gamma && var && rows >= 0 && C > 0 && dx_c0 >= 0 && dx_c0 + C <= dx_pitch && C <= dy_pitch,
                  "srgan_affine_bwd: bad arguments");
    if (rows == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = vec_ok(C, dx_pitch, dx_c0, dy_pitch);
    dim3 grid;
#define AFFINE_BWD2D(W_)                                                                                                   \
    do {                                                                                                                   \
        const Ew2d e = ew2d_plan(rows, C, W_, grid);                                                                       \
        if (accumulate) affine_bwd2d_kernel<T, W_, true><<<grid, 256, 0, st>>>((const T*)dy, dy_pitch, (T*)dx, dx_pitch, dx_c0, rows, C, gamma, var, eps, e); \
        else affine_bwd2d_kernel<T, W_, false><<<grid, 256, 0, st>>>((const T*)dy, dy_pitch, (T*)dx, dx_pitch, dx_c0, rows, C, gamma, var, eps, e); \
    } while (0)
    DISPATCH_T(dtype,
               if (vec8_ok(dtype, C, dx_pitch, dx_c0, dy_pitch)) AFFINE_BWD2D(8);
               else if (vec) AFFINE_BWD2D(4);
               else affine_bwd_kernel<T, false><<<ew_grid(rows * C), 256, 0, st>>>((const T*)dy, dy_pitch, (T*)dx, dx_pitch, dx_c0, rows, C, gamma, var, eps, accumulate));
#undef AFFINE_BWD2D
    SRGAN_CHECK_LAUNCH("affine_bwd_kernel");
    return SRGAN_OK;
}

int srgan_affine_grad(const void* dy, int dy_pitch, const void* x, int x_pitch, int x_c0, long long rows, int C, const float* mean,
                      const float* var, float eps, float* dgamma, float* dbeta, int subtract_mean, int dtype, void* stream) {
    SRGAN_REQUIRE(dy && x && var && dgamma && rows >= 0 && C > 0 && x_c0 >= 0 && x_c0 + C <= x_pitch && C <= dy_pitch && (!subtract_mean || mean),
                  "srgan_affine_grad: bad arguments");
    if (rows == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = vec_ok(C, x_pitch, x_c0, dy_pitch);
    dim3 grid;
    if (vec) {
        DISPATCH_T(dtype,
                   if (vec8_ok(dtype, C, x_pitch, x_c0, dy_pitch)) { const Ew2d e = ew2d_plan(rows, C, 8, grid, 2); affine_grad2d_kernel<T, 8, false><<<grid, 256, 0, st>>>((const T*)dy, dy_pitch, (const T*)x, nullptr, x_pitch, x_c0, rows, C, nullptr, mean, var, eps, dgamma, dbeta, subtract_mean, 0, e); }
                   else { const Ew2d e = ew2d_plan(rows, C, 4, grid, 2); affine_grad2d_kernel<T, 4, false><<<grid, 256, 0, st>>>((const T*)dy, dy_pitch, (const T*)x, nullptr, x_pitch, x_c0, rows, C, nullptr, mean, var, eps, dgamma, dbeta, subtract_mean, 0, e); });
    } else {
        const int gx = (C + 31) / 32;
        long long want = (4LL * kNumSMs + gx - 1) / gx;                 // ~4 CTAs per SM in total
        long long rpb = (rows + want - 1) / want;
        if (rpb < 64) rpb = 64;
        grid = dim3(gx, (unsigned)((rows + rpb - 1) / rpb));
        DISPATCH_T(dtype, affine_grad_kernel<T, false><<<grid, 256, 0, st>>>((const T*)dy, dy_pitch, (const T*)x, x_pitch, x_c0, rows, C, mean, var, eps, dgamma, dbeta, subtract_mean, rpb));
    }
    SRGAN_CHECK_LAUNCH("affine_grad_kernel");
    return SRGAN_OK;
}

int srgan_affine_bwd_grad(const void* dy, int dy_pitch, const void* x, void* dx, int x_pitch, int x_c0, long long rows, int C,
                          const float* gamma, const float* mean, const float* var, float eps, float* dgamma, float* dbeta,
                          int accumulate, int dtype, void* stream) {
    SRGAN_REQUIRE(dy && x && dx && gamma && mean && var && dgamma && dbeta && rows >= 0 && C > 0 && x_c0 >= 0 && x_c0 + C <= x_pitch &&
                      C <= dy_pitch, "srgan_affine_bwd_grad: bad arguments");
    if (rows == 0) return SRGAN_OK;
    if (!vec_ok(C, x_pitch, x_c0, dy_pitch)) {            // unaligned slices: the two separate kernels
        int rc = srgan_affine_grad(dy, dy_pitch, x, x_pitch, x_c0, rows, C, mean, var, eps, dgamma, dbeta, 1, dtype, stream);
        if (rc) return rc;
        return srgan_affine_bwd(dy, dy_pitch, dx, x_pitch, x_c0, rows, C, gamma, var, eps, accumulate, dtype, stream);
    }
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid;
    DISPATCH_T(dtype,
               if (vec8_ok(dtype, C, x_pitch, x_c0, dy_pitch)) { const Ew2d e = ew2d_plan(rows, C, 8, grid, 2); affine_grad2d_kernel<T, 8, true><<<grid, 256, 0, st>>>((const T*)dy, dy_pitch, (const T*)x, (T*)dx, x_pitch, x_c0, rows, C, gamma, mean, var, eps, dgamma, dbeta, 1, accumulate, e); }
               else { const Ew2d e = ew2d_plan(rows, C, 4, grid, 2); affine_grad2d_kernel<T, 4, true><<<grid, 256, 0, st>>>((const T*)dy, dy_pitch, (const T*)x, (T*)dx, x_pitch, x_c0, rows, C, gamma, mean, var, eps, dgamma, dbeta, 1, accumulate, e); });
    SRGAN_CHECK_LAUNCH("affine_bwd_grad_kernel");
    return SRGAN_OK;
}

int srgan_copy2d(const void* src, int src_pitch, int src_c0, void* dst, int dst_pitch, int dst_c0, long long rows, int C,
                 int accumulate, int dtype, void* stream) {
    SRGAN_REQUIRE(src && dst && rows >= 0 && C > 0 && src_c0 >= 0 && dst_c0 >= 0 && src_c0 + C <= src_pitch && dst_c0 + C <= dst_pitch,
                  "srgan_copy2d: bad arguments");
    if (rows == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = vec_ok(C, src_pitch, src_c0, dst_pitch, dst_c0);
    dim3 grid;
#define COPY2D2D(W_)                                                                                                       \
    do {                                                                                                                   \
        const Ew2d e = ew2d_plan(rows, C, W_, grid);                                                                       \
        if (accumulate) copy2d2d_kernel<T, W_, true><<<grid, 256, 0, st>>>((const T*)src, src_pitch, src_c0, (T*)dst, dst_pitch, dst_c0, rows, C, e); \
        else copy2d2d_kernel<T, W_, false><<<grid, 256, 0, st>>>((const T*)src, src_pitch, src_c0, (T*)dst, dst_pitch, dst_c0, rows, C, e); \
    } while (0)
    DISPATCH_T(dtype,
               if (vec8_ok(dtype, C, src_pitch, src_c0, dst_pitch, dst_c0)) COPY2D2D(8);
               else if (vec) COPY2D2D(4);
               else copy2d_kernel<T, false><<<ew_grid(rows * C), 256, 0, st>>>((const T*)src, src_pitch, src_c0, (T*)dst, dst_pitch, dst_c0, rows, C, accumulate));
#undef COPY2D2D
    SRGAN_CHECK_LAUNCH("copy2d_kernel");
    return SRGAN_OK;
}

// ---- 16-byte versions of the index-map kernels (bf16, C a multiple of 8): one block per image row of the output (forward /
// tangent) or input (backward), threads over (column, 8-channel group) with 32-bit index arithmetic only.  The 4-element
// kernels above spent their time in 64-bit divisions (ncu launch list: 928 us for the stem pool backward of 256 samples =
// 1.05 TB/s of the ~1 GB it moves).
__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(h[q]); v[2 * q] = f.x; v[2 * q + 1] = f.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
    return r;
}

__global__ void __launch_bounds__(256) maxpool_idx8_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int y_pitch, int y_c0,
                                                           unsigned char* __restrict__ idx, int write_idx, int H, int W, int C,
                                                           int Ho, int Wo, int k, int s, int p) {
    const int CV = C >> 3;
    const int ho = blockIdx.x % Ho, b = blockIdx.x / Ho;
    const bf16* xb = x + (long long)b * H * W * C;
    const long long orow = ((long long)b * Ho + ho) * Wo;
    const int h0 = ho * s - p;
    for (int t = threadIdx.x; t < Wo * CV; t += 256) {
        const int wo = t / CV, c = (t - wo * CV) << 3;
        const int w0 = wo * s - p;
        float out[8];
        if (write_idx) {
            float best[8];
            unsigned char win[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) { best[q] = -INFINITY; win[q] = 255; out[q] = 0.f; }
            for (int dh = 0; dh < k; ++dh) {
                const int h = h0 + dh;
                if (h < 0 || h >= H) continue;
                for (int dw = 0; dw < k; ++dw) {
                    const int w = w0 + dw;
                    if (w < 0 || w >= W) continue;
                    float v[8];
                    unpack8(*reinterpret_cast<const uint4*>(xb + ((long long)h * W + w) * C + c), v);
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (v[q] > best[q] || win[q] == 255) { best[q] = v[q]; win[q] = (unsigned char)(dh * k + dw); out[q] = v[q]; }
                }
            }
            uint2 wv;
            wv.x = win[0] | (win[1] << 8) | (win[2] << 16) | ((unsigned)win[3] << 24);
            wv.y = win[4] | (win[5] << 8) | (win[6] << 16) | ((unsigned)win[7] << 24);
            *reinterpret_cast<uint2*>(idx + (orow + wo) * C + c) = wv;
        } else {
            const uint2 wv = *reinterpret_cast<const uint2*>(idx + (orow + wo) * C + c);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int wq = ((q < 4 ? wv.x : wv.y) >> (8 * (q & 3))) & 255;
                out[q] = wq == 255 ? 0.f : to_f(xb[((long long)(h0 + wq / k) * W + (w0 + wq % k)) * C + c + q]);
            }
        }
        *reinterpret_cast<uint4*>(y + (orow + wo) * y_pitch + y_c0 + c) = pack8(out);
    }
}

__global__ void __launch_bounds__(256) maxpool_bwd_idx8_kernel(const bf16* __restrict__ xref, const unsigned char* __restrict__ idx,
                                                               const bf16* __restrict__ dy, int dy_pitch, int dy_c0,
                                                               bf16* __restrict__ dx, int H, int W, int C, int Ho, int Wo, int k,
                                                               int s, int p, int act, float slope) {
    const int CV = C >> 3;
    const int h = blockIdx.x % H, b = blockIdx.x / H;
    int ho_lo = (h + p - k + 1 + s - 1) / s, ho_hi = (h + p) / s;
    if (h + p - k + 1 < 0) ho_lo = 0;
    ho_hi = min(ho_hi, Ho - 1);
    const long long xrow = ((long long)b * H + h) * W;
    for (int t = threadIdx.x; t < W * CV; t += 256) {
        const int w = t / CV, c = (t - w * CV) << 3;
        int wo_lo = (w + p - k + 1 + s - 1) / s, wo_hi = (w + p) / s;
        if (w + p - k + 1 < 0) wo_lo = 0;
        wo_hi = min(wo_hi, Wo - 1);
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.f;
        for (int ho = ho_lo; ho <= ho_hi; ++ho)
            for (int wo = wo_lo; wo <= wo_hi; ++wo) {
                const unsigned me = (unsigned)((h - (ho * s - p)) * k + (w - (wo * s - p)));   // this input's position inside that window
                const long long oi = ((long long)b * Ho + ho) * Wo + wo;
                const uint2 wv = *reinterpret_cast<const uint2*>(idx + oi * C + c);
                const unsigned ex = wv.x ^ (me * 0x01010101u), ey = wv.y ^ (me * 0x01010101u);      // a zero byte = a match
                if ((((ex - 0x01010101u) & ~ex) | ((ey - 0x01010101u) & ~ey)) & 0x80808080u) {} else
                    continue;                                   // no byte of the index map equals `me`: this window routes nothing here
                float d[8];
                unpack8(*reinterpret_cast<const uint4*>(dy + oi * dy_pitch + dy_c0 + c), d);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if ((((q < 4 ? wv.x : wv.y) >> (8 * (q & 3))) & 255u) == me) acc[q] += d[q];
            }
        float hx[8], o[8];
        unpack8(*reinterpret_cast<const uint4*>(xref + (xrow + w) * C + c), hx);
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = acc[q] * act_bwd(hx[q], act, slope);
        *reinterpret_cast<uint4*>(dx + (xrow + w) * C + c) = pack8(o);
    }
}

int srgan_maxpool(const void* x, const void* xref, void* y, int y_pitch, int y_c0, unsigned char* idx, int idx_mode, int n, int H,
                  int W, int C, int k, int stride, int pad, int dtype, void* stream) {
    SRGAN_REQUIRE(x && y && n >= 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0 && pad >= 0 && 2 * pad <= k && y_c0 >= 0 &&
                      y_c0 + C <= y_pitch, "srgan_maxpool: bad arguments");
    if (n == 0) return SRGAN_OK;
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = vec_ok(C, y_pitch, y_c0);
    SRGAN_REQUIRE(idx_mode >= 0 && idx_mode <= 2 && (idx_mode == 0 || idx) && k * k < 255, "srgan_maxpool: bad index-map arguments");
    if (idx_mode != 0 && dtype == SRGAN_BF16 && ((C | y_pitch | y_c0) & 7) == 0 && (long long)n * Ho < 0x7fffffffLL) {
        maxpool_idx8_kernel<<<(unsigned)(n * Ho), 256, 0, st>>>((const bf16*)x, (bf16*)y, y_pitch, y_c0, idx, idx_mode == 1, H, W, C, Ho, Wo,
                                                              k, stride, pad);
        SRGAN_CHECK_LAUNCH("maxpool_idx8_kernel");
        return SRGAN_OK;
    }
    if (idx_mode != 0) {
        DISPATCH_T(dtype,
                   if (vec) maxpool_idx_kernel<T, 4><<<ew_grid((long long)n * Ho * Wo * C / 4), 256, 0, st>>>((const T*)x, (T*)y, y_pitch, y_c0, idx, idx_mode == 1, n, H, W, C, Ho, Wo, k, stride, pad);
                   else maxpool_idx_kernel<T, 1><<<ew_grid((long long)n * Ho * Wo * C), 256, 0, st>>>((const T*)x, (T*)y, y_pitch, y_c0, idx, idx_mode == 1, n, H, W, C, Ho, Wo, k, stride, pad));
        SRGAN_CHECK_LAUNCH("maxpool_idx_kernel");
        return SRGAN_OK;
    }
    DISPATCH_T(dtype,
               if (vec) maxpool4_kernel<T><<<ew_grid((long long)n * Ho * Wo * C / 4), 256, 0, st>>>((const T*)x, (const T*)(xref ? xref : x), (T*)y, y_pitch, y_c0, n, H, W, C, Ho, Wo, k, stride, pad);
               else maxpool_kernel<T><<<ew_grid((long long)n * Ho * Wo * C), 256, 0, st>>>((const T*)x, (const T*)(xref ? xref : x), (T*)y, y_pitch, y_c0, n, H, W, C, Ho, Wo, k, stride, pad));
    SRGAN_CHECK_LAUNCH("maxpool_kernel");
    return SRGAN_OK;
}

int srgan_maxpool_bwd(const void* xref, const unsigned char* idx, const void* dy, int dy_pitch, int dy_c0, void* dx, int n, int H, int W,
                      int C, int k, int stride, int pad, int act, float slope, int dtype, void* stream) {
    SRGAN_REQUIRE(xref && dy && dx && n >= 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0 && pad >= 0 && 2 * pad <= k && dy_c0 >= 0 &&
                      dy_c0 + C <= dy_pitch, "srgan_maxpool_bwd: bad arguments");
    if (n == 0) return SRGAN_OK;
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = vec_ok(C, dy_pitch, dy_c0);
    if (idx && dtype == SRGAN_BF16 && ((C | dy_pitch | dy_c0) & 7) == 0 && (long long)n * H < 0x7fffffffLL) {
        maxpool_bwd_idx8_kernel<<<(unsigned)(n * H), 256, 0, st>>>((const bf16*)xref, idx, (const bf16*)dy, dy_pitch, dy_c0, (bf16*)dx, H, W,
                                                                 C, Ho, Wo, k, stride, pad, act, slope);
        SRGAN_CHECK_LAUNCH("maxpool_bwd_idx8_kernel");
        return SRGAN_OK;
    }
    if (idx) {
        DISPATCH_T(dtype,
                   if (vec) maxpool_bwd_idx_kernel<T, 4><<<ew_grid((long long)n * H * W * C / 4), 256, 0, st>>>((const T*)xref, idx, (const T*)dy, dy_pitch, dy_c0, (T*)dx, n, H, W, C, Ho, Wo, k, stride, pad, act, slope);
                   else maxpool_bwd_idx_kernel<T, 1><<<ew_grid((long long)n * H * W * C), 256, 0, st>>>((const T*)xref, idx, (const T*)dy, dy_pitch, dy_c0, (T*)dx, n, H, W, C, Ho, Wo, k, stride, pad, act, slope));
        SRGAN_CHECK_LAUNCH("maxpool_bwd_idx_kernel");
        return SRGAN_OK;
    }
    DISPATCH_T(dtype,
               if (vec) maxpool_bwd4_kernel<T><<<ew_grid((long long)n * H * W * C / 4), 256, 0, st>>>((const T*)xref, (const T*)dy, dy_pitch, dy_c0, (T*)dx, n, H, W, C, Ho, Wo, k, stride, pad, act, slope);
               else maxpool_bwd_kernel<T><<<ew_grid((long long)n * H * W * C), 256, 0, st>>>((const T*)xref, (const T*)dy, dy_pitch, dy_c0, (T*)dx, n, H, W, C, Ho, Wo, k, stride, pad, act, slope));
    SRGAN_CHECK_LAUNCH("maxpool_bwd_kernel");
    return SRGAN_OK;
}

int srgan_avgpool(const void* x, int x_pitch, void* y, int y_pitch, int y_c0, int n, int H, int W, int C, int k, int dtype,
                  void* stream) {
    SRGAN_REQUIRE(x && y && n >= 0 && H > 0 && W > 0 && C > 0 && k > 0 && H % k == 0 && W % k == 0 && y_c0 >= 0 && y_c0 + C <= y_pitch &&
                      C <= x_pitch,
                  "srgan_avgpool: bad arguments (the window must tile the input)");
    if (n == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_T(dtype,
               if (vec8_ok(dtype, C, x_pitch, y_pitch, y_c0)) avgpool_vec_kernel<T, 8><<<ew_grid((long long)n * (H / k) * (W / k) * C / 8), 256, 0, st>>>((const T*)x, x_pitch, (T*)y, y_pitch, y_c0, n, H, W, C, k);
               else if (vec_ok(C, x_pitch, y_pitch, y_c0)) avgpool_vec_kernel<T, 4><<<ew_grid((long long)n * (H / k) * (W / k) * C / 4), 256, 0, st>>>((const T*)x, x_pitch, (T*)y, y_pitch, y_c0, n, H, W, C, k);
               else avgpool_kernel<T><<<ew_grid((long long)n * (H / k) * (W / k) * C), 256, 0, st>>>((const T*)x, x_pitch, (T*)y, y_pitch, y_c0, n, H, W, C, k));
    SRGAN_CHECK_LAUNCH("avgpool_kernel");
    return SRGAN_OK;
}

int srgan_avgpool_bwd(const void* dy, int dy_pitch, int dy_c0, void* dx, int x_pitch, int n, int H, int W, int C, int k,
                      const void* href, int act, float slope, int dtype, void* stream) {
    SRGAN_REQUIRE(dy && dx && n >= 0 && H > 0 && W > 0 && C > 0 && k > 0 && H % k == 0 && W % k == 0 && dy_c0 >= 0 && dy_c0 + C <= dy_pitch &&
                      C <= x_pitch,
                  "srgan_avgpool_bwd: bad arguments");
    if (n == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const void* hr = act == SRGAN_ACT_NONE ? nullptr : href;
    DISPATCH_T(dtype,
               if (vec8_ok(dtype, C, x_pitch, dy_pitch, dy_c0)) avgpool_bwd_vec_kernel<T, 8><<<ew_grid((long long)n * H * W * C / 8), 256, 0, st>>>((const T*)dy, dy_pitch, dy_c0, (T*)dx, x_pitch, n, H, W, C, k, (const T*)hr, act, slope);
               else if (vec_ok(C, x_pitch, dy_pitch, dy_c0)) avgpool_bwd_vec_kernel<T, 4><<<ew_grid((long long)n * H * W * C / 4), 256, 0, st>>>((const T*)dy, dy_pitch, dy_c0, (T*)dx, x_pitch, n, H, W, C, k, (const T*)hr, act, slope);
               else avgpool_bwd_kernel<T><<<ew_grid((long long)n * H * W * C), 256, 0, st>>>((const T*)dy, dy_pitch, dy_c0, (T*)dx, x_pitch, n, H, W, C, k, (const T*)hr, act, slope));
    SRGAN_CHECK_LAUNCH("avgpool_bwd_kernel");
    return SRGAN_OK;
}

int srgan_depth_to_space(const void* src, void* dst, int n, int Hs, int Ws, int k, int inverse, int dtype, void* stream) {
    SRGAN_REQUIRE(src && dst && n >= 0 && Hs > 0 && Ws > 0 && k > 0, "srgan_depth_to_space: bad arguments");
    if (n == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_T(dtype, depth_to_space_kernel<T><<<ew_grid((long long)n * Hs * Ws * k * k), 256, 0, st>>>((const T*)src, (T*)dst, n, Hs, Ws, k, inverse));
    SRGAN_CHECK_LAUNCH("depth_to_space_kernel");
    return SRGAN_OK;
}

int srgan_crowd_loss(const float* pred, const float* density, const void* const* maps, int nmaps, const float* map_label, int B,
                     long long HW, int order, float scale, float map_mult, float* loss, float* dpred, float* dm, int dtype,
                     void* stream) {
    SRGAN_REQUIRE(pred && density && maps && map_label && loss && dpred && dm && B >= 0 && HW > 0 && nmaps >= 1 && nmaps <= 4 && order >= 1,
                  "srgan_crowd_loss: bad arguments");
    if (B == 0) return SRGAN_OK;
    MapPtrs mp;
    for (int j = 0; j < 4; ++j) mp.p[j] = j < nmaps ? maps[j] : nullptr;
    for (int j = 0; j < nmaps; ++j) SRGAN_REQUIRE(mp.p[j], "srgan_crowd_loss: null map pointer");
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_T(dtype, crowd_loss_kernel<T><<<B, 1024, 0, st>>>(pred, density, mp, nmaps, map_label, HW, order, scale, map_mult, loss, dpred, dm));
    SRGAN_CHECK_LAUNCH("crowd_loss_kernel");
    return SRGAN_OK;
}

int srgan_crowd_map_grad(const void* map, const float* map_label, const float* dm, void* delta, int B, long long HW, int nmaps, int act,
                         float slope, int dtype, void* stream) {
    SRGAN_REQUIRE(map && map_label && dm && delta && B >= 0 && HW > 0 && nmaps >= 1, "srgan_crowd_map_grad: bad arguments");
    if (B == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_T(dtype, crowd_map_grad_kernel<T><<<ew_grid((long long)B * HW), 256, 0, st>>>((const T*)map, map_label, dm, (T*)delta, B, HW, 1.f / (float)nmaps, act, slope));
    SRGAN_CHECK_LAUNCH("crowd_map_grad_kernel");
    return SRGAN_OK;
}

}  // extern "C"
