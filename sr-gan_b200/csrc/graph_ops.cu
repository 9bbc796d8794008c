// Bandwidth-bound kernels of the graph discriminator (crowd KnnDenseNetCat, crowd/models.py:1049-1166):
//   eval-mode BatchNorm affine (+ReLU) forward / tangent / backward / parameter gradients (srgan.py:538-542 freezes the
//   statistics, weight and bias stay trainable), channel-slice copies (the in-place concat of the dense blocks), max / average
//   pooling with their transposes, and the crowd labeled loss (crowd/srgan.py:247-254).
// All activations are NHWC rows; a "slice" is the channel range [c0, c0+C) of rows with pitch `pitch` elements.
// Vectorised (4 elements per thread) whenever C, the pitches and the offsets are multiples of 4.
#include "common.cuh"

namespace {

__device__ __forceinline__ float4 ld4f(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float bn_scale(float gamma, float var, float eps) { return gamma / sqrtf(var + eps); }

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
                o.x = act_fwd((xv.x - m.x) * s.x + b.x, act, slope); o.y = act_fwd((xv.y - m.y) * s.y + b.y, act, slope);
                o.z = act_fwd((xv.z - m.z) * s.z + b.z, act, slope); o.w = act_fwd((xv.w - m.w) * s.w + b.w, act, slope);
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
            if (mode == 0) o = act_fwd((xv - mean[c]) * s + beta[c], act, slope);
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

// fused backward of one eval-BatchNorm(+ReLU) node: the parameter gradients AND the data gradient from ONE pass over dy
//   dgamma[c] += sum_r dy[r,c]*(x[r,c0+c]-mean[c])/sqrt(var[c]+eps) ; dbeta[c] += sum_r dy[r,c] ; dx[r,c0+c] (+)= dy[r,c]*gamma[c]/sqrt(..)
// same thread mapping as affine_grad_kernel (x and dx are the same slice of the activation / delta buffers)
template <typename T>
__global__ void __launch_bounds__(256) affine_bwd_grad_kernel(const T* __restrict__ dy, int dy_pitch, const T* __restrict__ x,
                                                              T* __restrict__ dx, int x_pitch, int x_c0, long long rows, int C,
                                                              const float* __restrict__ gamma, const float* __restrict__ mean,
                                                              const float* __restrict__ var, float eps, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, int accumulate, long long rows_per_block) {
    __shared__ float sg[8][129], sb[8][129];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + lane) * 4;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(rows, r0 + rows_per_block);
    float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag;
    if (c < C) {
        const float4 mu = ld4f(mean + c), g = ld4f(gamma + c), vr = ld4f(var + c);
        const float4 s = make_float4(bn_scale(g.x, vr.x, eps), bn_scale(g.y, vr.y, eps), bn_scale(g.z, vr.z, eps), bn_scale(g.w, vr.w, eps));
        for (long long r = r0 + w; r < r1; r += 8) {
            const float4 d = ld4(dy + r * dy_pitch + c), xv = ld4(x + r * x_pitch + x_c0 + c);
            ag.x = fmaf(d.x, xv.x - mu.x, ag.x); ag.y = fmaf(d.y, xv.y - mu.y, ag.y);
            ag.z = fmaf(d.z, xv.z - mu.z, ag.z); ag.w = fmaf(d.w, xv.w - mu.w, ag.w);
            ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w;
            T* xp = dx + r * x_pitch + x_c0 + c;
            float4 o = make_float4(d.x * s.x, d.y * s.y, d.z * s.z, d.w * s.w);
            if (accumulate) { const float4 p = ld4(xp); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
            st4(xp, o);
        }
    }
    sg[w][lane * 4] = ag.x; sg[w][lane * 4 + 1] = ag.y; sg[w][lane * 4 + 2] = ag.z; sg[w][lane * 4 + 3] = ag.w;
    sb[w][lane * 4] = ab.x; sb[w][lane * 4 + 1] = ab.y; sb[w][lane * 4 + 2] = ab.z; sb[w][lane * 4 + 3] = ab.w;
    __syncthreads();
    if (threadIdx.x < 128) {
        const int cc = blockIdx.x * 128 + threadIdx.x;
        if (cc < C) {
            float gsum = 0.f, bsum = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) { gsum += sg[k][threadIdx.x]; bsum += sb[k][threadIdx.x]; }
            atomicAdd(dgamma + cc, gsum / sqrtf(var[cc] + eps));
            atomicAdd(dbeta + cc, bsum);
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
// scale / shift are computed once) and walks down the rows; a warp covers 32 consecutive channel groups of a row (512
// contiguous bytes in bf16 with W = 8: one 16-byte access per thread).  grid.x = channel-group tiles, grid.y = row tiles.
// When a row has fewer than 32 groups the spare lanes take further rows: thread t -> (group t & (cvp-1), row t >> lg).
constexpr int EW_RI = 8;                      // row sweeps per block
struct Ew2d { int lg; };                      // log2(cvp), cvp = min(32, next power of two >= C/W)
template <int W>
__device__ __forceinline__ void ew2d_map(const Ew2d e, int& c, long long& r0, int& rstep) {
    const int t = threadIdx.x;
    c = (blockIdx.x * 32 + (t & ((1 << e.lg) - 1))) * W;
    rstep = 256 >> e.lg;
    r0 = (long long)blockIdx.y * (rstep * EW_RI) + (t >> e.lg);
}
// W consecutive elements <-> W floats (16-byte accesses for 8 x bf16 and 4 x fp32)
template <int W> __device__ __forceinline__ void ldw(const float* p, float (&v)[W]) {
#pragma unroll
    for (int q = 0; q < W / 4; ++q) { const float4 a = *reinterpret_cast<const float4*>(p + 4 * q); v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w; }
}
template <int W> __device__ __forceinline__ void ldw(const bf16* p, float (&v)[W]) {
    if (W == 8) {
        const uint4 r = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(h[q]); v[2 * q] = f.x; v[(2 * q + 1) % W] = f.y; }
    } else {
        const float4 a = ld4(p);
        v[0] = a.x; v[1 % W] = a.y; v[2 % W] = a.z; v[3 % W] = a.w;
    }
}
template <int W> __device__ __forceinline__ void stw(float* p, const float (&v)[W]) {
#pragma unroll
    for (int q = 0; q < W / 4; ++q) *reinterpret_cast<float4*>(p + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
template <int W> __device__ __forceinline__ void stw(bf16* p, const float (&v)[W]) {
    if (W == 8) {
        uint4 r;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
        for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(v[2 * q], v[(2 * q + 1) % W]);
        *reinterpret_cast<uint4*>(p) = r;
    } else {
        st4(p, make_float4(v[0], v[1 % W], v[2 % W], v[3 % W]));
    }
}

template <typename T, int W>
__global__ void __launch_bounds__(256) affine2d_kernel(const T* __restrict__ x, int x_pitch, int x_c0, T* __restrict__ y, int y_pitch,
                                                       long long rows, int C, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const float* __restrict__ mean,
                                                       const float* __restrict__ var, float eps, const T* __restrict__ href,
                                                       int mode, int act, float slope, Ew2d e) {
    int c, rstep; long long r0;
    ew2d_map<W>(e, c, r0, rstep);
    if (c >= C) return;
    float s[W], m[W], b[W];
#pragma unroll
    for (int q = 0; q < W; ++q) {
        s[q] = bn_scale(gamma[c + q], var[c + q], eps);
        m[q] = mode == 0 ? mean[c + q] : 0.f;
        b[q] = mode == 0 ? beta[c + q] : 0.f;
    }
#pragma unroll 4
    for (int j = 0; j < EW_RI; ++j) {
        const long long r = r0 + (long long)j * rstep;
        if (r >= rows) break;
        float xv[W], o[W];
        ldw<W>(x + r * x_pitch + x_c0 + c, xv);
        if (mode == 0) {
#pragma unroll
            for (int q = 0; q < W; ++q) o[q] = act_fwd((xv[q] - m[q]) * s[q] + b[q], act, slope);
        } else {
            float h[W];
            ldw<W>(href + r * y_pitch + c, h);
#pragma unroll
            for (int q = 0; q < W; ++q) o[q] = xv[q] * s[q] * act_bwd(h[q], act, slope);
        }
        stw<W>(y + r * y_pitch + c, o);
    }
}

template <typename T, int W>
__global__ void __launch_bounds__(256) affine_bwd2d_kernel(const T* __restrict__ dy, int dy_pitch, T* __restrict__ dx, int dx_pitch,
                                                           int dx_c0, long long rows, int C, const float* __restrict__ gamma,
                                                           const float* __restrict__ var, float eps, int accumulate, Ew2d e) {
    int c, rstep; long long r0;
    ew2d_map<W>(e, c, r0, rstep);
    if (c >= C) return;
    float s[W];
#pragma unroll
    for (int q = 0; q < W; ++q) s[q] = bn_scale(gamma[c + q], var[c + q], eps);
#pragma unroll 4
    for (int j = 0; j < EW_RI; ++j) {
        const long long r = r0 + (long long)j * rstep;
        if (r >= rows) break;
        float d[W], o[W];
        ldw<W>(dy + r * dy_pitch + c, d);
        T* xp = dx + r * dx_pitch + dx_c0 + c;
#pragma unroll
        for (int q = 0; q < W; ++q) o[q] = d[q] * s[q];
        if (accumulate) {
            float pv[W];
            ldw<W>(xp, pv);
#pragma unroll
            for (int q = 0; q < W; ++q) o[q] += pv[q];
        }
        stw<W>(xp, o);
    }
}

template <typename T, int W>
__global__ void __launch_bounds__(256) copy2d2d_kernel(const T* __restrict__ src, int src_pitch, int src_c0, T* __restrict__ dst,
                                                       int dst_pitch, int dst_c0, long long rows, int C, int accumulate, Ew2d e) {
    int c, rstep; long long r0;
    ew2d_map<W>(e, c, r0, rstep);
    if (c >= C) return;
#pragma unroll 4
    for (int j = 0; j < EW_RI; ++j) {
        const long long r = r0 + (long long)j * rstep;
        if (r >= rows) break;
        float v[W];
        ldw<W>(src + r * src_pitch + src_c0 + c, v);
        T* dp = dst + r * dst_pitch + dst_c0 + c;
        if (accumulate) {
            float pv[W];
            ldw<W>(dp, pv);
#pragma unroll
            for (int q = 0; q < W; ++q) v[q] += pv[q];
        }
        stw<W>(dp, v);
    }
}

inline Ew2d ew2d_map_for(int C, int W) {
    int lg = 0;
    while (lg < 5 && (1 << lg) < C / W) ++lg;
    return Ew2d{lg};
}
inline dim3 ew2d_grid(long long rows, int C, int W) {
    const long long rpb = (long long)(256 >> ew2d_map_for(C, W).lg) * EW_RI;
    return dim3((unsigned)((C / W + 31) / 32), (unsigned)((rows + rpb - 1) / rpb));
}
inline bool ew2d_ok(long long rows, int C, int W) {
    const long long rpb = (long long)(256 >> ew2d_map_for(C, W).lg) * EW_RI;
    return (rows + rpb - 1) / rpb <= 65535;
}
// 8 elements per thread need 16-byte alignment in bf16 (and 32-byte runs in fp32): everything a multiple of 8
inline bool vec8_ok(int C, int p0, int o0, int p1 = 0, int o1 = 0) { return ((C | p0 | o0 | p1 | o1) & 7) == 0; }

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
    DISPATCH_T(dtype,
               if (vec8_ok(C, x_pitch, x_c0, y_pitch) && ew2d_ok(rows, C, 8)) affine2d_kernel<T, 8><<<ew2d_grid(rows, C, 8), 256, 0, st>>>((const T*)x, x_pitch, x_c0, (T*)y, y_pitch, rows, C, gamma, beta, mean, var, eps, (const T*)href, mode, act, slope, ew2d_map_for(C, 8));
               else if (vec && ew2d_ok(rows, C, 4)) affine2d_kernel<T, 4><<<ew2d_grid(rows, C, 4), 256, 0, st>>>((const T*)x, x_pitch, x_c0, (T*)y, y_pitch, rows, C, gamma, beta, mean, var, eps, (const T*)href, mode, act, slope, ew2d_map_for(C, 4));
               else if (vec) affine_kernel<T, true><<<ew_grid(rows * C / 4), 256, 0, st>>>((const T*)x, x_pitch, x_c0, (T*)y, y_pitch, rows, C, gamma, beta, mean, var, eps, (const T*)href, mode, act, slope);
               else affine_kernel<T, false><<<ew_grid(rows * C), 256, 0, st>>>((const T*)x, x_pitch, x_c0, (T*)y, y_pitch, rows, C, gamma, beta, mean, var, eps, (const T*)href, mode, act, slope));
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
    DISPATCH_T(dtype,
               if (vec8_ok(C, dx_pitch, dx_c0, dy_pitch) && ew2d_ok(rows, C, 8)) affine_bwd2d_kernel<T, 8><<<ew2d_grid(rows, C, 8), 256, 0, st>>>((const T*)dy, dy_pitch, (T*)dx, dx_pitch, dx_c0, rows, C, gamma, var, eps, accumulate, ew2d_map_for(C, 8));
               else if (vec && ew2d_ok(rows, C, 4)) affine_bwd2d_kernel<T, 4><<<ew2d_grid(rows, C, 4), 256, 0, st>>>((const T*)dy, dy_pitch, (T*)dx, dx_pitch, dx_c0, rows, C, gamma, var, eps, accumulate, ew2d_map_for(C, 4));
               else if (vec) affine_bwd_kernel<T, true><<<ew_grid(rows * C / 4), 256, 0, st>>>((const T*)dy, dy_pitch, (T*)dx, dx_pitch, dx_c0, rows, C, gamma, var, eps, accumulate);
               else affine_bwd_kernel<T, false><<<ew_grid(rows * C), 256, 0, st>>>((const T*)dy, dy_pitch, (T*)dx, dx_pitch, dx_c0, rows, C, gamma, var, eps, accumulate));
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
    const int V = vec ? 4 : 1;
    const int gx = (C + 32 * V - 1) / (32 * V);
    long long want = (4LL * kNumSMs + gx - 1) / gx;                 // ~4 CTAs per SM in total
    long long rpb = (rows + want - 1) / want;
    if (rpb < 64) rpb = 64;
    const long long gy = (rows + rpb - 1) / rpb;
    dim3 grid(gx, (unsigned)gy);
    DISPATCH_T(dtype,
               if (vec) affine_grad_kernel<T, true><<<grid, 256, 0, st>>>((const T*)dy, dy_pitch, (const T*)x, x_pitch, x_c0, rows, C, mean, var, eps, dgamma, dbeta, subtract_mean, rpb);
               else affine_grad_kernel<T, false><<<grid, 256, 0, st>>>((const T*)dy, dy_pitch, (const T*)x, x_pitch, x_c0, rows, C, mean, var, eps, dgamma, dbeta, subtract_mean, rpb));
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
    const int gx = (C + 127) / 128;
    long long want = (8LL * kNumSMs + gx - 1) / gx;
    long long rpb = (rows + want - 1) / want;
    if (rpb < 64) rpb = 64;
    dim3 grid(gx, (unsigned)((rows + rpb - 1) / rpb));
    DISPATCH_T(dtype, affine_bwd_grad_kernel<T><<<grid, 256, 0, st>>>((const T*)dy, dy_pitch, (const T*)x, (T*)dx, x_pitch, x_c0, rows, C, gamma, mean, var, eps, dgamma, dbeta, accumulate, rpb));
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
    DISPATCH_T(dtype,
               if (vec8_ok(C, src_pitch, src_c0, dst_pitch, dst_c0) && ew2d_ok(rows, C, 8)) copy2d2d_kernel<T, 8><<<ew2d_grid(rows, C, 8), 256, 0, st>>>((const T*)src, src_pitch, src_c0, (T*)dst, dst_pitch, dst_c0, rows, C, accumulate, ew2d_map_for(C, 8));
               else if (vec && ew2d_ok(rows, C, 4)) copy2d2d_kernel<T, 4><<<ew2d_grid(rows, C, 4), 256, 0, st>>>((const T*)src, src_pitch, src_c0, (T*)dst, dst_pitch, dst_c0, rows, C, accumulate, ew2d_map_for(C, 4));
               else if (vec) copy2d_kernel<T, true><<<ew_grid(rows * C / 4), 256, 0, st>>>((const T*)src, src_pitch, src_c0, (T*)dst, dst_pitch, dst_c0, rows, C, accumulate);
               else copy2d_kernel<T, false><<<ew_grid(rows * C), 256, 0, st>>>((const T*)src, src_pitch, src_c0, (T*)dst, dst_pitch, dst_c0, rows, C, accumulate));
    SRGAN_CHECK_LAUNCH("copy2d_kernel");
    return SRGAN_OK;
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
    DISPATCH_T(dtype, avgpool_kernel<T><<<ew_grid((long long)n * (H / k) * (W / k) * C), 256, 0, st>>>((const T*)x, x_pitch, (T*)y, y_pitch, y_c0, n, H, W, C, k));
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
    DISPATCH_T(dtype, avgpool_bwd_kernel<T><<<ew_grid((long long)n * H * W * C), 256, 0, st>>>((const T*)dy, dy_pitch, dy_c0, (T*)dx, x_pitch, n, H, W, C, k, (const T*)(act == SRGAN_ACT_NONE ? nullptr : href), act, slope));
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
