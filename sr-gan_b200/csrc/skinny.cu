// "Skinny" linear pair: few outputs (Ca <= 32) over a long reduction axis (K = R*S*Cb >= 512, the full extent of the large
// side), e.g. MapModule.linear1 = Conv2d(32, 20, kernel 28) on a 28x28 map (crowd/models.py:774, K = 25088) and
// final_count_feature_layer = Conv2d(1920, 20, 1) (:1132).  As GEMMs these are M = batch rows, N = 20: a 128x64-tiled kernel runs
// them on ONE CTA.  They are bandwidth problems: each kernel below streams the long operand once, coalesced.
//   down : S[n,a]  = epilogue(sum_k L[n,k] * Wd[a,k])          one CTA per ROWS rows, threads stride over k, packed loads first
//   up   : L[n,k]  = epilogue(sum_a S[n,a] * Wu[b(k)][tap(k)][a])  one thread per (n, k)
//   wgrad: dW[a,k] += sum_n S[n,a] * L[n,k]                    one thread per k and 32-row chunk, all a in registers
#include "common.cuh"

namespace {

constexpr int CA_MAX = 32;
constexpr int ROWS = 2;

// down: one CTA per ROWS rows; the 256 threads stride over K with 16-byte loads and every thread accumulates ALL outputs.  Per
// iteration the ROWS input pieces and the Ca weight pieces are loaded first, kept PACKED (4 registers per 16 bytes), and only
// then unpacked and multiplied: Ca + ROWS independent 16-byte loads in flight per thread.  (The first version unpacked each
// weight piece right after its load and the compiler serialised the load -> FMA pairs: 160 us per launch whatever the batch.)
template <typename T>
__global__ void __launch_bounds__(256) skinny_down_kernel(const T* __restrict__ L, const T* __restrict__ Wd, T* __restrict__ out,
                                                          const float* __restrict__ bias, int bias_mod, const T* __restrict__ href,
                                                          int epi, int act, float slope, int n, int Ca, long long K) {
    constexpr int V = 16 / sizeof(T);                 // elements per 16-byte load
    __shared__ float red[8][ROWS][CA_MAX];
    const int row0 = blockIdx.x * ROWS;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float acc[ROWS][CA_MAX];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int a = 0; a < CA_MAX; ++a) acc[r][a] = 0.f;
    auto unpack = [](const uint4& raw, float (&v)[V]) {
        if constexpr (sizeof(T) == 2) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
            for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(h[q]); v[2 * q] = f.x; v[2 * q + 1] = f.y; }
        } else {
            v[0] = __uint_as_float(raw.x); v[1] = __uint_as_float(raw.y); v[2] = __uint_as_float(raw.z); v[3] = __uint_as_float(raw.w);
        }
    };
    if ((K % V) == 0 && ((reinterpret_cast<uintptr_t>(L) | reinterpret_cast<uintptr_t>(Wd)) & 15) == 0) {
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        for (long long k = (long long)threadIdx.x * V; k < K; k += 256 * V) {
            uint4 xr[ROWS], wr[CA_MAX];
#pragma unroll
            for (int r = 0; r < ROWS; ++r)
                xr[r] = row0 + r < n ? *reinterpret_cast<const uint4*>(L + (long long)(row0 + r) * K + k) : zero;
#pragma unroll
            for (int a = 0; a < CA_MAX; ++a)
                wr[a] = a < Ca ? __ldg(reinterpret_cast<const uint4*>(Wd + (long long)a * K + k)) : zero;
            float x[ROWS][V];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) unpack(xr[r], x[r]);
#pragma unroll
            for (int a = 0; a < CA_MAX; ++a) {
                float wv[V];
                unpack(wr[a], wv);
#pragma unroll
                for (int r = 0; r < ROWS; ++r)
#pragma unroll
                    for (int q = 0; q < V; ++q) acc[r][a] = fmaf(x[r][q], wv[q], acc[r][a]);
            }
        }
    } else {
        for (long long k = threadIdx.x; k < K; k += 256) {
            float x[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) x[r] = row0 + r < n ? to_f(L[(long long)(row0 + r) * K + k]) : 0.f;
#pragma unroll
            for (int a = 0; a < CA_MAX; ++a) {
                if (a < Ca) {
                    const float wv = to_f(Wd[(long long)a * K + k]);
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) acc[r][a] = fmaf(x[r], wv, acc[r][a]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int a = 0; a < CA_MAX; ++a) {
            const float v = warp_sum(acc[r][a]);
            if (lane == 0) red[w][r][a] = v;
        }
    __syncthreads();
    if (threadIdx.x < ROWS * CA_MAX) {
        const int r = threadIdx.x / CA_MAX, a = threadIdx.x % CA_MAX;
        if (a < Ca && row0 + r < n) {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) v += red[k][r][a];
            const long long o = (long long)(row0 + r) * Ca + a;
            if (epi == SRGAN_EPI_BIAS_ACT) {
                if (bias) v += bias[bias_mod ? a % bias_mod : a];
                v = act_fwd(v, act, slope);
            } else if (href && act != SRGAN_ACT_NONE) {
                v *= act_bwd(to_f(href[o]), act, slope);
            }
            out[o] = from_f<T>(v);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) skinny_up_kernel(const T* __restrict__ S, const T* __restrict__ Wu, T* __restrict__ out,
                                                        const float* __restrict__ bias, int bias_mod, const T* __restrict__ href,
                                                        int epi, int act, float slope, int Ca, int Cb, long long K) {
    __shared__ float s[CA_MAX];
    const int row = blockIdx.y;
    if (threadIdx.x < Ca) s[threadIdx.x] = to_f(S[(long long)row * Ca + threadIdx.x]);
    __syncthreads();
    const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
    if (k >= K) return;
    // Wu is [b][r][s][a] while k runs over (r, s, b) like the large side's memory order
    const T* w = Wu + ((k % Cb) * (K / Cb) + k / Cb) * Ca;
    float v = 0.f;
    if ((Ca & 3) == 0) {
        for (int a = 0; a < Ca; a += 4) {
            const float4 ww = ld4(w + a);
            v += ww.x * s[a] + ww.y * s[a + 1] + ww.z * s[a + 2] + ww.w * s[a + 3];
        }
    } else {
        for (int a = 0; a < Ca; ++a) v = fmaf(to_f(w[a]), s[a], v);
    }
    const long long o = (long long)row * K + k;
    if (epi == SRGAN_EPI_BIAS_ACT) {
        if (bias) { const int b = (int)(k % Cb); v += bias[bias_mod ? b % bias_mod : b]; }     // bias per large-side channel
        v = act_fwd(v, act, slope);
    } else if (href && act != SRGAN_ACT_NONE) {
        v *= act_bwd(to_f(href[o]), act, slope);
    }
    out[o] = from_f<T>(v);
}

// wgrad: a thread owns one k, a CTA 256 consecutive k and NCH rows of the batch (blockIdx.y): the NCH x Ca block of S sits in
// shared memory (broadcast reads), the L column is read 8 rows at a time (independent loads), the partial sums of the row
// chunks are combined with fp32 atomics.  (The first version walked all n rows in one dependent load -> FMA chain per thread:
// 240 us per launch at 320 rows.)
template <typename T>
__global__ void __launch_bounds__(256) skinny_wgrad_kernel(const T* __restrict__ S, const T* __restrict__ L, float* __restrict__ dW,
                                                           int n, int Ca, long long K) {
    constexpr int NCH = 32, U = 8;
    __shared__ float s[NCH][CA_MAX];
    const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
    const int n0 = blockIdx.y * NCH;
    const int nn = min(NCH, n - n0);
    for (int i = threadIdx.x; i < NCH * CA_MAX; i += 256) {
        const int r = i / CA_MAX, a = i % CA_MAX;
        s[r][a] = (r < nn && a < Ca) ? to_f(S[(long long)(n0 + r) * Ca + a]) : 0.f;
    }
    __syncthreads();
    if (k >= K) return;
    float acc[CA_MAX];
#pragma unroll
    for (int a = 0; a < CA_MAX; ++a) acc[a] = 0.f;
    for (int r0 = 0; r0 < nn; r0 += U) {
        float x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = r0 + u < nn ? to_f(L[(long long)(n0 + r0 + u) * K + k]) : 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int a = 0; a < CA_MAX; ++a) acc[a] = fmaf(s[r0 + u][a], x[u], acc[a]);
    }
#pragma unroll
    for (int a = 0; a < CA_MAX; ++a)
        if (a < Ca) atomicAdd(dW + (long long)a * K + k, acc[a]);
}

}  // namespace

// full-extent pair with few outputs?  (the large side's whole H x W x Cb extent is one reduction axis)
bool skinny_eligible(const srgan_geom* g) {
    return g->Hs == 1 && g->Ws == 1 && g->R == g->Hl && g->S == g->Wl && g->pad == 0 && g->Ca <= CA_MAX &&
           (long long)g->R * g->S * g->Cb >= 512;
}

template <typename T>
static int skinny_conv_t(int mode, const void* src, const void* W, void* out, int n, const srgan_geom* g, const float* bias, int bias_mod,
                         const void* href, int epi, int act, float slope, cudaStream_t st) {
    const long long K = (long long)g->R * g->S * g->Cb;
    if (mode == 0) {
        skinny_down_kernel<T><<<(n + ROWS - 1) / ROWS, 256, 0, st>>>((const T*)src, (const T*)W, (T*)out, bias, bias_mod, (const T*)href, epi,
                                                                     act, slope, n, g->Ca, K);
        SRGAN_CHECK_LAUNCH("skinny_down_kernel");
    } else {
        dim3 grid((unsigned)((K + 255) / 256), n);
        skinny_up_kernel<T><<<grid, 256, 0, st>>>((const T*)src, (const T*)W, (T*)out, bias, bias_mod, (const T*)href, epi, act, slope, g->Ca, g->Cb, K);
        SRGAN_CHECK_LAUNCH("skinny_up_kernel");
    }
    return SRGAN_OK;
}

int skinny_conv(int mode, const void* src, const void* W, void* out, int n, const srgan_geom* g, const float* bias, int bias_mod,
                const void* href, int epi, int act, float slope, int dtype, cudaStream_t st) {
    if (dtype == SRGAN_F32) return skinny_conv_t<float>(mode, src, W, out, n, g, bias, bias_mod, href, epi, act, slope, st);
    return skinny_conv_t<bf16>(mode, src, W, out, n, g, bias, bias_mod, href, epi, act, slope, st);
}

int skinny_wgrad(const void* S, const void* L, float* dW, int n, const srgan_geom* g, int dtype, cudaStream_t st) {
    const long long K = (long long)g->R * g->S * g->Cb;
    const dim3 grid((unsigned)((K + 255) / 256), (unsigned)((n + 31) / 32));
    if (dtype == SRGAN_F32) skinny_wgrad_kernel<float><<<grid, 256, 0, st>>>((const float*)S, (const float*)L, dW, n, g->Ca, K);
    else skinny_wgrad_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)S, (const bf16*)L, dW, n, g->Ca, K);
    SRGAN_CHECK_LAUNCH("skinny_wgrad_kernel");
    return SRGAN_OK;
}
