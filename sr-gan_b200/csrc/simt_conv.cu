// SIMT (CUDA-core FMA, fp32 accumulate) implicit-GEMM kernels for the conv pair: down / up / wgrad.
// They are (1) the fp32 parity mode of the SR-GAN step (1e-4 vs the reference, BASELINE.json north_star) and (2) the
// path for shapes the tcgen05 kernels do not take (3-channel image layers, the 10-wide coefficient MLP, odd sizes).
// Replaces cuDNN fprop/dgrad/wgrad + cuBLAS addmm behind age/models.py:44-52,68-80 and coefficient/models.py:22-72.
#include "common.cuh"

struct ConvP {
    int n, Hs, Ws, Ca, Hl, Wl, Cb, R, S, stride, pad;
};

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;

// ------------------------------------------------------------------------------------------------------------
// down / up :  out[m, c] = epilogue( sum_k A[m,k] * B[c,k] )
//   down: m=(n,oh,ow) on the small side, c=a, k=(r,s,b):  A = L[n, oh*st-pad+r, ow*st-pad+s, b],  B = Wd[a][k]
//   up  : per output phase (ih%st, iw%st) = blockIdx.z: m=(n,i,j), c=b, k=(tr,ts,a) over the taps that hit the
//         phase:  A = S[n, i+qa-tr, j+qb-ts, a],  B = Wu[b][r0+st*tr][s0+st*ts][a]
// ------------------------------------------------------------------------------------------------------------
template <typename T, int MODE, bool VEC>
__global__ void __launch_bounds__(NT) conv_gemm_kernel(const T* __restrict__ src, const T* __restrict__ W,
                                                       T* __restrict__ out, const float* __restrict__ bias, int bias_mod,
                                                       const T* __restrict__ href, int epi, int act, float slope, ConvP p) {
    __shared__ float As[2][BK][BM + 4];
    __shared__ float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x;
    // ---- problem view for this block
    int M, N, K, Hm, Wm;           // Hm x Wm: the row grid (per sample)
    int Cin;                       // channels per tap on the K axis
    int r0 = 0, s0 = 0, qa = 0, qb = 0, Rt = p.R, St = p.S, pa = 0, pb = 0;
    if (MODE == 0) {
        Hm = p.Hs; Wm = p.Ws; N = p.Ca; Cin = p.Cb; K = p.R * p.S * p.Cb;
    } else {
        pa = blockIdx.z / p.stride; pb = blockIdx.z % p.stride;
        Hm = (p.Hl - pa + p.stride - 1) / p.stride;
        Wm = (p.Wl - pb + p.stride - 1) / p.stride;
        r0 = (pa + p.pad) % p.stride; s0 = (pb + p.pad) % p.stride;
        qa = (pa + p.pad - r0) / p.stride; qb = (pb + p.pad - s0) / p.stride;
        Rt = r0 < p.R ? (p.R - r0 + p.stride - 1) / p.stride : 0;
        St = s0 < p.S ? (p.S - s0 + p.stride - 1) / p.stride : 0;
        N = p.Cb; Cin = p.Ca; K = Rt * St * p.Ca;
    }
    M = p.n * Hm * Wm;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (m0 >= M) return;

    // ---- A loader: thread -> (row = tid/2, 8 consecutive k starting at (tid%2)*8)
    const int a_row = tid >> 1, a_kq = (tid & 1) * 8;
    const int am = m0 + a_row;
    const bool a_ok = am < M;
    int an = 0, ay = 0, ax = 0;
    if (a_ok) { an = am / (Hm * Wm); int rem = am - an * Hm * Wm; ay = rem / Wm; ax = rem - ay * Wm; }
    // ---- B loader: thread -> (col = tid/4, 4 consecutive k starting at (tid%4)*4)
    const int b_col = tid >> 2, b_kq = (tid & 3) * 4;
    const int bc = n0 + b_col;
    const bool b_ok = bc < N;

    auto a_addr = [&](int k, bool& valid) -> long long {      // element address of A[am, k]; k < K assumed
        int tap = k / Cin, ch = k - tap * Cin;
        int tr = tap / St, ts = tap - tr * St;
        if (MODE == 0) {
            int ih = ay * p.stride - p.pad + tr, iw = ax * p.stride - p.pad + ts;
            valid = a_ok && ih >= 0 && ih < p.Hl && iw >= 0 && iw < p.Wl;
            return (((long long)an * p.Hl + ih) * p.Wl + iw) * p.Cb + ch;
        } else {
            int oh = ay + qa - tr, ow = ax + qb - ts;
            valid = a_ok && oh >= 0 && oh < p.Hs && ow >= 0 && ow < p.Ws;
            return (((long long)an * p.Hs + oh) * p.Ws + ow) * p.Ca + ch;
        }
    };
    auto b_addr = [&](int k) -> long long {
        if (MODE == 0) return (long long)bc * K + k;
        int tap = k / Cin, ch = k - tap * Cin;
        int tr = tap / St, ts = tap - tr * St;
        int r = r0 + p.stride * tr, s = s0 + p.stride * ts;
        return (long long)bc * (p.R * p.S * p.Ca) + (long long)(r * p.S + s) * p.Ca + ch;
    };

    float4 ra[2], rb;
    auto load_global = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int k = k0 + a_kq + h * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (VEC) {
                if (k < K) { bool ok; long long ad = a_addr(k, ok); if (ok) v = ld4(src + ad); }
            } else {
                float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (k + e < K) { bool ok; long long ad = a_addr(k + e, ok); if (ok) t[e] = to_f(src[ad]); }
                v = make_float4(t[0], t[1], t[2], t[3]);
            }
            ra[h] = v;
        }
        {
            int k = k0 + b_kq;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b_ok) {
                if (VEC) { if (k < K) v = ld4(W + b_addr(k)); }
                else {
                    float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int e = 0; e < 4; ++e) if (k + e < K) t[e] = to_f(W[b_addr(k + e)]);
                    v = make_float4(t[0], t[1], t[2], t[3]);
                }
            }
            rb = v;
        }
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int kk = a_kq + h * 4;
            As[buf][kk + 0][a_row] = ra[h].x; As[buf][kk + 1][a_row] = ra[h].y;
            As[buf][kk + 2][a_row] = ra[h].z; As[buf][kk + 3][a_row] = ra[h].w;
        }
        Bs[buf][b_kq + 0][b_col] = rb.x; Bs[buf][b_kq + 1][b_col] = rb.y;
        Bs[buf][b_kq + 2][b_col] = rb.z; Bs[buf][b_kq + 3][b_col] = rb.w;
    };

    const int tx = tid & 15, ty = tid >> 4;      // 16 x 16 threads, each 8 rows x 4 cols
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int nk = (K + BK - 1) / BK;
    if (nk > 0) { load_global(0); store_smem(0); }
    __syncthreads();
    for (int kc = 0; kc < nk; ++kc) {
        const int cur = kc & 1;
        if (kc + 1 < nk) load_global((kc + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 8 + 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kc + 1 < nk) store_smem(cur ^ 1);
        __syncthreads();
    }

    // ---- epilogue
    const int c0 = n0 + tx * 4;
    if (c0 >= N) return;
    const bool cvec = (N % 4 == 0);               // c0 % 4 == 0 always
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (epi == SRGAN_EPI_BIAS_ACT && bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (c0 + j < N) bv[j] = bias[bias_mod ? (c0 + j) % bias_mod : (c0 + j)];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m0 + ty * 8 + i;
        if (m >= M) break;
        long long o;
        if (MODE == 0) {
            o = (long long)m * N + c0;
        } else {
            int nn = m / (Hm * Wm); int rem = m - nn * Hm * Wm; int yy = rem / Wm, xx = rem - yy * Wm;
            o = (((long long)nn * p.Hl + (yy * p.stride + pa)) * p.Wl + (xx * p.stride + pb)) * p.Cb + c0;
        }
        float v[4];
        if (epi == SRGAN_EPI_BIAS_ACT) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = act_fwd(acc[i][j] + bv[j], act, slope);
        } else {
            if (href != nullptr && act != SRGAN_ACT_NONE) {
                if (cvec) {
                    float4 h = ld4(href + o);
                    v[0] = acc[i][0] * act_bwd(h.x, act, slope); v[1] = acc[i][1] * act_bwd(h.y, act, slope);
                    v[2] = acc[i][2] * act_bwd(h.z, act, slope); v[3] = acc[i][3] * act_bwd(h.w, act, slope);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        v[j] = (c0 + j < N) ? acc[i][j] * act_bwd(to_f(href[o + j]), act, slope) : 0.f;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = acc[i][j];
            }
        }
        if (cvec) st4(out + o, make_float4(v[0], v[1], v[2], v[3]));
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (c0 + j < N) out[o + j] = from_f<T>(v[j]);
        }
    }
}

template <typename T>
static int launch_conv_gemm(int mode, const T* src, const T* W, T* out, int n, const srgan_geom* g, const float* bias,
                            int bias_mod, const T* href, int epi, int act, float slope, cudaStream_t st) {
    ConvP p{n, g->Hs, g->Ws, g->Ca, g->Hl, g->Wl, g->Cb, g->R, g->S, g->stride, g->pad};
    int Cin = mode == 0 ? g->Cb : g->Ca;
    bool vec = (Cin % 4 == 0);
    dim3 grid;
    if (mode == 0) {
        long long M = (long long)n * g->Hs * g->Ws;
        grid = dim3(cdiv(M, BM), cdiv(g->Ca, BN), 1);
    } else {
        int Hm = (g->Hl + g->stride - 1) / g->stride, Wm = (g->Wl + g->stride - 1) / g->stride;
        long long M = (long long)n * Hm * Wm;
        grid = dim3(cdiv(M, BM), cdiv(g->Cb, BN), g->stride * g->stride);
    }
    if (grid.x == 0 || grid.y == 0) return SRGAN_OK;
    if (mode == 0) {
        if (vec) conv_gemm_kernel<T, 0, true><<<grid, NT, 0, st>>>(src, W, out, bias, bias_mod, href, epi, act, slope, p);
        else conv_gemm_kernel<T, 0, false><<<grid, NT, 0, st>>>(src, W, out, bias, bias_mod, href, epi, act, slope, p);
    } else {
        if (vec) conv_gemm_kernel<T, 1, true><<<grid, NT, 0, st>>>(src, W, out, bias, bias_mod, href, epi, act, slope, p);
        else conv_gemm_kernel<T, 1, false><<<grid, NT, 0, st>>>(src, W, out, bias, bias_mod, href, epi, act, slope, p);
    }
    SRGAN_CHECK_LAUNCH("conv_gemm_kernel");
    return SRGAN_OK;
}

// ------------------------------------------------------------------------------------------------------------
// direct `up` for image-side layers (Cb <= 4 output channels, e.g. the data-gradient of the crowd stem Conv2d(3, 64, k7 s2 p3),
// crowd/models.py:1075): as a GEMM the N axis would be 3 wide.  One thread per large-side pixel, all Cb channels in
// registers, weights staged in shared memory, S rows read as 4-wide vectors.
// ------------------------------------------------------------------------------------------------------------
template <typename T, int CB>
__global__ void __launch_bounds__(256) direct_up_kernel(const T* __restrict__ S, const T* __restrict__ Wu, T* __restrict__ out,
                                                        const float* __restrict__ bias, int bias_mod, const T* __restrict__ href,
                                                        int epi, int act, float slope, ConvP p) {
    extern __shared__ float wsm[];                 // [CB][R][S][Ca]
    const int wn = CB * p.R * p.S * p.Ca;
    for (int i = threadIdx.x; i < wn; i += 256) wsm[i] = to_f(Wu[i]);
    __syncthreads();
    const long long total = (long long)p.n * p.Hl * p.Wl;
    const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
    if (pix >= total) return;
    const int iw = (int)(pix % p.Wl);
    const long long t = pix / p.Wl;
    const int ih = (int)(t % p.Hl);
    const long long b = t / p.Hl;
    float acc[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) acc[c] = 0.f;
    const int tap = p.R * p.S * p.Ca;              // stride between output channels in wsm
    for (int r = (ih + p.pad) % p.stride; r < p.R; r += p.stride) {
        const int oh = (ih + p.pad - r) / p.stride;
        if (ih + p.pad - r < 0) break;
        if (oh >= p.Hs) continue;
        for (int s = (iw + p.pad) % p.stride; s < p.S; s += p.stride) {
            const int ow = (iw + p.pad - s) / p.stride;
            if (iw + p.pad - s < 0) break;
            if (ow >= p.Ws) continue;
            const T* sp = S + ((b * p.Hs + oh) * p.Ws + ow) * p.Ca;
            const float* wp = wsm + (r * p.S + s) * p.Ca;
            for (int a = 0; a < p.Ca; a += 4) {
                const float4 v = ld4(sp + a);
#pragma unroll
                for (int c = 0; c < CB; ++c) {
                    const float4 w = *reinterpret_cast<const float4*>(wp + c * tap + a);
                    acc[c] += v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w;
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < CB; ++c) {
        const long long o = pix * CB + c;
        float v = acc[c];
        if (epi == SRGAN_EPI_BIAS_ACT) {
            if (bias) v += bias[bias_mod ? c % bias_mod : c];
            v = act_fwd(v, act, slope);
        } else if (href && act != SRGAN_ACT_NONE) {
            v *= act_bwd(to_f(href[o]), act, slope);
        }
        out[o] = from_f<T>(v);
    }
}

template <typename T>
static int launch_direct_up(const T* S, const T* Wu, T* out, int n, const srgan_geom* g, const float* bias, int bias_mod,
                            const T* href, int epi, int act, float slope, cudaStream_t st) {
    ConvP p{n, g->Hs, g->Ws, g->Ca, g->Hl, g->Wl, g->Cb, g->R, g->S, g->stride, g->pad};
    const size_t smem = (size_t)g->Cb * g->R * g->S * g->Ca * sizeof(float);
    const long long total = (long long)n * g->Hl * g->Wl;
    const unsigned grid = (unsigned)((total + 255) / 256);
    switch (g->Cb) {
        case 1: direct_up_kernel<T, 1><<<grid, 256, smem, st>>>(S, Wu, out, bias, bias_mod, href, epi, act, slope, p); break;
        case 2: direct_up_kernel<T, 2><<<grid, 256, smem, st>>>(S, Wu, out, bias, bias_mod, href, epi, act, slope, p); break;
        case 3: direct_up_kernel<T, 3><<<grid, 256, smem, st>>>(S, Wu, out, bias, bias_mod, href, epi, act, slope, p); break;
        default: direct_up_kernel<T, 4><<<grid, 256, smem, st>>>(S, Wu, out, bias, bias_mod, href, epi, act, slope, p); break;
    }
    SRGAN_CHECK_LAUNCH("direct_up_kernel");
    return SRGAN_OK;
}

static bool direct_up_eligible(int mode, const srgan_geom* g, int n) {
    return mode == 1 && g->Cb <= 4 && g->Ca % 4 == 0 && (size_t)g->Cb * g->R * g->S * g->Ca * sizeof(float) <= 48 * 1024 &&
           (long long)n * g->Hl * g->Wl >= 1024;
}

// ------------------------------------------------------------------------------------------------------------
// patch convolutions: kernel = stride = 2, no padding, a handful of channels (the crowd MapModule convs 1->8->16->32,
// crowd/models.py:770-776, and their data gradients).  Every output pixel reads its own 2x2 input patch and nothing else,
// the whole filter is <= 2048 weights: one thread per small-side pixel, the patch / the Ca deltas in registers, weights
// broadcast from shared memory as float4, contiguous runs moved with the widest aligned vector access.  As tiled GEMMs
// these ran at 0.2-0.5 TB/s (N = 8..32 columns of a 64-wide tile); they are pure streaming problems.
// ------------------------------------------------------------------------------------------------------------
template <typename T, int N> struct RunVec {
    static constexpr int BYTES = N * (int)sizeof(T);
    static constexpr int VB = BYTES % 16 == 0 ? 16 : (BYTES % 8 == 0 ? 8 : (BYTES % 4 == 0 ? 4 : (int)sizeof(T)));
};
template <typename T, int N>
__device__ __forceinline__ void ld_run(const T* __restrict__ p, float (&v)[N]) {
    constexpr int VB = RunVec<T, N>::VB, E = VB / (int)sizeof(T);
    T tmp[N];
#pragma unroll
    for (int i = 0; i < N / E; ++i) {
        if (VB == 16) *reinterpret_cast<uint4*>(tmp + i * E) = *reinterpret_cast<const uint4*>(p + i * E);
        else if (VB == 8) *reinterpret_cast<uint2*>(tmp + i * E) = *reinterpret_cast<const uint2*>(p + i * E);
        else if (VB == 4) *reinterpret_cast<uint32_t*>(tmp + i * E) = *reinterpret_cast<const uint32_t*>(p + i * E);
        else tmp[i] = p[i];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = to_f(tmp[i]);
}
template <typename T, int N>
__device__ __forceinline__ void st_run(T* __restrict__ p, const float (&v)[N]) {
    constexpr int VB = RunVec<T, N>::VB, E = VB / (int)sizeof(T);
    T tmp[N];
#pragma unroll
    for (int i = 0; i < N; ++i) tmp[i] = from_f<T>(v[i]);
#pragma unroll
    for (int i = 0; i < N / E; ++i) {
        if (VB == 16) *reinterpret_cast<uint4*>(p + i * E) = *reinterpret_cast<const uint4*>(tmp + i * E);
        else if (VB == 8) *reinterpret_cast<uint2*>(p + i * E) = *reinterpret_cast<const uint2*>(tmp + i * E);
        else if (VB == 4) *reinterpret_cast<uint32_t*>(p + i * E) = *reinterpret_cast<const uint32_t*>(tmp + i * E);
        else p[i] = tmp[i];
    }
}

// down: S[pix, a] = epi(sum_{r,s,b} L[2oh+r, 2ow+s, b] * Wd[a][r][s][b])
template <typename T, int CA, int CB>
__global__ void __launch_bounds__(128) patch_down_kernel(const T* __restrict__ L, const T* __restrict__ Wd, T* __restrict__ out,
                                                         const float* __restrict__ bias, int bias_mod, const T* __restrict__ href,
                                                         int epi, int act, float slope, ConvP p) {
    constexpr int KK = 4 * CB;                     // patch elements; (r, s, b) order = two runs of 2*CB contiguous elements
    __shared__ __align__(16) float wsm[CA * KK];
    for (int i = threadIdx.x; i < CA * KK; i += blockDim.x) wsm[i] = to_f(Wd[i]);
    __syncthreads();
    const long long total = (long long)p.n * p.Hs * p.Ws;
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
        const int ow = (int)(pix % p.Ws);
        const long long t = pix / p.Ws;
        const int oh = (int)(t % p.Hs);
        const long long b = t / p.Hs;
        float in[KK];
        {
            float r0[2 * CB], r1[2 * CB];
            const T* lp = L + ((b * p.Hl + 2 * oh) * p.Wl + 2 * ow) * CB;
            ld_run<T, 2 * CB>(lp, r0);
            ld_run<T, 2 * CB>(lp + (long long)p.Wl * CB, r1);
#pragma unroll
            for (int i = 0; i < 2 * CB; ++i) { in[i] = r0[i]; in[2 * CB + i] = r1[i]; }
        }
        // outputs in groups of 8 channels (one 16-byte store in bf16): bounded register use, the group loop stays rolled
        constexpr int G = CA < 8 ? CA : 8;
#pragma unroll 1
        for (int a0 = 0; a0 < CA; a0 += G) {
            float o[G];
#pragma unroll
            for (int a = 0; a < G; ++a) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < KK; k += 4) {
                    const float4 w = *reinterpret_cast<const float4*>(wsm + (a0 + a) * KK + k);
                    acc += in[k] * w.x + in[k + 1] * w.y + in[k + 2] * w.z + in[k + 3] * w.w;
                }
                o[a] = acc;
            }
            if (epi == SRGAN_EPI_BIAS_ACT) {
#pragma unroll
                for (int a = 0; a < G; ++a) {
                    float v = o[a];
                    if (bias) v += bias[bias_mod ? (a0 + a) % bias_mod : (a0 + a)];
                    o[a] = act_fwd(v, act, slope);
                }
            } else if (href && act != SRGAN_ACT_NONE) {
                float h[G];
                ld_run<T, G>(href + pix * CA + a0, h);
#pragma unroll
                for (int a = 0; a < G; ++a) o[a] *= act_bwd(h[a], act, slope);
            }
            st_run<T, G>(out + pix * CA + a0, o);
        }
    }
}

// up: L[2oh+r, 2ow+s, b] = epi(sum_a S[pix, a] * Wu[b][r][s][a])
template <typename T, int CA, int CB>
__global__ void __launch_bounds__(128) patch_up_kernel(const T* __restrict__ S, const T* __restrict__ Wu, T* __restrict__ out,
                                                       const float* __restrict__ bias, int bias_mod, const T* __restrict__ href,
                                                       int epi, int act, float slope, ConvP p) {
    __shared__ __align__(16) float wsm[4 * CB * CA];              // [(r, s, b)][a]
    for (int i = threadIdx.x; i < 4 * CB * CA; i += blockDim.x) {
        const int a = i % CA, q = i / CA, b = q % CB, rs = q / CB;        // q = (r*2+s)*CB + b
        wsm[i] = to_f(Wu[(b * 4 + rs) * CA + a]);
    }
    __syncthreads();
    const long long total = (long long)p.n * p.Hs * p.Ws;
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
        const int ow = (int)(pix % p.Ws);
        const long long t = pix / p.Ws;
        const int oh = (int)(t % p.Hs);
        const long long b = t / p.Hs;
        float sv[CA];
        ld_run<T, CA>(S + pix * CA, sv);
        constexpr int G = 2 * CB < 8 ? 2 * CB : 8;               // outputs per group: 8 of the 2*CB-long (s, b) run
#pragma unroll 1
        for (int rq = 0; rq < 4 * CB; rq += G) {                  // rq = r * 2*CB + q0
            const int r = rq / (2 * CB), q0 = rq - r * 2 * CB;
            const long long o0 = ((b * p.Hl + 2 * oh + r) * p.Wl + 2 * ow) * CB + q0;
            float o[G];
#pragma unroll
            for (int q = 0; q < G; ++q) {
                float acc = 0.f;
#pragma unroll
                for (int a = 0; a < CA; a += 4) {
                    const float4 w = *reinterpret_cast<const float4*>(wsm + (rq + q) * CA + a);
                    acc += sv[a] * w.x + sv[a + 1] * w.y + sv[a + 2] * w.z + sv[a + 3] * w.w;
                }
                o[q] = acc;
            }
            if (epi == SRGAN_EPI_BIAS_ACT) {
#pragma unroll
                for (int q = 0; q < G; ++q) {
                    float v = o[q];
                    const int bc = (q0 + q) % CB;
                    if (bias) v += bias[bias_mod ? bc % bias_mod : bc];
                    o[q] = act_fwd(v, act, slope);
                }
            } else if (href && act != SRGAN_ACT_NONE) {
                float h[G];
                ld_run<T, G>(href + o0, h);
#pragma unroll
                for (int q = 0; q < G; ++q) o[q] *= act_bwd(h[q], act, slope);
            }
            st_run<T, G>(out + o0, o);
        }
    }
}

static bool patch_eligible(const srgan_geom* g, int n) {
    return g->R == 2 && g->S == 2 && g->stride == 2 && g->pad == 0 && g->Hl == 2 * g->Hs && g->Wl == 2 * g->Ws &&
           (long long)n * g->Hs * g->Ws >= 4096 &&
           ((g->Ca == 8 && g->Cb == 1) || (g->Ca == 16 && g->Cb == 8) || (g->Ca == 32 && g->Cb == 16));
}

template <typename T>
static int launch_patch(int mode, const T* src, const T* W, T* out, int n, const srgan_geom* g, const float* bias, int bias_mod,
                        const T* href, int epi, int act, float slope, cudaStream_t st) {
    ConvP p{n, g->Hs, g->Ws, g->Ca, g->Hl, g->Wl, g->Cb, g->R, g->S, g->stride, g->pad};
    const long long total = (long long)n * g->Hs * g->Ws;
    long long blocks = (total + 127) / 128;
    if (blocks > 16LL * kNumSMs) blocks = 16LL * kNumSMs;
    const unsigned grid = (unsigned)blocks;
#define PATCH_CASE(CA_, CB_)                                                                                               \
    if (g->Ca == CA_ && g->Cb == CB_) {                                                                                    \
        if (mode == 0) patch_down_kernel<T, CA_, CB_><<<grid, 128, 0, st>>>(src, W, out, bias, bias_mod, href, epi, act, slope, p); \
        else patch_up_kernel<T, CA_, CB_><<<grid, 128, 0, st>>>(src, W, out, bias, bias_mod, href, epi, act, slope, p);    \
    }
    PATCH_CASE(8, 1) else PATCH_CASE(16, 8) else PATCH_CASE(32, 16)
#undef PATCH_CASE
    SRGAN_CHECK_LAUNCH("patch_conv_kernel");
    return SRGAN_OK;
}

int simt_conv(int mode, const void* src, const void* W, void* out, int n, const srgan_geom* g, const float* bias,
              int bias_mod, const void* href, int epi, int act, float slope, int dtype, cudaStream_t st) {
    if (patch_eligible(g, n) && !(mode == 1 && g->Cb == 1)) {       // (8, 1) up: the few-channel direct kernel below
        if (dtype == SRGAN_F32)
            return launch_patch<float>(mode, (const float*)src, (const float*)W, (float*)out, n, g, bias, bias_mod,
                                       (const float*)href, epi, act, slope, st);
        return launch_patch<bf16>(mode, (const bf16*)src, (const bf16*)W, (bf16*)out, n, g, bias, bias_mod, (const bf16*)href,
                                  epi, act, slope, st);
    }
    if (direct_up_eligible(mode, g, n)) {
        if (dtype == SRGAN_F32)
            return launch_direct_up<float>((const float*)src, (const float*)W, (float*)out, n, g, bias, bias_mod, (const float*)href,
                                           epi, act, slope, st);
        return launch_direct_up<bf16>((const bf16*)src, (const bf16*)W, (bf16*)out, n, g, bias, bias_mod, (const bf16*)href, epi,
                                      act, slope, st);
    }
    if (dtype == SRGAN_F32)
        return launch_conv_gemm<float>(mode, (const float*)src, (const float*)W, (float*)out, n, g, bias, bias_mod,
                                       (const float*)href, epi, act, slope, st);
    return launch_conv_gemm<bf16>(mode, (const bf16*)src, (const bf16*)W, (bf16*)out, n, g, bias, bias_mod,
                                  (const bf16*)href, epi, act, slope, st);
}

// ------------------------------------------------------------------------------------------------------------
// wgrad:  dW[a, (t,b)] += sum_{pixels p=(n,oh,ow)} S[p, a] * L[n, oh*st-pad+r, ow*st-pad+s, b]
//   GEMM M = Ca, N = R*S*Cb, K = pixels; split over K across blockIdx.z, fp32 atomics into dW.
// ------------------------------------------------------------------------------------------------------------
constexpr int WM = 64, WN = 64, WK = 16;

template <typename T, bool VECA, bool VECB>
__global__ void __launch_bounds__(NT) wgrad_kernel(const T* __restrict__ Ssrc, const T* __restrict__ Lsrc,
                                                   float* __restrict__ dW, ConvP p, long long pix_per_split) {
    __shared__ float As[2][WK][WM + 4];
    __shared__ float Bs[2][WK][WN + 4];
    const int tid = threadIdx.x;
    const int M = p.Ca, N = p.R * p.S * p.Cb;
    const long long P = (long long)p.n * p.Hs * p.Ws;
    const int n0 = blockIdx.x * WN, m0 = blockIdx.y * WM;
    const long long p_begin = (long long)blockIdx.z * pix_per_split;
    long long p_end = p_begin + pix_per_split;
    if (p_end > P) p_end = P;
    if (p_begin >= p_end) return;

    const int l_pix = tid >> 4, l_q = (tid & 15) * 4;     // thread -> (pixel within chunk, 4 consecutive a / cols)
    // B column decode is k-invariant
    const int col = n0 + l_q;
    int tap = 0, cb = 0, tr = 0, ts = 0;
    if (VECB) { if (col < N) { tap = col / p.Cb; cb = col - tap * p.Cb; tr = tap / p.S; ts = tap - tr * p.S; } }

    float4 ra, rb;
    auto load_global = [&](long long pk0) {
        long long pp = pk0 + l_pix;
        ra = make_float4(0.f, 0.f, 0.f, 0.f);
        rb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pp >= p_end) return;
        int a = m0 + l_q;
        if (VECA) { if (a < M) ra = ld4(Ssrc + pp * p.Ca + a); }
        else {
            float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 4; ++e) if (a + e < M) t[e] = to_f(Ssrc[pp * p.Ca + a + e]);
            ra = make_float4(t[0], t[1], t[2], t[3]);
        }
        int nn = (int)(pp / (p.Hs * p.Ws)); int rem = (int)(pp - (long long)nn * p.Hs * p.Ws);
        int oh = rem / p.Ws, ow = rem - oh * p.Ws;
        if (VECB) {
            if (col < N) {
                int ih = oh * p.stride - p.pad + tr, iw = ow * p.stride - p.pad + ts;
                if (ih >= 0 && ih < p.Hl && iw >= 0 && iw < p.Wl)
                    rb = ld4(Lsrc + (((long long)nn * p.Hl + ih) * p.Wl + iw) * p.Cb + cb);
            }
        } else {
            float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int c = col + e;
                if (c < N) {
                    int tp = c / p.Cb, b = c - tp * p.Cb, r = tp / p.S, s = tp - r * p.S;
                    int ih = oh * p.stride - p.pad + r, iw = ow * p.stride - p.pad + s;
                    if (ih >= 0 && ih < p.Hl && iw >= 0 && iw < p.Wl)
                        t[e] = to_f(Lsrc[(((long long)nn * p.Hl + ih) * p.Wl + iw) * p.Cb + b]);
                }
            }
            rb = make_float4(t[0], t[1], t[2], t[3]);
        }
    };
    auto store_smem = [&](int buf) {
        *reinterpret_cast<float4*>(&As[buf][l_pix][l_q]) = ra;
        *reinterpret_cast<float4*>(&Bs[buf][l_pix][l_q]) = rb;
    };

    const int tx = tid & 15, ty = tid >> 4;       // each thread 4 (a) x 4 (cols)
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const long long nk = (p_end - p_begin + WK - 1) / WK;
    load_global(p_begin); store_smem(0);
    __syncthreads();
    for (long long kc = 0; kc < nk; ++kc) {
        const int cur = (int)(kc & 1);
        if (kc + 1 < nk) load_global(p_begin + (kc + 1) * WK);
#pragma unroll
        for (int k = 0; k < WK; ++k) {
            float4 a = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kc + 1 < nk) store_smem(cur ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int a = m0 + ty * 4 + i;
        if (a >= M) break;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = n0 + tx * 4 + j;
            if (c < N) atomicAdd(dW + (long long)a * N + c, acc[i][j]);
        }
    }
}

template <typename T>
static int launch_wgrad(const T* S, const T* L, float* dW, int n, const srgan_geom* g, cudaStream_t st) {
    ConvP p{n, g->Hs, g->Ws, g->Ca, g->Hl, g->Wl, g->Cb, g->R, g->S, g->stride, g->pad};
    const int M = g->Ca, N = g->R * g->S * g->Cb;
    const long long P = (long long)n * g->Hs * g->Ws;
    if (P == 0 || M == 0 || N == 0) return SRGAN_OK;
    int tiles = cdiv(N, WN) * cdiv(M, WM);
    long long want = (4LL * kNumSMs + tiles - 1) / tiles;          // ~4 CTAs per SM in total
    long long max_splits = (P + 8 * WK - 1) / (8 * WK);            // at least 8 k-chunks per split
    long long splits = want < max_splits ? want : max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    long long pps = (P + splits - 1) / splits;
    pps = (pps + WK - 1) / WK * WK;
    splits = (P + pps - 1) / pps;
    dim3 grid(cdiv(N, WN), cdiv(M, WM), (unsigned)splits);
    bool va = (g->Ca % 4 == 0), vb = (g->Cb % 4 == 0);
    if (va && vb) wgrad_kernel<T, true, true><<<grid, NT, 0, st>>>(S, L, dW, p, pps);
    else if (va) wgrad_kernel<T, true, false><<<grid, NT, 0, st>>>(S, L, dW, p, pps);
    else if (vb) wgrad_kernel<T, false, true><<<grid, NT, 0, st>>>(S, L, dW, p, pps);
    else wgrad_kernel<T, false, false><<<grid, NT, 0, st>>>(S, L, dW, p, pps);
    SRGAN_CHECK_LAUNCH("wgrad_kernel");
    return SRGAN_OK;
}

// ------------------------------------------------------------------------------------------------------------
// direct wgrad for tiny filters (Ca * R * S * Cb <= 64 weights, e.g. MapModule.conv1 = Conv2d(1, 8, k2 s2), crowd/models.py:770):
// as a GEMM the output is 8 x 4 and the tiled kernel wastes 99 % of its work.  One thread per small-side pixel (grid-stride),
// all weights' partial sums in registers, one warp-shuffle reduction and 32 atomics per warp at the end.
// ------------------------------------------------------------------------------------------------------------
template <typename T, int CA, int R, int S, int CB>
__global__ void __launch_bounds__(256) direct_wgrad_small_kernel(const T* __restrict__ Ssrc, const T* __restrict__ Lsrc,
                                                                 float* __restrict__ dW, ConvP p) {
    constexpr int TT = R * S * CB, NW = CA * TT;
    float acc[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) acc[i] = 0.f;
    const long long total = (long long)p.n * p.Hs * p.Ws;
    for (long long pix = (long long)blockIdx.x * 256 + threadIdx.x; pix < total; pix += 256LL * gridDim.x) {
        const int ow = (int)(pix % p.Ws);
        const long long t = pix / p.Ws;
        const int oh = (int)(t % p.Hs);
        const long long b = t / p.Hs;
        float sv[CA], lv[TT];
#pragma unroll
        for (int a = 0; a < CA; ++a) sv[a] = to_f(Ssrc[pix * CA + a]);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int q = 0; q < S; ++q) {
                const int ih = oh * p.stride - p.pad + r, iw = ow * p.stride - p.pad + q;
                const bool ok = ih >= 0 && ih < p.Hl && iw >= 0 && iw < p.Wl;
#pragma unroll
                for (int c = 0; c < CB; ++c)
                    lv[(r * S + q) * CB + c] = ok ? to_f(Lsrc[((b * p.Hl + ih) * p.Wl + iw) * CB + c]) : 0.f;
            }
#pragma unroll
        for (int a = 0; a < CA; ++a)
#pragma unroll
            for (int k = 0; k < TT; ++k) acc[a * TT + k] = fmaf(sv[a], lv[k], acc[a * TT + k]);
    }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        const float v = warp_sum(acc[i]);
        if (lane == (i & 31)) atomicAdd(dW + i, v);          // dW is [a][r][s][b] = [a][k]
    }
}

// ------------------------------------------------------------------------------------------------------------
// wgrad of the patch convolutions (kernel = stride = 2, pad 0; MapModule conv2 / conv3):
//   dW[a][(r,s,b)] += sum_pix S[pix, a] * L[patch(pix), (r,s,b)]          CA x 4*CB outputs (512 / 2048), K = pixels.
// A CTA walks its pixel range in chunks of 64 staged in shared memory as fp32 (the next chunk's global loads are in
// flight while the current one is multiplied); thread (aq, kq) owns a TA x TK block of outputs; fp32 atomics at the end.
// ------------------------------------------------------------------------------------------------------------
template <typename T, int CA, int CB>
__global__ void __launch_bounds__(256) patch_wgrad_kernel(const T* __restrict__ Ssrc, const T* __restrict__ Lsrc,
                                                          float* __restrict__ dW, ConvP p, long long pix_per_block) {
    constexpr int KK = 4 * CB, PC = 64;
    constexpr int TK = KK / 16, TA = CA / 16;                     // 16 x 16 threads over (a, k)
    static_assert(TK >= 1 && TA >= 1 && TK * 16 == KK && TA * 16 == CA, "patch_wgrad tile");
    __shared__ __align__(16) float Ss[PC][CA];
    __shared__ __align__(16) float Ls[PC][KK];
    const long long total = (long long)p.n * p.Hs * p.Ws;
    const long long p0 = (long long)blockIdx.x * pix_per_block;
    const long long p1 = p0 + pix_per_block < total ? p0 + pix_per_block : total;
    const int t = threadIdx.x;
    const int kq = t & 15, aq = t >> 4;
    float acc[TA][TK];
#pragma unroll
    for (int i = 0; i < TA; ++i)
#pragma unroll
        for (int j = 0; j < TK; ++j) acc[i][j] = 0.f;
    // loader roles: threads 0..127 = one (pixel, patch row r) run of 2*CB elements of L; threads 128..191 = one S row
    float lrun[2 * CB], srow[CA];
    auto fetch = [&](long long base) {
        if (t < 2 * PC) {
            const long long pix = base + (t >> 1);
#pragma unroll
            for (int i = 0; i < 2 * CB; ++i) lrun[i] = 0.f;
            if (pix < p1) {
                const int ow = (int)(pix % p.Ws);
                const long long q = pix / p.Ws;
                const int oh = (int)(q % p.Hs);
                const long long b = q / p.Hs;
                ld_run<T, 2 * CB>(Lsrc + ((b * p.Hl + 2 * oh + (t & 1)) * p.Wl + 2 * ow) * CB, lrun);
            }
        } else if (t < 3 * PC) {
            const long long pix = base + (t - 2 * PC);
#pragma unroll
            for (int i = 0; i < CA; ++i) srow[i] = 0.f;
            if (pix < p1) ld_run<T, CA>(Ssrc + pix * CA, srow);
        }
    };
    if (p0 < p1) fetch(p0);
    for (long long base = p0; base < p1; base += PC) {
        if (t < 2 * PC) {
#pragma unroll
            for (int i = 0; i < 2 * CB; i += 4)
                *reinterpret_cast<float4*>(&Ls[t >> 1][(t & 1) * 2 * CB + i]) = make_float4(lrun[i], lrun[(i + 1) % (2 * CB)], lrun[(i + 2) % (2 * CB)], lrun[(i + 3) % (2 * CB)]);
        } else if (t < 3 * PC) {
#pragma unroll
            for (int i = 0; i < CA; i += 4)
                *reinterpret_cast<float4*>(&Ss[t - 2 * PC][i]) = make_float4(srow[i], srow[i + 1], srow[i + 2], srow[i + 3]);
        }
        __syncthreads();
        if (base + PC < p1) fetch(base + PC);
#pragma unroll 8
        for (int pp = 0; pp < PC; ++pp) {
            float sv[TA], lv[TK];
#pragma unroll
            for (int i = 0; i < TA; ++i) sv[i] = Ss[pp][aq * TA + i];
#pragma unroll
            for (int j = 0; j < TK; ++j) lv[j] = Ls[pp][kq * TK + j];
#pragma unroll
            for (int i = 0; i < TA; ++i)
#pragma unroll
                for (int j = 0; j < TK; ++j) acc[i][j] = fmaf(sv[i], lv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TA; ++i)
#pragma unroll
        for (int j = 0; j < TK; ++j) atomicAdd(dW + (aq * TA + i) * KK + kq * TK + j, acc[i][j]);
}

template <typename T>
static int launch_patch_wgrad(const T* S, const T* L, float* dW, int n, const srgan_geom* g, cudaStream_t st) {
    ConvP p{n, g->Hs, g->Ws, g->Ca, g->Hl, g->Wl, g->Cb, g->R, g->S, g->stride, g->pad};
    const long long total = (long long)n * g->Hs * g->Ws;
    long long blocks = 4LL * kNumSMs;
    long long ppb = (total + blocks - 1) / blocks;
    ppb = (ppb + 63) / 64 * 64;
    blocks = (total + ppb - 1) / ppb;
    if (g->Ca == 16 && g->Cb == 8) patch_wgrad_kernel<T, 16, 8><<<(unsigned)blocks, 256, 0, st>>>(S, L, dW, p, ppb);
    else patch_wgrad_kernel<T, 32, 16><<<(unsigned)blocks, 256, 0, st>>>(S, L, dW, p, ppb);
    SRGAN_CHECK_LAUNCH("patch_wgrad_kernel");
    return SRGAN_OK;
}

int simt_wgrad(const void* S, const void* L, float* dW, int n, const srgan_geom* g, int dtype, cudaStream_t st) {
    if (g->Ca == 8 && g->R == 2 && g->S == 2 && g->Cb == 1 && (long long)n * g->Hs * g->Ws >= 512) {
        ConvP p{n, g->Hs, g->Ws, g->Ca, g->Hl, g->Wl, g->Cb, g->R, g->S, g->stride, g->pad};
        const int grid = 4 * kNumSMs;
        if (dtype == SRGAN_F32) direct_wgrad_small_kernel<float, 8, 2, 2, 1><<<grid, 256, 0, st>>>((const float*)S, (const float*)L, dW, p);
        else direct_wgrad_small_kernel<bf16, 8, 2, 2, 1><<<grid, 256, 0, st>>>((const bf16*)S, (const bf16*)L, dW, p);
        SRGAN_CHECK_LAUNCH("direct_wgrad_small_kernel");
        return SRGAN_OK;
    }
    if (patch_eligible(g, n) && g->Cb >= 8) {
        if (dtype == SRGAN_F32) return launch_patch_wgrad<float>((const float*)S, (const float*)L, dW, n, g, st);
        return launch_patch_wgrad<bf16>((const bf16*)S, (const bf16*)L, dW, n, g, st);
    }
    if (dtype == SRGAN_F32) return launch_wgrad<float>((const float*)S, (const float*)L, dW, n, g, st);
    return launch_wgrad<bf16>((const bf16*)S, (const bf16*)L, dW, n, g, st);
}
