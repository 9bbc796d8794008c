// Bandwidth-bound kernels of the SR-GAN step: feature column sums, distance losses, seeds, interpolation, gradient-norm
// penalty, layout conversion, fused Adam.  All are coalesced / 128-bit vectorised where the shape allows, reduce with
// warp shuffles, and size their grids from the SM count (148).  Reference lines: see include/srgan_b200.h.
#include "common.cuh"

// ------------------------------------------------------------------------------------------------------------
// colsum: out[c % mod] += sum_r rowscale[r] * X[r, c]
// VEC (cols % 4 == 0): a lane owns 4 consecutive columns; a warp covers cvp = min(32, pow2 >= cols/4) column groups and
// 32/cvp rows per sweep (narrow matrices fold rows into the spare lanes), 8 warps per block, 4 independent row sweeps in
// flight per thread; grid.y splits the rows.  Narrow (cols <= 8, any alignment): one thread per row, the row's elements in
// registers (consecutive threads read consecutive rows: coalesced).  Otherwise: one column per lane.
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ X, long long rows, int cols,
                                                         float* __restrict__ out, int mod,
                                                         const float* __restrict__ rowscale, int lg) {
    __shared__ float red[8][32][4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int cvp = 1 << lg, rpw = 32 >> lg;                       // column groups per warp, rows per warp sweep
    const int c = (blockIdx.x * 32 + (lane & (cvp - 1))) * 4;
    const int rsub = lane >> lg;
    const long long rstep = (long long)gridDim.y * 8 * rpw;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    if (c < cols) {
        long long r = ((long long)blockIdx.y * 8 + w) * rpw + rsub;
        for (; r + 3 * rstep < rows; r += 4 * rstep) {
            const float4 v0 = ld4(X + r * cols + c), v1 = ld4(X + (r + rstep) * cols + c);
            const float4 v2 = ld4(X + (r + 2 * rstep) * cols + c), v3 = ld4(X + (r + 3 * rstep) * cols + c);
            float s0 = 1.f, s1 = 1.f, s2 = 1.f, s3 = 1.f;
            if (rowscale) { s0 = rowscale[r]; s1 = rowscale[r + rstep]; s2 = rowscale[r + 2 * rstep]; s3 = rowscale[r + 3 * rstep]; }
            a0.x = fmaf(s0, v0.x, a0.x); a0.y = fmaf(s0, v0.y, a0.y); a0.z = fmaf(s0, v0.z, a0.z); a0.w = fmaf(s0, v0.w, a0.w);
            a1.x = fmaf(s1, v1.x, a1.x); a1.y = fmaf(s1, v1.y, a1.y); a1.z = fmaf(s1, v1.z, a1.z); a1.w = fmaf(s1, v1.w, a1.w);
            a2.x = fmaf(s2, v2.x, a2.x); a2.y = fmaf(s2, v2.y, a2.y); a2.z = fmaf(s2, v2.z, a2.z); a2.w = fmaf(s2, v2.w, a2.w);
            a3.x = fmaf(s3, v3.x, a3.x); a3.y = fmaf(s3, v3.y, a3.y); a3.z = fmaf(s3, v3.z, a3.z); a3.w = fmaf(s3, v3.w, a3.w);
        }
        for (; r < rows; r += rstep) {
            const float4 v = ld4(X + r * cols + c);
            const float sc = rowscale ? rowscale[r] : 1.f;
            a0.x = fmaf(sc, v.x, a0.x); a0.y = fmaf(sc, v.y, a0.y); a0.z = fmaf(sc, v.z, a0.z); a0.w = fmaf(sc, v.w, a0.w);
        }
    }
    red[w][lane][0] = a0.x + a1.x + a2.x + a3.x; red[w][lane][1] = a0.y + a1.y + a2.y + a3.y;
    red[w][lane][2] = a0.z + a1.z + a2.z + a3.z; red[w][lane][3] = a0.w + a1.w + a2.w + a3.w;
    __syncthreads();
    // threads 0 .. cvp*4-1: column (group g, element j) summed over the 8 warps and the rpw row-lanes
    if ((int)threadIdx.x < cvp * 4) {
        const int g = threadIdx.x >> 2, j = threadIdx.x & 3;
        const int cc = (blockIdx.x * 32 + g) * 4 + j;
        if (cc < cols) {
            float sum = 0.f;
            for (int y = 0; y < 8; ++y)
                for (int q = 0; q < rpw; ++q) sum += red[y][g + q * cvp][j];
            atomicAdd(out + (mod ? cc % mod : cc), sum);
        }
    }
}

template <typename T, int COLS>
__global__ void __launch_bounds__(256) colsum_narrow_kernel(const T* __restrict__ X, long long rows, float* __restrict__ out,
                                                            int mod, const float* __restrict__ rowscale) {
    __shared__ float red[32];
    float acc[COLS];
#pragma unroll
    for (int j = 0; j < COLS; ++j) acc[j] = 0.f;
    for (long long r = (long long)blockIdx.x * 256 + threadIdx.x; r < rows; r += 256LL * gridDim.x) {
        const float sc = rowscale ? rowscale[r] : 1.f;
#pragma unroll
        for (int j = 0; j < COLS; ++j) acc[j] = fmaf(sc, to_f(X[r * COLS + j]), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < COLS; ++j) {
        const float v = block_sum(acc[j], red);
        if (threadIdx.x == 0) atomicAdd(out + (mod ? j % mod : j), v);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ X, long long rows, int cols,
                                                     float* __restrict__ out, int mod,
                                                     const float* __restrict__ rowscale) {
    __shared__ float red[8][33];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lx;
    float acc = 0.f;
    if (c < cols)
        for (long long r = (long long)blockIdx.y * 8 + ly; r < rows; r += 8LL * gridDim.y)
            acc = fmaf(rowscale ? rowscale[r] : 1.f, to_f(X[r * cols + c]), acc);
    red[ly][lx] = acc;
    __syncthreads();
    if (ly == 0 && c < cols) {
        float sum = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) sum += red[y][lx];
        atomicAdd(out + (mod ? c % mod : c), sum);
    }
}

template <typename T>
static int launch_colsum(const T* X, long long rows, int cols, float* out, int mod, const float* rowscale,
                         cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return SRGAN_OK;
    const bool vec = (cols % 4 == 0);
    if (!vec && cols <= 8 && rows >= 4096) {
        long long b = (rows + 256 * 8 - 1) / (256 * 8);
        const int grid = (int)(b > 8LL * kNumSMs ? 8LL * kNumSMs : b);
        switch (cols) {
            case 1: colsum_narrow_kernel<T, 1><<<grid, 256, 0, st>>>(X, rows, out, mod, rowscale); break;
            case 2: colsum_narrow_kernel<T, 2><<<grid, 256, 0, st>>>(X, rows, out, mod, rowscale); break;
            case 3: colsum_narrow_kernel<T, 3><<<grid, 256, 0, st>>>(X, rows, out, mod, rowscale); break;
            case 5: colsum_narrow_kernel<T, 5><<<grid, 256, 0, st>>>(X, rows, out, mod, rowscale); break;
            case 6: colsum_narrow_kernel<T, 6><<<grid, 256, 0, st>>>(X, rows, out, mod, rowscale); break;
            default: colsum_narrow_kernel<T, 7><<<grid, 256, 0, st>>>(X, rows, out, mod, rowscale); break;
        }
        SRGAN_CHECK_LAUNCH("colsum_narrow_kernel");
        return SRGAN_OK;
    }
    if (vec) {
        int lg = 0;
        while (lg < 5 && (1 << lg) < cols / 4) ++lg;
        const int rows_per_sweep = 8 * (32 >> lg);
        const int gx = cdiv(cols, 128);
        long long want = (8LL * kNumSMs + gx - 1) / gx;
        long long maxy = (rows + 4LL * rows_per_sweep - 1) / (4LL * rows_per_sweep);      // >= 4 sweeps per block
        long long gy = want < maxy ? want : maxy;
        if (gy < 1) gy = 1;
        if (gy > 65535) gy = 65535;
        colsum_vec_kernel<T><<<dim3(gx, (unsigned)gy), 256, 0, st>>>(X, rows, cols, out, mod, rowscale, lg);
        SRGAN_CHECK_LAUNCH("colsum_vec_kernel");
        return SRGAN_OK;
    }
    const int gx = cdiv(cols, 32);
    long long want = (4LL * kNumSMs + gx - 1) / gx;
    long long maxy = (rows + 31) / 32;                 // >= 4 rows per row-lane
    long long gy = want < maxy ? want : maxy;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    colsum_kernel<T><<<dim3(gx, (unsigned)gy), 256, 0, st>>>(X, rows, cols, out, mod, rowscale);
    SRGAN_CHECK_LAUNCH("colsum_kernel");
    return SRGAN_OK;
}

// ------------------------------------------------------------------------------------------------------------
// rowdot: out[r] = sum_c X[r,c]*w[c] + bias[idx]   (TPR threads per row)
// ------------------------------------------------------------------------------------------------------------
template <typename T, int TPR>
__global__ void __launch_bounds__(256) rowdot_kernel(const T* __restrict__ X, int rows, int cols,
                                                     const float* __restrict__ w, const float* __restrict__ bias,
                                                     int bias_index, float* __restrict__ out) {
    __shared__ float red[32];
    constexpr int RPB = 256 / TPR;
    const int r = blockIdx.x * RPB + threadIdx.x / TPR;
    const int t = threadIdx.x % TPR;
    float acc = 0.f;
    if (r < rows) {
        const T* x = X + (long long)r * cols;
        if (cols % 4 == 0) {
            for (int c = t * 4; c < cols; c += TPR * 4) {
                float4 v = ld4(x + c);
                float4 ww = *reinterpret_cast<const float4*>(w + c);
                acc += v.x * ww.x + v.y * ww.y + v.z * ww.z + v.w * ww.w;
            }
        } else {
            for (int c = t; c < cols; c += TPR) acc = fmaf(to_f(x[c]), w[c], acc);
        }
    }
    if (TPR == 32) {
        acc = warp_sum(acc);
        if (t == 0 && r < rows) out[r] = acc + bias[bias_index];
    } else {
        acc = block_sum(acc, red);
        if (t == 0 && r < rows) out[r] = acc + bias[bias_index];
    }
}

// ------------------------------------------------------------------------------------------------------------
// seed_rows: out[r,c] = (gvec[c] + rowscale[r]*wrow[c]) * act'(href[r,c])
// ------------------------------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) seed_rows_kernel(T* __restrict__ out, int rows, int cols,
                                                        const float* __restrict__ gvec,
                                                        const float* __restrict__ rowscale,
                                                        const float* __restrict__ wrow, const T* __restrict__ href,
                                                        int act, float slope) {
    constexpr int V = VEC ? 4 : 1;
    const long long total = (long long)rows * cols / V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long e = i * V;
        int r = (int)(e / cols), c = (int)(e - (long long)r * cols);
        float rs = rowscale ? rowscale[r] : 0.f;
        if (VEC) {
            float4 h = ld4(href + e);
            float4 g = gvec ? *reinterpret_cast<const float4*>(gvec + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (rowscale) {
                float4 w = *reinterpret_cast<const float4*>(wrow + c);
                g.x = fmaf(rs, w.x, g.x); g.y = fmaf(rs, w.y, g.y); g.z = fmaf(rs, w.z, g.z); g.w = fmaf(rs, w.w, g.w);
            }
            g.x *= act_bwd(h.x, act, slope); g.y *= act_bwd(h.y, act, slope);
            g.z *= act_bwd(h.z, act, slope); g.w *= act_bwd(h.w, act, slope);
            st4(out + e, g);
        } else {
            float g = gvec ? gvec[c] : 0.f;
            if (rowscale) g = fmaf(rs, wrow[c], g);
            out[e] = from_f<T>(g * act_bwd(to_f(href[e]), act, slope));
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// layout conversion
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int n,
                                                           int c, long long hw) {
    // one thread per (sample, pixel): reads are coalesced per channel plane, writes are c consecutive elements
    const long long total = (long long)n * hw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long nn = i / hw, px = i - nn * hw;
        const float* s = src + nn * c * hw + px;
        T* d = dst + i * c;
        for (int k = 0; k < c; ++k) d[k] = from_f<T>(s[(long long)k * hw]);
    }
}
template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int n,
                                                           int c, long long hw) {
    const long long total = (long long)n * hw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long nn = i / hw, px = i - nn * hw;
        float* d = dst + nn * c * hw + px;
        const T* s = src + i * c;
        for (int k = 0; k < c; ++k) d[(long long)k * hw] = to_f(s[k]);
    }
}
template <typename T>
__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ src, T* __restrict__ dst, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x)
        dst[i] = from_f<T>(src[i]);
}
template <typename T>
__global__ void __launch_bounds__(256) uncast_kernel(const T* __restrict__ src, float* __restrict__ dst, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x)
        dst[i] = to_f(src[i]);
}

// ------------------------------------------------------------------------------------------------------------
// im2col / col2im for thin (<= 4 channel) image layers.  One thread per (small-side pixel, 8-element chunk of the
// Kpad-wide row): eight 2-byte gathers (L1/L2 resident: the image is tiny next to the col buffer) and one 16-byte
// coalesced store; col2im is one thread per large-side pixel gathering its <= ceil(R/st)*ceil(S/st) taps.
// ------------------------------------------------------------------------------------------------------------
struct ThinP { int n, Hs, Ws, Hl, Wl, Cb, R, S, stride, pad, kpad; };

// The column index k -> (filter row r, filter column s, channel b) split costs three integer divisions; it is the same for
// every pixel, so each block tabulates it once in shared memory (Kpad <= 256: the crowd stem k7 s2, 3 channels, has 147 of
// 192 columns) and the element loop is one table read, two adds and the 2-byte gather (interior pixels skip the bounds
// checks).
constexpr int IM2COL_TAB = 256;
template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const T* __restrict__ L, T* __restrict__ col, ThinP p) {
    __shared__ int tab[IM2COL_TAB];                              // (r << 20) | (s << 10) | b, or -1 for the pad columns
    const int K = p.R * p.S * p.Cb;
    const bool tabbed = p.kpad <= IM2COL_TAB;
    if (tabbed) {
        for (int k = threadIdx.x; k < p.kpad; k += blockDim.x) {
            int v = -1;
            if (k < K) { const int tap = k / p.Cb, b = k - tap * p.Cb, r = tap / p.S, sx = tap - r * p.S; v = (r << 20) | (sx << 10) | b; }
            tab[k] = v;
        }
        __syncthreads();
    }
    const int cpr = p.kpad / 8;                                  // 8-element chunks per row
    const long long total = (long long)p.n * p.Hs * p.Ws * cpr;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long pix = i / cpr;
        int j = (int)(i - pix * cpr);
        int nn = (int)(pix / (p.Hs * p.Ws)); int rem = (int)(pix - (long long)nn * p.Hs * p.Ws);
        int oh = rem / p.Ws, ow = rem - oh * p.Ws;
        const int ih0 = oh * p.stride - p.pad, iw0 = ow * p.stride - p.pad;
        float v[8];
        if (tabbed) {
            const long long base = (((long long)nn * p.Hl + ih0) * p.Wl + iw0) * p.Cb;   // may be negative: only in-bounds offsets are read
            const bool interior = ih0 >= 0 && iw0 >= 0 && ih0 + p.R <= p.Hl && iw0 + p.S <= p.Wl;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int t = tab[j * 8 + e];
                float x = 0.f;
                if (t >= 0) {
                    const int r = t >> 20, sx = (t >> 10) & 1023, b = t & 1023;
                    if (interior || (ih0 + r >= 0 && ih0 + r < p.Hl && iw0 + sx >= 0 && iw0 + sx < p.Wl))
                        x = to_f(L[base + (r * p.Wl + sx) * p.Cb + b]);
                }
                v[e] = x;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                int k = j * 8 + e;
                float x = 0.f;
                if (k < K) {
                    int tap = k / p.Cb, b = k - tap * p.Cb, r = tap / p.S, sx = tap - r * p.S;
                    int ih = ih0 + r, iw = iw0 + sx;
                    if (ih >= 0 && ih < p.Hl && iw >= 0 && iw < p.Wl)
                        x = to_f(L[(((long long)nn * p.Hl + ih) * p.Wl + iw) * p.Cb + b]);
                }
                v[e] = x;
            }
        }
        T* d = col + pix * p.kpad + j * 8;
        st4(d, make_float4(v[0], v[1], v[2], v[3]));
        st4(d + 4, make_float4(v[4], v[5], v[6], v[7]));
    }
}

// The DCGAN image layers (3 channels, k4 s2 p1, kpad 64), bf16: one thread per (pixel, filter row r).  The 4 taps x 3
// channels of a filter row are 24 contiguous bytes of the NHWC image, copied as twelve 2-byte loads (the source is only
// 2-byte aligned) packed into three 8-byte stores; no integer division per element.  Lanes 4p..4p+3 write one 128-byte
// col row (96 B data + 32 B zero pad), so a warp emits eight full lines.
__global__ void __launch_bounds__(256) im2col_c3k4s2_kernel(const bf16* __restrict__ L, bf16* __restrict__ col, int n, int Hs,
                                                            int Ws, int Hl, int Wl) {
    const long long total = (long long)n * Hs * Ws * 4;
    const unsigned short* Lu = reinterpret_cast<const unsigned short*>(L);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long pix = i >> 2;
        const int r = (int)(i & 3);
        const int ow = (int)(pix % Ws);
        const long long t = pix / Ws;
        const int oh = (int)(t % Hs);
        const int nn = (int)(t / Hs);
        const int ih = oh * 2 - 1 + r, iw0 = ow * 2 - 1;
        unsigned short e[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) e[k] = 0;
        if (ih >= 0 && ih < Hl) {
            const unsigned short* src = Lu + (((long long)nn * Hl + ih) * Wl + iw0) * 3;
            if (iw0 >= 0 && iw0 + 3 < Wl) {
#pragma unroll
                for (int k = 0; k < 12; ++k) e[k] = __ldg(src + k);
            } else {
#pragma unroll
                for (int s = 0; s < 4; ++s)
                    if (iw0 + s >= 0 && iw0 + s < Wl) {
                        e[s * 3] = __ldg(src + s * 3); e[s * 3 + 1] = __ldg(src + s * 3 + 1); e[s * 3 + 2] = __ldg(src + s * 3 + 2);
                    }
            }
        }
        uint2* d = reinterpret_cast<uint2*>(col + pix * 64 + r * 12);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            uint2 w;
            w.x = (unsigned)e[q * 4] | ((unsigned)e[q * 4 + 1] << 16);
            w.y = (unsigned)e[q * 4 + 2] | ((unsigned)e[q * 4 + 3] << 16);
            d[q] = w;
        }
        if (r == 3) {
            uint4* z = reinterpret_cast<uint4*>(col + pix * 64 + 48);
            z[0] = make_uint4(0u, 0u, 0u, 0u);
            z[1] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) col2im_kernel(const T* __restrict__ col, T* __restrict__ out, ThinP p,
                                                     const float* __restrict__ bias, const T* __restrict__ href, int epi,
                                                     int act, float slope) {
    const long long total = (long long)p.n * p.Hl * p.Wl;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int nn = (int)(i / (p.Hl * p.Wl)); int rem = (int)(i - (long long)nn * p.Hl * p.Wl);
        int ih = rem / p.Wl, iw = rem - ih * p.Wl;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        // taps r with (ih + pad - r) % stride == 0
        for (int r = (ih + p.pad) % p.stride; r < p.R; r += p.stride) {
            int oh = (ih + p.pad - r) / p.stride;
            if (oh < 0 || oh >= p.Hs) continue;
            for (int s = (iw + p.pad) % p.stride; s < p.S; s += p.stride) {
                int ow = (iw + p.pad - s) / p.stride;
                if (ow < 0 || ow >= p.Ws) continue;
                const T* c = col + (((long long)nn * p.Hs + oh) * p.Ws + ow) * p.kpad + (r * p.S + s) * p.Cb;
                for (int b = 0; b < p.Cb; ++b) acc[b] += to_f(c[b]);
            }
        }
        T* o = out + i * p.Cb;
        for (int b = 0; b < p.Cb; ++b) {
            float v = acc[b];
            if (epi == SRGAN_EPI_BIAS_ACT) v = act_fwd(v + (bias ? bias[b] : 0.f), act, slope);
            else if (href != nullptr && act != SRGAN_ACT_NONE) v *= act_bwd(to_f(href[i * p.Cb + b]), act, slope);
            o[b] = from_f<T>(v);
        }
    }
}

// 3 channels, k4 s2 p1, kpad 64, bf16: one thread per large-side pixel, exactly 2 x 2 taps, no divisions
__global__ void __launch_bounds__(256) col2im_c3k4s2_kernel(const bf16* __restrict__ col, bf16* __restrict__ out, int n, int Hs,
                                                            int Ws, const float* __restrict__ bias,
                                                            const bf16* __restrict__ href, int epi, int act, float slope) {
    const int Hl = 2 * Hs, Wl = 2 * Ws;
    const long long total = (long long)n * Hl * Wl;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int iw = (int)(i % Wl);
        const long long t = i / Wl;
        const int ih = (int)(t % Hl);
        const int nn = (int)(t / Hl);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        const int r0 = (ih + 1) & 1, s0 = (iw + 1) & 1;
#pragma unroll
        for (int dr = 0; dr < 2; ++dr) {
            const int r = r0 + 2 * dr;
            const int oh = (ih + 1 - r) >> 1;
            if (oh < 0 || oh >= Hs) continue;
#pragma unroll
            for (int ds = 0; ds < 2; ++ds) {
                const int sx = s0 + 2 * ds;
                const int ow = (iw + 1 - sx) >> 1;
                if (ow < 0 || ow >= Ws) continue;
                const bf16* c = col + (((long long)nn * Hs + oh) * Ws + ow) * 64 + (r * 4 + sx) * 3;
                a0 += to_f(c[0]); a1 += to_f(c[1]); a2 += to_f(c[2]);
            }
        }
        float v[3] = {a0, a1, a2};
        if (epi == SRGAN_EPI_BIAS_ACT) {
#pragma unroll
            for (int b = 0; b < 3; ++b) v[b] = act_fwd(v[b] + (bias ? bias[b] : 0.f), act, slope);
        } else if (href != nullptr && act != SRGAN_ACT_NONE) {
#pragma unroll
            for (int b = 0; b < 3; ++b) v[b] *= act_bwd(to_f(href[i * 3 + b]), act, slope);
        }
        bf16* o = out + i * 3;
        o[0] = from_f<bf16>(v[0]); o[1] = from_f<bf16>(v[1]); o[2] = from_f<bf16>(v[2]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// interpolate: out[n,e] = alpha[n]*u[n,e] + (1-alpha[n])*fake[n,e]
// ------------------------------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) interpolate_kernel(const T* __restrict__ u, const T* __restrict__ f,
                                                          const float* __restrict__ alpha, T* __restrict__ out,
                                                          long long total, long long per_sample) {
    constexpr int V = VEC ? 4 : 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total / V;
         i += (long long)gridDim.x * blockDim.x) {
        long long e = i * V;
        float a = alpha[e / per_sample];
        if (VEC) {
            float4 x = ld4(u + e), y = ld4(f + e);
            float b = 1.f - a;
            st4(out + e, make_float4(a * x.x + b * y.x, a * x.y + b * y.y, a * x.z + b * y.z, a * x.w + b * y.w));
        } else {
            out[e] = from_f<T>(a * to_f(u[e]) + (1.f - a) * to_f(f[e]));
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// scalar losses (single block)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) labeled_loss_kernel(const float* __restrict__ pred, const float* __restrict__ y,
                                                           int n, int order, float scale, float* loss, float* dpred) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float d = pred[i] - y[i];
        float ad = fabsf(d);
        float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        float pw, dpw;
        if (order == 2) { pw = ad * ad; dpw = 2.f * ad; }
        else if (order == 1) { pw = ad; dpw = 1.f; }
        else { pw = powf(ad, (float)order); dpw = order * powf(ad, (float)(order - 1)); }
        acc += pw;
        dpred[i] = scale * dpw * sg;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(loss, scale * acc);
}

__global__ void __launch_bounds__(256) bce_logits_kernel(const float* __restrict__ x, int n, float target, float scale,
                                                         float* loss, float* dscore) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float v = x[i];
        acc += fmaxf(v, 0.f) - v * target + log1pf(expf(-fabsf(v)));
        dscore[i] = scale * (1.f / (1.f + expf(-v)) - target);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(loss, scale * acc);
}

// feature_distance_loss on the global sums (single block of 1024 threads; F <= a few 10^4)
__global__ void __launch_bounds__(1024) distance_kernel(const float* __restrict__ sb, const float* __restrict__ so, int F,
                                                        float inv_B, int kind, float mult, float* loss, float* gbase,
                                                        float* gother, int accumulate_base) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < F; i += blockDim.x) {
        float d = (sb[i] - so[i]) * inv_B;
        float ad = fabsf(d);
        if (kind == 0) acc += ad;
        else if (kind == 1) acc -= ad;
        else if (kind == 2) acc -= sqrtf(ad + 1.f);
        else if (kind == 3) acc -= logf(ad + 1.f);
        else acc += d * d;                       // 4: square_mean, 5: norm_mean (sum of squares)
    }
    acc = block_sum(acc, red);
    float nrm = 0.f, lossv;
    if (kind == 5) { nrm = sqrtf(acc); lossv = nrm; }
    else lossv = acc / (float)F;
    if (threadIdx.x == 0) atomicAdd(loss, mult * lossv);
    const float invF = 1.f / (float)F;
    const float k = mult * inv_B;
    for (int i = threadIdx.x; i < F; i += blockDim.x) {
        float d = (sb[i] - so[i]) * inv_B;
        float ad = fabsf(d);
        float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        float g;
        if (kind == 0) g = sg * invF;
        else if (kind == 1) g = -sg * invF;
        else if (kind == 2) g = -sg * invF * 0.5f * rsqrtf(ad + 1.f);
        else if (kind == 3) g = -sg * invF / (ad + 1.f);
        else if (kind == 4) g = 2.f * d * invF;
        else g = d / nrm;
        g *= k;
        if (accumulate_base) gbase[i] += g; else gbase[i] = g;
        gother[i] = -g;
    }
}

// multi-block version for the distance functions that are a mean of an element-wise function (kinds 0..4): every block
// reduces its slice, adds its share of the loss atomically and writes its slice of the gradients
__global__ void __launch_bounds__(256) distance_mb_kernel(const float* __restrict__ sb, const float* __restrict__ so, int F,
                                                          float inv_B, int kind, float mult, float* loss, float* gbase,
                                                          float* gother, int accumulate_base) {
    __shared__ float red[32];
    const float invF = 1.f / (float)F;
    const float k = mult * inv_B;
    float acc = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < F; i += gridDim.x * blockDim.x) {
        const float d = (sb[i] - so[i]) * inv_B;
        const float ad = fabsf(d);
        const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        float g;
        if (kind == 0) { acc += ad; g = sg * invF; }
        else if (kind == 1) { acc -= ad; g = -sg * invF; }
        else if (kind == 2) { const float r = sqrtf(ad + 1.f); acc -= r; g = -sg * invF * 0.5f / r; }
        else if (kind == 3) { acc -= logf(ad + 1.f); g = -sg * invF / (ad + 1.f); }
        else { acc += d * d; g = 2.f * d * invF; }
        g *= k;
        if (accumulate_base) gbase[i] += g; else gbase[i] = g;
        gother[i] = -g;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(loss, mult * acc * invF);
}

// ------------------------------------------------------------------------------------------------------------
// gradient-penalty kernels (one block per sample / row)
// ------------------------------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(512) feature_norm_seed_kernel(const T* __restrict__ h, int cols, float* __restrict__ s_out,
                                                                T* __restrict__ gamma, int act, float slope) {
    __shared__ float red[32];
    const T* x = h + (long long)blockIdx.x * cols;
    T* g = gamma + (long long)blockIdx.x * cols;
    float acc = 0.f;
    if (VEC) {
        for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
            float4 v = ld4(x + c);
            acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
    } else {
        for (int c = threadIdx.x; c < cols; c += blockDim.x) { float v = to_f(x[c]); acc = fmaf(v, v, acc); }
    }
    acc = block_sum(acc, red);
    const float s = sqrtf(acc);
    if (threadIdx.x == 0) s_out[blockIdx.x] = s;
    const float inv = 1.f / s;
    if (VEC) {
        for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
            float4 v = ld4(x + c);          // second pass: the 64 KB row is L1/L2 resident
            st4(g + c, make_float4(v.x * inv * act_bwd(v.x, act, slope), v.y * inv * act_bwd(v.y, act, slope),
                                   v.z * inv * act_bwd(v.z, act, slope), v.w * inv * act_bwd(v.w, act, slope)));
        }
    } else {
        for (int c = threadIdx.x; c < cols; c += blockDim.x) {
            float v = to_f(x[c]);
            g[c] = from_f<T>(v * inv * act_bwd(v, act, slope));
        }
    }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(512) gp_feature_seed_kernel(const T* __restrict__ uL, const T* __restrict__ hL,
                                                              const float* __restrict__ s, T* __restrict__ out, int cols,
                                                              int act, float slope) {
    __shared__ float red[32];
    const T* u = uL + (long long)blockIdx.x * cols;
    const T* h = hL + (long long)blockIdx.x * cols;
    T* o = out + (long long)blockIdx.x * cols;
    const float inv = 1.f / s[blockIdx.x];
    float acc = 0.f;
    if (VEC) {
        for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
            float4 hv = ld4(h + c), uv = ld4(u + c);
            acc += (hv.x * uv.x + hv.y * uv.y + hv.z * uv.z + hv.w * uv.w) * inv;
        }
    } else {
        for (int c = threadIdx.x; c < cols; c += blockDim.x) acc = fmaf(to_f(h[c]) * inv, to_f(u[c]), acc);
    }
    const float dot = block_sum(acc, red);
    if (VEC) {
        for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
            float4 hv = ld4(h + c), uv = ld4(u + c);
            float4 r;
            r.x = (uv.x - hv.x * inv * dot) * inv * act_bwd(hv.x, act, slope);
            r.y = (uv.y - hv.y * inv * dot) * inv * act_bwd(hv.y, act, slope);
            r.z = (uv.z - hv.z * inv * dot) * inv * act_bwd(hv.z, act, slope);
            r.w = (uv.w - hv.w * inv * dot) * inv * act_bwd(hv.w, act, slope);
            st4(o + c, r);
        }
    } else {
        for (int c = threadIdx.x; c < cols; c += blockDim.x) {
            float hv = to_f(h[c]);
            float g = hv * inv;
            o[c] = from_f<T>((to_f(u[c]) - g * dot) * inv * act_bwd(hv, act, slope));
        }
    }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(512) gradnorm_penalty_kernel(const T* __restrict__ g0, long long per_sample,
                                                               float lam_over_B, float inv_B, float* __restrict__ gnorm,
                                                               float* penalty, float* gnorm_mean, T* __restrict__ u0) {
    __shared__ float red[32];
    const T* g = g0 + (long long)blockIdx.x * per_sample;
    T* u = u0 + (long long)blockIdx.x * per_sample;
    float acc = 0.f;
    if (VEC) {
        for (long long e = (long long)threadIdx.x * 4; e < per_sample; e += (long long)blockDim.x * 4) {
            float4 v = ld4(g + e);
            acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
    } else {
        for (long long e = threadIdx.x; e < per_sample; e += blockDim.x) { float v = to_f(g[e]); acc = fmaf(v, v, acc); }
    }
    acc = block_sum(acc, red);
    const float r = sqrtf(acc);
    const float ex = fmaxf(r - 1.f, 0.f);
    if (threadIdx.x == 0) {
        gnorm[blockIdx.x] = r;
        atomicAdd(penalty, lam_over_B * ex * ex);
        atomicAdd(gnorm_mean, inv_B * r);
    }
    const float coef = r > 0.f ? 2.f * lam_over_B * ex / r : 0.f;
    if (VEC) {
        for (long long e = (long long)threadIdx.x * 4; e < per_sample; e += (long long)blockDim.x * 4) {
            float4 v = ld4(g + e);      // second pass hits L2 (per-sample slab <= 600 KB)
            st4(u + e, make_float4(v.x * coef, v.y * coef, v.z * coef, v.w * coef));
        }
    } else {
        for (long long e = threadIdx.x; e < per_sample; e += blockDim.x) u[e] = from_f<T>(to_f(g[e]) * coef);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Adam + kernel-layout rewrite
// ------------------------------------------------------------------------------------------------------------
struct Dims4 { int d[4]; long long gs[4], s1[4], s2[4]; };

template <typename TO, bool UPDATE>
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ param, const float* __restrict__ grad,
                                                   float* __restrict__ m, float* __restrict__ v, Dims4 dm,
                                                   TO* __restrict__ out1, TO* __restrict__ out2,
                                                   const float* __restrict__ state, float beta1, float beta2, float eps,
                                                   float wd) {
    float step_size = 0.f, inv_sqrt_bc2 = 0.f;
    if (UPDATE) { step_size = state[1]; inv_sqrt_bc2 = state[2]; }
    const long long total = (long long)dm.d[0] * dm.d[1] * dm.d[2] * dm.d[3];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        // 32-bit index arithmetic (a tensor has < 2^32 elements: checked by the launcher); 64-bit divisions were the cost of
        // this kernel (2.1 TB/s on the 8.4 M-element generator fc weight)
        unsigned t = (unsigned)i;
        const unsigned q3 = t / (unsigned)dm.d[3]; const int i3 = (int)(t - q3 * (unsigned)dm.d[3]); t = q3;
        const unsigned q2 = t / (unsigned)dm.d[2]; const int i2 = (int)(t - q2 * (unsigned)dm.d[2]); t = q2;
        const unsigned q1 = t / (unsigned)dm.d[1]; const int i1 = (int)(t - q1 * (unsigned)dm.d[1]);
        const int i0 = (int)q1;
        float p = param[i];
        if (UPDATE) {
            float g = grad[i0 * dm.gs[0] + i1 * dm.gs[1] + i2 * dm.gs[2] + i3 * dm.gs[3]];
            if (wd != 0.f) g = fmaf(wd, p, g);
            float mm = beta1 * m[i] + (1.f - beta1) * g;
            float vv = beta2 * v[i] + (1.f - beta2) * g * g;
            m[i] = mm; v[i] = vv;
            float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
            p = p - step_size * (mm / denom);
            param[i] = p;
        }
        if (out1) out1[i0 * dm.s1[0] + i1 * dm.s1[1] + i2 * dm.s1[2] + i3 * dm.s1[3]] = from_f<TO>(p);
        if (out2) out2[i0 * dm.s2[0] + i1 * dm.s2[1] + i2 * dm.s2[2] + i3 * dm.s2[3]] = from_f<TO>(p);
    }
}

__global__ void adam_prepare_kernel(float* state, double lr, double beta1, double beta2) {
    const double t = (double)state[0] + 1.0;
    state[0] = (float)t;
    state[1] = (float)(lr / (1.0 - pow(beta1, t)));
    state[2] = (float)(1.0 / sqrt(1.0 - pow(beta2, t)));
}

static inline int ew_grid(long long work_items, int block) {
    long long b = (work_items + block - 1) / block;
    long long cap = 16LL * kNumSMs;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ============================================================================================================
// C ABI
// ============================================================================================================
// ------------------------------------------------------------------------------------------------------------
// adam_multi: torch.optim.Adam on many small tensors without kernel-layout copies (biases, BatchNorm weight / bias) in ONE
// launch: block = tensor (table row: {param pointer, gradient offset, moment offset, element count})
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) adam_multi_kernel(const long long* __restrict__ table, const float* __restrict__ grad,
                                                         float* __restrict__ m, float* __restrict__ v,
                                                         const float* __restrict__ state, float b1, float b2, float eps, float wd) {
    const long long* row = table + 4LL * blockIdx.x;
    float* p = reinterpret_cast<float*>(row[0]);
    const float* g = grad + row[1];
    float* mm = m + row[2];
    float* vv = v + row[2];
    const int n = (int)row[3];
    const float step_size = state[1], isb = state[2];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float gr = g[i];
        const float pv = p[i];
        if (wd != 0.f) gr = fmaf(wd, pv, gr);
        const float m1 = b1 * mm[i] + (1.f - b1) * gr;
        const float v1 = b2 * vv[i] + (1.f - b2) * gr * gr;
        mm[i] = m1; vv[i] = v1;
        p[i] = pv - step_size * (m1 / (sqrtf(v1) * isb + eps));
    }
}

// Adam + kernel-layout rewrite for MANY tensors in one launch (the crowd discriminator has 200 convolution weights: one
// adam_kernel launch each was 437 launches / 3 ms per step, and the D update sits on the step's critical path between the
// discriminator backward and the generator step).  Table row (26 x int64) per tensor:
//   [0] param ptr  [1] grad offset  [2] moment offset  [3..6] dims  [7..10] grad strides  [11] out1 ptr  [12..15] out1 strides
//   [16] out2 ptr  [17..20] out2 strides  [21] 1 = the layout copies are fp32 (prediction heads), 0 = bf16/activation dtype
//   [22] first block of this tensor  [23] number of elements
// A block handles ADAM_LM_CHUNK consecutive elements of one tensor (binary search of its first-block column).
constexpr int ADAM_LM_ROW = 26;
constexpr int ADAM_LM_CHUNK = 2048;
template <typename TO>
__global__ void __launch_bounds__(256) adam_layout_multi_kernel(const long long* __restrict__ table, int n_tensors,
                                                                const float* __restrict__ grad, float* __restrict__ m,
                                                                float* __restrict__ v, const float* __restrict__ state, float beta1,
                                                                float beta2, float eps, float wd) {
    int lo = 0, hi = n_tensors - 1;
    while (lo < hi) {                                    // last tensor whose first block <= blockIdx.x
        const int mid = (lo + hi + 1) >> 1;
        if (table[(long long)mid * ADAM_LM_ROW + 22] <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const long long* row = table + (long long)lo * ADAM_LM_ROW;
    float* param = reinterpret_cast<float*>(row[0]);
    const float* g0 = grad + row[1];
    float* mm = m + row[2];
    float* vv = v + row[2];
    const int d1 = (int)row[4], d2 = (int)row[5], d3 = (int)row[6];
    const long long gs0 = row[7], gs1 = row[8], gs2 = row[9], gs3 = row[10];
    void* out1 = reinterpret_cast<void*>(row[11]);
    void* out2 = reinterpret_cast<void*>(row[16]);
    const bool f32 = row[21] != 0;
    const long long n = row[23];
    const long long base = ((long long)blockIdx.x - row[22]) * ADAM_LM_CHUNK;
    const float step_size = state[1], inv_sqrt_bc2 = state[2];
    for (int k = threadIdx.x; k < ADAM_LM_CHUNK; k += 256) {
        const long long i = base + k;
        if (i >= n) break;
        unsigned t = (unsigned)i;                        // 32-bit index arithmetic: n < 2^32 per tensor
        const unsigned q3 = t / (unsigned)d3; const int i3 = (int)(t - q3 * (unsigned)d3); t = q3;
        const unsigned q2 = t / (unsigned)d2; const int i2 = (int)(t - q2 * (unsigned)d2); t = q2;
        const unsigned q1 = t / (unsigned)d1; const int i1 = (int)(t - q1 * (unsigned)d1);
        const int i0 = (int)q1;
        float p = param[i];
        float g = g0[i0 * gs0 + i1 * gs1 + i2 * gs2 + i3 * gs3];
        if (wd != 0.f) g = fmaf(wd, p, g);
        const float m1 = beta1 * mm[i] + (1.f - beta1) * g;
        const float v1 = beta2 * vv[i] + (1.f - beta2) * g * g;
        mm[i] = m1; vv[i] = v1;
        p = p - step_size * (m1 / (sqrtf(v1) * inv_sqrt_bc2 + eps));
        param[i] = p;
        if (out1) {
            const long long o = i0 * row[12] + i1 * row[13] + i2 * row[14] + i3 * row[15];
            if (f32) reinterpret_cast<float*>(out1)[o] = p; else reinterpret_cast<TO*>(out1)[o] = from_f<TO>(p);
        }
        if (out2) {
            const long long o = i0 * row[17] + i1 * row[18] + i2 * row[19] + i3 * row[20];
            if (f32) reinterpret_cast<float*>(out2)[o] = p; else reinterpret_cast<TO*>(out2)[o] = from_f<TO>(p);
        }
    }
}

extern "C" {

int srgan_adam_layout_multi(const long long* table, int n_tensors, long long total_blocks, const float* grad, float* m, float* v,
                            const float* state3, float beta1, float beta2, float eps, float weight_decay, int out_dtype,
                            void* stream) {
    SRGAN_REQUIRE(table && grad && m && v && state3 && n_tensors >= 0 && total_blocks >= 0 && total_blocks < 0x7fffffffLL,
                  "srgan_adam_layout_multi: bad arguments");
    SRGAN_REQUIRE(out_dtype == SRGAN_F32 || out_dtype == SRGAN_BF16, "srgan_adam_layout_multi: unknown dtype %d", out_dtype);
    if (n_tensors == 0 || total_blocks == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (out_dtype == SRGAN_F32)
        adam_layout_multi_kernel<float><<<(unsigned)total_blocks, 256, 0, st>>>(table, n_tensors, grad, m, v, state3, beta1, beta2, eps, weight_decay);
    else
        adam_layout_multi_kernel<bf16><<<(unsigned)total_blocks, 256, 0, st>>>(table, n_tensors, grad, m, v, state3, beta1, beta2, eps, weight_decay);
    SRGAN_CHECK_LAUNCH("adam_layout_multi_kernel");
    return SRGAN_OK;
}

int srgan_colsum(const void* X, long long rows, int cols, float* out, int mod, const float* rowscale, int dtype,
                 void* stream) {
    SRGAN_REQUIRE(X && out && cols >= 0 && rows >= 0, "srgan_colsum: bad arguments");
    SRGAN_REQUIRE(mod == 0 || cols % mod == 0, "srgan_colsum: cols %% mod != 0");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SRGAN_F32) return launch_colsum<float>((const float*)X, rows, cols, out, mod, rowscale, st);
    return launch_colsum<bf16>((const bf16*)X, rows, cols, out, mod, rowscale, st);
}

int srgan_rowdot(const void* X, int rows, int cols, const float* w, const float* bias, int bias_index, float* out,
                 int dtype, void* stream) {
    SRGAN_REQUIRE(X && w && bias && out && rows >= 0 && cols > 0, "srgan_rowdot: bad arguments");
    if (rows == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (cols >= 2048) {
        if (dtype == SRGAN_F32) rowdot_kernel<float, 256><<<rows, 256, 0, st>>>((const float*)X, rows, cols, w, bias, bias_index, out);
        else rowdot_kernel<bf16, 256><<<rows, 256, 0, st>>>((const bf16*)X, rows, cols, w, bias, bias_index, out);
    } else {
        int grid = cdiv(rows, 8);
        if (dtype == SRGAN_F32) rowdot_kernel<float, 32><<<grid, 256, 0, st>>>((const float*)X, rows, cols, w, bias, bias_index, out);
        else rowdot_kernel<bf16, 32><<<grid, 256, 0, st>>>((const bf16*)X, rows, cols, w, bias, bias_index, out);
    }
    SRGAN_CHECK_LAUNCH("rowdot_kernel");
    return SRGAN_OK;
}

int srgan_seed_rows(void* out, int rows, int cols, const float* gvec, const float* rowscale, const float* wrow,
                    const void* href, int act, float slope, int dtype, void* stream) {
    SRGAN_REQUIRE(out && href && rows >= 0 && cols > 0, "srgan_seed_rows: bad arguments");
    SRGAN_REQUIRE((rowscale == nullptr) == (wrow == nullptr), "srgan_seed_rows: rowscale and wrow go together");
    if (rows == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    bool vec = cols % 4 == 0;
    long long items = (long long)rows * cols / (vec ? 4 : 1);
    int grid = ew_grid(items, 256);
    if (dtype == SRGAN_F32) {
        if (vec) seed_rows_kernel<float, true><<<grid, 256, 0, st>>>((float*)out, rows, cols, gvec, rowscale, wrow, (const float*)href, act, slope);
        else seed_rows_kernel<float, false><<<grid, 256, 0, st>>>((float*)out, rows, cols, gvec, rowscale, wrow, (const float*)href, act, slope);
    } else {
        if (vec) seed_rows_kernel<bf16, true><<<grid, 256, 0, st>>>((bf16*)out, rows, cols, gvec, rowscale, wrow, (const bf16*)href, act, slope);
        else seed_rows_kernel<bf16, false><<<grid, 256, 0, st>>>((bf16*)out, rows, cols, gvec, rowscale, wrow, (const bf16*)href, act, slope);
    }
    SRGAN_CHECK_LAUNCH("seed_rows_kernel");
    return SRGAN_OK;
}

int srgan_nchw_to_nhwc(const float* src, void* dst, int n, int c, int h, int w, int dtype, void* stream) {
    SRGAN_REQUIRE(src && dst && n >= 0 && c > 0 && h > 0 && w > 0, "srgan_nchw_to_nhwc: bad arguments");
    if (n == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    long long hw = (long long)h * w;
    if (hw == 1 || c == 1) {
        long long total = (long long)n * c * hw;
        if (dtype == SRGAN_F32) cast_kernel<float><<<ew_grid(total, 256), 256, 0, st>>>(src, (float*)dst, total);
        else cast_kernel<bf16><<<ew_grid(total, 256), 256, 0, st>>>(src, (bf16*)dst, total);
    } else {
        long long total = (long long)n * hw;
        if (dtype == SRGAN_F32) nchw_to_nhwc_kernel<float><<<ew_grid(total, 256), 256, 0, st>>>(src, (float*)dst, n, c, hw);
        else nchw_to_nhwc_kernel<bf16><<<ew_grid(total, 256), 256, 0, st>>>(src, (bf16*)dst, n, c, hw);
    }
    SRGAN_CHECK_LAUNCH("nchw_to_nhwc_kernel");
    return SRGAN_OK;
}

int srgan_nhwc_to_nchw(const void* src, float* dst, int n, int c, int h, int w, int dtype, void* stream) {
    SRGAN_REQUIRE(src && dst && n >= 0 && c > 0 && h > 0 && w > 0, "srgan_nhwc_to_nchw: bad arguments");
    if (n == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    long long hw = (long long)h * w;
    if (hw == 1 || c == 1) {
        long long total = (long long)n * c * hw;
        if (dtype == SRGAN_F32) uncast_kernel<float><<<ew_grid(total, 256), 256, 0, st>>>((const float*)src, dst, total);
        else uncast_kernel<bf16><<<ew_grid(total, 256), 256, 0, st>>>((const bf16*)src, dst, total);
    } else {
        long long total = (long long)n * hw;
        if (dtype == SRGAN_F32) nhwc_to_nchw_kernel<float><<<ew_grid(total, 256), 256, 0, st>>>((const float*)src, dst, n, c, hw);
        else nhwc_to_nchw_kernel<bf16><<<ew_grid(total, 256), 256, 0, st>>>((const bf16*)src, dst, n, c, hw);
    }
    SRGAN_CHECK_LAUNCH("nhwc_to_nchw_kernel");
    return SRGAN_OK;
}

static int thin_params(ThinP& p, const char* who, int n, const srgan_geom* g, int kpad) {
    if (!g || n < 0 || g->Cb < 1 || g->Cb > 4 || kpad % 8 != 0 || kpad < g->R * g->S * g->Cb) {
        srgan_set_error("%s: needs 1 <= Cb <= 4 and kpad %% 8 == 0, kpad >= R*S*Cb", who);
        return SRGAN_ERR_ARG;
    }
    p = ThinP{n, g->Hs, g->Ws, g->Hl, g->Wl, g->Cb, g->R, g->S, g->stride, g->pad, kpad};
    return SRGAN_OK;
}

int srgan_im2col(const void* L, void* col, int n, const srgan_geom* g, int kpad, int dtype, void* stream) {
    SRGAN_REQUIRE(L && col, "srgan_im2col: null pointer");
    ThinP p;
    int rc = thin_params(p, "srgan_im2col", n, g, kpad);
    if (rc) return rc;
    if (n == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    long long items = (long long)n * g->Hs * g->Ws * (kpad / 8);
    if (dtype == SRGAN_F32) im2col_kernel<float><<<ew_grid(items, 256), 256, 0, st>>>((const float*)L, (float*)col, p);
    else if (g->Cb == 3 && g->R == 4 && g->S == 4 && g->stride == 2 && g->pad == 1 && kpad == 64 && g->Hl == 2 * g->Hs &&
             g->Wl == 2 * g->Ws)
        im2col_c3k4s2_kernel<<<ew_grid((long long)n * g->Hs * g->Ws * 4, 256), 256, 0, st>>>((const bf16*)L, (bf16*)col, n, g->Hs,
                                                                                           g->Ws, g->Hl, g->Wl);
    else im2col_kernel<bf16><<<ew_grid(items, 256), 256, 0, st>>>((const bf16*)L, (bf16*)col, p);
    SRGAN_CHECK_LAUNCH("im2col_kernel");
    return SRGAN_OK;
}

int srgan_col2im(const void* col, void* L_out, int n, const srgan_geom* g, int kpad, const float* bias, const void* href,
                 int epi, int act, float slope, int dtype, void* stream) {
    SRGAN_REQUIRE(col && L_out, "srgan_col2im: null pointer");
    ThinP p;
    int rc = thin_params(p, "srgan_col2im", n, g, kpad);
    if (rc) return rc;
    if (n == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    long long items = (long long)n * g->Hl * g->Wl;
    if (dtype == SRGAN_F32) col2im_kernel<float><<<ew_grid(items, 256), 256, 0, st>>>((const float*)col, (float*)L_out, p, bias, (const float*)href, epi, act, slope);
    else if (g->Cb == 3 && g->R == 4 && g->S == 4 && g->stride == 2 && g->pad == 1 && kpad == 64 && g->Hl == 2 * g->Hs &&
             g->Wl == 2 * g->Ws)
        col2im_c3k4s2_kernel<<<ew_grid(items, 256), 256, 0, st>>>((const bf16*)col, (bf16*)L_out, n, g->Hs, g->Ws, bias,
                                                                 (const bf16*)href, epi, act, slope);
    else col2im_kernel<bf16><<<ew_grid(items, 256), 256, 0, st>>>((const bf16*)col, (bf16*)L_out, p, bias, (const bf16*)href, epi, act, slope);
    SRGAN_CHECK_LAUNCH("col2im_kernel");
    return SRGAN_OK;
}

int srgan_interpolate(const void* u, const void* fake, const float* alpha, void* out, int n, long long per_sample,
                      int dtype, void* stream) {
    SRGAN_REQUIRE(u && fake && alpha && out && n >= 0 && per_sample > 0, "srgan_interpolate: bad arguments");
    if (n == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    long long total = (long long)n * per_sample;
    bool vec = per_sample % 4 == 0;
    int grid = ew_grid(total / (vec ? 4 : 1), 256);
    if (dtype == SRGAN_F32) {
        if (vec) interpolate_kernel<float, true><<<grid, 256, 0, st>>>((const float*)u, (const float*)fake, alpha, (float*)out, total, per_sample);
        else interpolate_kernel<float, false><<<grid, 256, 0, st>>>((const float*)u, (const float*)fake, alpha, (float*)out, total, per_sample);
    } else {
        if (vec) interpolate_kernel<bf16, true><<<grid, 256, 0, st>>>((const bf16*)u, (const bf16*)fake, alpha, (bf16*)out, total, per_sample);
        else interpolate_kernel<bf16, false><<<grid, 256, 0, st>>>((const bf16*)u, (const bf16*)fake, alpha, (bf16*)out, total, per_sample);
    }
    SRGAN_CHECK_LAUNCH("interpolate_kernel");
    return SRGAN_OK;
}

int srgan_labeled_loss(const float* pred, const float* y, int n, int order, float scale, float* loss, float* dpred,
                       void* stream) {
    SRGAN_REQUIRE(pred && y && loss && dpred && n > 0 && order >= 1, "srgan_labeled_loss: bad arguments");
    labeled_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pred, y, n, order, scale, loss, dpred);
    SRGAN_CHECK_LAUNCH("labeled_loss_kernel");
    return SRGAN_OK;
}

int srgan_bce_logits(const float* scores, int n, float target, float scale, float* loss, float* dscore, void* stream) {
    SRGAN_REQUIRE(scores && loss && dscore && n > 0, "srgan_bce_logits: bad arguments");
    bce_logits_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(scores, n, target, scale, loss, dscore);
    SRGAN_CHECK_LAUNCH("bce_logits_kernel");
    return SRGAN_OK;
}

int srgan_distance(const float* sum_base, const float* sum_other, int F, float inv_B, int kind, float mult,
                   float* loss, float* gbase, float* gother, int accumulate_base, void* stream) {
    SRGAN_REQUIRE(sum_base && sum_other && loss && gbase && gother && F > 0, "srgan_distance: bad arguments");
    SRGAN_REQUIRE(kind >= 0 && kind <= 5, "srgan_distance: unknown distance kind %d", kind);
    if (kind != 5 && F >= 4096) {
        int grid = (F + 1023) / 1024;
        if (grid > 2 * kNumSMs) grid = 2 * kNumSMs;
        distance_mb_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sum_base, sum_other, F, inv_B, kind, mult, loss, gbase,
                                                                   gother, accumulate_base);
    } else
        distance_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(sum_base, sum_other, F, inv_B, kind, mult, loss, gbase, gother,
                                                              accumulate_base);
    SRGAN_CHECK_LAUNCH("distance_kernel");
    return SRGAN_OK;
}

int srgan_feature_norm_seed(const void* h, int rows, int cols, float* s_out, void* gamma_out, int act, float slope,
                            int dtype, void* stream) {
    SRGAN_REQUIRE(h && s_out && gamma_out && rows >= 0 && cols > 0, "srgan_feature_norm_seed: bad arguments");
    if (rows == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int th = cols >= 2048 ? 512 : 128;
    if (cols % 4 == 0) {
        if (dtype == SRGAN_F32) feature_norm_seed_kernel<float, true><<<rows, th, 0, st>>>((const float*)h, cols, s_out, (float*)gamma_out, act, slope);
        else feature_norm_seed_kernel<bf16, true><<<rows, th, 0, st>>>((const bf16*)h, cols, s_out, (bf16*)gamma_out, act, slope);
    } else {
        if (dtype == SRGAN_F32) feature_norm_seed_kernel<float, false><<<rows, th, 0, st>>>((const float*)h, cols, s_out, (float*)gamma_out, act, slope);
        else feature_norm_seed_kernel<bf16, false><<<rows, th, 0, st>>>((const bf16*)h, cols, s_out, (bf16*)gamma_out, act, slope);
    }
    SRGAN_CHECK_LAUNCH("feature_norm_seed_kernel");
    return SRGAN_OK;
}

int srgan_gp_feature_seed(const void* uL, const void* hL, const float* s, void* out, int rows, int cols, int act,
                          float slope, int dtype, void* stream) {
    SRGAN_REQUIRE(uL && hL && s && out && rows >= 0 && cols > 0, "srgan_gp_feature_seed: bad arguments");
    if (rows == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int th = cols >= 2048 ? 512 : 128;
    if (cols % 4 == 0) {
        if (dtype == SRGAN_F32) gp_feature_seed_kernel<float, true><<<rows, th, 0, st>>>((const float*)uL, (const float*)hL, s, (float*)out, cols, act, slope);
        else gp_feature_seed_kernel<bf16, true><<<rows, th, 0, st>>>((const bf16*)uL, (const bf16*)hL, s, (bf16*)out, cols, act, slope);
    } else {
        if (dtype == SRGAN_F32) gp_feature_seed_kernel<float, false><<<rows, th, 0, st>>>((const float*)uL, (const float*)hL, s, (float*)out, cols, act, slope);
        else gp_feature_seed_kernel<bf16, false><<<rows, th, 0, st>>>((const bf16*)uL, (const bf16*)hL, s, (bf16*)out, cols, act, slope);
    }
    SRGAN_CHECK_LAUNCH("gp_feature_seed_kernel");
    return SRGAN_OK;
}

int srgan_gradnorm_penalty(const void* g0, int n, long long per_sample, float lam_over_B, float inv_B, float* gnorm,
                           float* penalty, float* gnorm_mean, void* u0, int dtype, void* stream) {
    SRGAN_REQUIRE(g0 && gnorm && penalty && gnorm_mean && u0 && n >= 0 && per_sample > 0,
                  "srgan_gradnorm_penalty: bad arguments");
    if (n == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    bool vec = per_sample % 4 == 0;
    if (dtype == SRGAN_F32) {
        if (vec) gradnorm_penalty_kernel<float, true><<<n, 512, 0, st>>>((const float*)g0, per_sample, lam_over_B, inv_B, gnorm, penalty, gnorm_mean, (float*)u0);
        else gradnorm_penalty_kernel<float, false><<<n, 512, 0, st>>>((const float*)g0, per_sample, lam_over_B, inv_B, gnorm, penalty, gnorm_mean, (float*)u0);
    } else {
        if (vec) gradnorm_penalty_kernel<bf16, true><<<n, 512, 0, st>>>((const bf16*)g0, per_sample, lam_over_B, inv_B, gnorm, penalty, gnorm_mean, (bf16*)u0);
        else gradnorm_penalty_kernel<bf16, false><<<n, 512, 0, st>>>((const bf16*)g0, per_sample, lam_over_B, inv_B, gnorm, penalty, gnorm_mean, (bf16*)u0);
    }
    SRGAN_CHECK_LAUNCH("gradnorm_penalty_kernel");
    return SRGAN_OK;
}

static int fill_dims(Dims4& dm, const int* dims4, const long long* gs, const long long* s1, const long long* s2) {
    for (int i = 0; i < 4; ++i) {
        if (dims4[i] <= 0) return -1;
        dm.d[i] = dims4[i];
        dm.gs[i] = gs ? gs[i] : 0;
        dm.s1[i] = s1 ? s1[i] : 0;
        dm.s2[i] = s2 ? s2[i] : 0;
    }
    return 0;
}

int srgan_adam_prepare(float* state3, double lr, double beta1, double beta2, void* stream) {
    SRGAN_REQUIRE(state3 != nullptr, "srgan_adam_prepare: null state");
    adam_prepare_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state3, lr, beta1, beta2);
    SRGAN_CHECK_LAUNCH("adam_prepare_kernel");
    return SRGAN_OK;
}

int srgan_adam(float* param, const float* grad, float* m, float* v, const int* dims4, const long long* gstrides4,
               void* out1, const long long* o1strides4, void* out2, const long long* o2strides4, int out_dtype,
               const float* state3, float beta1, float beta2, float eps, float weight_decay, void* stream) {
    SRGAN_REQUIRE(param && grad && m && v && dims4 && gstrides4 && state3, "srgan_adam: bad arguments");
    SRGAN_REQUIRE((out1 == nullptr) || o1strides4, "srgan_adam: out1 without strides");
    SRGAN_REQUIRE((out2 == nullptr) || o2strides4, "srgan_adam: out2 without strides");
    Dims4 dm;
    SRGAN_REQUIRE(fill_dims(dm, dims4, gstrides4, o1strides4, o2strides4) == 0, "srgan_adam: non-positive dim");
    long long total = (long long)dm.d[0] * dm.d[1] * dm.d[2] * dm.d[3];
    SRGAN_REQUIRE(total < 0xffffffffLL, "srgan_adam / srgan_repack: tensors of 2^32 or more elements are not supported");
    cudaStream_t st = (cudaStream_t)stream;
    int grid = ew_grid(total, 256);
    if (out_dtype == SRGAN_F32)
        adam_kernel<float, true><<<grid, 256, 0, st>>>(param, grad, m, v, dm, (float*)out1, (float*)out2, state3, beta1, beta2, eps, weight_decay);
    else
        adam_kernel<bf16, true><<<grid, 256, 0, st>>>(param, grad, m, v, dm, (bf16*)out1, (bf16*)out2, state3, beta1, beta2, eps, weight_decay);
    SRGAN_CHECK_LAUNCH("adam_kernel");
    return SRGAN_OK;
}

int srgan_repack(const float* param, const int* dims4, void* out1, const long long* o1strides4, void* out2,
                 const long long* o2strides4, int out_dtype, void* stream) {
    SRGAN_REQUIRE(param && dims4, "srgan_repack: bad arguments");
    Dims4 dm;
    SRGAN_REQUIRE(fill_dims(dm, dims4, nullptr, o1strides4, o2strides4) == 0, "srgan_repack: non-positive dim");
    long long total = (long long)dm.d[0] * dm.d[1] * dm.d[2] * dm.d[3];
    SRGAN_REQUIRE(total < 0xffffffffLL, "srgan_adam / srgan_repack: tensors of 2^32 or more elements are not supported");
    cudaStream_t st = (cudaStream_t)stream;
    int grid = ew_grid(total, 256);
    if (out_dtype == SRGAN_F32)
        adam_kernel<float, false><<<grid, 256, 0, st>>>(const_cast<float*>(param), nullptr, nullptr, nullptr, dm, (float*)out1, (float*)out2, nullptr, 0.f, 0.f, 0.f, 0.f);
    else
        adam_kernel<bf16, false><<<grid, 256, 0, st>>>(const_cast<float*>(param), nullptr, nullptr, nullptr, dm, (bf16*)out1, (bf16*)out2, nullptr, 0.f, 0.f, 0.f, 0.f);
    SRGAN_CHECK_LAUNCH("repack_kernel");
    return SRGAN_OK;
}


int srgan_adam_multi(const long long* table, int n_tensors, const float* grad, float* m, float* v, const float* state3,
                     float beta1, float beta2, float eps, float weight_decay, void* stream) {
    SRGAN_REQUIRE(table && grad && m && v && state3 && n_tensors >= 0, "srgan_adam_multi: bad arguments");
    if (n_tensors == 0) return SRGAN_OK;
    adam_multi_kernel<<<n_tensors, 128, 0, (cudaStream_t)stream>>>(table, grad, m, v, state3, beta1, beta2, eps, weight_decay);
    SRGAN_CHECK_LAUNCH("adam_multi_kernel");
    return SRGAN_OK;
}

}  // extern "C"
