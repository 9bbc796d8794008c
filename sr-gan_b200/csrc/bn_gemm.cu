// DenseNet dense-layer GEMMs with the eval-mode BatchNorm + ReLU that precedes them fused in (crowd/models.py:335-353:
// norm1 -> relu1 -> conv1 of _DenseLayer, and norm -> relu -> conv of _Transition :363-371).
//
// The trunk's 1x1 convolutions are [pixels x C] GEMMs over the first C channels of the block's concat buffer; C grows to
// 1920 while the other side of the GEMM stays 128 wide, so these launches are HBM-bound on the C-wide streams.  Unfused, a
// dense layer moves 10 C-wide streams per pixel (affine: read cat, write n1 | GEMM: read n1 || data gradient: read n1 as the
// mask, write dn1 | BatchNorm backward: read dn1, read cat, read + write dcat | weight gradient: read n1).
//
//   bn_dgrad_kernel   dcat[:, :C] (+)= ((dy . W) * [bn(cat) > 0]) * gamma/sigma, and the BatchNorm parameter gradients
//                     dgamma += sum d * (cat - mean) / sigma, dbeta += sum d      (d = the masked product), in the epilogue
//                     of the [pixels x K] x [K x C] data-gradient GEMM: 3 C-wide streams (read cat, read + write dcat)
//                     instead of 6, one launch instead of two.  The GEMM itself is short (K = 128): the kernel is its
//                     epilogue, laid out for it -- accumulator rows are rounded to bf16, transposed through a per-warp
//                     shared-memory buffer so that a lane owns 8 channels x 4 rows (16-byte global accesses, 8 rows x 64
//                     contiguous bytes per warp instruction), the cat / dcat pieces of the NEXT tile are already in flight
//                     (cp.async into a lane-private ring of landing slots), the per-channel scale / mean / shift tables
//                     live in shared memory and the partial sums of dgamma / dbeta in registers until the channel tile of
//                     the CTA's (contiguous) tile range changes.
#include "umma_common.cuh"

namespace {

// gamma / sqrt(var + eps): the same branch-free form as the streaming BatchNorm kernels (graph_ops.cu), so that the mask
// recomputed here agrees bit for bit with the forward pass that produced n1
__device__ __forceinline__ float bn_scale_f(float gamma, float var, float eps) {
    const float v = var + eps;
    float y = rsqrtf(v);
    y = y * (1.5f - 0.5f * v * y * y);
    return gamma * y;
}

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
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }     // the 8 epilogue warps

constexpr int BD_THREADS = 320;                  // warp 0 TMA, warp 1 MMA + TMEM, warps 2..9 epilogue
constexpr int BD_BN = 128;                       // output channels per tile
constexpr int BD_STAGE_BYTES = A_STAGE_BYTES + BD_BN * KCH * 2;
constexpr int BD_NCH = BD_BN / 64;               // 32-column chunks per epilogue warp and tile
constexpr int BD_ZONE_WARP = BD_NCH * 2 * EPI_STG_BYTES;      // per warp: chunks x {cat, dcat} x (4 row groups x 32 lanes x 16 B)

struct BnDgradParams {
    long long rows;               // GEMM rows (pixels x samples)
    int K;                        // reduction length = channels of dy (a multiple of 64)
    int C;                        // output channels that exist (the BatchNorm's channels; a multiple of 8)
    int Cpad;                     // C rounded up to 32: length of the shared-memory tables
    int pitch;                    // elements between consecutive rows of x / dx (the concat buffer's channel count)
    const bf16* x;                // BatchNorm input (the concat buffer), rows aligned with the GEMM rows
    bf16* dx;                     // its delta
    const float *gamma, *beta, *mean, *var;
    float eps;
    float *dgamma, *dbeta;        // nullptr: data gradient only
    bf16* d_out;                  // nullptr, or: also store the masked product d (the delta w.r.t. the BatchNorm output's
    int d_pitch;                  // pre-activation), rows of d_pitch elements -- the gradient-penalty chain keeps it for the
                                  // tangent block's BatchNorm-scale gradient
    int accumulate;               // dx += (else dx =)
    int m_tiles, total_tiles, stages;
};

__global__ void __launch_bounds__(BD_THREADS, 1) bn_dgrad_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB, const BnDgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_slot;

    constexpr int BN = BD_BN;
    constexpr int TMEM_COLS = 2 * BN;
    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stages = p.stages;
    const int nch = p.K / KCH;
    // a CTA walks a contiguous range of the tile list (row tiles fastest): its consecutive tiles cover the same channels
    const int tile_begin = (int)((long long)p.total_tiles * blockIdx.x / gridDim.x);
    const int tile_end = (int)((long long)p.total_tiles * (blockIdx.x + 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tmem_full_bar[b]), 1); mbar_init(smem_u32(&tmem_empty_bar[b]), 8); }
        fence_barrier_init();
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                const int mt = tile % p.m_tiles, ny = tile / p.m_tiles;
                for (int ch = 0; ch < nch; ++ch) {
                    mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    mbar_expect_tx(fb, BD_STAGE_BYTES);
                    const uint32_t dst = tiles + s * BD_STAGE_BYTES;
                    tma_load_2d(dst, &tmA, fb, ch * KCH, mt * TILE_M);
                    tma_load_2d(dst + A_STAGE_BYTES, &tmB, fb, ch * KCH, ny * BN);
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(TILE_M, BN, 0, 0);
            const uint64_t desc0 = make_desc(0, 16, 1024);
            int s = 0, tl = 0;
            uint32_t ph = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile, ++tl) {
                const int buf = tl & 1;
                const uint32_t bph = (tl >> 1) & 1;
                mbar_wait(smem_u32(&tmem_empty_bar[buf]), bph ^ 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + buf * BN;
                for (int k_it = 0; k_it < nch; ++k_it) {
                    mbar_wait(smem_u32(&full_bar[s]), ph);
                    tc_fence_after();
                    const uint32_t a_s = tiles + s * BD_STAGE_BYTES;
                    const uint64_t ad0 = desc0 + (uint64_t)(a_s >> 4);
                    const uint64_t bd0 = desc0 + (uint64_t)((a_s + A_STAGE_BYTES) >> 4);
#pragma unroll
                    for (int k = 0; k < KCH / 16; ++k)
                        umma_f16(acc, ad0 + (uint64_t)(k * 2), bd0 + (uint64_t)(k * 2), idesc, (k_it > 0 || k > 0) ? 1u : 0u);
                    umma_commit(smem_u32(&empty_bar[s]));
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
                umma_commit(smem_u32(&tmem_full_bar[buf]));
            }
        }
    } else {
        // ================= epilogue (8 warps) =================
        const int ew = warp - 2;
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int half = ew >> 2;                // which half of the tile's 32-column chunks
        const int t_unit = lane & 3, t_row = lane >> 2;       // transposed role: 16-byte unit t_unit of rows 8*it + t_row
        const int te = threadIdx.x - 64;         // 0..255 among the epilogue threads
        const uint32_t after_ring = (uint32_t)stages * BD_STAGE_BYTES;
        const uint32_t stg = tiles + after_ring + (uint32_t)ew * EPI_STG_BYTES;
        const uint32_t zone = tiles + after_ring + 8u * EPI_STG_BYTES + (uint32_t)ew * BD_ZONE_WARP;
        float* tab = reinterpret_cast<float*>(smem_raw + (tiles - smem_u32(smem_raw)) + after_ring + 8u * EPI_STG_BYTES + 8u * BD_ZONE_WARP);
        float* t_s = tab;                        // gamma / sigma
        float* t_mu = tab + p.Cpad;
        float* t_be = tab + 2 * p.Cpad;
        const bool grads = p.dgamma != nullptr;
        for (int c = te; c < p.Cpad; c += 256) {
            float s = 0.f, mu = 0.f, be = 0.f;
            if (c < p.C) { s = bn_scale_f(__ldg(p.gamma + c), __ldg(p.var + c), p.eps); mu = __ldg(p.mean + c); be = __ldg(p.beta + c); }
            t_s[c] = s; t_mu[c] = mu; t_be[c] = be;
        }
        epi_bar_sync();

        // The cat / dcat pieces travel through a lane-private ring of 8 slots per warp (slot k = chunk k/4, row group k%4;
        // 32 lanes x 16 bytes x 2 streams each), one cp.async group per slot: a slot is refilled with the NEXT tile's piece
        // right after it has been consumed, so every piece is in flight for a whole tile time and ~all of the zone (64 KB
        // per CTA) is outstanding at any moment (Little: 44 GB/s per SM x ~1.5 us).
        auto issue = [&](int tile, int k) {
            if (tile < tile_end) {
                const int mt = tile % p.m_tiles, ny = tile / p.m_tiles;
                const int cb = ny * BN + (half + 2 * (k >> 2)) * 32 + t_unit * 8;
                const long long gr = (long long)mt * TILE_M + q * 32 + 8 * (k & 3) + t_row;
                const bool ok = gr < p.rows && cb < p.C;
                const long long off = ok ? gr * p.pitch + cb : 0;
                cp_async16(zone + (uint32_t)(k * 1024 + lane * 16), p.x + off, ok ? 16u : 0u);
                if (p.accumulate) cp_async16(zone + (uint32_t)(k * 1024 + 512 + lane * 16), p.dx + off, ok ? 16u : 0u);
            }
            cp_async_commit();                   // always: the group count per slot stays fixed
        };
#pragma unroll
        for (int k = 0; k < 4 * BD_NCH; ++k) issue(tile_begin, k);

        // partial sums of dgamma / dbeta stay in registers while consecutive tiles cover the same channels (a CTA walks a
        // contiguous range of tiles, row tiles fastest): [chunk][0..7] sum d * (x - mean), [chunk][8..15] sum d over the rows
        // this lane has seen; combined across the 8 lanes that share the channels and added to global memory when the
        // channel tile changes
        float acc[BD_NCH][16];
#pragma unroll
        for (int jj = 0; jj < BD_NCH; ++jj)
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[jj][e] = 0.f;
        auto flush = [&](int ny) {
#pragma unroll
            for (int jj = 0; jj < BD_NCH; ++jj) {
                float* a = acc[jj];
                {   // recursive halving over lane bits 4, 3, 2: kind = bit 4 (0: dgamma term, 1: dbeta term),
                    // channel = 4*bit3 + 2*bit2 + {0, 1} of this lane's 8
                    const bool hi = (lane & 16) != 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float send = hi ? a[i] : a[i + 8], keep = hi ? a[i + 8] : a[i];
                        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
                }
                {
                    const bool hi = (lane & 8) != 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float send = hi ? a[i] : a[i + 4], keep = hi ? a[i + 4] : a[i];
                        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
                }
                {
                    const bool hi = (lane & 4) != 0;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float send = hi ? a[i] : a[i + 2], keep = hi ? a[i + 2] : a[i];
                        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                }
                const int c = ny * BN + (half + 2 * jj) * 32 + t_unit * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (c + i < p.C && a[i] != 0.f) {
                        if (lane & 16) { if (p.dbeta != nullptr) atomicAdd(p.dbeta + c + i, a[i]); }
                        else atomicAdd(p.dgamma + c + i, a[i] / sqrtf(__ldg(p.var + c + i) + p.eps));
                    }
                }
#pragma unroll
                for (int e = 0; e < 16; ++e) a[e] = 0.f;
            }
        };
        int tl = 0, ny_acc = tile_begin / p.m_tiles;
        for (int tile = tile_begin; tile < tile_end; ++tile, ++tl) {
            const int mt = tile % p.m_tiles, ny = tile / p.m_tiles;
            const int buf = tl & 1;
            const uint32_t bph = (tl >> 1) & 1;
            if (grads && ny != ny_acc) { flush(ny_acc); ny_acc = ny; }
            mbar_wait(smem_u32(&tmem_full_bar[buf]), bph);
            tc_fence_after();
#pragma unroll
            for (int jj = 0; jj < BD_NCH; ++jj) {
                const int j = half + 2 * jj;
                const int cbase = ny * BN + j * 32;
                const bool live = cbase < p.C;   // warp-uniform: chunks beyond the last channel (partial last N tile) carry no data
                const int cb = cbase + t_unit * 8;
                const bool cok = cb < p.C;
                float s8[8], mu8[8], be8[8];
                if (live) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + j * 32, v);
                    tmem_ld_wait();
                    // round to bf16 (what the unfused data-gradient kernel stores) and transpose: row-per-lane -> 8 channels x 4 rows
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 w;
                        __nv_bfloat162 b0 = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
                        __nv_bfloat162 b1 = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
                        __nv_bfloat162 b3 = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
                        w.x = *reinterpret_cast<uint32_t*>(&b0); w.y = *reinterpret_cast<uint32_t*>(&b1);
                        w.z = *reinterpret_cast<uint32_t*>(&b2); w.w = *reinterpret_cast<uint32_t*>(&b3);
                        sts128(stg_addr(stg, lane, g), w);
                    }
                    __syncwarp();
                    const int ct = cok ? cb : 0;
                    const float4* ps = reinterpret_cast<const float4*>(t_s + ct);
                    const float4* pm = reinterpret_cast<const float4*>(t_mu + ct);
                    const float4* pb = reinterpret_cast<const float4*>(t_be + ct);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 a = ps[h], b = pm[h], c = pb[h];
                        s8[4 * h] = a.x; s8[4 * h + 1] = a.y; s8[4 * h + 2] = a.z; s8[4 * h + 3] = a.w;
                        mu8[4 * h] = b.x; mu8[4 * h + 1] = b.y; mu8[4 * h + 2] = b.z; mu8[4 * h + 3] = b.w;
                        be8[4 * h] = c.x; be8[4 * h + 1] = c.y; be8[4 * h + 2] = c.z; be8[4 * h + 3] = c.w;
                    }
                }
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int k = jj * 4 + it;
                    cp_async_wait_group<4 * BD_NCH - 1>();      // the oldest slot = this one has landed (a thread reads back its own copies)
                    if (live) {
                        const uint4 dr = lds128(stg_addr(stg, 8 * it + t_row, t_unit));
                        const uint4 xr = lds128(zone + (uint32_t)(k * 1024 + lane * 16));
                        float d[8], xv[8], o[8], dm[8];
                        unpack8(dr, d);
                        unpack8(xr, xv);
                        if (p.accumulate) {
                            const uint4 cr = lds128(zone + (uint32_t)(k * 1024 + 512 + lane * 16));
                            unpack8(cr, o);
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) o[e] = 0.f;
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float xm = xv[e] - mu8[e];
                            const float dd = fmaf(xm, s8[e], be8[e]) > 0.f ? d[e] : 0.f;
                            acc[jj][e] = fmaf(dd, xm, acc[jj][e]);
                            acc[jj][8 + e] += dd;
                            o[e] = fmaf(dd, s8[e], o[e]);
                            dm[e] = dd;
                        }
                        const long long gr = (long long)mt * TILE_M + q * 32 + 8 * it + t_row;
                        if (cok && gr < p.rows) {
                            *reinterpret_cast<uint4*>(p.dx + gr * p.pitch + cb) = pack8(o);
                            if (p.d_out != nullptr) *reinterpret_cast<uint4*>(p.d_out + gr * p.d_pitch + cb) = pack8(dm);
                        }
                    }
                    issue(tile + 1, k);
                }
                __syncwarp();                    // the transposition buffer is rewritten by the next chunk
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[buf]));
        }
        cp_async_wait_all();
        if (grads) flush(ny_acc);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

// returns 1 = launched, 0 = shape not eligible, <0 = error
int bn_dgrad(const void* dy, const void* Wu, void* dx, const void* x, long long rows, int K, int Cout, int C, int pitch,
             const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* dgamma, float* dbeta,
             void* d_out, int d_pitch, int accumulate, cudaStream_t st) {
    if (K % KCH != 0 || Cout % 64 != 0 || C > Cout || C <= 0 || (C & 7) || (pitch & 7) || pitch < C || rows <= 0) return 0;
    if (((uintptr_t)dy | (uintptr_t)Wu | (uintptr_t)dx | (uintptr_t)x | (uintptr_t)d_out) & 15) return 0;
    if (d_out != nullptr && ((d_pitch & 7) || d_pitch < C)) return 0;
    BnDgradParams p;
    p.rows = rows; p.K = K; p.C = C; p.Cpad = (C + 31) / 32 * 32; p.pitch = pitch;
    p.x = (const bf16*)x; p.dx = (bf16*)dx;
    p.gamma = gamma; p.beta = beta; p.mean = mean; p.var = var; p.eps = eps;
    p.dgamma = dgamma; p.dbeta = dbeta; p.accumulate = accumulate;
    p.d_out = (bf16*)d_out; p.d_pitch = d_pitch;
    const long long m_tiles = (rows + TILE_M - 1) / TILE_M;
    const long long n_tiles = (C + BD_BN - 1) / BD_BN;
    if (m_tiles * n_tiles > 0x7fffffffLL) return 0;
    p.m_tiles = (int)m_tiles; p.total_tiles = (int)(m_tiles * n_tiles);
    const int fixed = 8 * EPI_STG_BYTES + 8 * BD_ZONE_WARP + 3 * p.Cpad * 4 + 1024;
    int stages = (226 * 1024 - fixed) / BD_STAGE_BYTES;
    if (stages > 8) stages = 8;
    if (stages < 2) return 0;
    p.stages = stages;
    const size_t smem = (size_t)stages * BD_STAGE_BYTES + fixed;
    CUtensorMap tmA, tmB;
    int rc = encode_mat(&tmA, dy, rows, K, TILE_M);
    if (rc) return rc;
    rc = encode_mat(&tmB, Wu, Cout, K, BD_BN);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(bn_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) { srgan_set_error("cudaFuncSetAttribute(bn_dgrad_kernel): %s", cudaGetErrorString(e)); return SRGAN_ERR_CUDA; }
        attr_set = true;
    }
    const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
    bn_dgrad_kernel<<<grid, BD_THREADS, smem, st>>>(tmA, tmB, p);
    SRGAN_CHECK_LAUNCH("bn_dgrad_kernel");
    return 1;
}
