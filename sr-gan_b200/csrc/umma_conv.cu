// tcgen05 (5th-gen tensor core) implicit-GEMM kernels for the conv pair, bf16 operands, fp32 accumulation in TMEM.
//
//   down / up  (umma_conv_kernel):   D[m, c] = sum_{tap} sum_{ch} A_tap[m, ch] * W[c, (tap, ch)]
//     * 128-row M tile = a TW x TH x TN patch of the output row grid (pixels x samples); for every filter tap the A
//       operand is ONE tiled TMA load of the input at the tap's shifted coordinates (TMA zero-fills the padding halo;
//       stride-2 convs use the tensor map's element strides), landing in shared memory in the K-major SWIZZLE_128B
//       layout tcgen05.mma consumes -- no im2col buffer ever exists in HBM;
//     * the weight operand is a plain 2-D TMA tile of Wd[a][(r,s,b)] / Wu[b][(r,s,a)];
//     * warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2-5 = epilogue
//       (tcgen05.ld -> bias + LeakyReLU/tanh, or multiply by act'(href) for the backward / tangent passes -> bf16 NHWC);
//   wgrad (umma_wgrad_kernel): see below.
//
// Replaces cuDNN fprop / dgrad behind age/models.py:44-52,68-80 and crowd/models.py:139-147.
#include "umma_common.cuh"

namespace {


struct UmmaConvParams {
    int mode;                     // 0 down, 1 up
    int n;                        // samples
    int Hm, Wm;                   // output row grid per sample (per phase for up)
    int TW, TH, TN;               // tile patch: TW*TH*TN == 128
    int tiles_w, tiles_h;
    int Cin, Cout;
    int R, S, stride, pad;
    int Hout, Wout;               // full output spatial extent (== Hm, Wm for down)
    const float* bias;
    int bias_mod;
    const bf16* href;
    bf16* out;
    int out_pitch;                // elements between consecutive output pixels (= Cout for a dense output; the row pitch of a
    int out_valid;                // concat buffer when the output is a channel window of it) and the channels that exist there
    int epi, act;
    float slope;
    int stages;
    int href_smem;                // 1: the epilogue's href pieces of tile i+1 are fetched with cp.async into a shared-memory
                                  // landing zone while tile i is processed (short-K GEMMs: the MMAs do not hide the latency)
};

// Epilogue math on one 32-column chunk of one accumulator row (all branches are warp-uniform and hoisted out of the
// element loops; bias is read as float4 runs, href arrives as the row's four 16-byte pieces), packed to 64 bytes of bf16.
__device__ __forceinline__ void epilogue_math(const uint32_t (&v)[32], const UmmaConvParams& p, int cbase, const uint4 (&hv)[4],
                                              uint4 (&w)[4]) {
    float f[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
    if (p.epi == SRGAN_EPI_BIAS_ACT) {
        if (p.bias != nullptr) {
            if (p.bias_mod == 1) {                // one scalar bias for every column (ConvTranspose2d with out_channels = 1)
                const float b = __ldg(p.bias);
#pragma unroll
                for (int e = 0; e < 32; ++e) f[e] += b;
            } else {
                const int bi = p.bias_mod ? cbase % p.bias_mod : cbase;
                const float4* bp = reinterpret_cast<const float4*>(p.bias + bi);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float4 b = __ldg(bp + g);
                    f[g * 4] += b.x; f[g * 4 + 1] += b.y; f[g * 4 + 2] += b.z; f[g * 4 + 3] += b.w;
                }
            }
        }
        if (p.act == SRGAN_ACT_LEAKY) {
            const float sl = p.slope;
#pragma unroll
            for (int e = 0; e < 32; ++e) f[e] = f[e] > 0.f ? f[e] : f[e] * sl;
        } else if (p.act == SRGAN_ACT_TANH) {
#pragma unroll
            for (int e = 0; e < 32; ++e) f[e] = tanhf(f[e]);
        }
    } else if (p.href != nullptr && p.act == SRGAN_ACT_LEAKY && p.slope == 0.f) {
        // ReLU mask (the DenseNet trunk's data gradients: HBM-bound GEMMs whose cost is this epilogue): round first, then
        // AND with the packed comparison mask -- 3 instructions per element pair instead of 8, same bits as the fp32 path
        // (the mask multiplies by exactly 1 or 0)
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(hv);
        const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
        uint32_t* wo = reinterpret_cast<uint32_t*>(w);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const __nv_bfloat162 d2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
            wo[e] = *reinterpret_cast<const uint32_t*>(&d2) & __hgt2_mask(h2[e], zero2);
        }
        return;
    } else if (p.href != nullptr && p.act != SRGAN_ACT_NONE) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(hv);
        if (p.act == SRGAN_ACT_LEAKY) {
            const float sl = p.slope;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float2 hf = __bfloat1622float2(h2[e]);
                f[2 * e] *= hf.x > 0.f ? 1.f : sl;
                f[2 * e + 1] *= hf.y > 0.f ? 1.f : sl;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float2 hf = __bfloat1622float2(h2[e]);
                f[2 * e] *= 1.f - hf.x * hf.x;
                f[2 * e + 1] *= 1.f - hf.y * hf.y;
            }
        }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        __nv_bfloat162 b0 = __floats2bfloat162_rn(f[g * 8 + 0], f[g * 8 + 1]);
        __nv_bfloat162 b1 = __floats2bfloat162_rn(f[g * 8 + 2], f[g * 8 + 3]);
        __nv_bfloat162 b2 = __floats2bfloat162_rn(f[g * 8 + 4], f[g * 8 + 5]);
        __nv_bfloat162 b3 = __floats2bfloat162_rn(f[g * 8 + 6], f[g * 8 + 7]);
        w[g].x = *reinterpret_cast<uint32_t*>(&b0); w[g].y = *reinterpret_cast<uint32_t*>(&b1);
        w[g].z = *reinterpret_cast<uint32_t*>(&b2); w[g].w = *reinterpret_cast<uint32_t*>(&b3);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Persistent kernel: one CTA per SM walks a static round-robin list of output tiles.
//   * CTA tile = MT x 128 rows (MT sub-tiles share every weight tile: halves the weight traffic per FLOP) x BN columns;
//   * a deep TMA ring (as many 128-byte-swizzled stages as fit in ~212 KB) covers the L2/HBM latency that the
//     3-stage version exposed (ncu: tensor pipe 22 %, nothing else above 30 %);
//   * the accumulator is double-buffered in TMEM (2 x MT x BN <= 512 columns): the 8 epilogue warps drain tile i while
//     the MMA thread already accumulates tile i+1.
// ------------------------------------------------------------------------------------------------------------
constexpr int PC_THREADS = 320;                  // warp 0 TMA, warp 1 MMA + TMEM, warps 2..9 epilogue


struct UmmaConvParamsP {
    UmmaConvParams c;
    int m_subtiles;               // 128-row sub-tiles per phase
    int n_tiles;                  // Cout / BN
    int total_tiles;              // phases * n_tiles * (m_subtiles / MT)
};

template <int MT, int BN>
__global__ void __launch_bounds__(PC_THREADS, 1) umma_conv_persistent_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                            const __grid_constant__ CUtensorMap tmB,
                                                                            const UmmaConvParamsP pp) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[12];
    __shared__ __align__(8) uint64_t empty_bar[12];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_slot;

    const UmmaConvParams& p = pp.c;
    constexpr int B_STAGE_BYTES = BN * KCH * 2;
    constexpr int STAGE_BYTES = MT * A_STAGE_BYTES + B_STAGE_BYTES;
    constexpr int ACC_COLS = MT * BN;
    constexpr int TMEM_COLS = 2 * ACC_COLS;
    static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS >= 32, "TMEM budget");
    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stages = p.stages;
    const int nch = p.Cin / KCH;
    const int m_tiles = pp.m_subtiles / MT;
    const int sub_per_n = p.tiles_w * p.tiles_h;

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

    // per-tile decode shared by the three roles
    struct Tile { int mt, ny, pa, pb, r0, s0, qa, qb, Rt, St; };
    auto decode = [&](int t) {
        Tile T;
        T.mt = t % m_tiles; t /= m_tiles;
        T.ny = t % pp.n_tiles; t /= pp.n_tiles;
        T.pa = 0; T.pb = 0; T.r0 = 0; T.s0 = 0; T.qa = 0; T.qb = 0; T.Rt = p.R; T.St = p.S;
        if (p.mode == 1) {
            T.pa = t / p.stride; T.pb = t % p.stride;
            T.r0 = (T.pa + p.pad) % p.stride; T.s0 = (T.pb + p.pad) % p.stride;
            T.qa = (T.pa + p.pad - T.r0) / p.stride; T.qb = (T.pb + p.pad - T.s0) / p.stride;
            T.Rt = (p.R - T.r0 + p.stride - 1) / p.stride;
            T.St = (p.S - T.s0 + p.stride - 1) / p.stride;
        }
        return T;
    };

    if (warp == 0) {
        // ================= TMA producer (one elected thread) =================
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            const int mul = p.mode == 0 ? p.stride : 1;
            for (int tile = blockIdx.x; tile < pp.total_tiles; tile += gridDim.x) {
                const Tile T = decode(tile);
                int cw[MT], chh[MT], cn[MT];
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    int ms = T.mt * MT + i;
                    cw[i] = (ms % p.tiles_w) * p.TW * mul;
                    chh[i] = ((ms / p.tiles_w) % p.tiles_h) * p.TH * mul;
                    cn[i] = (ms / sub_per_n) * p.TN;
                }
                for (int tr = 0; tr < T.Rt; ++tr)
                    for (int ts = 0; ts < T.St; ++ts) {
                        int dw, dh, kcol;
                        if (p.mode == 0) { dw = -p.pad + ts; dh = -p.pad + tr; kcol = (tr * p.S + ts) * p.Cin; }
                        else {
                            dw = T.qb - ts; dh = T.qa - tr;
                            kcol = ((T.r0 + p.stride * tr) * p.S + (T.s0 + p.stride * ts)) * p.Cin;
                        }
                        for (int ch = 0; ch < nch; ++ch) {
                            mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
                            const uint32_t fb = smem_u32(&full_bar[s]);
                            mbar_expect_tx(fb, STAGE_BYTES);
                            const uint32_t dst = tiles + s * STAGE_BYTES;
#pragma unroll
                            for (int i = 0; i < MT; ++i)
                                tma_load_4d(dst + i * A_STAGE_BYTES, &tmA, fb, ch * KCH, cw[i] + dw, chh[i] + dh, cn[i]);
                            tma_load_2d(dst + MT * A_STAGE_BYTES, &tmB, fb, kcol + ch * KCH, T.ny * BN);
                            if (++s == stages) { s = 0; ph ^= 1; }
                        }
                    }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one elected thread) =================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(TILE_M, BN, 0, 0);
            const uint64_t desc0 = make_desc(0, 16, 1024);      // the 14-bit address field is added per operand below
            int s = 0, tl = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < pp.total_tiles; tile += gridDim.x, ++tl) {
                const Tile T = decode(tile);
                const int n_iters = T.Rt * T.St * nch;
                const int buf = tl & 1;
                const uint32_t bph = (tl >> 1) & 1;
                mbar_wait(smem_u32(&tmem_empty_bar[buf]), bph ^ 1);     // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t acc = tmem_base + buf * ACC_COLS;
                for (int k_it = 0; k_it < n_iters; ++k_it) {
                    mbar_wait(smem_u32(&full_bar[s]), ph);
                    tc_fence_after();
                    const uint32_t a_s = tiles + s * STAGE_BYTES;
                    const uint64_t ad0 = desc0 + (uint64_t)(a_s >> 4);
                    const uint64_t bd0 = desc0 + (uint64_t)((a_s + MT * A_STAGE_BYTES) >> 4);
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
#pragma unroll
                        for (int k = 0; k < KCH / 16; ++k)
                            umma_f16(acc + i * BN, ad0 + (uint64_t)(i * (A_STAGE_BYTES >> 4) + k * 2), bd0 + (uint64_t)(k * 2), idesc,
                                     (k_it > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(smem_u32(&empty_bar[s]));
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
                umma_commit(smem_u32(&tmem_full_bar[buf]));
            }
        }
    } else {
        // ================= epilogue (8 warps): TMEM -> registers -> smem transposition -> bf16 NHWC =================
        const int ew = warp - 2;
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int half = ew >> 2;                // which half of the 32-column chunks
        const int row = q * 32 + lane;
        const int tw = row % p.TW, th = (row / p.TW) % p.TH, tn = row / (p.TW * p.TH);
        const uint32_t stg = tiles + (uint32_t)stages * STAGE_BYTES + (uint32_t)ew * EPI_STG_BYTES;
        const int t_unit = lane & 3, t_row = lane >> 2;       // transposed role: unit t_unit of rows 8*it + t_row
        const bool use_href = p.epi != SRGAN_EPI_BIAS_ACT && p.href != nullptr && p.act != SRGAN_ACT_NONE;
        constexpr int NCH = (BN / 32 + 1) / 2;   // chunks per sub-tile handled by this warp (BN = 64: one)
        const bool hsm = use_href && p.href_smem;
        // landing zone of the cp.async href prefetch: per warp MT x NCH blocks of 32 rows x 64 bytes, XOR-swizzled like stg
        const uint32_t hz = tiles + (uint32_t)stages * STAGE_BYTES + 8u * EPI_STG_BYTES + (uint32_t)ew * (MT * NCH * EPI_STG_BYTES);
        // element offsets of the rows this lane moves in the transposed accesses (-1: row beyond the last sample)
        auto row_offsets = [&](const Tile& T, long long (&o_t)[MT][4]) {
            const int c0 = T.ny * BN;
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const int ms = T.mt * MT + i;
                const int tw_i = ms % p.tiles_w, th_i = (ms / p.tiles_w) % p.tiles_h, tn_i = ms / sub_per_n;
                const int sample = tn_i * p.TN + tn;
                const int oy = th_i * p.TH + th, ox = tw_i * p.TW + tw;
                long long o;
                if (p.mode == 0) o = (((long long)sample * p.Hm + oy) * p.Wm + ox) * p.out_pitch + c0;
                else o = (((long long)sample * p.Hout + (oy * p.stride + T.pa)) * p.Wout + (ox * p.stride + T.pb)) * p.out_pitch + c0;
                if (sample >= p.n) o = -1;
#pragma unroll
                for (int it = 0; it < 4; ++it) o_t[i][it] = __shfl_sync(0xffffffffu, o, 8 * it + t_row);
            }
        };
        // issue the asynchronous copies of one tile's href pieces into the landing zone (out-of-range pieces are zero-filled)
        auto href_prefetch = [&](int tile) {
            const Tile T = decode(tile);
            const int c0 = T.ny * BN;
            long long o_n[MT][4];
            row_offsets(T, o_n);
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int jj = 0; jj < NCH; ++jj) {
                    const int j = half + 2 * jj;
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const bool ok = j < BN / 32 && c0 + j * 32 + t_unit * 8 < p.out_valid && o_n[i][it] >= 0;
                        const bf16* src = ok ? p.href + o_n[i][it] + j * 32 + t_unit * 8 : p.href;
                        cp_async16(stg_addr(hz + (uint32_t)(i * NCH + jj) * EPI_STG_BYTES, 8 * it + t_row, t_unit), src, ok ? 16u : 0u);
                    }
                }
            cp_async_commit();
        };
        if (hsm && (int)blockIdx.x < pp.total_tiles) href_prefetch(blockIdx.x);
        int tl = 0;
        for (int tile = blockIdx.x; tile < pp.total_tiles; tile += gridDim.x, ++tl) {
            const Tile T = decode(tile);
            const int buf = tl & 1;
            const uint32_t bph = (tl >> 1) & 1;
            const int c0 = T.ny * BN;
            long long o_t[MT][4];
            row_offsets(T, o_t);
            // hreg: [hsm] this thread's OWN row of every chunk (4 x 16 bytes), read from the landing zone that was filled
            // while the previous tile was processed; the copies of the NEXT tile are issued right after, so that they are
            // in flight for a whole tile time.  [!hsm] the transposed pieces, requested before the accumulator is waited
            // for (their latency hides behind a long K loop).
            uint4 hreg[MT][NCH][4];
            if (hsm) {
                cp_async_wait_all();
                __syncwarp();
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int jj = 0; jj < NCH; ++jj)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            hreg[i][jj][k] = lds128(stg_addr(hz + (uint32_t)(i * NCH + jj) * EPI_STG_BYTES, lane, k));
                __syncwarp();
                if (tile + (int)gridDim.x < pp.total_tiles) href_prefetch(tile + gridDim.x);
            } else if (use_href) {
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int jj = 0; jj < NCH; ++jj) {
                        const int j = half + 2 * jj;
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            hreg[i][jj][it] = make_uint4(0u, 0u, 0u, 0u);
                            if (j < BN / 32 && c0 + j * 32 + t_unit * 8 < p.out_valid && o_t[i][it] >= 0)
                                hreg[i][jj][it] = __ldg(reinterpret_cast<const uint4*>(p.href + o_t[i][it] + j * 32 + t_unit * 8));
                        }
                    }
            }
            mbar_wait(smem_u32(&tmem_full_bar[buf]), bph);
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < MT; ++i) {
#pragma unroll
                for (int jj = 0; jj < NCH; ++jj) {
                    const int j = half + 2 * jj;
                    if (j >= BN / 32 || c0 + j * 32 >= p.out_valid) continue;      // warp-uniform; partial last N tile / channel window
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * ACC_COLS + i * BN + j * 32, v);
                    uint4 hv[4];
                    if (hsm) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) hv[k] = hreg[i][jj][k];
                    } else if (use_href) {                    // transposed pieces -> this thread's own row
#pragma unroll
                        for (int it = 0; it < 4; ++it) sts128(stg_addr(stg, 8 * it + t_row, t_unit), hreg[i][jj][it]);
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < 4; ++k) hv[k] = lds128(stg_addr(stg, lane, k));
                        __syncwarp();
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) hv[k] = make_uint4(0u, 0u, 0u, 0u);
                    }
                    tmem_ld_wait();
                    uint4 w[4];
                    epilogue_math(v, p, c0 + j * 32, hv, w);
#pragma unroll
                    for (int k = 0; k < 4; ++k) sts128(stg_addr(stg, lane, k), w[k]);
                    __syncwarp();
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const uint4 x = lds128(stg_addr(stg, 8 * it + t_row, t_unit));
                        if (o_t[i][it] >= 0 && c0 + j * 32 + t_unit * 8 < p.out_valid)
                            *reinterpret_cast<uint4*>(p.out + o_t[i][it] + j * 32 + t_unit * 8) = x;
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[buf]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
// wgrad:  dW[a, tap, b] += sum_{pixels} S[pix, a] * L[pix shifted by tap, b]        (fp32 atomics into dW)
//   M = 128 channels a (A operand MN-major: the TMA box [64 a x 64 pixels] x 2 is already "rows = K index"),
//   N = BN channels b per tap (B operand MN-major), K = pixels, NT taps accumulate side by side in TMEM
//   (NT*BN <= 512 columns) so one S tile feeds NT MMAs.  K is split across CTAs (blockIdx.z).
// ------------------------------------------------------------------------------------------------------------
// pixels (K) per stage: 32 (40 KB stages, five in flight) for BN >= 128; 64 for BN = 64, where a stage is ten small TMA
// boxes and the per-box issue cost matters more than ring depth
constexpr int WG_THREADS = 192;

struct UmmaWgradParams {
    int n, Hs, Ws, Ca, Cb, R, S, stride, pad;
    int TW, TH, TN;               // pixel patch per stage: TW*TH*TN == 64
    int tiles_w, tiles_h, tiles_n;
    int NT;                       // taps per CTA
    int NB;                       // b tiles (BN channels each) per CTA: NT*NB accumulators share every S tile
    int chunks_per_split;
    float* dW;
    int stages;
};

template <int BN, int WG_PIX>
__global__ void __launch_bounds__(WG_THREADS) umma_wgrad_kernel(const __grid_constant__ CUtensorMap tmS,
                                                                const __grid_constant__ CUtensorMap tmL,
                                                                const UmmaWgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_slot;

    constexpr int A_BYTES = 2 * WG_PIX * 128;            // two 64-channel column groups of [64 pixels x 128 B]
    constexpr int B_TAP_BYTES = (BN / 64) * WG_PIX * 128;
    const int NT = p.NT, NB = p.NB, NS = NT * NB;            // NS sub-problems (tap, b tile) per CTA
    const int STAGE_BYTES = A_BYTES + NS * B_TAP_BYTES;
    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stages = p.stages;

    // blockIdx.x -> (a tile, tap group, b tile); blockIdx.z -> K split
    const int b_tiles = (p.Cb + NB * BN - 1) / (NB * BN), tap_groups = (p.R * p.S) / NT;
    int t = blockIdx.x;
    const int bt = t % b_tiles; t /= b_tiles;
    const int tg = t % tap_groups; t /= tap_groups;
    const int at = t;
    const int a0 = at * 128, b0 = bt * NB * BN, tap0 = tg * NT;     // channels >= Cb of the last group: TMA zero fill
    const int total_chunks = p.tiles_w * p.tiles_h * p.tiles_n;
    const int ch_begin = blockIdx.z * p.chunks_per_split;
    int ch_end = ch_begin + p.chunks_per_split;
    if (ch_end > total_chunks) ch_end = total_chunks;
    const int n_iters = ch_end - ch_begin;
    const uint32_t ncols = NS * BN <= 32 ? 32 : (NS * BN <= 64 ? 64 : (NS * BN <= 128 ? 128 : (NS * BN <= 256 ? 256 : 512)));

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        mbar_init(smem_u32(&tmem_full_bar), 1);
        fence_barrier_init();
        prefetch_tmap(&tmS);
        prefetch_tmap(&tmL);
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), ncols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (n_iters > 0) {
        if (warp == 0) {
            if (elect_one()) {
                int s = 0;
                uint32_t ph = 0;
                for (int it = 0; it < n_iters; ++it) {
                    int c = ch_begin + it;
                    const int tw_i = c % p.tiles_w; c /= p.tiles_w;
                    const int th_i = c % p.tiles_h; c /= p.tiles_h;
                    const int tn_i = c;
                    mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    mbar_expect_tx(fb, STAGE_BYTES);
                    const uint32_t dst = tiles + s * STAGE_BYTES;
                    tma_load_4d(dst, &tmS, fb, a0, tw_i * p.TW, th_i * p.TH, tn_i * p.TN);
                    tma_load_4d(dst + WG_PIX * 128, &tmS, fb, a0 + 64, tw_i * p.TW, th_i * p.TH, tn_i * p.TN);
                    for (int k = 0; k < NS; ++k) {
                        const int tap = tap0 + k / NB, r = tap / p.S, sx = tap % p.S, bk = b0 + (k % NB) * BN;
                        const int lw = tw_i * p.TW * p.stride - p.pad + sx, lh = th_i * p.TH * p.stride - p.pad + r;
                        for (int g = 0; g < BN / 64; ++g)
                            tma_load_4d(dst + A_BYTES + k * B_TAP_BYTES + g * WG_PIX * 128, &tmL, fb, bk + g * 64, lw, lh,
                                        tn_i * p.TN);
                    }
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            }
        } else if (warp == 1) {
            if (elect_one()) {
                constexpr uint32_t idesc = make_idesc(128, BN, 1, 1);
                const uint64_t desc0 = make_desc(0, WG_PIX * 128, 1024);
                int s = 0;
                uint32_t ph = 0;
                for (int it = 0; it < n_iters; ++it) {
                    mbar_wait(smem_u32(&full_bar[s]), ph);
                    tc_fence_after();
                    const uint32_t a_s = tiles + s * STAGE_BYTES;
                    const uint64_t ad0 = desc0 + (uint64_t)(a_s >> 4);
                    for (int k = 0; k < NS; ++k) {
                        const uint64_t bd0 = desc0 + (uint64_t)((a_s + A_BYTES + k * B_TAP_BYTES) >> 4);
#pragma unroll
                        for (int kk = 0; kk < WG_PIX / 16; ++kk)      // 16 K rows (pixels) = 2048 B further into every column group
                            umma_f16(tmem_base + k * BN, ad0 + (uint64_t)(kk * 128), bd0 + (uint64_t)(kk * 128), idesc,
                                     (it > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(smem_u32(&empty_bar[s]));
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
                umma_commit(smem_u32(&tmem_full_bar));
            }
        } else {
            const int q = warp & 3;
            const int a = a0 + q * 32 + lane;                 // this thread's output row (channel a)
            mbar_wait(smem_u32(&tmem_full_bar), 0);
            tc_fence_after();
            const long long rowN = (long long)p.R * p.S * p.Cb;
#pragma unroll 1
            for (int k = 0; k < NS; ++k) {
#pragma unroll 1
                for (int j = 0; j < BN / 32; ++j) {
                    const int bcol = b0 + (k % NB) * BN + j * 32;
                    if (bcol >= p.Cb) continue;               // warp-uniform: zero-filled tail of the last b group
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + k * BN + j * 32, v);
                    tmem_ld_wait();
                    if (a >= p.Ca) continue;                  // zero-filled half tile (Ca = 64 mod 128)
                    float* dst = p.dW + (long long)a * rowN + (long long)(tap0 + k / NB) * p.Cb + bcol;
#pragma unroll
                    for (int e = 0; e < 32; e += 4)
                        red_add_v4(dst + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]),
                                   __uint_as_float(v[e + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------------------
// wgrad, few small-side channels (DenseNet dense-layer conv2: 3x3, stride 1, Ca = growth_rate = 32 real channels):
//   dW[a, tap, b] = sum_o S[o, a] * L[o - pad + tap, b] = sum_i L[i, b] * S[i + pad - tap, a]
// The kernel above would pad a to a 128-row MMA (4x the tensor work) and load the 128-channel L tile once per tap (9x the
// L2 -> SM traffic; ncu: 492 us per launch at 56x56, 320 samples = 0.5 TB/s algorithmic).  Here the roles are swapped:
//   M = 128 channels b (A operand = the L tile, loaded ONCE per pixel chunk, unshifted),
//   N = 32 channels a per tap (B operand = the S tile shifted by pad - tap: 64 real bytes per pixel; out-of-image pixels
//       and the channels beyond the window are TMA zero fill), all R*S taps side by side in TMEM (R*S*32 <= 512 columns),
//   K = pixels, split across one wave of CTAs; fp32 reductions into dW (lanes = consecutive b: coalesced).
// ------------------------------------------------------------------------------------------------------------
struct UmmaWgradSwapParams {
    int n, H, W, Ca_valid, Cb, R, S, pad;
    int TW, TH, TN;               // pixel patch per stage: TW*TH*TN == WS_PIX
    int tiles_w, tiles_h, tiles_n;
    int chunks_per_split;
    float* dW;
    int stages;
};
constexpr int WS_PIX = 32;
constexpr int WS_N = 32;

__global__ void __launch_bounds__(WG_THREADS) umma_wgrad_swap_kernel(const __grid_constant__ CUtensorMap tmS,
                                                                     const __grid_constant__ CUtensorMap tmL,
                                                                     const UmmaWgradSwapParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_slot;

    constexpr int A_BYTES = 2 * WS_PIX * 128;            // two 64-channel column groups of the L tile
    constexpr int B_TAP_BYTES = WS_PIX * 128;            // one 64-channel box of S per tap (32 of them used by the MMA)
    const int taps = p.R * p.S;
    const int STAGE_BYTES = A_BYTES + taps * B_TAP_BYTES;
    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stages = p.stages;
    const int b0 = blockIdx.x * 128;
    const int total_chunks = p.tiles_w * p.tiles_h * p.tiles_n;
    const int ch_begin = blockIdx.z * p.chunks_per_split;
    int ch_end = ch_begin + p.chunks_per_split;
    if (ch_end > total_chunks) ch_end = total_chunks;
    const int n_iters = ch_end - ch_begin;
    const uint32_t ncols = taps * WS_N <= 32 ? 32 : (taps * WS_N <= 64 ? 64 : (taps * WS_N <= 128 ? 128 : (taps * WS_N <= 256 ? 256 : 512)));

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        mbar_init(smem_u32(&tmem_full_bar), 1);
        fence_barrier_init();
        prefetch_tmap(&tmS);
        prefetch_tmap(&tmL);
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), ncols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (n_iters > 0) {
        if (warp == 0) {
            if (elect_one()) {
                int s = 0;
                uint32_t ph = 0;
                for (int it = 0; it < n_iters; ++it) {
                    int c = ch_begin + it;
                    const int tw_i = c % p.tiles_w; c /= p.tiles_w;
                    const int th_i = c % p.tiles_h; c /= p.tiles_h;
                    const int tn_i = c;
                    mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    mbar_expect_tx(fb, STAGE_BYTES);
                    const uint32_t dst = tiles + s * STAGE_BYTES;
                    const int w0 = tw_i * p.TW, h0 = th_i * p.TH, n0 = tn_i * p.TN;
                    tma_load_4d(dst, &tmL, fb, b0, w0, h0, n0);
                    tma_load_4d(dst + WS_PIX * 128, &tmL, fb, b0 + 64, w0, h0, n0);
                    for (int k = 0; k < taps; ++k) {
                        const int r = k / p.S, sx = k % p.S;
                        tma_load_4d(dst + A_BYTES + k * B_TAP_BYTES, &tmS, fb, 0, w0 + p.pad - sx, h0 + p.pad - r, n0);
                    }
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            }
        } else if (warp == 1) {
            if (elect_one()) {
                constexpr uint32_t idesc = make_idesc(128, WS_N, 1, 1);
                const uint64_t desc0 = make_desc(0, WS_PIX * 128, 1024);
                int s = 0;
                uint32_t ph = 0;
                for (int it = 0; it < n_iters; ++it) {
                    mbar_wait(smem_u32(&full_bar[s]), ph);
                    tc_fence_after();
                    const uint32_t a_s = tiles + s * STAGE_BYTES;
                    const uint64_t ad0 = desc0 + (uint64_t)(a_s >> 4);
                    for (int k = 0; k < taps; ++k) {
                        const uint64_t bd0 = desc0 + (uint64_t)((a_s + A_BYTES + k * B_TAP_BYTES) >> 4);
#pragma unroll
                        for (int kk = 0; kk < WS_PIX / 16; ++kk)
                            umma_f16(tmem_base + k * WS_N, ad0 + (uint64_t)(kk * 128), bd0 + (uint64_t)(kk * 128), idesc,
                                     (it > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(smem_u32(&empty_bar[s]));
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
                umma_commit(smem_u32(&tmem_full_bar));
            }
        } else {
            const int q = warp & 3;
            const int b = b0 + q * 32 + lane;                 // this thread's accumulator row = channel b
            mbar_wait(smem_u32(&tmem_full_bar), 0);
            tc_fence_after();
            const long long rowN = (long long)taps * p.Cb;
#pragma unroll 1
            for (int k = 0; k < taps; ++k) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + k * WS_N, v);
                tmem_ld_wait();
                if (b >= p.Cb) continue;
                float* dst = p.dW + (long long)k * p.Cb + b;       // dW[a][tap][b]: a warp's 32 lanes are 32 consecutive b
#pragma unroll
                for (int a = 0; a < WS_N; ++a)
                    if (a < p.Ca_valid) atomicAdd(dst + (long long)a * rowN, __uint_as_float(v[a]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

template <int MT, int BN>
int launch_conv_persistent(const CUtensorMap& tmA, const CUtensorMap& tmB, UmmaConvParamsP& pp, cudaStream_t st) {
    constexpr int stage_bytes = MT * A_STAGE_BYTES + BN * KCH * 2;
    // short-K GEMMs with a masking epilogue (the DenseNet trunk's data gradients: K = 128): the epilogue is the critical
    // path and the href latency is not hidden by the MMAs (ncu: 15 % of all stall samples on the first use of the href
    // registers) -> prefetch the next tile's href through a shared-memory landing zone (one tile of href bytes)
    const int k_iters = pp.c.R * pp.c.S * (pp.c.Cin / KCH) / (pp.c.mode == 1 ? pp.c.stride * pp.c.stride : 1);
    const bool use_href = pp.c.epi != SRGAN_EPI_BIAS_ACT && pp.c.href != nullptr && pp.c.act != SRGAN_ACT_NONE;
    static const bool href_smem_on = [] { const char* e = getenv("SRGAN_NO_HREF_SMEM"); return !(e && e[0] == '1'); }();
    pp.c.href_smem = (use_href && k_iters <= 4 && href_smem_on) ? 1 : 0;
    const int zone = pp.c.href_smem ? MT * TILE_M * BN * 2 : 0;
    int stages = (212 * 1024 - zone) / stage_bytes;
    if (stages > 12) stages = 12;
    if (stages < 2) { pp.c.href_smem = 0; stages = (212 * 1024) / stage_bytes; }
    pp.c.stages = stages;
    size_t smem = (size_t)stages * stage_bytes + 8 * EPI_STG_BYTES + (pp.c.href_smem ? zone : 0) + 1024;   // ring + epilogue transposition buffers + href landing zone + alignment
    static srgan_per_device_once attr_set;
    if (attr_set.need()) {
        cudaError_t e = cudaFuncSetAttribute(umma_conv_persistent_kernel<MT, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             226 * 1024);
        if (e != cudaSuccess) { srgan_set_error("cudaFuncSetAttribute(umma_conv_persistent_kernel): %s", cudaGetErrorString(e)); return SRGAN_ERR_CUDA; }
        attr_set.done();
    }
    int grid = pp.total_tiles < kNumSMs ? pp.total_tiles : kNumSMs;
    umma_conv_persistent_kernel<MT, BN><<<grid, PC_THREADS, smem, st>>>(tmA, tmB, pp);
    SRGAN_CHECK_LAUNCH("umma_conv_persistent_kernel");
    return 1;
}

template <int BN, int WG_PIX>
int launch_wgrad(const CUtensorMap& tmS, const CUtensorMap& tmL, UmmaWgradParams& p, dim3 grid, cudaStream_t st) {
    const int stage_bytes = 2 * WG_PIX * 128 + p.NT * p.NB * (BN / 64) * WG_PIX * 128;
    p.stages = (200 * 1024) / stage_bytes;
    if (p.stages > 8) p.stages = 8;
    size_t smem = (size_t)p.stages * stage_bytes + 1024;
    static srgan_per_device_once attr_set;
    if (attr_set.need()) {
        cudaError_t e = cudaFuncSetAttribute(umma_wgrad_kernel<BN, WG_PIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
        if (e != cudaSuccess) { srgan_set_error("cudaFuncSetAttribute(umma_wgrad_kernel): %s", cudaGetErrorString(e)); return SRGAN_ERR_CUDA; }
        attr_set.done();
    }
    umma_wgrad_kernel<BN, WG_PIX><<<grid, WG_THREADS, smem, st>>>(tmS, tmL, p);
    SRGAN_CHECK_LAUNCH("umma_wgrad_kernel");
    return 1;
}

}  // namespace

// flat3x3.cu
int flat3x3_conv(const void* src, const void* Wd, void* out, int n, const srgan_geom* g, int in_pitch, int in_valid,
                 int out_pitch, int out_valid, cudaStream_t st);

// returns 1 = launched on the tensor cores, 0 = shape not eligible (caller uses the SIMT kernel), <0 = error
int umma_conv(int mode, const void* src, const void* W, void* out, int n, const srgan_geom* g, const float* bias,
              int bias_mod, const void* href, int epi, int act, float slope, const srgan_views* vw, cudaStream_t st) {
    const int Cin = mode == 0 ? g->Cb : g->Ca, Cout = mode == 0 ? g->Ca : g->Cb;
    // channel windows: the input side is the TMA-loaded operand, the output side is written (and href read) by the epilogue
    int in_pitch = 0, in_valid = 0, out_pitch = Cout, out_valid = Cout;
    if (vw) {
        in_pitch = mode == 0 ? vw->L_pitch : vw->S_pitch; in_valid = mode == 0 ? vw->L_valid : vw->S_valid;
        const int op = mode == 0 ? vw->S_pitch : vw->L_pitch, ov = mode == 0 ? vw->S_valid : vw->L_valid;
        if (op > 0) out_pitch = op;
        if (ov > 0) out_valid = ov;
        if ((in_pitch | in_valid | out_pitch | out_valid) & 7) return 0;          // 16-byte granularity
        if (in_valid > Cin || out_valid > Cout || (in_pitch > 0 && in_pitch < (in_valid > 0 ? in_valid : Cin)) || out_pitch < out_valid) return 0;
    }
    if (Cin % KCH != 0) return 0;
    // 3x3 / stride 1 / same size with few output channels and a plain epilogue (dense-layer conv2): every input pixel once
    if (mode == 0 && bias == nullptr && (epi == SRGAN_EPI_BIAS_ACT ? act == SRGAN_ACT_NONE : (href == nullptr || act == SRGAN_ACT_NONE))) {
        const int took = flat3x3_conv(src, W, out, n, g, in_pitch, in_valid, out_pitch, out_valid, st);
        if (took != 0) return took;
    }
    int BN = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : 0));
    if (BN == 0) return 0;
    // Cout = 64 (mod 128) with three or more 64-wide tiles (the DenseNet trunk's 1x1 data gradients: every second concat
    // width): 256-wide tiles with a partial last one instead -- the weight rows beyond Cout are TMA zero fill and the
    // epilogue skips their 32-column chunks -- so the A operand is re-read from L2 Cout/256 instead of Cout/64 times
    if (BN == 64 && Cout >= 192) BN = 256;
    int Hm, Wm, phases = 1;
    if (mode == 0) { Hm = g->Hs; Wm = g->Ws; }
    else {
        if (g->Hl % g->stride || g->Wl % g->stride) return 0;
        if (g->R < g->stride || g->S < g->stride) return 0;
        Hm = g->Hl / g->stride; Wm = g->Wl / g->stride; phases = g->stride * g->stride;
    }
    if (g->stride > 8) return 0;
    UmmaConvParams p;
    if (!pick_patch(Wm, Hm, TILE_M, 16, p.TW, p.TH, p.TN)) return 0;
    if (mode == 0 && (p.TW * g->stride > 256 || p.TH * g->stride > 256)) return 0;
    if (((uintptr_t)src & 15) || ((uintptr_t)W & 15) || ((uintptr_t)out & 15) || (href && ((uintptr_t)href & 15))) return 0;
    if (bias && bias_mod != 1 && (((uintptr_t)bias & 15) || (bias_mod % 32) != 0)) return 0;   // epilogue reads bias as float4 runs of 32
    p.mode = mode; p.n = n; p.Hm = Hm; p.Wm = Wm;
    p.tiles_w = Wm / p.TW; p.tiles_h = Hm / p.TH;
    const int tiles_n = (n + p.TN - 1) / p.TN;
    p.Cin = Cin; p.Cout = Cout; p.R = g->R; p.S = g->S; p.stride = g->stride; p.pad = g->pad;
    p.Hout = mode == 0 ? g->Hs : g->Hl; p.Wout = mode == 0 ? g->Ws : g->Wl;
    p.bias = bias; p.bias_mod = bias_mod; p.href = (const bf16*)href; p.out = (bf16*)out;
    p.out_pitch = out_pitch; p.out_valid = out_valid;
    p.epi = epi; p.act = act; p.slope = slope;
    CUtensorMap tmA, tmB;
    int rc;
    if (mode == 0) rc = encode_act(&tmA, src, n, g->Hl, g->Wl, g->Cb, p.TW, p.TH, p.TN, g->stride, in_pitch, in_valid);
    else rc = encode_act(&tmA, src, n, g->Hs, g->Ws, g->Ca, p.TW, p.TH, p.TN, 1, in_pitch, in_valid);
    if (rc) return rc;
    rc = encode_mat(&tmB, W, Cout, (long long)g->R * g->S * Cin, BN);
    if (rc) return rc;
    long long mtiles = (long long)p.tiles_w * p.tiles_h * tiles_n;
    if (mtiles > 0x7fffffffLL) return 0;
    {
        UmmaConvParamsP pp;
        pp.c = p;
        pp.m_subtiles = (int)mtiles;
        pp.n_tiles = (Cout + BN - 1) / BN;
        // 256-row CTA tiles (two M sub-tiles share every weight tile) unless the wave quantisation on 148 persistent CTAs
        // makes 128-row tiles finish sooner: time ~ rounds x rows per tile
        int MT = (BN <= 128 && mtiles % 2 == 0) ? 2 : 1;
        if (MT == 2) {
            const long long t1 = (long long)phases * pp.n_tiles * mtiles, t2 = t1 / 2;
            const long long r1 = (t1 + kNumSMs - 1) / kNumSMs, r2 = (t2 + kNumSMs - 1) / kNumSMs;
            if (r1 < 2 * r2) MT = 1;
        }
        long long total = (long long)phases * pp.n_tiles * (mtiles / MT);
        if (total > 0x7fffffffLL) return 0;
        pp.total_tiles = (int)total;
        if (BN == 256) return launch_conv_persistent<1, 256>(tmA, tmB, pp, st);
        if (BN == 128) return MT == 2 ? launch_conv_persistent<2, 128>(tmA, tmB, pp, st) : launch_conv_persistent<1, 128>(tmA, tmB, pp, st);
        return MT == 2 ? launch_conv_persistent<2, 64>(tmA, tmB, pp, st) : launch_conv_persistent<1, 64>(tmA, tmB, pp, st);
    }
}

int umma_wgrad(const void* S, const void* L, float* dW, int n, const srgan_geom* g, const srgan_views* vw, cudaStream_t st) {
    if (g->Ca % 64 != 0 || g->Cb % 64 != 0) return 0;   // Ca = 64 (mod 128): the upper half tile is TMA zero fill
    if (vw && ((vw->S_pitch | vw->S_valid | vw->L_pitch | vw->L_valid) & 7)) return 0;
    if (vw && (vw->S_valid > g->Ca || vw->L_valid > g->Cb)) return 0;
    {
        // few real small-side channels, stride 1 (dense-layer conv2): the swapped formulation
        const int a_valid = (vw && vw->S_valid > 0) ? vw->S_valid : g->Ca;
        static const bool swap_on = [] { const char* e = getenv("SRGAN_NO_WGRAD_SWAP"); return !(e && e[0] == '1'); }();
        if (swap_on && a_valid <= WS_N && g->stride == 1 && g->R * g->S * WS_N <= 512 && g->Cb % 64 == 0 && g->Hs == g->Hl &&
            g->Ws == g->Wl && !(((uintptr_t)S | (uintptr_t)L | (uintptr_t)dW) & 15)) {
            UmmaWgradSwapParams p;
            if (pick_patch(g->Wl, g->Hl, WS_PIX, 8, p.TW, p.TH, p.TN)) {
                p.n = n; p.H = g->Hl; p.W = g->Wl; p.Ca_valid = a_valid; p.Cb = g->Cb; p.R = g->R; p.S = g->S; p.pad = g->pad;
                p.tiles_w = g->Wl / p.TW; p.tiles_h = g->Hl / p.TH; p.tiles_n = (n + p.TN - 1) / p.TN;
                p.dW = dW;
                const int total_chunks = p.tiles_w * p.tiles_h * p.tiles_n;
                const int m_tiles = (g->Cb + 127) / 128;
                int splits = kNumSMs / m_tiles;
                const int max_splits = (total_chunks + 3) / 4;             // at least 4 stages of work per CTA
                if (splits > max_splits) splits = max_splits;
                if (splits < 1) splits = 1;
                p.chunks_per_split = (total_chunks + splits - 1) / splits;
                splits = (total_chunks + p.chunks_per_split - 1) / p.chunks_per_split;
                CUtensorMap tmS, tmL;
                int rc = encode_act(&tmS, S, n, g->Hs, g->Ws, g->Ca, p.TW, p.TH, p.TN, 1, vw ? vw->S_pitch : 0, vw ? vw->S_valid : 0);
                if (rc) return rc;
                rc = encode_act(&tmL, L, n, g->Hl, g->Wl, g->Cb, p.TW, p.TH, p.TN, 1, vw ? vw->L_pitch : 0, vw ? vw->L_valid : 0);
                if (rc) return rc;
                const int stage_bytes = 2 * WS_PIX * 128 + g->R * g->S * WS_PIX * 128;
                p.stages = (200 * 1024) / stage_bytes;
                if (p.stages > 8) p.stages = 8;
                const size_t smem = (size_t)p.stages * stage_bytes + 1024;
                static srgan_per_device_once attr_set;
                if (attr_set.need()) {
                    cudaError_t e = cudaFuncSetAttribute(umma_wgrad_swap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
                    if (e != cudaSuccess) { srgan_set_error("cudaFuncSetAttribute(umma_wgrad_swap_kernel): %s", cudaGetErrorString(e)); return SRGAN_ERR_CUDA; }
                    attr_set.done();
                }
                umma_wgrad_swap_kernel<<<dim3(m_tiles, 1, splits), WG_THREADS, smem, st>>>(tmS, tmL, p);
                SRGAN_CHECK_LAUNCH("umma_wgrad_swap_kernel");
                return 1;
            }
        }
    }
    int BN = g->Cb % 256 == 0 ? 256 : (g->Cb % 128 == 0 ? 128 : 64);
    const int taps = g->R * g->S;
    int NT = 512 / BN;                                         // largest divisor of the tap count that fits TMEM (3 for 3x3)
    while (NT > 1 && taps % NT != 0) --NT;
    if (taps % NT != 0) return 0;
    // 1x1 (the DenseNet trunk: S = the 128-channel delta, L = up to 1920 input channels): TMEM columns that the taps do
    // not use can hold further b tiles (NB), so one S tile feeds all of them, and tiles may be wider than Cb's
    // factorisation allows (the tail beyond Cb is TMA zero fill, skipped in the epilogue).  Fewer output tiles t mean S
    // is re-read t times instead of Cb/64 times, but the K split over one wave of CTAs then adds 148/t partial sums into
    // dW with fp32 reductions: pick (BN, NB) by that traffic model (a reduced byte weighted 4x a read byte).
    int NB = 1;
    if (taps == 1) {
        const double rows = (double)n * g->Hs * g->Ws;
        const double s_bytes = rows * g->Ca * 2.0, dw_bytes = 4.0 * g->Ca * g->Cb;
        const int a_tiles = (g->Ca + 127) / 128;
        double best = -1.0;
        int best_bn = BN, best_nb = 1;
        for (int bn = 64; bn <= 256; bn *= 2)
            for (int nb = 1; nb * bn <= 512; ++nb) {
                if (nb > 1 && (nb - 1) * bn >= g->Cb) break;          // an entirely empty b tile
                if (bn == BN && nb == 1) {}                              // the exact-fit default is always a candidate
                else if (bn < BN) continue;                              // narrower than the exact fit never helps
                const int t = a_tiles * ((g->Cb + nb * bn - 1) / (nb * bn));
                const double chunks = rows / (bn == 64 ? 64 : 32);
                double splits = kNumSMs / t < 1 ? 1 : kNumSMs / t;
                if (splits > (chunks + 7) / 8) splits = (chunks + 7) / 8;
                if (splits < 1) splits = 1;
                const double cost = s_bytes * t + 4.0 * dw_bytes * splits;
                if (best < 0 || cost < best) { best = cost; best_bn = bn; best_nb = nb; }
            }
        BN = best_bn; NB = best_nb;
    }
    if (g->stride > 8) return 0;
    UmmaWgradParams p;
    const int WG_PIX = BN == 64 ? 64 : 32;
    if (!pick_patch(g->Ws, g->Hs, WG_PIX, 8, p.TW, p.TH, p.TN)) return 0;
    if (p.TW * g->stride > 256 || p.TH * g->stride > 256) return 0;
    if (((uintptr_t)S & 15) || ((uintptr_t)L & 15) || ((uintptr_t)dW & 15)) return 0;
    p.n = n; p.Hs = g->Hs; p.Ws = g->Ws; p.Ca = g->Ca; p.Cb = g->Cb; p.R = g->R; p.S = g->S; p.stride = g->stride; p.pad = g->pad;
    p.tiles_w = g->Ws / p.TW; p.tiles_h = g->Hs / p.TH; p.tiles_n = (n + p.TN - 1) / p.TN;
    p.NT = NT; p.NB = NB; p.dW = dW;
    const int total_chunks = p.tiles_w * p.tiles_h * p.tiles_n;
    const int out_tiles = ((g->Ca + 127) / 128) * (taps / NT) * ((g->Cb + NB * BN - 1) / (NB * BN));
    int splits = kNumSMs / out_tiles;                         // one CTA per SM (512 TMEM columns each): never a second wave
    int max_splits = (total_chunks + 7) / 8;                  // at least 8 stages of work per CTA
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.chunks_per_split = (total_chunks + splits - 1) / splits;
    splits = (total_chunks + p.chunks_per_split - 1) / p.chunks_per_split;
    CUtensorMap tmS, tmL;
    int rc = encode_act(&tmS, S, n, g->Hs, g->Ws, g->Ca, p.TW, p.TH, p.TN, 1, vw ? vw->S_pitch : 0, vw ? vw->S_valid : 0);
    if (rc) return rc;
    rc = encode_act(&tmL, L, n, g->Hl, g->Wl, g->Cb, p.TW, p.TH, p.TN, g->stride, vw ? vw->L_pitch : 0, vw ? vw->L_valid : 0);
    if (rc) return rc;
    dim3 grid(out_tiles, 1, splits);
    if (BN == 256) return launch_wgrad<256, 32>(tmS, tmL, p, grid, st);
    if (BN == 128) return launch_wgrad<128, 32>(tmS, tmL, p, grid, st);
    return launch_wgrad<64, 64>(tmS, tmL, p, grid, st);
}
