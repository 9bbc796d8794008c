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
constexpr int BD_EW = 8;                         // epilogue warps of bn_dgrad_kernel (16 = four per TMEM lane quarter was measured: no gain, 91 vs 90 us, and spills)
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(BD_EW * 32) : "memory"); }     // the epilogue warps

int bn_packed_scale() {
    static const int v = [] { const char* e = getenv("SRGAN_BN_PACKED_SCALE"); return e ? atoi(e) : 0; }();
    return v;
}

constexpr int BD_THREADS = 64 + BD_EW * 32;      // warp 0 TMA, warp 1 MMA + TMEM, warps 2.. epilogue
constexpr int BD_BN = 128;                       // output channels per tile
constexpr int BD_STAGE_BYTES = A_STAGE_BYTES + BD_BN * KCH * 2;
constexpr int BD_CS = BD_EW / 4;                 // chunk stride: warp ew owns the 32-column chunks (ew >> 2) + BD_CS * jj
constexpr int BD_NCH = BD_BN / 32 / BD_CS;       // 32-column chunks per epilogue warp and tile
constexpr int BD_ZONE_WARP = BD_NCH * 2 * EPI_STG_BYTES;      // per warp: chunks x {cat, dcat} x (4 row groups x 32 lanes x 16 B)

struct BnDgradParams {
    long long rows;               // GEMM rows (pixels x samples)
    int K;                        // reduction length = channels of dy (a multiple of 64)
    int C;                        // output channels that exist (the BatchNorm's channels; a multiple of 8)
    int Cpad;                     // C rounded up to 32
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
    int packed_s;                 // 0 (default): dx (+)= d * s with the fp32 scale, bit-identical to the unfused affine backward;
                                  // 1 (SRGAN_BN_PACKED_SCALE=1): one packed bf16 fma per channel pair with s rounded to bf16 --
                                  // 2 % faster on the crowd step, but the gradient-penalty error of the reduced crowd case
                                  // grows from 2.1e-2 to 3.4e-2 (profiles/r2_bf16_parity_errors.txt)
    int resb;                     // 1: the weight tile of the current channel tile (K/64 x 16 KB) stays RESIDENT in shared memory
                                  // and is re-loaded only when the CTA's tile range moves on to the next channel tile; the ring
                                  // stages then carry the dy tile only (GEMM variant, K <= 256)
    // CONV variant: dy is an NHWC activation [n, H, W, .] and the product a stride-1 transposed convolution over R x S taps
    // (the data gradient of a same-size convolution); a row tile is a TW x TH x TN patch (powers of two) of pixels x samples
    int n, H, W, R, S, pad, Cin;  // Cin = channels of dy per tap as the weight matrix counts them (a multiple of 64)
    int TW, TH, TN, lgTW, lgTH, tiles_w, tiles_h;
    int k32;                      // CONV: at most 32 channels of dy exist per tap (a dense layer's growth): the stage holds 64-byte
                                  // rows (SWIZZLE_64B boxes of 32 channels, two K steps) -- half the L2 -> SM bytes per tap
};

template <bool CONV>
__global__ void __launch_bounds__(BD_THREADS, 1) bn_dgrad_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB, const BnDgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ __align__(8) uint64_t b_full_bar, b_empty_bar;
    __shared__ uint32_t tmem_slot;

    constexpr int BN = BD_BN;
    constexpr int TMEM_COLS = 2 * BN;
    constexpr int B_TILE_BYTES = BN * KCH * 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stages = p.stages;
    const int nch = p.K / KCH;
    const bool resb = !CONV && p.resb != 0;
    // shared memory: [resident weight tiles] [ring] [epilogue buffers]
    const uint32_t b_res = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t tiles = b_res + (resb ? (uint32_t)(nch * B_TILE_BYTES) : 0u);
    const bool k32 = CONV && p.k32 != 0;
    const int stage_bytes = resb ? A_STAGE_BYTES : (k32 ? BD_STAGE_BYTES / 2 : BD_STAGE_BYTES);
    // a CTA walks a contiguous range of the tile list (row tiles fastest): its consecutive tiles cover the same channels
    const int tile_begin = (int)((long long)p.total_tiles * blockIdx.x / gridDim.x);
    const int tile_end = (int)((long long)p.total_tiles * (blockIdx.x + 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tmem_full_bar[b]), 1); mbar_init(smem_u32(&tmem_empty_bar[b]), BD_EW); }
        mbar_init(smem_u32(&b_full_bar), 1);
        mbar_init(smem_u32(&b_empty_bar), 1);
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
            int s = 0, ny_res = -1;
            uint32_t ph = 0, bph = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile) {
                const int mt = tile % p.m_tiles, ny = tile / p.m_tiles;
                if (resb && ny != ny_res) {
                    // next channel tile: its weights replace the resident ones once every MMA that reads those has completed
                    if (ny_res >= 0) { mbar_wait(smem_u32(&b_empty_bar), bph); bph ^= 1; }
                    const uint32_t bb = smem_u32(&b_full_bar);
                    mbar_expect_tx(bb, (uint32_t)(nch * B_TILE_BYTES));
                    for (int ch = 0; ch < nch; ++ch) tma_load_2d(b_res + (uint32_t)(ch * B_TILE_BYTES), &tmB, bb, ch * KCH, ny * BN);
                    ny_res = ny;
                }
                if constexpr (CONV) {
                    const int w0 = (mt % p.tiles_w) * p.TW, h0 = ((mt / p.tiles_w) % p.tiles_h) * p.TH;
                    const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.TN;
                    const int cpt = p.Cin / KCH;             // K chunks per tap
                    for (int tr = 0; tr < p.R; ++tr)
                        for (int ts = 0; ts < p.S; ++ts)
                            for (int ch = 0; ch < cpt; ++ch) {
                                mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
                                const uint32_t fb = smem_u32(&full_bar[s]);
                                mbar_expect_tx(fb, (uint32_t)stage_bytes);
                                const uint32_t dst = tiles + s * stage_bytes;
                                tma_load_4d(dst, &tmA, fb, ch * KCH, w0 + p.pad - ts, h0 + p.pad - tr, n0);
                                tma_load_2d(dst + (k32 ? A_STAGE_BYTES / 2 : A_STAGE_BYTES), &tmB, fb, (tr * p.S + ts) * p.Cin + ch * KCH, ny * BN);
                                if (++s == stages) { s = 0; ph ^= 1; }
                            }
                } else {
                    for (int ch = 0; ch < nch; ++ch) {
                        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
                        const uint32_t fb = smem_u32(&full_bar[s]);
                        mbar_expect_tx(fb, (uint32_t)stage_bytes);
                        const uint32_t dst = tiles + s * stage_bytes;
                        tma_load_2d(dst, &tmA, fb, ch * KCH, mt * TILE_M);
                        if (!resb) tma_load_2d(dst + A_STAGE_BYTES, &tmB, fb, ch * KCH, ny * BN);
                        if (++s == stages) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(TILE_M, BN, 0, 0);
            // 128-byte rows in 1024-byte swizzle atoms, or (k32) 64-byte rows in 512-byte atoms of SWIZZLE_64B
            const uint64_t desc0 = k32 ? ((make_desc(0, 16, 512) & ~((uint64_t)7 << 61)) | ((uint64_t)4 << 61)) : make_desc(0, 16, 1024);
            const int a_bytes = k32 ? A_STAGE_BYTES / 2 : A_STAGE_BYTES, ksteps = k32 ? 2 : KCH / 16;
            int s = 0, tl = 0, ny_res = -1;
            uint32_t ph = 0, bfph = 0;
            for (int tile = tile_begin; tile < tile_end; ++tile, ++tl) {
                const int buf = tl & 1;
                const uint32_t bph = (tl >> 1) & 1;
                const int ny = tile / p.m_tiles;
                mbar_wait(smem_u32(&tmem_empty_bar[buf]), bph ^ 1);
                if (resb && ny != ny_res) { mbar_wait(smem_u32(&b_full_bar), bfph); bfph ^= 1; ny_res = ny; }
                tc_fence_after();
                const uint32_t acc = tmem_base + buf * BN;
                for (int k_it = 0; k_it < nch; ++k_it) {
                    mbar_wait(smem_u32(&full_bar[s]), ph);
                    tc_fence_after();
                    const uint32_t a_s = tiles + s * stage_bytes;
                    const uint64_t ad0 = desc0 + (uint64_t)(a_s >> 4);
                    const uint64_t bd0 = desc0 + (uint64_t)((resb ? b_res + (uint32_t)(k_it * B_TILE_BYTES) : a_s + a_bytes) >> 4);
#pragma unroll 4
                    for (int k = 0; k < ksteps; ++k)
                        umma_f16(acc, ad0 + (uint64_t)(k * 2), bd0 + (uint64_t)(k * 2), idesc, (k_it > 0 || k > 0) ? 1u : 0u);
                    umma_commit(smem_u32(&empty_bar[s]));
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
                umma_commit(smem_u32(&tmem_full_bar[buf]));
                // last tile of this channel tile in the CTA's range: the resident weights may be replaced when these MMAs are done
                if (resb && tile + 1 < tile_end && (tile + 1) / p.m_tiles != ny) umma_commit(smem_u32(&b_empty_bar));
            }
        }
    } else {
        // ================= epilogue (8 warps) =================
        const int ew = warp - 2;
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int half = ew >> 2;                // first of the tile's 32-column chunks this warp owns (then every BD_CS-th)
        const int t_unit = lane & 3, t_row = lane >> 2;       // transposed role: 16-byte unit t_unit of rows 8*it + t_row
        const int te = threadIdx.x - 64;         // 0..255 among the epilogue threads
        const uint32_t after_ring = (uint32_t)(stages * stage_bytes);
        const uint32_t stg = tiles + after_ring + (uint32_t)ew * EPI_STG_BYTES;
        const uint32_t zone = tiles + after_ring + (uint32_t)(BD_EW * EPI_STG_BYTES) + (uint32_t)ew * BD_ZONE_WARP + (uint32_t)lane * 16u;
        // Per-channel tables, all bf16 so that the hot loop stays in packed arithmetic (the epilogue is bound by the number of
        // instructions its 8 warps execute, not by bytes):
        //   t_sg = sign of the scale s (+-1), t_T = the ReLU threshold in u = x * sign: bn(x) rounded to bf16 > 0  <=>  u >= T,
        //          found EXACTLY per channel below (the predicate is monotone in u), so the mask is still the stored
        //          activation's sign bit for bit but costs a packed multiply + compare per channel pair;
        //   t_sb = s rounded to bf16 (packed_s: dx (+)= d * s as one packed fma per channel pair); t_sf = s in fp32 (default).
        uint16_t* tab = reinterpret_cast<uint16_t*>(smem_raw + (tiles - smem_u32(smem_raw)) + after_ring + (uint32_t)(BD_EW * EPI_STG_BYTES) + (uint32_t)(BD_EW * BD_ZONE_WARP));
        // The tables hold the BN channels of the CURRENT channel tile only (a CTA's tile range is contiguous, row tiles fastest:
        // it crosses a channel-tile boundary a few times per launch at most) and are rebuilt there by the first 128 epilogue
        // threads -- filling them for all C channels up front cost 3-4 us of every launch (7 dependent parameter loads per thread).
        uint16_t* t_T = tab;
        uint16_t* t_sg = tab + BN;
        uint16_t* t_sb = tab + 2 * BN;
        float* t_sf = reinterpret_cast<float*>(tab + 3 * BN);          // s in fp32
        const bool grads = p.dgamma != nullptr;
        const bool accum = p.accumulate != 0;
        auto fill_tables = [&](int ny_t) {
            epi_bar_sync();                      // nobody still reads the previous channel tile's entries
            if (te < BN) {
            const int cl = te, c = ny_t * BN + te;
            uint16_t Tb = 0x7F80, sgb = 0x3F80, sbb = 0;          // +inf: never active; +1; 0
            if (c < p.C) {
                const float s = bn_scale_f(__ldg(p.gamma + c), __ldg(p.var + c), p.eps);
                const float t = bn_shift(__ldg(p.beta + c), __ldg(p.mean + c), s);
                sbb = __bfloat16_as_ushort(__float2bfloat16_rn(s));
                auto val = [](int m) { return __bfloat162float(__ushort_as_bfloat16((uint16_t)(m < 0 ? (0x8000 | (-m)) : m))); };
                const float sg = s < 0.f ? -1.f : 1.f;
                auto pred = [&](int m) { return __bfloat162float(__float2bfloat16_rn(bn_apply(val(m) * sg, s, t))) > 0.f; };
                if (s == 0.f) {
                    Tb = t > 0.f && __bfloat162float(__float2bfloat16_rn(t)) > 0.f ? 0xFF80 : 0x7F80;            // -inf / +inf
                } else {
                    if (s < 0.f) sgb = 0xBF80;
                    const float u0 = -t / fabsf(s);
                    int m;                         // bf16 values in monotone integer order: m = +-(magnitude bits)
                    if (!(u0 == u0)) m = 0x7F80;
                    else {
                        const uint16_t b = __bfloat16_as_ushort(__float2bfloat16_rn(u0));
                        m = (b & 0x8000) ? -(int)(b & 0x7FFF) : (int)(b & 0x7FFF);
                        if (m > 0x7F80) m = 0x7F80;
                        if (m < -0x7F80) m = -0x7F80;
                    }
                    for (int i = 0; i < 64 && m < 0x7F80 && !pred(m); ++i) ++m;
                    for (int i = 0; i < 64 && m > -0x7F80 && pred(m - 1); ++i) --m;
                    if (!pred(m)) m = 0x7F80;
                    Tb = (uint16_t)(m < 0 ? (0x8000 | (-m)) : m);
                }
            }
            t_T[cl] = Tb; t_sg[cl] = sgb; t_sb[cl] = sbb;
            t_sf[cl] = c < p.C ? bn_scale_f(__ldg(p.gamma + c), __ldg(p.var + c), p.eps) : 0.f;
            }
            epi_bar_sync();
        };

        // Per tile a lane touches 8 pieces ("slots" k = 4*jj + it: chunk half + 2*jj, rows 8*it + t_row of the warp's 32) of 16
        // bytes of cat and of dcat.  They travel through a lane-private ring of 8 landing slots per warp (32 lanes x 16
        // bytes x 2 streams each), one cp.async group per slot: a slot is refilled with the NEXT tile's piece right after
        // it has been consumed, so every piece is in flight for a whole tile time and ~all of the zone (64 KB per CTA) is
        // outstanding at any moment (Little: 44 GB/s per SM x ~1.5 us).
        // Geometry of a tile for this lane: element offset of slot 0, rows left below the lane's first row, channel validity.
        // (pix[it] = index of the pixel row that the lane's slot `it` belongs to, ok = bit it set when that row exists)
        struct Geo { long long off[4]; int pix[4]; int ok; int c0; };     // off = element offset of (pixel row, channel c0) in x / dx
        auto geo = [&](int mt, int ny) {
            Geo g;
            g.ok = 0;
            g.c0 = ny * BN + half * 32 + t_unit * 8;      // slot chunk jj adds BD_CS * 32 channels
            if (CONV) {
                const int w0 = (mt % p.tiles_w) * p.TW, h0 = ((mt / p.tiles_w) % p.tiles_h) * p.TH;
                const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.TN;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int r = q * 32 + 8 * it + t_row;
                    const int sample = n0 + (r >> (p.lgTW + p.lgTH));
                    g.pix[it] = (sample * p.H + h0 + ((r >> p.lgTW) & (p.TH - 1))) * p.W + w0 + (r & (p.TW - 1));
                    if (sample < p.n) g.ok |= 1 << it;
                }
            } else {
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    g.pix[it] = mt * TILE_M + q * 32 + 8 * it + t_row;
                    if (g.pix[it] < p.rows) g.ok |= 1 << it;
                }
            }
#pragma unroll
            for (int it = 0; it < 4; ++it) g.off[it] = (long long)g.pix[it] * p.pitch + g.c0;
            return g;
        };
        auto issue = [&](const Geo& g, bool live_tile, int k) {
            const int jj = k >> 2, it = k & 3;
            const bool ok = live_tile && ((g.ok >> it) & 1) && g.c0 + jj * (BD_CS * 32) < p.C;
            const long long o = ok ? g.off[it] + jj * (BD_CS * 32) : 0;
            cp_async16(zone + (uint32_t)(k * 1024), p.x + o, ok ? 16u : 0u);
            if (accum) cp_async16(zone + (uint32_t)(k * 1024 + 512), p.dx + o, ok ? 16u : 0u);
            cp_async_commit();                   // always: the group count per slot stays fixed
        };
        int mt = tile_begin % p.m_tiles, ny = tile_begin / p.m_tiles;
        Geo gc = geo(mt, ny);
#pragma unroll
        for (int k = 0; k < 4 * BD_NCH; ++k) issue(gc, tile_begin < tile_end, k);

        // partial sums of dgamma / dbeta stay in registers while consecutive tiles cover the same channels (a CTA walks a
        // contiguous range of tiles, row tiles fastest): [chunk][0..7] sum d * x, [chunk][8..15] sum d over the rows this
        // lane has seen; (sum d*x - mean * sum d) combined across the 8 lanes that share the channels and added to global
        // memory when the channel tile changes
        float acc[BD_NCH][16];
#pragma unroll
        for (int jj = 0; jj < BD_NCH; ++jj)
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[jj][e] = 0.f;
        auto flush = [&](int ny) {
#pragma unroll
            for (int jj = 0; jj < BD_NCH; ++jj) {
                float* a = acc[jj];
                const int cl = ny * BN + (half + BD_CS * jj) * 32 + t_unit * 8;
#pragma unroll
                for (int e = 0; e < 8; ++e) a[e] = fmaf(-(cl + e < p.C ? __ldg(p.mean + cl + e) : 0.f), a[8 + e], a[e]);
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
                const int c = cl + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2;
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
        int tl = 0, ny_tab = -1;
        for (int tile = tile_begin; tile < tile_end; ++tile, ++tl) {
            const int buf = tl & 1;
            const uint32_t bph = (tl >> 1) & 1;
            if (ny != ny_tab) { fill_tables(ny); ny_tab = ny; }
            // the next tile of this CTA's range (row tiles fastest)
            int mt_n = mt + 1, ny_n = ny;
            if (mt_n == p.m_tiles) { mt_n = 0; ++ny_n; }
            const bool live_n = tile + 1 < tile_end;
            const Geo gn = geo(mt_n, ny_n);
            mbar_wait(smem_u32(&tmem_full_bar[buf]), bph);
            tc_fence_after();
#pragma unroll
            for (int jj = 0; jj < BD_NCH; ++jj) {
                const int j = half + BD_CS * jj;
                const bool live = ny * BN + j * 32 < p.C;   // warp-uniform: chunks beyond the last channel (partial last N tile) carry no data
                const bool cok = gc.c0 + jj * (BD_CS * 32) < p.C;
                uint4 T8 = make_uint4(0u, 0u, 0u, 0u), sg8 = T8, sb8 = T8;
                float s8[8];
                if (live) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + j * 32, v);
                    const int ct = cok ? gc.c0 + jj * (BD_CS * 32) - ny * BN : 0;     // index within the channel tile
                    T8 = *reinterpret_cast<const uint4*>(t_T + ct);
                    sg8 = *reinterpret_cast<const uint4*>(t_sg + ct);
                    if (p.packed_s) sb8 = *reinterpret_cast<const uint4*>(t_sb + ct);
                    else {
                        const float4 a = *reinterpret_cast<const float4*>(t_sf + ct), b = *reinterpret_cast<const float4*>(t_sf + ct + 4);
                        s8[0] = a.x; s8[1] = a.y; s8[2] = a.z; s8[3] = a.w; s8[4] = b.x; s8[5] = b.y; s8[6] = b.z; s8[7] = b.w;
                    }
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
                }
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int k = jj * 4 + it;
                    cp_async_wait_group<4 * BD_NCH - 1>();      // the oldest slot = this one has landed (a thread reads back its own copies)
                    if (live) {
                        uint4 dr = lds128(stg_addr(stg, 8 * it + t_row, t_unit));
                        const uint4 xr = lds128(zone + (uint32_t)(k * 1024));
                        uint4 orw = make_uint4(0u, 0u, 0u, 0u);
                        if (accum) orw = lds128(zone + (uint32_t)(k * 1024 + 512));
                        // ReLU mask ANDed into the packed delta, then dx (+)= d * s: packed bf16 pairs throughout
                        {
                            uint32_t* dw = reinterpret_cast<uint32_t*>(&dr);
                            const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&xr);
                            const __nv_bfloat162* T2 = reinterpret_cast<const __nv_bfloat162*>(&T8);
                            const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&sg8);
                            const __nv_bfloat162* s2 = reinterpret_cast<const __nv_bfloat162*>(&sb8);
                            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&orw);
#pragma unroll
                            for (int h = 0; h < 4; ++h) dw[h] &= __hge2_mask(__hmul2(x2[h], g2[h]), T2[h]);
                            if (p.packed_s) {
#pragma unroll
                                for (int h = 0; h < 4; ++h) {
                                    const __nv_bfloat162 d2 = *reinterpret_cast<const __nv_bfloat162*>(&dw[h]);
                                    o2[h] = accum ? __hfma2(d2, s2[h], o2[h]) : __hmul2(d2, s2[h]);
                                }
                            }
                        }
                        if (grads || !p.packed_s) {
                            float d[8];
                            unpack8(dr, d);
                            if (!p.packed_s) {
                                float o[8];
                                unpack8(orw, o);
#pragma unroll
                                for (int e = 0; e < 8; ++e) o[e] = fmaf(d[e], s8[e], o[e]);
                                orw = pack8(o);
                            }
                            if (grads) {
                                float xv[8];
                                unpack8(xr, xv);
#pragma unroll
                                for (int e = 0; e < 8; ++e) { acc[jj][e] = fmaf(d[e], xv[e], acc[jj][e]); acc[jj][8 + e] += d[e]; }
                            }
                        }
                        if (cok && ((gc.ok >> it) & 1)) {
                            *reinterpret_cast<uint4*>(p.dx + gc.off[it] + jj * (BD_CS * 32)) = orw;
                            if (p.d_out != nullptr)
                                *reinterpret_cast<uint4*>(p.d_out + (long long)gc.pix[it] * p.d_pitch + gc.c0 + jj * (BD_CS * 32)) = dr;
                        }
                    }
                    issue(gn, live_n, k);
                }
                __syncwarp();                    // the transposition buffer is rewritten by the next chunk
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[buf]));
            if (grads && live_n && ny_n != ny) flush(ny);
            mt = mt_n; ny = ny_n; gc = gn;
        }
        cp_async_wait_all();
        if (grads && tile_end > tile_begin) flush(ny - (mt == 0 ? 1 : 0));
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}


// ------------------------------------------------------------------------------------------------------------
// bn_conv_down_kernel   out[r, :] = relu(bn(x[r, :C])) . Wd^T      (norm1 -> relu1 -> conv1 forward, crowd/models.py:339-341)
//   The A operand is the RAW concat buffer: every K chunk (128 rows x 64 channels) is TMA-loaded into the swizzled ring
//   stage, four transform warps apply the BatchNorm affine + ReLU to it IN PLACE in shared memory (generic-proxy writes,
//   fence.proxy.async, then a second mbarrier releases the stage to the MMA thread), so the normalised activation n1 never
//   has to exist in HBM: 1 C-wide stream per pixel instead of 3 (affine: read cat, write n1 | GEMM: read n1).  When a later
//   consumer still wants n1 (the weight-gradient GEMM and the tangent pass of the rows they cover), the transform threads
//   also store their transformed 16-byte units to n1_out (2 C-wide streams).
//   Optional second BatchNorm + ReLU in the epilogue (norm2 -> relu2 on the 128 bottleneck channels): out2 = relu(bn2(out)).
//   warp 0 TMA, warp 1 MMA + TMEM, warps 2..9 epilogue, warps 10.. transform (BF_NG groups of 4).
// ------------------------------------------------------------------------------------------------------------
constexpr int BF_NG = 2;                         // transform groups of 4 warps; group g owns the K chunks c = g (mod BF_NG)
constexpr int BF_THREADS = 320 + 128 * BF_NG;
constexpr int BF_BN = 128;

struct BnFpropParams {
    long long rows;
    int C;                        // BatchNorm channels = real K
    int Kpad;                     // K rounded up to 64 (the weight matrix's row length; rows of Wd beyond C are zero)
    int pitch;                    // elements between rows of x
    int Cout;                     // output channels that exist (a multiple of 8)
    int out_pitch;
    const float *gamma, *beta, *mean, *var;
    float eps;
    bf16* out;
    bf16* n1_out;                 // nullptr, or: the transformed operand, rows of n1_pitch elements (columns < Kpad written),
    int n1_pitch;                 // stored for the GEMM rows >= n1_first_row only
    long long n1_first_row;
    const float *gamma2, *beta2, *mean2, *var2;      // nullptr, or the BatchNorm that follows the convolution ...
    bf16* out2;                   // ... and where relu(bn2(out)) goes (same pitch as out); channels >= C2 get zeros
    int C2;
    int m_tiles, n_tiles, total_tiles, stages;
    int dbg_noxf;                 // SRGAN_BF_NOXF=1 (measurement only): the transform groups pass the stages on untouched
};

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int MT>
__global__ void __launch_bounds__(BF_THREADS, 1) bn_conv_down_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmB,
                                                                    const BnFpropParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];
    __shared__ __align__(8) uint64_t ready_bar[8];
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_slot;

    constexpr int BN = BF_BN;
    constexpr int STAGE_BYTES = MT * A_STAGE_BYTES + BN * KCH * 2;
    constexpr int ACC_COLS = MT * BN;
    constexpr int TMEM_COLS = 2 * ACC_COLS;
    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const tiles_g = smem_raw + (tiles - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stages = p.stages;
    const int nch = p.Kpad / KCH;
    float* const t_s = reinterpret_cast<float*>(tiles_g + (size_t)stages * STAGE_BYTES + 8 * EPI_STG_BYTES);
    float* const t_t = t_s + p.Kpad;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&ready_bar[s]), 4); mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tmem_full_bar[b]), 1); mbar_init(smem_u32(&tmem_empty_bar[b]), 8); }
        fence_barrier_init();
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), TMEM_COLS);
    // per-channel tables of the fused BatchNorm (channels >= C: scale = shift = 0, so the zero-filled K tail stays zero)
#pragma unroll 4
    for (int c = threadIdx.x; c < p.Kpad; c += BF_THREADS) {
        float s = 0.f, t = 0.f;
        if (c < p.C) { s = bn_scale_f(__ldg(p.gamma + c), __ldg(p.var + c), p.eps); t = bn_shift(__ldg(p.beta + c), __ldg(p.mean + c), s); }
        t_s[c] = s; t_t[c] = t;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int ny = tile % p.n_tiles, mt = tile / p.n_tiles;
                for (int ch = 0; ch < nch; ++ch) {
                    mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    mbar_expect_tx(fb, STAGE_BYTES);
                    const uint32_t dst = tiles + s * STAGE_BYTES;
#pragma unroll
                    for (int i = 0; i < MT; ++i) tma_load_2d(dst + i * A_STAGE_BYTES, &tmA, fb, ch * KCH, (mt * MT + i) * TILE_M);
                    tma_load_2d(dst + MT * A_STAGE_BYTES, &tmB, fb, ch * KCH, ny * BN);
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
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl) {
                const int buf = tl & 1;
                const uint32_t bph = (tl >> 1) & 1;
                mbar_wait(smem_u32(&tmem_empty_bar[buf]), bph ^ 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + buf * ACC_COLS;
                for (int k_it = 0; k_it < nch; ++k_it) {
                    mbar_wait(smem_u32(&ready_bar[s]), ph);          // transformed (implies landed)
                    tc_fence_after();
                    const uint32_t a_s = tiles + s * STAGE_BYTES;
                    const uint64_t ad0 = desc0 + (uint64_t)(a_s >> 4);
                    const uint64_t bd0 = desc0 + (uint64_t)((a_s + MT * A_STAGE_BYTES) >> 4);
#pragma unroll
                    for (int i = 0; i < MT; ++i)
#pragma unroll
                        for (int k = 0; k < KCH / 16; ++k)
                            umma_f16(acc + i * BN, ad0 + (uint64_t)(i * (A_STAGE_BYTES >> 4) + k * 2), bd0 + (uint64_t)(k * 2), idesc,
                                     (k_it > 0 || k > 0) ? 1u : 0u);
                    umma_commit(smem_u32(&empty_bar[s]));
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
                umma_commit(smem_u32(&tmem_full_bar[buf]));
            }
        }
    } else if (warp >= 10) {
        // ================= transform (BF_NG x 4 warps): BatchNorm affine + ReLU on the landed A tile, in place =================
        // thread -> 16-byte unit u (8 channels) of rows r0 + 16*i: the unit's swizzled position (u ^ (row & 7)) is the same
        // for all of its rows, the eight lanes of a row cover its 128 bytes (bank-conflict free)
        // (consecutive chunks are transformed concurrently by different groups: the per-chunk chain wait -> LDS -> math -> STS ->
        // proxy fence -> arrive is latency-bound for a single group)
        const int tt = (threadIdx.x - 320) & 127, grp = (threadIdx.x - 320) >> 7;
        const int u = tt & 7, r0 = tt >> 3;
        const uint32_t toff = (uint32_t)(r0 * 128 + ((u ^ (r0 & 7)) << 4));
        const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
        int s = 0, turn = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int ny = tile % p.n_tiles, mt = tile / p.n_tiles;
            const bool keep = p.n1_out != nullptr && ny == 0 && (long long)(mt + 1) * MT * TILE_M > p.n1_first_row;
            for (int ch = 0; ch < nch; ++ch) {
                const bool mine = turn == grp;
                if (++turn == BF_NG) turn = 0;
                if (!mine) {
                    if (++s == stages) { s = 0; ph ^= 1; }
                    continue;
                }
                const int cb = ch * KCH + u * 8;
                float s8[8], t8[8];
                {
                    const float4* ps = reinterpret_cast<const float4*>(t_s + cb);
                    const float4* pt = reinterpret_cast<const float4*>(t_t + cb);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 a = ps[h], b = pt[h];
                        s8[4 * h] = a.x; s8[4 * h + 1] = a.y; s8[4 * h + 2] = a.z; s8[4 * h + 3] = a.w;
                        t8[4 * h] = b.x; t8[4 * h + 1] = b.y; t8[4 * h + 2] = b.z; t8[4 * h + 3] = b.w;
                    }
                }
                mbar_wait(smem_u32(&full_bar[s]), ph);
                const uint32_t base = tiles + s * STAGE_BYTES + toff;
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    if (p.dbg_noxf) break;
                    const long long gr0 = (long long)(mt * MT + i) * TILE_M + r0;
                    bf16* const n1p = keep ? p.n1_out + gr0 * p.n1_pitch + cb : nullptr;
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {
                        uint4 raw[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) raw[k] = lds128(base + (uint32_t)(i * A_STAGE_BYTES + (hb * 4 + k) * 2048));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            float xv[8];
                            unpack8(raw[k], xv);
                            uint4 w;
                            uint32_t* ww = reinterpret_cast<uint32_t*>(&w);
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                const __nv_bfloat162 y = __hmax2(__floats2bfloat162_rn(bn_apply(xv[2 * h], s8[2 * h], t8[2 * h]),
                                                                                       bn_apply(xv[2 * h + 1], s8[2 * h + 1], t8[2 * h + 1])), zero2);
                                ww[h] = *reinterpret_cast<const uint32_t*>(&y);
                            }
                            sts128(base + (uint32_t)(i * A_STAGE_BYTES + (hb * 4 + k) * 2048), w);
                            if (keep && gr0 + 16 * (hb * 4 + k) < p.rows && gr0 + 16 * (hb * 4 + k) >= p.n1_first_row && cb < p.n1_pitch)
                                *reinterpret_cast<uint4*>(n1p + (long long)(16 * (hb * 4 + k)) * p.n1_pitch) = w;
                        }
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ready_bar[s]));
                if (++s == stages) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ================= epilogue (8 warps) =================
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;
        const int t_unit = lane & 3, t_row = lane >> 2;
        const uint32_t stg = tiles + (uint32_t)stages * STAGE_BYTES + (uint32_t)ew * EPI_STG_BYTES;
        const bool bn2 = p.out2 != nullptr;
        int tl = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl) {
            const int ny = tile % p.n_tiles, mt = tile / p.n_tiles;
            const int buf = tl & 1;
            const uint32_t bph = (tl >> 1) & 1;
            mbar_wait(smem_u32(&tmem_full_bar[buf]), bph);
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < MT; ++i) {
#pragma unroll
                for (int jj = 0; jj < BN / 64; ++jj) {
                    const int j = half + 2 * jj;
                    const int cbase = ny * BN + j * 32;
                    if (cbase >= p.Cout) continue;              // warp-uniform
                    const int cb = cbase + t_unit * 8;
                    const bool cok = cb < p.Cout;
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * ACC_COLS + i * BN + j * 32, v);
                    float s8[8], t8[8];
                    if (bn2) {
                        const int ct = cok ? cb : 0;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            s8[e] = 0.f; t8[e] = 0.f;
                            if (ct + e < p.C2) {
                                s8[e] = bn_scale_f(__ldg(p.gamma2 + ct + e), __ldg(p.var2 + ct + e), p.eps);
                                t8[e] = bn_shift(__ldg(p.beta2 + ct + e), __ldg(p.mean2 + ct + e), s8[e]);
                            }
                        }
                    }
                    tmem_ld_wait();
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
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const uint4 x = lds128(stg_addr(stg, 8 * it + t_row, t_unit));
                        const long long gr = (long long)(mt * MT + i) * TILE_M + q * 32 + 8 * it + t_row;
                        if (cok && gr < p.rows) {
                            *reinterpret_cast<uint4*>(p.out + gr * p.out_pitch + cb) = x;
                            if (bn2) {
                                float xv[8];
                                unpack8(x, xv);
#pragma unroll
                                for (int e = 0; e < 8; ++e) xv[e] = fmaxf(bn_apply(xv[e], s8[e], t8[e]), 0.f);
                                *reinterpret_cast<uint4*>(p.out2 + gr * p.out_pitch + cb) = pack8(xv);
                            }
                        }
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

// [rows x cols] bf16 matrix whose rows are `pitch` elements apart (a channel window of a wider buffer) as a 2-D map
int encode_mat_pitch(CUtensorMap* tm, const void* base, long long rows, long long cols, long long pitch, int brows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { srgan_set_error("cuTensorMapEncodeTiled is not available from the driver"); return SRGAN_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)brows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        srgan_set_error("cuTensorMapEncodeTiled(matrix %lldx%lld pitch %lld box %d) failed: %d", rows, cols, pitch, brows, (int)r);
        return SRGAN_ERR_CUDA;
    }
    return SRGAN_OK;
}

// 32-channel boxes with SWIZZLE_64B (64-byte rows): an NHWC activation window whose first `valid` <= 32 channels exist ...
int encode_act32(CUtensorMap* tm, const void* base, int n, int H, int W, int bw, int bh, int bn, int pitch, int valid) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { srgan_set_error("cuTensorMapEncodeTiled is not available from the driver"); return SRGAN_ERR_CUDA; }
    const cuuint64_t P = pitch;
    cuuint64_t dims[4] = {(cuuint64_t)valid, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    cuuint64_t strides[3] = {P * 2, (cuuint64_t)W * P * 2, (cuuint64_t)H * W * P * 2};
    cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { srgan_set_error("cuTensorMapEncodeTiled(activation32) failed: %d", (int)r); return SRGAN_ERR_CUDA; }
    return SRGAN_OK;
}
// ... and the matching weight boxes: 32 of every tap's 64 K columns, brows rows
int encode_mat32(CUtensorMap* tm, const void* base, long long rows, long long cols, int brows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { srgan_set_error("cuTensorMapEncodeTiled is not available from the driver"); return SRGAN_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {32, (cuuint32_t)brows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { srgan_set_error("cuTensorMapEncodeTiled(matrix32) failed: %d", (int)r); return SRGAN_ERR_CUDA; }
    return SRGAN_OK;
}

template <int MT>
int launch_bn_conv_down(const CUtensorMap& tmA, const CUtensorMap& tmB, BnFpropParams& p, cudaStream_t st) {
    constexpr int stage_bytes = MT * A_STAGE_BYTES + BF_BN * KCH * 2;
    const int fixed = 8 * EPI_STG_BYTES + 2 * p.Kpad * 4 + 1024;
    int stages = (220 * 1024 - fixed) / stage_bytes;
    if (stages > 8) stages = 8;
    static const int dbg_stages = [] { const char* e = getenv("SRGAN_BF_STAGES"); return e ? atoi(e) : 0; }();
    static const int dbg_noxf = [] { const char* e = getenv("SRGAN_BF_NOXF"); return e ? atoi(e) : 0; }();
    if (dbg_stages > 0 && dbg_stages < stages) stages = dbg_stages;
    p.dbg_noxf = dbg_noxf;
    // a multiple of the number of transform groups: every ring stage then belongs to ONE group, which sees all of that
    // stage's mbarrier phases in order (a group that skipped a phase could pass a parity wait one phase early)
    stages -= stages % BF_NG;
    if (stages < 2) return 0;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + fixed;
    static srgan_per_device_once attr_set;
    if (attr_set.need()) {
        cudaError_t e = cudaFuncSetAttribute(bn_conv_down_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) { srgan_set_error("cudaFuncSetAttribute(bn_conv_down_kernel): %s", cudaGetErrorString(e)); return SRGAN_ERR_CUDA; }
        attr_set.done();
    }
    const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
    bn_conv_down_kernel<MT><<<grid, BF_THREADS, smem, st>>>(tmA, tmB, p);
    SRGAN_CHECK_LAUNCH("bn_conv_down_kernel");
    return 1;
}


// ------------------------------------------------------------------------------------------------------------
// bn_conv_wgrad_kernel   dW[a, b] += sum_r dy[r, a] * relu(bn(x[r, b]))      (weight gradient of conv1 from the RAW concat buffer)
//   M = 128 channels a of dy (A operand MN-major: two [WG pixels x 64 channels] boxes), N = BN channels b per sub-problem
//   (B operand MN-major), NS sub-problems side by side in TMEM (NS * BN <= 512 columns) so that one dy tile feeds all of
//   them, K = pixels, split across one wave of CTAs; fp32 reductions into dW.  The x boxes are normalised + rectified in
//   place in shared memory by the transform groups (same scheme as bn_conv_down_kernel; a thread's 16 units per stage all
//   belong to the same 8 channels, so its scale / shift live in registers for the whole kernel).
//   warp 0 TMA, warp 1 MMA + TMEM, warps 2.. transform (BW_NG groups of 4 warps); warps 2..5 also drain TMEM at the end.
// ------------------------------------------------------------------------------------------------------------
constexpr int BW_NG = 2;
constexpr int BW_THREADS = 64 + 128 * BW_NG;
constexpr int BW_PIX = 32;                       // pixels (K) per stage
constexpr int BW_BOX = BW_PIX * 128;             // one [32 pixels x 64 channels] box: 4 KB

struct BnWgradParams {
    long long rows;
    int C, Kpad;                  // BatchNorm channels (real b), dW row length
    int BN, NS;                   // channels per sub-problem (64 / 128 / 256), sub-problems per CTA
    int chunks_per_split, total_chunks;
    const float *gamma, *beta, *mean, *var;
    float eps;
    float* dW;                    // [Ca][Kpad] fp32
    int Ca;                       // channels of dy that exist (128 per blockIdx.y)
    int stages;
};

__global__ void __launch_bounds__(BW_THREADS, 1) bn_conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmS,
                                                                     const __grid_constant__ CUtensorMap tmL,
                                                                     const BnWgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];
    __shared__ __align__(8) uint64_t ready_bar[8];
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_slot;

    constexpr int A_BYTES = 2 * BW_BOX;
    const int nbox = p.NS * (p.BN / 64);                 // x boxes per stage (<= 8)
    const int STAGE_BYTES = A_BYTES + nbox * BW_BOX;
    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stages = p.stages;
    const int b0 = blockIdx.x * p.NS * p.BN;
    const int a0 = blockIdx.y * 128;
    const int ch_begin = blockIdx.z * p.chunks_per_split;
    int ch_end = ch_begin + p.chunks_per_split;
    if (ch_end > p.total_chunks) ch_end = p.total_chunks;
    const int n_iters = ch_end - ch_begin;
    const int used = p.NS * p.BN;
    const uint32_t ncols = used <= 32 ? 32 : (used <= 64 ? 64 : (used <= 128 ? 128 : (used <= 256 ? 256 : 512)));

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&ready_bar[s]), 4); mbar_init(smem_u32(&empty_bar[s]), 1);
        }
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
                    const int r = (ch_begin + it) * BW_PIX;
                    mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    mbar_expect_tx(fb, STAGE_BYTES);
                    const uint32_t dst = tiles + s * STAGE_BYTES;
                    tma_load_2d(dst, &tmS, fb, a0, r);
                    tma_load_2d(dst + BW_BOX, &tmS, fb, a0 + 64, r);
                    for (int k = 0; k < nbox; ++k) tma_load_2d(dst + A_BYTES + k * BW_BOX, &tmL, fb, b0 + k * 64, r);
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            }
        } else if (warp == 1) {
            if (elect_one()) {
                const uint32_t idesc = make_idesc(128, p.BN, 1, 1);
                const uint64_t desc0 = make_desc(0, BW_BOX, 1024);
                const int sub_bytes = (p.BN / 64) * BW_BOX;
                int s = 0;
                uint32_t ph = 0;
                for (int it = 0; it < n_iters; ++it) {
                    mbar_wait(smem_u32(&ready_bar[s]), ph);
                    tc_fence_after();
                    const uint32_t a_s = tiles + s * STAGE_BYTES;
                    const uint64_t ad0 = desc0 + (uint64_t)(a_s >> 4);
                    for (int k = 0; k < p.NS; ++k) {
                        const uint64_t bd0 = desc0 + (uint64_t)((a_s + A_BYTES + k * sub_bytes) >> 4);
#pragma unroll
                        for (int kk = 0; kk < BW_PIX / 16; ++kk)      // 16 K rows (pixels) = 2048 B further into every column group
                            umma_f16(tmem_base + k * p.BN, ad0 + (uint64_t)(kk * 128), bd0 + (uint64_t)(kk * 128), idesc,
                                     (it > 0 || kk > 0) ? 1u : 0u);
                    }
                    umma_commit(smem_u32(&empty_bar[s]));
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
                umma_commit(smem_u32(&tmem_full_bar));
            }
        } else {
            // ================= transform: thread -> box bx, 16-byte unit u, rows rh*16 .. rh*16+15 =================
            {
                const int tt = (threadIdx.x - 64) & 127, grp = (threadIdx.x - 64) >> 7;
                const int u = tt & 7, rh = (tt >> 3) & 1, bx = tt >> 4;
                const int cb = b0 + bx * 64 + u * 8;
                float s8[8], t8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    s8[e] = 0.f; t8[e] = 0.f;
                    if (cb + e < p.C) {
                        s8[e] = bn_scale_f(__ldg(p.gamma + cb + e), __ldg(p.var + cb + e), p.eps);
                        t8[e] = bn_shift(__ldg(p.beta + cb + e), __ldg(p.mean + cb + e), s8[e]);
                    }
                }
                const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
                const bool active = bx < nbox;
                int s = 0, turn = 0;
                uint32_t ph = 0;
                for (int it = 0; it < n_iters; ++it) {
                    const bool mine = turn == grp;
                    if (++turn == BW_NG) turn = 0;
                    if (mine) {
                        mbar_wait(smem_u32(&full_bar[s]), ph);
                        if (active) {
                            const uint32_t base = tiles + s * STAGE_BYTES + A_BYTES + bx * BW_BOX + rh * 2048;
#pragma unroll
                            for (int hb = 0; hb < 4; ++hb) {
                                uint4 raw[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const int r = hb * 4 + k;                  // row within the half box; swizzle by the row's low 3 bits
                                    raw[k] = lds128(base + (uint32_t)(r * 128 + ((u ^ (r & 7)) << 4)));
                                }
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const int r = hb * 4 + k;
                                    float xv[8];
                                    unpack8(raw[k], xv);
                                    uint4 w;
                                    uint32_t* ww = reinterpret_cast<uint32_t*>(&w);
#pragma unroll
                                    for (int h = 0; h < 4; ++h) {
                                        const __nv_bfloat162 y = __hmax2(__floats2bfloat162_rn(bn_apply(xv[2 * h], s8[2 * h], t8[2 * h]),
                                                                                               bn_apply(xv[2 * h + 1], s8[2 * h + 1], t8[2 * h + 1])), zero2);
                                        ww[h] = *reinterpret_cast<const uint32_t*>(&y);
                                    }
                                    sts128(base + (uint32_t)(r * 128 + ((u ^ (r & 7)) << 4)), w);
                                }
                            }
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&ready_bar[s]));
                    }
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            }
            if (warp < 6) {
                // ================= drain: TMEM -> fp32 reductions into dW =================
                const int q = warp & 3;
                const int a = a0 + q * 32 + lane;                 // this thread's output row (channel a of dy)
                mbar_wait(smem_u32(&tmem_full_bar), 0);
                tc_fence_after();
#pragma unroll 1
                for (int j = 0; j < used / 32; ++j) {
                    const int bcol = b0 + j * 32;
                    if (bcol >= p.Kpad) break;                    // warp-uniform: zero-filled tail of the last b group
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + j * 32, v);
                    tmem_ld_wait();
                    if (a >= p.Ca) continue;
                    float* dst = p.dW + (long long)a * p.Kpad + bcol;
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
    p.packed_s = bn_packed_scale();
    const int fixed = BD_EW * EPI_STG_BYTES + BD_EW * BD_ZONE_WARP + 3 * BD_BN * 2 + BD_BN * 4 + 1024;
    static const bool resb_on = [] { const char* e = getenv("SRGAN_NO_RESIDENT_B"); return !(e && e[0] == '1'); }();
    p.resb = (resb_on && K <= 256) ? 1 : 0;
    const int b_bytes = p.resb ? (K / KCH) * BD_BN * KCH * 2 : 0;
    const int stage_bytes = p.resb ? A_STAGE_BYTES : BD_STAGE_BYTES;
    int stages = (226 * 1024 - fixed - b_bytes) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) return 0;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + b_bytes + fixed;
    CUtensorMap tmA, tmB;
    int rc = encode_mat(&tmA, dy, rows, K, TILE_M);
    if (rc) return rc;
    rc = encode_mat(&tmB, Wu, Cout, K, BD_BN);
    if (rc) return rc;
    p.k32 = 0;
    p.n = p.H = p.W = p.R = p.S = p.pad = p.Cin = p.TW = p.TH = p.TN = p.lgTW = p.lgTH = p.tiles_w = p.tiles_h = 0;
    static srgan_per_device_once attr_set;
    if (attr_set.need()) {
        cudaError_t e = cudaFuncSetAttribute(bn_dgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) { srgan_set_error("cudaFuncSetAttribute(bn_dgrad_kernel): %s", cudaGetErrorString(e)); return SRGAN_ERR_CUDA; }
        attr_set.done();
    }
    const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
    bn_dgrad_kernel<false><<<grid, BD_THREADS, smem, st>>>(tmA, tmB, p);
    SRGAN_CHECK_LAUNCH("bn_dgrad_kernel");
    return 1;
}

// The same with dy an NHWC activation and the product a stride-1 R x S transposed convolution (the data gradient of a same-size
// convolution: conv2 of a dense layer, whose input is relu2(norm2(.))).  dy may be a channel window (dy_pitch elements between
// pixels, dy_valid channels exist, the rest of the Cin channels per tap is zero fill).  returns 1 / 0 / <0 as above
int bn_conv_dgrad(const void* dy, int dy_pitch, int dy_valid, const void* Wu, void* dx, const void* x, int n, int H, int W, int R,
                  int S, int pad, int Cin, int Cout, int C, int pitch, const float* gamma, const float* beta, const float* mean,
                  const float* var, float eps, float* dgamma, float* dbeta, void* d_out, int d_pitch, int accumulate, cudaStream_t st) {
    if (Cin % KCH != 0 || Cout % 64 != 0 || C > Cout || C <= 0 || (C & 7) || (pitch & 7) || pitch < C || n <= 0) return 0;
    if ((dy_pitch | dy_valid) & 7) return 0;
    if (dy_valid > Cin || (dy_pitch > 0 && dy_pitch < (dy_valid > 0 ? dy_valid : Cin))) return 0;
    if (((uintptr_t)dy | (uintptr_t)Wu | (uintptr_t)dx | (uintptr_t)x | (uintptr_t)d_out) & 15) return 0;
    if (d_out != nullptr && ((d_pitch & 7) || d_pitch < C)) return 0;
    if (R <= 0 || S <= 0 || pad < 0 || pad >= R || pad >= S || 2 * pad + 1 != R || 2 * pad + 1 != S) return 0;     // same-size, stride 1
    BnDgradParams p;
    if (!pick_patch(W, H, TILE_M, 16, p.TW, p.TH, p.TN)) return 0;
    if ((p.TW & (p.TW - 1)) || (p.TH & (p.TH - 1))) return 0;
    p.lgTW = 0; while ((1 << p.lgTW) < p.TW) ++p.lgTW;
    p.lgTH = 0; while ((1 << p.lgTH) < p.TH) ++p.lgTH;
    p.tiles_w = W / p.TW; p.tiles_h = H / p.TH;
    const long long tiles_n = (n + p.TN - 1) / p.TN;
    const long long m_tiles = (long long)p.tiles_w * p.tiles_h * tiles_n;
    const long long n_tiles = (C + BD_BN - 1) / BD_BN;
    if (m_tiles * n_tiles > 0x7fffffffLL || (long long)(tiles_n * p.TN) * H * W > 0x7fffffffLL) return 0;
    p.n = n; p.H = H; p.W = W; p.R = R; p.S = S; p.pad = pad; p.Cin = Cin;
    p.rows = (long long)n * H * W; p.K = R * S * Cin; p.C = C; p.Cpad = (C + 31) / 32 * 32; p.pitch = pitch;
    p.x = (const bf16*)x; p.dx = (bf16*)dx;
    p.gamma = gamma; p.beta = beta; p.mean = mean; p.var = var; p.eps = eps;
    p.dgamma = dgamma; p.dbeta = dbeta; p.accumulate = accumulate;
    p.d_out = (bf16*)d_out; p.d_pitch = d_pitch;
    p.m_tiles = (int)m_tiles; p.total_tiles = (int)(m_tiles * n_tiles); p.resb = 0;
    p.packed_s = bn_packed_scale();
    static const bool k32_on = [] { const char* e = getenv("SRGAN_NO_K32"); return !(e && e[0] == '1'); }();
    p.k32 = (k32_on && Cin == KCH && dy_valid > 0 && dy_valid <= 32) ? 1 : 0;
    const int stage_bytes = p.k32 ? BD_STAGE_BYTES / 2 : BD_STAGE_BYTES;
    const int fixed = BD_EW * EPI_STG_BYTES + BD_EW * BD_ZONE_WARP + 3 * BD_BN * 2 + BD_BN * 4 + 1024;
    int stages = (226 * 1024 - fixed) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) return 0;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + fixed;
    CUtensorMap tmA, tmB;
    int rc = p.k32 ? encode_act32(&tmA, dy, n, H, W, p.TW, p.TH, p.TN, dy_pitch, dy_valid)
                   : encode_act(&tmA, dy, n, H, W, Cin, p.TW, p.TH, p.TN, 1, dy_pitch, dy_valid);
    if (rc) return rc;
    rc = p.k32 ? encode_mat32(&tmB, Wu, Cout, (long long)R * S * Cin, BD_BN) : encode_mat(&tmB, Wu, Cout, (long long)R * S * Cin, BD_BN);
    if (rc) return rc;
    static srgan_per_device_once attr_set;
    if (attr_set.need()) {
        cudaError_t e = cudaFuncSetAttribute(bn_dgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) { srgan_set_error("cudaFuncSetAttribute(bn_dgrad_kernel<conv>): %s", cudaGetErrorString(e)); return SRGAN_ERR_CUDA; }
        attr_set.done();
    }
    const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
    bn_dgrad_kernel<true><<<grid, BD_THREADS, smem, st>>>(tmA, tmB, p);
    SRGAN_CHECK_LAUNCH("bn_dgrad_kernel<conv>");
    return 1;
}

// returns 1 = launched, 0 = shape not eligible, <0 = error
int bn_conv_down(const void* x, const void* Wd, void* out, long long rows, int Kpad, int Cout, int C, int pitch, const float* gamma,
                 const float* beta, const float* mean, const float* var, float eps, void* n1_out, int n1_pitch, long long n1_first_row,
                 const float* gamma2, const float* beta2, const float* mean2, const float* var2, void* out2, int C2, cudaStream_t st) {
    if (Kpad % KCH != 0 || C > Kpad || C <= 0 || (C & 7) || (pitch & 7) || pitch < C || (Cout & 7) || Cout <= 0 || rows <= 0) return 0;
    if (((uintptr_t)x | (uintptr_t)Wd | (uintptr_t)out | (uintptr_t)n1_out | (uintptr_t)out2) & 15) return 0;
    if (n1_out != nullptr && ((n1_pitch & 7) || n1_pitch < C)) return 0;
    if (Kpad > 4096) return 0;
    BnFpropParams p;
    p.rows = rows; p.C = C; p.Kpad = Kpad; p.pitch = pitch; p.Cout = Cout; p.out_pitch = Cout;
    p.gamma = gamma; p.beta = beta; p.mean = mean; p.var = var; p.eps = eps;
    p.out = (bf16*)out; p.n1_out = (bf16*)n1_out; p.n1_pitch = n1_pitch; p.n1_first_row = n1_first_row;
    p.gamma2 = gamma2; p.beta2 = beta2; p.mean2 = mean2; p.var2 = var2; p.out2 = (bf16*)out2; p.C2 = C2;
    const long long sub = (rows + TILE_M - 1) / TILE_M;
    p.n_tiles = (Cout + BF_BN - 1) / BF_BN;
    // 256-row CTA tiles share every weight tile between two sub-tiles (measured: 10 % less time per sub-tile), unless 128-row
    // tiles finish sooner because of the wave quantisation on 148 CTAs: time ~ rounds x rows per tile x (0.9 for 256 rows)
    int MT = 2;
    {
        const long long t1 = sub * p.n_tiles, t2 = ((sub + 1) / 2) * p.n_tiles;
        const long long r1 = (t1 + kNumSMs - 1) / kNumSMs, r2 = (t2 + kNumSMs - 1) / kNumSMs;
        if (10 * r1 < 18 * r2) MT = 1;
        static const int dbg_mt = [] { const char* e = getenv("SRGAN_BF_MT"); return e ? atoi(e) : 0; }();
        if (dbg_mt == 1 || dbg_mt == 2) MT = dbg_mt;
    }
    const long long m_tiles = (sub + MT - 1) / MT;
    if (m_tiles * p.n_tiles > 0x7fffffffLL) return 0;
    p.m_tiles = (int)m_tiles; p.total_tiles = (int)(m_tiles * p.n_tiles);
    CUtensorMap tmA, tmB;
    int rc = encode_mat_pitch(&tmA, x, rows, C, pitch, TILE_M);
    if (rc) return rc;
    rc = encode_mat(&tmB, Wd, Cout, Kpad, BF_BN);
    if (rc) return rc;
    return MT == 2 ? launch_bn_conv_down<2>(tmA, tmB, p, st) : launch_bn_conv_down<1>(tmA, tmB, p, st);
}

// returns 1 = launched, 0 = shape not eligible, <0 = error
int bn_conv_wgrad(const void* dy, const void* x, float* dW, long long rows, int Ca, int Kpad, int C, int pitch, const float* gamma,
                  const float* beta, const float* mean, const float* var, float eps, cudaStream_t st) {
    if (Kpad % KCH != 0 || C > Kpad || C <= 0 || (C & 7) || (pitch & 7) || pitch < C || Ca <= 0 || (Ca & 7) || rows <= 0) return 0;
    if (((uintptr_t)dy | (uintptr_t)x | (uintptr_t)dW) & 15) return 0;
    BnWgradParams p;
    p.rows = rows; p.C = C; p.Kpad = Kpad; p.Ca = Ca;
    p.gamma = gamma; p.beta = beta; p.mean = mean; p.var = var; p.eps = eps; p.dW = dW;
    const long long chunks = (rows + BW_PIX - 1) / BW_PIX;
    if (chunks > 0x7fffffffLL) return 0;
    p.total_chunks = (int)chunks;
    // (BN, NS) by the traffic model of umma_wgrad: t output tiles re-read dy t times, the K split over one wave of CTAs adds
    // 148 / t partial sums into dW with fp32 reductions (a reduced byte weighted 4x a read byte)
    const int a_tiles = (Ca + 127) / 128;
    double best = -1.0;
    int best_bn = 64, best_ns = 1;
    for (int bn = 64; bn <= 256; bn *= 2)
        for (int ns = 1; ns * bn <= 512; ++ns) {
            if (ns > 1 && (ns - 1) * bn >= Kpad) break;
            if (bn > 64 && bn / 2 >= Kpad) continue;
            const int t = a_tiles * ((Kpad + ns * bn - 1) / (ns * bn));
            double splits = kNumSMs / t < 1 ? 1 : kNumSMs / t;
            if (splits > (chunks + 7) / 8) splits = (double)((chunks + 7) / 8);
            if (splits < 1) splits = 1;
            const double cost = (double)rows * 128 * 2.0 * t + 4.0 * (4.0 * Ca * Kpad) * splits;
            if (best < 0 || cost < best) { best = cost; best_bn = bn; best_ns = ns; }
        }
    p.BN = best_bn; p.NS = best_ns;
    const int out_tiles = (Kpad + p.NS * p.BN - 1) / (p.NS * p.BN);
    int splits = kNumSMs / (out_tiles * a_tiles);
    const int max_splits = (int)((chunks + 7) / 8);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.chunks_per_split = (p.total_chunks + splits - 1) / splits;
    splits = (p.total_chunks + p.chunks_per_split - 1) / p.chunks_per_split;
    const int stage_bytes = 2 * BW_BOX + p.NS * (p.BN / 64) * BW_BOX;
    p.stages = (200 * 1024) / stage_bytes;
    if (p.stages > 8) p.stages = 8;
    p.stages -= p.stages % BW_NG;                // every ring stage belongs to one transform group (see launch_bn_conv_down)
    const size_t smem = (size_t)p.stages * stage_bytes + 1024;
    CUtensorMap tmS, tmL;
    int rc = encode_mat_pitch(&tmS, dy, rows, Ca, Ca, BW_PIX);
    if (rc) return rc;
    rc = encode_mat_pitch(&tmL, x, rows, C, pitch, BW_PIX);
    if (rc) return rc;
    static srgan_per_device_once attr_set;
    if (attr_set.need()) {
        cudaError_t e = cudaFuncSetAttribute(bn_conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
        if (e != cudaSuccess) { srgan_set_error("cudaFuncSetAttribute(bn_conv_wgrad_kernel): %s", cudaGetErrorString(e)); return SRGAN_ERR_CUDA; }
        attr_set.done();
    }
    bn_conv_wgrad_kernel<<<dim3(out_tiles, a_tiles, splits), BW_THREADS, smem, st>>>(tmS, tmL, p);
    SRGAN_CHECK_LAUNCH("bn_conv_wgrad_kernel");
    return 1;
}
