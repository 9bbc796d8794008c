// 3x3, stride-1, same-size convolutions with few output channels (conv2 of a DenseNet dense layer, crowd/models.py:345-346:
// 128 -> growth_rate = 32 channels) as a tcgen05 implicit GEMM that reads every input pixel from L2 ONCE.
//
// The tap-per-stage kernel (umma_conv.cu) loads the shifted input tile once per filter tap: 9 x 16 KB of L2 -> SM traffic per
// 64 input channels for 16 KB of distinct data; at ~50 GB/s of ingest per SM these launches ran at < 1 TB/s of algorithmic
// bytes.  Here a ring stage holds a zero-padded PATCH of the image -- (W+1) x (TH+2) x TN pixels of 64 channels, one TMA box
// whose out-of-image coordinates arrive as zeros -- laid out flat: pixel f of the patch is the 128-byte row f of the stage
// (K-major, SWIZZLE_128B).  With one shared zero column between image rows a filter tap (r, s) is a CONSTANT row offset
// r*(W+1) + s in that flat order, so the A operand of the tap is the same shared memory, 128 consecutive rows starting
// r*(W+1) + s rows further: a UMMA descriptor whose start address is not 1024-byte aligned (the swizzle pattern is a function
// of the absolute shared-memory address, so it stays anchored to the stage).  Rows whose flat position falls on the
// zero column / the two halo rows / beyond the last sample produce garbage accumulator rows that the epilogue skips.
// All 9 x (Cin / 64) weight tiles (N = 32 rows) stay resident in shared memory for the whole kernel.
//
//   warp 0 TMA, warp 1 MMA + TMEM, warps 2..5 epilogue (row per thread: 32 channels = 64 contiguous bytes per pixel).
#include "umma_common.cuh"

namespace {

constexpr int F3_N = 32;                         // output channels per launch (the accumulator's columns per row tile)
constexpr int F3_THREADS = 192;
constexpr int F3_MAX_TILES = 8;                  // row tiles of 128 flat positions per patch (2 x 8 x 32 = 512 TMEM columns)

struct Flat3Params {
    int n, H, W;                  // samples, image extent (input = output)
    int nk;                       // input channel chunks of 64
    int BW, BH, TH, TN;           // patch box: BW = W + 1 columns, BH = TH + 2 rows, TN samples; TH output rows per patch
    int P;                        // flat positions per patch = BW * BH * TN
    int F;                        // last candidate output position of a patch = P - 2 * BW - 2 ... see config(): P - 2 * BW - 1
    int tiles;                    // ceil((F + 1) / 128)
    int bands, total_patches;     // ceil(H / TH) row bands per image; bands * ceil(n / TN) patches
    int stage_bytes, stages;
    bf16* out;
    int out_pitch, out_valid;     // elements between output pixels, output channels that exist (<= 32, a multiple of 8)
};

// SWIZZLE_128B K-major descriptor whose start address is 128-byte but not 1024-byte aligned.  Measured on the B200 (tools/
// f3dbg.py, one tap at a time): the swizzle is applied to the ABSOLUTE shared-memory address bits, so a start address shifted by
// whole 128-byte rows needs nothing but the shifted address -- with the descriptor's base-offset field set to (address >> 7) & 7
// every tap whose row offset is not a multiple of 8 came out wrong, with the field left 0 all nine are exact.
__device__ __forceinline__ uint64_t make_desc_off(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(16 >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(F3_THREADS, 1) flat3x3_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB, const Flat3Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[4];
    __shared__ __align__(8) uint64_t empty_bar[4];
    __shared__ __align__(8) uint64_t w_bar;
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_slot;

    constexpr int N = F3_N;
    constexpr int W_TILE_BYTES = N * KCH * 2;            // one (tap, chunk) weight tile: 4 KB
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_smem = base;                        // 9 * nk weight tiles
    const uint32_t ring = base + (uint32_t)(9 * p.nk * W_TILE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stages = p.stages;
    const int acc_cols = p.tiles * N;
    uint32_t ncols = 32;
    while ((int)ncols < 2 * acc_cols) ncols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        mbar_init(smem_u32(&w_bar), 1);
        for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tmem_full_bar[b]), 1); mbar_init(smem_u32(&tmem_empty_bar[b]), 4); }
        fence_barrier_init();
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), ncols);
    // rows P .. of every stage are never written by the TMA, but row P is read as the right-hand zero padding of the patch's
    // last output pixel (and the rows after it by garbage positions): zero them once
    {
        const int slack16 = (p.stage_bytes - p.P * 128) / 16;
        for (int s = 0; s < stages; ++s)
            for (int i = threadIdx.x; i < slack16; i += F3_THREADS)
                sts128(ring + (uint32_t)(s * p.stage_bytes + p.P * 128 + i * 16), make_uint4(0u, 0u, 0u, 0u));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            const uint32_t wb = smem_u32(&w_bar);
            mbar_expect_tx(wb, (uint32_t)(9 * p.nk * W_TILE_BYTES));
            for (int t = 0; t < 9; ++t)
                for (int c = 0; c < p.nk; ++c)
                    tma_load_2d(w_smem + (uint32_t)((t * p.nk + c) * W_TILE_BYTES), &tmB, wb, (t * p.nk + c) * KCH, 0);
            int s = 0;
            uint32_t ph = 0;
            for (int patch = blockIdx.x; patch < p.total_patches; patch += gridDim.x) {
                const int y0 = (patch % p.bands) * p.TH, n0 = (patch / p.bands) * p.TN;
                for (int c = 0; c < p.nk; ++c) {
                    mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    mbar_expect_tx(fb, (uint32_t)(p.P * 128));
                    tma_load_4d(ring + (uint32_t)(s * p.stage_bytes), &tmA, fb, c * KCH, -1, y0 - 1, n0);
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(TILE_M, N, 0, 0);
            mbar_wait(smem_u32(&w_bar), 0);
            int s = 0, pl = 0;
            uint32_t ph = 0;
            for (int patch = blockIdx.x; patch < p.total_patches; patch += gridDim.x, ++pl) {
                const int buf = pl & 1;
                const uint32_t bph = (pl >> 1) & 1;
                mbar_wait(smem_u32(&tmem_empty_bar[buf]), bph ^ 1);
                tc_fence_after();
                for (int c = 0; c < p.nk; ++c) {
                    mbar_wait(smem_u32(&full_bar[s]), ph);
                    tc_fence_after();
                    const uint32_t a_s = ring + (uint32_t)(s * p.stage_bytes);
                    for (int j = 0; j < p.tiles; ++j) {
                        const uint32_t acc = tmem_base + (uint32_t)(buf * acc_cols + j * N);
                        for (int t = 0; t < 9; ++t) {
                            const int off = (t / 3) * p.BW + (t % 3);              // the tap's constant row offset in the flat patch
                            const uint32_t a_row = a_s + (uint32_t)((j * TILE_M + off) * 128);
                            const uint32_t b_t = w_smem + (uint32_t)((t * p.nk + c) * W_TILE_BYTES);
#pragma unroll
                            for (int k = 0; k < KCH / 16; ++k)
                                umma_f16(acc, make_desc_off(a_row + k * 32), make_desc_off(b_t + k * 32), idesc,
                                         (c > 0 || t > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(smem_u32(&empty_bar[s]));
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
                umma_commit(smem_u32(&tmem_full_bar[buf]));
            }
        }
    } else {
        // ================= epilogue (4 warps): accumulator row = flat position -> its pixel, 64 bytes per row =================
        const int q = warp & 3;
        const int per_sample = p.BW * p.BH;
        int pl = 0;
        for (int patch = blockIdx.x; patch < p.total_patches; patch += gridDim.x, ++pl) {
            const int buf = pl & 1;
            const uint32_t bph = (pl >> 1) & 1;
            const int y0 = (patch % p.bands) * p.TH, n0 = (patch / p.bands) * p.TN;
            mbar_wait(smem_u32(&tmem_full_bar[buf]), bph);
            tc_fence_after();
            for (int j = 0; j < p.tiles; ++j) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * acc_cols + j * N), v);
                const int f = j * TILE_M + q * 32 + lane;
                const int sl = f / per_sample, rem = f - sl * per_sample;
                const int yy = rem / p.BW, xx = rem - yy * p.BW;
                const bool ok = f <= p.F && yy < p.TH && xx < p.W && y0 + yy < p.H && n0 + sl < p.n;
                tmem_ld_wait();
                if (ok) {
                    bf16* dst = p.out + ((long long)((n0 + sl) * p.H + y0 + yy) * p.W + xx) * p.out_pitch;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (g * 8 < p.out_valid) {
                            uint4 w;
                            __nv_bfloat162 b0 = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
                            __nv_bfloat162 b1 = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
                            __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
                            __nv_bfloat162 b3 = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
                            w.x = *reinterpret_cast<uint32_t*>(&b0); w.y = *reinterpret_cast<uint32_t*>(&b1);
                            w.z = *reinterpret_cast<uint32_t*>(&b2); w.w = *reinterpret_cast<uint32_t*>(&b3);
                            *reinterpret_cast<uint4*>(dst + g * 8) = w;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[buf]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// zero-padded patch map: NHWC activation [n, H, W, C] as (C, W, H, n) with box (64, BW, BH, TN); negative / beyond-extent
// coordinates are zero fill.  in_pitch / in_valid describe a channel window as in encode_act.
int encode_patch(CUtensorMap* tm, const void* base, int n, int H, int W, int C, int BW, int BH, int TN, int pitch, int valid) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { srgan_set_error("cuTensorMapEncodeTiled is not available from the driver"); return SRGAN_ERR_CUDA; }
    const cuuint64_t P = pitch > 0 ? pitch : C;
    cuuint64_t dims[4] = {(cuuint64_t)(valid > 0 ? valid : C), (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    cuuint64_t strides[3] = {P * 2, (cuuint64_t)W * P * 2, (cuuint64_t)H * W * P * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)TN};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        srgan_set_error("cuTensorMapEncodeTiled(patch n=%d H=%d W=%d C=%d box=%dx%dx%d) failed: %d", n, H, W, C, BW, BH, TN, (int)r);
        return SRGAN_ERR_CUDA;
    }
    return SRGAN_OK;
}

}  // namespace

// returns 1 = launched, 0 = shape not eligible (the caller continues with the tap-per-stage kernel), <0 = error
int flat3x3_conv(const void* src, const void* Wd, void* out, int n, const srgan_geom* g, int in_pitch, int in_valid,
                 int out_pitch, int out_valid, cudaStream_t st) {
    static const bool on = [] { const char* e = getenv("SRGAN_NO_FLAT3X3"); return !(e && e[0] == '1'); }();
    if (!on) return 0;
    if (g->R != 3 || g->S != 3 || g->stride != 1 || g->pad != 1 || g->Hs != g->Hl || g->Ws != g->Wl) return 0;
    if (g->Cb % KCH != 0 || out_valid <= 0 || out_valid > F3_N || (out_valid & 7) || (out_pitch & 7) || g->Ca < F3_N) return 0;
    if (((uintptr_t)src | (uintptr_t)Wd | (uintptr_t)out) & 15) return 0;
    const int H = g->Hl, W = g->Wl, nk = g->Cb / KCH;
    if (W + 1 > 256 || H + 2 > 256 || n <= 0) return 0;
    const int w_bytes = 9 * nk * F3_N * KCH * 2;
    Flat3Params p;
    p.n = n; p.H = H; p.W = W; p.nk = nk; p.BW = W + 1;
    p.stages = 2;
    // patch = TH output rows x TN samples; fits when two stages (patch + the rows the last tile's shifted reads run into) and
    // the weights fit in shared memory and its row tiles fit in TMEM
    auto config = [&](int th, int tn) {
        p.TH = th; p.TN = tn; p.BH = th + 2;
        p.P = p.BW * p.BH * tn;
        p.F = p.P - 2 * p.BW - 1;
        p.tiles = (p.F + TILE_M) / TILE_M;         // positions 0 .. F
        const int rows_needed = p.tiles * TILE_M + 2 * p.BW + 2;
        p.stage_bytes = ((rows_needed > p.P + 1 ? rows_needed : p.P + 1) * 128 + 1023) / 1024 * 1024;
        return p.tiles >= 1 && p.tiles <= F3_MAX_TILES && p.BH <= 256 && tn <= 256 &&
               (size_t)w_bytes + (size_t)p.stages * p.stage_bytes + 1024 <= (size_t)224 * 1024;
    };
    if (config(H, 1)) {                          // whole images: as many samples per patch as fit, but one wave of patches first
        int tn = n < 256 ? n : 256;
        while (tn > 1 && (!config(H, tn) || (n + tn - 1) / tn < kNumSMs)) --tn;
        if (!config(H, tn)) return 0;
    } else {                                     // row bands of one image
        int th = H;
        while (th > 1 && !config(th, 1)) --th;
        if (!config(th, 1)) return 0;
        const int bands = (H + th - 1) / th;
        if (!config((H + bands - 1) / bands, 1)) return 0;
    }
    p.bands = (H + p.TH - 1) / p.TH;
    const long long total = (long long)p.bands * ((n + p.TN - 1) / p.TN);
    if (total > 0x7fffffffLL) return 0;
    p.total_patches = (int)total;
    p.out = (bf16*)out; p.out_pitch = out_pitch; p.out_valid = out_valid;
    const size_t smem = (size_t)w_bytes + (size_t)p.stages * p.stage_bytes + 1024;
    if (smem > 226 * 1024) return 0;
    CUtensorMap tmA, tmB;
    int rc = encode_patch(&tmA, src, n, H, W, g->Cb, p.BW, p.BH, p.TN, in_pitch, in_valid);
    if (rc) return rc;
    rc = encode_mat(&tmB, Wd, g->Ca, (long long)9 * g->Cb, F3_N);
    if (rc) return rc;
    static srgan_per_device_once attr_set;
    if (attr_set.need()) {
        cudaError_t e = cudaFuncSetAttribute(flat3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) { srgan_set_error("cudaFuncSetAttribute(flat3x3_kernel): %s", cudaGetErrorString(e)); return SRGAN_ERR_CUDA; }
        attr_set.done();
    }
    const int grid = p.total_patches < kNumSMs ? p.total_patches : kNumSMs;
    flat3x3_kernel<<<grid, F3_THREADS, smem, st>>>(tmA, tmB, p);
    SRGAN_CHECK_LAUNCH("flat3x3_kernel");
    return 1;
}
