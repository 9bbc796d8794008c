// PTX wrappers (mbarrier, TMA, tcgen05 / TMEM), UMMA descriptors, the epilogue transposition-buffer helpers and the host-side
// tensor-map encoders shared by the tcgen05 kernels (umma_conv.cu, bn_gemm.cu).  Everything lives in an anonymous namespace:
// each translation unit gets its own copy.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace {


// ------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a CUDA error, never as a hung GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t i = 0; i < (1u << 24); ++i)
        if (mbar_try_wait(bar, parity)) return;
    printf("srgan umma: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
    __trap();
}
// one elected lane of a converged warp (lets the compiler keep descriptors / barrier addresses in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xFFFFFFFF;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B, Blackwell version 1.
//   K-major : rows of 128 B (64 bf16 of K), 8-row swizzle atoms stacked every SBO = 1024 B; LBO unused.
//   MN-major: rows of 128 B (64 bf16 of M/N) per K index, 8-K atoms every SBO = 1024 B; next 64-wide M/N chunk at LBO.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;      // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;      // SWIZZLE_128B
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, M x N, majorness bits.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int TILE_M = 128;
constexpr int KCH = 64;                        // bf16 elements per K chunk = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = TILE_M * KCH * 2;  // 16 KB

// Per-warp transposition buffer of the epilogue: 32 rows x 64 bytes (one 32-column bf16 chunk of 32 accumulator rows).
// A thread owns accumulator row `lane`, but a warp-wide 16-byte access with one ROW per lane touches 32 different
// 128-byte lines (32 LSU wavefronts per instruction: the k1s1 data-gradient GEMMs were bound by exactly that).  Through
// this buffer the global accesses are issued "transposed": lane l moves 16-byte unit (l & 3) of rows 8*it + (l >> 2),
// so one instruction covers 8 rows x 64 contiguous bytes.  Units are XOR-swizzled by (row >> 1) & 3: both access
// patterns are bank-conflict free.
constexpr int EPI_STG_BYTES = 32 * 64;
__device__ __forceinline__ uint32_t stg_addr(uint32_t base, int row, int unit) {
    return base + (uint32_t)(row * 64 + ((unit ^ ((row >> 1) & 3)) << 4));
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// 16-byte asynchronous global -> shared copy; src_bytes = 0 writes zeros (out-of-range rows / channels)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// host side: tensor maps
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_encode_once;

EncodeTiledFn get_encode() {
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            g_encode = (EncodeTiledFn)fn;
    });
    return g_encode;
}

// NHWC activation [n, H, W, C] bf16 as a 4-D map (C, W, H, n); box (64, bw, bh, bn) with element strides (1, es, es, 1)
// pitch (elements between pixels, 0 = C) and valid (channels that exist in memory, 0 = C) describe a channel window of a
// wider buffer (a slice of a DenseNet concat buffer): the box still spans 64-channel chunks of the logical C channels, the
// channels beyond `valid` are out of bounds for the map and arrive as zeros.
int act_l2_promotion() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SRGAN_ACT_L2_PROMOTION"); v = e ? atoi(e) : 128; }
    return v;
}
int encode_act(CUtensorMap* tm, const void* base, int n, int H, int W, int C, int bw, int bh, int bn, int es, int pitch = 0,
               int valid = 0) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { srgan_set_error("cuTensorMapEncodeTiled is not available from the driver"); return SRGAN_ERR_CUDA; }
    const cuuint64_t P = pitch > 0 ? pitch : C;
    cuuint64_t dims[4] = {(cuuint64_t)(valid > 0 ? valid : C), (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    cuuint64_t strides[3] = {P * 2, (cuuint64_t)W * P * 2, (cuuint64_t)H * W * P * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(bw * es), (cuuint32_t)(bh * es), (cuuint32_t)bn};
    cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
    // L2 promotion of the activation maps: 128 bytes (one box row).  256 bytes was measured (SRGAN_ACT_L2_PROMOTION=256): no
    // change for the trunk's [pixels x C] GEMMs (3.21 vs 3.22 TB/s) and -7 % on the stride-2 DCGAN convolutions, whose boxes
    // skip every other pixel (the promoted half is the skipped pixel)
    const int promo = act_l2_promotion();
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     promo >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (promo >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE),
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        srgan_set_error("cuTensorMapEncodeTiled(activation n=%d H=%d W=%d C=%d box=%dx%dx%d es=%d) failed: %d", n, H, W, C, bw, bh,
                        bn, es, (int)r);
        return SRGAN_ERR_CUDA;
    }
    return SRGAN_OK;
}
// weight matrix [rows, cols] bf16 row-major as a 2-D map (cols, rows); box (64, brows)
int encode_mat(CUtensorMap* tm, const void* base, long long rows, long long cols, int brows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { srgan_set_error("cuTensorMapEncodeTiled is not available from the driver"); return SRGAN_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)brows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        srgan_set_error("cuTensorMapEncodeTiled(matrix %lldx%lld box %d) failed: %d", rows, cols, brows, (int)r);
        return SRGAN_ERR_CUDA;
    }
    return SRGAN_OK;
}

int gcd(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

// patch TW x TH x TN with TW*TH*TN == rows (a power of two), TW | W, TH | H
bool pick_patch(int W, int H, int rows, int max_w, int& TW, int& TH, int& TN) {
    TW = gcd(W, max_w);
    TH = gcd(H, rows / TW);
    TN = rows / (TW * TH);
    return TW * TH * TN == rows && TN <= 256;
}

}  // namespace
