// Shared helpers for the sm_100a kernels of the SR-GAN step.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "../../include/srgan_b200.h"

typedef __nv_bfloat16 bf16;

// ---- error plumbing (api.cu owns the storage)
void srgan_set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
#define SRGAN_COUNT_LAUNCH() (g_launches.fetch_add(1, std::memory_order_relaxed))

#define SRGAN_CHECK_LAUNCH(name)                                                        \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) {                                                       \
            srgan_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
            return SRGAN_ERR_CUDA;                                                      \
        }                                                                               \
        SRGAN_COUNT_LAUNCH();                                                           \
    } while (0)

#define SRGAN_REQUIRE(cond, ...)               \
    do {                                       \
        if (!(cond)) {                         \
            srgan_set_error(__VA_ARGS__);      \
            return SRGAN_ERR_ARG;              \
        }                                      \
    } while (0)

// ---- a kernel attribute (cudaFuncSetAttribute: dynamic shared memory opt-in) belongs to the DEVICE the call was made on:
// remembered per device, not per process (two devices driven from one process each need their own call)
struct srgan_per_device_once {
    std::atomic<unsigned long long> mask{0};
    static unsigned long long bit() {
        int d = 0;
        cudaGetDevice(&d);
        return 1ull << (d & 63);
    }
    bool need() const { return !(mask.load(std::memory_order_acquire) & bit()); }
    void done() { mask.fetch_or(bit(), std::memory_order_release); }
};

// ---- element conversion
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements <-> float4 (caller guarantees 4-element alignment of the address)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
}

// ---- activations
__device__ __forceinline__ float act_fwd(float x, int act, float slope) {
    if (act == SRGAN_ACT_LEAKY) return x > 0.f ? x : x * slope;
    if (act == SRGAN_ACT_TANH) return tanhf(x);
    return x;
}
// derivative expressed through the stored POST-activation value h (leaky: sign(h)==sign(a); tanh: 1-h^2)
__device__ __forceinline__ float act_bwd(float h, int act, float slope) {
    if (act == SRGAN_ACT_LEAKY) return h > 0.f ? 1.f : slope;
    if (act == SRGAN_ACT_TANH) return 1.f - h * h;
    return 1.f;
}

// ---- eval-mode BatchNorm forward, ONE formula for every kernel that evaluates it (the streaming affine kernels, the
// transform-on-load of bn_conv_down / bn_conv_wgrad and the ReLU mask that bn_dgrad recomputes must agree bit for bit):
//   y = fma(x, s, t),  s = gamma / sqrt(var + eps),  t = fma(-mean, s, beta)
__device__ __forceinline__ float bn_shift(float beta, float mean, float s) { return fmaf(-mean, s, beta); }
__device__ __forceinline__ float bn_apply(float x, float s, float t) { return fmaf(x, s, t); }

// ---- reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum broadcast to every thread; `red` = __shared__ float[32]
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
    if (w == 0) {
        t = warp_sum(t);
        if (lane == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
constexpr int kNumSMs = 148;
