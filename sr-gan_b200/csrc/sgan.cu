// SGAN method (sgan.py:18-67): the discriminator's head has K class logits (K = number_of_bins)
// instead of one regression output.  Labeled loss = cross entropy against the bin of the real label; the GAN terms are
// BCE-with-logits on logsumexp(logits); the gradient penalty differentiates BCE(logsumexp(logits(x_hat)), 0), a NONLINEAR
// function of the logits, so its double backward needs the Hessian-vector product of that loss.  The trunk passes are the
// SR-GAN step's (umma_conv.cu etc.); these kernels are the K-wide head: HBM-bound reads of the [rows, F] feature block
// (age: F = 32 768) and per-sample K-vector arithmetic.  Logit-shaped arrays are stored TRANSPOSED, [K][rows], so that one
// output's per-sample column is contiguous (it doubles as the row-scale vector of srgan_colsum for the head gradients).
#include "common.cuh"

namespace {

constexpr int kMaxK = 16;

// logitsT[k][r] = sum_c X[r,c] * W[k][c] + bias[k].  A warp owns (one row, one range of kLogitCols columns): it reads its piece of
// the row once, keeps K partial dot products per lane, warp-reduces them and writes them to partials[range][k][row]; a second
// small kernel adds the ranges in a FIXED order (+ bias), so the logits are bit-reproducible (no atomics).  The 8 warps of a
// block take 8 ROWS of the SAME column range, so the K weight rows of that range (K x 8 KB) are fetched once per block and
// served from L1 to the other warps.  First version: one block per row -- 100 CTAs each re-reading all K weight rows from
// L2: 124 us for 100 x 32 768 (ncu); this one is grid = column ranges x row groups.
constexpr int kLogitCols = 2048;

template <typename T>
__global__ void __launch_bounds__(256) head_logits_kernel(const T* __restrict__ X, int rows, int cols, const float* __restrict__ W,
                                                          int K, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int r = blockIdx.y * 8 + w;
    if (r >= rows) return;
    const int c0 = blockIdx.x * kLogitCols, c1 = min(cols, c0 + kLogitCols);
    const T* x = X + (long long)r * cols;
    float acc[kMaxK];
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) acc[k] = 0.f;
    if (cols % 4 == 0) {
        for (int c = c0 + 4 * lane; c < c1; c += 128) {
            const float4 v = ld4(x + c);
#pragma unroll
            for (int k = 0; k < kMaxK; ++k) {
                if (k < K) {
                    const float4 wv = ld4(W + (long long)k * cols + c);
                    acc[k] = fmaf(v.x, wv.x, fmaf(v.y, wv.y, fmaf(v.z, wv.z, fmaf(v.w, wv.w, acc[k]))));
                }
            }
        }
    } else {
        for (int c = c0 + lane; c < c1; c += 32) {
            const float v = to_f(x[c]);
#pragma unroll
            for (int k = 0; k < kMaxK; ++k)
                if (k < K) acc[k] = fmaf(v, W[(long long)k * cols + c], acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) {
        if (k < K) {
            const float s = warp_sum(acc[k]);
            if (lane == 0) out[((long long)blockIdx.x * K + k) * rows + r] = s;       // out = partials[range][k][row]
        }
    }
}

__global__ void __launch_bounds__(256) logits_reduce_kernel(const float* __restrict__ partials, int ranges, int K, int rows,
                                                            const float* __restrict__ bias, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * rows) return;
    float s = 0.f;
    for (int g = 0; g < ranges; ++g) s += partials[(long long)g * K * rows + i];
    out[i] = s + (bias ? bias[i / rows] : 0.f);
}

__device__ __forceinline__ float sigmoidf_(float z) { return z >= 0.f ? 1.f / (1.f + expf(-z)) : expf(z) / (1.f + expf(z)); }

// softmax p[k] and z = logsumexp(l) of one sample's K logits (column r of the transposed array)
__device__ __forceinline__ float softmax_lse(const float* __restrict__ lT, int K, int n, int r, float* p) {
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) m = fmaxf(m, lT[(long long)k * n + r]);
    float s = 0.f;
    for (int k = 0; k < K; ++k) {
        p[k] = expf(lT[(long long)k * n + r] - m);
        s += p[k];
    }
    const float inv = 1.f / s;
    for (int k = 0; k < K; ++k) p[k] *= inv;
    return m + logf(s);
}

// mode 0: cross entropy against the bin of y (nn.CrossEntropyLoss, sgan.py:20-31; bin = first minimum of |y - bins[k]|,
//         utility.py:141-144): loss += scale * (lse - l[bin]), dl = scale * (p - onehot)
// mode 1: nn.BCEWithLogitsLoss(logsumexp(l), target) (sgan.py:33-67): loss += scale * (softplus(z) - target * z),
//         dl = scale * (sigmoid(z) - target) * p
__global__ void __launch_bounds__(256) sgan_loss_kernel(const float* __restrict__ lT, int K, int n, int mode,
                                                        const float* __restrict__ y, const float* __restrict__ bins, float target,
                                                        float scale, float* __restrict__ loss, float* __restrict__ dlT) {
    __shared__ float red[32];
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    float local = 0.f;
    if (r < n) {
        float p[kMaxK];
        const float z = softmax_lse(lT, K, n, r, p);
        if (mode == 0) {
            int bin = 0;
            float best = fabsf(y[r] - bins[0]);
            for (int k = 1; k < K; ++k) {
                const float d = fabsf(y[r] - bins[k]);
                if (d < best) best = d, bin = k;
            }
            local = z - lT[(long long)bin * n + r];
            if (dlT)
                for (int k = 0; k < K; ++k) dlT[(long long)k * n + r] = scale * (p[k] - (k == bin ? 1.f : 0.f));
        } else {
            local = fmaxf(z, 0.f) - target * z + log1pf(expf(-fabsf(z)));
            const float dz = sigmoidf_(z) - target;
            if (dlT)
                for (int k = 0; k < K; ++k) dlT[(long long)k * n + r] = scale * dz * p[k];
        }
    }
    const float s = block_sum(local, red);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, scale * s);
}

// Hessian-vector product of L = c * softplus(logsumexp(l)) per sample: with p = softmax(l), sg = sigmoid(z), a = p . t,
//   q = c * [ sg (1 - sg) a p + sg (p * t - a p) ]      (d/dl of sum_k s_k(l) t_k, s = dL/dl = c sg p)
__global__ void __launch_bounds__(256) sgan_gp_second_kernel(const float* __restrict__ lT, const float* __restrict__ tT, int K, int n,
                                                             float c, float* __restrict__ qT) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float p[kMaxK];
    const float z = softmax_lse(lT, K, n, r, p);
    const float sg = sigmoidf_(z);
    float a = 0.f;
    for (int k = 0; k < K; ++k) a = fmaf(p[k], tT[(long long)k * n + r], a);
    for (int k = 0; k < K; ++k) {
        const float t = tT[(long long)k * n + r];
        qT[(long long)k * n + r] = c * (sg * (1.f - sg) * a * p[k] + sg * (p[k] * t - a * p[k]));
    }
}

// out[r,c] = (sum_k dT[k][r] * W[k][c]) * act'(href[r,c]) : the K-output form of srgan_seed_rows.  A thread owns 4 columns: the K
// weight quads stay in registers over the thread's rows (blockIdx.y = row range), the K per-row coefficients are warp-uniform
// loads.  First version: one thread per element re-reading K weights per element -- 38 us for 100 x 32 768 (ncu).
template <typename T>
__global__ void __launch_bounds__(256) seed_rows_multi_kernel(T* __restrict__ out, int rows, int cols, const float* __restrict__ dT,
                                                              const float* __restrict__ W, int K, const T* __restrict__ href, int act,
                                                              float slope) {
    const int r0 = (int)(((long long)rows * blockIdx.y) / gridDim.y), r1 = (int)(((long long)rows * (blockIdx.y + 1)) / gridDim.y);
    if (cols % 4 == 0) {
        const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
        if (c >= cols) return;
        float4 wv[kMaxK];
#pragma unroll
        for (int k = 0; k < kMaxK; ++k) wv[k] = k < K ? ld4(W + (long long)k * cols + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = r0; r < r1; ++r) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < kMaxK; ++k) {
                if (k < K) {
                    const float d = dT[(long long)k * rows + r];
                    s.x = fmaf(d, wv[k].x, s.x); s.y = fmaf(d, wv[k].y, s.y); s.z = fmaf(d, wv[k].z, s.z); s.w = fmaf(d, wv[k].w, s.w);
                }
            }
            const float4 h = ld4(href + (long long)r * cols + c);
            st4(out + (long long)r * cols + c, make_float4(s.x * act_bwd(h.x, act, slope), s.y * act_bwd(h.y, act, slope),
                                                          s.z * act_bwd(h.z, act, slope), s.w * act_bwd(h.w, act, slope)));
        }
    } else {
        const int c = blockIdx.x * blockDim.x + threadIdx.x;
        if (c >= cols) return;
        for (int r = r0; r < r1; ++r) {
            float s = 0.f;
            for (int k = 0; k < K; ++k) s = fmaf(dT[(long long)k * rows + r], W[(long long)k * cols + c], s);
            out[(long long)r * cols + c] = from_f<T>(s * act_bwd(to_f(href[(long long)r * cols + c]), act, slope));
        }
    }
}

// dW[k][c] += sum_r dT[k][r] * X[r,c]  and  db[k] += sum_r dT[k][r]  for all K outputs in ONE pass over the feature block (as K
// row-scaled column sums the block was read K times).  A thread owns 4 columns x K accumulators; the 8 warps of a block take
// interleaved rows of the block's row range; partial sums meet in shared memory and leave with one atomicAdd per (k, column).
template <typename T>
__global__ void __launch_bounds__(256) head_wgrad_kernel(const T* __restrict__ X, int rows, int cols, const float* __restrict__ dT, int K,
                                                         float* __restrict__ dW, float* __restrict__ db) {
    __shared__ float red[8][32][4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + lane) * 4;
    const int r0 = (int)(((long long)rows * blockIdx.y) / gridDim.y), r1 = (int)(((long long)rows * (blockIdx.y + 1)) / gridDim.y);
    float4 acc[kMaxK];
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < cols) {
        for (int r = r0 + w; r < r1; r += 8) {
            const float4 v = ld4(X + (long long)r * cols + c);
#pragma unroll
            for (int k = 0; k < kMaxK; ++k) {
                if (k < K) {
                    const float sc = dT[(long long)k * rows + r];
                    acc[k].x = fmaf(sc, v.x, acc[k].x); acc[k].y = fmaf(sc, v.y, acc[k].y);
                    acc[k].z = fmaf(sc, v.z, acc[k].z); acc[k].w = fmaf(sc, v.w, acc[k].w);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) {
        if (k < K) {                                          // K is uniform over the block
            __syncthreads();
            red[w][lane][0] = acc[k].x; red[w][lane][1] = acc[k].y; red[w][lane][2] = acc[k].z; red[w][lane][3] = acc[k].w;
            __syncthreads();
            if (threadIdx.x < 128) {
                const int g = threadIdx.x >> 2, j = threadIdx.x & 3;
                const int cc = (blockIdx.x * 32 + g) * 4 + j;
                if (cc < cols) {
                    float sum = 0.f;
#pragma unroll
                    for (int y = 0; y < 8; ++y) sum += red[y][g][j];
                    atomicAdd(dW + (long long)k * cols + cc, sum);
                }
            }
        }
    }
    if (db && blockIdx.x == 0 && blockIdx.y == 0 && w < K && w < 8) {        // bias sums: warp k (and k + 8) over the rows
        for (int k = w; k < K; k += 8) {
            float s = 0.f;
            for (int r = lane; r < rows; r += 32) s += dT[(long long)k * rows + r];
            s = warp_sum(s);
            if (lane == 0) atomicAdd(db + k, s);
        }
    }
}

}  // namespace

extern "C" {

int srgan_head_wgrad(const void* X, int rows, int cols, const float* dT, int K, float* dW, float* db, int dtype, void* stream) {
    SRGAN_REQUIRE(X && dT && dW && rows >= 0 && cols > 0, "srgan_head_wgrad: bad arguments");
    SRGAN_REQUIRE(K >= 1 && K <= kMaxK, "srgan_head_wgrad: K = %d outside 1..%d", K, kMaxK);
    SRGAN_REQUIRE(cols % 4 == 0, "srgan_head_wgrad: cols = %d is not a multiple of 4", cols);
    if (rows == 0) return SRGAN_OK;
    const int gx = cdiv(cols, 128);
    long long gy = (2LL * kNumSMs + gx - 1) / gx;
    if (gy > (rows + 15) / 16) gy = (rows + 15) / 16;          // at least ~2 rows per warp
    if (gy < 1) gy = 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SRGAN_F32) head_wgrad_kernel<float><<<dim3(gx, (unsigned)gy), 256, 0, st>>>((const float*)X, rows, cols, dT, K, dW, db);
    else head_wgrad_kernel<bf16><<<dim3(gx, (unsigned)gy), 256, 0, st>>>((const bf16*)X, rows, cols, dT, K, dW, db);
    SRGAN_CHECK_LAUNCH("head_wgrad_kernel");
    return SRGAN_OK;
}

size_t srgan_head_logits_workspace_bytes(int rows, int cols, int K) {
    return sizeof(float) * (size_t)cdiv(cols > 0 ? cols : 1, kLogitCols) * (size_t)(K > 0 ? K : 0) * (size_t)(rows > 0 ? rows : 0);
}

int srgan_head_logits(const void* X, int rows, int cols, const float* W, const float* bias, int K, float* logitsT, void* workspace,
                      size_t workspace_bytes, int dtype, void* stream) {
    SRGAN_REQUIRE(X && W && logitsT && rows >= 0 && cols > 0, "srgan_head_logits: bad arguments");
    SRGAN_REQUIRE(K >= 1 && K <= kMaxK, "srgan_head_logits: K = %d outside 1..%d", K, kMaxK);
    if (rows == 0) return SRGAN_OK;
    SRGAN_REQUIRE(workspace && workspace_bytes >= srgan_head_logits_workspace_bytes(rows, cols, K), "srgan_head_logits: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int ranges = cdiv(cols, kLogitCols);
    const dim3 grid(ranges, cdiv(rows, 8));
    float* partials = (float*)workspace;
    if (dtype == SRGAN_F32) head_logits_kernel<float><<<grid, 256, 0, st>>>((const float*)X, rows, cols, W, K, partials);
    else head_logits_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)X, rows, cols, W, K, partials);
    SRGAN_CHECK_LAUNCH("head_logits_kernel");
    logits_reduce_kernel<<<cdiv((long long)K * rows, 256), 256, 0, st>>>(partials, ranges, K, rows, bias, logitsT);
    SRGAN_CHECK_LAUNCH("logits_reduce_kernel");
    return SRGAN_OK;
}

int srgan_sgan_loss(const float* logitsT, int K, int n, int mode, const float* y, const float* bins, float target, float scale,
                    float* loss, float* dlogitsT, void* stream) {
    SRGAN_REQUIRE(logitsT && n >= 0, "srgan_sgan_loss: bad arguments");
    SRGAN_REQUIRE(K >= 1 && K <= kMaxK, "srgan_sgan_loss: K = %d outside 1..%d", K, kMaxK);
    SRGAN_REQUIRE(mode == 1 || (mode == 0 && y && bins), "srgan_sgan_loss: the cross-entropy mode needs labels and bins");
    if (n == 0) return SRGAN_OK;
    sgan_loss_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(logitsT, K, n, mode, y, bins, target, scale, loss, dlogitsT);
    SRGAN_CHECK_LAUNCH("sgan_loss_kernel");
    return SRGAN_OK;
}

int srgan_sgan_gp_second(const float* logitsT, const float* tangentT, int K, int n, float c, float* qT, void* stream) {
    SRGAN_REQUIRE(logitsT && tangentT && qT && n >= 0, "srgan_sgan_gp_second: bad arguments");
    SRGAN_REQUIRE(K >= 1 && K <= kMaxK, "srgan_sgan_gp_second: K = %d outside 1..%d", K, kMaxK);
    if (n == 0) return SRGAN_OK;
    sgan_gp_second_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(logitsT, tangentT, K, n, c, qT);
    SRGAN_CHECK_LAUNCH("sgan_gp_second_kernel");
    return SRGAN_OK;
}

int srgan_seed_rows_multi(void* out, int rows, int cols, const float* dT, const float* W, int K, const void* href, int act,
                          float slope, int dtype, void* stream) {
    SRGAN_REQUIRE(out && dT && W && href && rows >= 0 && cols > 0, "srgan_seed_rows_multi: bad arguments");
    SRGAN_REQUIRE(K >= 1 && K <= kMaxK, "srgan_seed_rows_multi: K = %d outside 1..%d", K, kMaxK);
    if (rows == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int gx = cdiv(cols % 4 == 0 ? cols / 4 : cols, 256);
    long long gy = (4LL * kNumSMs + gx - 1) / gx;               // row ranges: ~4 waves of blocks, at least 4 rows each
    if (gy > (rows + 3) / 4) gy = (rows + 3) / 4;
    if (gy < 1) gy = 1;
    const dim3 grid(gx, (unsigned)gy);
    if (dtype == SRGAN_F32)
        seed_rows_multi_kernel<float><<<grid, 256, 0, st>>>((float*)out, rows, cols, dT, W, K, (const float*)href, act, slope);
    else
        seed_rows_multi_kernel<bf16><<<grid, 256, 0, st>>>((bf16*)out, rows, cols, dT, W, K, (const bf16*)href, act, slope);
    SRGAN_CHECK_LAUNCH("seed_rows_multi_kernel");
    return SRGAN_OK;
}

}  // extern "C"
