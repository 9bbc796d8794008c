// Crowd label preprocessing on the device: the point density label and the k-nearest-neighbour
// distance maps the reference builds once per image on the host with a scikit-learn ball tree
// (crowd/database_preprocessor.py:64-99,258-290: every pixel of the label queries its k nearest head annotations).
//
// A ball tree is a pointer-chasing CPU structure; on the GPU the whole query is a brute-force sweep: a thread owns one pixel,
// the head positions stream through shared memory in tiles, and the thread keeps the KMAX smallest squared distances in
// registers.  768 x 1024 pixels x 2 000 heads = 1.6 G distance evaluations of ~6 fp64 operations each.  All arithmetic is
// float64 and un-contracted (no FMA) so that the distances carry the bits of sklearn's EuclideanDistance (sum of squared
// coordinate differences, y first, then sqrt) and of numpy's mean over the k columns.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int kKnnMax = 8;            // largest k (the preprocessor writes k = 1..5, database_preprocessor.py:91)
constexpr int kKnnThreads = 256;
constexpr int kHeadTile = 1024;       // head positions per shared-memory tile (16 KB of double2)

template <int KMAX>
__global__ void __launch_bounds__(kKnnThreads)
knn_maps_kernel(const double* __restrict__ head_yx, int n_heads, int H, int W, int kmax, double upper_bound,
                double* __restrict__ knn, __half* __restrict__ iknn, double epsilon) {
    __shared__ double2 tile[kHeadTile];
    const long long total = (long long)H * W;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < total;
    const double Y = (double)(live ? i / W : 0), X = (double)(live ? i % W : 0);
    double best[KMAX];                  // ascending squared distances
#pragma unroll
    for (int k = 0; k < KMAX; ++k) best[k] = __longlong_as_double(0x7ff0000000000000LL);
    for (int base = 0; base < n_heads; base += kHeadTile) {
        const int n = min(kHeadTile, n_heads - base);
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += blockDim.x)
            tile[j] = make_double2(head_yx[2 * (long long)(base + j)], head_yx[2 * (long long)(base + j) + 1]);
        __syncthreads();
        for (int j = 0; j < n; ++j) {
            const double dy = __dsub_rn(Y, tile[j].x), dx = __dsub_rn(X, tile[j].y);
            double d = __dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx));
            if (d < best[KMAX - 1]) {
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {       // insertion: bubble the new value down the sorted list
                    const double lo = fmin(best[k], d);
                    d = fmax(best[k], d);
                    best[k] = lo;
                }
            }
        }
    }
    if (!live) return;
    // neighbor_distances[:, :k] -> clip(a_max = upper_bound) -> mean(axis=1)   (database_preprocessor.py:282-286), k = 1..kmax;
    // generate_knn_map uses min(k, number of heads) neighbours (:278)
    double sum = 0.0;
    const int have = min(n_heads, KMAX);
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        if (k >= kmax) break;
        if (k < have) {
            double dist = sqrt(best[k]);
            if (upper_bound > 0.0 && dist > upper_bound) dist = upper_bound;
            sum = __dadd_rn(sum, dist);
        }
        const double mean = __ddiv_rn(sum, (double)min(k + 1, have));
        if (knn) knn[(long long)k * total + i] = mean;
        if (iknn) iknn[(long long)k * total + i] = __double2half(__ddiv_rn(1.0, __dadd_rn(mean, epsilon)));   // :98-99
    }
}

// generate_point_density_map (database_preprocessor.py:246-256): density_map[int(round(y)), int(round(x))] += 1 with Python's
// round (half to even) and Python's indexing (a negative index wraps once; beyond that the head is out of bounds and counted).
__global__ void point_density_kernel(const double* __restrict__ head_yx, int n_heads, int H, int W, float* __restrict__ density,
                                     int* __restrict__ out_of_bounds) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_heads) return;
    long long y = (long long)rint(head_yx[2 * j]), x = (long long)rint(head_yx[2 * j + 1]);
    if (y < -H || y >= H || x < -W || x >= W) {
        atomicAdd(out_of_bounds, 1);
        return;
    }
    if (y < 0) y += H;
    if (x < 0) x += W;
    atomicAdd(density + y * W + x, 1.0f);           // whole numbers: exact in any order
}

}  // namespace

extern "C" {

int srgan_knn_maps(const double* head_yx, int n_heads, int H, int W, int kmax, double upper_bound, double epsilon, double* knn,
                   void* iknn_f16, void* stream) {
    SRGAN_REQUIRE(head_yx && (knn || iknn_f16), "srgan_knn_maps: null pointer");
    SRGAN_REQUIRE(n_heads > 0, "srgan_knn_maps: no head positions (the reference's NearestNeighbors.fit raises on an empty set)");
    SRGAN_REQUIRE(H > 0 && W > 0, "srgan_knn_maps: empty label");
    SRGAN_REQUIRE(kmax >= 1 && kmax <= kKnnMax, "srgan_knn_maps: k = %d outside 1..%d", kmax, kKnnMax);
    const long long total = (long long)H * W;
    const int blocks = (int)((total + kKnnThreads - 1) / kKnnThreads);
    if (kmax <= 2)
        knn_maps_kernel<2><<<blocks, kKnnThreads, 0, (cudaStream_t)stream>>>(head_yx, n_heads, H, W, kmax, upper_bound, knn,
                                                                             (__half*)iknn_f16, epsilon);
    else
        knn_maps_kernel<kKnnMax><<<blocks, kKnnThreads, 0, (cudaStream_t)stream>>>(head_yx, n_heads, H, W, kmax, upper_bound, knn,
                                                                                   (__half*)iknn_f16, epsilon);
    SRGAN_CHECK_LAUNCH("srgan_knn_maps");
    return 0;
}

int srgan_point_density_map(const double* head_yx, int n_heads, int H, int W, float* density, int* out_of_bounds, void* stream) {
    SRGAN_REQUIRE(head_yx && density && out_of_bounds, "srgan_point_density_map: null pointer");
    SRGAN_REQUIRE(n_heads >= 0 && H > 0 && W > 0, "srgan_point_density_map: bad sizes");
    cudaError_t e = cudaMemsetAsync(density, 0, sizeof(float) * (size_t)H * W, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(out_of_bounds, 0, sizeof(int), (cudaStream_t)stream);
    if (e != cudaSuccess) {
        srgan_set_error("srgan_point_density_map: memset failed: %s", cudaGetErrorString(e));
        return SRGAN_ERR_CUDA;
    }
    if (n_heads == 0) return 0;
    point_density_kernel<<<(n_heads + 255) / 256, 256, 0, (cudaStream_t)stream>>>(head_yx, n_heads, H, W, density, out_of_bounds);
    SRGAN_CHECK_LAUNCH("srgan_point_density_map");
    return 0;
}

}  // extern "C"
