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

// ---------------------------------------------------------------------------------------------------------------------
// generate_density_label (crowd/database_preprocessor.py:113-225) as generate_labels_for_example calls it (:81-89; run.py:67
// trains the crowd application on the beta = 0.3 maps): every head adds a normalised square Gaussian whose standard deviation
// is beta x the mean distance to its min(11, n) nearest heads (itself included), clipped at the label's borders; the sum is
// rescaled to the head count.  Three launches: per-head geometry (thread per head, brute-force neighbours), per-head kernel
// sum (block per head), then a GATHER per pixel over the heads in annotation order -- the order the reference's
// `label += person_label` runs in, so the fp32 accumulation is the same sequence of additions.
struct head_geom {
    long long y, x;       // np.rint(position).astype(np.uint32)
    int off, skip;        // kernel half-width int(2 sigma); 'Offset out of head gaussian bounds' (:189-191)
    double two_s2, sum;   // 2 sigma^2; sum of the full (unclipped) kernel
};

constexpr int kSpreadK = 11;

__global__ void __launch_bounds__(128)
head_geometry_kernel(const double* __restrict__ head_yx, int n, int H, int W, double beta, head_geom* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double py = head_yx[2 * i], px = head_yx[2 * i + 1];
    double best[kSpreadK];
#pragma unroll
    for (int k = 0; k < kSpreadK; ++k) best[k] = __longlong_as_double(0x7ff0000000000000LL);
    for (int j = 0; j < n; ++j) {
        const double dy = __dsub_rn(py, head_yx[2 * j]), dx = __dsub_rn(px, head_yx[2 * j + 1]);
        double d = __dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx));
        if (d < best[kSpreadK - 1]) {
#pragma unroll
            for (int k = 0; k < kSpreadK; ++k) {
                const double lo = fmin(best[k], d);
                d = fmax(best[k], d);
                best[k] = lo;
            }
        }
    }
    const int k = n < kSpreadK ? n : kSpreadK;
    double a[kSpreadK];
#pragma unroll
    for (int j = 0; j < kSpreadK; ++j) a[j] = j < k ? sqrt(best[j]) : 0.0;
    // numpy's mean over the k contiguous columns: a plain loop below 8 elements, else 8 running sums combined pairwise and
    // the remainder added in order (numpy/core/src/umath/loops_utils.h pairwise sum)
    double sum;
    if (k < 8) {
        sum = 0.0;
        for (int j = 0; j < k; ++j) sum = __dadd_rn(sum, a[j]);
    } else {
        sum = __dadd_rn(__dadd_rn(__dadd_rn(a[0], a[1]), __dadd_rn(a[2], a[3])), __dadd_rn(__dadd_rn(a[4], a[5]), __dadd_rn(a[6], a[7])));
        for (int j = 8; j < k; ++j) sum = __dadd_rn(sum, a[j]);
    }
    const double sigma = __dmul_rn(__ddiv_rn(sum, (double)k), beta);
    head_geom g;
    g.y = (long long)(unsigned int)(long long)rint(py);
    g.x = (long long)(unsigned int)(long long)rint(px);
    g.off = (int)__dmul_rn(sigma, 2.0);
    g.two_s2 = __dmul_rn(2.0, __dmul_rn(sigma, sigma));
    const long long size = 2LL * g.off + 1;
    const long long y0 = g.off - g.y > 0 ? g.off - g.y : 0, y1 = g.y + g.off + 1 - H > 0 ? g.y + g.off + 1 - H : 0;
    const long long x0 = g.off - g.x > 0 ? g.off - g.x : 0, x1 = g.x + g.off + 1 - W > 0 ? g.x + g.off + 1 - W : 0;
    g.skip = (size <= (y0 > y1 ? y0 : y1)) || (size <= (x0 > x1 ? x0 : x1));
    g.sum = 0.0;
    out[i] = g;
}

__device__ __forceinline__ double gaussian_value(long long dy, long long dx, double two_s2) {
    const double fy = (double)dy, fx = (double)dx;       // exp(-(x^2 / (2 s^2) + y^2 / (2 s^2))), :238-241
    return exp(-__dadd_rn(__ddiv_rn(__dmul_rn(fx, fx), two_s2), __ddiv_rn(__dmul_rn(fy, fy), two_s2)));
}

__global__ void __launch_bounds__(128) head_kernel_sum_kernel(head_geom* __restrict__ geoms) {
    __shared__ double red[4];
    head_geom& g = geoms[blockIdx.x];
    const int size = 2 * g.off + 1;
    double local = 0.0;
    for (int e = threadIdx.x; e < size * size; e += blockDim.x)
        local += gaussian_value(e / size - g.off, e % size - g.off, g.two_s2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) g.sum = (red[0] + red[1]) + (red[2] + red[3]);
}

constexpr int kDensityThreads = 256;
constexpr int kGeomTile = 256;

__global__ void __launch_bounds__(kDensityThreads)
density_label_kernel(const head_geom* __restrict__ geoms, int n, int H, int W, float* __restrict__ label,
                     double* __restrict__ partials) {
    __shared__ head_geom tile[kGeomTile];
    __shared__ double red[kDensityThreads / 32];
    const long long total = (long long)H * W;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < total;
    const long long Y = live ? i / W : 0, X = live ? i % W : 0;
    float acc = 0.f;
    for (int base = 0; base < n; base += kGeomTile) {
        const int m = min(kGeomTile, n - base);
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += blockDim.x) tile[j] = geoms[base + j];
        __syncthreads();
        for (int j = 0; j < m; ++j) {
            const head_geom& g = tile[j];
            if (g.skip) continue;
            const long long dy = Y - g.y, dx = X - g.x;
            if (dy < -g.off || dy > g.off || dx < -g.off || dx > g.off) continue;
            // person_label (float32 zeros) += gaussian / gaussian.sum()  ->  label += person_label
            acc = __fadd_rn(acc, (float)__ddiv_rn(gaussian_value(dy, dx, g.two_s2), g.sum));
        }
    }
    if (live) label[i] = acc;
    double local = live ? (double)acc : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < kDensityThreads / 32; ++w) s += red[w];
        partials[blockIdx.x] = s;
    }
}

// label = head_count * (label / label.sum())  in float32 (:223-224); one block first folds the per-block partial sums
__global__ void __launch_bounds__(256)
density_normalise_kernel(float* __restrict__ label, long long total, const double* __restrict__ partials, int n_partials,
                         float head_count, __half* __restrict__ label_f16) {
    __shared__ double red[8];
    __shared__ float s_sum;
    double local = 0.0;
    for (int j = threadIdx.x; j < n_partials; j += blockDim.x) local += partials[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        s_sum = (float)s;
    }
    __syncthreads();
    const float sum = s_sum;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float v = __fmul_rn(head_count, __fdiv_rn(label[i], sum));
        label[i] = v;
        if (label_f16) label_f16[i] = __float2half_rn(v);
    }
}

}  // namespace

extern "C" {

size_t srgan_density_label_workspace_bytes(int n_heads, int H, int W) {
    const long long blocks = ((long long)H * W + kDensityThreads - 1) / kDensityThreads;
    return sizeof(head_geom) * (size_t)(n_heads > 0 ? n_heads : 0) + sizeof(double) * (size_t)blocks;
}

int srgan_density_label(const double* head_yx, int n_heads, int H, int W, double beta, float* label, void* label_f16,
                        void* workspace, size_t workspace_bytes, void* stream) {
    SRGAN_REQUIRE(head_yx && label && workspace, "srgan_density_label: null pointer");
    SRGAN_REQUIRE(n_heads > 0 && H > 0 && W > 0, "srgan_density_label: empty problem");
    SRGAN_REQUIRE(workspace_bytes >= srgan_density_label_workspace_bytes(n_heads, H, W), "srgan_density_label: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)H * W;
    const int blocks = (int)((total + kDensityThreads - 1) / kDensityThreads);
    double* partials = (double*)workspace;                                  // 8-byte aligned first, the head table behind it
    head_geom* geoms = (head_geom*)(partials + blocks);
    head_geometry_kernel<<<(n_heads + 127) / 128, 128, 0, st>>>(head_yx, n_heads, H, W, beta, geoms);
    SRGAN_CHECK_LAUNCH("srgan_density_label(geometry)");
    head_kernel_sum_kernel<<<n_heads, 128, 0, st>>>(geoms);
    SRGAN_CHECK_LAUNCH("srgan_density_label(kernel sums)");
    density_label_kernel<<<blocks, kDensityThreads, 0, st>>>(geoms, n_heads, H, W, label, partials);
    SRGAN_CHECK_LAUNCH("srgan_density_label(gather)");
    density_normalise_kernel<<<2 * kNumSMs, 256, 0, st>>>(label, total, partials, blocks, (float)n_heads, (__half*)label_f16);
    SRGAN_CHECK_LAUNCH("srgan_density_label(normalise)");
    return 0;
}

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
