// Crowd input pipeline and sliding-window inference on the device.
//
// The reference feeds the training step from a 4-worker DataLoader: every sample is an np.load(mmap) of a full image, a
// numpy crop / pad / flip / normalise on the host and a per-step host -> device copy (crowd/shanghai_tech_data.py:73-104,
// crowd/data.py:92-128,370-492, srgan.py:107-117).  Here the full images, density labels and kNN maps stay resident in
// HBM (ShanghaiTech part A: 300 images, ~0.7 GB as uint8 + 2 x 0.9 GB fp32) and a batch of patches is ONE gather launch
// driven by a [B,4] position table; nothing but the table crosses PCIe.  All of it is HBM-bound byte work: per 224x224
// patch 150 KB of uint8 and 2 x 200 KB of fp32 are read, 602 KB + 2 x 200 KB are written.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// ExtractPatchForPosition(allow_padded=True) -> RandomHorizontalFlip -> NegativeOneToOneNormalizeImage ->
// NumpyArraysToTorchTensors.  A thread owns 4 consecutive OUTPUT pixels of one patch row: 12 source bytes + 2 x 4 source
// floats in, five 16-byte stores out (three image planes, label, map).
// Patch row r / column c of a patch centred at (y, x) is source pixel (y - half + r, x - half + c): the reference pads the
// example so that the window exists (crowd/data.py:391-400) and shifts y / x by the top / left padding, which leaves this
// mapping unchanged; padding is the constant 0 for the image (BEFORE normalisation: -1 after it), the label and the map
// (:442-452).  The flip reverses the patch's columns (np.flip(axis=1), :105-107).
__global__ void __launch_bounds__(256)
extract_patches_kernel(const uint8_t* __restrict__ images, const float* __restrict__ labels, const float* __restrict__ maps,
                       const long long* __restrict__ pixel_offset, const int* __restrict__ heights,
                       const int* __restrict__ widths, const int* __restrict__ pos, int B, int P, float* __restrict__ img_out,
                       float* __restrict__ label_out, float* __restrict__ map_out) {
    const int quads = P >> 2;
    const long long total = (long long)B * P * quads;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(i % quads);
        const int r = (int)((i / quads) % P);
        const int b = (int)(i / ((long long)quads * P));
        const int4 p = reinterpret_cast<const int4*>(pos)[b];          // {image, y, x, flip}
        const int H = heights[p.x], W = widths[p.x];
        const long long base = pixel_offset[p.x];
        const int half = P >> 1;
        const int sy = p.y - half + r;
        float im[3][4], lb[4], mp[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = 4 * q + j;                                     // output column
            const int sc = p.w ? (P - 1 - c) : c;                       // patch column before the flip
            const int sx = p.z - half + sc;
            const bool in = sy >= 0 && sy < H && sx >= 0 && sx < W;
            const long long pix = base + (long long)sy * W + sx;
            unsigned v0 = 0, v1 = 0, v2 = 0;
            float l = 0.f, m = 0.f;
            if (in) {
                const uint8_t* s = images + 3 * pix;
                v0 = s[0], v1 = s[1], v2 = s[2];
                if (labels) l = labels[pix];
                if (maps) m = maps[pix];
            }
            // (image.astype(float32) / (255 / 2)) - 1 : IEEE fp32 division by 127.5f, then the subtraction (crowd/data.py:127)
            im[0][j] = __fsub_rn(__fdiv_rn((float)v0, 127.5f), 1.f);
            im[1][j] = __fsub_rn(__fdiv_rn((float)v1, 127.5f), 1.f);
            im[2][j] = __fsub_rn(__fdiv_rn((float)v2, 127.5f), 1.f);
            lb[j] = l;
            mp[j] = m;
        }
        const long long plane = (long long)P * P;
        const long long o = (long long)r * P + 4 * q;
        float* io = img_out + (long long)b * 3 * plane + o;
        st4(io, make_float4(im[0][0], im[0][1], im[0][2], im[0][3]));
        st4(io + plane, make_float4(im[1][0], im[1][1], im[1][2], im[1][3]));
        st4(io + 2 * plane, make_float4(im[2][0], im[2][1], im[2][2], im[2][3]));
        if (label_out) st4(label_out + (long long)b * plane + o, make_float4(lb[0], lb[1], lb[2], lb[3]));
        if (map_out) st4(map_out + (long long)b * plane + o, make_float4(mp[0], mp[1], mp[2], mp[3]));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// CrowdExperiment.predict_full_example (crowd/srgan.py:332-395): every patch adds its predicted density label and a
// constant count / patch^2 into the full-size sums over the rows / columns it covers and counts a hit; the result is the
// per-pixel mean.  As a GATHER: a thread owns one pixel of the full image and walks the (few) window positions that cover
// it, in patch-index order -- the order the reference's `+=` run in, so the per-pixel fp32 sums carry the same bits.
// Patch (yi, xi) has index yi * nx + xi (np.unravel_index, crowd/data.py:553) and covers rows [ys[yi]-half, ys[yi]+half).
constexpr int kMergeThreads = 256;

__global__ void __launch_bounds__(kMergeThreads)
sliding_merge_kernel(const float* __restrict__ patch_labels, const float* __restrict__ patch_counts,
                     const int* __restrict__ ys, int ny, const int* __restrict__ xs, int nx, int H, int W, int P,
                     float* __restrict__ full_label, double* __restrict__ partials) {
    __shared__ double red[kMergeThreads / 32];
    const int half = P >> 1;
    const float size = (float)(P * P);
    double local = 0.0;
    const long long total = (long long)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int Y = (int)(i / W), X = (int)(i % W);
        float sum_density = 0.f, sum_count = 0.f;
        int hits = 0;
        for (int yi = 0; yi < ny; ++yi) {
            const int r = Y - (ys[yi] - half);
            if (r < 0 || r >= P) continue;
            for (int xi = 0; xi < nx; ++xi) {
                const int c = X - (xs[xi] - half);
                if (c < 0 || c >= P) continue;
                const int p = yi * nx + xi;
                if (patch_labels) sum_density = __fadd_rn(sum_density, patch_labels[((long long)p * P + r) * P + c]);
                sum_count = __fadd_rn(sum_count, __fdiv_rn(patch_counts[p], size));        // np.full(.., count / label.size)
                ++hits;
            }
        }
        const float h = (float)(hits ? hits : 1);                                          // hit_predicted_label[.. == 0] = 1
        full_label[i] = __fdiv_rn(sum_density, h);
        local += (double)__fdiv_rn(sum_count, h);
    }
    // fixed-order block sum (warp shuffle tree, then the warps in order): partials[block]
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < kMergeThreads / 32; ++w) s += red[w];
        partials[blockIdx.x] = s;
    }
}

__global__ void sum_partials_kernel(const double* __restrict__ partials, int n, double* __restrict__ out, int n_out_stride) {
    // one warp per output: out[k] = sum_j partials[k * n_out_stride + j], j < n, in a fixed order
    const int k = blockIdx.x;
    double s = 0.0;
    for (int j = threadIdx.x; j < n; j += 32) s += partials[(long long)k * n_out_stride + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[k] = s;
}

// ---------------------------------------------------------------------------------------------------------------------
// CrowdExperiment.evaluation_epoch (crowd/srgan.py:149-191): the reductions behind ME / MAE / MSE of the counts and the kNN
// map MAE / MSE.  numpy accumulates them in float64 (the arrays are concatenated onto np.array([])), so the sums are
// double here too.  blockIdx.y = sample; out rows: [0] sum(density_b), [1] sum_c sum_hw |map_hat - map|, [2] the squares.
constexpr int kEvalThreads = 256;
constexpr int kEvalBlocksPerSample = 8;

__global__ void __launch_bounds__(kEvalThreads)
eval_sums_kernel(const float* __restrict__ densities, const float* __restrict__ pred_maps, int nmaps,
                 const float* __restrict__ maps, long long HW, double* __restrict__ partials) {
    __shared__ double red[3][kEvalThreads / 32];
    const int b = blockIdx.y;
    double s[3] = {0.0, 0.0, 0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
        if (densities) s[0] += (double)densities[(long long)b * HW + i];
        if (pred_maps) {
            const float m = maps[(long long)b * HW + i];
            for (int c = 0; c < nmaps; ++c) {
                const double d = fabs((double)pred_maps[((long long)b * nmaps + c) * HW + i] - (double)m);   // float64 arrays there
                s[1] += d;
                s[2] += d * d;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = s[k];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int w = 0; w < kEvalThreads / 32; ++w) t += red[threadIdx.x][w];
        partials[((long long)threadIdx.x * gridDim.y + b) * kEvalBlocksPerSample + blockIdx.x] = t;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Age / driving datasets (age/data.py:52-60, driving/data.py:44-51): a sample is one decoded 3 x S x S uint8 image, moved to
// CHW, cast to fp32 and mapped to [-1, 1] by utility.to_normalized_range ((x / 127.5) - 1), plus its scalar label.  With the
// whole decoded dataset resident (IMDB-WIKI at 128 x 128: 49 KB per image, 230 k images = 11 GB of the 180 GB) a batch is a
// gather by sample index.  A thread owns 4 consecutive output elements of one plane.
__global__ void __launch_bounds__(256)
image_batch_kernel(const uint8_t* __restrict__ images, int hwc, const long long* __restrict__ index, int B, int C, int H, int W,
                   float* __restrict__ out, const float* __restrict__ labels, float* __restrict__ labels_out) {
    const long long plane = (long long)H * W, per = (long long)C * plane;
    const long long total = (long long)B * per / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long e = 4 * i;
        const int b = (int)(e / per);
        const long long r = e % per;                       // offset inside the CHW sample
        const uint8_t* src = images + index[b] * per;
        float v[4];
        if (hwc) {
            const int c = (int)(r / plane);
            const long long px = r % plane;
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (float)src[(px + j) * C + c];
        } else {
            const uchar4 q = *reinterpret_cast<const uchar4*>(src + r);
            v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
        }
        st4(out + e, make_float4(__fsub_rn(__fdiv_rn(v[0], 127.5f), 1.f), __fsub_rn(__fdiv_rn(v[1], 127.5f), 1.f),
                                 __fsub_rn(__fdiv_rn(v[2], 127.5f), 1.f), __fsub_rn(__fdiv_rn(v[3], 127.5f), 1.f)));
        if (labels_out && r == 0) labels_out[b] = labels[index[b]];
    }
}

}  // namespace

extern "C" {

int srgan_image_batch(const uint8_t* images, int hwc, const long long* index, int B, int C, int H, int W, float* out,
                      const float* labels, float* labels_out, void* stream) {
    SRGAN_REQUIRE(images && index && out, "srgan_image_batch: null pointer");
    SRGAN_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "srgan_image_batch: empty batch");
    SRGAN_REQUIRE(((long long)H * W) % 4 == 0, "srgan_image_batch: H*W = %lld is not a multiple of 4", (long long)H * W);
    SRGAN_REQUIRE((labels != nullptr) == (labels_out != nullptr), "srgan_image_batch: labels and labels_out go together");
    const long long total = (long long)B * C * H * W / 4;
    const int blocks = (int)((total + 255) / 256 < 8LL * kNumSMs * 4 ? (total + 255) / 256 : 8LL * kNumSMs * 4);
    image_batch_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(images, hwc, index, B, C, H, W, out, labels, labels_out);
    SRGAN_CHECK_LAUNCH("srgan_image_batch");
    return 0;
}

int srgan_crowd_extract_patches(const uint8_t* images, const float* labels, const float* maps, const long long* pixel_offset,
                                const int* heights, const int* widths, int n_images, const int* pos, int B, int patch,
                                float* img_out, float* label_out, float* map_out, void* stream) {
    SRGAN_REQUIRE(images && pixel_offset && heights && widths && pos && img_out, "srgan_crowd_extract_patches: null pointer");
    SRGAN_REQUIRE(B > 0 && n_images > 0, "srgan_crowd_extract_patches: empty batch or store (B=%d, images=%d)", B, n_images);
    SRGAN_REQUIRE(patch > 0 && patch % 4 == 0, "srgan_crowd_extract_patches: patch size %d is not a multiple of 4", patch);
    SRGAN_REQUIRE((labels != nullptr) == (label_out != nullptr) && (maps != nullptr) == (map_out != nullptr),
                  "srgan_crowd_extract_patches: a label / map output needs its source and vice versa");
    const long long total = (long long)B * patch * (patch / 4);
    const int blocks = (int)((total + 255) / 256 < 8LL * kNumSMs * 4 ? (total + 255) / 256 : 8LL * kNumSMs * 4);
    extract_patches_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(images, labels, maps, pixel_offset, heights, widths, pos, B,
                                                                     patch, img_out, label_out, map_out);
    SRGAN_CHECK_LAUNCH("srgan_crowd_extract_patches");
    return 0;
}

size_t srgan_sliding_window_workspace_bytes(void) { return sizeof(double) * 4 * kNumSMs; }

int srgan_sliding_window_merge(const float* patch_labels, const float* patch_counts, const int* ys, int ny, const int* xs,
                               int nx, int H, int W, int patch, float* full_label, double* full_count, void* workspace,
                               size_t workspace_bytes, void* stream) {
    SRGAN_REQUIRE(patch_counts && ys && xs && full_label && full_count && workspace, "srgan_sliding_window_merge: null pointer");
    SRGAN_REQUIRE(ny > 0 && nx > 0 && H > 0 && W > 0 && patch > 0, "srgan_sliding_window_merge: empty problem");
    SRGAN_REQUIRE(workspace_bytes >= srgan_sliding_window_workspace_bytes(), "srgan_sliding_window_merge: workspace too small");
    const long long total = (long long)H * W;
    int blocks = (int)((total + kMergeThreads - 1) / kMergeThreads);
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    sliding_merge_kernel<<<blocks, kMergeThreads, 0, (cudaStream_t)stream>>>(patch_labels, patch_counts, ys, ny, xs, nx, H, W,
                                                                             patch, full_label, (double*)workspace);
    SRGAN_CHECK_LAUNCH("srgan_sliding_window_merge");
    sum_partials_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const double*)workspace, blocks, full_count, blocks);
    SRGAN_CHECK_LAUNCH("srgan_sliding_window_merge(sum)");
    return 0;
}

size_t srgan_crowd_eval_workspace_bytes(int n) { return sizeof(double) * 3 * (size_t)(n > 0 ? n : 0) * kEvalBlocksPerSample; }

int srgan_crowd_eval_sums(const float* densities, const float* pred_maps, int nmaps, const float* maps, int n, long long HW,
                          double* out, void* workspace, size_t workspace_bytes, void* stream) {
    SRGAN_REQUIRE(out && workspace, "srgan_crowd_eval_sums: null pointer");
    SRGAN_REQUIRE(n > 0 && HW > 0, "srgan_crowd_eval_sums: empty problem");
    SRGAN_REQUIRE((pred_maps != nullptr) == (maps != nullptr) && (!pred_maps || nmaps > 0),
                  "srgan_crowd_eval_sums: predicted maps need the map labels and a map count");
    SRGAN_REQUIRE(workspace_bytes >= srgan_crowd_eval_workspace_bytes(n), "srgan_crowd_eval_sums: workspace too small");
    dim3 grid(kEvalBlocksPerSample, n);
    eval_sums_kernel<<<grid, kEvalThreads, 0, (cudaStream_t)stream>>>(densities, pred_maps, nmaps, maps, HW, (double*)workspace);
    SRGAN_CHECK_LAUNCH("srgan_crowd_eval_sums");
    // out[k * n + b] = sum over the sample's blocks
    sum_partials_kernel<<<3 * n, 32, 0, (cudaStream_t)stream>>>((const double*)workspace, kEvalBlocksPerSample, out,
                                                                kEvalBlocksPerSample);
    SRGAN_CHECK_LAUNCH("srgan_crowd_eval_sums(sum)");
    return 0;
}

}  // extern "C"
