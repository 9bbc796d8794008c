// One persistent cooperative kernel for the coefficient-application training step
// (BASELINE configs[0]; coefficient/models.py:12-72 MLPs, srgan.py:259-320 + coefficient/dggan.py:22-64):
//   DNN step, discriminator step (labeled / unlabeled / fake losses, gradient penalty with its double backward),
//   generator step and the three Adam updates -- ~850 ATen launches in the reference, ~150 launches on this library's
//   generic kernels, ONE launch per step method here (`phases`: 1 = dnn_training_step, 2 = gan_training_step, 3 = both).
// Layout of the work: the three networks (2.4 k parameters) live in shared memory; every thread owns one sample per
// round and keeps only 10-wide vectors in registers; 50-wide vectors (inputs, fake, x_hat, dLoss/dinput) are streamed
// through the thread's own column of a shared staging tile ([slot][sample]: conflict-free, and 4 samples per LDS.128 in the
// outer-product sums); weight matrices sit in shared memory with rows padded to 12 floats (three LDS.128 per row).  Batch-wide sums (feature sums, losses) are per-CTA partials in
// a caller workspace, combined after a grid-wide sync in a fixed order (bit-reproducible); weight gradients are
// CTA-level outer-product sums over the staging tile (one table-driven loop for every tensor of a net), flushed with
// fp32 atomics into the engine's flat gradient buffers, which the in-kernel Adam reads and re-zeroes.
// Per-sample forwards are recomputed after a sync instead of being stored (a forward is 710 MACs).
// Arithmetic is fp32 in both precision modes (the layers are 10 wide: nothing for a tensor core to do).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int H = 10;          // hidden width
constexpr int NIN = 50;        // discriminator input = generator output
constexpr int NZ = 10;         // generator input
constexpr int CT = 64;         // threads (= samples per round) per CTA
constexpr float SLOPE = 0.01f;
constexpr int MAXG = kNumSMs;  // grid size bound (one CTA per SM)

// staging tile [slot][sample]: per sample, slots [0,50) = a 50-wide vector, then seven 10-wide groups, then 3 scalars.
constexpr int NSLOT = 123;
constexpr int HP = 12;         // padded row length of the shared-memory weight matrices (10 -> 12: three float4)
constexpr int O_IN = 0;
__host__ __device__ constexpr int O_S(int k) { return NIN + H * k; }
constexpr int O_HV0 = 120, O_HV1 = 121, O_ONE = 122;
// discriminator-shaped pass: dLoss/da_l and h_l of the three hidden layers, head row scales
constexpr int O_D1 = O_S(0), O_H1 = O_S(1), O_D2 = O_S(2), O_H2 = O_S(3), O_D3 = O_S(4), O_H3 = O_S(5);
// generator pass: IN = dLoss/d(G output)
constexpr int O_Z = O_S(0), O_E1 = O_S(1), O_G1 = O_S(2), O_E2 = O_S(3), O_G2 = O_S(4), O_E3 = O_S(5), O_G3 = O_S(6);

constexpr int KD = (730 + 11 * 2 + CT - 1) / CT;   // gradient elements per thread, discriminator-shaped net (<= 752)
constexpr int KG = (880 + CT - 1) / CT;            // generator (880)

// workspace (floats): per-CTA partials
constexpr int WS_FA = 0;                           // [MAXG][30] feature sums of x, u, fake (D step)
constexpr int WS_FC = WS_FA + MAXG * 30;           // [MAXG][20] feature sums of fake2, u (G step)
constexpr int WS_SC = WS_FC + MAXG * 20;           // [7][MAXG] scalar partials
constexpr int WS_FLOATS = WS_SC + 7 * MAXG;

struct Lin { float* W; float* b; float* gW; float* gb; float* mW; float* mb; float* vW; float* vb; };
struct NetP { Lin l[4]; float* state; };     // state: [t, lr/(1-b1^t), 1/sqrt(1-b2^t)] (srgan_adam_prepare layout)

struct CoefParams {
    NetP D, G, DNN;
    const float *x, *y, *u, *z, *alpha, *z2;
    int B;                      // local batch
    float inv_Bg;               // 1 / global batch
    int dggan, head_out;
    int order;
    float labeled_mult, unl_mult, fake_mult, gen_mult, gp_lambda;   // already including srgan/dggan multipliers
    int kind_match, kind_contrast;
    float lr, lr_dnn, wd, beta1, beta2, eps;
    int phases, train_g;
    float* ws;                  // WS_FLOATS floats, no initialisation required
    float* scalars;             // engine scalar slots
    // optional (may be null): [4][B][10] features of x | u | fake | x_hat as the D step saw them (the fake block is
    // replaced by the generator step's G(z2) under the updated discriminator), then [B] gradient norms: the tensors
    // srgan.py:332-386 leaves in self.*_features / self.gradient_norm
    float* publish;
};

__device__ __forceinline__ void publish10(float* pub, int block, int B, int sample, bool act, const float* h) {
    if (pub != nullptr && act) {
        float* d = pub + ((long long)block * B + sample) * H;
#pragma unroll
        for (int i = 0; i < H; ++i) d[i] = h[i];
    }
}

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : v * SLOPE; }
__device__ __forceinline__ float dleaky(float h) { return h > 0.f ? 1.f : SLOPE; }

// shared-memory copy of one 4-layer MLP.  Matrices are row-padded to HP floats; for discriminator-shaped nets W1 is stored
// TRANSPOSED ([input j][output o]): one padded row serves both W1 x and W1^T y at input j.
struct SW { const float *W1, *b1, *W2, *b2, *W3, *b3, *W4, *b4; };

// this thread's column of the staging tile: row[slot]
struct Row {
    float* p;
    __device__ __forceinline__ float& operator[](int i) const { return p[i * CT]; }
    __device__ __forceinline__ Row operator+(int off) const { return Row{p + off * CT}; }
};
__device__ __forceinline__ void ldrow(const float* __restrict__ w, float (&r)[HP]) {
    const float4 a = *reinterpret_cast<const float4*>(w), b = *reinterpret_cast<const float4*>(w + 4),
                 c = *reinterpret_cast<const float4*>(w + 8);
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
    r[8] = c.x; r[9] = c.y; r[10] = c.z; r[11] = c.w;
}

// y[o] = sum_i W[o][i] x[i] (+ b[o]);  W row-major [H][HP] in shared memory
__device__ __forceinline__ void mv(const float* __restrict__ W, const float* __restrict__ b, const float* x, float* y) {
#pragma unroll
    for (int o = 0; o < H; ++o) {
        float r[HP];
        ldrow(W + o * HP, r);
        float a = b ? b[o] : 0.f;
#pragma unroll
        for (int i = 0; i < H; ++i) a = fmaf(r[i], x[i], a);
        y[o] = a;
    }
}
// x[i] = sum_o W[o][i] y[o]
__device__ __forceinline__ void mvt(const float* __restrict__ W, const float* y, float* x) {
#pragma unroll
    for (int i = 0; i < H; ++i) x[i] = 0.f;
#pragma unroll
    for (int o = 0; o < H; ++o) {
        float r[HP];
        ldrow(W + o * HP, r);
        const float v = y[o];
#pragma unroll
        for (int i = 0; i < H; ++i) x[i] = fmaf(r[i], v, x[i]);
    }
}
__device__ __forceinline__ void leaky10(float* h) {
#pragma unroll
    for (int i = 0; i < H; ++i) h[i] = leaky(h[i]);
}
// hidden layers 2 and 3 of a discriminator-shaped net from the layer-1 pre-activation (in h1, activated in place)
__device__ __forceinline__ void d_tail(const SW& w, float* h1, float* h2, float* h3) {
    leaky10(h1);
    mv(w.W2, w.b2, h1, h2);
    leaky10(h2);
    mv(w.W3, w.b3, h2, h3);
    leaky10(h3);
}
// forward of a discriminator-shaped net on the 50-vector in `row` (shared memory)
__device__ __forceinline__ void d_fwd_row(const SW& w, const Row row, float* h1, float* h2, float* h3) {
#pragma unroll
    for (int o = 0; o < H; ++o) h1[o] = w.b1[o];
#pragma unroll 2
    for (int j = 0; j < NIN; ++j) {
        const float v = row[j];
        float r[HP];
        ldrow(w.W1 + j * HP, r);
#pragma unroll
        for (int o = 0; o < H; ++o) h1[o] = fmaf(r[o], v, h1[o]);
    }
    d_tail(w, h1, h2, h3);
}
// generator hidden layers: z -> g1, g2, g3 (the 50-wide output layer is streamed by the callers)
__device__ __forceinline__ void g_hidden(const SW& w, const float* zin, float* g1, float* g2, float* g3) {
    mv(w.W1, w.b1, zin, g1);        // NZ == H
    leaky10(g1);
    mv(w.W2, w.b2, g1, g2);
    leaky10(g2);
    mv(w.W3, w.b3, g2, g3);
    leaky10(g3);
}
// one element of the generator output (no activation on the last layer, coefficient/models.py:26)
__device__ __forceinline__ float g_out(const SW& w, const float* g3, int j) {
    float a = w.b4[j], r[HP];
    ldrow(w.W4 + j * HP, r);
#pragma unroll
    for (int i = 0; i < H; ++i) a = fmaf(r[i], g3[i], a);
    return a;
}
// D(G(z)) with the fake sample streamed: optionally stored into `row`
template <bool STORE>
__device__ __forceinline__ void d_fwd_fake(const SW& wd, const SW& wg, const float* g3, const Row row, float* h1, float* h2,
                                           float* h3) {
#pragma unroll
    for (int o = 0; o < H; ++o) h1[o] = wd.b1[o];
#pragma unroll 2
    for (int j = 0; j < NIN; ++j) {
        const float v = g_out(wg, g3, j);
        if (STORE) row[j] = v;
        float r[HP];
        ldrow(wd.W1 + j * HP, r);
#pragma unroll
        for (int o = 0; o < H; ++o) h1[o] = fmaf(r[o], v, h1[o]);
    }
    d_tail(wd, h1, h2, h3);
}
// dLoss/da_2, dLoss/da_1 from dLoss/da_3
__device__ __forceinline__ void d_bwd_hidden(const SW& w, const float* h1, const float* h2, const float* da3, float* da2,
                                             float* da1) {
    float t[H];
    mvt(w.W3, da3, t);
#pragma unroll
    for (int i = 0; i < H; ++i) da2[i] = t[i] * dleaky(h2[i]);
    mvt(w.W2, da2, t);
#pragma unroll
    for (int i = 0; i < H; ++i) da1[i] = t[i] * dleaky(h1[i]);
}
__device__ __forceinline__ void put10(const Row dst, const float* v, float k = 1.f) {
#pragma unroll
    for (int i = 0; i < H; ++i) dst[i] = k * v[i];
}

// labeled loss term and its derivative for one sample: srgan.py:414-417
__device__ __forceinline__ void labeled_term(float pred, float y, int order, float scale, float& loss, float& dpred) {
    const float d = pred - y, ad = fabsf(d), sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    float pw, dpw;
    if (order == 2) { pw = ad * ad; dpw = 2.f * ad; }
    else if (order == 1) { pw = ad; dpw = 1.f; }
    else { pw = powf(ad, (float)order); dpw = order * powf(ad, (float)(order - 1)); }
    loss = scale * pw;
    dpred = scale * dpw * sg;
}
// BCE-with-logits against a constant target: coefficient/dggan.py:39-40,48-49,62-63
__device__ __forceinline__ void bce_term(float s, float target, float scale, float& loss, float& ds) {
    loss = scale * (fmaxf(s, 0.f) - s * target + log1pf(expf(-fabsf(s))));
    ds = scale * (1.f / (1.f + expf(-s)) - target);
}

// distance function of utility.py:201-243 on a 10-vector of mean differences: returns loss, writes dloss/dd
__device__ __forceinline__ float distance10(const float* d, int kind, float* g) {
    float acc = 0.f;
    const float invF = 1.f / (float)H;
    if (kind == 5) {
#pragma unroll
        for (int i = 0; i < H; ++i) acc = fmaf(d[i], d[i], acc);
        const float nrm = sqrtf(acc);
#pragma unroll
        for (int i = 0; i < H; ++i) g[i] = nrm > 0.f ? d[i] / nrm : 0.f;
        return nrm;
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const float v = d[i], ad = fabsf(v);
        const float sg = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);
        if (kind == 0) { acc += ad; g[i] = sg * invF; }
        else if (kind == 1) { acc -= ad; g[i] = -sg * invF; }
        else if (kind == 2) { const float r = sqrtf(ad + 1.f); acc -= r; g[i] = -sg * invF * 0.5f / r; }
        else if (kind == 3) { acc -= logf(ad + 1.f); g[i] = -sg * invF / (ad + 1.f); }
        else { acc += v * v; g[i] = 2.f * v * invF; }
    }
    return acc * invF;
}

// ---- gradient element tables: element e of a net's gradient = sum over the tile rows of row[oa] * row[ob]
__device__ __forceinline__ int d_elem(int e, int HO) {       // discriminator-shaped net, returns oa | ob << 8
    int oa, ob;
    if (e < 500) { oa = O_D1 + e / NIN; ob = O_IN + e % NIN; }
    else if (e < 510) { oa = O_D1 + (e - 500); ob = O_ONE; }
    else if (e < 610) { oa = O_D2 + (e - 510) / H; ob = O_H1 + (e - 510) % H; }
    else if (e < 620) { oa = O_D2 + (e - 610); ob = O_ONE; }
    else if (e < 720) { oa = O_D3 + (e - 620) / H; ob = O_H2 + (e - 620) % H; }
    else if (e < 730) { oa = O_D3 + (e - 720); ob = O_ONE; }
    else if (e < 730 + H * HO) { oa = O_HV0 + (e - 730) / H; ob = O_H3 + (e - 730) % H; }
    else { oa = O_HV0 + (e - 730 - H * HO); ob = O_ONE; }
    return oa | (ob << 8);
}
__device__ __forceinline__ float* d_grad_ptr(const NetP& n, int e, int HO) {
    if (e < 500) return n.l[0].gW + e;
    if (e < 510) return n.l[0].gb + (e - 500);
    if (e < 610) return n.l[1].gW + (e - 510);
    if (e < 620) return n.l[1].gb + (e - 610);
    if (e < 720) return n.l[2].gW + (e - 620);
    if (e < 730) return n.l[2].gb + (e - 720);
    if (e < 730 + H * HO) return n.l[3].gW + (e - 730);
    return n.l[3].gb + (e - 730 - H * HO);
}
__device__ __forceinline__ int g_elem(int e) {
    int oa, ob;
    if (e < 100) { oa = O_E1 + e / NZ; ob = O_Z + e % NZ; }
    else if (e < 110) { oa = O_E1 + (e - 100); ob = O_ONE; }
    else if (e < 210) { oa = O_E2 + (e - 110) / H; ob = O_G1 + (e - 110) % H; }
    else if (e < 220) { oa = O_E2 + (e - 210); ob = O_ONE; }
    else if (e < 320) { oa = O_E3 + (e - 220) / H; ob = O_G2 + (e - 220) % H; }
    else if (e < 330) { oa = O_E3 + (e - 320); ob = O_ONE; }
    else if (e < 830) { oa = O_IN + (e - 330) / H; ob = O_G3 + (e - 330) % H; }
    else { oa = O_IN + (e - 830); ob = O_ONE; }
    return oa | (ob << 8);
}
__device__ __forceinline__ float* g_grad_ptr(const NetP& n, int e) {
    if (e < 100) return n.l[0].gW + e;
    if (e < 110) return n.l[0].gb + (e - 100);
    if (e < 210) return n.l[1].gW + (e - 110);
    if (e < 220) return n.l[1].gb + (e - 210);
    if (e < 320) return n.l[2].gW + (e - 220);
    if (e < 330) return n.l[2].gb + (e - 320);
    if (e < 830) return n.l[3].gW + (e - 330);
    return n.l[3].gb + (e - 830);
}

// acc[k] += sum over the CT rows of the tile of row[oa_k] * row[ob_k]   (barriers on both sides)
template <int K>
__device__ __forceinline__ void accumulate(const float* stage, const int (&pk)[K], float (&acc)[K], int nelem) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if ((int)threadIdx.x + k * CT < nelem) {
            const float4* pa = reinterpret_cast<const float4*>(stage + (pk[k] & 0xff) * CT);
            const float4* pb = reinterpret_cast<const float4*>(stage + (pk[k] >> 8) * CT);
            float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
            for (int s = 0; s < CT / 4; s += 2) {
                const float4 x0 = pa[s], y0 = pb[s], x1 = pa[s + 1], y1 = pb[s + 1];
                a0 += x0.x * y0.x + x0.y * y0.y + x0.z * y0.z + x0.w * y0.w;
                a1 += x1.x * y1.x + x1.y * y1.y + x1.z * y1.z + x1.w * y1.w;
            }
            acc[k] += a0 + a1;
        }
    }
    __syncthreads();
}

// this thread's sample of a [B][NIN] global tensor -> the IN slots of its staging column (rows are 200 bytes: float2 loads);
// threads past the batch end get zeros (they are masked everywhere)
__device__ __forceinline__ void row_load(const Row row, const float* __restrict__ src, long long sample, bool active) {
    const float2* p = reinterpret_cast<const float2*>(src + sample * NIN);
#pragma unroll 5
    for (int q = 0; q < NIN / 2; ++q) {
        const float2 v = active ? __ldg(p + q) : make_float2(0.f, 0.f);
        row[O_IN + 2 * q] = v.x;
        row[O_IN + 2 * q + 1] = v.y;
    }
}

// shared-memory layout of a net (floats): W1 | b1 | W2 | b2 | W3 | b3 | W4 | b4, matrices row-padded to HP, vectors to 4
__host__ __device__ constexpr int pad4(int n) { return (n + 3) / 4 * 4; }
__host__ __device__ constexpr int net_floats(int l1_in, int l4_out, bool w1_transposed) {
    return (w1_transposed ? l1_in : H) * HP + pad4(H) + 2 * (H * HP + pad4(H)) + l4_out * HP + pad4(l4_out);
}
// dst[r * HP + c] = src[r][c] (or src[c][r] when `transposed`), the pad columns are zero
__device__ __forceinline__ void load_mat(float* dst, const float* src, int rows, int cols, bool transposed) {
    for (int i = threadIdx.x; i < rows * HP; i += CT) {
        const int r = i / HP, c = i % HP;
        dst[i] = c < cols ? __ldcg(transposed ? src + c * rows + r : src + r * cols + c) : 0.f;   // L2: other CTAs update them
    }
}
__device__ __forceinline__ void load_vec(float* dst, const float* src, int n) {
    for (int i = threadIdx.x; i < pad4(n); i += CT) dst[i] = i < n ? __ldcg(src + i) : 0.f;
}
__device__ __forceinline__ SW load_net(float* base, const NetP& n, int l1_in, int l4_out, bool w1_transposed) {
    float* q = base;
    SW w;
    w.W1 = q;
    if (w1_transposed) { load_mat(q, n.l[0].W, l1_in, H, true); q += l1_in * HP; }     // [input][output]
    else { load_mat(q, n.l[0].W, H, l1_in, false); q += H * HP; }
    w.b1 = q; load_vec(q, n.l[0].b, H); q += pad4(H);
    w.W2 = q; load_mat(q, n.l[1].W, H, H, false); q += H * HP;
    w.b2 = q; load_vec(q, n.l[1].b, H); q += pad4(H);
    w.W3 = q; load_mat(q, n.l[2].W, H, H, false); q += H * HP;
    w.b3 = q; load_vec(q, n.l[2].b, H); q += pad4(H);
    w.W4 = q; load_mat(q, n.l[3].W, l4_out, H, false); q += l4_out * HP;
    w.b4 = q; load_vec(q, n.l[3].b, l4_out);
    return w;
}

// Adam for one flat tensor (torch.optim.Adam semantics); the whole GRID strides over it; grad re-zeroed
__device__ __forceinline__ void adam_tensor(float* p, float* g, float* m, float* v, int n, float step_size, float isb,
                                            float b1, float b2, float eps, float wd, int gtid, int gthreads) {
    for (int i = gtid; i < n; i += gthreads) {
        float gr = __ldcg(g + i);
        const float pv = p[i];
        if (wd != 0.f) gr = fmaf(wd, pv, gr);
        const float mm = b1 * m[i] + (1.f - b1) * gr;
        const float vv = b2 * v[i] + (1.f - b2) * gr * gr;
        m[i] = mm; v[i] = vv;
        p[i] = pv - step_size * (mm / (sqrtf(vv) * isb + eps));
        g[i] = 0.f;
    }
}
// t = the step number this update is (1-based); bias corrections in double like torch does on the host
__device__ void adam_net(const NetP& n, int l1_in, int l4_out, float t, double lr, float b1, float b2, float eps, float wd,
                         int gtid, int gthreads, float* bc /* shared [2] */) {
    __syncthreads();
    if (threadIdx.x == 0) {
        bc[0] = (float)(lr / (1.0 - pow((double)b1, (double)t)));
        bc[1] = (float)(1.0 / sqrt(1.0 - pow((double)b2, (double)t)));
    }
    __syncthreads();
    const float step_size = bc[0], isb = bc[1];
    const int szW[4] = {H * l1_in, H * H, H * H, l4_out * H};
    const int szb[4] = {H, H, H, l4_out};
    for (int l = 0; l < 4; ++l) {
        adam_tensor(n.l[l].W, n.l[l].gW, n.l[l].mW, n.l[l].vW, szW[l], step_size, isb, b1, b2, eps, wd, gtid, gthreads);
        adam_tensor(n.l[l].b, n.l[l].gb, n.l[l].mb, n.l[l].vb, szb[l], step_size, isb, b1, b2, eps, wd, gtid, gthreads);
    }
}

// per-CTA partial of a per-thread scalar -> ws slot; combined by combine_scalar after a grid sync
__device__ __forceinline__ void scalar_partial(float v, float* red, float* ws, int slot) {
    v = block_sum(v, red);
    if (threadIdx.x == 0) ws[WS_SC + slot * MAXG + blockIdx.x] = v;
}
__device__ __forceinline__ void combine_scalars(const float* ws, float* scalars, unsigned mask) {
    if (blockIdx.x == 0 && threadIdx.x < 7 && ((mask >> threadIdx.x) & 1u)) {
        float a = 0.f;
        for (int b = 0; b < (int)gridDim.x; ++b) a += __ldcg(ws + WS_SC + threadIdx.x * MAXG + b);
        scalars[threadIdx.x] = a;
    }
}

__global__ void __launch_bounds__(CT) coef_step_kernel(const CoefParams P) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) float smem[];
    __shared__ float red[32];
    __shared__ float fsum[30];
    __shared__ float bc[2];
    const int HO = P.head_out;
    const int tid = threadIdx.x;
    const int gtid = blockIdx.x * CT + tid, gthreads = gridDim.x * CT;
    const int rounds = (P.B + gthreads - 1) / gthreads;
    const bool do_dnn = P.phases & 1, do_gan = P.phases & 2, dggan = P.dggan != 0;
    const int ND = 730 + 11 * HO;

    // shared-memory carve-up: DNN | D | G weights, then the staging tile
    float* wDNN = smem;
    float* wD = wDNN + net_floats(NIN, 2, true);
    float* wG = wD + net_floats(NIN, 2, true);
    float* stage = wG + net_floats(NZ, NIN, false);
    const Row row{stage + tid};
    SW sDNN = {}, sD = {}, sG = {};
    if (do_dnn) sDNN = load_net(wDNN, P.DNN, NIN, HO, true);
    if (do_gan) { sD = load_net(wD, P.D, NIN, HO, true); sG = load_net(wG, P.G, NZ, NIN, false); }
    // step numbers of the updates this launch performs (read before anybody advances them)
    const float tD = P.D.state[0] + 1.f, tG = P.G.state[0] + 1.f, tDNN = P.DNN.state[0] + 1.f;
    __syncthreads();
    const float* W4r1 = sD.W4 + HP;                // DG-GAN fake-score head row

    int pkD[KD];
#pragma unroll
    for (int k = 0; k < KD; ++k) pkD[k] = d_elem(min(tid + k * CT, ND - 1), HO);
    const float lscale = P.labeled_mult * P.inv_Bg;

    // =========================================================================================== phase A
    // DNN step (srgan.py:259-271) and the feature sums of the discriminator forwards (srgan.py:329-358)
    {
        float acc[KD];
#pragma unroll
        for (int k = 0; k < KD; ++k) acc[k] = 0.f;
        float loss_dnn = 0.f, fs0[H], fs1[H], fs2[H];
#pragma unroll
        for (int i = 0; i < H; ++i) fs0[i] = fs1[i] = fs2[i] = 0.f;
        const bool sums = do_gan && !dggan;
        for (int rd = 0; rd < rounds; ++rd) {
            const int first = rd * gthreads + blockIdx.x * CT;
            const bool act = first + tid < P.B;
            const int sc = act ? first + tid : 0;
            const float k = act ? 1.f : 0.f;
            float h1[H], h2[H], h3[H];
            row_load(row, P.x, sc, act);
            if (do_dnn) {
                d_fwd_row(sDNN, row, h1, h2, h3);
                float pred = sDNN.b4[0];
#pragma unroll
                for (int i = 0; i < H; ++i) pred = fmaf(sDNN.W4[i], h3[i], pred);
                float lt, dpred;
                labeled_term(pred, P.y[sc], P.order, lscale, lt, dpred);
                loss_dnn += k * lt;
                dpred *= k;
                float da3[H], da2[H], da1[H];
#pragma unroll
                for (int i = 0; i < H; ++i) da3[i] = dpred * sDNN.W4[i] * dleaky(h3[i]);
                d_bwd_hidden(sDNN, h1, h2, da3, da2, da1);
                put10(row + O_D1, da1); put10(row + O_H1, h1); put10(row + O_D2, da2); put10(row + O_H2, h2);
                put10(row + O_D3, da3); put10(row + O_H3, h3);
                row[O_HV0] = dpred; row[O_HV1] = 0.f; row[O_ONE] = 1.f;
                accumulate<KD>(stage, pkD, acc, ND);
            }
            if (sums) {
                d_fwd_row(sD, row, h1, h2, h3);
#pragma unroll
                for (int i = 0; i < H; ++i) fs0[i] += k * h3[i];
                row_load(row, P.u, sc, act);
                d_fwd_row(sD, row, h1, h2, h3);
#pragma unroll
                for (int i = 0; i < H; ++i) fs1[i] += k * h3[i];
                float zz[NZ], g1[H], g2[H], g3[H];
#pragma unroll
                for (int i = 0; i < NZ; ++i) zz[i] = P.z[(long long)sc * NZ + i];
                g_hidden(sG, zz, g1, g2, g3);
                d_fwd_fake<false>(sD, sG, g3, row, h1, h2, h3);
#pragma unroll
                for (int i = 0; i < H; ++i) fs2[i] += k * h3[i];
            }
        }
        if (do_dnn) {
#pragma unroll
            for (int k = 0; k < KD; ++k) {
                const int e = tid + k * CT;
                if (e < ND) atomicAdd(d_grad_ptr(P.DNN, e, HO), acc[k]);
            }
            scalar_partial(loss_dnn, red, P.ws, 0);
        }
        if (sums) {
            put10(row + O_S(0), fs0); put10(row + O_S(1), fs1); put10(row + O_S(2), fs2);
            __syncthreads();
            if (tid < 30) {
                float a = 0.f;
                for (int s = 0; s < CT; ++s) a += stage[(O_S(0) + tid) * CT + s];
                P.ws[WS_FA + blockIdx.x * 30 + tid] = a;
            }
        }
    }
    __threadfence();
    grid.sync();
    if (do_dnn) combine_scalars(P.ws, P.scalars, 1u);
    if (do_gan && !dggan) {
        if (tid < 30) {
            float a = 0.f;
            for (int b = 0; b < (int)gridDim.x; ++b) a += __ldcg(P.ws + WS_FA + b * 30 + tid);
            fsum[tid] = a;
        }
        __syncthreads();
    }

    // =========================================================================================== phase B
    // discriminator step: losses, seeds, gradient penalty, one backward; gradients -> global
    if (do_gan) {
        float gx[H], gu[H], gf[H];
#pragma unroll
        for (int i = 0; i < H; ++i) gx[i] = gu[i] = gf[i] = 0.f;
        if (!dggan) {
            float d[H], g[H];
#pragma unroll
            for (int i = 0; i < H; ++i) d[i] = (fsum[H + i] - fsum[i]) * P.inv_Bg;                // mean_u - mean_x
            const float l_unl = P.unl_mult * distance10(d, P.kind_match, g);
#pragma unroll
            for (int i = 0; i < H; ++i) { gu[i] = P.unl_mult * P.inv_Bg * g[i]; gx[i] = -gu[i]; }
#pragma unroll
            for (int i = 0; i < H; ++i) d[i] = (fsum[H + i] - fsum[2 * H + i]) * P.inv_Bg;        // mean_u - mean_fake
            const float l_fake = P.fake_mult * distance10(d, P.kind_contrast, g);
#pragma unroll
            for (int i = 0; i < H; ++i) { gu[i] += P.fake_mult * P.inv_Bg * g[i]; gf[i] = -P.fake_mult * P.inv_Bg * g[i]; }
            if (gtid == 0) { P.scalars[2] = l_unl; P.scalars[3] = l_fake; }
        }
        float acc[KD];
#pragma unroll
        for (int k = 0; k < KD; ++k) acc[k] = 0.f;
        float l_lab = 0.f, l_pen = 0.f, l_gn = 0.f, l_unl_s = 0.f, l_fake_s = 0.f;
        const float lam = P.gp_lambda * P.inv_Bg;
        for (int rd = 0; rd < rounds; ++rd) {
            const int first = rd * gthreads + blockIdx.x * CT;
            const bool act = first + tid < P.B;
            const int sc = act ? first + tid : 0;
            const float k = act ? 1.f : 0.f;
            float h1[H], h2[H], h3[H], da3[H], da2[H], da1[H];
            // ---- x: labeled loss + matching seed
            row_load(row, P.x, sc, act);
            d_fwd_row(sD, row, h1, h2, h3);
            publish10(P.publish, 0, P.B, sc, act, h3);
            {
                float pred = sD.b4[0];
#pragma unroll
                for (int i = 0; i < H; ++i) pred = fmaf(sD.W4[i], h3[i], pred);
                float lt, dpred;
                labeled_term(pred, P.y[sc], P.order, lscale, lt, dpred);
                l_lab += k * lt;
#pragma unroll
                for (int i = 0; i < H; ++i) da3[i] = k * (gx[i] + dpred * sD.W4[i]) * dleaky(h3[i]);
                row[O_HV0] = k * dpred; row[O_HV1] = 0.f; row[O_ONE] = 1.f;
            }
            d_bwd_hidden(sD, h1, h2, da3, da2, da1);
            put10(row + O_D1, da1); put10(row + O_H1, h1); put10(row + O_D2, da2); put10(row + O_H2, h2);
            put10(row + O_D3, da3); put10(row + O_H3, h3);
            accumulate<KD>(stage, pkD, acc, ND);
            // ---- u
            row_load(row, P.u, sc, act);
            d_fwd_row(sD, row, h1, h2, h3);
            publish10(P.publish, 1, P.B, sc, act, h3);
            {
                float dsu = 0.f;
                if (dggan) {
                    float s1 = sD.b4[1], lt;
#pragma unroll
                    for (int i = 0; i < H; ++i) s1 = fmaf(W4r1[i], h3[i], s1);
                    bce_term(s1, 0.f, P.unl_mult * P.inv_Bg, lt, dsu);
                    l_unl_s += k * lt;
                    dsu *= k;
#pragma unroll
                    for (int i = 0; i < H; ++i) da3[i] = dsu * W4r1[i] * dleaky(h3[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < H; ++i) da3[i] = k * gu[i] * dleaky(h3[i]);
                }
                row[O_HV0] = 0.f; row[O_HV1] = dsu; row[O_ONE] = 1.f;
            }
            d_bwd_hidden(sD, h1, h2, da3, da2, da1);
            put10(row + O_D1, da1); put10(row + O_H1, h1); put10(row + O_D2, da2); put10(row + O_H2, h2);
            put10(row + O_D3, da3); put10(row + O_H3, h3);
            accumulate<KD>(stage, pkD, acc, ND);
            // ---- fake = G(z), stored in the IN slot (no gradient to G here: srgan.py:352 detaches, the DG-GAN G
            //      gradients of this pass are zeroed at srgan.py:300 before they are used)
            {
                float zz[NZ], g1[H], g2[H], g3[H];
#pragma unroll
                for (int i = 0; i < NZ; ++i) zz[i] = P.z[(long long)sc * NZ + i];
                g_hidden(sG, zz, g1, g2, g3);
                d_fwd_fake<true>(sD, sG, g3, row + O_IN, h1, h2, h3);
                publish10(P.publish, 2, P.B, sc, act, h3);
                float dsf = 0.f;
                if (dggan) {
                    float s1 = sD.b4[1], lt;
#pragma unroll
                    for (int i = 0; i < H; ++i) s1 = fmaf(W4r1[i], h3[i], s1);
                    bce_term(s1, 1.f, P.fake_mult * P.inv_Bg, lt, dsf);
                    l_fake_s += k * lt;
                    dsf *= k;
#pragma unroll
                    for (int i = 0; i < H; ++i) da3[i] = dsf * W4r1[i] * dleaky(h3[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < H; ++i) da3[i] = k * gf[i] * dleaky(h3[i]);
                }
                row[O_HV0] = 0.f; row[O_HV1] = dsf; row[O_ONE] = 1.f;
            }
            d_bwd_hidden(sD, h1, h2, da3, da2, da1);
            put10(row + O_D1, da1); put10(row + O_H1, h1); put10(row + O_D2, da2); put10(row + O_H2, h2);
            put10(row + O_D3, da3); put10(row + O_H3, h3);
            accumulate<KD>(stage, pkD, acc, ND);
            // ---- gradient penalty on x_hat = alpha*u + (1-alpha)*fake (srgan.py:360-375); IN still holds fake
            {
                const float a = P.alpha[sc];
#pragma unroll
                for (int o = 0; o < H; ++o) h1[o] = sD.b1[o];
#pragma unroll 2
                for (int j = 0; j < NIN; ++j) {
                    const float v = a * P.u[(long long)sc * NIN + j] + (1.f - a) * row[O_IN + j];
                    row[O_IN + j] = v;
                    float r[HP];
                    ldrow(sD.W1 + j * HP, r);
#pragma unroll
                    for (int o = 0; o < H; ++o) h1[o] = fmaf(r[o], v, h1[o]);
                }
                d_tail(sD, h1, h2, h3);
                publish10(P.publish, 3, P.B, sc, act, h3);
                float g3v[H], gm3[H], gm2[H], gm1[H], t[H];
                float snorm = 1.f, inv_s = 1.f;
                if (dggan) {
#pragma unroll
                    for (int i = 0; i < H; ++i) g3v[i] = W4r1[i];
                } else {
                    float ss = 0.f;
#pragma unroll
                    for (int i = 0; i < H; ++i) ss = fmaf(h3[i], h3[i], ss);
                    snorm = sqrtf(ss);
                    inv_s = snorm > 0.f ? 1.f / snorm : 0.f;
#pragma unroll
                    for (int i = 0; i < H; ++i) g3v[i] = h3[i] * inv_s;
                }
#pragma unroll
                for (int i = 0; i < H; ++i) gm3[i] = g3v[i] * dleaky(h3[i]);
                d_bwd_hidden(sD, h1, h2, gm3, gm2, gm1);
                // g0 = W1^T gm1 streamed: r = ||g0||, t = W1 g0
                float rr = 0.f;
#pragma unroll
                for (int o = 0; o < H; ++o) t[o] = 0.f;
#pragma unroll 2
                for (int j = 0; j < NIN; ++j) {
                    float g0 = 0.f, r[HP];
                    ldrow(sD.W1 + j * HP, r);
#pragma unroll
                    for (int o = 0; o < H; ++o) g0 = fmaf(r[o], gm1[o], g0);
                    rr = fmaf(g0, g0, rr);
#pragma unroll
                    for (int o = 0; o < H; ++o) t[o] = fmaf(r[o], g0, t[o]);
                }
                rr = sqrtf(rr);
                if (P.publish != nullptr && act) P.publish[(long long)4 * P.B * H + sc] = rr;
                const float ex = fmaxf(rr - 1.f, 0.f);
                l_pen += k * lam * ex * ex;
                l_gn += k * P.inv_Bg * rr;
                const float coef = (act && rr > 0.f) ? 2.f * lam * ex / rr : 0.f;
                // tangent chain u_l = (W_l u_{l-1}) * act'(h_l), u_0 = coef * g0
                float u1[H], u2[H], u3[H];
#pragma unroll
                for (int i = 0; i < H; ++i) u1[i] = coef * t[i] * dleaky(h1[i]);
                mv(sD.W2, nullptr, u1, u2);
#pragma unroll
                for (int i = 0; i < H; ++i) u2[i] *= dleaky(h2[i]);
                mv(sD.W3, nullptr, u2, u3);
#pragma unroll
                for (int i = 0; i < H; ++i) u3[i] *= dleaky(h3[i]);
                if (!dggan) {
                    // ordinary backward of the x_hat forward seeded with the Jacobian of f/||f|| applied to u_L
                    float dot = 0.f;
#pragma unroll
                    for (int i = 0; i < H; ++i) dot = fmaf(g3v[i], u3[i], dot);
#pragma unroll
                    for (int i = 0; i < H; ++i) da3[i] = (u3[i] - g3v[i] * dot) * inv_s * dleaky(h3[i]);
                    d_bwd_hidden(sD, h1, h2, da3, da2, da1);
                    put10(row + O_D1, da1); put10(row + O_H1, h1); put10(row + O_D2, da2); put10(row + O_H2, h2);
                    put10(row + O_D3, da3); put10(row + O_H3, h3);
                    row[O_HV0] = 0.f; row[O_HV1] = 0.f; row[O_ONE] = 1.f;
                    accumulate<KD>(stage, pkD, acc, ND);
                }
                // tangent block: weight gradients wgrad(u_{l-1}, gamma_l), no bias term (ONE = 0); the DG-GAN head row
                // gets sum_n u_L,n (the target is linear in the features: the Jacobian term vanishes)
#pragma unroll 2
                for (int j = 0; j < NIN; ++j) {
                    float g0 = 0.f, r[HP];
                    ldrow(sD.W1 + j * HP, r);
#pragma unroll
                    for (int o = 0; o < H; ++o) g0 = fmaf(r[o], gm1[o], g0);
                    row[O_IN + j] = g0;
                }
                put10(row + O_D1, gm1, coef); put10(row + O_H1, u1); put10(row + O_D2, gm2, k); put10(row + O_H2, u2);
                put10(row + O_D3, gm3, k); put10(row + O_H3, u3);
                row[O_HV0] = 0.f; row[O_HV1] = (dggan && act) ? 1.f : 0.f; row[O_ONE] = 0.f;
                accumulate<KD>(stage, pkD, acc, ND);
            }
        }
#pragma unroll
        for (int k = 0; k < KD; ++k) {
            const int e = tid + k * CT;
            if (e < ND) atomicAdd(d_grad_ptr(P.D, e, HO), acc[k]);
        }
        scalar_partial(l_lab, red, P.ws, 1);
        scalar_partial(l_pen, red, P.ws, 4);
        scalar_partial(l_gn, red, P.ws, 5);
        if (dggan) { scalar_partial(l_unl_s, red, P.ws, 2); scalar_partial(l_fake_s, red, P.ws, 3); }
    }
    __threadfence();
    grid.sync();
    if (do_gan) combine_scalars(P.ws, P.scalars, dggan ? 0x3Eu : 0x32u);

    // =========================================================================================== Adam: D and DNN
    if (do_gan) adam_net(P.D, NIN, HO, tD, (double)P.lr, P.beta1, P.beta2, P.eps, P.wd, gtid, gthreads, bc);
    if (do_dnn) adam_net(P.DNN, NIN, HO, tDNN, (double)P.lr_dnn, P.beta1, P.beta2, P.eps, P.wd, gtid, gthreads, bc);
    if (gtid == 0) { if (do_gan) P.D.state[0] = tD; if (do_dnn) P.DNN.state[0] = tDNN; }
    if (!do_gan || !P.train_g) return;
    __threadfence();
    grid.sync();

    // =========================================================================================== phase C: G step, sums
    sD = load_net(wD, P.D, NIN, HO, true);         // the UPDATED discriminator
    __syncthreads();
    if (!dggan) {
        float fs0[H], fs1[H];
#pragma unroll
        for (int i = 0; i < H; ++i) fs0[i] = fs1[i] = 0.f;
        for (int rd = 0; rd < rounds; ++rd) {
            const int first = rd * gthreads + blockIdx.x * CT;
            const bool act = first + tid < P.B;
            const int sc = act ? first + tid : 0;
            const float k = act ? 1.f : 0.f;
            float zz[NZ], g1[H], g2[H], g3[H], h1[H], h2[H], h3[H];
#pragma unroll
            for (int i = 0; i < NZ; ++i) zz[i] = P.z2[(long long)sc * NZ + i];
            g_hidden(sG, zz, g1, g2, g3);
            d_fwd_fake<false>(sD, sG, g3, row, h1, h2, h3);
#pragma unroll
            for (int i = 0; i < H; ++i) fs0[i] += k * h3[i];
            row_load(row, P.u, sc, act);
            d_fwd_row(sD, row, h1, h2, h3);
#pragma unroll
            for (int i = 0; i < H; ++i) fs1[i] += k * h3[i];
        }
        put10(row + O_S(0), fs0); put10(row + O_S(1), fs1);
        __syncthreads();
        if (tid < 20) {
            float a = 0.f;
            for (int s = 0; s < CT; ++s) a += stage[(O_S(0) + tid) * CT + s];
            P.ws[WS_FC + blockIdx.x * 20 + tid] = a;
        }
        __threadfence();
        grid.sync();
        if (tid < 20) {
            float a = 0.f;
            for (int b = 0; b < (int)gridDim.x; ++b) a += __ldcg(P.ws + WS_FC + b * 20 + tid);
            fsum[tid] = a;
        }
        __syncthreads();
    }

    // =========================================================================================== phase D: G backward
    {
        float gf2[H];
#pragma unroll
        for (int i = 0; i < H; ++i) gf2[i] = 0.f;
        if (!dggan) {
            float d[H], g[H];
#pragma unroll
            for (int i = 0; i < H; ++i) d[i] = (fsum[H + i] - fsum[i]) * P.inv_Bg;                // mean_u - mean_fake2
            const float l_gen = P.gen_mult * distance10(d, P.kind_match, g);
#pragma unroll
            for (int i = 0; i < H; ++i) gf2[i] = -P.gen_mult * P.inv_Bg * g[i];
            if (gtid == 0) P.scalars[6] = l_gen;
        }
        int pkG[KG];
        float acc[KG];
#pragma unroll
        for (int k = 0; k < KG; ++k) { pkG[k] = g_elem(min(tid + k * CT, 879)); acc[k] = 0.f; }
        float l_gen_s = 0.f;
        for (int rd = 0; rd < rounds; ++rd) {
            const int first = rd * gthreads + blockIdx.x * CT;
            const bool act = first + tid < P.B;
            const int sc = act ? first + tid : 0;
            const float k = act ? 1.f : 0.f;
            float zz[NZ], g1[H], g2[H], g3[H], h1[H], h2[H], h3[H], da3[H], da2[H], da1[H];
#pragma unroll
            for (int i = 0; i < NZ; ++i) zz[i] = P.z2[(long long)sc * NZ + i];
            g_hidden(sG, zz, g1, g2, g3);
            d_fwd_fake<false>(sD, sG, g3, row, h1, h2, h3);
            publish10(P.publish, 2, P.B, sc, act, h3);                                  // srgan.py:386
            if (dggan) {
                float s1 = sD.b4[1], lt, ds;
#pragma unroll
                for (int i = 0; i < H; ++i) s1 = fmaf(W4r1[i], h3[i], s1);
                bce_term(s1, 0.f, P.gen_mult * P.inv_Bg, lt, ds);                       // dggan.py:59-64
                l_gen_s += k * lt;
#pragma unroll
                for (int i = 0; i < H; ++i) da3[i] = k * ds * W4r1[i] * dleaky(h3[i]);
            } else {
#pragma unroll
                for (int i = 0; i < H; ++i) da3[i] = k * gf2[i] * dleaky(h3[i]);
            }
            // D data-backward only, streamed into dLoss/d(G output) = IN; G's last layer has no activation
            d_bwd_hidden(sD, h1, h2, da3, da2, da1);
            float t[H], e3[H], e2[H], e1[H];
#pragma unroll
            for (int i = 0; i < H; ++i) t[i] = 0.f;
#pragma unroll 2
            for (int j = 0; j < NIN; ++j) {
                float dj = 0.f, r[HP], q[HP];
                ldrow(sD.W1 + j * HP, r);
#pragma unroll
                for (int o = 0; o < H; ++o) dj = fmaf(r[o], da1[o], dj);
                row[O_IN + j] = dj;
                ldrow(sG.W4 + j * HP, q);
#pragma unroll
                for (int i = 0; i < H; ++i) t[i] = fmaf(q[i], dj, t[i]);
            }
#pragma unroll
            for (int i = 0; i < H; ++i) e3[i] = t[i] * dleaky(g3[i]);
            mvt(sG.W3, e3, t);
#pragma unroll
            for (int i = 0; i < H; ++i) e2[i] = t[i] * dleaky(g2[i]);
            mvt(sG.W2, e2, t);
#pragma unroll
            for (int i = 0; i < H; ++i) e1[i] = t[i] * dleaky(g1[i]);
            put10(row + O_Z, zz); put10(row + O_E1, e1); put10(row + O_G1, g1); put10(row + O_E2, e2); put10(row + O_G2, g2);
            put10(row + O_E3, e3); put10(row + O_G3, g3);
            row[O_ONE] = 1.f;
            accumulate<KG>(stage, pkG, acc, 880);
        }
#pragma unroll
        for (int k = 0; k < KG; ++k) {
            const int e = tid + k * CT;
            if (e < 880) atomicAdd(g_grad_ptr(P.G, e), acc[k]);
        }
        if (dggan) scalar_partial(l_gen_s, red, P.ws, 6);
    }
    __threadfence();
    grid.sync();
    if (dggan) combine_scalars(P.ws, P.scalars, 0x40u);
    adam_net(P.G, NZ, NIN, tG, (double)P.lr, P.beta1, P.beta2, P.eps, 0.f, gtid, gthreads, bc);   // no weight decay on G (srgan.py:137)
    if (gtid == 0) P.G.state[0] = tG;
}

}  // namespace

extern "C" size_t srgan_coefficient_step_workspace_bytes(void) { return (size_t)WS_FLOATS * sizeof(float); }

extern "C" int srgan_coefficient_step(const float* const* d_ptrs, const float* const* g_ptrs, const float* const* dnn_ptrs,
                                      float* d_state, float* g_state, float* dnn_state, const float* x, const float* y,
                                      const float* u, const float* z, const float* alpha, const float* z2, int B,
                                      float inv_Bg, int dggan, int order, float labeled_mult, float unl_mult,
                                      float fake_mult, float gen_mult, float gp_lambda, int kind_match, int kind_contrast,
                                      float lr, float lr_dnn, float wd, float beta1, float beta2, float eps, int phases,
                                      int train_g, void* workspace, size_t workspace_bytes, float* scalars, float* publish,
                                      void* stream) {
    SRGAN_REQUIRE(d_ptrs && g_ptrs && dnn_ptrs && d_state && g_state && dnn_state && x && y && scalars && B > 0,
                  "srgan_coefficient_step: bad arguments");
    SRGAN_REQUIRE(phases >= 1 && phases <= 3, "srgan_coefficient_step: phases must be 1 (dnn), 2 (gan) or 3 (both)");
    SRGAN_REQUIRE(!(phases & 2) || (u && z && alpha && z2), "srgan_coefficient_step: the GAN step needs u, z, alpha, z2");
    SRGAN_REQUIRE(kind_match >= 0 && kind_match <= 5 && kind_contrast >= 0 && kind_contrast <= 5,
                  "srgan_coefficient_step: unknown distance kind");
    SRGAN_REQUIRE(workspace && workspace_bytes >= (size_t)WS_FLOATS * sizeof(float),
                  "srgan_coefficient_step: workspace too small (srgan_coefficient_step_workspace_bytes)");
    CoefParams P;
    auto fill = [](NetP& n, const float* const* p, float* state) {
        // p: 4 layers x {W, b, gW, gb, mW, mb, vW, vb}
        for (int l = 0; l < 4; ++l) {
            float** f = reinterpret_cast<float**>(&n.l[l]);
            for (int k = 0; k < 8; ++k) f[k] = const_cast<float*>(p[l * 8 + k]);
        }
        n.state = state;
    };
    fill(P.D, d_ptrs, d_state); fill(P.G, g_ptrs, g_state); fill(P.DNN, dnn_ptrs, dnn_state);
    for (int l = 0; l < 4; ++l)
        for (const NetP* n : {&P.D, &P.G, &P.DNN}) {
            const float* const* f = reinterpret_cast<const float* const*>(&n->l[l]);
            for (int k = 0; k < 8; ++k) SRGAN_REQUIRE(f[k], "srgan_coefficient_step: null parameter pointer");
        }
    P.x = x; P.y = y; P.u = u; P.z = z; P.alpha = alpha; P.z2 = z2;
    P.B = B; P.inv_Bg = inv_Bg; P.dggan = dggan; P.head_out = dggan ? 2 : 1; P.order = order;
    P.labeled_mult = labeled_mult; P.unl_mult = unl_mult; P.fake_mult = fake_mult; P.gen_mult = gen_mult; P.gp_lambda = gp_lambda;
    P.kind_match = kind_match; P.kind_contrast = kind_contrast;
    P.lr = lr; P.lr_dnn = lr_dnn; P.wd = wd; P.beta1 = beta1; P.beta2 = beta2; P.eps = eps;
    P.phases = phases; P.train_g = train_g;
    P.ws = static_cast<float*>(workspace); P.scalars = scalars; P.publish = publish;
    const size_t smem = (size_t)(2 * net_floats(NIN, 2, true) + net_floats(NZ, NIN, false) + NSLOT * CT) * sizeof(float);
    int grid = (B + CT - 1) / CT;
    if (grid > MAXG) grid = MAXG;
    void* args[] = {(void*)&P};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)coef_step_kernel, dim3(grid), dim3(CT), args, smem, (cudaStream_t)stream);
    if (e != cudaSuccess) { srgan_set_error("coef_step_kernel: cooperative launch failed: %s", cudaGetErrorString(e)); return SRGAN_ERR_CUDA; }
    SRGAN_COUNT_LAUNCH();
    return SRGAN_OK;
}
