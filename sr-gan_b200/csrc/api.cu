// C ABI: library plumbing + dispatch of the dense contractions between the tcgen05 path and the SIMT path.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

std::atomic<long long> g_launches{0};
static thread_local char t_err[512] = "";
static thread_local int t_last_tensor = 0;
static std::atomic<int> g_force_simt{0};
static std::atomic<long long> g_tensor_calls{0}, g_simt_fallbacks{0};

// A bf16 contraction of tensor-core size that is not eligible for the tcgen05 kernels runs ~30x slower on the fp32-FMA
// kernels: count it and say so once per distinct shape (up to a few lines), never silently.
static void note_simt_fallback(const char* who, const srgan_geom* g, int n) {
    if (g->Ca < 64 || g->Cb < 64) return;            // narrow layers (MLPs, map heads, image stems) are SIMT by design
    g_simt_fallbacks.fetch_add(1, std::memory_order_relaxed);
    static std::atomic<int> lines{0};
    static std::atomic<unsigned long long> seen[8];
    const unsigned long long key = ((unsigned long long)g->Ca << 44) ^ ((unsigned long long)g->Cb << 28) ^
                                   ((unsigned long long)g->Hs << 16) ^ ((unsigned long long)g->R << 8) ^ (unsigned)g->stride ^
                                   ((unsigned long long)(who[11] == 'w') << 60);
    for (auto& s : seen)
        if (s.load() == key) return;
    const int i = lines.fetch_add(1);
    if (i >= 8) return;
    seen[i].store(key);
    const char* q = getenv("SRGAN_QUIET_FALLBACK");
    if (q && q[0] == '1') return;
    fprintf(stderr, "srgan_b200: %s n=%d %dx%dx%d <- %dx%dx%d k%d s%d p%d (bf16) is not tcgen05-eligible: running on the fp32-FMA "
            "kernel\n", who, n, g->Hs, g->Ws, g->Ca, g->Hl, g->Wl, g->Cb, g->R, g->stride, g->pad);
}

void srgan_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

// simt_conv.cu
int simt_conv(int mode, const void* src, const void* W, void* out, int n, const srgan_geom* g, const float* bias,
              int bias_mod, const void* href, int epi, int act, float slope, int dtype, cudaStream_t st);
int simt_wgrad(const void* S, const void* L, float* dW, int n, const srgan_geom* g, int dtype, cudaStream_t st);
// umma_conv.cu : return 1 if the tensor-core path took the call, 0 if the shape is not eligible, <0 on error
int umma_conv(int mode, const void* src, const void* W, void* out, int n, const srgan_geom* g, const float* bias,
              int bias_mod, const void* href, int epi, int act, float slope, const srgan_views* views, cudaStream_t st);
int umma_wgrad(const void* S, const void* L, float* dW, int n, const srgan_geom* g, const srgan_views* views, cudaStream_t st);

// bn_gemm.cu : dense-layer GEMMs with the preceding eval-mode BatchNorm + ReLU fused in (1 = launched, 0 = not eligible)
int bn_dgrad(const void* dy, const void* Wu, void* dx, const void* x, long long rows, int K, int Cout, int C, int pitch,
             const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* dgamma, float* dbeta,
             void* d_out, int d_pitch, int accumulate, cudaStream_t st);

int bn_conv_dgrad(const void* dy, int dy_pitch, int dy_valid, const void* Wu, void* dx, const void* x, int n, int H, int W, int R,
                  int S, int pad, int Cin, int Cout, int C, int pitch, const float* gamma, const float* beta, const float* mean,
                  const float* var, float eps, float* dgamma, float* dbeta, void* d_out, int d_pitch, int accumulate, cudaStream_t st);
int bn_conv_down(const void* x, const void* Wd, void* out, long long rows, int Kpad, int Cout, int C, int pitch, const float* gamma,
                 const float* beta, const float* mean, const float* var, float eps, void* n1_out, int n1_pitch, long long n1_first_row,
                 const float* gamma2, const float* beta2, const float* mean2, const float* var2, void* out2, int C2, cudaStream_t st);
int bn_conv_wgrad(const void* dy, const void* x, float* dW, long long rows, int Ca, int Kpad, int C, int pitch, const float* gamma,
                  const float* beta, const float* mean, const float* var, float eps, cudaStream_t st);

// skinny.cu : few outputs over a long full-extent reduction axis (MapModule.linear1, the count feature layer)
bool skinny_eligible(const srgan_geom* g);
int skinny_conv(int mode, const void* src, const void* W, void* out, int n, const srgan_geom* g, const float* bias, int bias_mod,
                const void* href, int epi, int act, float slope, int dtype, cudaStream_t st);
int skinny_wgrad(const void* S, const void* L, float* dW, int n, const srgan_geom* g, int dtype, cudaStream_t st);

static int check_geom(const char* who, const srgan_geom* g, int n) {
    if (!g || n < 0 || g->Hs <= 0 || g->Ws <= 0 || g->Ca <= 0 || g->Hl <= 0 || g->Wl <= 0 || g->Cb <= 0 || g->R <= 0 ||
        g->S <= 0 || g->stride <= 0 || g->pad < 0) {
        srgan_set_error("%s: bad geometry", who);
        return SRGAN_ERR_ARG;
    }
    // every small-side position must read inside the padded large side: the pair is a valid strided conv
    if ((g->Hs - 1) * g->stride - g->pad + g->R - 1 >= g->Hl + g->pad || (g->Ws - 1) * g->stride - g->pad + g->S - 1 >= g->Wl + g->pad) {
        srgan_set_error("%s: small side %dx%d does not fit large side %dx%d (k%d s%d p%d)", who, g->Hs, g->Ws, g->Hl,
                        g->Wl, g->R, g->stride, g->pad);
        return SRGAN_ERR_ARG;
    }
    return SRGAN_OK;
}

extern "C" {

int srgan_version(void) { return 100; }
const char* srgan_last_error(void) { return t_err; }
long long srgan_launch_count(void) { return g_launches.load(); }
int srgan_last_path_tensor(void) { return t_last_tensor; }
long long srgan_tensor_launch_count(void) { return g_tensor_calls.load(); }
long long srgan_simt_fallback_count(void) { return g_simt_fallbacks.load(); }
void srgan_set_force_simt(int on) { g_force_simt.store(on); }

static bool has_views(const srgan_views* v) { return v && (v->S_pitch || v->S_valid || v->L_pitch || v->L_valid); }

static int conv_common(int mode, const char* who, const void* src, const void* W, void* out, int n, const srgan_geom* g,
                       const float* bias, int bias_mod, const void* href, int epi, int act, float slope, int dtype,
                       const srgan_views* views, void* stream) {
    if (!has_views(views)) views = nullptr;
    int rc = check_geom(who, g, n);
    if (rc) return rc;
    SRGAN_REQUIRE(src && W && out, "%s: null pointer", who);
    SRGAN_REQUIRE(dtype == SRGAN_F32 || dtype == SRGAN_BF16, "%s: unknown dtype %d", who, dtype);
    SRGAN_REQUIRE(epi == SRGAN_EPI_BIAS_ACT || epi == SRGAN_EPI_DACT, "%s: unknown epilogue %d", who, epi);
    SRGAN_REQUIRE(act >= SRGAN_ACT_NONE && act <= SRGAN_ACT_TANH, "%s: unknown activation %d", who, act);
    int ncols = mode == 0 ? g->Ca : g->Cb;
    SRGAN_REQUIRE(bias_mod == 0 || ncols % bias_mod == 0, "%s: bias_mod does not divide the channel count", who);
    if (n == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    t_last_tensor = 0;
    if (views == nullptr && skinny_eligible(g) && !g_force_simt.load())
        return skinny_conv(mode, src, W, out, n, g, bias, bias_mod, href, epi, act, slope, dtype, st);
    if (dtype == SRGAN_BF16 && !g_force_simt.load()) {
        int took = umma_conv(mode, src, W, out, n, g, bias, bias_mod, href, epi, act, slope, views, st);
        if (took < 0) return took;
        if (took == 1) { t_last_tensor = 1; g_tensor_calls.fetch_add(1, std::memory_order_relaxed); return SRGAN_OK; }
        note_simt_fallback(who, g, n);
    }
    // channel windows are a feature of the TMA-fed kernels (tensor-map strides / epilogue pitch): no second implementation
    SRGAN_REQUIRE(views == nullptr, "%s: channel windows (srgan_views) need a tcgen05-eligible bf16 shape: both channel counts "
                  "multiples of 64, pitches / windows multiples of 8", who);
    return simt_conv(mode, src, W, out, n, g, bias, bias_mod, href, epi, act, slope, dtype, st);
}

int srgan_conv_down(const void* L, const void* Wd, void* S_out, int n, const srgan_geom* g, const float* bias,
                    int bias_mod, const void* href, int epi, int act, float slope, int dtype, const srgan_views* views,
                    void* stream) {
    return conv_common(0, "srgan_conv_down", L, Wd, S_out, n, g, bias, bias_mod, href, epi, act, slope, dtype, views, stream);
}

int srgan_conv_up(const void* S, const void* Wu, void* L_out, int n, const srgan_geom* g, const float* bias,
                  int bias_mod, const void* href, int epi, int act, float slope, int dtype, const srgan_views* views,
                  void* stream) {
    return conv_common(1, "srgan_conv_up", S, Wu, L_out, n, g, bias, bias_mod, href, epi, act, slope, dtype, views, stream);
}

int srgan_conv_wgrad(const void* S, const void* L, float* dW, int n, const srgan_geom* g, int dtype, const srgan_views* views,
                     void* stream) {
    if (!has_views(views)) views = nullptr;
    int rc = check_geom("srgan_conv_wgrad", g, n);
    if (rc) return rc;
    SRGAN_REQUIRE(S && L && dW, "srgan_conv_wgrad: null pointer");
    SRGAN_REQUIRE(dtype == SRGAN_F32 || dtype == SRGAN_BF16, "srgan_conv_wgrad: unknown dtype %d", dtype);
    if (n == 0) return SRGAN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    t_last_tensor = 0;
    if (views == nullptr && skinny_eligible(g) && !g_force_simt.load()) return skinny_wgrad(S, L, dW, n, g, dtype, st);
    if (dtype == SRGAN_BF16 && !g_force_simt.load()) {
        int took = umma_wgrad(S, L, dW, n, g, views, st);
        if (took < 0) return took;
        if (took == 1) { t_last_tensor = 1; g_tensor_calls.fetch_add(1, std::memory_order_relaxed); return SRGAN_OK; }
        note_simt_fallback("srgan_conv_wgrad", g, n);
    }
    SRGAN_REQUIRE(views == nullptr, "srgan_conv_wgrad: channel windows (srgan_views) need a tcgen05-eligible bf16 shape");
    return simt_wgrad(S, L, dW, n, g, dtype, st);
}

int srgan_bn_dgrad(const void* dy, const void* Wu, void* dx, const void* x, long long rows, int K, int Cout, int C, int pitch,
                   const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* dgamma,
                   float* dbeta, void* d_out, int d_pitch, int accumulate, int dtype, void* stream) {
    SRGAN_REQUIRE(dy && Wu && dx && x && gamma && beta && mean && var, "srgan_bn_dgrad: null pointer");
    SRGAN_REQUIRE(dtype == SRGAN_BF16, "srgan_bn_dgrad: the fused dense-layer kernels are bf16 / tcgen05 only");
    SRGAN_REQUIRE(rows >= 0 && K > 0 && C > 0 && Cout >= C && pitch >= C, "srgan_bn_dgrad: bad sizes");
    if (rows == 0) return SRGAN_OK;
    int took = bn_dgrad(dy, Wu, dx, x, rows, K, Cout, C, pitch, gamma, beta, mean, var, eps, dgamma, dbeta, d_out, d_pitch,
                        accumulate, (cudaStream_t)stream);
    if (took < 0) return took;
    if (took == 0) {
        srgan_set_error("srgan_bn_dgrad: shape not eligible (K %% 64, Cout %% 64, C %% 8, pitch %% 8, 16-byte aligned pointers)");
        return SRGAN_ERR_UNSUPPORTED;
    }
    t_last_tensor = 1;
    g_tensor_calls.fetch_add(1, std::memory_order_relaxed);
    return SRGAN_OK;
}

int srgan_bn_conv_dgrad(const void* dy, int dy_pitch, int dy_valid, const void* Wu, void* dx, const void* x, int n, int H, int W,
                        int R, int S, int pad, int Cin, int Cout, int C, int pitch, const float* gamma, const float* beta,
                        const float* mean, const float* var, float eps, float* dgamma, float* dbeta, void* d_out, int d_pitch,
                        int accumulate, int dtype, void* stream) {
    SRGAN_REQUIRE(dy && Wu && dx && x && gamma && beta && mean && var, "srgan_bn_conv_dgrad: null pointer");
    SRGAN_REQUIRE(dtype == SRGAN_BF16, "srgan_bn_conv_dgrad: the fused dense-layer kernels are bf16 / tcgen05 only");
    SRGAN_REQUIRE(n >= 0 && H > 0 && W > 0 && Cin > 0 && C > 0 && Cout >= C && pitch >= C, "srgan_bn_conv_dgrad: bad sizes");
    if (n == 0) return SRGAN_OK;
    int took = bn_conv_dgrad(dy, dy_pitch, dy_valid, Wu, dx, x, n, H, W, R, S, pad, Cin, Cout, C, pitch, gamma, beta, mean, var, eps,
                             dgamma, dbeta, d_out, d_pitch, accumulate, (cudaStream_t)stream);
    if (took < 0) return took;
    if (took == 0) {
        srgan_set_error("srgan_bn_conv_dgrad: shape not eligible (odd R = S = 2 pad + 1, Cin %% 64, Cout %% 64, C %% 8, pitches %% 8, "
                        "power-of-two pixel patches, 16-byte aligned pointers)");
        return SRGAN_ERR_UNSUPPORTED;
    }
    t_last_tensor = 1;
    g_tensor_calls.fetch_add(1, std::memory_order_relaxed);
    return SRGAN_OK;
}

int srgan_bn_conv_down(const void* x, const void* Wd, void* out, long long rows, int Kpad, int Cout, int C, int pitch,
                       const float* gamma, const float* beta, const float* mean, const float* var, float eps, void* n1_out,
                       int n1_pitch, long long n1_first_row, const float* gamma2, const float* beta2, const float* mean2,
                       const float* var2, void* out2, int C2, int dtype, void* stream) {
    SRGAN_REQUIRE(x && Wd && out && gamma && beta && mean && var, "srgan_bn_conv_down: null pointer");
    SRGAN_REQUIRE(dtype == SRGAN_BF16, "srgan_bn_conv_down: the fused dense-layer kernels are bf16 / tcgen05 only");
    SRGAN_REQUIRE(rows >= 0 && Kpad >= C && C > 0 && Cout > 0 && pitch >= C, "srgan_bn_conv_down: bad sizes");
    SRGAN_REQUIRE((out2 == nullptr) == (gamma2 == nullptr) && (gamma2 == nullptr) == (beta2 == nullptr) &&
                      (gamma2 == nullptr) == (mean2 == nullptr) && (gamma2 == nullptr) == (var2 == nullptr),
                  "srgan_bn_conv_down: the second BatchNorm needs gamma2, beta2, mean2, var2 and out2 together");
    SRGAN_REQUIRE(out2 == nullptr || (C2 > 0 && C2 <= Cout), "srgan_bn_conv_down: C2 must be in (0, Cout]");
    if (rows == 0) return SRGAN_OK;
    int took = bn_conv_down(x, Wd, out, rows, Kpad, Cout, C, pitch, gamma, beta, mean, var, eps, n1_out, n1_pitch, n1_first_row,
                            gamma2, beta2, mean2, var2, out2, C2, (cudaStream_t)stream);
    if (took < 0) return took;
    if (took == 0) {
        srgan_set_error("srgan_bn_conv_down: shape not eligible (Kpad %% 64, C %% 8, Cout %% 8, pitch %% 8, 16-byte aligned pointers)");
        return SRGAN_ERR_UNSUPPORTED;
    }
    t_last_tensor = 1;
    g_tensor_calls.fetch_add(1, std::memory_order_relaxed);
    return SRGAN_OK;
}

int srgan_bn_conv_wgrad(const void* dy, const void* x, float* dW, long long rows, int Ca, int Kpad, int C, int pitch,
                        const float* gamma, const float* beta, const float* mean, const float* var, float eps, int dtype,
                        void* stream) {
    SRGAN_REQUIRE(dy && x && dW && gamma && beta && mean && var, "srgan_bn_conv_wgrad: null pointer");
    SRGAN_REQUIRE(dtype == SRGAN_BF16, "srgan_bn_conv_wgrad: the fused dense-layer kernels are bf16 / tcgen05 only");
    SRGAN_REQUIRE(rows >= 0 && Kpad >= C && C > 0 && Ca > 0 && pitch >= C, "srgan_bn_conv_wgrad: bad sizes");
    if (rows == 0) return SRGAN_OK;
    int took = bn_conv_wgrad(dy, x, dW, rows, Ca, Kpad, C, pitch, gamma, beta, mean, var, eps, (cudaStream_t)stream);
    if (took < 0) return took;
    if (took == 0) {
        srgan_set_error("srgan_bn_conv_wgrad: shape not eligible (Kpad %% 64, C %% 8, Ca %% 8, pitch %% 8, 16-byte aligned pointers)");
        return SRGAN_ERR_UNSUPPORTED;
    }
    t_last_tensor = 1;
    g_tensor_calls.fetch_add(1, std::memory_order_relaxed);
    return SRGAN_OK;
}

}  // extern "C"
