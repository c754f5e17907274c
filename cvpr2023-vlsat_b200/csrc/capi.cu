// C-ABI glue: version / error strings and the engine dispatch for the dense projections and the
// streaming attention. See include/vlsat_b200.h for the contract.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

namespace vlsat {
long long g_launch_count = 0;
// Arithmetic of the tensor-core kernels: 0 = BF16x3 (three bf16 MMAs per product: fp32 parity, the default),
// 1 = single-pass bf16 (one MMA per product on the hi halves: BASELINE configs #3 / #4, looser stated tolerance).
int g_precision = [] { const char* e = getenv("VLSAT_PRECISION"); return (e && (e[0] == 'b' || e[0] == '1')) ? 1 : 0; }();
int tc_passes() { return g_precision == 1 ? 1 : 3; }
bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("VLSAT_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
    return on != 0;
}
int linear_simt(const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int64_t, int64_t,
                const vlsat_epilogue*, cudaStream_t);
bool linear_tc_eligible(const float*, int64_t, const float*, int64_t, int64_t, int64_t, int64_t, int);
size_t linear_tc_workspace_bytes(int64_t, int64_t, int64_t, bool, bool);
int tf32_split(const float*, int64_t, int64_t, int64_t, float*, float*, cudaStream_t);
int linear_tc(const void*, const void*, const void*, const void*, float*, int64_t, int64_t, int64_t, int64_t,
              const vlsat_epilogue*, int, int, cudaStream_t);
bool pointnet_tc_eligible(int, int, int, int, int64_t);
int pointnet_tc(const float*, int64_t, int, int64_t, const float*, const float*, const float*, const float*, const float*,
                const float*, int, float*, int32_t*, cudaStream_t);
int flash_attn_tc(const float*, const float*, int64_t, const float*, const float*, int64_t, const float*, const float*,
                  int64_t, float*, int64_t, float*, int64_t, int64_t, int, int, cudaStream_t);
int bf16_split(const float*, int64_t, int64_t, int64_t, uint16_t*, uint16_t*, int64_t, cudaStream_t);
int flash_attn_bf16(const uint16_t*, const uint16_t*, int64_t, const uint16_t*, const uint16_t*, int64_t, const uint16_t*,
                    const uint16_t*, int64_t, float*, int64_t, float*, int64_t, int64_t, int, int, void*, size_t, cudaStream_t);
size_t flash_attn_bf16_workspace_bytes(int64_t, int64_t, int, int*);
int flash_attn_simt(const float*, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, float*,
                    int64_t, int64_t, int, int, cudaStream_t);
}  // namespace vlsat

namespace vlsat { extern long long* g_trace; }
using namespace vlsat;

// Debug only (not part of the public header): device buffer of >= 128 int64 that CTA (0,0) of the next
// tensor-core GEMM launches fills with clock64 stamps; pass NULL to switch tracing off.
extern "C" void vlsat_debug_set_trace(long long* device_buffer) { g_trace = device_buffer; }

extern "C" int vlsat_version(void) { return 100; }

extern "C" const char* vlsat_error_string(int status) {
    switch (status) {
        case VLSAT_OK: return "ok";
        case VLSAT_ERR_INVALID_ARG: return "invalid argument (null pointer, negative size or inconsistent dims)";
        case VLSAT_ERR_UNSUPPORTED: return "shape not supported by the sm_100a kernels (no fallback exists)";
        case VLSAT_ERR_LAUNCH: return "CUDA launch failed";
        case VLSAT_ERR_WORKSPACE: return "workspace too small";
        default: return "unknown status";
    }
}

extern "C" const char* vlsat_gemm_engine(void) {
    return g_precision == 1 ? "tcgen05-bf16 single pass (one kind::f16 MMA per product on the hi halves of the operand pairs)"
                            : "tcgen05-bf16x3 (tcgen05-3xtf32 / ffma-fp32 for shapes bf16 pairs cannot address)";
}

extern "C" int64_t vlsat_launch_count(void) { return g_launch_count; }

extern "C" int vlsat_set_precision(int mode) {
    if (mode != VLSAT_PRECISION_FP32 && mode != VLSAT_PRECISION_BF16) return VLSAT_ERR_INVALID_ARG;
    g_precision = mode;
    return VLSAT_OK;
}
extern "C" int vlsat_get_precision(void) { return g_precision; }

extern "C" size_t vlsat_linear_workspace_bytes(int64_t M, int64_t N, int64_t K, int need_x_split, int need_w_split) {
    return linear_tc_workspace_bytes(M, N, K, need_x_split != 0, need_w_split != 0);
}

extern "C" int vlsat_tf32_split(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* hi, float* lo, void* stream) {
    VLSAT_REQUIRE(rows >= 0 && cols >= 0);
    if (rows == 0 || cols == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && hi && lo && ldx >= cols);
    VLSAT_SUPPORT(cols % 4 == 0 && ldx % 4 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)hi % 16 == 0) && ((uintptr_t)lo % 16 == 0));
    return tf32_split(x, ldx, rows, cols, hi, lo, (cudaStream_t)stream);
}

extern "C" int vlsat_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, float* y, int64_t ldy,
                                int64_t M, int64_t N, int64_t K, const vlsat_epilogue* epi,
                                const vlsat_linear_opts* opts, void* stream) {
    VLSAT_REQUIRE(M >= 0 && N >= 1 && K >= 1);
    if (M == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && w && ldx >= K && ldw >= K);
    VLSAT_REQUIRE((y && ldy >= N) || (epi && epi->split_hi));
    if (epi) {
        VLSAT_REQUIRE((epi->split_hi == nullptr) == (epi->split_lo == nullptr));
        VLSAT_REQUIRE(!epi->split_hi || epi->ld_split >= N);
        VLSAT_REQUIRE(!epi->gather_a || epi->idx_a);
        VLSAT_REQUIRE(!epi->gather_b || epi->idx_b);
        VLSAT_REQUIRE(!(epi->gather_a || epi->gather_b) || epi->ld_gather >= N);
        VLSAT_REQUIRE(!epi->residual || epi->ld_res >= N);
    }
    if (epi && epi->split_hi) VLSAT_REQUIRE(epi->split_fmt == VLSAT_SPLIT_TF32 || epi->split_fmt == VLSAT_SPLIT_BF16);
    cudaStream_t st = (cudaStream_t)stream;
    int engine = opts ? opts->engine : VLSAT_ENGINE_AUTO;
    VLSAT_REQUIRE(engine >= VLSAT_ENGINE_AUTO && engine <= VLSAT_ENGINE_TC_BF16X3);
    if (engine == VLSAT_ENGINE_AUTO) {
        if (!opts) engine = VLSAT_ENGINE_SIMT;
        else if (linear_tc_eligible(x, ldx, w, ldw, M, N, K, 1)) engine = VLSAT_ENGINE_TC_BF16X3;
        else if (linear_tc_eligible(x, ldx, w, ldw, M, N, K, 0)) engine = VLSAT_ENGINE_TC;
        else engine = VLSAT_ENGINE_SIMT;
    }
    if (engine == VLSAT_ENGINE_SIMT) return linear_simt(x, ldx, w, ldw, y, ldy, M, N, K, epi, st);
    const int kind = engine == VLSAT_ENGINE_TC_BF16X3 ? 1 : 0;
    VLSAT_SUPPORT(linear_tc_eligible(x, ldx, w, ldw, M, N, K, kind));
    const int passes = engine == VLSAT_ENGINE_TC_1PASS ? 1 : 3;
    const void *xh = opts->x_hi, *xl = opts->x_lo, *wh = opts->w_hi, *wl = opts->w_lo;
    VLSAT_REQUIRE((xh == nullptr) == (xl == nullptr) && (wh == nullptr) == (wl == nullptr));
    const size_t need = linear_tc_workspace_bytes(M, N, K, xh == nullptr, wh == nullptr);
    if (need > 0 && (!opts->workspace || opts->workspace_bytes < need)) return VLSAT_ERR_WORKSPACE;
    float* ws = (float*)opts->workspace;
    if (need > 0) VLSAT_SUPPORT((uintptr_t)ws % 16 == 0);
    auto split = [&](const float* src, int64_t ld, int64_t rows, const void*& hi, const void*& lo) -> int {
        // both formats: hi then lo, rows * K elements each; a pair takes rows * K floats
        if (kind == 1) {
            uint16_t* h = (uint16_t*)ws; uint16_t* l = h + rows * K;
            hi = h; lo = l; ws += rows * K;
            return bf16_split(src, ld, rows, K, h, l, K, st);
        }
        float* h = ws; float* l = ws + rows * K;
        hi = h; lo = l; ws += 2 * rows * K;
        return tf32_split(src, ld, rows, K, h, l, st);
    };
    if (!xh) { int rc = split(x, ldx, M, xh, xl); if (rc) return rc; }
    if (!wh) { int rc = split(w, ldw, N, wh, wl); if (rc) return rc; }
    return linear_tc(xh, xl, wh, wl, y, ldy, M, N, K, epi, passes, kind, st);
}

extern "C" int vlsat_flash_attn_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                    float* out, int64_t ldo, float* lse, int64_t nq, int64_t nk, int n_heads, int dk,
                                    void* stream) {
    VLSAT_REQUIRE(nq >= 0 && nk >= 1 && n_heads >= 1);
    if (nq == 0) return VLSAT_OK;
    VLSAT_REQUIRE(q && k && v && out);
    VLSAT_SUPPORT(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0);
    VLSAT_SUPPORT(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) && ((uintptr_t)out % 16 == 0));
    return flash_attn_simt(q, ldq, k, ldk, v, ldv, out, ldo, lse, nq, nk, n_heads, dk, (cudaStream_t)stream);
}

extern "C" int vlsat_flash_attn_tc_fwd(const float* q_hi, const float* q_lo, int64_t ldq, const float* k_hi,
                                       const float* k_lo, int64_t ldk, const float* vt_hi, const float* vt_lo,
                                       int64_t ldvt, float* out, int64_t ldo, float* lse, int64_t nq, int64_t nk,
                                       int n_heads, int dk, void* stream) {
    VLSAT_REQUIRE(nq >= 0 && nk >= 1 && n_heads >= 1);
    if (nq == 0) return VLSAT_OK;
    VLSAT_REQUIRE(q_hi && q_lo && k_hi && k_lo && vt_hi && vt_lo && out);
    VLSAT_REQUIRE(ldq >= (int64_t)n_heads * dk && ldk >= (int64_t)n_heads * dk && ldvt >= nk && ldo >= (int64_t)n_heads * dk);
    const uintptr_t all = (uintptr_t)q_hi | (uintptr_t)q_lo | (uintptr_t)k_hi | (uintptr_t)k_lo | (uintptr_t)vt_hi |
                          (uintptr_t)vt_lo | (uintptr_t)out;
    VLSAT_SUPPORT(all % 16 == 0);
    return flash_attn_tc(q_hi, q_lo, ldq, k_hi, k_lo, ldk, vt_hi, vt_lo, ldvt, out, ldo, lse, nq, nk, n_heads, dk,
                         (cudaStream_t)stream);
}

extern "C" int vlsat_pointnet_tc_fwd(const float* x, int64_t n_obj, int c_in, int64_t n_pts,
                                     const float* w1, const float* b1, int c1, const float* w2, const float* b2, int c2,
                                     const float* w3, const float* b3, int c_out, float* out, int32_t* argmax, void* stream) {
    VLSAT_REQUIRE(x && w1 && b1 && w2 && b2 && w3 && b3 && out);
    VLSAT_REQUIRE(n_obj >= 0 && n_pts >= 1 && c_in >= 1 && c_out >= 1);
    VLSAT_SUPPORT(pointnet_tc_eligible(c_in, c1, c2, c_out, n_pts) && n_pts < (1ll << 31));
    if (n_obj == 0) return VLSAT_OK;
    return pointnet_tc(x, n_obj, c_in, n_pts, w1, b1, w2, b2, w3, b3, c_out, out, argmax, (cudaStream_t)stream);
}

extern "C" int vlsat_bf16_split(const float* x, int64_t ldx, int64_t rows, int64_t cols, void* hi, void* lo, int64_t ld_out,
                                void* stream) {
    VLSAT_REQUIRE(rows >= 0 && cols >= 0);
    if (rows == 0 || cols == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && hi && lo && ldx >= cols && ld_out >= cols);
    VLSAT_SUPPORT(ld_out % 8 == 0 && ((uintptr_t)hi % 16 == 0) && ((uintptr_t)lo % 16 == 0));
    return bf16_split(x, ldx, rows, cols, (uint16_t*)hi, (uint16_t*)lo, ld_out, (cudaStream_t)stream);
}

extern "C" int vlsat_flash_attn_bf16x3_fwd(const void* q_hi, const void* q_lo, int64_t ldq, const void* k_hi, const void* k_lo,
                                           int64_t ldk, const void* vt_hi, const void* vt_lo, int64_t ldvt, float* out,
                                           int64_t ldo, float* lse, int64_t nq, int64_t nk, int n_heads, int dk,
                                           void* workspace, size_t workspace_bytes, void* stream) {
    VLSAT_REQUIRE(nq >= 0 && nk >= 1 && n_heads >= 1);
    if (nq == 0) return VLSAT_OK;
    VLSAT_REQUIRE(q_hi && q_lo && k_hi && k_lo && vt_hi && vt_lo && out);
    VLSAT_REQUIRE(ldq >= (int64_t)n_heads * dk && ldk >= (int64_t)n_heads * dk && ldvt >= nk && ldo >= (int64_t)n_heads * dk);
    const uintptr_t all = (uintptr_t)q_hi | (uintptr_t)q_lo | (uintptr_t)k_hi | (uintptr_t)k_lo | (uintptr_t)vt_hi |
                          (uintptr_t)vt_lo | (uintptr_t)out | (uintptr_t)workspace;
    VLSAT_SUPPORT(all % 16 == 0);
    return flash_attn_bf16((const uint16_t*)q_hi, (const uint16_t*)q_lo, ldq, (const uint16_t*)k_hi, (const uint16_t*)k_lo, ldk,
                           (const uint16_t*)vt_hi, (const uint16_t*)vt_lo, ldvt, out, ldo, lse, nq, nk, n_heads, dk,
                           workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" size_t vlsat_flash_attn_bf16x3_workspace_bytes(int64_t nq, int64_t nk, int n_heads) {
    return flash_attn_bf16_workspace_bytes(nq, nk, n_heads, nullptr);
}
