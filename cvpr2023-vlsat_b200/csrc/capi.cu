// C-ABI glue: version / error strings and the engine dispatch for the dense projections and the
// streaming attention. See include/vlsat_b200.h for the contract.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

namespace vlsat {
long long g_launch_count = 0;
int linear_simt(const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int64_t, int64_t,
                const vlsat_epilogue*, cudaStream_t);
int flash_attn_simt(const float*, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, float*,
                    int64_t, int64_t, int, int, cudaStream_t);
}  // namespace vlsat

using namespace vlsat;

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

extern "C" const char* vlsat_gemm_engine(void) { return "simt-fp32"; }

extern "C" int64_t vlsat_launch_count(void) { return g_launch_count; }

extern "C" int vlsat_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, float* y, int64_t ldy,
                                int64_t M, int64_t N, int64_t K, const vlsat_epilogue* epi, void* stream) {
    VLSAT_REQUIRE(M >= 0 && N >= 1 && K >= 1);
    if (M == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && w && y && ldx >= K && ldw >= K && ldy >= N);
    if (epi) {
        VLSAT_REQUIRE(!epi->gather_a || epi->idx_a);
        VLSAT_REQUIRE(!epi->gather_b || epi->idx_b);
        VLSAT_REQUIRE(!(epi->gather_a || epi->gather_b) || epi->ld_gather >= N);
        VLSAT_REQUIRE(!epi->residual || epi->ld_res >= N);
    }
    return linear_simt(x, ldx, w, ldw, y, ldy, M, N, K, epi, (cudaStream_t)stream);
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
