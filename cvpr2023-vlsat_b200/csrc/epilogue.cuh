// Arguments and fused epilogue shared by the two dense-projection engines (FFMA and tcgen05).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace vlsat {

struct LinearArgs {
    const float* x; int64_t ldx;
    const float* w; int64_t ldw;
    float* y; int64_t ldy;
    int64_t M, N, K;
    vlsat_epilogue epi;
    long long* trace;      // debug: per-phase clock64 stamps of CTA (0,0); nullptr in normal use
    int tma_store;         // tensor-core engine: outputs leave through TMA stores (all rows 16-byte addressable)
};

__device__ __forceinline__ float apply_act(float t, int act) {
    if (act == VLSAT_ACT_RELU) return fmaxf(t, 0.f);
    if (act == VLSAT_ACT_SIGMOID) {
        // opaque to if-conversion: keep the exp / divide off the ReLU and identity paths
        float r;
        asm volatile("" ::: "memory");
        r = 1.f / (1.f + expf(-t));
        return r;
    }
    return t;
}

// Shared epilogue for one output element group; used by both GEMM engines.
__device__ __forceinline__ float epilogue_one(const vlsat_epilogue& e, float acc, int64_t m, int64_t n,
                                              int64_t ia, int64_t ib, float post_scale) {
    float t = acc;
    if (e.bias) t += __ldg(e.bias + (e.bias_per_row ? m : n));
    if (e.gather_a) t += __ldg(e.gather_a + ia * e.ld_gather + n);
    if (e.gather_b) t += __ldg(e.gather_b + ib * e.ld_gather + n);
    t = apply_act(t, e.act);
    if (e.residual) t = e.alpha * t + e.beta * __ldg(e.residual + m * e.ld_res + n);
    else if (e.alpha != 1.f) t *= e.alpha;
    return t * post_scale;
}

// tf32 split of one value: hi = round-to-nearest tf32, lo = v - hi
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    uint32_t u;
    u = tc::tf32_rna_bits(v);
    hi = __uint_as_float(u);
    lo = v - hi;
}

}  // namespace vlsat
