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
    // split reduction (backward GEMMs with a tall reduction, csrc/gemm_tc.cu): work item t covers K blocks
    // [sp * kb_per_split, ...) of tile t / splits, sp = t % splits, and stores at row offset sp * slab_rows of the y map
    int splits = 1;
    int kb_per_split = 0;
    int64_t slab_rows = 0;
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

// bf16 pair of two values packed as {lo16 = a, hi16 = b} (cvt.rn.bf16x2: first source goes to the upper half)
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// (hi, lo) bf16 pairs of two values: hi = bf16(v), lo = bf16(v - hi), packed {low half = a, high half = b}.
// Done with full-rate integer ops instead of cvt.rn.bf16x2: the conversion instruction issues on the XU pipe (16 lanes
// per clock, shared with ex2 / rcp) and a 128 x 128 tile needs 2 x 8192 of them - measured as the busiest pipe of the
// epilogues that emit pairs. Rounding is half-up on the magnitude (add half an ulp of the 8-bit mantissa, keep the upper
// 16 bits); lo = v - hi is exact in fp32 and absorbs the difference to round-to-nearest-even.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const uint32_t ha = (__float_as_uint(a) + 0x8000u) & 0xffff0000u, hb = (__float_as_uint(b) + 0x8000u) & 0xffff0000u;
    const uint32_t la = __float_as_uint(a - __uint_as_float(ha)) + 0x8000u, lb = __float_as_uint(b - __uint_as_float(hb)) + 0x8000u;
    hi = __byte_perm(ha, hb, 0x7632);      // {hb.hi16, ha.hi16}: lower address = first value
    lo = __byte_perm(la, lb, 0x7632);
}

// scalar store of one result's split in the format the epilogue asks for
__device__ __forceinline__ void store_split(const vlsat_epilogue& e, float v, int64_t m, int64_t n) {
    if (e.split_fmt == VLSAT_SPLIT_BF16) {
        uint32_t hi, lo;
        split_bf16x2(v, 0.f, hi, lo);
        reinterpret_cast<uint16_t*>(e.split_hi)[m * e.ld_split + n] = (uint16_t)(hi & 0xffffu);
        reinterpret_cast<uint16_t*>(e.split_lo)[m * e.ld_split + n] = (uint16_t)(lo & 0xffffu);
    } else {
        split_tf32(v, reinterpret_cast<float*>(e.split_hi)[m * e.ld_split + n], reinterpret_cast<float*>(e.split_lo)[m * e.ld_split + n]);
    }
}

}  // namespace vlsat
