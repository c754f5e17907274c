// MultiHeadAttention with the DENSE additive / multiplicative weights and 0/1 mask of the reference signature
// (ScaledDotProductAttention.forward, src/model/transformer/attention.py:41-78):
//     att = q k^T / sqrt(dk);  att = att * w | att + w;  att[mask == 0] = -inf;  out = softmax(att) v
// This is what the reference's own MMG.forward hands to its attention modules (network_MMG.py:183-205,217-218: a
// [1, H, N, N] distance-bias tensor and a [1, 1, N, N] block-diagonal mask). The fast path of this implementation never
// builds those tensors (attend_scenes evaluates the bias per scene); this kernel exists so that a MultiHeadAttention of
// this package can be swapped in, module by module, under the reference's MMG. Exact fp32 FFMA.
//
// One warp per (query, head): keys in chunks of 32 (lane = key) for the scores, online softmax across chunks, then lane =
// two output dims for the weighted sum of values (the probabilities of the chunk are broadcast by shuffles). Sizes are node
// counts (hundreds to a few thousand), so this is latency-, not throughput-critical.
#include "common.cuh"
#include <float.h>

namespace vlsat {

template <int DK>
__global__ void __launch_bounds__(256)
dense_attn_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk, const float* __restrict__ v,
                  int64_t ldv, const float* __restrict__ w, int64_t w_head_stride, int way, const float* __restrict__ mask,
                  int64_t mask_head_stride, float* __restrict__ out, int64_t ldo, int64_t nq, int64_t nk, int n_heads, float scale) {
    pdl_entry();
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= nq * n_heads) return;
    const int64_t qi = item / n_heads;
    const int h = (int)(item % n_heads);
    constexpr int R = DK / 32;                                    // query / output dims per lane
    float qv[DK];                                                 // the whole query row of this head in every lane
#pragma unroll
    for (int d = 0; d < DK; d += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(q + qi * ldq + h * DK + d));
        qv[d] = t.x; qv[d + 1] = t.y; qv[d + 2] = t.z; qv[d + 3] = t.w;
    }
    const float* wrow = w ? w + h * w_head_stride + qi * nk : nullptr;
    const float* mrow = mask ? mask + h * mask_head_stride + qi * nk : nullptr;
    float m_run = -INFINITY, l_run = 0.f, acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    for (int64_t k0 = 0; k0 < nk; k0 += 32) {
        const int64_t kj = k0 + lane;
        float s = -INFINITY;
        if (kj < nk) {
            const float4* kr = reinterpret_cast<const float4*>(k + kj * ldk + h * DK);
            float dot = 0.f;
#pragma unroll
            for (int d = 0; d < DK / 4; ++d) {
                const float4 t = __ldg(kr + d);
                dot = fmaf(qv[4 * d], t.x, dot); dot = fmaf(qv[4 * d + 1], t.y, dot);
                dot = fmaf(qv[4 * d + 2], t.z, dot); dot = fmaf(qv[4 * d + 3], t.w, dot);
            }
            s = dot * scale;
            if (wrow) s = way == 1 ? s * __ldg(wrow + kj) : s + __ldg(wrow + kj);
            if (mrow && __ldg(mrow + kj) == 0.f) s = -INFINITY;
        }
        const float m_new = fmaxf(m_run, warp_max(s));
        // all keys so far masked: keep the state empty (exp(-inf - -inf) would be NaN)
        const float corr = (m_new == -INFINITY) ? 1.f : __expf(m_run - m_new);
        const float p = (m_new == -INFINITY || s == -INFINITY) ? 0.f : __expf(s - m_new);
        l_run = l_run * corr + warp_sum(p);
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] *= corr;
        const int cnt = (int)min((int64_t)32, nk - k0);
        for (int j = 0; j < cnt; ++j) {
            const float pj = __shfl_sync(0xffffffffu, p, j);
            if (pj != 0.f) {                                      // warp-uniform
                const float* vr = v + (k0 + j) * ldv + h * DK;
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = fmaf(pj, __ldg(vr + lane + 32 * r), acc[r]);
            }
        }
        m_run = m_new;
    }
    // a fully masked row is 0 / 0 = NaN, as torch.softmax of a row of -inf is in the reference
    const float inv = 1.f / l_run;
#pragma unroll
    for (int r = 0; r < R; ++r) out[qi * ldo + h * DK + lane + 32 * r] = acc[r] * inv;
}

}  // namespace vlsat

using namespace vlsat;

extern "C" int vlsat_dense_attn_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                    const float* weights, int64_t weights_head_stride, int way, const float* mask,
                                    int64_t mask_head_stride, float* out, int64_t ldo, int64_t nq, int64_t nk, int n_heads,
                                    int dk, void* stream) {
    VLSAT_REQUIRE(nq >= 0 && nk >= 0 && n_heads >= 1 && (way == 0 || way == 1 || way == 2));
    if (nq == 0) return VLSAT_OK;
    VLSAT_REQUIRE(q && out && (nk == 0 || (k && v)) && (way == 0 || weights));
    const int64_t D = (int64_t)n_heads * dk;
    VLSAT_REQUIRE(ldq >= D && ldk >= D && ldv >= D && ldo >= D);
    VLSAT_SUPPORT(dk == 64 || dk == 32 || dk == 128);
    VLSAT_SUPPORT((ldq | ldk) % 4 == 0 && (((uintptr_t)q | (uintptr_t)k) & 15) == 0);
    VLSAT_SUPPORT(nq * n_heads < (1ll << 31) * 8);
    const float scale = 1.f / sqrtf((float)dk);
    const float* w = way == 0 ? nullptr : weights;
    dim3 grid((unsigned)ceil_div(nq * n_heads, 8)), block(256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dk == 64) launch_k(dense_attn_kernel<64>, grid, block, 0, st, q, ldq, k, ldk, v, ldv, w, weights_head_stride, way, mask, mask_head_stride, out, ldo, nq, nk, n_heads, scale);
    else if (dk == 32) launch_k(dense_attn_kernel<32>, grid, block, 0, st, q, ldq, k, ldk, v, ldv, w, weights_head_stride, way, mask, mask_head_stride, out, ldo, nq, nk, n_heads, scale);
    else launch_k(dense_attn_kernel<128>, grid, block, 0, st, q, ldq, k, ldk, v, ldv, w, weights_head_stride, way, mask, mask_head_stride, out, ldo, nq, nk, n_heads, scale);
    return finish_launch();
}
