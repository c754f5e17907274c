// SURVEY 8(f) N3 - the ranking of --mode eval (src/utils/eva_utils_acc.py:27-79, 137-211; called from
// Mmgnet.process_val, SGFN_MMG/model.py:463-472). The reference loops over edges in python, builds a [160, 160, 26] score
// tensor per edge and SORTS its 665,600 entries on the CPU to find the position of the ground truth. Every rank it
// derives is  1 + #{scores strictly greater than the ground truth's}  (capped at topk + 1), so nothing needs sorting or
// materialising: one CTA per edge streams the products (s_i * o_j) * r_k - the reference's association order, correctly
// rounded multiplies, so its exact float equality test (:184) is reproduced bit for bit - and counts.
// Per-edge lists are emitted sorted with the reference's "i-th smallest minus i" adjustment (:73-78, :205-210):
// ranks_out[e, pos] for pos < number of entries of edge e, INT32_MIN beyond.
// NOT YET RUN ON HARDWARE (round 1 ended without GPU budget): tests carry the gpu_next marker.
#include "common.cuh"
#include <limits.h>

namespace vlsat {

// ---- F.softmax(objs_pred, dim=-1) (:143-145), one warp per row ------------------------------------------------------
__global__ void softmax_rows_kernel(const float* __restrict__ x, int64_t ld, int64_t R, int C, float* __restrict__ y) {
    pdl_entry();
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const float* xr = x + r * ld;
    float m = -3.402823466e+38f;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, xr[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(xr[c] - m);
    s = warp_sum(s);
    for (int c = lane; c < C; c += 32) y[r * C + c] = expf(xr[c] - m) / s;
}

// ---- evaluate_topk_object (:27-39): rank[n] = min(1 + #{c: pred[n, c] > pred[n, gt[n]]}, topk + 1) --------------------
__global__ void topk_object_ranks_kernel(const float* __restrict__ pred, int64_t ld, const int64_t* __restrict__ gt,
                                         int64_t N, int C, int topk, int32_t* __restrict__ out) {
    pdl_entry();
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= N) return;
    const float* p = pred + r * ld;
    int64_t g = gt[r];
    g = g < 0 ? 0 : (g >= C ? C - 1 : g);
    const float ref = p[g];
    int cnt = 0;
    for (int c = lane; c < C; c += 32) cnt += p[c] > ref;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) out[r] = min(cnt + 1, topk + 1);
}

// sorted-minus-counter over the (<= 64) entries of one edge, done by one thread: tiny
__device__ void emit_adjusted(const int* ranks, int m, int32_t* out, int C) {
    for (int a = 0; a < m; ++a) {
        int pos = 0;
        for (int b = 0; b < m; ++b) pos += (ranks[b] < ranks[a]) || (ranks[b] == ranks[a] && b < a);
        out[pos] = ranks[a] - pos;
    }
    for (int a = m; a < C; ++a) out[a] = INT_MIN;             // adjusted ranks of tied labels can reach 0 and below: no small sentinel
}

// ---- evaluate_topk_predicate (:42-79) with get_gt's multi-label targets (:6-24); one warp per edge, C <= 64 ------------
__global__ void topk_predicate_ranks_kernel(const float* __restrict__ rel, const float* __restrict__ gt_rel, int64_t E, int C,
                                            int topk, float threshold, int32_t* __restrict__ out) {
    pdl_entry();
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (e >= E) return;
    const float* p = rel + e * C;
    const float* y = gt_rel + e * C;
    int ranks[64];
    int m = 0;
    for (int g = 0; g < C; ++g) {
        if (y[g] != 1.f) continue;                         // warp-uniform
        const float ref = p[g];
        int cnt = 0;
        for (int c = lane; c < C; c += 32) cnt += p[c] > ref;
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        ranks[m++] = min(cnt + 1, topk + 1);
    }
    if (m == 0) {                                          // no ground-truth relation: first position below the threshold
        int above = 0;
        for (int c = lane; c < C; c += 32) above += p[c] >= threshold;
        above = __reduce_add_sync(0xffffffffu, above);
        ranks[m++] = above == C ? topk + 1 : above + 1;
    }
    if (lane == 0) emit_adjusted(ranks, m, out + e * C, C);
}

// ---- evaluate_triplet_topk (:137-211), ranks only; one CTA per edge -----------------------------------------------------
constexpr int TR_THREADS = 256;
__global__ void __launch_bounds__(TR_THREADS)
topk_triplet_ranks_kernel(const float* __restrict__ obj_prob, int No, const float* __restrict__ rel, int C,
                          const int64_t* __restrict__ gt_cls, const float* __restrict__ gt_rel, const int64_t* __restrict__ edges,
                          int64_t n_nodes, int topk, float threshold, int32_t* __restrict__ out) {
    pdl_entry();
    extern __shared__ float sm[];
    float* s_sub = sm;                 // [No]
    float* s_obj = sm + No;            // [No]
    float* s_rel = sm + 2 * No;        // [C]
    float* s_thr = s_rel + C;          // [C] ground-truth confidences (m of them)
    __shared__ int s_cnt[64];
    __shared__ int s_m;
    const int64_t e = blockIdx.x;
    int64_t a = edges[2 * e], b = edges[2 * e + 1];
    a = a < 0 ? 0 : (a >= n_nodes ? n_nodes - 1 : a);
    b = b < 0 ? 0 : (b >= n_nodes ? n_nodes - 1 : b);
    for (int i = threadIdx.x; i < No; i += TR_THREADS) { s_sub[i] = obj_prob[a * No + i]; s_obj[i] = obj_prob[b * No + i]; }
    for (int k = threadIdx.x; k < C; k += TR_THREADS) s_rel[k] = rel[e * C + k];
    if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t gs = gt_cls[a], go = gt_cls[b];
        gs = gs < 0 ? 0 : (gs >= No ? No - 1 : gs);
        go = go < 0 ? 0 : (go >= No ? No - 1 : go);
        const float node = __fmul_rn(s_sub[gs], s_obj[go]);
        int m = 0;
        for (int k = 0; k < C; ++k)
            if (gt_rel[e * C + k] == 1.f) s_thr[m++] = __fmul_rn(node, s_rel[k]);
        s_m = m;
    }
    __syncthreads();
    const int m = s_m;
    int cnt[4] = {0, 0, 0, 0};                                           // thresholds in groups of four per sweep
    for (int t0 = 0; t0 < (m == 0 ? 1 : m); t0 += 4) {
        float thr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { thr[u] = (t0 + u < m) ? s_thr[t0 + u] : 3.402823466e+38f; cnt[u] = 0; }
        for (int ij = threadIdx.x; ij < No * No; ij += TR_THREADS) {
            const float node = __fmul_rn(s_sub[ij / No], s_obj[ij % No]);
            for (int k = 0; k < C; ++k) {
                const float conf = __fmul_rn(node, s_rel[k]);
                if (m == 0) cnt[0] += conf >= threshold;
                else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) cnt[u] += conf > thr[u];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (t0 + u < max(m, 1)) atomicAdd(&s_cnt[t0 + u], cnt[u]);  // integer adds: order-independent
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int ranks[64];
        const int mm = max(m, 1);
        for (int t = 0; t < mm; ++t) ranks[t] = s_cnt[t] < topk ? s_cnt[t] + 1 : topk + 1;
        emit_adjusted(ranks, mm, out + e * C, C);
    }
}

// 100 * #{valid ranks <= t_j} / #{valid ranks} for three thresholds, one CTA, fixed summation order (deterministic):
// the "R@k" figures process_train logs after every step (SGFN_MMG/model.py:422-432). INT32_MIN marks an empty slot.
__global__ void recall_at_kernel(const int32_t* __restrict__ ranks, int64_t n, int t0, int t1, int t2, float* __restrict__ out) {
    pdl_entry();
    __shared__ int sm[4][32];
    int c[4] = {0, 0, 0, 0};
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const int r = ranks[i];
        if (r != INT_MIN) { ++c[3]; c[0] += r <= t0; c[1] += r <= t1; c[2] += r <= t2; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int v = c[j];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sm[j][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
        int tot[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int v = lane < (int)(blockDim.x >> 5) ? sm[j][lane] : 0;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            tot[j] = v;
        }
        if (lane < 3) out[lane] = tot[3] > 0 ? 100.f * (float)tot[lane] / (float)tot[3] : 0.f;
    }
}

}  // namespace vlsat

using namespace vlsat;

extern "C" int vlsat_softmax_rows(const float* x, int64_t ld, int64_t R, int C, float* y, void* stream) {
    VLSAT_REQUIRE(R >= 0 && C >= 1 && ld >= C);
    if (R == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && y);
    launch_k(softmax_rows_kernel, dim3((unsigned)ceil_div(R * 32, 256)), dim3(256), 0, (cudaStream_t)stream, x, ld, R, C, y);
    return finish_launch();
}

extern "C" int vlsat_topk_object_ranks(const float* pred, int64_t ld, const int64_t* target, int64_t N, int C, int topk,
                                       int32_t* ranks, void* stream) {
    VLSAT_REQUIRE(N >= 0 && C >= 1 && ld >= C && topk >= 1);
    if (N == 0) return VLSAT_OK;
    VLSAT_REQUIRE(pred && target && ranks);
    launch_k(topk_object_ranks_kernel, dim3((unsigned)ceil_div(N * 32, 256)), dim3(256), 0, (cudaStream_t)stream, pred, ld, target, N, C, topk, ranks);
    return finish_launch();
}

extern "C" int vlsat_topk_predicate_ranks(const float* rel_prob, const float* gt_rel, int64_t E, int C, int topk, float threshold,
                                          int32_t* ranks, void* stream) {
    VLSAT_REQUIRE(E >= 0 && C >= 1 && topk >= 1);
    VLSAT_SUPPORT(C <= 64);
    if (E == 0) return VLSAT_OK;
    VLSAT_REQUIRE(rel_prob && gt_rel && ranks);
    launch_k(topk_predicate_ranks_kernel, dim3((unsigned)ceil_div(E * 32, 256)), dim3(256), 0, (cudaStream_t)stream, rel_prob, gt_rel, E, C, topk, threshold, ranks);
    return finish_launch();
}

extern "C" int vlsat_topk_triplet_ranks(const float* obj_prob, int64_t n_nodes, int n_obj_cls, const float* rel_prob, int n_rel_cls,
                                        const int64_t* gt_cls, const float* gt_rel, const int64_t* edges, int64_t E, int topk,
                                        float threshold, int32_t* ranks, void* stream) {
    VLSAT_REQUIRE(E >= 0 && n_nodes >= 1 && n_obj_cls >= 1 && n_rel_cls >= 1 && topk >= 1);
    VLSAT_SUPPORT(n_rel_cls <= 64 && n_obj_cls <= 4096 && (int64_t)n_obj_cls * n_obj_cls * n_rel_cls < (1ll << 31));
    if (E == 0) return VLSAT_OK;
    VLSAT_REQUIRE(obj_prob && rel_prob && gt_cls && gt_rel && edges && ranks);
    VLSAT_SUPPORT(E < (1ll << 31));
    const size_t smem = (size_t)(2 * n_obj_cls + 2 * n_rel_cls) * sizeof(float);
    launch_k(topk_triplet_ranks_kernel, dim3((unsigned)E), dim3(TR_THREADS), smem, (cudaStream_t)stream, obj_prob, n_obj_cls, rel_prob, n_rel_cls,
             gt_cls, gt_rel, edges, n_nodes, topk, threshold, ranks);
    return finish_launch();
}

extern "C" int vlsat_recall_at(const int32_t* ranks, int64_t n, int t0, int t1, int t2, float* out3, void* stream) {
    VLSAT_REQUIRE(n >= 0 && out3 && (n == 0 || ranks));
    launch_k(recall_at_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, ranks, n, t0, t1, t2, out3);
    return finish_launch();
}
