// SURVEY 8(f) N1 - the training-step glue around the path:
//   * the losses of Mmgnet.process_train (src/model/SGFN_MMG/model.py:343-418): two cross-entropies over the object
//     logits, two weighted binary cross-entropies over the relationship probabilities with the per-batch DYNAMIC class
//     weights (:353-366), the cosine-margin "mimic" loss between the 3D and 2D object features (:257-258, :402-404) and
//     the L1 distance between the unit-normalised 2D edge feature and a text embedding (:409-410);
//   * Mmgnet.backward's optimiser step (:483-488): AdamW over the 13 parameter groups of :143-156 with the cosine
//     learning-rate schedule of :157, as ONE multi-tensor kernel.
// Every loss is one warp per row writing a row loss; a single-CTA pass sums the rows in a fixed order (bitwise
// reproducible, no float atomics). Backward kernels read the upstream gradient from device memory (no host sync).
// All of it is HBM-bound row work: 16 B loaded + 12 B stored per parameter element for AdamW, one read of each loss
// operand forward and one read + one write backward.
#include "common.cuh"
#include <float.h>

namespace vlsat {

constexpr int GL_THREADS = 256;                 // 8 rows (warps) per CTA

__device__ __forceinline__ int64_t gl_row() { return ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; }
static inline unsigned gl_grid(int64_t rows) { return (unsigned)ceil_div(rows * 32, GL_THREADS); }

// ---------------------------------------------------------------------------------------------------- cross entropy
__global__ void cross_entropy_fwd_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ target,
                                         int64_t R, int C, float* __restrict__ row_loss, float* __restrict__ row_lse) {
    pdl_entry();
    const int64_t r = gl_row(); const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const float* x = logits + r * ld;
    float m = -FLT_MAX;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, x[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(x[c] - m);
    s = warp_sum(s);
    const float lse = m + logf(s);
    if (lane == 0) {
        const int64_t t = target[r];
        row_lse[r] = lse;
        row_loss[r] = (t >= 0 && t < C) ? lse - x[t] : 0.f;
    }
}

__global__ void cross_entropy_bwd_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ target,
                                         const float* __restrict__ row_lse, const float* __restrict__ gout, float coef,
                                         float* __restrict__ dlogits, int64_t ldd, int64_t R, int C) {
    pdl_entry();
    const int64_t r = gl_row(); const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const int64_t t = target[r];
    const bool ok = t >= 0 && t < C;
    const float g = ok ? gout[0] * coef : 0.f, lse = row_lse[r];
    const float* x = logits + r * ld;
    float* d = dlogits + r * ldd;
    for (int c = lane; c < C; c += 32) d[c] = g * (expf(x[c] - lse) - (c == t ? 1.f : 0.f));
}

// ----------------------------------------------------------------------------- DYNAMIC class weights (model.py:353-366)
// weight[c] = scale / (log(sum_e gt[e, c] + 1) + 1). The "none" slot the reference prepends (count of label-free edges)
// is dropped again by weight[1:], and no weight can be zero, so it never reaches the loss. Counts of 0/1 labels are exact
// in fp32: the sum order does not matter. One CTA: 16 row lanes x 64 class lanes.
__global__ void rel_class_weights_kernel(const float* __restrict__ gt, int64_t E, int C, float scale, float* __restrict__ weight) {
    pdl_entry();
    __shared__ float part[16][65];
    const int c = threadIdx.x & 63, rl = threadIdx.x >> 6;
    float acc = 0.f;
    if (c < C)
        for (int64_t e = rl; e < E; e += 16) acc += gt[e * C + c];
    part[rl][c] = acc;
    __syncthreads();
    if (rl == 0 && c < C) {
        float cnt = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) cnt += part[i][c];
        weight[c] = fabsf(scale / (logf(cnt + 1.f) + 1.f));
    }
}

// -------------------------------------------------------------------------------- weighted binary cross entropy
// F.binary_cross_entropy(p, y, weight): -w_c (y max(log p, -100) + (1 - y) max(log1p(-p), -100)), mean over E*C
__global__ void bce_fwd_kernel(const float* __restrict__ p, const float* __restrict__ y, const float* __restrict__ w,
                               int64_t E, int C, float* __restrict__ row_loss) {
    pdl_entry();
    const int64_t r = gl_row(); const int lane = threadIdx.x & 31;
    if (r >= E) return;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float pv = p[r * C + c], yv = y[r * C + c];
        const float l = (yv - 1.f) * fmaxf(log1pf(-pv), -100.f) - yv * fmaxf(logf(pv), -100.f);
        s += (w ? w[c] : 1.f) * l;
    }
    s = warp_sum(s);
    if (lane == 0) row_loss[r] = s;
}
// d/dp = w_c (p - y) / max((1 - p) p, 1e-12)   (ATen's binary_cross_entropy_backward)
__global__ void bce_bwd_kernel(const float* __restrict__ p, const float* __restrict__ y, const float* __restrict__ w,
                               const float* __restrict__ gout, float coef, float* __restrict__ dp, int64_t n, int C) {
    pdl_entry();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float pv = p[i], yv = y[i];
    dp[i] = gout[0] * coef * (w ? w[i % C] : 1.f) * (pv - yv) / fmaxf((1.f - pv) * pv, 1e-12f);
}

// ---------------------------------------------------------------------------------------- cosine margin ("mimic")
// row loss = max(margin - cos(a, b), 0); the reference normalises both rows first (model.py:402-403), which leaves the
// cosine unchanged
__global__ void cosine_margin_fwd_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb,
                                         int64_t R, int D, float margin, float* __restrict__ row_loss) {
    pdl_entry();
    const int64_t r = gl_row(); const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const float* x = a + r * lda; const float* z = b + r * ldb;
    float dot = 0.f, na = 0.f, nb = 0.f;
    for (int c = lane; c < D; c += 32) { const float u = x[c], v = z[c]; dot += u * v; na += u * u; nb += v * v; }
    dot = warp_sum(dot); na = warp_sum(na); nb = warp_sum(nb);
    const float cs = dot / (fmaxf(sqrtf(na), 1e-8f) * fmaxf(sqrtf(nb), 1e-8f));
    if (lane == 0) row_loss[r] = fmaxf(margin - cs, 0.f);
}
__global__ void cosine_margin_bwd_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb,
                                         const float* __restrict__ gout, float coef, float margin,
                                         float* __restrict__ da, int64_t ldda, float* __restrict__ db, int64_t lddb, int64_t R, int D) {
    pdl_entry();
    const int64_t r = gl_row(); const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const float* x = a + r * lda; const float* z = b + r * ldb;
    float dot = 0.f, na = 0.f, nb = 0.f;
    for (int c = lane; c < D; c += 32) { const float u = x[c], v = z[c]; dot += u * v; na += u * u; nb += v * v; }
    dot = warp_sum(dot); na = warp_sum(na); nb = warp_sum(nb);
    const float la = fmaxf(sqrtf(na), 1e-8f), lb = fmaxf(sqrtf(nb), 1e-8f);
    const float cs = dot / (la * lb);
    const float g = (margin - cs > 0.f) ? -gout[0] * coef : 0.f;          // d loss / d cos = -1 where the margin is active
    const float inv = 1.f / (la * lb), ca = cs / (la * la), cb = cs / (lb * lb);
    for (int c = lane; c < D; c += 32) {
        const float u = x[c], v = z[c];
        if (da) da[r * ldda + c] = g * (v * inv - ca * u);
        if (db) db[r * lddb + c] = g * (u * inv - cb * v);
    }
}

// ------------------------------------------------------------------------- L1 between x / |x| and a target (model.py:409-410)
__global__ void l1_unit_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ t, int64_t ldt,
                                   int64_t R, int D, float* __restrict__ row_loss) {
    pdl_entry();
    const int64_t r = gl_row(); const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const float* xr = x + r * ldx; const float* tr = t + r * ldt;
    float n2 = 0.f;
    for (int c = lane; c < D; c += 32) n2 += xr[c] * xr[c];
    const float n = sqrtf(warp_sum(n2));
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += fabsf(xr[c] / n - tr[c]);
    s = warp_sum(s);
    if (lane == 0) row_loss[r] = s;
}
__global__ void l1_unit_bwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ t, int64_t ldt,
                                   const float* __restrict__ gout, float coef, float* __restrict__ dx, int64_t lddx, int64_t R, int D) {
    pdl_entry();
    const int64_t r = gl_row(); const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const float* xr = x + r * ldx; const float* tr = t + r * ldt;
    float n2 = 0.f;
    for (int c = lane; c < D; c += 32) n2 += xr[c] * xr[c];
    const float n = sqrtf(warp_sum(n2));
    const float g = gout[0] * coef;
    float s = 0.f;                                             // sum_c sign_c xhat_c
    for (int c = lane; c < D; c += 32) {
        const float xh = xr[c] / n, d = xh - tr[c];
        s += (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * xh;
    }
    s = warp_sum(s);
    for (int c = lane; c < D; c += 32) {
        const float xh = xr[c] / n, d = xh - tr[c];
        const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        dx[r * lddx + c] = g * (sg - s * xh) / n;
    }
}

// --------------------------------------------------------------------------- fixed-order sum of a row-loss buffer
__global__ void sum_rows_kernel(const float* __restrict__ v, int64_t n, float scale, float* __restrict__ out_term,
                                float* __restrict__ out_total, float total_coef, int accumulate) {
    pdl_entry();
    __shared__ float red[1024];
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += 1024) s += v[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float term = red[0] * scale;
        if (out_term) out_term[0] = term;
        if (out_total) out_total[0] = (accumulate ? out_total[0] : 0.f) + total_coef * term;
    }
}

// ------------------------------------------------------------------------------------------------ multi-tensor AdamW
// torch.optim.AdamW semantics (decoupled decay, bias corrections, optional amsgrad) for every chunk of every tensor of
// the table in one launch; the learning rate of a tensor is its group's base rate times the CosineAnnealingLR factor
// (eta_min = 0) of the step counter kept in device memory, so the launch is CUDA-graph replayable.
__global__ void adamw_kernel(const vlsat_adamw_tensor* __restrict__ tab, const int32_t* __restrict__ chunk_tensor,
                             const int32_t* __restrict__ chunk_index, int chunk_elems, double beta1, double beta2, float eps,
                             const int64_t* __restrict__ step, int64_t t_max) {
    pdl_entry();
    const int64_t k = step[0] + 1;                                   // this is the k-th step (1-based)
    const double sched = t_max > 0 ? 0.5 * (1.0 + cospi((double)(k - 1) / (double)t_max)) : 1.0;
    const double bc1 = 1.0 - pow(beta1, (double)k), bc2 = 1.0 - pow(beta2, (double)k);
    const vlsat_adamw_tensor T = tab[chunk_tensor[blockIdx.x]];
    const float lr = (float)((double)T.lr * sched);
    const float decay = 1.f - lr * T.weight_decay, step_size = (float)((double)lr / bc1), rs_bc2 = (float)(1.0 / sqrt(bc2));
    const float b1 = (float)beta1, b2 = (float)beta2;
    const int64_t base = (int64_t)chunk_index[blockIdx.x] * chunk_elems;
    const int64_t end = base + chunk_elems < T.n ? base + chunk_elems : T.n;
    auto update = [&](float p, float g, float& m, float& v, float& vm) {
        p *= decay;
        m = m + (g - m) * (1.f - b1);
        v = v * b2 + (1.f - b2) * g * g;
        float den = v;
        if (T.vmax) { vm = fmaxf(vm, v); den = vm; }
        return p - step_size * m / (sqrtf(den) * rs_bc2 + eps);
    };
    const bool vec = ((((uintptr_t)T.p | (uintptr_t)T.g | (uintptr_t)T.m | (uintptr_t)T.v | (uintptr_t)T.vmax) & 15) == 0) && (chunk_elems % 4 == 0);
    int64_t i0 = base;
    if (vec) {
        const int64_t n4 = (end - base) >> 2;
        for (int64_t j = threadIdx.x; j < n4; j += blockDim.x) {
            const int64_t i = base + 4 * j;
            float4 p = *reinterpret_cast<const float4*>(T.p + i), m = *reinterpret_cast<const float4*>(T.m + i), v = *reinterpret_cast<const float4*>(T.v + i);
            const float4 g = __ldg(reinterpret_cast<const float4*>(T.g + i));
            float4 vm = T.vmax ? *reinterpret_cast<const float4*>(T.vmax + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            p.x = update(p.x, g.x, m.x, v.x, vm.x); p.y = update(p.y, g.y, m.y, v.y, vm.y);
            p.z = update(p.z, g.z, m.z, v.z, vm.z); p.w = update(p.w, g.w, m.w, v.w, vm.w);
            *reinterpret_cast<float4*>(T.p + i) = p; *reinterpret_cast<float4*>(T.m + i) = m; *reinterpret_cast<float4*>(T.v + i) = v;
            if (T.vmax) *reinterpret_cast<float4*>(T.vmax + i) = vm;
        }
        i0 = base + 4 * n4;
    }
    for (int64_t i = i0 + threadIdx.x; i < end; i += blockDim.x) {
        float m = T.m[i], v = T.v[i], vm = T.vmax ? T.vmax[i] : 0.f;
        T.p[i] = update(T.p[i], T.g[i], m, v, vm);
        T.m[i] = m; T.v[i] = v;
        if (T.vmax) T.vmax[i] = vm;
    }
}
// dst_t[i] = scale * src_t[i] for every chunk of every tensor of the table, one launch: packs the gradients of a step
// into the flat buffer the NCCL all-reduce runs on, with the 1 / world_size of the mean folded in (train_glue / dist.py).
__global__ void pack_scale_kernel(const vlsat_copy_tensor* __restrict__ tab, const int32_t* __restrict__ chunk_tensor,
                                  const int32_t* __restrict__ chunk_index, int chunk_elems, float scale) {
    pdl_entry();
    const vlsat_copy_tensor T = tab[chunk_tensor[blockIdx.x]];
    const int64_t base = (int64_t)chunk_index[blockIdx.x] * chunk_elems;
    const int64_t end = base + chunk_elems < T.n ? base + chunk_elems : T.n;
    int64_t i0 = base;
    if (((((uintptr_t)T.dst | (uintptr_t)T.src) & 15) == 0) && (chunk_elems % 4 == 0)) {
        const int64_t n4 = (end - base) >> 2;
        for (int64_t j = threadIdx.x; j < n4; j += blockDim.x) {
            float4 v = __ldg(reinterpret_cast<const float4*>(T.src + base + 4 * j));
            v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
            *reinterpret_cast<float4*>(T.dst + base + 4 * j) = v;
        }
        i0 = base + 4 * n4;
    }
    for (int64_t i = i0 + threadIdx.x; i < end; i += blockDim.x) T.dst[i] = T.src[i] * scale;
}
// N4: the text supervision target of an edge from a cached table of prompt features (get_rel_emb, SGFN_MMG/model.py:221-255):
// mean over the edge's ground-truth predicates of table[cls[subject], cls[object], r, :] (row R = "no relation" when it has
// none), then L2 normalisation. One warp per edge, D / 32 features per lane, fixed summation order.
template <int PER>
__global__ void rel_text_embed_kernel(const float* __restrict__ table, int n_obj, int n_rel, const int64_t* __restrict__ gt_cls,
                                      const float* __restrict__ gt_rel, int64_t ld_rel, const int64_t* __restrict__ edges,
                                      int64_t E, float* __restrict__ out, int64_t ld_out) {
    pdl_entry();
    const int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= E) return;
    const int lane = threadIdx.x & 31;
    constexpr int D = PER * 32;
    const int64_t s = gt_cls[edges[2 * e]], o = gt_cls[edges[2 * e + 1]];
    const float* base = table + ((s * n_obj + o) * (int64_t)(n_rel + 1)) * D;
    float acc[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) acc[i] = 0.f;
    int cnt = 0;
    for (int r = 0; r < n_rel; ++r) {
        if (gt_rel[e * ld_rel + r] == 1.f) {                     // warp-uniform
            ++cnt;
#pragma unroll
            for (int i = 0; i < PER; ++i) acc[i] += __ldg(base + (int64_t)r * D + lane + 32 * i);
        }
    }
    if (cnt == 0) {
#pragma unroll
        for (int i = 0; i < PER; ++i) acc[i] = __ldg(base + (int64_t)n_rel * D + lane + 32 * i);
    } else {
        const float inv = 1.f / (float)cnt;
#pragma unroll
        for (int i = 0; i < PER; ++i) acc[i] *= inv;
    }
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) ss += acc[i] * acc[i];
    const float rn = 1.f / sqrtf(warp_sum(ss));
#pragma unroll
    for (int i = 0; i < PER; ++i) out[e * ld_out + lane + 32 * i] = acc[i] * rn;
}
__global__ void bump_step_kernel(int64_t* step) {
    pdl_entry();
    if (threadIdx.x == 0 && blockIdx.x == 0) step[0] += 1;
}

}  // namespace vlsat

using namespace vlsat;

extern "C" int vlsat_cross_entropy_fwd(const float* logits, int64_t ld, const int64_t* target, int64_t R, int C,
                                       float* row_loss, float* row_lse, void* stream) {
    VLSAT_REQUIRE(R >= 0 && C >= 1 && ld >= C);
    if (R == 0) return VLSAT_OK;
    VLSAT_REQUIRE(logits && target && row_loss && row_lse);
    launch_k(cross_entropy_fwd_kernel, dim3(gl_grid(R)), dim3(GL_THREADS), 0, (cudaStream_t)stream, logits, ld, target, R, C, row_loss, row_lse);
    return finish_launch();
}

extern "C" int vlsat_cross_entropy_bwd(const float* logits, int64_t ld, const int64_t* target, const float* row_lse,
                                       const float* gout, float coef, float* dlogits, int64_t ldd, int64_t R, int C, void* stream) {
    VLSAT_REQUIRE(R >= 0 && C >= 1 && ld >= C && ldd >= C);
    if (R == 0) return VLSAT_OK;
    VLSAT_REQUIRE(logits && target && row_lse && gout && dlogits);
    launch_k(cross_entropy_bwd_kernel, dim3(gl_grid(R)), dim3(GL_THREADS), 0, (cudaStream_t)stream, logits, ld, target, row_lse, gout, coef, dlogits, ldd, R, C);
    return finish_launch();
}

extern "C" int vlsat_rel_class_weights(const float* gt, int64_t E, int C, float scale, float* weight, void* stream) {
    VLSAT_REQUIRE(E >= 0 && C >= 1);
    VLSAT_SUPPORT(C <= 64);
    VLSAT_REQUIRE(weight && (E == 0 || gt));
    launch_k(rel_class_weights_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, gt, E, C, scale, weight);
    return finish_launch();
}

extern "C" int vlsat_bce_fwd(const float* p, const float* y, const float* weight, int64_t E, int C, float* row_loss, void* stream) {
    VLSAT_REQUIRE(E >= 0 && C >= 1);
    if (E == 0) return VLSAT_OK;
    VLSAT_REQUIRE(p && y && row_loss);
    launch_k(bce_fwd_kernel, dim3(gl_grid(E)), dim3(GL_THREADS), 0, (cudaStream_t)stream, p, y, weight, E, C, row_loss);
    return finish_launch();
}

extern "C" int vlsat_bce_bwd(const float* p, const float* y, const float* weight, const float* gout, float coef, float* dp,
                             int64_t E, int C, void* stream) {
    VLSAT_REQUIRE(E >= 0 && C >= 1);
    if (E == 0) return VLSAT_OK;
    VLSAT_REQUIRE(p && y && gout && dp);
    launch_k(bce_bwd_kernel, dim3((unsigned)ceil_div(E * C, 256)), dim3(256), 0, (cudaStream_t)stream, p, y, weight, gout, coef, dp, E * C, C);
    return finish_launch();
}

extern "C" int vlsat_cosine_margin_fwd(const float* a, int64_t lda, const float* b, int64_t ldb, int64_t R, int D, float margin,
                                       float* row_loss, void* stream) {
    VLSAT_REQUIRE(R >= 0 && D >= 1 && lda >= D && ldb >= D);
    if (R == 0) return VLSAT_OK;
    VLSAT_REQUIRE(a && b && row_loss);
    launch_k(cosine_margin_fwd_kernel, dim3(gl_grid(R)), dim3(GL_THREADS), 0, (cudaStream_t)stream, a, lda, b, ldb, R, D, margin, row_loss);
    return finish_launch();
}

extern "C" int vlsat_cosine_margin_bwd(const float* a, int64_t lda, const float* b, int64_t ldb, const float* gout, float coef,
                                       float margin, float* da, int64_t ldda, float* db, int64_t lddb, int64_t R, int D, void* stream) {
    VLSAT_REQUIRE(R >= 0 && D >= 1 && lda >= D && ldb >= D);
    if (R == 0 || (!da && !db)) return VLSAT_OK;
    VLSAT_REQUIRE(a && b && gout && (!da || ldda >= D) && (!db || lddb >= D));
    launch_k(cosine_margin_bwd_kernel, dim3(gl_grid(R)), dim3(GL_THREADS), 0, (cudaStream_t)stream, a, lda, b, ldb, gout, coef, margin, da, ldda, db, lddb, R, D);
    return finish_launch();
}

extern "C" int vlsat_l1_unit_fwd(const float* x, int64_t ldx, const float* target, int64_t ldt, int64_t R, int D, float* row_loss, void* stream) {
    VLSAT_REQUIRE(R >= 0 && D >= 1 && ldx >= D && ldt >= D);
    if (R == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && target && row_loss);
    launch_k(l1_unit_fwd_kernel, dim3(gl_grid(R)), dim3(GL_THREADS), 0, (cudaStream_t)stream, x, ldx, target, ldt, R, D, row_loss);
    return finish_launch();
}

extern "C" int vlsat_l1_unit_bwd(const float* x, int64_t ldx, const float* target, int64_t ldt, const float* gout, float coef,
                                 float* dx, int64_t lddx, int64_t R, int D, void* stream) {
    VLSAT_REQUIRE(R >= 0 && D >= 1 && ldx >= D && ldt >= D && lddx >= D);
    if (R == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && target && gout && dx);
    launch_k(l1_unit_bwd_kernel, dim3(gl_grid(R)), dim3(GL_THREADS), 0, (cudaStream_t)stream, x, ldx, target, ldt, gout, coef, dx, lddx, R, D);
    return finish_launch();
}

extern "C" int vlsat_sum_rows(const float* v, int64_t n, float scale, float* out_term, float* out_total, float total_coef,
                              int accumulate, void* stream) {
    VLSAT_REQUIRE(n >= 0 && (out_term || out_total) && (n == 0 || v));
    launch_k(sum_rows_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, v, n, scale, out_term, out_total, total_coef, accumulate);
    return finish_launch();
}

extern "C" int vlsat_adamw_step(const vlsat_adamw_tensor* tensors, const int32_t* chunk_tensor, const int32_t* chunk_index,
                                int64_t n_chunks, int chunk_elems, double beta1, double beta2, float eps, int64_t* step,
                                int64_t t_max, void* stream) {
    VLSAT_REQUIRE(n_chunks >= 0 && chunk_elems >= 1 && step);
    VLSAT_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.f);
    VLSAT_SUPPORT(n_chunks < (1ll << 31));
    int launches = 1;
    if (n_chunks > 0) {
        VLSAT_REQUIRE(tensors && chunk_tensor && chunk_index);
        launch_k(adamw_kernel, dim3((unsigned)n_chunks), dim3(256), 0, (cudaStream_t)stream, tensors, chunk_tensor, chunk_index, chunk_elems,
                 beta1, beta2, eps, (const int64_t*)step, t_max);
        ++launches;
    }
    launch_k(bump_step_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, step);
    return finish_launch(launches);
}

extern "C" int vlsat_pack_scale(const vlsat_copy_tensor* tensors, const int32_t* chunk_tensor, const int32_t* chunk_index,
                                int64_t n_chunks, int chunk_elems, float scale, void* stream) {
    VLSAT_REQUIRE(n_chunks >= 0 && chunk_elems >= 1);
    VLSAT_SUPPORT(n_chunks < (1ll << 31));
    if (n_chunks == 0) return VLSAT_OK;
    VLSAT_REQUIRE(tensors && chunk_tensor && chunk_index);
    launch_k(pack_scale_kernel, dim3((unsigned)n_chunks), dim3(256), 0, (cudaStream_t)stream, tensors, chunk_tensor, chunk_index, chunk_elems, scale);
    return finish_launch();
}

extern "C" int vlsat_rel_text_embed(const float* table, int n_obj_cls, int n_rel_cls, int dim, const int64_t* gt_cls,
                                    const float* gt_rel, int64_t ld_rel, const int64_t* edges, int64_t E, float* out,
                                    int64_t ld_out, void* stream) {
    VLSAT_REQUIRE(E >= 0 && n_obj_cls >= 1 && n_rel_cls >= 1 && ld_rel >= n_rel_cls && ld_out >= dim);
    if (E == 0) return VLSAT_OK;
    VLSAT_REQUIRE(table && gt_cls && gt_rel && edges && out);
    VLSAT_SUPPORT(dim == 512 || dim == 768 || dim == 256);
    dim3 grid((unsigned)ceil_div(E, 8)), block(256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dim == 512) launch_k(rel_text_embed_kernel<16>, grid, block, 0, st, table, n_obj_cls, n_rel_cls, gt_cls, gt_rel, ld_rel, edges, E, out, ld_out);
    else if (dim == 768) launch_k(rel_text_embed_kernel<24>, grid, block, 0, st, table, n_obj_cls, n_rel_cls, gt_cls, gt_rel, ld_rel, edges, E, out, ld_out);
    else launch_k(rel_text_embed_kernel<8>, grid, block, 0, st, table, n_obj_cls, n_rel_cls, gt_cls, gt_rel, ld_rel, edges, E, out, ld_out);
    return finish_launch();
}
