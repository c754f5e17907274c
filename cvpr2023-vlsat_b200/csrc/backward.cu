// Backward / training-mode kernels of the path that are not GEMMs. The GEMM-shaped parts of every backward
// (dX = dZ W, dW = dZ^T X) reuse vlsat_linear_fwd on transposed operands (vlsat_transpose below), so this file
// holds the memory-bound pieces: activation/bias backward, row scatter-add (backward of the row gathers),
// LayerNorm backward, the graph-attention softmax + aggregation (forward with saved arg-max, and backward),
// the score-matrix stage of the edge cross-attention backward, PointNet max-pool backward, dropout,
// BatchNorm1d with batch statistics, and row-L2-normalisation backward.
//
// Reference semantics: autograd of the modules cited in include/vlsat_b200.h (the reference has no hand-written
// backward; torch.autograd over network_MMG.py / network_PointNet.py / attention.py defines it).
#include "common.cuh"
#include "epilogue.cuh"
#include <float.h>

namespace vlsat {

// ------------------------------------------------------------------------------------------ transpose
// out[b, c, r] = in[b, r, c]; columns [rows, ld_out) of every output row are zero-filled so the result can
// feed a GEMM whose reduction length is rounded up to a multiple of 4.
__global__ void transpose_kernel(const float* __restrict__ in, int64_t ld_in, int64_t bs_in, float* __restrict__ out,
                                 int64_t ld_out, int64_t bs_out, int64_t rows, int64_t cols) {
    pdl_entry();
    __shared__ float tile[32][33];
    const int64_t b = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    const float* ip = in + b * bs_in;
    float* op = out + b * bs_out;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int64_t r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? ip[r * ld_in + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int64_t c = c0 + i, r = r0 + threadIdx.x;
        if (c < cols && r < ld_out) op[c * ld_out + r] = tile[threadIdx.x][i];    // r >= rows -> 0 from the load guard
    }
}

// ------------------------------------------------------------------------------------ activation backward
// dz = dy * act'(y) * scale (* exp(*scale_ptr));  dbias[n] += sum_m dz[m, n] (atomic; caller zero-fills).
// y is the forward OUTPUT of the activation (relu: y > 0; sigmoid: y (1 - y)).
constexpr int AB_ROWS = 64;       // rows per block
__global__ void act_bwd_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ y, int64_t ldy,
                               int act, float scale, const float* __restrict__ scale_ptr, float* __restrict__ dz,
                               int64_t lddz, float* __restrict__ dbias, int64_t M, int64_t N) {
    pdl_entry();
    __shared__ float part[8][33];
    const int64_t n = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.y * AB_ROWS;
    const float sc = scale * (scale_ptr ? expf(__ldg(scale_ptr)) : 1.f);
    float acc = 0.f;
    if (n < N) {
        for (int i = threadIdx.y; i < AB_ROWS; i += 8) {
            const int64_t m = m0 + i;
            if (m >= M) break;
            float g = dy[m * lddy + n] * sc;
            if (act == VLSAT_ACT_RELU) { if (!(y[m * ldy + n] > 0.f)) g = 0.f; }
            else if (act == VLSAT_ACT_SIGMOID) { const float t = y[m * ldy + n]; g *= t * (1.f - t); }
            if (dz) dz[m * lddz + n] = g;
            acc += g;
        }
    }
    if (dbias) {
        part[threadIdx.y][threadIdx.x] = acc;
        __syncthreads();
        if (threadIdx.y == 0 && n < N) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += part[i][threadIdx.x];
            atomicAdd(dbias + n, s);
        }
    }
}

// Vectorised form (N % 4 == 0, 16-byte aligned rows): four columns per thread, and optionally the bf16 (hi, lo) pair of
// dz written in the same pass - dz feeds the two backward GEMMs of its projection (dX = dZ W, dW = dZ^T X), which read bf16
// pairs; emitting them here removes one split launch per projection backward.
__global__ void act_bwd_vec_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ y, int64_t ldy,
                                   int act, float scale, const float* __restrict__ scale_ptr, float* __restrict__ dz,
                                   int64_t lddz, float* __restrict__ dbias, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                   int64_t ld_pair, int64_t M, int64_t N) {
    pdl_entry();
    __shared__ float4 part[8][33];
    const int64_t n = ((int64_t)blockIdx.x * 32 + threadIdx.x) * 4;
    const int64_t m0 = (int64_t)blockIdx.y * AB_ROWS;
    const float sc = scale * (scale_ptr ? expf(__ldg(scale_ptr)) : 1.f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) {
        for (int i = threadIdx.y; i < AB_ROWS; i += 8) {
            const int64_t m = m0 + i;
            if (m >= M) break;
            float4 g = __ldg(reinterpret_cast<const float4*>(dy + m * lddy + n));
            g.x *= sc; g.y *= sc; g.z *= sc; g.w *= sc;
            if (act != VLSAT_ACT_NONE) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(y + m * ldy + n));
                if (act == VLSAT_ACT_RELU) {
                    if (!(t.x > 0.f)) g.x = 0.f; if (!(t.y > 0.f)) g.y = 0.f; if (!(t.z > 0.f)) g.z = 0.f; if (!(t.w > 0.f)) g.w = 0.f;
                } else {
                    g.x *= t.x * (1.f - t.x); g.y *= t.y * (1.f - t.y); g.z *= t.z * (1.f - t.z); g.w *= t.w * (1.f - t.w);
                }
            }
            if (dz) *reinterpret_cast<float4*>(dz + m * lddz + n) = g;
            if (hi) {
                uint32_t h0, l0, h1, l1;
                split_bf16x2(g.x, g.y, h0, l0); split_bf16x2(g.z, g.w, h1, l1);
                *reinterpret_cast<uint2*>(hi + m * ld_pair + n) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(lo + m * ld_pair + n) = make_uint2(l0, l1);
            }
            acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
        }
    }
    if (dbias) {
        part[threadIdx.y][threadIdx.x] = acc;
        __syncthreads();
        if (threadIdx.y == 0 && n < N) {
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float4 t = part[i][threadIdx.x]; s4.x += t.x; s4.y += t.y; s4.z += t.z; s4.w += t.w; }
            atomicAdd(dbias + n, s4.x); atomicAdd(dbias + n + 1, s4.y); atomicAdd(dbias + n + 2, s4.z); atomicAdd(dbias + n + 3, s4.w);
        }
    }
}

// --------------------------------------------------------------------------------------- row scatter-add
// out[idx[i] * idx_mul + idx_add(i), :] += in[i, :]. With rows_per_idx = R > 1, row i uses idx[i / R] * R + i % R
// (rows (e, h) of a head-major edge tensor scattering onto rows (node, h)).
__global__ void scatter_add_rows_kernel(const float* __restrict__ in, int64_t ld_in, const int64_t* __restrict__ idx,
                                        int rows_per_idx, int64_t rows, int cols, float* __restrict__ out, int64_t ld_out) {
    pdl_entry();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = t / cols; const int c = (int)(t % cols);
    if (r >= rows) return;
    const int64_t tr = idx[r / rows_per_idx] * rows_per_idx + (r % rows_per_idx);
    const float v = in[r * ld_in + c];
    if (v != 0.f) atomicAdd(out + tr * ld_out + c, v);
}

// out[i, :] = in[idx[i / R] * R + i % R, :]   (forward of the above; also the row gather of head-major operands)
// Four columns per thread, one vector atomic (red.global.add.v4.f32, sm_90+) per 16 bytes: a quarter of the atomic
// instructions and of the index arithmetic of the scalar kernel.
__global__ void scatter_add_rows_vec_kernel(const float* __restrict__ in, int64_t ld_in, const int64_t* __restrict__ idx,
                                            int rows_per_idx, int64_t rows, int cols4, float* __restrict__ out, int64_t ld_out) {
    pdl_entry();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = t / cols4; const int c = (int)(t - r * cols4) * 4;
    if (r >= rows) return;
    const int64_t tr = idx[r / rows_per_idx] * rows_per_idx + (r % rows_per_idx);
    const float4 v = __ldg(reinterpret_cast<const float4*>(in + r * ld_in + c));
    if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) atomicAdd(reinterpret_cast<float4*>(out + tr * ld_out + c), v);
}

__global__ void gather_rows_kernel(const float* __restrict__ in, int64_t ld_in, const int64_t* __restrict__ idx,
                                   int rows_per_idx, int64_t rows, int cols, float* __restrict__ out, int64_t ld_out) {
    pdl_entry();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = t / cols; const int c = (int)(t % cols);
    if (r >= rows) return;
    const int64_t sr = idx[r / rows_per_idx] * rows_per_idx + (r % rows_per_idx);
    out[r * ld_out + c] = in[sr * ld_in + c];
}

// ------------------------------------------------------------------------------------- LayerNorm backward
// y = [relu](LN(x + res) * gamma + beta). One warp per row (grid-stride); D <= 1024.
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma [masked by y > 0]; dgamma += dy' * xhat; dbeta += dy'.
// NPL = columns per lane (template: the 512-wide rows of the path need 16, the 32-wide distance-bias MLP 1; a fixed 32 cost
// 128 accumulator registers per thread and half-empty loops - 71 us per call in the ncu launch list of a training step).
constexpr int LNB_MAX = 32;
template <int NPL>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ x, int64_t ldx,
                     const float* __restrict__ res, int64_t ldr, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ dx, int64_t lddx, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, int64_t M, int D, float eps, int relu) {
    pdl_entry();
    extern __shared__ float sacc[];               // [2 * D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    float dg[NPL], db[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) { dg[i] = 0.f; db[i] = 0.f; }
    for (int64_t row = (int64_t)blockIdx.x * nwarp + warp; row < M; row += (int64_t)gridDim.x * nwarp) {
        float v[NPL], g[NPL];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const int c = lane + 32 * i;
            float t = 0.f;
            if (c < D) { t = x[row * ldx + c]; if (res) t += res[row * ldr + c]; }
            v[i] = t; sum += t;
        }
        const float mean = warp_sum(sum) / (float)D;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NPL; ++i) { const int c = lane + 32 * i; if (c < D) { const float d = v[i] - mean; sq += d * d; } }
        const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const int c = lane + 32 * i;
            g[i] = 0.f;
            if (c < D) {
                const float xh = (v[i] - mean) * rstd;
                const float gm = __ldg(gamma + c);
                float d = dy[row * lddy + c];
                if (relu && !(xh * gm + __ldg(beta + c) > 0.f)) d = 0.f;
                dg[i] += d * xh; db[i] += d;
                g[i] = d * gm;
                v[i] = xh;
                s1 += g[i]; s2 += g[i] * xh;
            }
        }
        s1 = warp_sum(s1) / (float)D; s2 = warp_sum(s2) / (float)D;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const int c = lane + 32 * i;
            if (c < D) dx[row * lddx + c] = rstd * (g[i] - s1 - v[i] * s2);
        }
    }
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const int c = lane + 32 * i;
        if (c < D) { atomicAdd(sacc + c, dg[i]); atomicAdd(sacc + D + c, db[i]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        if (dgamma) atomicAdd(dgamma + i, sacc[i]);
        if (dbeta) atomicAdd(dbeta + i, sacc[D + i]);
    }
}

// ---------------------------------------------------------------- graph attention: softmax + aggregation
// Edges are in CSR (source-sorted) order. t [E*H, d_o] = attention-MLP output for row (e, h); v row (n, h) =
// v + n * ldv + h * d_o (head-major proj_value). One warp per (node, head).
//   p = softmax_c(t[(e,h), :]);  m[e, c*H + h] = p[c] * v[dst(e), h, c]
//   xx[n, c*H + h] = aggr_{e in seg(n)} m   (max: 0 for an empty segment; add; mean)
// Saved for backward: p [E*H, d_o] and, for max, arg [n, c*H + h] = winning edge (-1: empty).
template <int CPL>   // channels per lane: d_o <= 32 * CPL
__global__ void gat_softmax_aggr_fwd_kernel(const float* __restrict__ t, const float* __restrict__ v, int64_t ldv,
                                            const int64_t* __restrict__ dst, const int32_t* __restrict__ row_ptr,
                                            int64_t n_nodes, int H, int d_o, int aggr, float* __restrict__ xx,
                                            int64_t ldxx, float* __restrict__ p_out, int32_t* __restrict__ arg) {
    pdl_entry();
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_nodes * H) return;
    const int64_t n = w / H; const int h = (int)(w % H);
    const int e0 = row_ptr[n], e1 = row_ptr[n + 1];
    float best[CPL]; int barg[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) { best[i] = (aggr == VLSAT_AGGR_MAX) ? -FLT_MAX : 0.f; barg[i] = -1; }
    for (int e = e0; e < e1; ++e) {
        const int64_t r = (int64_t)e * H + h;
        const int64_t d = dst[e];
        float tv[CPL]; float mx = -FLT_MAX;
#pragma unroll
        for (int i = 0; i < CPL; ++i) { const int c = lane + 32 * i; tv[i] = (c < d_o) ? t[r * d_o + c] : -FLT_MAX; mx = fmaxf(mx, tv[i]); }
        mx = warp_max(mx);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) { const int c = lane + 32 * i; tv[i] = (c < d_o) ? expf(tv[i] - mx) : 0.f; s += tv[i]; }
        const float inv = 1.f / warp_sum(s);
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int c = lane + 32 * i;
            if (c < d_o) {
                const float p = tv[i] * inv;
                if (p_out) p_out[r * d_o + c] = p;
                const float m = p * __ldg(v + d * ldv + h * d_o + c);
                if (aggr == VLSAT_AGGR_MAX) { if (m > best[i]) { best[i] = m; barg[i] = e; } }
                else best[i] += m;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        const int c = lane + 32 * i;
        if (c < d_o) {
            float o = best[i];
            if (e1 == e0) o = 0.f;
            else if (aggr == VLSAT_AGGR_MEAN) o /= (float)(e1 - e0);
            xx[n * ldxx + (int64_t)c * H + h] = o;
            if (arg) arg[n * ((int64_t)H * d_o) + (int64_t)c * H + h] = barg[i];
        }
    }
}

// backward: dt[(e,h), c] = p (dp - sum_c p dp), dp = dm * v, dm = routed dxx; dv[dst(e), h, c] += dm * p (atomic).
template <int CPL>
__global__ void gat_softmax_aggr_bwd_kernel(const float* __restrict__ dxx, int64_t lddxx, const float* __restrict__ p,
                                            const float* __restrict__ v, int64_t ldv, const int64_t* __restrict__ dst,
                                            const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ arg,
                                            int64_t n_nodes, int H, int d_o, int aggr, float* __restrict__ dt,
                                            float* __restrict__ dv, int64_t lddv) {
    pdl_entry();
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_nodes * H) return;
    const int64_t n = w / H; const int h = (int)(w % H);
    const int e0 = row_ptr[n], e1 = row_ptr[n + 1];
    if (e1 == e0) return;
    float g[CPL]; int a[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        const int c = lane + 32 * i;
        g[i] = 0.f; a[i] = -1;
        if (c < d_o) {
            g[i] = dxx[n * lddxx + (int64_t)c * H + h];
            if (aggr == VLSAT_AGGR_MEAN) g[i] /= (float)(e1 - e0);
            if (aggr == VLSAT_AGGR_MAX) a[i] = arg[n * ((int64_t)H * d_o) + (int64_t)c * H + h];
        }
    }
    for (int e = e0; e < e1; ++e) {
        const int64_t r = (int64_t)e * H + h;
        const int64_t d = dst[e];
        float pv[CPL], dp[CPL]; float dot = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int c = lane + 32 * i;
            pv[i] = 0.f; dp[i] = 0.f;
            if (c < d_o) {
                pv[i] = p[r * d_o + c];
                const float dm = (aggr == VLSAT_AGGR_MAX) ? (a[i] == e ? g[i] : 0.f) : g[i];
                dp[i] = dm * __ldg(v + d * ldv + h * d_o + c);
                if (dm != 0.f) atomicAdd(dv + d * lddv + h * d_o + c, dm * pv[i]);
                dot += pv[i] * dp[i];
            }
        }
        dot = warp_sum(dot);
#pragma unroll
        for (int i = 0; i < CPL; ++i) { const int c = lane + 32 * i; if (c < d_o) dt[r * d_o + c] = pv[i] * (dp[i] - dot); }
    }
}

// ------------------------------------------------------------- edge cross-attention backward, score stage
// For one head and one block of queries: s = Q K^T (raw, unscaled) and dp = dO V^T are [nq, nk] GEMM outputs.
//   P = exp(scale * s - lse[i]);  dS = P * (dp - delta[i]) * scale
// Written: ds [nq, nk] (may alias dp), ds_t [nk, ld_t] = dS^T and p_t [nk, ld_t] = P^T (columns >= nq zero-filled).
__global__ void attn_prob_bwd_kernel(const float* __restrict__ s, const float* __restrict__ dp, int64_t ld,
                                     const float* __restrict__ lse, const float* __restrict__ delta, float scale,
                                     float* __restrict__ ds, float* __restrict__ ds_t, float* __restrict__ p_t, int64_t ld_t,
                                     int64_t nq, int64_t nk) {
    pdl_entry();
    __shared__ float tp[32][33], td[32][33];
    const int64_t i0 = (int64_t)blockIdx.y * 32, j0 = (int64_t)blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int64_t i = i0 + r, j = j0 + threadIdx.x;
        float pv = 0.f, dv = 0.f;
        if (i < nq && j < nk) {
            pv = expf(scale * s[i * ld + j] - __ldg(lse + i));
            dv = pv * (dp[i * ld + j] - __ldg(delta + i)) * scale;
            ds[i * ld + j] = dv;
        }
        tp[r][threadIdx.x] = pv; td[r][threadIdx.x] = dv;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int64_t j = j0 + r, i = i0 + threadIdx.x;
        if (j < nk && i < ld_t) { p_t[j * ld_t + i] = tp[threadIdx.x][r]; ds_t[j * ld_t + i] = td[threadIdx.x][r]; }
    }
}

// Same stage with every output as the bf16 (hi, lo) pair the three following products read (dV += P^T dO, dK += dS^T Q,
// dQ = dS K): the fp32 [nq, nk] matrices are never written and never read back to be split. ds_* [nq, ld_ds],
// dst_* / pt_* [nk, ld_t] (columns >= nq zero-filled).
__device__ __forceinline__ void bwd_split16(float v, uint16_t& hi, uint16_t& lo) {
    const uint32_t h = (__float_as_uint(v) + 0x8000u) & 0xffff0000u;
    hi = (uint16_t)(h >> 16);
    lo = (uint16_t)((__float_as_uint(v - __uint_as_float(h)) + 0x8000u) >> 16);
}
__global__ void attn_prob_bwd_pairs_kernel(const float* __restrict__ s, const float* __restrict__ dp, int64_t ld,
                                           const float* __restrict__ lse, const float* __restrict__ delta, float scale,
                                           uint16_t* __restrict__ ds_hi, uint16_t* __restrict__ ds_lo, int64_t ld_ds,
                                           uint16_t* __restrict__ dst_hi, uint16_t* __restrict__ dst_lo,
                                           uint16_t* __restrict__ pt_hi, uint16_t* __restrict__ pt_lo, int64_t ld_t,
                                           int64_t nq, int64_t nk) {
    pdl_entry();
    __shared__ float tp[32][33], td[32][33];
    const int64_t i0 = (int64_t)blockIdx.y * 32, j0 = (int64_t)blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int64_t i = i0 + r, j = j0 + threadIdx.x;
        float pv = 0.f, dv = 0.f;
        if (i < nq && j < nk) {
            pv = expf(scale * s[i * ld + j] - __ldg(lse + i));
            dv = pv * (dp[i * ld + j] - __ldg(delta + i)) * scale;
            bwd_split16(dv, ds_hi[i * ld_ds + j], ds_lo[i * ld_ds + j]);
        }
        tp[r][threadIdx.x] = pv; td[r][threadIdx.x] = dv;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int64_t j = j0 + r, i = i0 + threadIdx.x;
        if (j < nk && i < ld_t) {
            bwd_split16(tp[threadIdx.x][r], pt_hi[j * ld_t + i], pt_lo[j * ld_t + i]);
            bwd_split16(td[threadIdx.x][r], dst_hi[j * ld_t + i], dst_lo[j * ld_t + i]);
        }
    }
}

// delta[h, i] = sum_d a[i, h*dk + d] * b[i, h*dk + d]; one warp per (i, h)
__global__ void rowdot_heads_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb,
                                    float* __restrict__ out, int64_t M, int H, int dk) {
    pdl_entry();
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= M * H) return;
    const int64_t i = w / H; const int h = (int)(w % H);
    float acc = 0.f;
    for (int d = lane; d < dk; d += 32) acc += a[i * lda + h * dk + d] * b[i * ldb + h * dk + d];
    acc = warp_sum(acc);
    if (lane == 0) out[(int64_t)h * M + i] = acc;
}

// ------------------------------------------------------------------------------- PointNet max-pool backward
// dz3 [n_obj, C3] = dOut masked by out > 0; arg [n_obj, C3] = winning point; h2 [n_obj * P, C2] = post-ReLU layer-2
// activations (recomputed by two projections); w3 [C3, C2].
//   dW3[c, k] += dz3[o, c] * h2[(o, arg[o,c]), k];   dh2[(o, arg[o,c]), k] += dz3[o, c] * w3[c, k]
// Block = C2 threads (k); blockIdx.x = group of PB_CH channels; blockIdx.y strides over objects.
constexpr int PB_CH = 8;
__global__ void pointnet_pool_bwd_kernel(const float* __restrict__ dz3, const int32_t* __restrict__ arg,
                                         const float* __restrict__ h2, const float* __restrict__ w3, int64_t n_obj,
                                         int64_t n_pts, int C3, int C2, float* __restrict__ dw3, float* __restrict__ dh2) {
    pdl_entry();
    const int k = threadIdx.x;
    const int c0 = blockIdx.x * PB_CH;
    float acc[PB_CH], wv[PB_CH];
#pragma unroll
    for (int u = 0; u < PB_CH; ++u) { acc[u] = 0.f; wv[u] = (c0 + u < C3) ? __ldg(w3 + (int64_t)(c0 + u) * C2 + k) : 0.f; }
    for (int64_t o = blockIdx.y; o < n_obj; o += gridDim.y) {
#pragma unroll
        for (int u = 0; u < PB_CH; ++u) {
            const int c = c0 + u;
            if (c >= C3) break;
            const float g = __ldg(dz3 + o * C3 + c);
            if (g == 0.f) continue;                                   // block-uniform
            const int64_t row = o * n_pts + __ldg(arg + o * C3 + c);
            acc[u] = fmaf(g, __ldg(h2 + row * C2 + k), acc[u]);
            atomicAdd(dh2 + row * C2 + k, g * wv[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < PB_CH; ++u)
        if (c0 + u < C3 && acc[u] != 0.f) atomicAdd(dw3 + (int64_t)(c0 + u) * C2 + k, acc[u]);
}

// ------------------------------------------------------------------------------------------------ dropout
// Counter-based mask: element i of call (seed, offset) is kept iff hash(seed, offset + i) >= p * 2^32.
// y = keep ? x / (1 - p) : 0. Backward applies the same call to the gradient.
__device__ __forceinline__ uint32_t mix32(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 16);
}
__global__ void dropout_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ y, int64_t ldy, int64_t rows,
                               int64_t cols, uint32_t thresh, float inv_keep, uint64_t seed, uint64_t offset,
                               const uint64_t* __restrict__ device_step) {
    pdl_entry();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i / cols, c = i % cols;
    // device_step: a counter in device memory that a replayed CUDA graph bumps, so that replays draw fresh masks
    const uint64_t step = device_step ? *device_step : 0ull;
    const bool keep = mix32((seed + step * 0x9E3779B97F4A7C15ull) * 0xD6E8FEB86659FD93ull + offset + (uint64_t)i) >= thresh;
    y[r * ldy + c] = keep ? x[r * ldx + c] * inv_keep : 0.f;
}

// Four columns per thread (cols % 4 == 0, 16-byte aligned rows): the same per-element mask as the scalar kernel (the counter
// of element (r, c) is r * cols + c), without its 64-bit division per element.
__global__ void dropout_vec_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ y, int64_t ldy, int64_t rows,
                                   int cols4, uint32_t thresh, float inv_keep, uint64_t seed, uint64_t offset,
                                   const uint64_t* __restrict__ device_step, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                   int64_t ld_pair) {
    pdl_entry();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * cols4) return;
    const int64_t r = t / cols4;
    const int c = (int)(t - r * cols4) * 4;
    const uint64_t step = device_step ? *device_step : 0ull;
    const uint64_t base = (seed + step * 0x9E3779B97F4A7C15ull) * 0xD6E8FEB86659FD93ull + offset + (uint64_t)(r * (int64_t)cols4 * 4 + c);
    float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ldx + c));
    v.x = mix32(base) >= thresh ? v.x * inv_keep : 0.f;
    v.y = mix32(base + 1) >= thresh ? v.y * inv_keep : 0.f;
    v.z = mix32(base + 2) >= thresh ? v.z * inv_keep : 0.f;
    v.w = mix32(base + 3) >= thresh ? v.w * inv_keep : 0.f;
    *reinterpret_cast<float4*>(y + r * ldy + c) = v;
    if (hi) {                                                // the consumer is a projection: its bf16 (hi, lo) operand pair from this pass
        uint32_t h0, l0, h1, l1;
        split_bf16x2(v.x, v.y, h0, l0); split_bf16x2(v.z, v.w, h1, l1);
        *reinterpret_cast<uint2*>(hi + r * ld_pair + c) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(lo + r * ld_pair + c) = make_uint2(l0, l1);
    }
}

// ------------------------------------------------------------------- BatchNorm1d with batch statistics
// x [M, N]; block = 32 columns x 8 row-lanes. stats: mean[n], rstd[n] (biased variance, as F.batch_norm normalises);
// running stats updated with the unbiased variance (momentum), as nn.BatchNorm1d does in training mode.
__global__ void bn_stats_kernel(const float* __restrict__ x, int64_t ldx, int64_t M, int64_t N, float eps, float* __restrict__ mean,
                                float* __restrict__ rstd, float* __restrict__ running_mean, float* __restrict__ running_var,
                                float momentum) {
    pdl_entry();
    __shared__ float part[8][33];
    const int64_t n = (int64_t)blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (n < N) for (int64_t m = threadIdx.y; m < M; m += 8) s += x[m * ldx + n];
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    float mu = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) mu += part[i][threadIdx.x];
    mu /= (float)M;
    __syncthreads();
    float q = 0.f;
    if (n < N) for (int64_t m = threadIdx.y; m < M; m += 8) { const float d = x[m * ldx + n] - mu; q += d * d; }
    part[threadIdx.y][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.y == 0 && n < N) {
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) var += part[i][threadIdx.x];
        mean[n] = mu;
        rstd[n] = rsqrtf(var / (float)M + eps);
        if (running_mean) {
            running_mean[n] = (1.f - momentum) * running_mean[n] + momentum * mu;
            running_var[n] = (1.f - momentum) * running_var[n] + momentum * (M > 1 ? var / (float)(M - 1) : var);
        }
    }
}
// y = (x - mean) * rstd * gamma + beta, optional ReLU
__global__ void bn_apply_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ mean, const float* __restrict__ rstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int relu, float* __restrict__ y,
                                int64_t ldy, int64_t M, int64_t N) {
    pdl_entry();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * N) return;
    const int64_t m = i / N, n = i % N;
    float t = (x[m * ldx + n] - __ldg(mean + n)) * __ldg(rstd + n) * __ldg(gamma + n) + __ldg(beta + n);
    if (relu) t = fmaxf(t, 0.f);
    y[m * ldy + n] = t;
}
// backward: g = dy [masked by y > 0]; dgamma = sum g xhat; dbeta = sum g;
//   batch stats: dx = gamma rstd (g - dbeta/M - xhat dgamma/M);  running stats (eval): dx = gamma rstd g
__global__ void bn_bwd_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ x, int64_t ldx,
                              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                              const float* __restrict__ beta, int relu, int batch_stats, float* __restrict__ dx, int64_t lddx,
                              float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t M, int64_t N) {
    pdl_entry();
    __shared__ float p1[8][33], p2[8][33];
    const int64_t n = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const bool ok = n < N;
    const float mu = ok ? mean[n] : 0.f, rs = ok ? rstd[n] : 0.f, gm = ok ? gamma[n] : 0.f, bt = ok ? beta[n] : 0.f;
    float s1 = 0.f, s2 = 0.f;
    if (ok) for (int64_t m = threadIdx.y; m < M; m += 8) {
        const float xh = (x[m * ldx + n] - mu) * rs;
        float g = dy[m * lddy + n];
        if (relu && !(xh * gm + bt > 0.f)) g = 0.f;
        s1 += g; s2 += g * xh;
    }
    p1[threadIdx.y][threadIdx.x] = s1; p2[threadIdx.y][threadIdx.x] = s2;
    __syncthreads();
    s1 = 0.f; s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1 += p1[i][threadIdx.x]; s2 += p2[i][threadIdx.x]; }
    if (ok && threadIdx.y == 0) { dbeta[n] = s1; dgamma[n] = s2; }
    if (ok && dx) for (int64_t m = threadIdx.y; m < M; m += 8) {
        const float xh = (x[m * ldx + n] - mu) * rs;
        float g = dy[m * lddy + n];
        if (relu && !(xh * gm + bt > 0.f)) g = 0.f;
        dx[m * lddx + n] = batch_stats ? gm * rs * (g - s1 / (float)M - xh * s2 / (float)M) : gm * rs * g;
    }
}

// ------------------------------------------------------------------------------ row L2 normalisation backward
// y = x / |x|:  dx = (dy - y (y . dy)) / |x|
__global__ void row_l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, int64_t M, int D) {
    pdl_entry();
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    float sq = 0.f, dot = 0.f;
    for (int c = lane; c < D; c += 32) { const float t = x[row * D + c]; sq += t * t; dot += t * dy[row * D + c]; }
    sq = warp_sum(sq); dot = warp_sum(dot);
    const float inv = 1.f / sqrtf(sq);
    const float k = dot * inv * inv * inv;            // (x . dy) / |x|^3
    for (int c = lane; c < D; c += 32) dx[row * D + c] = dy[row * D + c] * inv - x[row * D + c] * k;
}

// out[0] += sum_i a[i] * b[i]   (gradient of the logit scale: d/ds (e^s z) . dy = y . dy)
__global__ void dot_accum_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ out) {
    pdl_entry();
    __shared__ float part[8];
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += a[i] * b[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { float s = 0.f; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += part[i]; atomicAdd(out, s); }
}

// pair features of the distance-bias MLP (network_MMG.py:189-196): for query a and key b of the same scene,
// row pair_off[a] + (b - seg_start[a]) = [c_b - c_a, |c_b - c_a|]
__global__ void pair_features_kernel(const float* __restrict__ centres, int64_t ldc, const int32_t* __restrict__ seg_start,
                                     const int32_t* __restrict__ seg_end, const int64_t* __restrict__ pair_off, int64_t n_nodes,
                                     float* __restrict__ out) {
    pdl_entry();
    const int64_t a = blockIdx.x;
    if (a >= n_nodes) return;
    const int s0 = seg_start[a], s1 = seg_end[a];
    const float ax = centres[a * ldc], ay = centres[a * ldc + 1], az = centres[a * ldc + 2];
    for (int b = s0 + threadIdx.x; b < s1; b += blockDim.x) {
        const float dx = centres[(int64_t)b * ldc] - ax, dy = centres[(int64_t)b * ldc + 1] - ay, dz = centres[(int64_t)b * ldc + 2] - az;
        float4 o = make_float4(dx, dy, dz, sqrtf(dx * dx + dy * dy + dz * dz));
        reinterpret_cast<float4*>(out)[pair_off[a] + (b - s0)] = o;
    }
}

// ---------------------------------------------------------- weight gradient of a small projection (N, K <= 128)
// dW[n, k] += sum_m dz[m, n] * x[m, k] for a tall operand pair (M = edges x heads or points): every CTA reduces a slab
// of rows in registers (16 x 16 threads, 8 x 8 outputs each, exact FP32 FMA) and adds its partial sum atomically.
// The tensor-core GEMM would run such a [N, K] result as one tile on one SM.
constexpr int WG_ROWS = 32;
__global__ void __launch_bounds__(256)
wgrad_small_kernel(const float* __restrict__ dz, int64_t lddz, const float* __restrict__ x, int64_t ldx, int64_t M, int N, int K,
                   int64_t rows_per_cta, float* __restrict__ dw, int64_t lddw) {
    pdl_entry();
    __shared__ float sz[WG_ROWS][128], sx[WG_ROWS][128];
    const int tn = threadIdx.x >> 4, tk = threadIdx.x & 15;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int64_t m_begin = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t m_end = (m_begin + rows_per_cta < M) ? m_begin + rows_per_cta : M;
    for (int64_t m0 = m_begin; m0 < m_end; m0 += WG_ROWS) {
        for (int i = threadIdx.x; i < WG_ROWS * 128; i += 256) {
            const int r = i >> 7, c = i & 127;
            const int64_t m = m0 + r;
            sz[r][c] = (m < m_end && c < N) ? dz[m * lddz + c] : 0.f;
            sx[r][c] = (m < m_end && c < K) ? x[m * ldx + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < WG_ROWS; ++r) {
            float a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = sz[r][tn + 16 * i]; b[i] = sx[r][tk + 16 * i]; }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = tn + 16 * i, k = tk + 16 * j;
            if (n < N && k < K && acc[i][j] != 0.f) atomicAdd(dw + (int64_t)n * lddw + k, acc[i][j]);
        }
}

}  // namespace vlsat

using namespace vlsat;

extern "C" int vlsat_wgrad_small(const float* dz, int64_t lddz, const float* x, int64_t ldx, int64_t M, int N, int K, float* dw,
                                 int64_t lddw, void* stream) {
    VLSAT_REQUIRE(M >= 0 && N >= 1 && K >= 1);
    if (M == 0) return VLSAT_OK;
    VLSAT_REQUIRE(dz && x && dw && lddz >= N && ldx >= K && lddw >= K);
    VLSAT_SUPPORT(N <= 128 && K <= 128);
    int64_t rows_per_cta = ceil_div(ceil_div(M, 2 * kNumSMs), WG_ROWS) * WG_ROWS;
    if (rows_per_cta < WG_ROWS) rows_per_cta = WG_ROWS;
    launch_k(wgrad_small_kernel, dim3((unsigned)ceil_div(M, rows_per_cta)), dim3(256), 0, (cudaStream_t)stream, dz, lddz, x, ldx, M, N, K, rows_per_cta, dw, lddw);
    return finish_launch();
}

extern "C" int vlsat_transpose(const float* in, int64_t ld_in, int64_t batch_stride_in, float* out, int64_t ld_out,
                               int64_t batch_stride_out, int64_t batch, int64_t rows, int64_t cols, void* stream) {
    VLSAT_REQUIRE(batch >= 0 && rows >= 0 && cols >= 0 && ld_in >= cols && ld_out >= rows);
    if (batch == 0 || cols == 0 || ld_out == 0) return VLSAT_OK;
    VLSAT_REQUIRE(in && out);
    VLSAT_SUPPORT(batch <= 65535 && ceil_div(cols, 32) <= 65535);
    dim3 grid((unsigned)ceil_div(ld_out, 32), (unsigned)ceil_div(cols, 32), (unsigned)batch);
    launch_k(transpose_kernel, grid, dim3(32, 8), 0, (cudaStream_t)stream, in, ld_in, batch_stride_in, out, ld_out, batch_stride_out, rows, cols);
    return finish_launch();
}

extern "C" int vlsat_act_bwd(const float* dy, int64_t lddy, const float* y, int64_t ldy, int act, float scale,
                             const float* scale_ptr, float* dz, int64_t lddz, float* dbias, int64_t M, int64_t N, void* stream) {
    VLSAT_REQUIRE(M >= 0 && N >= 0 && act >= VLSAT_ACT_NONE && act <= VLSAT_ACT_SIGMOID);
    if (M == 0 || N == 0) return VLSAT_OK;
    VLSAT_REQUIRE(dy && lddy >= N && (dz || dbias) && (!dz || lddz >= N) && (act == VLSAT_ACT_NONE || (y && ldy >= N)));
    VLSAT_SUPPORT(ceil_div(M, AB_ROWS) <= 65535);
    dim3 grid((unsigned)ceil_div(N, 32), (unsigned)ceil_div(M, AB_ROWS));
    launch_k(act_bwd_kernel, grid, dim3(32, 8), 0, (cudaStream_t)stream, dy, lddy, y, ldy, act, scale, scale_ptr, dz, lddz, dbias, M, N);
    return finish_launch();
}

extern "C" int vlsat_act_bwd_pair(const float* dy, int64_t lddy, const float* y, int64_t ldy, int act, float scale,
                                  const float* scale_ptr, float* dz, int64_t lddz, float* dbias, void* split_hi, void* split_lo,
                                  int64_t ld_split, int64_t M, int64_t N, void* stream) {
    VLSAT_REQUIRE(M >= 0 && N >= 0 && act >= VLSAT_ACT_NONE && act <= VLSAT_ACT_SIGMOID);
    if (M == 0 || N == 0) return VLSAT_OK;
    VLSAT_REQUIRE(dy && lddy >= N && (dz || dbias || split_hi) && (!dz || lddz >= N) && (act == VLSAT_ACT_NONE || (y && ldy >= N)));
    VLSAT_REQUIRE((split_hi == nullptr) == (split_lo == nullptr) && (!split_hi || ld_split >= N));
    VLSAT_SUPPORT(ceil_div(M, AB_ROWS) <= 65535);
    const uintptr_t al = (uintptr_t)dy | (uintptr_t)y | (uintptr_t)dz;
    VLSAT_SUPPORT(N % 4 == 0 && lddy % 4 == 0 && (act == VLSAT_ACT_NONE || ldy % 4 == 0) && (!dz || lddz % 4 == 0) && (al & 15) == 0);
    VLSAT_SUPPORT(!split_hi || (ld_split % 4 == 0 && (((uintptr_t)split_hi | (uintptr_t)split_lo) & 7) == 0));
    dim3 grid((unsigned)ceil_div(N, 128), (unsigned)ceil_div(M, AB_ROWS));
    launch_k(act_bwd_vec_kernel, grid, dim3(32, 8), 0, (cudaStream_t)stream, dy, lddy, y, ldy, act, scale, scale_ptr, dz, lddz, dbias,
             (uint16_t*)split_hi, (uint16_t*)split_lo, ld_split, M, N);
    return finish_launch();
}

extern "C" int vlsat_scatter_add_rows(const float* in, int64_t ld_in, const int64_t* idx, int rows_per_idx, int64_t rows,
                                      int cols, float* out, int64_t ld_out, void* stream) {
    VLSAT_REQUIRE(rows >= 0 && cols >= 0 && rows_per_idx >= 1);
    if (rows == 0 || cols == 0) return VLSAT_OK;
    VLSAT_REQUIRE(in && idx && out && ld_in >= cols && ld_out >= cols);
    if (cols % 4 == 0 && ld_in % 4 == 0 && ld_out % 4 == 0 && ((((uintptr_t)in | (uintptr_t)out) & 15) == 0))
        launch_k(scatter_add_rows_vec_kernel, dim3((unsigned)ceil_div(rows * (cols / 4), 256)), dim3(256), 0, (cudaStream_t)stream, in, ld_in, idx,
                 rows_per_idx, rows, cols / 4, out, ld_out);
    else
        launch_k(scatter_add_rows_kernel, dim3((unsigned)ceil_div(rows * cols, 256)), dim3(256), 0, (cudaStream_t)stream, in, ld_in, idx, rows_per_idx, rows, cols, out, ld_out);
    return finish_launch();
}

extern "C" int vlsat_gather_rows(const float* in, int64_t ld_in, const int64_t* idx, int rows_per_idx, int64_t rows,
                                 int cols, float* out, int64_t ld_out, void* stream) {
    VLSAT_REQUIRE(rows >= 0 && cols >= 0 && rows_per_idx >= 1);
    if (rows == 0 || cols == 0) return VLSAT_OK;
    VLSAT_REQUIRE(in && idx && out && ld_in >= cols && ld_out >= cols);
    launch_k(gather_rows_kernel, dim3((unsigned)ceil_div(rows * cols, 256)), dim3(256), 0, (cudaStream_t)stream, in, ld_in, idx, rows_per_idx, rows, cols, out, ld_out);
    return finish_launch();
}

extern "C" int vlsat_add_layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* res, int64_t ld_res,
                                       const float* gamma, const float* beta, float* dx, int64_t lddx, float* dgamma, float* dbeta,
                                       int64_t M, int D, float eps, int relu, void* stream) {
    VLSAT_REQUIRE(M >= 0 && D >= 1);
    if (M == 0) return VLSAT_OK;
    VLSAT_REQUIRE(dy && x && gamma && beta && dx && lddy >= D && ldx >= D && lddx >= D && (!res || ld_res >= D));
    VLSAT_SUPPORT(D <= 32 * LNB_MAX);
    const unsigned grid = (unsigned)(ceil_div(M, 8) < 4 * kNumSMs ? ceil_div(M, 8) : 4 * kNumSMs);
#define LNB_LAUNCH(NPL_) launch_k(layernorm_bwd_kernel<NPL_>, grid, dim3(256), 2 * D * sizeof(float), (cudaStream_t)stream, dy, lddy, x, ldx, res, \
                                  ld_res, gamma, beta, dx, lddx, dgamma, dbeta, M, D, eps, relu)
    if (D <= 32) LNB_LAUNCH(1); else if (D <= 128) LNB_LAUNCH(4); else if (D <= 512) LNB_LAUNCH(16); else LNB_LAUNCH(32);
#undef LNB_LAUNCH
    return finish_launch();
}

extern "C" int vlsat_gat_softmax_aggr_fwd(const float* t, const float* v, int64_t ldv, const int64_t* dst_sorted,
                                          const int32_t* row_ptr, int64_t n_nodes, int64_t n_edges, int n_heads, int d_o,
                                          int aggr, float* xx, int64_t ld_xx, float* prob, int32_t* argmax, void* stream) {
    VLSAT_REQUIRE(n_nodes >= 0 && n_edges >= 0 && n_heads >= 1 && d_o >= 1 && aggr >= VLSAT_AGGR_MAX && aggr <= VLSAT_AGGR_MEAN);
    if (n_nodes == 0) return VLSAT_OK;
    VLSAT_REQUIRE(v && row_ptr && xx && ld_xx >= (int64_t)n_heads * d_o && (n_edges == 0 || (t && dst_sorted)));
    VLSAT_SUPPORT(d_o <= 128);
    const unsigned grid = (unsigned)ceil_div(n_nodes * n_heads * 32, 256);
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(C_) launch_k(gat_softmax_aggr_fwd_kernel<C_>, grid, dim3(256), 0, st, t, v, ldv, dst_sorted, row_ptr, n_nodes, n_heads, d_o, aggr, xx, ld_xx, prob, argmax)
    if (d_o <= 32) LAUNCH(1); else if (d_o <= 64) LAUNCH(2); else LAUNCH(4);
#undef LAUNCH
    return finish_launch();
}

extern "C" int vlsat_gat_softmax_aggr_bwd(const float* dxx, int64_t ld_dxx, const float* prob, const float* v, int64_t ldv,
                                          const int64_t* dst_sorted, const int32_t* row_ptr, const int32_t* argmax,
                                          int64_t n_nodes, int64_t n_edges, int n_heads, int d_o, int aggr, float* dt,
                                          float* dv, int64_t ld_dv, void* stream) {
    VLSAT_REQUIRE(n_nodes >= 0 && n_edges >= 0 && n_heads >= 1 && d_o >= 1 && aggr >= VLSAT_AGGR_MAX && aggr <= VLSAT_AGGR_MEAN);
    if (n_nodes == 0 || n_edges == 0) return VLSAT_OK;
    VLSAT_REQUIRE(dxx && prob && v && dst_sorted && row_ptr && dt && dv && (aggr != VLSAT_AGGR_MAX || argmax));
    VLSAT_SUPPORT(d_o <= 128);
    const unsigned grid = (unsigned)ceil_div(n_nodes * n_heads * 32, 256);
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(C_) launch_k(gat_softmax_aggr_bwd_kernel<C_>, grid, dim3(256), 0, st, dxx, ld_dxx, prob, v, ldv, dst_sorted, row_ptr, argmax, n_nodes, n_heads, d_o, aggr, dt, dv, ld_dv)
    if (d_o <= 32) LAUNCH(1); else if (d_o <= 64) LAUNCH(2); else LAUNCH(4);
#undef LAUNCH
    return finish_launch();
}

extern "C" int vlsat_attn_prob_bwd(const float* s, const float* dp, int64_t ld, const float* lse, const float* delta, float scale,
                                   float* ds, float* ds_t, float* p_t, int64_t ld_t, int64_t nq, int64_t nk, void* stream) {
    VLSAT_REQUIRE(nq >= 0 && nk >= 0);
    if (nq == 0 || nk == 0) return VLSAT_OK;
    VLSAT_REQUIRE(s && dp && lse && delta && ds && ds_t && p_t && ld >= nk && ld_t >= nq);
    VLSAT_SUPPORT(ceil_div(ld_t, 32) <= 65535);
    dim3 grid((unsigned)ceil_div(nk, 32), (unsigned)ceil_div(ld_t, 32));
    launch_k(attn_prob_bwd_kernel, grid, dim3(32, 8), 0, (cudaStream_t)stream, s, dp, ld, lse, delta, scale, ds, ds_t, p_t, ld_t, nq, nk);
    return finish_launch();
}

extern "C" int vlsat_attn_prob_bwd_pairs(const float* s, const float* dp, int64_t ld, const float* lse, const float* delta,
                                         float scale, void* ds_hi, void* ds_lo, int64_t ld_ds, void* dst_hi, void* dst_lo,
                                         void* pt_hi, void* pt_lo, int64_t ld_t, int64_t nq, int64_t nk, void* stream) {
    VLSAT_REQUIRE(nq >= 0 && nk >= 0);
    if (nq == 0 || nk == 0) return VLSAT_OK;
    VLSAT_REQUIRE(s && dp && lse && delta && ds_hi && ds_lo && dst_hi && dst_lo && pt_hi && pt_lo && ld >= nk && ld_ds >= nk && ld_t >= nq);
    VLSAT_SUPPORT(ceil_div(ld_t, 32) <= 65535);
    dim3 grid((unsigned)ceil_div(nk, 32), (unsigned)ceil_div(ld_t, 32));
    launch_k(attn_prob_bwd_pairs_kernel, grid, dim3(32, 8), 0, (cudaStream_t)stream, s, dp, ld, lse, delta, scale, (uint16_t*)ds_hi, (uint16_t*)ds_lo,
                                                                               ld_ds, (uint16_t*)dst_hi, (uint16_t*)dst_lo, (uint16_t*)pt_hi,
                                                                               (uint16_t*)pt_lo, ld_t, nq, nk);
    return finish_launch();
}

extern "C" int vlsat_rowdot_heads(const float* a, int64_t lda, const float* b, int64_t ldb, float* out, int64_t M, int n_heads,
                                  int dk, void* stream) {
    VLSAT_REQUIRE(M >= 0 && n_heads >= 1 && dk >= 1);
    if (M == 0) return VLSAT_OK;
    VLSAT_REQUIRE(a && b && out && lda >= (int64_t)n_heads * dk && ldb >= (int64_t)n_heads * dk);
    launch_k(rowdot_heads_kernel, dim3((unsigned)ceil_div(M * n_heads * 32, 256)), dim3(256), 0, (cudaStream_t)stream, a, lda, b, ldb, out, M, n_heads, dk);
    return finish_launch();
}

extern "C" int vlsat_pointnet_pool_bwd(const float* dz3, const int32_t* argmax, const float* h2, const float* w3, int64_t n_obj,
                                       int64_t n_pts, int c_out, int c2, float* dw3, float* dh2, void* stream) {
    VLSAT_REQUIRE(n_obj >= 0 && n_pts >= 1 && c_out >= 1 && c2 >= 1);
    if (n_obj == 0) return VLSAT_OK;
    VLSAT_REQUIRE(dz3 && argmax && h2 && w3 && dw3 && dh2);
    VLSAT_SUPPORT(c2 <= 1024);
    const unsigned gy = (unsigned)(n_obj < 32 ? n_obj : 32);
    dim3 grid((unsigned)ceil_div(c_out, PB_CH), gy);
    launch_k(pointnet_pool_bwd_kernel, grid, dim3(c2), 0, (cudaStream_t)stream, dz3, argmax, h2, w3, n_obj, n_pts, c_out, c2, dw3, dh2);
    return finish_launch();
}

static int dropout_launch(const float* x, int64_t ldx, float* y, int64_t ldy, int64_t rows, int64_t cols, float p, uint64_t seed,
                          uint64_t offset, const uint64_t* device_step, uint16_t* hi, uint16_t* lo, int64_t ld_pair, void* stream) {
    const double th = (double)p * 4294967296.0;
    const uint32_t thresh = th >= 4294967295.0 ? 4294967295u : (uint32_t)th;
    const bool vec = cols % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && ((((uintptr_t)x | (uintptr_t)y) & 15) == 0) && cols / 4 < (1ll << 31);
    if (hi && !(vec && ld_pair % 4 == 0 && ((((uintptr_t)hi | (uintptr_t)lo) & 7) == 0))) return VLSAT_ERR_UNSUPPORTED;
    if (vec)
        launch_k(dropout_vec_kernel, dim3((unsigned)ceil_div(rows * (cols / 4), 256)), dim3(256), 0, (cudaStream_t)stream, x, ldx, y, ldy, rows,
                 (int)(cols / 4), thresh, 1.f / (1.f - p), seed, offset, device_step, hi, lo, ld_pair);
    else
        launch_k(dropout_kernel, dim3((unsigned)ceil_div(rows * cols, 256)), dim3(256), 0, (cudaStream_t)stream, x, ldx, y, ldy, rows, cols, thresh, 1.f / (1.f - p), seed, offset, device_step);
    return finish_launch();
}

extern "C" int vlsat_dropout(const float* x, int64_t ldx, float* y, int64_t ldy, int64_t rows, int64_t cols, float p,
                             uint64_t seed, uint64_t offset, const uint64_t* device_step, void* stream) {
    VLSAT_REQUIRE(rows >= 0 && cols >= 0 && p >= 0.f && p < 1.f);
    if (rows == 0 || cols == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && y && ldx >= cols && ldy >= cols);
    return dropout_launch(x, ldx, y, ldy, rows, cols, p, seed, offset, device_step, nullptr, nullptr, 0, stream);
}

extern "C" int vlsat_dropout_pair(const float* x, int64_t ldx, float* y, int64_t ldy, int64_t rows, int64_t cols, float p,
                                  uint64_t seed, uint64_t offset, const uint64_t* device_step, void* pair_hi, void* pair_lo,
                                  int64_t ld_pair, void* stream) {
    VLSAT_REQUIRE(rows >= 0 && cols >= 0 && p >= 0.f && p < 1.f);
    if (rows == 0 || cols == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && y && pair_hi && pair_lo && ldx >= cols && ldy >= cols && ld_pair >= cols);
    return dropout_launch(x, ldx, y, ldy, rows, cols, p, seed, offset, device_step, (uint16_t*)pair_hi, (uint16_t*)pair_lo, ld_pair, stream);
}

extern "C" int vlsat_batchnorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float* mean, float* rstd,
                                   float* running_mean, float* running_var, float momentum, float eps, int batch_stats, int relu,
                                   float* y, int64_t ldy, int64_t M, int64_t N, void* stream) {
    VLSAT_REQUIRE(M >= 0 && N >= 1);
    if (M == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && gamma && beta && mean && rstd && y && ldx >= N && ldy >= N);
    cudaStream_t st = (cudaStream_t)stream;
    int launches = 1;
    if (batch_stats) {
        launch_k(bn_stats_kernel, dim3((unsigned)ceil_div(N, 32)), dim3(32, 8), 0, st, x, ldx, M, N, eps, mean, rstd, running_mean, running_var, momentum);
        ++launches;
    }
    launch_k(bn_apply_kernel, dim3((unsigned)ceil_div(M * N, 256)), dim3(256), 0, st, x, ldx, mean, rstd, gamma, beta, relu, y, ldy, M, N);
    return finish_launch(launches);
}

extern "C" int vlsat_batchnorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd,
                                   const float* gamma, const float* beta, int relu, int batch_stats, float* dx, int64_t lddx,
                                   float* dgamma, float* dbeta, int64_t M, int64_t N, void* stream) {
    VLSAT_REQUIRE(M >= 0 && N >= 1);
    if (M == 0) return VLSAT_OK;
    VLSAT_REQUIRE(dy && x && mean && rstd && gamma && beta && dgamma && dbeta && lddy >= N && ldx >= N && (!dx || lddx >= N));
    launch_k(bn_bwd_kernel, dim3((unsigned)ceil_div(N, 32)), dim3(32, 8), 0, (cudaStream_t)stream, dy, lddy, x, ldx, mean, rstd, gamma, beta, relu, batch_stats,
                                                                                       dx, lddx, dgamma, dbeta, M, N);
    return finish_launch();
}

extern "C" int vlsat_row_l2norm_bwd(const float* dy, const float* x, float* dx, int64_t M, int D, void* stream) {
    VLSAT_REQUIRE(M >= 0 && D >= 1);
    if (M == 0) return VLSAT_OK;
    VLSAT_REQUIRE(dy && x && dx);
    launch_k(row_l2norm_bwd_kernel, dim3((unsigned)ceil_div(M * 32, 256)), dim3(256), 0, (cudaStream_t)stream, dy, x, dx, M, D);
    return finish_launch();
}

extern "C" int vlsat_dot_accum(const float* a, const float* b, int64_t n, float* out, void* stream) {
    VLSAT_REQUIRE(n >= 0 && out);
    if (n == 0) return VLSAT_OK;
    VLSAT_REQUIRE(a && b);
    const unsigned grid = (unsigned)(ceil_div(n, 256) < 4 * kNumSMs ? ceil_div(n, 256) : 4 * kNumSMs);
    launch_k(dot_accum_kernel, grid, dim3(256), 0, (cudaStream_t)stream, a, b, n, out);
    return finish_launch();
}

extern "C" int vlsat_pair_features(const float* centres, int64_t ld_centres, const int32_t* seg_start, const int32_t* seg_end,
                                   const int64_t* pair_off, int64_t n_nodes, float* out, void* stream) {
    VLSAT_REQUIRE(n_nodes >= 0);
    if (n_nodes == 0) return VLSAT_OK;
    VLSAT_REQUIRE(centres && seg_start && seg_end && pair_off && out && ld_centres >= 3);
    VLSAT_SUPPORT((uintptr_t)out % 16 == 0);
    launch_k(pair_features_kernel, dim3((unsigned)n_nodes), dim3(64), 0, (cudaStream_t)stream, centres, ld_centres, seg_start, seg_end, pair_off, n_nodes, out);
    return finish_launch();
}
