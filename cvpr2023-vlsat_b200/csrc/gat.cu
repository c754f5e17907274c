// A8 core: CSR construction and the fused graph-attention edge kernel.
//
// Reference: Gen_Index gathers x[src], x[dst] into [E,512] tensors, five per-edge GEMMs / convs run on
// them, and torch_scatter aggregates with atomics (network_MMG.py:34-41,96-104; network_util.py:50-73).
// Here the node-side projections q = proj_query(x), v = proj_value(x) are computed once per NODE, the
// edges are grouped by source node (CSR), and one kernel does
//   gather q[src], k[e], v[dst] -> per-(edge, head) MLP -> softmax over d_o -> * value -> aggregate
// with the aggregate kept in registers per source node: no atomics, no [E, D_a] message tensor, every
// node's q row read once, deterministic for max / add / mean alike.
#include "common.cuh"
#include <float.h>
#include <algorithm>

namespace vlsat {

// ---------------------------------------------------------------------------------------- CSR build
__global__ void csr_zero_kernel(int32_t* counts, int64_t n) {
    pdl_entry();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) counts[i] = 0;
}
__global__ void csr_count_kernel(const int64_t* __restrict__ row, int64_t n_edges, int64_t n_nodes, int32_t* counts) {
    pdl_entry();
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int64_t r = row[e];
    if (r >= 0 && r < n_nodes) atomicAdd(counts + r, 1);
}
// single-CTA exclusive scan: row_ptr[0..n] from counts[0..n-1]; cursor <- row_ptr (for the fill pass)
__global__ void csr_scan_kernel(int32_t* counts_cursor, int64_t n, int32_t* row_ptr) {
    pdl_entry();
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += blockDim.x) {
        const int64_t i = base + tid;
        const int32_t c = (i < n) ? counts_cursor[i] : 0;
        int32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int32_t w = (lane < (int)(blockDim.x >> 5)) ? warp_tot[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            warp_tot[lane] = w;     // inclusive totals
        }
        __syncthreads();
        const int32_t carry = carry_s;
        const int32_t excl = carry + (warp ? warp_tot[warp - 1] : 0) + inc - c;
        if (i < n) { row_ptr[i] = excl; counts_cursor[i] = excl; }
        __syncthreads();
        if (tid == blockDim.x - 1) carry_s = carry + warp_tot[(blockDim.x >> 5) - 1];
        __syncthreads();
    }
    if (tid == 0) row_ptr[n] = carry_s;
}
__global__ void csr_fill_kernel(const int64_t* __restrict__ row, int64_t n_edges, int64_t n_nodes,
                                int32_t* cursor, int32_t* tmp) {
    pdl_entry();
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int64_t r = row[e];
    if (r >= 0 && r < n_nodes) tmp[atomicAdd(cursor + r, 1)] = (int32_t)e;
}
// one warp per node: rank-sort its segment ascending so the permutation is the STABLE sort by row
__global__ void csr_sort_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ tmp,
                                int32_t* __restrict__ perm, int64_t n_nodes) {
    pdl_entry();
    const int64_t node = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (node >= n_nodes) return;
    const int s = row_ptr[node], e = row_ptr[node + 1];
    for (int i = s + lane; i < e; i += 32) {
        const int32_t me = tmp[i];
        int rank = 0;
        for (int j = s; j < e; ++j) rank += (tmp[j] < me);
        perm[s + rank] = me;
    }
}

// ------------------------------------------------------------------------------- fused edge kernel
constexpr int GAT_THREADS = 256;
constexpr int GAT_EB = 4;          // edges of one source node processed together

struct GatDims {
    int H, d_n, d_e, d_o, hid, din;   // din = d_n + (use_edge ? d_e : 0)
    int D_n, D_e, D_a;
    int s_c1, s_c2, s_q, s_k;         // padded smem row strides
};

__host__ __device__ inline int pad4(int x) { return ((x + 3) & ~3) + 4; }

// WS = true: the shared per-head MLP weights are staged in shared memory (mmgnet.json dims: 80 KB);
// WS = false: they are too large for it (e.g. 4 heads x 128 channels) and are read through L1/L2.
template <bool WS>
__global__ void __launch_bounds__(GAT_THREADS, 1)
gat_edge_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ v, int64_t ldv,
                const float* __restrict__ k, int64_t ldk, const int64_t* __restrict__ edge_index, const int32_t* __restrict__ row_ptr,
                const int32_t* __restrict__ perm,
                const float* __restrict__ c1, const float* __restrict__ c1b,
                const float* __restrict__ c2, const float* __restrict__ c2b,
                int64_t n_nodes, int64_t n_edges, GatDims g, int aggr, int use_edge,
                float* __restrict__ xx, int64_t ld_xx, float* __restrict__ prob, int32_t* __restrict__ argmax) {
    pdl_entry();
    extern __shared__ __align__(16) float sm[];
    const float* c1s;                                  // [hid][s_c1]   C1 rows (K-major)
    const float* c2s;                                  // [d_o][s_c2]
    float* c1bs;                                       // [hid]
    if constexpr (WS) {
        float* c1w = sm;
        float* c2w = c1w + g.hid * g.s_c1;
        c1bs = c2w + g.d_o * g.s_c2;
        for (int i = threadIdx.x; i < g.hid * g.din; i += GAT_THREADS) c1w[(i / g.din) * g.s_c1 + (i % g.din)] = __ldg(c1 + i);
        for (int i = threadIdx.x; i < g.d_o * g.hid; i += GAT_THREADS) c2w[(i / g.hid) * g.s_c2 + (i % g.hid)] = __ldg(c2 + i);
        c1s = c1w; c2s = c2w;
    } else {
        c1s = c1; c2s = c2; c1bs = sm;
    }
    float* c2bs = c1bs + g.hid;                        // [d_o]
    float* qt = c2bs + ((g.d_o + 3) & ~3);             // [H][s_q]      q of this node, de-interleaved
    float* qc = qt + g.H * g.s_q;                      // [H][hid]      C1q . q + c1 bias
    float* kt = qc + g.H * g.hid;                      // [EB][H][s_k]  edge rows, de-interleaved
    float* hs = kt + GAT_EB * g.H * g.s_k;             // [EB][H][hid]  hidden
    float* ms = hs + GAT_EB * g.H * g.hid;             // [EB][D_a]     probabilities, then messages
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NW = GAT_THREADS / 32;

    for (int i = tid; i < g.hid; i += GAT_THREADS) c1bs[i] = __ldg(c1b + i);
    for (int i = tid; i < g.d_o; i += GAT_THREADS) c2bs[i] = __ldg(c2b + i);

    for (int64_t node = blockIdx.x; node < n_nodes; node += gridDim.x) {
        const int e_lo = row_ptr[node], e_hi = row_ptr[node + 1];
        const int deg = e_hi - e_lo;
        __syncthreads();                               // weights staged / previous node finished
        if (deg > 0) {
            for (int f = tid; f < g.D_n; f += GAT_THREADS)
                qt[(f % g.H) * g.s_q + (f / g.H)] = __ldg(q + node * ldq + f);
            __syncthreads();
            // qc[h][j] = c1b[j] + sum_c C1[j][c] * q3[c][h]
            for (int idx = tid; idx < g.H * g.hid; idx += GAT_THREADS) {
                const int j = idx % g.hid, h = idx / g.hid;
                const float* wr = c1s + j * g.s_c1;
                const float* qr = qt + h * g.s_q;
                float acc = c1bs[j];
                for (int c = 0; c < g.d_n; c += 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wr + c);
                    const float4 q4 = *reinterpret_cast<const float4*>(qr + c);
                    acc = fmaf(w4.x, q4.x, acc); acc = fmaf(w4.y, q4.y, acc);
                    acc = fmaf(w4.z, q4.z, acc); acc = fmaf(w4.w, q4.w, acc);
                }
                qc[h * g.hid + j] = acc;
            }
        }
        // per-thread running aggregate of features f = tid, tid + 256, ...
        constexpr int FPT = 4;                         // supports D_a <= 1024
        float agg[FPT]; int agg_i[FPT];
#pragma unroll
        for (int r = 0; r < FPT; ++r) { agg[r] = (aggr == VLSAT_AGGR_MAX) ? -FLT_MAX : 0.f; agg_i[r] = -1; }

        for (int eb = e_lo; eb < e_hi; eb += GAT_EB) {
            const int nb = min(GAT_EB, e_hi - eb);
            __syncthreads();                           // qc ready / previous batch consumed
            if (use_edge) {
                for (int idx = tid; idx < nb * g.D_e; idx += GAT_THREADS) {
                    const int b = idx / g.D_e, f = idx % g.D_e;
                    const int64_t e = perm[eb + b];
                    kt[(b * g.H + (f % g.H)) * g.s_k + (f / g.H)] = __ldg(k + e * ldk + f);
                }
                __syncthreads();
            }
            // hidden[b][h][j] = relu(qc[h][j] + sum_c C1[j][d_n + c] * k3[b][c][h])
            for (int idx = tid; idx < nb * g.H * g.hid; idx += GAT_THREADS) {
                const int j = idx % g.hid, bh = idx / g.hid;
                float acc = qc[(bh % g.H) * g.hid + j];
                if (use_edge) {
                    const float* wr = c1s + j * g.s_c1 + g.d_n;
                    const float* kr = kt + bh * g.s_k;
                    for (int c = 0; c < g.d_e; c += 4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(wr + c);
                        const float4 k4 = *reinterpret_cast<const float4*>(kr + c);
                        acc = fmaf(w4.x, k4.x, acc); acc = fmaf(w4.y, k4.y, acc);
                        acc = fmaf(w4.z, k4.z, acc); acc = fmaf(w4.w, k4.w, acc);
                    }
                }
                hs[bh * g.hid + j] = fmaxf(acc, 0.f);
            }
            __syncthreads();
            // logits + softmax over d_o per (b, h): one warp per (b, h) pair, lanes over the d_o channels
            for (int bh = warp; bh < nb * g.H; bh += NW) {
                const int b = bh / g.H, h = bh % g.H;
                const float* hr = hs + bh * g.hid;
                float t[4];                            // supports d_o <= 128
                float mx = -FLT_MAX;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int o = lane + 32 * r;
                    t[r] = -FLT_MAX;
                    if (o < g.d_o) {
                        const float* wr = c2s + o * g.s_c2;
                        float acc = c2bs[o];
                        for (int j = 0; j < g.hid; j += 4) {
                            const float4 w4 = *reinterpret_cast<const float4*>(wr + j);
                            const float4 h4 = *reinterpret_cast<const float4*>(hr + j);
                            acc = fmaf(w4.x, h4.x, acc); acc = fmaf(w4.y, h4.y, acc);
                            acc = fmaf(w4.z, h4.z, acc); acc = fmaf(w4.w, h4.w, acc);
                        }
                        t[r] = acc;
                    }
                    mx = fmaxf(mx, t[r]);
                }
                mx = warp_max(mx);
                float sum = 0.f;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int o = lane + 32 * r;
                    t[r] = (o < g.d_o) ? expf(t[r] - mx) : 0.f;
                    sum += t[r];
                }
                sum = warp_sum(sum);
                const float inv = 1.f / sum;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int o = lane + 32 * r;
                    if (o < g.d_o) ms[b * g.D_a + o * g.H + h] = t[r] * inv;
                }
            }
            __syncthreads();
            // message = prob * value[dst]; fold into the running aggregate (feature-major, coalesced)
#pragma unroll
            for (int r = 0; r < FPT; ++r) {
                const int f = tid + r * GAT_THREADS;
                if (f < g.D_a) {
                    for (int b = 0; b < nb; ++b) {
                        const int64_t e = perm[eb + b];
                        const int64_t dst = edge_index[n_edges + e];
                        const float p = ms[b * g.D_a + f];
                        if (prob) prob[e * g.D_a + f] = p;
                        const float m = p * __ldg(v + dst * ldv + f);
                        if (aggr == VLSAT_AGGR_MAX) { if (m > agg[r]) { agg[r] = m; agg_i[r] = (int)e; } }
                        else agg[r] += m;
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < FPT; ++r) {
            const int f = tid + r * GAT_THREADS;
            if (f < g.D_a) {
                float res = agg[r];
                if (deg == 0) res = 0.f;
                else if (aggr == VLSAT_AGGR_MEAN) res /= (float)deg;
                xx[node * ld_xx + f] = res;
                if (argmax) argmax[node * g.D_a + f] = (aggr == VLSAT_AGGR_MAX) ? agg_i[r] : -1;
            }
        }
    }
}

}  // namespace vlsat

using namespace vlsat;

extern "C" int vlsat_build_csr(const int64_t* index_row, int64_t n_edges, int64_t n_nodes, int32_t* row_ptr,
                               int32_t* perm, void* workspace, size_t workspace_bytes, void* stream) {
    VLSAT_REQUIRE(row_ptr && n_edges >= 0 && n_nodes >= 0);
    VLSAT_REQUIRE(n_edges == 0 || (index_row && perm));
    VLSAT_SUPPORT(n_edges < 0x7fffffff && n_nodes < 0x7fffffff);
    if ((size_t)(n_nodes + 1 + n_edges) * sizeof(int32_t) > workspace_bytes || !workspace) return VLSAT_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* cursor = (int32_t*)workspace;
    int32_t* tmp = cursor + n_nodes + 1;
    int launches = 2;
    launch_k(csr_zero_kernel, dim3((unsigned)ceil_div(n_nodes + 1, 256)), dim3(256), 0, st, cursor, n_nodes + 1);
    if (n_edges) { launch_k(csr_count_kernel, dim3((unsigned)ceil_div(n_edges, 256)), dim3(256), 0, st, index_row, n_edges, n_nodes, cursor); ++launches; }
    launch_k(csr_scan_kernel, dim3(1), dim3(1024), 0, st, cursor, n_nodes, row_ptr);
    if (n_edges) {
        launch_k(csr_fill_kernel, dim3((unsigned)ceil_div(n_edges, 256)), dim3(256), 0, st, index_row, n_edges, n_nodes, cursor, tmp);
        launch_k(csr_sort_kernel, dim3((unsigned)ceil_div(n_nodes * 32, 256)), dim3(256), 0, st, row_ptr, tmp, perm, n_nodes);
        launches += 2;
    }
    return finish_launch(launches);
}

extern "C" int vlsat_gat_edge_fwd(const float* q, int64_t ldq, const float* v, int64_t ldv, const float* k, int64_t ldk,
                                  const int64_t* edge_index, const int32_t* row_ptr, const int32_t* perm,
                                  const float* c1, const float* c1_bias, const float* c2, const float* c2_bias,
                                  int64_t n_nodes, int64_t n_edges, int n_heads, int d_n, int d_e, int d_o, int hid,
                                  int aggr, int use_edge, float* xx, int64_t ld_xx, float* prob, int32_t* argmax, void* stream) {
    VLSAT_REQUIRE(q && v && row_ptr && c1 && c1_bias && c2 && c2_bias && xx && n_nodes >= 0 && n_edges >= 0);
    VLSAT_REQUIRE(n_edges == 0 || (edge_index && perm));
    VLSAT_REQUIRE(!use_edge || k || n_edges == 0);
    VLSAT_REQUIRE(aggr >= 0 && aggr <= 2 && n_heads >= 1 && d_n >= 1 && d_o >= 1 && hid >= 1);
    VLSAT_SUPPORT(d_n % 4 == 0 && (!use_edge || d_e % 4 == 0) && hid % 4 == 0 && d_o <= 128 && n_heads * d_o <= 1024);
    VLSAT_SUPPORT(n_nodes < 0x7fffffff && n_edges < 0x7fffffff);
    if (n_nodes == 0) return VLSAT_OK;
    VLSAT_REQUIRE(ldq >= n_heads * d_n && ldv >= n_heads * d_o && ld_xx >= n_heads * d_o && (!use_edge || ldk >= n_heads * d_e));
    GatDims g;
    g.H = n_heads; g.d_n = d_n; g.d_e = use_edge ? d_e : 0; g.d_o = d_o; g.hid = hid; g.din = d_n + g.d_e;
    g.D_n = n_heads * d_n; g.D_e = n_heads * g.d_e; g.D_a = n_heads * d_o;
    g.s_c1 = pad4(g.din); g.s_c2 = pad4(hid); g.s_q = pad4(d_n); g.s_k = pad4(g.d_e > 0 ? g.d_e : 4);
    const size_t work_floats = (size_t)hid + ((d_o + 3) & ~3) + (size_t)g.H * g.s_q + (size_t)g.H * hid +
                               (size_t)GAT_EB * g.H * g.s_k + (size_t)GAT_EB * g.H * hid + (size_t)GAT_EB * g.D_a;
    const size_t weight_floats = (size_t)hid * g.s_c1 + (size_t)d_o * g.s_c2;
    const bool ws = (work_floats + weight_floats) * sizeof(float) <= 200 * 1024;
    if (!ws) { g.s_c1 = g.din; g.s_c2 = hid; VLSAT_SUPPORT(((uintptr_t)c1 % 16 == 0) && ((uintptr_t)c2 % 16 == 0)); }
    const size_t smem = (work_floats + (ws ? weight_floats : 0)) * sizeof(float);
    VLSAT_SUPPORT(smem <= 227 * 1024);
    auto kern = ws ? gat_edge_kernel<true> : gat_edge_kernel<false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t ctas_per_sm = std::max<int64_t>(1, (int64_t)(227 * 1024) / (int64_t)(smem + 1024));
    const unsigned grid = (unsigned)std::min<int64_t>(n_nodes, ctas_per_sm * kNumSMs);
    launch_k(kern, grid, dim3(GAT_THREADS), smem, (cudaStream_t)stream, 
        q, ldq, v, ldv, k, ldk, edge_index, row_ptr, perm, c1, c1_bias, c2, c2_bias, n_nodes, n_edges, g, aggr, use_edge, xx, ld_xx, prob, argmax);
    return finish_launch();
}
