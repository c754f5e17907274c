// A8 on the tensor cores: the fused graph-attention edge kernel (max aggregation, use_edge=True).
//
// Rows are (edge, head) pairs in source-sorted (CSR) edge order: row = e * H + h. With the head-major
// operand layouts prepared by the host (proj_edge / proj_value / the C1q.proj_query fold with permuted weight
// rows) every per-row operand is a contiguous segment:
//   K'  [E*H, d_e]   : proj_edge(e) de-interleaved, tf32 hi/lo split          (TMA, 2-D, K-major)
//   QC  [N*H, hid]   : C1[:, :d_n] . q3[n, :, h] + c1_bias  per (node, head)    (gathered by src)
//   V'  [N*H, d_o]   : proj_value(x) de-interleaved                             (gathered by dst)
// Per 128-row tile (= 128/H edges), persistent CTAs:
//   MMA1 (SS, 3xTF32)   acc1[128, hid]  = K' C1k^T                      C1k = C1[:, d_n:]  resident in smem
//   epilogue 1          hidden = relu(acc1 + QC[src, h]) -> tf32 hi/lo -> TMEM (A operand of MMA2)
//   MMA2 (TS, 3xTF32)   acc2[128, d_o] = hidden C2^T                    C2 resident in smem
//   epilogue 2          p = softmax_c(acc2 + c2_bias); m = p * V'[dst, h]; segmented max over the edges of a
//                       source inside the tile, then one atomicMax (order-preserving int encoding) per
//                       (segment, feature) into xx_enc[N, H*d_o].
// A finalize kernel decodes xx_enc (untouched rows -> 0, "empty max = 0" semantics of the reference),
// and writes xx back in the reference's interleaved feature order c*H + h.
// HBM traffic = the algorithmic minimum: K' once (hi+lo), QC/V' rows (L2-resident re-reads), xx once.
#include "common.cuh"
#include "tc_common.cuh"
#include <float.h>
#include <limits.h>

namespace vlsat {

using namespace tc;

constexpr int GT_THREADS = 192;
constexpr int GT_ROWS = 128;

struct GatTcParams {
    const float* qc; int64_t ld_qc;      // QC row of (node n, head h) = qc + n * ld_qc + h * hid
    const float* v; int64_t ld_v;        // V' row                      = v  + n * ld_v  + h * d_o
    const int64_t* src; const int64_t* dst;   // CSR-sorted edge endpoints [E]
    const float* c2_bias;
    int* xx_enc;                         // [N, H * d_o] order-preserving int encoding, pre-set to INT_MIN
    float* prob;                         // optional [E, d_o, H]
    int64_t n_edges;
    int H, d_e, hid, d_o;
};

__device__ __forceinline__ int enc_ordered(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }

__device__ __forceinline__ void tmem_st_32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int DO>      // d_o (channels per head of the attention output): 32 or 64
__global__ void __launch_bounds__(GT_THREADS, 1)
gat_edge_tc_kernel(const __grid_constant__ CUtensorMap tm_khi, const __grid_constant__ CUtensorMap tm_klo,
                   const __grid_constant__ CUtensorMap tm_c1hi, const __grid_constant__ CUtensorMap tm_c1lo,
                   const __grid_constant__ CUtensorMap tm_c2hi, const __grid_constant__ CUtensorMap tm_c2lo,
                   const GatTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int kh = p.d_e / 32;                         // 128-byte column blocks of K' / C1k
    const int hc = p.hid / 32;                         // 128-byte column blocks of C2
    const uint32_t c1_box = (uint32_t)p.hid * 128, c2_box = (uint32_t)p.d_o * 128, k_box = GT_ROWS * 128;
    uint8_t* c1hi_s = smem;
    uint8_t* c1lo_s = c1hi_s + kh * c1_box;
    uint8_t* c2hi_s = c1lo_s + kh * c1_box;
    uint8_t* c2lo_s = c2hi_s + hc * c2_box;
    uint8_t* khi_s = c2lo_s + hc * c2_box;
    uint8_t* klo_s = khi_s + kh * k_box;
    float* msg = reinterpret_cast<float*>(klo_s + kh * k_box);           // [128][DO + 1]
    int* s_src = reinterpret_cast<int*>(msg + GT_ROWS * (DO + 1));       // [128] source node of each edge of the tile
    float* s_c2b = reinterpret_cast<float*>(s_src + GT_ROWS);
    uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_c2b + DO) + 7) & ~(uintptr_t)7);
    uint64_t* w_full = bars; uint64_t* k_full = bars + 1; uint64_t* k_empty = bars + 2;
    uint64_t* acc1_full = bars + 3; uint64_t* a2_ready = bars + 4; uint64_t* acc2_full = bars + 5;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t total_rows = p.n_edges * p.H;
    const int n_tiles = (int)((total_rows + GT_ROWS - 1) / GT_ROWS);
    const int ept = GT_ROWS / p.H;                     // edges per tile

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_khi); prefetch_tmap(&tm_klo); prefetch_tmap(&tm_c1hi);
        prefetch_tmap(&tm_c1lo); prefetch_tmap(&tm_c2hi); prefetch_tmap(&tm_c2lo);
        mbar_init(w_full, 1); mbar_init(k_full, 1); mbar_init(k_empty, 1);
        mbar_init(acc1_full, 1); mbar_init(a2_ready, 128); mbar_init(acc2_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_holder, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t t_acc1 = tmem_base, t_a2hi = tmem_base + p.hid, t_a2lo = tmem_base + 2 * p.hid, t_acc2 = tmem_base + 3 * p.hid;

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(w_full, 2 * kh * c1_box + 2 * hc * c2_box);
            for (int b = 0; b < kh; ++b) {
                tma_load_2d(c1hi_s + b * c1_box, &tm_c1hi, w_full, b * 32, 0);
                tma_load_2d(c1lo_s + b * c1_box, &tm_c1lo, w_full, b * 32, 0);
            }
            for (int b = 0; b < hc; ++b) {
                tma_load_2d(c2hi_s + b * c2_box, &tm_c2hi, w_full, b * 32, 0);
                tma_load_2d(c2lo_s + b * c2_box, &tm_c2lo, w_full, b * 32, 0);
            }
        }
        __syncwarp();
        int it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            mbar_wait(k_empty, (it & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(k_full, 2 * kh * k_box);
                for (int b = 0; b < kh; ++b) {
                    tma_load_2d(khi_s + b * k_box, &tm_khi, k_full, b * 32, t * GT_ROWS);
                    tma_load_2d(klo_s + b * k_box, &tm_klo, k_full, b * 32, t * GT_ROWS);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        const uint32_t idesc1 = make_idesc<Kind::TF32>(GT_ROWS, p.hid), idesc2 = make_idesc<Kind::TF32>(GT_ROWS, p.d_o);
        const uint64_t d_khi = make_sdesc_k128(smem_u32(khi_s)), d_klo = make_sdesc_k128(smem_u32(klo_s));
        const uint64_t d_c1hi = make_sdesc_k128(smem_u32(c1hi_s)), d_c1lo = make_sdesc_k128(smem_u32(c1lo_s));
        const uint64_t d_c2hi = make_sdesc_k128(smem_u32(c2hi_s)), d_c2lo = make_sdesc_k128(smem_u32(c2lo_s));
        const int ks1 = p.d_e / 8, ks2 = p.hid / 8;
        auto issue_mma1 = [&](int it) {
            mbar_wait(k_full, it & 1);
            tc_fence_after();
            if (elect_one()) {
                for (int kk = 0; kk < ks1; ++kk) {
                    const uint32_t oa = (kk >> 2) * (k_box >> 4) + (kk & 3) * 2, ob = (kk >> 2) * (c1_box >> 4) + (kk & 3) * 2;
                    mma_ss<Kind::TF32>(t_acc1, d_klo + oa, d_c1hi + ob, idesc1, kk > 0);
                    mma_ss<Kind::TF32>(t_acc1, d_khi + oa, d_c1lo + ob, idesc1, 1);
                    mma_ss<Kind::TF32>(t_acc1, d_khi + oa, d_c1hi + ob, idesc1, 1);
                }
                tc_commit(k_empty);
                tc_commit(acc1_full);
            }
            __syncwarp();
        };
        mbar_wait(w_full, 0);
        int n_local = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) ++n_local;
        if (n_local > 0) issue_mma1(0);
        for (int it = 0; it < n_local; ++it) {
            mbar_wait(a2_ready, it & 1);                 // hidden (hi, lo) is in TMEM; acc1 has been drained
            tc_fence_after();
            if (elect_one()) {
                for (int kk = 0; kk < ks2; ++kk) {
                    const uint32_t ob = (kk >> 2) * (c2_box >> 4) + (kk & 3) * 2;
                    mma_ts(t_acc2, t_a2lo + kk * 8, d_c2hi + ob, idesc2, kk > 0);
                    mma_ts(t_acc2, t_a2hi + kk * 8, d_c2lo + ob, idesc2, 1);
                    mma_ts(t_acc2, t_a2hi + kk * 8, d_c2hi + ob, idesc2, 1);
                }
                tc_commit(acc2_full);
            }
            __syncwarp();
            if (it + 1 < n_local) issue_mma1(it + 1);    // overlaps epilogue 2 of this tile
        }
    } else {
        const int qd = warp & 3;
        const int r = qd * 32 + lane;                    // row of the tile = TMEM lane
        const int et = threadIdx.x - 64;                 // 0..127 among the epilogue threads
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const int D_a = p.H * DO;
        constexpr int MS = DO + 1;
        for (int i = et; i < DO; i += 128) s_c2b[i] = __ldg(p.c2_bias + i);
        int it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int64_t row_g = (int64_t)t * GT_ROWS + r;
            const bool valid = row_g < total_rows;
            const int64_t e = valid ? row_g / p.H : 0;
            const int h = (int)(row_g % p.H);
            const int64_t e0 = (int64_t)t * ept;
            if (et < ept) s_src[et] = (e0 + et < p.n_edges) ? (int)p.src[e0 + et] : -1;
            const int64_t src = valid ? p.src[e] : 0, dst = valid ? p.dst[e] : 0;
            // operands gathered per row: issue the loads now, they land while the tensor core runs MMA1
            const float4* qrow = reinterpret_cast<const float4*>(p.qc + src * p.ld_qc + (int64_t)h * p.hid);
            const float4* vrow = reinterpret_cast<const float4*>(p.v + dst * p.ld_v + (int64_t)h * DO);
            float4 vv[DO / 4], qv[8];
#pragma unroll
            for (int j = 0; j < DO / 4; ++j) vv[j] = valid ? __ldg(vrow + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) qv[j] = valid ? __ldg(qrow + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            // ---- epilogue 1: hidden = relu(acc1 + QC[src, h]) -> tf32 hi / lo -> TMEM
            mbar_wait(acc1_full, it & 1);
            tc_fence_after();
            for (int c0 = 0; c0 < p.hid; c0 += 32) {
                uint32_t a[32], lo[32];
                tmem_ld_32x32(t_acc1 + lane_off + c0, a);
                float4 qn[8];                            // next chunk's QC values, in flight while this chunk is processed
                const bool more = c0 + 32 < p.hid;
#pragma unroll
                for (int j = 0; j < 8; ++j) qn[j] = (valid && more) ? __ldg(qrow + (c0 + 32) / 4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float hv[4] = {fmaxf(__uint_as_float(a[4 * j]) + qv[j].x, 0.f), fmaxf(__uint_as_float(a[4 * j + 1]) + qv[j].y, 0.f),
                                         fmaxf(__uint_as_float(a[4 * j + 2]) + qv[j].z, 0.f), fmaxf(__uint_as_float(a[4 * j + 3]) + qv[j].w, 0.f)};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint32_t hi;
                        hi = tc::tf32_rna_bits(hv[u]);
                        a[4 * j + u] = hi; lo[4 * j + u] = __float_as_uint(hv[u] - __uint_as_float(hi));
                    }
                }
                tmem_st_32(t_a2hi + lane_off + c0, a);
                tmem_st_32(t_a2lo + lane_off + c0, lo);
#pragma unroll
                for (int j = 0; j < 8; ++j) qv[j] = qn[j];
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(a2_ready);
            // ---- epilogue 2: softmax over d_o in registers, times value, segmented max
            mbar_wait(acc2_full, it & 1);
            tc_fence_after();
            float tl[DO];
#pragma unroll
            for (int c0 = 0; c0 < DO; c0 += 32) {
                uint32_t a[32];
                tmem_ld_32x32(t_acc2 + lane_off + c0, a);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) tl[c0 + j] = __uint_as_float(a[j]) + s_c2b[c0 + j];
            }
            tc_fence_before();
            float mx = -FLT_MAX;
#pragma unroll
            for (int c = 0; c < DO; ++c) mx = fmaxf(mx, tl[c]);
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < DO; ++c) { tl[c] = __expf(tl[c] - mx); sum += tl[c]; }
            const float inv = 1.f / sum;
#pragma unroll
            for (int c = 0; c < DO; ++c) tl[c] *= inv;
            if (p.prob && valid) {
#pragma unroll
                for (int c = 0; c < DO; ++c) p.prob[(e * DO + c) * p.H + h] = tl[c];
            }
#pragma unroll
            for (int j = 0; j < DO / 4; ++j) {
                msg[r * MS + 4 * j] = tl[4 * j] * vv[j].x; msg[r * MS + 4 * j + 1] = tl[4 * j + 1] * vv[j].y;
                msg[r * MS + 4 * j + 2] = tl[4 * j + 2] * vv[j].z; msg[r * MS + 4 * j + 3] = tl[4 * j + 3] * vv[j].w;
            }
            epi_barrier();
            for (int f = et; f < D_a; f += 128) {
                const int fh = f / DO, fc = f % DO;      // DO is a power of two: shifts
                float best = -FLT_MAX;
                int cur = s_src[0];
                for (int i = 0; i < ept; ++i) {
                    const int sn = s_src[i];
                    if (sn != cur) {
                        if (cur >= 0) atomicMax(p.xx_enc + (int64_t)cur * D_a + f, enc_ordered(best));
                        cur = sn; best = -FLT_MAX;
                    }
                    if (sn >= 0) best = fmaxf(best, msg[(i * p.H + fh) * MS + fc]);
                }
                if (cur >= 0) atomicMax(p.xx_enc + (int64_t)cur * D_a + f, enc_ordered(best));
            }
            epi_barrier();                               // msg / s_src are reused by the next tile
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

__global__ void fill_int_kernel(int* p, int64_t n, int v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// xx[n, c*H + h] = decode(xx_enc[n, h*d_o + c]); untouched (INT_MIN) -> 0
__global__ void gat_finalize_kernel(const int* __restrict__ enc, float* __restrict__ xx, int64_t ld_xx, int64_t n_nodes, int H, int d_o) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int D_a = H * d_o;
    if (idx >= n_nodes * D_a) return;
    const int64_t n = idx / D_a; const int f = (int)(idx % D_a);      // f = c*H + h (output order, coalesced writes)
    const int c = f / H, h = f % H;
    const int v = enc[n * D_a + h * d_o + c];
    xx[n * ld_xx + f] = (v == INT_MIN) ? 0.f : __int_as_float(v >= 0 ? v : v ^ 0x7fffffff);
}

// out[i, :] = in[idx[i], :]  (gather == 1)   or   out[idx[i], :] = in[i, :]  (gather == 0)
__global__ void permute_rows_kernel(const float* __restrict__ in, int64_t ld_in, const int32_t* __restrict__ idx,
                                    int64_t rows, int cols4, float* __restrict__ out, int64_t ld_out, int gather) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * cols4) return;
    const int64_t i = t / cols4; const int c = (int)(t % cols4) * 4;
    const int64_t j = idx[i];
    const int64_t ri = gather ? j : i, ro = gather ? i : j;
    *reinterpret_cast<float4*>(out + ro * ld_out + c) = __ldg(reinterpret_cast<const float4*>(in + ri * ld_in + c));
}
__global__ void permute_edges_kernel(const int64_t* __restrict__ ei, const int32_t* __restrict__ perm, int64_t n_edges, int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_edges) return;
    const int64_t j = perm[i];
    out[i] = ei[j];
    out[n_edges + i] = ei[n_edges + j];
}

}  // namespace vlsat

using namespace vlsat;

extern "C" int vlsat_permute_rows(const float* in, int64_t ld_in, const int32_t* idx, int64_t rows, int cols,
                                  float* out, int64_t ld_out, int gather, void* stream) {
    VLSAT_REQUIRE(rows >= 0 && cols >= 0);
    if (rows == 0 || cols == 0) return VLSAT_OK;
    VLSAT_REQUIRE(in && idx && out && ld_in >= cols && ld_out >= cols);
    VLSAT_SUPPORT(cols % 4 == 0 && ld_in % 4 == 0 && ld_out % 4 == 0 && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0));
    const int64_t n = rows * (cols / 4);
    permute_rows_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(in, ld_in, idx, rows, cols / 4, out, ld_out, gather);
    return finish_launch();
}

extern "C" int vlsat_permute_edges(const int64_t* edge_index, const int32_t* perm, int64_t n_edges, int64_t* out, void* stream) {
    VLSAT_REQUIRE(n_edges >= 0);
    if (n_edges == 0) return VLSAT_OK;
    VLSAT_REQUIRE(edge_index && perm && out);
    permute_edges_kernel<<<(unsigned)ceil_div(n_edges, 256), 256, 0, (cudaStream_t)stream>>>(edge_index, perm, n_edges, out);
    return finish_launch();
}

extern "C" int vlsat_gat_edge_tc_fwd(const float* k_hi, const float* k_lo, const float* qc, int64_t ld_qc,
                                     const float* v, int64_t ld_v, const int64_t* src_sorted, const int64_t* dst_sorted,
                                     const float* c1k_hi, const float* c1k_lo, const float* c2_hi, const float* c2_lo,
                                     const float* c2_bias, int64_t n_nodes, int64_t n_edges, int n_heads, int d_e, int hid,
                                     int d_o, float* xx, int64_t ld_xx, float* prob, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    VLSAT_REQUIRE(n_nodes >= 0 && n_edges >= 0 && n_heads >= 1);
    if (n_nodes == 0) return VLSAT_OK;
    VLSAT_REQUIRE(xx && ld_xx >= (int64_t)n_heads * d_o);
    VLSAT_SUPPORT(GT_ROWS % n_heads == 0 && d_e % 32 == 0 && d_e >= 32 && d_e <= 256 && hid % 32 == 0 && hid >= 32 &&
                  (d_o == 32 || d_o == 64) && 3 * hid + d_o <= 512 && hid <= 256);
    VLSAT_SUPPORT(n_edges * n_heads < (1ll << 31) && n_nodes * n_heads * d_o < (1ll << 31));
    const int D_a = n_heads * d_o;
    const size_t need = (size_t)n_nodes * D_a * sizeof(int);
    if (!workspace || workspace_bytes < need) return VLSAT_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int* enc = (int*)workspace;
    int launches = 2;
    fill_int_kernel<<<(unsigned)ceil_div(n_nodes * D_a, 256), 256, 0, st>>>(enc, n_nodes * D_a, INT_MIN);
    if (n_edges > 0) {
        VLSAT_REQUIRE(k_hi && k_lo && qc && v && src_sorted && dst_sorted && c1k_hi && c1k_lo && c2_hi && c2_lo && c2_bias);
        VLSAT_SUPPORT(ld_qc % 4 == 0 && ld_v % 4 == 0 && ((uintptr_t)qc % 16 == 0) && ((uintptr_t)v % 16 == 0));
        CUtensorMap tk, tkl, t1, t1l, t2, t2l;
        const auto F32 = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
        const uint64_t rows = (uint64_t)n_edges * n_heads;
        bool ok = make_tmap_2d(&tk, k_hi, F32, 4, rows, d_e, d_e, 32, GT_ROWS) && make_tmap_2d(&tkl, k_lo, F32, 4, rows, d_e, d_e, 32, GT_ROWS) &&
                  make_tmap_2d(&t1, c1k_hi, F32, 4, hid, d_e, d_e, 32, hid) && make_tmap_2d(&t1l, c1k_lo, F32, 4, hid, d_e, d_e, 32, hid) &&
                  make_tmap_2d(&t2, c2_hi, F32, 4, d_o, hid, hid, 32, d_o) && make_tmap_2d(&t2l, c2_lo, F32, 4, d_o, hid, hid, 32, d_o);
        if (!ok) return VLSAT_ERR_UNSUPPORTED;
        const int kh = d_e / 32, hc = hid / 32;
        const size_t smem = (size_t)2 * kh * hid * 128 + (size_t)2 * hc * d_o * 128 + (size_t)2 * kh * GT_ROWS * 128 +
                            (size_t)GT_ROWS * (d_o + 1) * 4 + 16 + GT_ROWS * 8 + (size_t)d_o * 4 + 128 + 1024;
        VLSAT_SUPPORT(smem <= 227 * 1024);
        auto kern = d_o == 32 ? gat_edge_tc_kernel<32> : gat_edge_tc_kernel<64>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        GatTcParams p;
        p.qc = qc; p.ld_qc = ld_qc; p.v = v; p.ld_v = ld_v; p.src = src_sorted; p.dst = dst_sorted; p.c2_bias = c2_bias;
        p.xx_enc = enc; p.prob = prob; p.n_edges = n_edges; p.H = n_heads; p.d_e = d_e; p.hid = hid; p.d_o = d_o;
        const int64_t n_tiles = ceil_div(n_edges * n_heads, GT_ROWS);
        const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, kNumSMs);
        kern<<<grid, GT_THREADS, smem, st>>>(tk, tkl, t1, t1l, t2, t2l, p);
        ++launches;
    }
    gat_finalize_kernel<<<(unsigned)ceil_div(n_nodes * D_a, 256), 256, 0, st>>>(enc, xx, ld_xx, n_nodes, n_heads, d_o);
    return finish_launch(launches);
}
