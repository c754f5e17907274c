// A8 on the tensor cores: the fused graph-attention edge kernel (max aggregation, use_edge=True), BF16x3.
//
// Rows are (edge, head) pairs in source-sorted (CSR) edge order: row = e * H + h. With the head-major
// operand layouts prepared by the host (proj_edge / proj_value / the C1q.proj_query fold with permuted weight
// rows) every per-row operand is a contiguous segment:
//   K'  [E*H, 64]    : proj_edge(e) de-interleaved, bf16 (hi, lo) pair = 4 B / element, exactly the algorithmic
//                      edge bytes; emitted by the producing GEMM epilogue                     (TMA, 128-byte rows)
//   QC  [N*H, hid]   : C1[:, :d_n] . q3[n, :, h] + c1_bias  per (node, head)    (gathered by src; CSR order makes
//                      the 16 edges of a tile share 1-2 sources, so these rows are L1 hits)
//   V'  [N*H, d_o]   : proj_value(x) de-interleaved                             (gathered by dst)
// Per 128-row tile (= 128/H edges), persistent CTAs, every product as lo*hi + hi*lo + hi*hi (kind::f16, bf16):
//   MMA1 (SS)    acc1[128, hid]  = K' C1k^T                      C1k = C1[:, d_n:]  resident in smem
//   epilogue 1   hidden = relu(acc1 + QC[src, h]) -> bf16 pair written IN PLACE over the accumulator columns
//                (32 fp32 columns -> 16 packed hi words + 16 packed lo words): the A operand of MMA2 in TMEM
//   MMA2 (TS)    acc2[128, d_o] = hidden C2^T                    C2 resident in smem
//   epilogue 2   p = softmax_c(acc2 + c2_bias); m = p * V'[dst, h]; segmented max over the edges of a
//                source inside the tile, then one atomicMax (order-preserving int encoding) per
//                (segment, feature) into xx_enc[N, H*d_o].
// Pipelining: GT_WG epilogue warpgroups, each with its own TMEM buffer (acc1/hidden + acc2) and message buffer,
// take tiles round-robin and run both epilogues of their tile; the MMA warp issues MMA1 of the next tiles while
// they work, so a tile's latency chain (TMA -> MMA1 -> epi1 -> MMA2 -> epi2) overlaps with GT_WG - 1 others.
// A finalize kernel decodes xx_enc (untouched rows -> 0, "empty max = 0" semantics of the reference), restores the
// INT_MIN fill of the workspace (a caller that keeps it passes workspace_ready = 1 and saves the fill) and writes xx back in the reference's interleaved feature order c*H + h.
// HBM traffic = the algorithmic minimum: K' once, QC/V' rows (cache-resident re-reads), xx once.
#include "common.cuh"
#include "epilogue.cuh"
#include "tc_common.cuh"
#include <float.h>
#include <limits.h>
#include <stdlib.h>

namespace vlsat {

using namespace tc;

// epilogue warpgroups = tiles in flight per CTA: three where the TMEM columns (GT_WG * (hid + d_o) <= 512) and the
// register file (448 threads -> 128 registers each; the d_o = 32 kernel fits with 48 bytes of spill) allow it
// (VLSAT_GAT_WG=2 forces two, for A/B timing)
static int gt_wg(int d_o) {
    static const int forced = [] { const char* e = getenv("VLSAT_GAT_WG"); return e ? atoi(e) : 0; }();
    return (d_o <= 32 && forced != 2) ? 3 : 2;
}
constexpr int GT_ROWS = 128;
constexpr int GT_DE = 64;                     // proj_edge channels per head: one 128-byte bf16 row
constexpr int GT_STAGES = 3;                  // K' tile ring
constexpr int GT_K_TILE = 2 * GT_ROWS * 128;  // K'_hi | K'_lo
constexpr int GT_QN = 4;                      // QC rows of up to this many consecutive source nodes are staged in smem per tile

struct GatTcParams {
    const float* qc; int64_t ld_qc;      // QC row of (node n, head h) = qc + n * ld_qc + h * hid
    const float* v; int64_t ld_v;        // V' row                      = v  + n * ld_v  + h * d_o
    const int64_t* src; const int64_t* dst;   // CSR-sorted edge endpoints [E]
    const float* c2_bias;
    int* xx_enc;                         // [N, H * d_o] order-preserving int encoding, holds INT_MIN on entry
    float* prob;                         // optional [E, d_o, H]
    int64_t n_edges;
    int H, hid, d_o;
    int dbg;                             // debug experiments (VLSAT_GAT_DBG): 1 no atomics, 2 no QC loads, 4 no smem staging of QC rows
    long long* trace;                    // debug: globaltimer stamps of CTA 0 / warpgroup 0 (6 per tile), nullptr in normal use
};

extern long long* g_trace;               // csrc/gemm_tc.cu (vlsat_debug_set_trace)
__device__ __forceinline__ long long gt_time() { unsigned long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); return (long long)g; }
#define GT_STAMP(k) do { if (p.trace && blockIdx.x == 0 && g == 0 && et == 0 && j < 16) p.trace[j * 12 + (k)] = gt_time(); } while (0)

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
// D[tmem] (+)= A[tmem] . B[smem]^T, bf16 inputs: the A operand is read from tensor memory (lane = row, each 32-bit
// column holds two consecutive K elements)
__device__ __forceinline__ void mma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_st_16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// d_o (channels per head of the attention output): 32 or 64; epilogue warpgroups; PASSES: 3 = BF16x3, 1 = the hi halves
// alone (single-pass bf16 mode: K'_lo, C1_lo, C2_lo are neither loaded nor multiplied, the hidden layer keeps its hi words)
template <int DO, int GT_WG, int PASSES>
__global__ void __launch_bounds__(64 + 128 * GT_WG, 1)
gat_edge_tc_kernel(const __grid_constant__ CUtensorMap tm_khi, const __grid_constant__ CUtensorMap tm_klo,
                   const __grid_constant__ CUtensorMap tm_c1hi, const __grid_constant__ CUtensorMap tm_c1lo,
                   const __grid_constant__ CUtensorMap tm_c2hi, const __grid_constant__ CUtensorMap tm_c2lo,
                   const GatTcParams p) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array keeps LDS/STS
    const int hc = (p.hid + 63) / 64;                        // 128-byte (64 bf16) column blocks of C2
    const uint32_t c1_box = (uint32_t)p.hid * 128, c2_box = (uint32_t)DO * 128;
    uint8_t* c1hi_s = smem;                                  // [hid rows x 128 B]
    uint8_t* c1lo_s = c1hi_s + c1_box;
    uint8_t* c2hi_s = c1lo_s + c1_box;                       // hc x [DO rows x 128 B]
    uint8_t* c2lo_s = c2hi_s + hc * c2_box;
    uint8_t* k_s = c2lo_s + hc * c2_box;                     // GT_STAGES x (K'_hi | K'_lo)
    float* s_c2b = reinterpret_cast<float*>(k_s + GT_STAGES * GT_K_TILE);   // [DO] second-layer bias
    // per-warpgroup scratch, two uses that never overlap in time: the staged QC rows [GT_QN nodes][H heads][hid + 4]
    // (written at the top of a tile, last read in epilogue 1) and the run-maxima transpose [4 warps][32 rows][DO + 1]
    // of epilogue 2 - a warp enters epilogue 2 only after acc2_full, i.e. after all 128 threads arrived on hid_ready
    float* qs_all = s_c2b + DO;
    const int qs_words = max(GT_QN * p.H * (p.hid + 4), 4 * 32 * (DO + 1));
    uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(qs_all + GT_WG * qs_words) + 7) & ~(uintptr_t)7);
    uint64_t* w_full = bars;
    uint64_t* k_full = bars + 1;                             // [GT_STAGES]
    uint64_t* k_empty = k_full + GT_STAGES;                  // [GT_STAGES]
    uint64_t* acc1_full = k_empty + GT_STAGES;               // [GT_WG]
    uint64_t* hid_ready = acc1_full + GT_WG;                 // [GT_WG]
    uint64_t* acc2_full = hid_ready + GT_WG;                 // [GT_WG]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc2_full + GT_WG);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t total_rows = p.n_edges * p.H;          // < 2^31 (checked by the launcher)
    const int n_tiles = (int)((total_rows + GT_ROWS - 1) / GT_ROWS);
    const int ept = GT_ROWS / p.H;                     // edges per tile
    int n_local = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) ++n_local;
    const uint32_t buf_cols = (uint32_t)p.hid + DO;    // TMEM columns of one buffer: acc1 / hidden pair, then acc2

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_khi); prefetch_tmap(&tm_klo); prefetch_tmap(&tm_c1hi);
        prefetch_tmap(&tm_c1lo); prefetch_tmap(&tm_c2hi); prefetch_tmap(&tm_c2lo);
        mbar_init(w_full, 1);
        for (int s = 0; s < GT_STAGES; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
        for (int g = 0; g < GT_WG; ++g) { mbar_init(&acc1_full[g], 1); mbar_init(&hid_ready[g], 128); mbar_init(&acc2_full[g], 1); }
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_holder, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();                                            // everything above is local to the CTA

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(w_full, (PASSES == 3 ? 2 : 1) * (c1_box + hc * c2_box));
            tma_load_2d(c1hi_s, &tm_c1hi, w_full, 0, 0);
            if (PASSES == 3) tma_load_2d(c1lo_s, &tm_c1lo, w_full, 0, 0);
            for (int b = 0; b < hc; ++b) {
                tma_load_2d(c2hi_s + b * c2_box, &tm_c2hi, w_full, b * 64, 0);
                if (PASSES == 3) tma_load_2d(c2lo_s + b * c2_box, &tm_c2lo, w_full, b * 64, 0);
            }
        }
        __syncwarp();
        int it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int s = it % GT_STAGES;
            mbar_wait(&k_empty[s], ((it / GT_STAGES) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&k_full[s], PASSES == 3 ? GT_K_TILE : GT_K_TILE / 2);
                tma_load_2d(k_s + s * GT_K_TILE, &tm_khi, &k_full[s], 0, t * GT_ROWS);
                if (PASSES == 3) tma_load_2d(k_s + s * GT_K_TILE + GT_ROWS * 128, &tm_klo, &k_full[s], 0, t * GT_ROWS);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        const uint32_t idesc1 = make_idesc<Kind::BF16>(GT_ROWS, p.hid), idesc2 = make_idesc<Kind::BF16>(GT_ROWS, DO);
        const uint64_t d_k0 = make_sdesc_k128(smem_u32(k_s));
        const uint64_t d_c1hi = make_sdesc_k128(smem_u32(c1hi_s)), d_c1lo = make_sdesc_k128(smem_u32(c1lo_s));
        const uint64_t d_c2hi = make_sdesc_k128(smem_u32(c2hi_s)), d_c2lo = make_sdesc_k128(smem_u32(c2lo_s));
        const int ks2 = p.hid / 16;
        auto issue_mma1 = [&](int it) {
            const int s = it % GT_STAGES, b = it % GT_WG;
            mbar_wait(&k_full[s], (it / GT_STAGES) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t d_khi = d_k0 + (uint64_t)(s * (GT_K_TILE >> 4)), d_klo = d_khi + ((GT_ROWS * 128) >> 4);
                const uint32_t t_acc1 = tmem_base + b * buf_cols;
#pragma unroll
                for (int kk = 0; kk < GT_DE / 16; ++kk) {       // 16 channels (32 bytes) per MMA
                    if (PASSES == 3) {
                        mma_ss<Kind::BF16>(t_acc1, d_klo + 2 * kk, d_c1hi + 2 * kk, idesc1, kk > 0);
                        mma_ss<Kind::BF16>(t_acc1, d_khi + 2 * kk, d_c1lo + 2 * kk, idesc1, 1);
                    }
                    mma_ss<Kind::BF16>(t_acc1, d_khi + 2 * kk, d_c1hi + 2 * kk, idesc1, PASSES == 3 ? 1u : (kk > 0 ? 1u : 0u));
                }
                tc_commit(&k_empty[s]);
                tc_commit(&acc1_full[b]);
            }
            __syncwarp();
        };
        mbar_wait(w_full, 0);
        for (int it = 0; it < GT_WG && it < n_local; ++it) issue_mma1(it);
        for (int it = 0; it < n_local; ++it) {
            const int b = it % GT_WG;
            mbar_wait(&hid_ready[b], (it / GT_WG) & 1);        // hidden (hi, lo) is in TMEM; acc1 / acc2 of this buffer are drained
            tc_fence_after();
            if (elect_one()) {
                const uint32_t t_hid = tmem_base + b * buf_cols, t_acc2 = t_hid + p.hid;
                for (int kk = 0; kk < ks2; ++kk) {              // 16 hidden units per MMA = 8 packed columns
                    const uint32_t a_hi = t_hid + 32 * (kk >> 1) + 8 * (kk & 1), a_lo = a_hi + 16;
                    const uint32_t ob = (kk >> 2) * (c2_box >> 4) + (kk & 3) * 2;
                    if (PASSES == 3) {
                        mma_ts_bf16(t_acc2, a_lo, d_c2hi + ob, idesc2, kk > 0);
                        mma_ts_bf16(t_acc2, a_hi, d_c2lo + ob, idesc2, 1);
                    }
                    mma_ts_bf16(t_acc2, a_hi, d_c2hi + ob, idesc2, PASSES == 3 ? 1u : (kk > 0 ? 1u : 0u));
                }
                tc_commit(&acc2_full[b]);
            }
            __syncwarp();
            if (it + GT_WG < n_local) issue_mma1(it + GT_WG);  // ordered behind MMA2(it): may overwrite its hidden operand
        }
    } else {
        const int g = (warp - 2) >> 2;                       // epilogue warpgroup = TMEM buffer
        const int qd = warp & 3;
        const int r = qd * 32 + lane;                        // row of the tile = TMEM lane
        const int et = (threadIdx.x - 64) & 127;             // 0..127 inside the warpgroup
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const uint32_t t_acc1 = tmem_base + g * buf_cols + lane_off, t_acc2 = t_acc1 + p.hid;
        const int D_a = p.H * DO;
        float* tps = qs_all + g * qs_words + ((warp - 2) & 3) * 32 * (DO + 1);
        if (g == 0) for (int i = et; i < DO; i += 128) s_c2b[i] = __ldg(p.c2_bias + i);
        asm volatile("bar.sync 15, %0;" ::"n"(128 * GT_WG) : "memory");     // c2 bias visible to every epilogue warpgroup
        const int hshift = 31 - __clz(p.H);                  // H divides 128: a power of two
        int j = 0;                                           // this warpgroup's tile counter
        // row metadata of a tile: (valid, edge, head) of this thread's row and its two gathered rows
        auto locate = [&](int it_, bool& valid_, int& e_, int& h_, int64_t& src_, int64_t& dst_) {
            const int row_g = (blockIdx.x + it_ * (int)gridDim.x) * GT_ROWS + r;      // < 2^31 (checked by the launcher)
            valid_ = it_ < n_local && row_g < total_rows;
            e_ = valid_ ? row_g >> hshift : 0;
            h_ = row_g & (p.H - 1);
            src_ = valid_ ? p.src[e_] : 0; dst_ = valid_ ? p.dst[e_] : 0;
        };
        bool valid, valid_n; int e, e_n, h, h_n; int64_t src, src_n, dst, dst_n;
        locate(g, valid, e, h, src, dst);
        // QC staging. The edges of a tile come from a run of consecutive source nodes (CSR order; 1-3 nodes at the graph
        // densities of the path), so instead of every (edge, head) row gathering its own 4*hid bytes - 16 rows per source
        // asking for the same lines, measured as > 1 us of exposed load latency per tile - the warpgroup copies the QC rows
        // of nodes [first, first + cnt) once, coalesced, one tile ahead (registers -> padded smem rows at the top of the
        // tile), and epilogue 1 reads them with conflict-free LDS.128. Tiles spanning more than GT_QN nodes gather directly.
        const int ppn = p.H * p.hid / 4;                     // 16-byte pieces per node row
        const bool stage_ok = (ppn & (ppn - 1)) == 0 && !(p.dbg & 4);
        const int lg_ppn = 31 - __clz(ppn), lg_hp = 31 - __clz(p.hid / 4);
        float* qs_g = qs_all + g * qs_words;
        auto wg_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); };
        // source ids of the first / last edge of a tile: loaded two tiles ahead, so that sizing the next block never waits
        auto ends_of = [&](int it_, int64_t& sa_, int64_t& sb_) {
            sa_ = 0; sb_ = GT_QN;                            // a span the staging rejects
            if (it_ < n_local && stage_ok) {
                const int64_t ea = (int64_t)(blockIdx.x + it_ * (int)gridDim.x) * ept, eb = min(ea + ept, p.n_edges) - 1;
                sa_ = p.src[ea]; sb_ = p.src[eb];
            }
        };
        auto block_of = [&](int64_t sa_, int64_t sb_, int64_t& first_, int& cnt_) {   // staged node range (cnt 0: gather directly)
            first_ = sa_;
            cnt_ = (sb_ - sa_) < GT_QN ? (int)(sb_ - sa_) + 1 : 0;
        };
        float4 qreg[8];                                      // this thread's pieces of the NEXT tile's block, in flight
        auto fetch_block = [&](int64_t first_, int cnt_) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = et + 128 * k;
                if (i < cnt_ * ppn)
                    qreg[k] = __ldg(reinterpret_cast<const float4*>(p.qc + (first_ + (i >> lg_ppn)) * p.ld_qc) + (i & (ppn - 1)));
            }
        };
        int64_t blk_first, blk_first_n, sa_n, sb_n; int blk_cnt, blk_cnt_n;
        ends_of(g, sa_n, sb_n);
        block_of(sa_n, sb_n, blk_first, blk_cnt);
        fetch_block(blk_first, blk_cnt);
        ends_of(g + GT_WG, sa_n, sb_n);
        for (int it = g; it < n_local; it += GT_WG, ++j) {
            GT_STAMP(0);
            wg_sync();                                       // every reader of the previous tile's staged rows is done
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = et + 128 * k;
                if (i < blk_cnt * ppn) {
                    const int off = i & (ppn - 1);
                    *reinterpret_cast<float4*>(qs_g + (((i >> lg_ppn) * p.H + (off >> lg_hp)) * (p.hid + 4)) + (off & (p.hid / 4 - 1)) * 4) = qreg[k];
                }
            }
            wg_sync();
            block_of(sa_n, sb_n, blk_first_n, blk_cnt_n);
            fetch_block(blk_first_n, blk_cnt_n);             // lands during this tile
            ends_of(it + 2 * GT_WG, sa_n, sb_n);
            locate(it + GT_WG, valid_n, e_n, h_n, src_n, dst_n);     // next tile's endpoints: consumed after epilogue 1
            const float4* qrow = reinterpret_cast<const float4*>(p.qc + src * p.ld_qc + (int64_t)h * p.hid);
            const float4* vrow = reinterpret_cast<const float4*>(p.v + dst * p.ld_v + (int64_t)h * DO);
            const float4* qs_row = reinterpret_cast<const float4*>(qs_g + ((valid ? (int)(src - blk_first) : 0) * p.H + h) * (p.hid + 4));
            const bool staged = blk_cnt > 0;
            // ---- epilogue 1: hidden = relu(acc1 + QC[src, h]) -> bf16 pair, in place over the accumulator columns
            mbar_wait(&acc1_full[g], j & 1);
            tc_fence_after();
            GT_STAMP(1);
            for (int c0 = 0; c0 < p.hid; c0 += 32) {
                uint32_t a[32];
                tmem_ld_32x32(t_acc1 + c0, a);
                float4 qv[8];
                if (staged) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) qv[u] = qs_row[c0 / 4 + u];
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) qv[u] = (valid && !(p.dbg & 2)) ? __ldg(qrow + c0 / 4 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                tmem_ld_wait();
                if (c0 == 0) GT_STAMP(6);
                uint32_t w[32];                              // 16 packed hi words, then 16 packed lo words
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float h0 = fmaxf(__uint_as_float(a[4 * u]) + qv[u].x, 0.f), h1 = fmaxf(__uint_as_float(a[4 * u + 1]) + qv[u].y, 0.f);
                    const float h2 = fmaxf(__uint_as_float(a[4 * u + 2]) + qv[u].z, 0.f), h3 = fmaxf(__uint_as_float(a[4 * u + 3]) + qv[u].w, 0.f);
                    split_bf16x2(h0, h1, w[2 * u], w[16 + 2 * u]);
                    split_bf16x2(h2, h3, w[2 * u + 1], w[16 + 2 * u + 1]);
                }
                if (c0 == 0) GT_STAMP(7);
                if (PASSES == 3) tmem_st_32(t_acc1 + c0, w);
                else tmem_st_16(t_acc1 + c0, w);                 // hi words only (split_bf16x2 rounds hi to nearest)
            }
            float4 vv[DO / 4];                               // value row: lands while the tensor core runs MMA2
#pragma unroll
            for (int u = 0; u < DO / 4; ++u) vv[u] = valid ? __ldg(vrow + u) : make_float4(0.f, 0.f, 0.f, 0.f);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&hid_ready[g]);
            GT_STAMP(2);
            // ---- epilogue 2: softmax over d_o in registers, times value, segmented max
            mbar_wait(&acc2_full[g], j & 1);
            tc_fence_after();
            GT_STAMP(3);
            float tl[DO];
#pragma unroll
            for (int c0 = 0; c0 < DO; c0 += 32) {
                uint32_t a[32];
                tmem_ld_32x32(t_acc2 + c0, a);
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float4 b4 = *reinterpret_cast<const float4*>(s_c2b + c0 + 4 * u);
                    tl[c0 + 4 * u] = __uint_as_float(a[4 * u]) + b4.x; tl[c0 + 4 * u + 1] = __uint_as_float(a[4 * u + 1]) + b4.y;
                    tl[c0 + 4 * u + 2] = __uint_as_float(a[4 * u + 2]) + b4.z; tl[c0 + 4 * u + 3] = __uint_as_float(a[4 * u + 3]) + b4.w;
                }
            }
            tc_fence_before();
            // softmax over the d_o channels of this (edge, head) row; four partial maxima / sums keep the chains short
            // (one row per thread, two warps per scheduler: dependent-issue latency is what this epilogue pays for)
            float mxp[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
#pragma unroll
            for (int c = 0; c < DO; ++c) mxp[c & 3] = fmaxf(mxp[c & 3], tl[c]);
            const float neg_m = -fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3])) * 1.4426950408889634f;
            float smp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < DO; ++c) {
                float ex;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fmaf(tl[c], 1.4426950408889634f, neg_m)));
                tl[c] = ex; smp[c & 3] += ex;
            }
            const float inv = 1.f / ((smp[0] + smp[1]) + (smp[2] + smp[3]));
#pragma unroll
            for (int c = 0; c < DO; ++c) tl[c] *= inv;
            if (p.prob && valid) {
#pragma unroll
                for (int c = 0; c < DO; ++c) p.prob[((int64_t)e * DO + c) * p.H + h] = tl[c];
            }
#pragma unroll
            for (int u = 0; u < DO / 4; ++u) { tl[4 * u] *= vv[u].x; tl[4 * u + 1] *= vv[u].y; tl[4 * u + 2] *= vv[u].z; tl[4 * u + 3] *= vv[u].w; }
            GT_STAMP(4);
            // Max over the edges of a source, in registers: the lanes of a warp are (edge, head) pairs, edges in CSR order,
            // so a segmented max-scan along the edge axis (lane distance H, 2H, ...) leaves the run maximum in the last
            // edge of every run inside the warp; that lane sends one atomicMax per channel (order-preserving int encoding).
            const int my_src = valid ? (int)src : -1;
            if (p.H < 32) {
                for (int d = p.H; d < 32; d <<= 1) {
                    const int o_src = __shfl_up_sync(0xffffffffu, my_src, d);
                    const bool take = lane >= d && o_src == my_src;
#pragma unroll
                    for (int c = 0; c < DO; ++c) {
                        const float o = __shfl_up_sync(0xffffffffu, tl[c], d);
                        tl[c] = take ? fmaxf(tl[c], o) : tl[c];
                    }
                }
            }
            const int nx_src = __shfl_down_sync(0xffffffffu, my_src, p.H & 31);
            const bool last = my_src >= 0 && (p.H >= 32 || lane + p.H >= 32 || nx_src != my_src);
            // The run maxima leave through a per-warp smem transpose so that one atomic instruction covers 32 consecutive
            // words of xx_enc (a lane owns the d_o channels of ONE head; sent directly, every lane would hit its own line).
            const int hs = min(p.H, 32), hh = h & 31;                         // heads per warp row group, head inside it
            // index of my run in the warp = closed runs among the edges before mine (head-0 lanes mark them)
            const int run = __popc(__ballot_sync(0xffffffffu, last && hh == 0) & ((1u << (lane & ~(hs - 1))) - 1u));
            float* tp = tps + (run * hs + hh) * (DO + 1);
            if (last) {
#pragma unroll
                for (int c = 0; c < DO; ++c) tp[c] = tl[c];
            }
            const unsigned run_heads = __ballot_sync(0xffffffffu, last);      // one bit per (run, head) row written
            __syncwarp();
            for (unsigned m = run_heads; m && !(p.dbg & 1); m &= m - 1) {
                const int owner = __ffs(m) - 1;                               // lane that owns this (run, head) row
                const int o_src = __shfl_sync(0xffffffffu, my_src, owner);
                const int o_h = __shfl_sync(0xffffffffu, h, owner);
                const int o_run = __shfl_sync(0xffffffffu, run, owner);
                const float* row = tps + (o_run * hs + (o_h & 31)) * (DO + 1);
                int* const dst_row = p.xx_enc + (int64_t)o_src * D_a + o_h * DO;
#pragma unroll
                for (int c = lane; c < DO; c += 32) atomicMax(dst_row + c, enc_ordered(row[c]));
            }
            __syncwarp();                                                     // the transpose buffer is reused by the next tile
            GT_STAMP(5);
            valid = valid_n; e = e_n; h = h_n; src = src_n; dst = dst_n; blk_first = blk_first_n; blk_cnt = blk_cnt_n;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

__global__ void fill_int_kernel(int* p, int64_t n, int v) {
    pdl_entry();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// xx[n, c*H + h] = decode(xx_enc[n, h*d_o + c]); untouched (INT_MIN) -> 0
// and restore the INT_MIN fill, so the same workspace can be handed to the next call with workspace_ready = 1
__global__ void gat_finalize_kernel(int* __restrict__ enc, float* __restrict__ xx, int64_t ld_xx, int64_t n_nodes, int H, int d_o) {
    pdl_entry();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int D_a = H * d_o;
    if (idx >= n_nodes * D_a) return;
    const int64_t n = idx / D_a; const int f = (int)(idx % D_a);      // f = c*H + h (output order, coalesced writes)
    const int c = f / H, h = f % H;
    const int v = enc[n * D_a + h * d_o + c];
    enc[n * D_a + h * d_o + c] = INT_MIN;
    xx[n * ld_xx + f] = (v == INT_MIN) ? 0.f : __int_as_float(v >= 0 ? v : v ^ 0x7fffffff);
}

// out[i, :] = in[idx[i], :]  (gather == 1)   or   out[idx[i], :] = in[i, :]  (gather == 0); optionally also the bf16
// (hi, lo) pair of the permuted rows (compact [rows, cols]), the operand format of the projections that read them next
__global__ void permute_rows_kernel(const float* __restrict__ in, int64_t ld_in, const int32_t* __restrict__ idx,
                                    int64_t rows, int cols4, float* __restrict__ out, int64_t ld_out, int gather,
                                    uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    pdl_entry();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * cols4) return;
    const int64_t i = t / cols4; const int c = (int)(t % cols4) * 4;
    const int64_t j = idx[i];
    const int64_t ri = gather ? j : i, ro = gather ? i : j;
    const float4 v = __ldg(reinterpret_cast<const float4*>(in + ri * ld_in + c));
    *reinterpret_cast<float4*>(out + ro * ld_out + c) = v;
    if (hi) {
        uint32_t h0, l0, h1, l1;
        split_bf16x2(v.x, v.y, h0, l0); split_bf16x2(v.z, v.w, h1, l1);
        *reinterpret_cast<uint2*>(hi + ro * (int64_t)cols4 * 4 + c) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(lo + ro * (int64_t)cols4 * 4 + c) = make_uint2(l0, l1);
    }
}
__global__ void permute_edges_kernel(const int64_t* __restrict__ ei, const int32_t* __restrict__ perm, int64_t n_edges, int64_t* __restrict__ out) {
    pdl_entry();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_edges) return;
    const int64_t j = perm[i];
    out[i] = ei[j];
    out[n_edges + i] = ei[n_edges + j];
}

}  // namespace vlsat

using namespace vlsat;

extern "C" int vlsat_permute_rows(const float* in, int64_t ld_in, const int32_t* idx, int64_t rows, int cols,
                                  float* out, int64_t ld_out, int gather, void* split_hi, void* split_lo, void* stream) {
    VLSAT_REQUIRE(rows >= 0 && cols >= 0);
    if (rows == 0 || cols == 0) return VLSAT_OK;
    VLSAT_REQUIRE(in && idx && out && ld_in >= cols && ld_out >= cols);
    VLSAT_SUPPORT(cols % 4 == 0 && ld_in % 4 == 0 && ld_out % 4 == 0 && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0));
    const int64_t n = rows * (cols / 4);
    VLSAT_REQUIRE((split_hi == nullptr) == (split_lo == nullptr));
    VLSAT_SUPPORT(!split_hi || (((uintptr_t)split_hi | (uintptr_t)split_lo) % 8 == 0));
    launch_k(permute_rows_kernel, dim3((unsigned)ceil_div(n, 256)), dim3(256), 0, (cudaStream_t)stream, in, ld_in, idx, rows, cols / 4, out, ld_out, gather,
                                                                                      (uint16_t*)split_hi, (uint16_t*)split_lo);
    return finish_launch();
}

extern "C" int vlsat_permute_edges(const int64_t* edge_index, const int32_t* perm, int64_t n_edges, int64_t* out, void* stream) {
    VLSAT_REQUIRE(n_edges >= 0);
    if (n_edges == 0) return VLSAT_OK;
    VLSAT_REQUIRE(edge_index && perm && out);
    launch_k(permute_edges_kernel, dim3((unsigned)ceil_div(n_edges, 256)), dim3(256), 0, (cudaStream_t)stream, edge_index, perm, n_edges, out);
    return finish_launch();
}

extern "C" int vlsat_gat_edge_tc_fwd(const void* k_hi, const void* k_lo, const float* qc, int64_t ld_qc,
                                     const float* v, int64_t ld_v, const int64_t* src_sorted, const int64_t* dst_sorted,
                                     const void* c1k_hi, const void* c1k_lo, const void* c2_hi, const void* c2_lo,
                                     const float* c2_bias, int64_t n_nodes, int64_t n_edges, int n_heads, int d_e, int hid,
                                     int d_o, float* xx, int64_t ld_xx, float* prob, void* workspace, size_t workspace_bytes,
                                     int workspace_ready, void* stream) {
    VLSAT_REQUIRE(n_nodes >= 0 && n_edges >= 0 && n_heads >= 1);
    if (n_nodes == 0) return VLSAT_OK;
    VLSAT_REQUIRE(xx && ld_xx >= (int64_t)n_heads * d_o);
    VLSAT_SUPPORT(GT_ROWS % n_heads == 0 && d_e == GT_DE && hid % 32 == 0 && hid >= 32 && hid <= 128 &&
                  (d_o == 32 || d_o == 64) && gt_wg(d_o) * (hid + d_o) <= 512);
    VLSAT_SUPPORT(n_edges * n_heads < (1ll << 31) && n_nodes * n_heads * d_o < (1ll << 31));
    const int D_a = n_heads * d_o;
    const size_t need = (size_t)n_nodes * D_a * sizeof(int);
    if (!workspace || workspace_bytes < need) return VLSAT_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int* enc = (int*)workspace;
    int launches = 1;
    if (!workspace_ready) {
        launch_k(fill_int_kernel, dim3((unsigned)ceil_div(n_nodes * D_a, 256)), dim3(256), 0, st, enc, n_nodes * D_a, INT_MIN);
        ++launches;
    }
    if (n_edges > 0) {
        VLSAT_REQUIRE(k_hi && k_lo && qc && v && src_sorted && dst_sorted && c1k_hi && c1k_lo && c2_hi && c2_lo && c2_bias);
        VLSAT_SUPPORT(ld_qc % 4 == 0 && ld_v % 4 == 0 && ((uintptr_t)qc % 16 == 0) && ((uintptr_t)v % 16 == 0));
        VLSAT_SUPPORT((((uintptr_t)k_hi | (uintptr_t)k_lo | (uintptr_t)c1k_hi | (uintptr_t)c1k_lo | (uintptr_t)c2_hi | (uintptr_t)c2_lo) & 15) == 0);
        CUtensorMap tk, tkl, t1, t1l, t2, t2l;
        const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
        const uint64_t rows = (uint64_t)n_edges * n_heads;
        bool ok = make_tmap_2d(&tk, k_hi, BF, 2, rows, d_e, d_e, 64, GT_ROWS) && make_tmap_2d(&tkl, k_lo, BF, 2, rows, d_e, d_e, 64, GT_ROWS) &&
                  make_tmap_2d(&t1, c1k_hi, BF, 2, hid, d_e, d_e, 64, hid) && make_tmap_2d(&t1l, c1k_lo, BF, 2, hid, d_e, d_e, 64, hid) &&
                  make_tmap_2d(&t2, c2_hi, BF, 2, d_o, hid, hid, 64, d_o) && make_tmap_2d(&t2l, c2_lo, BF, 2, d_o, hid, hid, 64, d_o);
        if (!ok) return VLSAT_ERR_UNSUPPORTED;
        const int hc = (hid + 63) / 64;
        const size_t smem = (size_t)2 * hid * 128 + (size_t)2 * hc * d_o * 128 + (size_t)GT_STAGES * GT_K_TILE +
                            (size_t)d_o * 4 +
                            (size_t)gt_wg(d_o) * std::max(GT_QN * n_heads * (hid + 4), 4 * 32 * (d_o + 1)) * 4 + 256 + 1024;
        VLSAT_SUPPORT(smem <= 227 * 1024);
        const int wg = gt_wg(d_o);
        auto kern = tc_passes() == 1 ? (d_o == 64 ? gat_edge_tc_kernel<64, 2, 1> : wg == 3 ? gat_edge_tc_kernel<32, 3, 1> : gat_edge_tc_kernel<32, 2, 1>)
                                     : (d_o == 64 ? gat_edge_tc_kernel<64, 2, 3> : wg == 3 ? gat_edge_tc_kernel<32, 3, 3> : gat_edge_tc_kernel<32, 2, 3>);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        GatTcParams p;
        p.qc = qc; p.ld_qc = ld_qc; p.v = v; p.ld_v = ld_v; p.src = src_sorted; p.dst = dst_sorted; p.c2_bias = c2_bias;
        p.xx_enc = enc; p.prob = prob; p.trace = g_trace;
        static const int dbg_env = [] { const char* d = getenv("VLSAT_GAT_DBG"); return d ? atoi(d) : 0; }();   // read once, not per launch
        p.dbg = dbg_env;
        p.n_edges = n_edges; p.H = n_heads; p.hid = hid; p.d_o = d_o;
        const int64_t n_tiles = ceil_div(n_edges * n_heads, GT_ROWS);
        const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, kNumSMs);
        launch_k(kern, dim3(grid), dim3(64 + 128 * wg), smem, st, tk, tkl, t1, t1l, t2, t2l, p);
        ++launches;
    }
    launch_k(gat_finalize_kernel, dim3((unsigned)ceil_div(n_nodes * D_a, 256)), dim3(256), 0, st, enc, xx, ld_xx, n_nodes, n_heads, d_o);
    return finish_launch(launches);
}
