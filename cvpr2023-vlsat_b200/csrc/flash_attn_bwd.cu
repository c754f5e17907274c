// A9 backward on the tensor cores (BF16x3): streaming gradients of softmax(Q K^T / sqrt(dk)) V. The [nq, nk] score and
// probability matrices never leave the SM (round 1 materialised them in 8192-query blocks: 32 of the 35 ms of a
// forward + backward at config #2).
//
// Reference: autograd through ScaledDotProductAttention.forward (src/model/transformer/attention.py:41-78) as called by
// cross_attn_rel (src/model/model_utils/network_MMG.py:231) - no mask, no bias.
//
//   P = exp(scale S - lse),  dP = dO V^T,  dS = scale P o (dP - delta),  delta = rowsum(dO o O)
//   dV = P^T dO,  dK = dS^T Q,  dQ = dS K
//
// One kernel template, two launches (the FlashAttention-2 split: deterministic, no atomics):
//   KV mode  CTA = 128 keys of one head (K_j, V_j resident in TMEM as A operands), streams 64-query tiles:
//            S^T = K_j Q_i^T, dP^T = V_j dO_i^T  ->  P^T, dS^T written IN PLACE over them as bf16 (hi, lo) pairs
//            ->  dV_j += P^T dO_i,  dK_j += dS^T Q_i  (TS MMAs, accumulators live in TMEM for the whole query range)
//   Q mode   CTA = 128 queries of one head (Q_i, dO_i resident in TMEM), streams 64-key tiles:
//            S = Q_i K_j^T, dP = dO_i V_j^T  ->  dS in place  ->  dQ_i += dS K_j
// Every product is a TS MMA (A from tensor memory) whose B operand is a K-major 128B-swizzled [64 x 64] bf16 tile written
// by TMA - exactly the operand forms of the forward kernel (csrc/flash_attn_bf16.cu). The products that reduce over the
// streamed axis read transposed copies (dO^T, Q^T, K^T: [H*64, n]) made once per call by vlsat_bf16_split_t.
//
// TMEM (512 columns): T1[b] (S / P) at 64 b | T2[b] (dP / dS) at 128 + 64 b | acc1 (dV) at 256 | acc2 (dK or dQ) at 320 |
//                     X1 (hi 32 words, lo 32 words) at 384 | X2 at 448.  b = tile parity: the scores of tile i + 1 are
//                     issued before the accumulation of tile i, so the tensor pipe never waits for the conversion warps.
// Warps: 0 = producer of the score operands (Y ring), 1 = MMA issue, 2..17 = conversion (FOUR threads per row, one 16-column
//        quarter of every score tile each: the conversion is a chain of dependent TMEM loads, ex2 and splits, and with
//        two warps per scheduler ncu showed it pacing the kernel - tensor pipe 55 % active, issue 33 %, long-scoreboard
//        stalls; four warps per scheduler hide that latency), 18 = producer of the accumulation operands (Z ring).
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <float.h>
#include <stdlib.h>

namespace vlsat {

using namespace tc;

constexpr int FW_ROWS = 128, FW_T = 64, FW_DK = 64;
constexpr int FW_CONV_WARPS = 16;                         // conversion warps: 4 per TMEM lane quarter
constexpr int FW_THREADS = (3 + FW_CONV_WARPS) * 32;
constexpr int FW_TILE = FW_T * 128;                       // one [64 x 64] bf16 tile: 64 rows of 128 bytes
constexpr int FW_PAIR = 2 * FW_TILE;                      // hi | lo
constexpr int FW_YST = 3, FW_ZST = 3;                     // ring depths
constexpr int FW_NSTAT = FW_YST + 2;                      // per-column statistics slots (KV mode), see the producer
constexpr uint32_t FW_TMEM_COLS = 512;
constexpr float FW_LSE_PAD = 1e30f;                       // lse2 of a padded query: exp2(0 - 1e30) = 0

__device__ __forceinline__ float fw_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ void fw_tmem_st_32(uint32_t taddr, const uint32_t (&r)[32]) {
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
__device__ __forceinline__ void fw_tmem_st_16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void fw_tmem_st_8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void fw_tmem_ld_16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T, bf16 inputs, A read from tensor memory (lane = row, two consecutive K elements per word)
__device__ __forceinline__ void fw_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct FlashBwdArgs {
    // stationary operands (A of the score products), bf16 pairs [n_stat, H*64]: KV mode K | V, Q mode Q | dO
    const uint16_t *x1_hi, *x1_lo, *x2_hi, *x2_lo;
    int64_t ldx1, ldx2;
    // statistics of the QUERY axis, [H, ld_stat] with ld_stat % 64 == 0: lse2 = lse * log2(e) (padding: FW_LSE_PAD), delta (padding 0)
    const float *lse2, *delta;
    int64_t ld_stat;
    float *out1, *out2;                  // KV mode: dV, dK. Q mode: out2 = dQ. fp32 [n_stat, H*64], row stride ldo
    int64_t ldo, split_stride;           // split z writes row r at out + z * split_stride + (r - row_base) * ldo
    int n_stat, n_stream, tiles_per_split;
    int n_heads, item_offset, row_base;  // this launch covers work items (stationary tile, head), heads fastest, from item_offset on
    float scale_log2e, scale;
};

// (hi, lo) truncation split of two fp32 values into packed bf16x2 words (low half = first value): hi = upper 16 bits,
// lo = upper 16 bits of x - hi (exact in fp32); hi + lo misses x by < 2^-16 |x|, towards zero.
__device__ __forceinline__ void fw_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
    const float2 l = fsub2(make_float2(a, b), make_float2(__uint_as_float(ua & 0xffff0000u), __uint_as_float(ub & 0xffff0000u)));
    hi = __byte_perm(ua, ub, 0x7632);
    lo = __byte_perm(__float_as_uint(l.x), __float_as_uint(l.y), 0x7632);
}

// packed bf16x2, round to nearest even, low half = a. One issue slot per pair: the conversion warps are bound by the
// instructions they issue, and the integer form (two adds + a permute) measured slower than this XU-pipe conversion.
__device__ __forceinline__ uint32_t fw_pack_rn(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

template <bool KV, int PASSES>      // PASSES: 3 = BF16x3, 1 = the hi halves alone (single-pass bf16 mode)
__global__ void __launch_bounds__(FW_THREADS, 1)
flash_attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_y1h, const __grid_constant__ CUtensorMap tm_y1l,
                      const __grid_constant__ CUtensorMap tm_y2h, const __grid_constant__ CUtensorMap tm_y2l,
                      const __grid_constant__ CUtensorMap tm_z1h, const __grid_constant__ CUtensorMap tm_z1l,
                      const __grid_constant__ CUtensorMap tm_z2h, const __grid_constant__ CUtensorMap tm_z2l,
                      const FlashBwdArgs a) {
    pdl_launch_dependents();
    constexpr int Y_STAGE = 2 * FW_PAIR;                       // Y1 (hi, lo) | Y2 (hi, lo)
    constexpr int Z_STAGE = KV ? 2 * FW_PAIR : FW_PAIR;        // [Z1 (hi, lo)] | Z2 (hi, lo)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* y_smem = smem;
    uint8_t* z_smem = y_smem + FW_YST * Y_STAGE;
    float* stat_smem = reinterpret_cast<float*>(z_smem + FW_ZST * Z_STAGE);          // [FW_NSTAT][128]: -lse2 x 64 | delta x 64
    uint64_t* bars = reinterpret_cast<uint64_t*>(stat_smem + FW_NSTAT * 128);
    uint64_t* x_full = bars;
    uint64_t* y_full = bars + 1;            uint64_t* y_empty = y_full + FW_YST;
    uint64_t* z_full = y_empty + FW_YST;    uint64_t* z_empty = z_full + FW_ZST;
    uint64_t* t_full = z_empty + FW_ZST;    // [2]
    uint64_t* p_ready = t_full + 2;         // [2]
    uint64_t* stat_full = p_ready + 2;      // [FW_NSTAT]
    uint64_t* acc_done = stat_full + FW_NSTAT;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x + a.item_offset;
    const int head = item % a.n_heads;
    const int r0 = (item / a.n_heads) * FW_ROWS;
    const int n_tiles_all = (a.n_stream + FW_T - 1) / FW_T;
    const int tile_begin = blockIdx.z * a.tiles_per_split;
    const int n_tiles = max(0, min(n_tiles_all, tile_begin + a.tiles_per_split) - tile_begin);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_y1h); prefetch_tmap(&tm_y1l); prefetch_tmap(&tm_y2h); prefetch_tmap(&tm_y2l);
        prefetch_tmap(&tm_z2h); prefetch_tmap(&tm_z2l);
        if (KV) { prefetch_tmap(&tm_z1h); prefetch_tmap(&tm_z1l); }
        mbar_init(x_full, 32 * FW_CONV_WARPS);
        for (int s = 0; s < FW_YST; ++s) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], 1); }
        for (int s = 0; s < FW_ZST; ++s) { mbar_init(&z_full[s], 1); mbar_init(&z_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&p_ready[s], 32 * FW_CONV_WARPS); }
        for (int s = 0; s < FW_NSTAT; ++s) mbar_init(&stat_full[s], 1);
        mbar_init(acc_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_holder, FW_TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();

    if (warp == 0) {
        // ---- score operands: Y1 = Q_i (KV) / K_j (Q mode), Y2 = dO_i / V_j: rows = streamed index, 64 dims = 128 bytes
        for (int i = 0; i < n_tiles; ++i) {
            const int s = i % FW_YST;
            const int c0 = (tile_begin + i) * FW_T;
            mbar_wait(&y_empty[s], ((i / FW_YST) & 1) ^ 1);
            if (KV) {
                // per-column statistics of this query tile. Slot reuse: passing y_empty for tile i means the scores of tile
                // i - 3 have completed, hence were issued, hence p_ready of tile i - 5 had been waited for: the conversion
                // warps are done with slot (i - 5) % 5.
                float* st = stat_smem + (i % FW_NSTAT) * 128;
                const float* src = (lane < 16 ? a.lse2 : a.delta) + (int64_t)head * a.ld_stat + c0 + 4 * (lane & 15);
                float4 sv = __ldg(reinterpret_cast<const float4*>(src));
                if (lane < 16) sv = make_float4(-sv.x, -sv.y, -sv.z, -sv.w);      // -lse2: the addend of the packed FMA below
                *reinterpret_cast<float4*>(st + 4 * lane) = sv;
                __syncwarp();
                if (lane == 0) mbar_arrive(&stat_full[i % FW_NSTAT]);
            }
            if (elect_one()) {
                uint8_t* st = y_smem + s * Y_STAGE;
                mbar_arrive_expect_tx(&y_full[s], PASSES == 3 ? Y_STAGE : Y_STAGE / 2);
                tma_load_2d(st, &tm_y1h, &y_full[s], head * FW_DK, c0);
                if (PASSES == 3) tma_load_2d(st + FW_TILE, &tm_y1l, &y_full[s], head * FW_DK, c0);
                tma_load_2d(st + 2 * FW_TILE, &tm_y2h, &y_full[s], head * FW_DK, c0);
                if (PASSES == 3) tma_load_2d(st + 3 * FW_TILE, &tm_y2l, &y_full[s], head * FW_DK, c0);
            }
            __syncwarp();
        }
    } else if (warp == 2 + FW_CONV_WARPS) {
        // ---- accumulation operands (transposed copies): rows = 64 dims, 64 streamed elements = 128 bytes
        for (int i = 0; i < n_tiles; ++i) {
            const int s = i % FW_ZST;
            const int c0 = (tile_begin + i) * FW_T;
            mbar_wait(&z_empty[s], ((i / FW_ZST) & 1) ^ 1);
            if (elect_one()) {
                uint8_t* st = z_smem + s * Z_STAGE;
                mbar_arrive_expect_tx(&z_full[s], PASSES == 3 ? Z_STAGE : Z_STAGE / 2);
                if (KV) {
                    tma_load_2d(st, &tm_z1h, &z_full[s], c0, head * FW_DK);
                    if (PASSES == 3) tma_load_2d(st + FW_TILE, &tm_z1l, &z_full[s], c0, head * FW_DK);
                    st += FW_PAIR;
                }
                tma_load_2d(st, &tm_z2h, &z_full[s], c0, head * FW_DK);
                if (PASSES == 3) tma_load_2d(st + FW_TILE, &tm_z2l, &z_full[s], c0, head * FW_DK);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc<Kind::BF16>(FW_ROWS, 64);
        constexpr uint64_t LO = (uint64_t)(FW_TILE >> 4);                   // hi -> lo tile, in descriptor units (16 B)
        const uint64_t dy0 = make_sdesc_k128(smem_u32(y_smem));
        const uint64_t dz0 = make_sdesc_k128(smem_u32(z_smem));
        const uint32_t tx1 = tmem_base + 384, tx2 = tmem_base + 448;
        // three MMAs per 16-wide reduction step: A_lo B_hi + A_hi B_lo + A_hi B_hi
        auto mma3 = [&](uint32_t td, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint32_t acc) {
            if (PASSES == 3) {
                fw_mma_ts(td, a_lo, b_hi, idesc, acc);
                fw_mma_ts(td, a_hi, b_hi + LO, idesc, 1);
                fw_mma_ts(td, a_hi, b_hi, idesc, 1);
            } else {
                fw_mma_ts(td, a_hi, b_hi, idesc, acc);
            }
        };
        auto issue_scores = [&](int i) {
            const int b = i & 1, ys = i % FW_YST;
            mbar_wait(&y_full[ys], (i / FW_YST) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dy = dy0 + (uint64_t)(ys * (Y_STAGE >> 4));
                const uint32_t t1 = tmem_base + 64 * b, t2 = tmem_base + 128 + 64 * b;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) mma3(t1, tx1 + 8 * kk, tx1 + 32 + 8 * kk, dy + 2 * kk, kk > 0);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) mma3(t2, tx2 + 8 * kk, tx2 + 32 + 8 * kk, dy + 2 * LO + 2 * kk, kk > 0);
                tc_commit(&t_full[b]);
                tc_commit(&y_empty[ys]);
            }
            __syncwarp();
        };
        mbar_wait(x_full, 0);
        tc_fence_after();
        for (int i = 0; i < 2 && i < n_tiles; ++i) issue_scores(i);
        for (int i = 0; i < n_tiles; ++i) {
            const int b = i & 1, zs = i % FW_ZST;
            mbar_wait(&z_full[zs], (i / FW_ZST) & 1);
            mbar_wait(&p_ready[b], (i >> 1) & 1);               // P / dS of tile i sit in TMEM over T1[b] / T2[b]
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dz = dz0 + (uint64_t)(zs * (Z_STAGE >> 4));
                const uint32_t tp = tmem_base + 64 * b, tds = tmem_base + 128 + 64 * b;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {                // 16 streamed elements per MMA = 8 packed columns
                    const uint32_t off = 16 * kk;                  // quarter kk of the tile: 8 hi words, then its 8 lo words
                    const uint32_t acc = (i > 0 || kk > 0) ? 1u : 0u;
                    if (KV) {
                        mma3(tmem_base + 256, tp + off, tp + off + 8, dz + 2 * kk, acc);                  // dV += P^T dO
                        mma3(tmem_base + 320, tds + off, tds + off + 8, dz + 2 * LO + 2 * kk, acc);       // dK += dS^T Q
                    } else {
                        mma3(tmem_base + 320, tds + off, tds + off + 8, dz + 2 * kk, acc);                // dQ += dS K
                    }
                }
                tc_commit(&z_empty[zs]);
                if (i == n_tiles - 1) tc_commit(acc_done);
            }
            __syncwarp();
            if (i + 2 < n_tiles) issue_scores(i + 2);            // T1[b] / T2[b] are free in pipe order behind the accumulation
        }
    } else {
        const int kq = ((warp - 2) >> 2) & 3;                    // 16-column quarter of every score tile owned by this thread
        const int qd = warp & 3;                                 // TMEM lane quarter this warp may access
        const int row_l = qd * 32 + lane;
        const int row = r0 + row_l;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        // stationary operands into TMEM, once: the four threads of a row copy X1 hi, X1 lo, X2 hi, X2 lo (32 words each)
        {
            uint32_t w[32];
            const bool live = row < a.n_stat && (PASSES == 3 || (kq & 1) == 0);
            const int64_t rr = row < a.n_stat ? row : 0;
            const uint16_t* base = kq == 0 ? a.x1_hi : kq == 1 ? a.x1_lo : kq == 2 ? a.x2_hi : a.x2_lo;
            const uint4* s1 = reinterpret_cast<const uint4*>(base + rr * (kq < 2 ? a.ldx1 : a.ldx2) + head * FW_DK);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint4 t = live ? __ldg(s1 + u) : make_uint4(0u, 0u, 0u, 0u);
                w[4 * u] = t.x; w[4 * u + 1] = t.y; w[4 * u + 2] = t.z; w[4 * u + 3] = t.w;
            }
            fw_tmem_st_32(tmem_base + 384 + lane_off + 32 * kq, w);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(x_full);
        }
        float nlse_r = -FW_LSE_PAD, delta_r = 0.f;               // Q mode: the statistics (-lse2, delta) belong to this thread's row
        if (!KV && row < a.n_stat) {
            nlse_r = -__ldg(a.lse2 + (int64_t)head * a.ld_stat + row);
            delta_r = __ldg(a.delta + (int64_t)head * a.ld_stat + row);
        }
        const float2 scale2 = make_float2(a.scale_log2e, a.scale_log2e);
        for (int i = 0; i < n_tiles; ++i) {
            const int b = i & 1;
            const int c0 = (tile_begin + i) * FW_T + 16 * kq;     // first streamed element of this thread's quarter tile
            const uint32_t t1 = tmem_base + 64 * b + lane_off + 16 * kq;
            const uint32_t t2 = tmem_base + 128 + 64 * b + lane_off + 16 * kq;
            uint32_t s[16], d[16];
            mbar_wait(&t_full[b], (i >> 1) & 1);
            tc_fence_after();
            fw_tmem_ld_16(t1, s);
            fw_tmem_ld_16(t2, d);
            tmem_ld_wait();
            const float* st = stat_smem + (i % FW_NSTAT) * 128 + 16 * kq;
            if (KV) mbar_wait(&stat_full[i % FW_NSTAT], (i / FW_NSTAT) & 1);
            const bool ragged = !KV && (c0 + 16 > a.n_stream);
            uint32_t ph[8], pl[8], dh[8], dl[8];
            // p = 2^(s scale - lse2), dS = p (dP - delta) on packed fp32 pairs (FFMA2 / FADD2 / FMUL2: one issued instruction per
            // two elements - the conversion warps are issue-bound, four of them per scheduler next to the MMA stream)
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {                     // four streamed elements per step
                float2 nl[2], e2[2];
                if (KV) {
                    const float4 lv = *reinterpret_cast<const float4*>(st + 4 * c4);          // broadcast reads
                    const float4 ev = *reinterpret_cast<const float4*>(st + 64 + 4 * c4);
                    nl[0] = make_float2(lv.x, lv.y); nl[1] = make_float2(lv.z, lv.w);
                    e2[0] = make_float2(ev.x, ev.y); e2[1] = make_float2(ev.z, ev.w);
                } else {
                    nl[0] = nl[1] = make_float2(nlse_r, nlse_r);
                    e2[0] = e2[1] = make_float2(delta_r, delta_r);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = 4 * c4 + 2 * h, w = 2 * c4 + h;
                    const float2 x = ffma2(make_float2(__uint_as_float(s[j]), __uint_as_float(s[j + 1])), scale2, nl[h]);
                    float2 p = make_float2(fw_ex2(x.x), fw_ex2(x.y));
                    if (ragged) {
                        if (c0 + j >= a.n_stream) p.x = 0.f;
                        if (c0 + j + 1 >= a.n_stream) p.y = 0.f;
                    }
                    const float2 g = fmul2(p, fsub2(make_float2(__uint_as_float(d[j]), __uint_as_float(d[j + 1])), e2[h]));
                    if (PASSES == 3) {
                        if (KV) fw_split2(p.x, p.y, ph[w], pl[w]);
                        fw_split2(g.x, g.y, dh[w], dl[w]);
                    } else {                                     // single pass: round to nearest (a truncated hi alone is biased)
                        if (KV) ph[w] = fw_pack_rn(p.x, p.y);
                        dh[w] = fw_pack_rn(g.x, g.y);
                    }
                }
            }
            // in place: this thread's 16 fp32 columns become 8 hi words + 8 lo words of the same 16 streamed elements
            if (KV) { fw_tmem_st_8(t1, ph); if (PASSES == 3) fw_tmem_st_8(t1 + 8, pl); }
            fw_tmem_st_8(t2, dh);
            if (PASSES == 3) fw_tmem_st_8(t2 + 8, dl);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_ready[b]);
        }
        // ---- epilogue: each thread stores its row's 16-column quarter of the accumulators
        float* o1 = a.out1 ? a.out1 + (int64_t)blockIdx.z * a.split_stride : nullptr;
        float* o2 = a.out2 + (int64_t)blockIdx.z * a.split_stride;
        uint32_t acc[16];
        if (n_tiles > 0) { mbar_wait(acc_done, 0); tc_fence_after(); }
        if (KV) {
            if (n_tiles > 0) { fw_tmem_ld_16(tmem_base + 256 + lane_off + 16 * kq, acc); tmem_ld_wait(); }
            else {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = 0u;
            }
            if (row < a.n_stat) {
                float* orow = o1 + (int64_t)(row - a.row_base) * a.ldo + head * FW_DK + 16 * kq;
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(orow + j) = make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]),
                                                                       __uint_as_float(acc[j + 2]), __uint_as_float(acc[j + 3]));
            }
        }
        if (n_tiles > 0) { fw_tmem_ld_16(tmem_base + 320 + lane_off + 16 * kq, acc); tmem_ld_wait(); }
        else {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0u;
        }
        if (row < a.n_stat) {
            float* orow = o2 + (int64_t)(row - a.row_base) * a.ldo + head * FW_DK + 16 * kq;
            const float sc = a.scale;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(orow + j) = make_float4(__uint_as_float(acc[j]) * sc, __uint_as_float(acc[j + 1]) * sc,
                                                                   __uint_as_float(acc[j + 2]) * sc, __uint_as_float(acc[j + 3]) * sc);
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, FW_TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------- helper kernels
// fp32 [n, D] -> bf16 (hi, lo) pairs [n, D] (round to nearest even: hi = bf16(x), lo = bf16(x - hi)) and, optionally, the
// transposed pairs [D, ldt] (columns n .. ldt - 1 zero) that the products reducing over the row axis read.
__global__ void bf16_split_t_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int64_t D,
                                    uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int64_t ld_out,
                                    uint16_t* __restrict__ hi_t, uint16_t* __restrict__ lo_t, int64_t ldt) {
    pdl_entry();
    __shared__ uint32_t tile[32][33];
    const int64_t r_base = (int64_t)blockIdx.x * 32, c_base = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;                // 32 x 8
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t r = r_base + ty + 8 * k, c = c_base + tx;
        float v = 0.f;
        if (r < n && c < D) v = __ldg(x + r * ldx + c);
        const uint16_t h = __bfloat16_as_ushort(__float2bfloat16_rn(v));
        const uint16_t l = __bfloat16_as_ushort(__float2bfloat16_rn(v - __uint_as_float((uint32_t)h << 16)));
        if (r < n && c < D) { hi[r * ld_out + c] = h; lo[r * ld_out + c] = l; }
        tile[ty + 8 * k][tx] = (uint32_t)h | ((uint32_t)l << 16);
    }
    if (!hi_t) return;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t c = c_base + ty + 8 * k, r = r_base + tx;
        if (c < D && r < ldt) {
            const uint32_t w = tile[tx][ty + 8 * k];
            hi_t[c * ldt + r] = (uint16_t)(w & 0xffffu);
            lo_t[c * ldt + r] = (uint16_t)(w >> 16);
        }
    }
}

// lse2 = lse log2(e) and delta = rowsum_head(dO o O), both [H, ld_stat] padded (FW_LSE_PAD / 0) to a multiple of 64 queries.
// One warp per query row: lane l covers columns [16 l, 16 l + 16) of the 512, i.e. head l / 4 (dk = 64).
__global__ void flash_bwd_stats_kernel(const float* __restrict__ dout, int64_t lddo, const float* __restrict__ out, int64_t ldo,
                                       const float* __restrict__ lse, int64_t ld_lse, float* __restrict__ lse2,
                                       float* __restrict__ delta, int64_t ld_stat, int64_t nq, int n_heads, int dk) {
    pdl_entry();
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= ld_stat) return;
    if (row >= nq) {
        for (int h = lane; h < n_heads; h += 32) { lse2[(int64_t)h * ld_stat + row] = FW_LSE_PAD; delta[(int64_t)h * ld_stat + row] = 0.f; }
        return;
    }
    const int per = dk / 4;                                      // lanes per head when every lane takes 4 floats per step
    for (int h0 = 0; h0 < n_heads; h0 += 32 / per) {             // 32 / per heads per pass (dk = 64: 2 heads ... generic)
        const int h = h0 + lane / per;
        float acc = 0.f;
        if (h < n_heads) {
            const int64_t col = (int64_t)h * dk + 4 * (lane % per);
            const float4 g = __ldg(reinterpret_cast<const float4*>(dout + row * lddo + col));
            const float4 o = __ldg(reinterpret_cast<const float4*>(out + row * ldo + col));
            acc = g.x * o.x + g.y * o.y + g.z * o.z + g.w * o.w;
        }
        for (int off = per >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (h < n_heads && lane % per == 0) {
            delta[(int64_t)h * ld_stat + row] = acc;
            lse2[(int64_t)h * ld_stat + row] = __ldg(lse + (int64_t)h * ld_lse + row) * 1.4426950408889634f;
        }
    }
}

// out = sum over `splits` slabs (float4 granularity)
__global__ void sum_slabs_kernel(const float* __restrict__ slabs, int64_t slab_stride, int splits, float* __restrict__ out,
                                 int64_t ldo, int64_t rows, int64_t cols4) {
    pdl_entry();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols4) return;
    const int64_t r = idx / cols4, c = (idx % cols4) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(slabs + (int64_t)s * slab_stride + r * (cols4 * 4) + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(out + r * ldo + c) = acc;
}

// ------------------------------------------------------------------------------------------------------ host
// Many slabs of a small output (weight gradients of narrow layers: up to 148 slabs of a few thousand float4): 16 threads share
// one output quad, thread g summing slabs g, g + 16, ... and the 16 partial sums combined in lane order through shared memory -
// a fixed association for a given slab count, so still deterministic; 16 x the loads in flight of the kernel above.
__global__ void __launch_bounds__(256)
sum_slabs_wide_kernel(const float* __restrict__ slabs, int64_t slab_stride, int splits, float* __restrict__ out,
                      int64_t ldo, int64_t rows, int64_t cols4) {
    __shared__ float4 part[16][17];
    pdl_entry();
    const int o = threadIdx.x & 15, g = threadIdx.x >> 4;
    const int64_t idx = (int64_t)blockIdx.x * 16 + o;
    const bool ok = idx < rows * cols4;
    const int64_t r = ok ? idx / cols4 : 0, c = ok ? (idx % cols4) * 4 : 0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok)
        for (int s = g; s < splits; s += 16) {
            const float4 v = *reinterpret_cast<const float4*>(slabs + (int64_t)s * slab_stride + r * (cols4 * 4) + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    part[g][o] = acc;
    __syncthreads();
    if (g == 0 && ok) {
#pragma unroll
        for (int j = 1; j < 16; ++j) { const float4 v = part[j][o]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        *reinterpret_cast<float4*>(out + r * ldo + c) = acc;
    }
}

// out [rows, cols] (row stride ldo) = sum of `splits` slabs of [slab rows >= rows, cols] fp32, slab_stride elements apart
int sum_slabs(const float* slabs, int64_t slab_stride, int splits, float* out, int64_t ldo, int64_t rows, int64_t cols, cudaStream_t st) {
    if (rows == 0 || cols == 0) return VLSAT_OK;
    if (cols % 4 || ldo % 4) return VLSAT_ERR_UNSUPPORTED;
    const int64_t n4 = rows * (cols / 4);
    if (splits >= 16 && n4 <= 16 * 1024)
        launch_k(sum_slabs_wide_kernel, dim3((unsigned)ceil_div(n4, 16)), dim3(256), 0, st, slabs, slab_stride, splits, out, ldo, rows, cols / 4);
    else
        launch_k(sum_slabs_kernel, dim3((unsigned)ceil_div(n4, 256)), dim3(256), 0, st, slabs, slab_stride, splits, out, ldo, rows, cols / 4);
    return finish_launch();
}

// Launch plan of one mode (same scheme as the forward, csrc/flash_attn_bf16.cu): work items = (128-row stationary tile, head).
//   uniform: every item split s ways over CTAs, each split writing its own slab of the WHOLE output, summed afterwards;
//   hybrid:  the first floor(items / 148) * 148 items unsplit in whole rounds, written directly; the leftover items split so
//            that together they fill one more round (config #2: 592 + 8 items x 18 splits instead of 3 x 600 and slab sums
//            of everything) - only the rows of the leftover tiles go through slabs.
struct FwPlan { int items_a, items_b, splits_b, splits_uniform, row_base; bool hybrid; };

static FwPlan fw_plan(int64_t n_stat, int64_t n_stream, int n_heads) {
    const int64_t items = ceil_div(n_stat, FW_ROWS) * n_heads;
    const int64_t n_tiles = ceil_div(n_stream, FW_T);
    FwPlan p{};
    const char* forced = getenv("VLSAT_FLASH_BWD_SPLITS");       // experiments only: forces the uniform scheme
    if (forced && atoi(forced) >= 1 && atoi(forced) <= 8) {
        p.splits_uniform = (int)min((int64_t)atoi(forced), max((int64_t)1, n_tiles));
        return p;
    }
    int best = 1; double best_cost = 1e30;
    for (int s = 1; s <= 8; ++s) {
        if (s > n_tiles) break;
        const double rounds = (double)ceil_div(items * s, kNumSMs);
        // tile-rounds + the per-CTA prologue / epilogue (about 3 tiles' worth) + writing and merging the extra slabs
        const double cost = rounds * ((double)ceil_div(n_tiles, s) + 3.0) + (s > 1 ? 2.0 + 0.02 * n_tiles : 0.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
    }
    p.splits_uniform = best;
    int64_t a = (items / kNumSMs) * kNumSMs;
    a -= a % n_heads;
    const int64_t b = items - a;
    if (a > 0 && b > 0 && b < kNumSMs) {
        const int sb = (int)max((int64_t)1, min(min((int64_t)kNumSMs / b, n_tiles), (int64_t)32));
        const double cost = (double)(a / kNumSMs) * ((double)n_tiles + 3.0) + (double)ceil_div(n_tiles, sb) + 3.0 + 5.0;
        if (sb > 1 && cost < best_cost - 1e-9) {
            p.hybrid = true; p.items_a = (int)a; p.items_b = (int)b; p.splits_b = sb;
            p.row_base = (int)(a / n_heads) * FW_ROWS;
        }
    }
    return p;
}

static size_t fw_plan_slab_floats(const FwPlan& p, int64_t n_stat, size_t D) {        // per output tensor
    if (p.hybrid) return (size_t)p.splits_b * (size_t)(n_stat - p.row_base) * D;
    return p.splits_uniform > 1 ? (size_t)p.splits_uniform * (size_t)n_stat * D : 0;
}

size_t flash_attn_bwd_workspace_bytes(int64_t nq, int64_t nk, int n_heads) {
    const size_t D = (size_t)n_heads * FW_DK;
    const size_t kv = 2 * fw_plan_slab_floats(fw_plan(nk, nq, n_heads), nk, D), q = fw_plan_slab_floats(fw_plan(nq, nk, n_heads), nq, D);
    return max(kv, q) * sizeof(float);
}

int bf16_split_t(const float* x, int64_t ldx, int64_t n, int64_t D, uint16_t* hi, uint16_t* lo, int64_t ld_out,
                 uint16_t* hi_t, uint16_t* lo_t, int64_t ldt, cudaStream_t st) {
    if (n == 0 || D == 0) return VLSAT_OK;
    if (hi_t && ldt < n) return VLSAT_ERR_INVALID_ARG;
    dim3 grid((unsigned)ceil_div(hi_t ? max(n, ldt) : n, 32), (unsigned)ceil_div(D, 32));
    launch_k(bf16_split_t_kernel, grid, dim3(32, 8), 0, st, x, ldx, n, D, hi, lo, ld_out, hi_t, lo_t, ldt);
    return finish_launch();
}

int flash_bwd_stats(const float* dout, int64_t lddo, const float* out, int64_t ldo, const float* lse, int64_t ld_lse, float* lse2,
                    float* delta, int64_t ld_stat, int64_t nq, int n_heads, int dk, cudaStream_t st) {
    if (dk != 64 && dk != 32 && dk != 16 && dk != 128) return VLSAT_ERR_UNSUPPORTED;
    if (ld_stat % 64 || ld_stat < nq || (lddo | ldo) % 4) return VLSAT_ERR_INVALID_ARG;
    if (ld_stat == 0) return VLSAT_OK;
    launch_k(flash_bwd_stats_kernel, dim3((unsigned)ceil_div(ld_stat, 8)), dim3(256), 0, st, dout, lddo, out, ldo, lse, ld_lse, lse2,
             delta, ld_stat, nq, n_heads, dk);
    return finish_launch();
}

// All pair operands bf16 (uint16 storage), row strides in elements (multiples of 8). q/k/v/do: [n, H*64]; qt/kt/dot: [H*64, n]
// (transposed copies). lse2/delta: [H, ld_stat], ld_stat % 64 == 0, padded as flash_bwd_stats writes them.
int flash_attn_bwd_bf16(const uint16_t* q_hi, const uint16_t* q_lo, int64_t ldq, const uint16_t* k_hi, const uint16_t* k_lo, int64_t ldk,
                        const uint16_t* v_hi, const uint16_t* v_lo, int64_t ldv, const uint16_t* do_hi, const uint16_t* do_lo, int64_t lddo,
                        const uint16_t* qt_hi, const uint16_t* qt_lo, int64_t ldqt, const uint16_t* kt_hi, const uint16_t* kt_lo, int64_t ldkt,
                        const uint16_t* dot_hi, const uint16_t* dot_lo, int64_t lddot, const float* lse2, const float* delta, int64_t ld_stat,
                        float* dq, int64_t lddq, float* dk, int64_t lddk, float* dv, int64_t lddv, int64_t nq, int64_t nk, int n_heads,
                        int dk_, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (dk_ != FW_DK || nq >= (1ll << 31) || nk >= (1ll << 31) || nq <= 0 || nk <= 0) return VLSAT_ERR_UNSUPPORTED;
    if ((ldq | ldk | ldv | lddo | ldqt | ldkt | lddot) % 8 || (lddq | lddk | lddv) % 4 || ld_stat % 64 || ld_stat < nq) return VLSAT_ERR_UNSUPPORTED;
    if (lddk != lddv) return VLSAT_ERR_UNSUPPORTED;
    const size_t need = flash_attn_bwd_workspace_bytes(nq, nk, n_heads);
    if (need > 0 && (!workspace || workspace_bytes < need)) return VLSAT_ERR_WORKSPACE;
    const uint64_t D = (uint64_t)n_heads * FW_DK;
    const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUtensorMap tq[2], tk[2], tv[2], tdo[2], tqt[2], tkt[2], tdot[2];
    bool ok = make_tmap_2d(&tq[0], q_hi, BF, 2, nq, D, ldq, 64, FW_T) && make_tmap_2d(&tq[1], q_lo, BF, 2, nq, D, ldq, 64, FW_T) &&
              make_tmap_2d(&tk[0], k_hi, BF, 2, nk, D, ldk, 64, FW_T) && make_tmap_2d(&tk[1], k_lo, BF, 2, nk, D, ldk, 64, FW_T) &&
              make_tmap_2d(&tv[0], v_hi, BF, 2, nk, D, ldv, 64, FW_T) && make_tmap_2d(&tv[1], v_lo, BF, 2, nk, D, ldv, 64, FW_T) &&
              make_tmap_2d(&tdo[0], do_hi, BF, 2, nq, D, lddo, 64, FW_T) && make_tmap_2d(&tdo[1], do_lo, BF, 2, nq, D, lddo, 64, FW_T) &&
              make_tmap_2d(&tqt[0], qt_hi, BF, 2, D, nq, ldqt, 64, FW_DK) && make_tmap_2d(&tqt[1], qt_lo, BF, 2, D, nq, ldqt, 64, FW_DK) &&
              make_tmap_2d(&tkt[0], kt_hi, BF, 2, D, nk, ldkt, 64, FW_DK) && make_tmap_2d(&tkt[1], kt_lo, BF, 2, D, nk, ldkt, 64, FW_DK) &&
              make_tmap_2d(&tdot[0], dot_hi, BF, 2, D, nq, lddot, 64, FW_DK) && make_tmap_2d(&tdot[1], dot_lo, BF, 2, D, nq, lddot, 64, FW_DK);
    if (!ok) return VLSAT_ERR_UNSUPPORTED;
    const float scale = 1.f / sqrtf((float)dk_);
    int launches = 0;
    auto smem_bytes = [](bool kv) {
        return (size_t)FW_YST * 2 * FW_PAIR + (size_t)FW_ZST * (kv ? 2 : 1) * FW_PAIR + FW_NSTAT * 128 * 4 + 256 + 1024;
    };
    // one mode = one or two kernel launches (+ slab sums) following its plan
    auto run_mode = [&](bool kv) {
        const int64_t n_stat = kv ? nk : nq, n_stream = kv ? nq : nk;
        const FwPlan plan = fw_plan(n_stat, n_stream, n_heads);
        const int n_tiles = (int)ceil_div(n_stream, FW_T);
        const int64_t items = ceil_div(n_stat, FW_ROWS) * n_heads;
        float* const dst1 = kv ? dv : nullptr;
        float* const dst2 = kv ? dk : dq;
        const int64_t ld1 = lddv, ld2 = kv ? lddk : lddq;
        const size_t smem = smem_bytes(kv);
        auto launch = [&](int offset, int count, int splits, int row_base) {
            FlashBwdArgs a{};
            if (kv) { a.x1_hi = k_hi; a.x1_lo = k_lo; a.ldx1 = ldk; a.x2_hi = v_hi; a.x2_lo = v_lo; a.ldx2 = ldv; }
            else { a.x1_hi = q_hi; a.x1_lo = q_lo; a.ldx1 = ldq; a.x2_hi = do_hi; a.x2_lo = do_lo; a.ldx2 = lddo; }
            a.lse2 = lse2; a.delta = delta; a.ld_stat = ld_stat;
            a.n_stat = (int)n_stat; a.n_stream = (int)n_stream; a.tiles_per_split = (int)ceil_div(n_tiles, splits);
            a.scale_log2e = 1.4426950408889634f * scale; a.scale = scale;
            a.n_heads = n_heads; a.item_offset = offset; a.row_base = 0;
            const int64_t rows = n_stat - row_base;                       // rows that go through slabs when splits > 1
            if (splits > 1) {
                a.row_base = row_base;
                a.split_stride = rows * (int64_t)D; a.ldo = (int64_t)D;
                a.out1 = kv ? (float*)workspace : nullptr;
                a.out2 = kv ? (float*)workspace + (size_t)splits * rows * D : (float*)workspace;
            } else {
                a.out1 = dst1; a.out2 = dst2; a.ldo = ld2; a.split_stride = 0;   // (lddk == lddv is checked above)
            }
            dim3 grid((unsigned)count, 1u, (unsigned)splits);
            const bool one = tc_passes() == 1;
            if (kv) {
                auto kern = one ? flash_attn_bwd_kernel<true, 1> : flash_attn_bwd_kernel<true, 3>;
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                launch_k(kern, grid, dim3(FW_THREADS), smem, st, tq[0], tq[1], tdo[0], tdo[1], tdot[0], tdot[1], tqt[0], tqt[1], a);
            } else {
                auto kern = one ? flash_attn_bwd_kernel<false, 1> : flash_attn_bwd_kernel<false, 3>;
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                launch_k(kern, grid, dim3(FW_THREADS), smem, st, tk[0], tk[1], tv[0], tv[1], tkt[0], tkt[1], tkt[0], tkt[1], a);
            }
            ++launches;
            if (splits > 1) {
                const int64_t n4 = rows * (int64_t)(D / 4);
                if (kv) {
                    launch_k(sum_slabs_kernel, dim3((unsigned)ceil_div(n4, 256)), dim3(256), 0, st, (const float*)a.out1, a.split_stride, splits,
                             dst1 + (int64_t)row_base * ld1, ld1, rows, (int64_t)(D / 4));
                    ++launches;
                }
                launch_k(sum_slabs_kernel, dim3((unsigned)ceil_div(n4, 256)), dim3(256), 0, st, (const float*)a.out2, a.split_stride, splits,
                         dst2 + (int64_t)row_base * ld2, ld2, rows, (int64_t)(D / 4));
                ++launches;
            }
        };
        if (plan.hybrid) {
            launch(0, plan.items_a, 1, 0);
            launch(plan.items_a, plan.items_b, plan.splits_b, plan.row_base);
        } else {
            launch(0, (int)items, plan.splits_uniform, 0);
        }
    };
    run_mode(true);      // dK, dV: key tiles stationary, queries streamed
    run_mode(false);     // dQ: query tiles stationary, keys streamed
    return finish_launch(launches);
}

}  // namespace vlsat

using namespace vlsat;

extern "C" int vlsat_bf16_split_t(const float* x, int64_t ldx, int64_t rows, int64_t cols, void* hi, void* lo, int64_t ld_out,
                                  void* hi_t, void* lo_t, int64_t ld_t, void* stream) {
    VLSAT_REQUIRE(rows >= 0 && cols >= 0);
    if (rows == 0 || cols == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && hi && lo && ldx >= cols && ld_out >= cols && ((hi_t == nullptr) == (lo_t == nullptr)));
    VLSAT_REQUIRE(!hi_t || ld_t >= rows);
    return bf16_split_t(x, ldx, rows, cols, (uint16_t*)hi, (uint16_t*)lo, ld_out, (uint16_t*)hi_t, (uint16_t*)lo_t, ld_t, (cudaStream_t)stream);
}

extern "C" int vlsat_flash_attn_bwd_stats(const float* dout, int64_t ld_dout, const float* out, int64_t ld_out, const float* lse,
                                          int64_t ld_lse, float* lse2, float* delta, int64_t ld_stat, int64_t nq, int n_heads, int dk,
                                          void* stream) {
    VLSAT_REQUIRE(nq >= 0 && n_heads >= 1 && dk >= 1);
    if (ld_stat == 0) return VLSAT_OK;
    VLSAT_REQUIRE(lse2 && delta && (nq == 0 || (dout && out && lse)) && ld_lse >= nq);
    VLSAT_REQUIRE(ld_dout >= (int64_t)n_heads * dk && ld_out >= (int64_t)n_heads * dk);
    VLSAT_SUPPORT((((uintptr_t)dout | (uintptr_t)out) % 16) == 0);
    return flash_bwd_stats(dout, ld_dout, out, ld_out, lse, ld_lse, lse2, delta, ld_stat, nq, n_heads, dk, (cudaStream_t)stream);
}

extern "C" int vlsat_flash_attn_bf16x3_bwd(const vlsat_flash_bwd_operands* o, float* dq, int64_t ld_dq, float* dk, int64_t ld_dk,
                                           float* dv, int64_t ld_dv, int64_t nq, int64_t nk, int n_heads, int head_dim,
                                           void* workspace, size_t workspace_bytes, void* stream) {
    VLSAT_REQUIRE(o && nq >= 0 && nk >= 0 && n_heads >= 1);
    if (nq == 0 && nk == 0) return VLSAT_OK;
    VLSAT_SUPPORT(nq >= 1 && nk >= 1);                  // an empty side has all-zero gradients: the caller fills them
    VLSAT_REQUIRE(dq && dk && dv && o->lse2 && o->delta);
    const vlsat_bf16_pair* ps[7] = {&o->q, &o->k, &o->v, &o->dout, &o->q_t, &o->k_t, &o->dout_t};
    uintptr_t all = (uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv | (uintptr_t)workspace | (uintptr_t)o->lse2 | (uintptr_t)o->delta;
    for (int i = 0; i < 7; ++i) {
        VLSAT_REQUIRE(ps[i]->hi && ps[i]->lo);
        VLSAT_REQUIRE(ps[i]->ld >= (i < 4 ? (int64_t)n_heads * head_dim : (i == 5 ? nk : nq)));
        all |= (uintptr_t)ps[i]->hi | (uintptr_t)ps[i]->lo;
    }
    VLSAT_SUPPORT(all % 16 == 0);
    const int64_t D = (int64_t)n_heads * head_dim;
    VLSAT_REQUIRE(ld_dq >= D && ld_dk >= D && ld_dv >= D);
    return flash_attn_bwd_bf16((const uint16_t*)o->q.hi, (const uint16_t*)o->q.lo, o->q.ld, (const uint16_t*)o->k.hi, (const uint16_t*)o->k.lo, o->k.ld,
                               (const uint16_t*)o->v.hi, (const uint16_t*)o->v.lo, o->v.ld, (const uint16_t*)o->dout.hi, (const uint16_t*)o->dout.lo, o->dout.ld,
                               (const uint16_t*)o->q_t.hi, (const uint16_t*)o->q_t.lo, o->q_t.ld, (const uint16_t*)o->k_t.hi, (const uint16_t*)o->k_t.lo, o->k_t.ld,
                               (const uint16_t*)o->dout_t.hi, (const uint16_t*)o->dout_t.lo, o->dout_t.ld, o->lse2, o->delta, o->ld_stat,
                               dq, ld_dq, dk, ld_dk, dv, ld_dv, nq, nk, n_heads, head_dim, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" size_t vlsat_flash_attn_bf16x3_bwd_workspace_bytes(int64_t nq, int64_t nk, int n_heads) {
    if (nq <= 0 || nk <= 0) return 0;
    return flash_attn_bwd_workspace_bytes(nq, nk, n_heads);
}
