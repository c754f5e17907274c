// A9 on the tensor cores, BF16x3 variant: streaming softmax(Q K^T / sqrt(dk)) V with every operand split into
// bf16 (hi, lo) pairs and three kind::f16 MMAs per K step (lo*hi + hi*lo + hi*hi) into fp32 TMEM accumulators.
//
// Why not the 3xTF32 split of the projections here: the attention kernel re-reads K and V once per 128-query tile,
// so it is bound by L2 -> SM bandwidth (~6.5 TB/s aggregate, measured); fp32 hi/lo pairs cost 8 B per element,
// bf16 pairs 4 B, and the bf16 MMA retires twice the K per instruction. The bf16x3 product error is 2^-17
// relative per term; through the softmax (64-term dot products scaled by 1/8, convex combination of V) that
// is <= ~1e-5 of the output scale, the same noise floor the fp32-accumulating tensor core already has.
//
// One CTA = 256 queries (two Q tiles) of one head (dk = 64), KV tiles of 64 keys on 2-stage TMA rings, optional KV split.
//   warp 0          TMA producer: per tile K (hi, lo) [64 keys x 64] and (warp 10) V^T (hi, lo) [64 dims x 64 keys],
//                   all K-major rows of exactly 128 bytes (64 bf16) with the 128B swizzle
//   warp 1          tcgen05.mma issue, both products with the A operand in TMEM (TS): S(t) = Q K(t)^T into S[t&1] with
//                   Q copied into TMEM once by the softmax threads, O += P(t) V(t) with P written over S(t)
//   warps 2-17      softmax: per Q tile two warpgroups, each owning one 32-key HALF of every S tile, so a query row is
//                   shared by two threads (they exchange the half-row maximum through smem once per tile, the row sums
//                   once at the end) and every scheduler holds four softmax warps instead of two - the loop is
//                   latency-bound (ncu: 38 % issue, long-scoreboard / fixed-latency stalls), not pipe-bound.
//                   tcgen05.ld S, online softmax in
//                   the exp2 domain, P split to bf16 hi/lo and written back IN PLACE over the S columns of TMEM (64 fp32
//                   columns -> 32 packed hi + 32 packed lo words): P.V is a TS MMA (A from TMEM) that accumulates into
//                   the warpgroup's O columns for the whole KV range. With S double-buffered the softmax of tile i+1
//                   never waits for P.V of tile i (the profile of the smem-P version showed the softmax warps stalled on
//                   exactly that barrier for 28 % of their samples).
#include "common.cuh"
#include "tc_common.cuh"
#include <algorithm>
#include <float.h>
#include <stdlib.h>

namespace vlsat {

using namespace tc;

constexpr int FB_BQ = 128, FB_BKV = 64, FB_DK = 64;
constexpr int FB_THREADS = 64 + 512 + 32;                // K producer, MMA, 16 softmax warps, V producer (warp 18)
constexpr int FB_STAGES = 4;                             // K and V^T tile rings (Q and P live in TMEM: shared memory is all theirs)
constexpr int FB_K_STAGE = 2 * FB_BKV * 128;             // K_hi | K_lo, 64 rows x 128 B
constexpr int FB_V_STAGE = 2 * FB_DK * 128;              // Vt_hi | Vt_lo
constexpr uint32_t FB_TMEM_COLS = 512;                   // S[g][b] x 64 at 64*(2g+b) | O[g] x 64 at 256 + 64 g | Q[g] (hi, lo) x 64 at 384 + 64 g

__device__ __forceinline__ float fb_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// two floats -> packed bf16x2 (round to nearest even); low half = first argument
__device__ __forceinline__ uint32_t fb_pack_bf16(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

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

__device__ __forceinline__ void tmem_st_16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// D[tmem] (+)= A[tmem] . B[smem]^T, bf16 inputs: A is read from tensor memory (lane = row, two consecutive K elements per word)
__device__ __forceinline__ void mma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void bf16_split_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int64_t cols,
                                  uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int64_t ld_out) {
    pdl_entry();
    const int64_t c4 = (cols + 3) >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * c4) return;
    const int64_t r = idx / c4, c = (idx % c4) * 4;
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (c + i < cols) ? __ldg(x + r * ldx + c + i) : 0.f;
    uint32_t h01 = fb_pack_bf16(v[0], v[1]), h23 = fb_pack_bf16(v[2], v[3]);
    const float l0 = v[0] - __uint_as_float(h01 << 16), l1 = v[1] - __uint_as_float(h01 & 0xffff0000u);
    const float l2 = v[2] - __uint_as_float(h23 << 16), l3 = v[3] - __uint_as_float(h23 & 0xffff0000u);
    const uint32_t q01 = fb_pack_bf16(l0, l1), q23 = fb_pack_bf16(l2, l3);
    if (c + 3 < ld_out) {
        *reinterpret_cast<uint2*>(hi + r * ld_out + c) = make_uint2(h01, h23);
        *reinterpret_cast<uint2*>(lo + r * ld_out + c) = make_uint2(q01, q23);
    }
}

// Partial softmax state of one KV split: o (un-normalised, relative to m), m (running max, exp2 domain), l (sum).
struct FlashPartial {
    float* o;          // [splits, rows, H*64]   rows = nq - row_base: the query rows this (split) launch covers
    float* m;          // [splits, H, rows]
    float* l;          // [splits, H, rows]
    int rows, row_base;
};

// One CTA = 256 queries (two 128-row Q tiles, one per softmax warpgroup) of one head over the KV tiles
// [tile_begin, tile_end) of split blockIdx.z: every K / V tile fetched from L2 serves both Q tiles, which halves
// the L2 -> SM traffic that bounds this kernel.
template <int PASSES>       // 3 = BF16x3 (lo*hi + hi*lo + hi*hi), 1 = the hi halves alone (single-pass bf16 mode)
__global__ void __launch_bounds__(FB_THREADS, 1)
flash_attn_bf16_kernel(const uint16_t* __restrict__ q_hi, const uint16_t* __restrict__ q_lo, int64_t ldq,
                       const __grid_constant__ CUtensorMap tm_khi, const __grid_constant__ CUtensorMap tm_klo,
                       const __grid_constant__ CUtensorMap tm_vhi, const __grid_constant__ CUtensorMap tm_vlo,
                       float* __restrict__ out, int64_t ldo, float* __restrict__ lse, FlashPartial part,
                       int nq, int nk, int n_heads, int tiles_per_split, float scale_log2e, int item_offset) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array keeps LDS/STS
    uint8_t* k_smem = smem;                                  // 2 stages x (K_hi | K_lo)
    uint8_t* v_smem = k_smem + FB_STAGES * FB_K_STAGE;       // FB_STAGES x (Vt_hi | Vt_lo)
    uint64_t* bars = reinterpret_cast<uint64_t*>(v_smem + FB_STAGES * FB_V_STAGE);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1; uint64_t* k_empty = k_full + FB_STAGES;
    uint64_t* v_full = k_empty + FB_STAGES; uint64_t* v_empty = v_full + FB_STAGES;
    uint64_t* s_full = v_empty + FB_STAGES;   // [g * 2 + b]
    uint64_t* p_ready = s_full + 4;           // [g]
    uint64_t* pv_full = p_ready + 2;          // [g]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(pv_full + 2);
    float* xch = reinterpret_cast<float*>(tmem_holder + 4);  // [parity 2][Q tile 2][key half 2][128 rows] half-row maxima / sums

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work item = (256-query block, head), heads fastest: launches cover a contiguous range of items starting at item_offset
    const int item = blockIdx.x + item_offset;
    const int head = item % n_heads;
    const int q0 = (item / n_heads) * 2 * FB_BQ;
    const int n_tiles_all = (nk + FB_BKV - 1) / FB_BKV;
    const int tile_begin = blockIdx.z * tiles_per_split;
    const int n_tiles = max(0, min(n_tiles_all, tile_begin + tiles_per_split) - tile_begin);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_khi);
        prefetch_tmap(&tm_klo); prefetch_tmap(&tm_vhi); prefetch_tmap(&tm_vlo);
        mbar_init(q_full, 512);                                  // every softmax thread has put its half of a Q row into TMEM
        for (int s = 0; s < FB_STAGES; ++s) {
            mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1); mbar_init(&s_full[2 + s], 1); mbar_init(&p_ready[s], 256); mbar_init(&pv_full[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_holder, FB_TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();

    if (warp == 0) {
        for (int i = 0; i < n_tiles; ++i) {
            const int s = i % FB_STAGES;
            mbar_wait(&k_empty[s], ((i / FB_STAGES) & 1) ^ 1);
            if (elect_one()) {
                uint8_t* st = k_smem + s * FB_K_STAGE;
                mbar_arrive_expect_tx(&k_full[s], PASSES == 3 ? FB_K_STAGE : FB_K_STAGE / 2);
                tma_load_2d(st, &tm_khi, &k_full[s], head * FB_DK, (tile_begin + i) * FB_BKV);          // rows = keys
                if (PASSES == 3) tma_load_2d(st + FB_BKV * 128, &tm_klo, &k_full[s], head * FB_DK, (tile_begin + i) * FB_BKV);
            }
            __syncwarp();
        }
    } else if (warp == 18) {
        // V producer on its own warp: S tiles are issued ahead of the P.V products, so one in-order producer would
        // stall the K ring behind the V ring while the MMA warp waits for K
        for (int i = 0; i < n_tiles; ++i) {
            const int s = i % FB_STAGES;
            mbar_wait(&v_empty[s], ((i / FB_STAGES) & 1) ^ 1);
            if (elect_one()) {
                uint8_t* st = v_smem + s * FB_V_STAGE;
                mbar_arrive_expect_tx(&v_full[s], PASSES == 3 ? FB_V_STAGE : FB_V_STAGE / 2);
                tma_load_2d(st, &tm_vhi, &v_full[s], (tile_begin + i) * FB_BKV, head * FB_DK);          // rows = dims, cols = keys
                if (PASSES == 3) tma_load_2d(st + FB_DK * 128, &tm_vlo, &v_full[s], (tile_begin + i) * FB_BKV, head * FB_DK);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc<Kind::BF16>(FB_BQ, 64);
        const uint64_t dk0 = make_sdesc_k128(smem_u32(k_smem));
        const uint64_t dv0 = make_sdesc_k128(smem_u32(v_smem));
        // S(g, i) = Q_g K(i)^T for one Q tile; the K stage is released after the second one
        auto issue_s = [&](int i, int g) {
            const int s = i & 1, ks = i % FB_STAGES;             // S buffer parity, K ring stage
            if (g == 0) { mbar_wait(&k_full[ks], (i / FB_STAGES) & 1); tc_fence_after(); }
            if (elect_one()) {
                const uint64_t dk = dk0 + (uint64_t)(ks * (FB_K_STAGE >> 4));
                const uint32_t tq = tmem_base + 384 + 64 * g;               // Q_g: hi words in columns [0, 32), lo in [32, 64)
                const uint32_t ts = tmem_base + 64 * (2 * g + s);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {                 // 16 dims per MMA = 8 packed columns
                    if (PASSES == 3) {
                        mma_ts_bf16(ts, tq + 32 + 8 * kk, dk + 2 * kk, idesc, kk > 0);                             // Q_lo K_hi
                        mma_ts_bf16(ts, tq + 8 * kk, dk + ((FB_BKV * 128) >> 4) + 2 * kk, idesc, 1);               // Q_hi K_lo
                    }
                    mma_ts_bf16(ts, tq + 8 * kk, dk + 2 * kk, idesc, PASSES == 3 ? 1u : (kk > 0 ? 1u : 0u));       // Q_hi K_hi
                }
                tc_commit(&s_full[2 * g + s]);
                if (g == 1) tc_commit(&k_empty[ks]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        tc_fence_after();
        for (int i = 0; i < 2 && i < n_tiles; ++i) { issue_s(i, 0); issue_s(i, 1); }
        for (int i = 0; i < n_tiles; ++i) {
            const int s = i & 1, vs = i % FB_STAGES;             // S / P buffer parity, V ring stage
            mbar_wait(&v_full[vs], (i / FB_STAGES) & 1);
            for (int g = 0; g < 2; ++g) {
                mbar_wait(&p_ready[g], i & 1);                   // P_g(i) is in TMEM (over S_g[s]); O_g is rescaled if needed
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t dv = dv0 + (uint64_t)(vs * (FB_V_STAGE >> 4));
                    const uint32_t tp = tmem_base + 64 * (2 * g + s);       // P_g(i): hi words in columns [0, 32), lo in [32, 64)
                    const uint32_t tpv = tmem_base + 256 + 64 * g;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {             // 16 keys per MMA = 8 packed columns
                        if (PASSES == 3) {
                            mma_ts_bf16(tpv, tp + 32 + 8 * kk, dv + 2 * kk, idesc, (kk > 0) || (i > 0));              // P_lo V_hi
                            mma_ts_bf16(tpv, tp + 8 * kk, dv + ((FB_DK * 128) >> 4) + 2 * kk, idesc, 1);               // P_hi V_lo
                        }
                        mma_ts_bf16(tpv, tp + 8 * kk, dv + 2 * kk, idesc, PASSES == 3 ? 1u : ((kk > 0 || i > 0) ? 1u : 0u));   // P_hi V_hi
                    }
                    tc_commit(&pv_full[g]);
                    if (g == 1) tc_commit(&v_empty[vs]);
                }
                __syncwarp();
                // S_g(i + 2) right behind P_g(i).V(i): its buffer (the one P_g(i) sits in) is free in pipe order, and the
                // Q tile that finished first gets its next scores without waiting for the other one
                if (i + 2 < n_tiles) issue_s(i + 2, g);
            }
        }
    } else {
        const int g = (warp - 2) >> 3;                           // Q tile
        const int kh = ((warp - 2) >> 2) & 1;                    // key half of every S tile owned by this thread
        const int qd = warp & 3;
        const int row_l = qd * 32 + lane;
        const int row = q0 + g * FB_BQ + row_l;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const uint32_t t_pv = tmem_base + 256 + 64 * g;
        auto pair_sync = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory"); };    // the two halves of Q tile g
        // The output accumulator O_g stays in TMEM: the P.V products of all tiles accumulate there (tensor-core fp32
        // accumulation). The running maximum is LAZY: it only moves (and O, l are rescaled through a TMEM
        // load / multiply / store) when some row of the warp exceeds it by more than 2^8; until then p = 2^(s - m_stale)
        // <= 256, exactly representable in the bf16 pair, and l accumulates with the same stale reference.
        float m_run = -FLT_MAX, l_run = 0.f;                     // l_run: this thread's half of the row sum
        constexpr float kRescaleAbove = 8.f;                     // log2 units
        // Q_g lives in TMEM as the A operand of S = Q K^T (TS MMA): an SS MMA of this shape reads 4 KB of Q and 2 KB of K
        // from shared memory for 32 clocks of math - 48 clocks at 128 B/clk, which is what the S products were paced by.
        // The two threads of a row copy its hi (key half 0) and lo (key half 1) words from global memory, once.
        {
            uint32_t qw[32];
            const uint4* src = reinterpret_cast<const uint4*>((kh ? q_lo : q_hi) + (int64_t)min(row, nq - 1) * ldq + head * FB_DK);
#pragma unroll
            for (int u = 0; u < 8; ++u) { const uint4 t = __ldg(src + u); qw[4 * u] = t.x; qw[4 * u + 1] = t.y; qw[4 * u + 2] = t.z; qw[4 * u + 3] = t.w; }
            tmem_st_32(tmem_base + 384 + 64 * g + lane_off + 32 * kh, qw);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(q_full);
        }
        for (int i = 0; i < n_tiles; ++i) {
            const int k0 = (tile_begin + i) * FB_BKV + 32 * kh;  // first key of this thread's half tile
            uint32_t r[32];
            const uint32_t t_s = tmem_base + 64 * (2 * g + (i & 1));
            mbar_wait(&s_full[2 * g + (i & 1)], (i >> 1) & 1);
            tc_fence_after();
            tmem_ld_32x32(t_s + lane_off + 32 * kh, r);
            tmem_ld_wait();
            tc_fence_before();
            if (k0 + 32 > nk) {                                  // ragged last tile
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (k0 + j >= nk) r[j] = __float_as_uint(-FLT_MAX);
            }
            // The loop is bound by the instructions its warps issue (four softmax warps per scheduler against one MMA stream), so
            // it is written for the 3-input FMNMX3 and the packed FFMA2 / FADD2 forms: half the issue slots of the scalar code
            // for the maximum, the scaling and the row sum. Independent partial maxima / sums keep the dependent chains short.
            float mxp[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
#pragma unroll
            for (int j = 0; j < 32; j += 2) mxp[(j >> 1) & 3] = fmaxf(mxp[(j >> 1) & 3], fmaxf(__uint_as_float(r[j]), __uint_as_float(r[j + 1])));
            const float mx_half = fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3]));
            // the row maximum is the larger of the two halves: exchanged through smem (double-buffered by tile parity)
            float* xm = xch + (((i & 1) * 2 + g) * 2) * 128;
            xm[kh * 128 + row_l] = mx_half;
            pair_sync();
            const float mx = fmaxf(mx_half, xm[(kh ^ 1) * 128 + row_l]);
            const float m_cand = fmaxf(m_run, mx * scale_log2e);
            // warp-uniform (TMEM ops are collective) and identical in the partner warp, which sees the same 32 rows
            const bool rescale = __any_sync(0xffffffffu, m_cand > m_run + kRescaleAbove);
            const float m_new = rescale ? m_cand : m_run;
            const float2 scale2 = make_float2(scale_log2e, scale_log2e), neg_m2 = make_float2(-m_new, -m_new);
            float2 rsp[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            // p = 2^(s * scale - m) -> bf16 (hi, lo) pair, packed two keys per word, written over this row's S columns
            // (hi words of the tile's 64 keys in columns [0, 32), lo words in [32, 64); this thread owns 16 of each)
            uint32_t ph[16], pl[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {                       // two keys per packed word
                const float2 x = ffma2(make_float2(__uint_as_float(r[2 * c]), __uint_as_float(r[2 * c + 1])), scale2, neg_m2);
                const float2 p = make_float2(fb_ex2(x.x), fb_ex2(x.y));
                rsp[c & 3] = fadd2(rsp[c & 3], p);
                if (PASSES == 3) {
                    // bf16 pair by TRUNCATION: hi = upper 16 bits of p (the byte permute takes them straight from the fp32 bit
                    // patterns), lo = upper 16 bits of p - hi (exact in fp32). hi + lo misses p by less than 2^-16 p, always
                    // from below: a -8e-6 relative bias of the weights against the exactly summed l, two orders inside the
                    // parity budget.
                    const uint32_t ua = __float_as_uint(p.x), ub = __float_as_uint(p.y);
                    const float2 l = fsub2(p, make_float2(__uint_as_float(ua & 0xffff0000u), __uint_as_float(ub & 0xffff0000u)));
                    ph[c] = __byte_perm(ua, ub, 0x7632);         // {pb.hi16, pa.hi16}: low half = even key
                    pl[c] = __byte_perm(__float_as_uint(l.x), __float_as_uint(l.y), 0x7632);
                } else {
                    ph[c] = fb_pack_bf16(p.x, p.y);              // single pass: round to nearest (a truncated hi alone is biased by 2^-9)
                }
            }
            tmem_st_16(t_s + lane_off + 16 * kh, ph);
            if (PASSES == 3) tmem_st_16(t_s + lane_off + 32 + 16 * kh, pl);
            tmem_st_wait();
            if (i > 0) {
                // P.V(i-1) was issued a whole tile ago: this wait is normally free. It orders the (rare) rescale of O_g
                // between P.V(i-1) and P.V(i), and keeps this thread in step with the barrier's phases.
                mbar_wait(&pv_full[g], (i - 1) & 1);
                tc_fence_after();
                if (rescale) {                                   // each half rescales its 32 output columns
                    const float corr = fb_ex2(m_run - m_new);
                    uint32_t a[32];
                    tmem_ld_32x32(t_pv + lane_off + 32 * kh, a);
                    tmem_ld_wait();
#pragma unroll
                    for (int d = 0; d < 32; ++d) a[d] = __float_as_uint(__uint_as_float(a[d]) * corr);
                    tmem_st_32(t_pv + lane_off + 32 * kh, a);
                    tmem_st_wait();
                    l_run *= corr;
                }
            }
            tc_fence_before();
            mbar_arrive(&p_ready[g]);
            {
                const float2 t = fadd2(fadd2(rsp[0], rsp[1]), fadd2(rsp[2], rsp[3]));
                l_run += t.x + t.y;
            }
            m_run = m_new;
        }
        // row sum = the two half sums (same reference maximum); each thread then finishes its 32 output dims
        float* xl = xch + ((n_tiles & 1) * 2 + g) * 2 * 128;
        xl[kh * 128 + row_l] = l_run;
        pair_sync();
        l_run += xl[(kh ^ 1) * 128 + row_l];
        float o[32];
        if (n_tiles > 0) {
            mbar_wait(&pv_full[g], (n_tiles - 1) & 1);
            tc_fence_after();
            uint32_t a[32];
            tmem_ld_32x32(t_pv + lane_off + 32 * kh, a);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < 32; ++d) o[d] = __uint_as_float(a[d]);
            tc_fence_before();
        } else {
#pragma unroll
            for (int d = 0; d < 32; ++d) o[d] = 0.f;
        }
        if (row < nq) {
            if (gridDim.z == 1) {
                const float inv = 1.f / l_run;
                float* orow = out + (int64_t)row * ldo + head * FB_DK + 32 * kh;
#pragma unroll
                for (int d = 0; d < 32; d += 4)
                    *reinterpret_cast<float4*>(orow + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
                if (lse && kh == 0) lse[(int64_t)head * nq + row] = (m_run + log2f(l_run)) * 0.6931471805599453f;
            } else {
                const int64_t D = (int64_t)n_heads * FB_DK;
                const int64_t pr = row - part.row_base;
                float* orow = part.o + ((int64_t)blockIdx.z * part.rows + pr) * D + head * FB_DK + 32 * kh;
#pragma unroll
                for (int d = 0; d < 32; d += 4)
                    *reinterpret_cast<float4*>(orow + d) = make_float4(o[d], o[d + 1], o[d + 2], o[d + 3]);
                if (kh == 0) {
                    part.m[((int64_t)blockIdx.z * n_heads + head) * part.rows + pr] = m_run;
                    part.l[((int64_t)blockIdx.z * n_heads + head) * part.rows + pr] = l_run;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, FB_TMEM_COLS);
}

// Combine the KV splits of the rows [row_base, nq): out = sum_s o_s 2^(m_s - m) / sum_s l_s 2^(m_s - m)
__global__ void flash_merge_kernel(FlashPartial part, int splits, int nq, int n_heads, float* __restrict__ out, int64_t ldo,
                                   float* __restrict__ lse) {
    pdl_entry();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // (row, head, 4-dim group)
    const int groups = FB_DK / 4;
    if (idx >= (int64_t)part.rows * n_heads * groups) return;
    const int gidx = (int)(idx % groups);
    const int head = (int)((idx / groups) % n_heads);
    const int64_t pr = idx / (groups * n_heads);
    const int64_t q = pr + part.row_base;
    if (q >= nq) return;
    const int64_t D = (int64_t)n_heads * FB_DK;
    float m = -FLT_MAX;
    for (int s = 0; s < splits; ++s) m = fmaxf(m, part.m[((int64_t)s * n_heads + head) * part.rows + pr]);
    float l = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
        const float w = exp2f(part.m[((int64_t)s * n_heads + head) * part.rows + pr] - m);
        l += part.l[((int64_t)s * n_heads + head) * part.rows + pr] * w;
        const float4 v = *reinterpret_cast<const float4*>(part.o + ((int64_t)s * part.rows + pr) * D + head * FB_DK + gidx * 4);
        acc.x += v.x * w; acc.y += v.y * w; acc.z += v.z * w; acc.w += v.w * w;
    }
    const float inv = 1.f / l;
    *reinterpret_cast<float4*>(out + q * ldo + head * FB_DK + gidx * 4) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    if (lse && gidx == 0) lse[(int64_t)head * nq + q] = (m + log2f(l)) * 0.6931471805599453f;
}

int bf16_split(const float* x, int64_t ldx, int64_t rows, int64_t cols, uint16_t* hi, uint16_t* lo, int64_t ld_out, cudaStream_t st) {
    const int64_t n = rows * ((cols + 3) / 4);
    if (n == 0) return VLSAT_OK;
    launch_k(bf16_split_kernel, dim3((unsigned)ceil_div(n, 256)), dim3(256), 0, st, x, ldx, rows, cols, hi, lo, ld_out);
    return finish_launch();
}

// Launch plan. Work items = (256-query block, head); every item costs n_tiles key tiles; one CTA per SM.
//   uniform: every item split s ways over CTAs (s = 1: no merge) - round 1's scheme; at config #2 (304 items, 150 tiles) the best
//            is s = 3: 7 rounds x 50 tiles = 350 tile-times + a merge of everything, against an ideal of 308;
//   hybrid:  the first floor(items / 148) * 148 items run UNSPLIT in whole rounds (no partial state, written directly), the
//            leftover items are split so that together they fill one more round (config #2: 296 + 8 items x 18 splits:
//            300 + 9 tile-times, and only the last 128 query rows go through the merge).
struct FlashPlan { int items_a, items_b, splits_b, splits_uniform; bool hybrid; int row_base; };

static FlashPlan flash_plan(int64_t nq, int64_t nk, int n_heads) {
    const int64_t items = ceil_div(nq, 2 * FB_BQ) * n_heads;
    const int64_t n_tiles = ceil_div(nk, FB_BKV);
    FlashPlan p{};
    int best = 1; double best_cost = 1e30;
    const char* forced = getenv("VLSAT_FLASH_SPLITS");           // experiments only: forces the uniform scheme
    for (int s = 1; s <= 4; ++s) {
        if (s > n_tiles) break;
        const double rounds = (double)ceil_div(items * s, kNumSMs);
        // tile-rounds of the main kernel + the partial-state write/merge traffic of every extra split
        const double cost = rounds * (double)ceil_div(n_tiles, s) + (s - 1) * (4.0 + 0.03 * n_tiles);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
    }
    if (forced && atoi(forced) >= 1 && atoi(forced) <= 4 && atoi(forced) <= n_tiles) best = atoi(forced);
    p.splits_uniform = best;
    // hybrid: whole rounds of unsplit items (a multiple of the head count so that launch A ends on a query-block boundary)
    int64_t a = (items / kNumSMs) * kNumSMs;
    a -= a % n_heads;
    const int64_t b = items - a;
    if (!forced && a > 0 && b > 0 && b < kNumSMs) {
        const int sb = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(kNumSMs / b, n_tiles), 32));
        // launch B: its CTAs pay the prologue (Q into TMEM, pipeline fill: ~4 tiles' worth) + a second launch + a small merge
        const double cost = (double)(a / kNumSMs) * n_tiles + (double)ceil_div(n_tiles, sb) + 4.0 + 6.0;
        if (sb > 1 && cost < best_cost - 1e-9) {
            p.hybrid = true; p.items_a = (int)a; p.items_b = (int)b; p.splits_b = sb;
            p.row_base = (int)(a / n_heads) * 2 * FB_BQ;
        }
    }
    return p;
}

size_t flash_attn_bf16_workspace_bytes(int64_t nq, int64_t nk, int n_heads, int* splits_out) {
    const FlashPlan p = flash_plan(nq, nk, n_heads);
    if (splits_out) *splits_out = p.hybrid ? p.splits_b : p.splits_uniform;
    const size_t per_row = ((size_t)n_heads * FB_DK + 2 * (size_t)n_heads) * sizeof(float);
    if (p.hybrid) return (size_t)p.splits_b * (size_t)(nq - p.row_base) * per_row;
    if (p.splits_uniform == 1) return 0;
    return (size_t)p.splits_uniform * (size_t)nq * per_row;
}

// q_* [nq, H*64] bf16 (row stride ldq elements), k_* [nk, H*64], vt_* [H*64, nk] (row stride ldvt, multiple of 8)
int flash_attn_bf16(const uint16_t* q_hi, const uint16_t* q_lo, int64_t ldq, const uint16_t* k_hi, const uint16_t* k_lo, int64_t ldk,
                    const uint16_t* vt_hi, const uint16_t* vt_lo, int64_t ldvt, float* out, int64_t ldo, float* lse,
                    int64_t nq, int64_t nk, int n_heads, int dk, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (dk != FB_DK || nq >= (1ll << 31) || nk >= (1ll << 31)) return VLSAT_ERR_UNSUPPORTED;
    if ((ldq | ldk | ldvt) % 8 || ldo % 4) return VLSAT_ERR_UNSUPPORTED;
    const FlashPlan plan = flash_plan(nq, nk, n_heads);
    const size_t need = flash_attn_bf16_workspace_bytes(nq, nk, n_heads, nullptr);
    if (need > 0 && (!workspace || workspace_bytes < need)) return VLSAT_ERR_WORKSPACE;
    CUtensorMap tk, tkl, tv, tvl;
    const uint64_t d = (uint64_t)n_heads * dk;
    const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    bool ok =
              make_tmap_2d(&tk, k_hi, BF, 2, nk, d, ldk, 64, FB_BKV) && make_tmap_2d(&tkl, k_lo, BF, 2, nk, d, ldk, 64, FB_BKV) &&
              make_tmap_2d(&tv, vt_hi, BF, 2, d, nk, ldvt, 64, FB_DK) && make_tmap_2d(&tvl, vt_lo, BF, 2, d, nk, ldvt, 64, FB_DK);
    if (!ok) return VLSAT_ERR_UNSUPPORTED;
    const int n_tiles = (int)ceil_div(nk, FB_BKV);
    const int64_t items = ceil_div(nq, 2 * FB_BQ) * n_heads;
    const size_t smem = FB_STAGES * FB_K_STAGE + FB_STAGES * FB_V_STAGE + 1024 + 256 + 2 * 2 * 2 * 128 * 4;
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)dk);
    const bool one_pass = tc_passes() == 1;
    cudaFuncSetAttribute(flash_attn_bf16_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(flash_attn_bf16_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int launches = 0;
    // one launch over items [offset, offset + count) with `splits` key ranges each; partial state for rows >= row_base
    auto run = [&](int offset, int count, int splits, int row_base) {
        FlashPartial part{nullptr, nullptr, nullptr, 0, row_base};
        if (splits > 1) {
            part.rows = (int)(nq - row_base);
            part.o = (float*)workspace;
            part.m = part.o + (size_t)splits * part.rows * d;
            part.l = part.m + (size_t)splits * n_heads * part.rows;
        }
        const int tiles_per_split = (int)ceil_div(n_tiles, splits);
        dim3 grid((unsigned)count, 1u, (unsigned)splits);
        if (one_pass)
            launch_k(flash_attn_bf16_kernel<1>, grid, dim3(FB_THREADS), smem, st, q_hi, q_lo, ldq, tk, tkl, tv, tvl, out, ldo, lse, part, (int)nq, (int)nk,
                     n_heads, tiles_per_split, scale_log2e, offset);
        else
            launch_k(flash_attn_bf16_kernel<3>, grid, dim3(FB_THREADS), smem, st, q_hi, q_lo, ldq, tk, tkl, tv, tvl, out, ldo, lse, part, (int)nq, (int)nk,
                     n_heads, tiles_per_split, scale_log2e, offset);
        ++launches;
        if (splits > 1) {
            const int64_t n = (int64_t)part.rows * n_heads * (FB_DK / 4);
            launch_k(flash_merge_kernel, dim3((unsigned)ceil_div(n, 256)), dim3(256), 0, st, part, splits, (int)nq, n_heads, out, ldo, lse);
            ++launches;
        }
    };
    if (plan.hybrid) {
        run(0, plan.items_a, 1, 0);
        run(plan.items_a, plan.items_b, plan.splits_b, plan.row_base);
    } else {
        run(0, (int)items, plan.splits_uniform, 0);
    }
    return finish_launch(launches);
}

}  // namespace vlsat
