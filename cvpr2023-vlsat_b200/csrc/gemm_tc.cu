// Dense projection on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32 in / fp32 out.
//
// fp32 parity on tensor cores: every operand is split once into  hi = rna_tf32(x)  and  lo = x - hi
// (exact in fp32; |lo| <= 2^-12 |x|) and each K step issues three kind::tf32 MMAs
//     D += A_hi B_hi + A_hi B_lo + A_lo B_hi
// into the same fp32 TMEM accumulator ("3xTF32"). The dropped lo*lo term and the hardware's tf32
// conversion of lo are both <= 2^-23 relative, i.e. fp32-level, so results agree with the FFMA engine
// to ~1e-6 while running on the tensor pipe.
//
// BF16x3 variant (KD = Kind::BF16, the default engine): operands are bf16 pairs  hi = bf16(x), lo = bf16(x - hi)  (same
// 4 bytes per element as fp32) and the three MMAs are kind::f16 with bf16 inputs: twice the tensor rate and twice the
// K per 128-byte smem row of the tf32 form; |x - hi - lo| <= 2^-17 |x| and the dropped lo*lo term is <= 2^-18, so one
// product is accurate to ~1e-5 relative (fp32 accumulation) - two orders inside the 1e-3 parity budget.
//
// Kernel shape: one 128 x BN output tile per CTA, K blocks of one 128-byte swizzle row (32 tf32 / 64 bf16).
//   warp 0      : TMA producer (4 operand tiles per stage: A_hi, A_lo, B_hi, B_lo) on an mbarrier ring
//   warp 1      : TMEM allocation + single-thread tcgen05.mma issue, tcgen05.commit releases the stage
//   warps 2..9  : epilogue, two warpgroups - tcgen05.ld the accumulator (one TMEM lane = one output row per thread,
//                 warpgroup g takes the 32-column chunks g, g + 2, ...), fused bias (staged in smem per tile) /
//                 row-gather / activation / residual / scale, results through swizzled smem + TMA stores.
//                 The epilogue, not the MMA loop, paces these GEMMs (K <= 1024): its latencies (TMEM load, L2 row
//                 gathers, store hand-off) only overlap across warps, hence eight of them; GATHER / RESID template
//                 flags keep the prefetch registers of unused operands out of the common instantiations.
// Tails in M, N and K are handled by TMA out-of-bounds zero fill plus masking in the epilogue.
#include "epilogue.cuh"
#include "tc_common.cuh"
#include <algorithm>

namespace vlsat {

using namespace tc;

__device__ __forceinline__ long long gtime() { unsigned long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); return (long long)g; }

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                 // fp32 elements per K block = 128 bytes
constexpr int TC_EPI_WARPS = 8;            // two epilogue warpgroups: each owns every other 32-column chunk of a tile
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_A_TILE = TC_BM * 128;    // bytes

// hi = round-to-nearest tf32 (low 13 mantissa bits zero), lo = x - hi
__global__ void tf32_split_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int64_t cols,
                                  float* __restrict__ hi, float* __restrict__ lo) {
    pdl_entry();
    const int64_t c4 = cols >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * c4) return;
    const int64_t r = idx / c4, c = (idx % c4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ldx + c));
    float in[4] = {v.x, v.y, v.z, v.w}, h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t u;
        u = tc::tf32_rna_bits(in[i]);
        h[i] = __uint_as_float(u);
        l[i] = in[i] - h[i];
    }
    *reinterpret_cast<float4*>(hi + r * cols + c) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(lo + r * cols + c) = make_float4(l[0], l[1], l[2], l[3]);
}

// Persistent kernel: gridDim.x CTAs walk the output tiles round-robin (n fastest, so the CTAs that run
// together share A row-tiles in L2). Two TMEM accumulators: the epilogue of tile i overlaps the main loop
// of tile i+1.
// MAJ: operand storage. Bit 0: A is MN-major (stored [K, M]: the reduction index runs over the ROWS of the stored matrix),
// bit 1: B is MN-major (stored [K, N]). Backward GEMMs read their operands as the forward left them, without transposes:
//   dX = dZ W      A = dZ [M, N] K-major, B = W [N, K] stored with the reduction (N) on rows -> MAJ = 2
//   dW = dZ^T X    A = dZ stored [M, N], B = X stored [M, K], reduction (M) on rows of both    -> MAJ = 3 (+ split reduction)
// An MN-major tile of one K block is [64 reduction rows x 64 elements (128 B)] TMA boxes with the 128B swizzle, one box per 64
// M/N elements, 8 KB apart (LBO); 8-row swizzle atoms are 1024 B apart along the reduction (SBO); one UMMA_K = 16 rows = 2 KB.
template <int BN, int STAGES, int PASSES, Kind KD, bool GATHER, bool RESID, int MAJ = 0>
__global__ void __launch_bounds__(TC_THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tm_ahi, const __grid_constant__ CUtensorMap tm_alo,
                 const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo,
                 const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_shi,
                 const __grid_constant__ CUtensorMap tm_slo, const LinearArgs a) {
    constexpr int B_TILE = BN * 128;
    constexpr int STAGE_BYTES = (PASSES == 3 ? 2 : 1) * (TC_A_TILE + B_TILE);
    constexpr int STG_BYTES = TC_BM * 128;                  // one staged [128 rows x 32 cols] output block
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array keeps LDS/STS
    uint8_t* staging = smem + STAGES * STAGE_BYTES;         // one block per epilogue warpgroup, swizzled, read by TMA stores
    float* bias_s = reinterpret_cast<float*>(staging + 2 * STG_BYTES);      // [2 warpgroups][BN / 2] bias of the current tile
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_s + BN);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;                // [2]
    uint64_t* acc_empty = acc_full + 2;                     // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 2);

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int BK = (KD == Kind::TF32) ? TC_BK : 2 * TC_BK;      // K elements per 128-byte block
    const int num_kb = (int)((a.K + BK - 1) / BK);
    const int tiles_n = (int)((a.N + BN - 1) / BN), tiles_m = (int)((a.M + TC_BM - 1) / TC_BM);
    const int n_tiles = tiles_m * tiles_n * a.splits;        // work items: (tile, reduction split)
    const int splits = a.splits;
    const int kb_per = splits > 1 ? a.kb_per_split : num_kb;
    constexpr bool A_MN = (MAJ & 1) != 0, B_MN = (MAJ & 2) != 0;
    static_assert(MAJ == 0 || KD == Kind::BF16, "MN-major operands: bf16 pairs only");

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_ahi); prefetch_tmap(&tm_bhi);
        if (PASSES == 3) { prefetch_tmap(&tm_alo); prefetch_tmap(&tm_blo); }
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 32 * TC_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_holder, 2 * BN); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();                                             // everything above is local; operands of the previous kernel below

    if (warp == 0) {
        // warp-uniform loops; one elected lane issues (keeps TMA / MMA issue on the uniform datapath)
        int g = 0;                                          // k-block counter across tiles -> ring position
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int tl = t / splits, sp = t - tl * splits;
            const int m0 = (tl / tiles_n) * TC_BM, n0 = (tl % tiles_n) * BN;
            const int kb_end = min(num_kb, (sp + 1) * kb_per);
            for (int kb = sp * kb_per; kb < kb_end; ++kb, ++g) {
                const int s = g % STAGES;
                mbar_wait(&empty_bar[s], ((g / STAGES) & 1) ^ 1);
                if (elect_one()) {
                    uint8_t* st = smem + s * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                    auto load_a = [&](uint8_t* dst, const CUtensorMap* tm) {
                        if constexpr (A_MN) {
#pragma unroll
                            for (int j = 0; j < TC_BM / 64; ++j) tma_load_2d(dst + j * 8192, tm, &full_bar[s], m0 + 64 * j, kb * BK);
                        } else {
                            tma_load_2d(dst, tm, &full_bar[s], kb * BK, m0);
                        }
                    };
                    auto load_b = [&](uint8_t* dst, const CUtensorMap* tm) {
                        if constexpr (B_MN) {
#pragma unroll
                            for (int j = 0; j < BN / 64; ++j) tma_load_2d(dst + j * 8192, tm, &full_bar[s], n0 + 64 * j, kb * BK);
                        } else {
                            tma_load_2d(dst, tm, &full_bar[s], kb * BK, n0);
                        }
                    };
                    load_a(st, &tm_ahi);
                    load_b(st + TC_A_TILE, &tm_bhi);
                    if (PASSES == 3) {
                        load_a(st + TC_A_TILE + B_TILE, &tm_alo);
                        load_b(st + 2 * TC_A_TILE + B_TILE, &tm_blo);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc<KD>(TC_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
        // descriptors differ only in the 14-bit start-address field: build one per operand form, then add offsets (>>4)
        const uint64_t desc0 = make_sdesc_k128(smem_u32(smem));
        const uint64_t desc0_mn = make_sdesc_mn128(smem_u32(smem), 8192, 1024);
        const uint64_t da0 = A_MN ? desc0_mn : desc0, db0 = B_MN ? desc0_mn : desc0;
        constexpr uint64_t a_step = A_MN ? (2048 >> 4) : 2, b_step = B_MN ? (2048 >> 4) : 2;    // one UMMA_K along the reduction
        int g = 0, i = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
            const int buf = i & 1;
            const int sp = t % splits;
            const int kb_begin = sp * kb_per, kb_end = min(num_kb, (sp + 1) * kb_per);
            mbar_wait(&acc_empty[buf], ((i >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
            tc_fence_after();
            if (a.trace && blockIdx.x == 0 && lane == 0 && i < 8) a.trace[i * 4 + 0] = gtime();
            const uint32_t tacc = tmem_base + buf * BN;
            for (int kb = kb_begin; kb < kb_end; ++kb, ++g) {
                const int s = g % STAGES;
                mbar_wait(&full_bar[s], (g / STAGES) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t so = (uint64_t)(s * (STAGE_BYTES >> 4));
                    const uint64_t dah = da0 + so;
                    const uint64_t dbh = db0 + so + (TC_A_TILE >> 4);
                    const uint64_t dal = da0 + so + ((TC_A_TILE + B_TILE) >> 4);
                    const uint64_t dbl = db0 + so + ((2 * TC_A_TILE + B_TILE) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {               // UMMA_K = 8 tf32 / 16 bf16 = 32 bytes = 2 x 16 B (K-major)
                        const uint32_t acc = (k > 0) ? 1u : (kb > kb_begin ? 1u : 0u);
                        if (PASSES == 3) {
                            // small terms first, then the leading product
                            mma_ss<KD>(tacc, dal + a_step * k, dbh + b_step * k, idesc, acc);
                            mma_ss<KD>(tacc, dah + a_step * k, dbl + b_step * k, idesc, 1);
                            mma_ss<KD>(tacc, dah + a_step * k, dbh + b_step * k, idesc, 1);
                        } else {
                            mma_ss<KD>(tacc, dah + a_step * k, dbh + b_step * k, idesc, acc);
                        }
                    }
                    tc_commit(&empty_bar[s]);                    // stage reusable once these MMAs retire
                    if (kb == kb_end - 1) tc_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
            if (a.trace && blockIdx.x == 0 && lane == 0 && i < 8) a.trace[i * 4 + 1] = gtime();
        }
    } else {
        const int q = warp & 3;                              // TMEM lane quarter owned by this warp
        const int wg = (warp - 2) >> 2;                      // epilogue warpgroup: chunks wg, wg + 2, ...
        const int et = (threadIdx.x - 64) & 127;             // 0..127 inside the warpgroup
        const int srow = q * 32 + lane;                      // row of the tile owned by this thread
        uint8_t* sb = staging + wg * STG_BYTES;              // this warpgroup's staging block
        float* bs = bias_s + wg * (BN / 2);
        const vlsat_epilogue& e = a.epi;
        if (a.tma_store && et == 0) { prefetch_tmap(&tm_y); prefetch_tmap(&tm_shi); prefetch_tmap(&tm_slo); }
        auto wg_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory"); };
        // stage one 128x32 block (this thread's row) and let one thread hand it to the TMA store engine
        auto stage_store = [&](const CUtensorMap* tm, const float4 (&vals)[8], int col0, int row0) {
            if (et == 0) bulk_wait_read<0>();                // the store that last read this block has drained it
            wg_sync();
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(sb + srow * 128 + ((c ^ (srow & 7)) << 4)) = vals[c];
            fence_proxy_async();
            wg_sync();
            if (et == 0) { tma_store_2d(tm, sb, col0, row0); bulk_commit(); }
        };
        // bf16 (hi, lo) pairs of one 128x32 block: two [128 rows x 64 B] halves of the staging block in the 64-byte
        // swizzle (16-byte chunk c of row r lives at c ^ ((r >> 1) & 3)), one TMA store each
        auto stage_store_bf16 = [&](const uint32_t (&hp)[16], const uint32_t (&lp)[16], int col0, int row0) {
            if (et == 0) bulk_wait_read<0>();
            wg_sync();
            const int sw = (srow >> 1) & 3;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                *reinterpret_cast<uint4*>(sb + srow * 64 + ((c ^ sw) << 4)) = make_uint4(hp[4 * c], hp[4 * c + 1], hp[4 * c + 2], hp[4 * c + 3]);
                *reinterpret_cast<uint4*>(sb + STG_BYTES / 2 + srow * 64 + ((c ^ sw) << 4)) = make_uint4(lp[4 * c], lp[4 * c + 1], lp[4 * c + 2], lp[4 * c + 3]);
            }
            fence_proxy_async();
            wg_sync();
            if (et == 0) { tma_store_2d(&tm_shi, sb, col0, row0); tma_store_2d(&tm_slo, sb + STG_BYTES / 2, col0, row0); bulk_commit(); }
        };
        auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
        const bool split_bf16 = e.split_fmt == VLSAT_SPLIT_BF16;
        const bool has_ga = GATHER && e.gather_a, has_gb = GATHER && e.gather_b, has_res = RESID && e.residual;
        // fully vectorised epilogue when every operand row is 16-byte addressable
        const bool vec_ok = a.tma_store && (BN % 4 == 0) &&
                            (!e.split_hi || (e.ld_split % (split_bf16 ? 8 : 4) == 0 && al16(e.split_hi) && al16(e.split_lo))) &&
                            (!(e.gather_a || e.gather_b) || (e.ld_gather % 4 == 0 && al16(e.gather_a) && al16(e.gather_b))) &&
                            (!e.residual || (e.ld_res % 4 == 0 && al16(e.residual)));
        const bool col_bias = e.bias && !e.bias_per_row;
        const float post_scale = e.scale_ptr ? expf(__ldg(e.scale_ptr)) : 1.f;
        int i = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
            const int buf = i & 1;
            const int tl = t / splits, sp = t - tl * splits;
            const int m0 = (tl / tiles_n) * TC_BM, n0 = (tl % tiles_n) * BN;
            const int row_out0 = m0 + sp * (int)a.slab_rows;     // split reductions store into their own slab of the y map
            const int64_t m_own = (int64_t)m0 + q * 32 + lane;
            const bool own_ok = m_own < a.M;
            const int64_t ia_own = (own_ok && e.gather_a) ? e.idx_a[m_own] : 0;
            const int64_t ib_own = (own_ok && e.gather_b) ? e.idx_b[m_own] : 0;
            const float row_bias = (own_ok && e.bias && e.bias_per_row) ? __ldg(e.bias + m_own) : 0.f;
            // this warpgroup's bias columns of the tile -> smem while the main loop still runs (et < BN / 2:
            // chunk (et / 32) * 2 + wg, column et % 32). The previous tile's readers are past their last store barrier.
            if (et < BN / 2) {
                const int64_t col = (int64_t)n0 + ((et >> 5) * 2 + wg) * 32 + (et & 31);
                bs[et] = (col_bias && col < a.N) ? __ldg(e.bias + col) : 0.f;
            }
            wg_sync();
            // row-gather / residual operands of this warpgroup's NEXT chunk are always in flight while the current one
            // is processed (and its first chunk's while the main loop of this tile still runs)
            float4 ga_n[GATHER ? 8 : 1], gb_n[GATHER ? 8 : 1], rs_n[RESID ? 8 : 1];
            auto issue_row_loads = [&](int64_t nbn) {
                const bool go = vec_ok && own_ok && nbn + 32 <= a.N;
                if constexpr (GATHER) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        ga_n[j] = (go && has_ga) ? __ldg(reinterpret_cast<const float4*>(e.gather_a + ia_own * e.ld_gather + nbn) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                        gb_n[j] = (go && has_gb) ? __ldg(reinterpret_cast<const float4*>(e.gather_b + ib_own * e.ld_gather + nbn) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                if constexpr (RESID) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        rs_n[j] = (go && has_res) ? __ldg(reinterpret_cast<const float4*>(e.residual + m_own * e.ld_res + nbn) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            issue_row_loads((int64_t)n0 + wg * 32);
            mbar_wait(&acc_full[buf], (i >> 1) & 1);
            tc_fence_after();
            if (a.trace && blockIdx.x == 0 && threadIdx.x == 64 && i < 8) a.trace[i * 4 + 2] = gtime();
            const uint32_t tacc = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int c0 = wg * 32; c0 < BN; c0 += 64) {
                const int64_t nb = (int64_t)n0 + c0;
                if (nb >= a.N) break;                        // warpgroup-uniform
                uint32_t r[32];
                tmem_ld_32x32(tacc + (uint32_t)c0, r);
                tmem_ld_wait();
                if (vec_ok && nb + 32 <= a.N) {
                    // One output row per thread (its TMEM lane), 32 consecutive columns. All global loads of the
                    // chunk are issued before any is consumed so their L2 latencies overlap; results leave through
                    // swizzled staging blocks + TMA stores (whole 128-byte lines, tails clipped by the TMA unit).
                    float4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        v[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                    if (own_ok) {
                        const float4* b4 = reinterpret_cast<const float4*>(bs + (c0 >> 6) * 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 bj = b4[j];
                            v[j].x += bj.x + row_bias; v[j].y += bj.y + row_bias; v[j].z += bj.z + row_bias; v[j].w += bj.w + row_bias;
                        }
                        if constexpr (GATHER) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                v[j].x += ga_n[j].x + gb_n[j].x; v[j].y += ga_n[j].y + gb_n[j].y;
                                v[j].z += ga_n[j].z + gb_n[j].z; v[j].w += ga_n[j].w + gb_n[j].w;
                            }
                        }
                        if (e.act == VLSAT_ACT_RELU) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) { v[j].x = fmaxf(v[j].x, 0.f); v[j].y = fmaxf(v[j].y, 0.f); v[j].z = fmaxf(v[j].z, 0.f); v[j].w = fmaxf(v[j].w, 0.f); }
                        } else if (e.act == VLSAT_ACT_SIGMOID) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) { v[j].x = apply_act(v[j].x, e.act); v[j].y = apply_act(v[j].y, e.act); v[j].z = apply_act(v[j].z, e.act); v[j].w = apply_act(v[j].w, e.act); }
                        }
                        if (RESID && has_res) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                v[j].x = e.alpha * v[j].x + e.beta * rs_n[j].x; v[j].y = e.alpha * v[j].y + e.beta * rs_n[j].y;
                                v[j].z = e.alpha * v[j].z + e.beta * rs_n[j].z; v[j].w = e.alpha * v[j].w + e.beta * rs_n[j].w;
                            }
                        } else if (e.alpha != 1.f) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) { v[j].x *= e.alpha; v[j].y *= e.alpha; v[j].z *= e.alpha; v[j].w *= e.alpha; }
                        }
                        if (e.scale_ptr) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) { v[j].x *= post_scale; v[j].y *= post_scale; v[j].z *= post_scale; v[j].w *= post_scale; }
                        }
                    }
                    // the prefetched row operands are consumed: refill them for this warpgroup's next chunk, in flight
                    // during the store hand-off below and the next TMEM load
                    if ((GATHER || RESID) && c0 + 64 < BN) issue_row_loads(nb + 64);
                    if (a.y) stage_store(&tm_y, v, (int)nb, row_out0);
                    if (e.split_hi && split_bf16) {
                        uint32_t hp[16], lp[16];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            split_bf16x2(v[j].x, v[j].y, hp[2 * j], lp[2 * j]);
                            split_bf16x2(v[j].z, v[j].w, hp[2 * j + 1], lp[2 * j + 1]);
                        }
                        stage_store_bf16(hp, lp, (int)nb, m0);
                    } else if (e.split_hi) {
                        float4 hi[8], lo[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            split_tf32(v[j].x, hi[j].x, lo[j].x); split_tf32(v[j].y, hi[j].y, lo[j].y);
                            split_tf32(v[j].z, hi[j].z, lo[j].z); split_tf32(v[j].w, hi[j].w, lo[j].w);
                        }
                        stage_store(&tm_shi, hi, (int)nb, m0);
                        stage_store(&tm_slo, lo, (int)nb, m0);
                    }
                } else if (own_ok) {
#pragma unroll 4
                    for (int j = 0; j < 32; ++j)
                        if (nb + j < a.N) {
                            const float v = epilogue_one(e, __uint_as_float(r[j]), m_own, nb + j, ia_own, ib_own, post_scale);
                            if (a.y) a.y[m_own * a.ldy + nb + j] = v;
                            if (e.split_hi) store_split(e, v, m_own, nb + j);
                        }
                }
            }
            tc_fence_before();
            if (a.trace && blockIdx.x == 0 && threadIdx.x == 64 && i < 8) a.trace[i * 4 + 3] = gtime();
            mbar_arrive(&acc_empty[buf]);                    // every epilogue thread's arrival frees the accumulator
        }
        if (et == 0) bulk_wait_all();                        // every TMA store has landed before the CTA retires
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

long long* g_trace = nullptr;

int bf16_split(const float*, int64_t, int64_t, int64_t, uint16_t*, uint16_t*, int64_t, cudaStream_t);

// kind: 0 = tf32 pairs (K % 4 == 0), 1 = bf16 pairs (K % 8 == 0: compact bf16 rows must stay 16-byte aligned)
bool linear_tc_eligible(const float* x, int64_t ldx, const float* w, int64_t ldw, int64_t M, int64_t N, int64_t K, int kind) {
    const bool common = K >= 32 && M >= 1 && N >= 8 && M < (1ll << 31) && N < (1ll << 31) && encode_fn() != nullptr;
    if (kind == 1) return common && (K % 8 == 0);
    return common && (K % 4 == 0) && (ldx % 4 == 0) && (ldw % 4 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)w % 16 == 0);
}

// both pair formats take 4 bytes per element and half
size_t linear_tc_workspace_bytes(int64_t M, int64_t N, int64_t K, bool need_x, bool need_w) {
    return (size_t)((need_x ? 2 * M * K : 0) + (need_w ? 2 * N * K : 0)) * sizeof(float);
}

int tf32_split(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* hi, float* lo, cudaStream_t st) {
    const int64_t n = rows * (cols / 4);
    if (n == 0) return VLSAT_OK;
    launch_k(tf32_split_kernel, dim3((unsigned)ceil_div(n, 256)), dim3(256), 0, st, x, ldx, rows, cols, hi, lo);
    return finish_launch();
}

template <int BN, int STAGES, int PASSES, Kind KD, bool GATHER, bool RESID, int MAJ = 0>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tal, const CUtensorMap& tb, const CUtensorMap& tbl,
                     const CUtensorMap& ty, const CUtensorMap& tsh, const CUtensorMap& tsl,
                     const LinearArgs& a, cudaStream_t st) {
    constexpr int STAGE_BYTES = (PASSES == 3 ? 2 : 1) * (TC_A_TILE + BN * 128);
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 2 * TC_BM * 128 /*store staging*/ + BN * 4 /*bias*/ + 1024 /*align*/ + 256 /*barriers*/;
    auto kern = linear_tc_kernel<BN, STAGES, PASSES, KD, GATHER, RESID, MAJ>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t n_tiles = ceil_div(a.N, BN) * ceil_div(a.M, TC_BM) * a.splits;
    const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, kNumSMs);
    launch_k(kern, dim3(grid), dim3(TC_THREADS), smem, st, ta, tal, tb, tbl, ty, tsh, tsl, a);
    return finish_launch();
}

// x_hi/x_lo and w_hi/w_lo: compact [rows, K] split operands (ld = K) in the pair format of `kind`
// (0 = tf32 floats, 1 = bf16). passes = 3 (x3 split product) or 1 (plain TF32, kind 0 only).
int linear_tc(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, float* y, int64_t ldy,
              int64_t M, int64_t N, int64_t K, const vlsat_epilogue* epi, int passes, int kind, cudaStream_t st) {
    LinearArgs a;
    a.x = (const float*)x_hi; a.ldx = K; a.w = (const float*)w_hi; a.ldw = K; a.y = y; a.ldy = ldy; a.M = M; a.N = N; a.K = K;
    if (epi) a.epi = *epi;
    else { a.epi = vlsat_epilogue{}; a.epi.alpha = 1.f; }
    a.trace = g_trace;
    // 128 x 64 tiles for narrow outputs and for small problems: when 128 x 128 tiles would leave more than half of the SMs
    // idle the GEMM is one tile-latency long, and that latency is the per-SM operand ingest (~64 B/clk): a 64-column B
    // tile is 25 % less to pull per K block, on twice as many SMs
    const int64_t tiles128 = ceil_div(M, TC_BM) * ceil_div(N, 128);
    const int bn = (N <= 64 || 2 * tiles128 <= kNumSMs) ? 64 : 128;
    const auto DT = kind == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const int eb = kind == 1 ? 2 : 4;
    const uint32_t bk = 128 / eb;                     // elements per 128-byte K block
    if (kind == 1) passes = tc_passes();              // bf16 pairs: BF16x3, or the hi halves alone in the single-pass mode
    CUtensorMap ta, tal, tb, tbl;
    bool ok = make_tmap_2d(&ta, x_hi, DT, eb, M, K, K, bk, TC_BM) && make_tmap_2d(&tb, w_hi, DT, eb, N, K, K, bk, bn);
    if (passes == 3)
        ok = ok && make_tmap_2d(&tal, x_lo, DT, eb, M, K, K, bk, TC_BM) && make_tmap_2d(&tbl, w_lo, DT, eb, N, K, K, bk, bn);
    else { tal = ta; tbl = tb; }
    if (!ok) return VLSAT_ERR_UNSUPPORTED;
    // output tensor maps (TMA stores): possible when every output row is 16-byte addressable
    auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
    const vlsat_epilogue& e = a.epi;
    const bool sbf = e.split_fmt == VLSAT_SPLIT_BF16;
    bool tma_out = (!y || (ldy % 4 == 0 && al16(y))) && (!e.split_hi || (e.ld_split % (sbf ? 8 : 4) == 0 && al16(e.split_hi) && al16(e.split_lo))) &&
                   (!e.bias || e.bias_per_row || al16(e.bias)) &&
                   (!(e.gather_a || e.gather_b) || (e.ld_gather % 4 == 0 && al16(e.gather_a) && al16(e.gather_b))) &&
                   (!e.residual || (e.ld_res % 4 == 0 && al16(e.residual)));
    CUtensorMap ty = ta, tsh = ta, tsl = ta;
    if (tma_out && y) tma_out = make_tmap_2d(&ty, y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, ldy, 32, TC_BM);
    if (tma_out && e.split_hi) {
        if (sbf)     // 64-byte wide boxes (32 bf16 columns) in the 64-byte swizzle
            tma_out = make_tmap_2d(&tsh, e.split_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, e.ld_split, 32, TC_BM, CU_TENSOR_MAP_SWIZZLE_64B) &&
                      make_tmap_2d(&tsl, e.split_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, e.ld_split, 32, TC_BM, CU_TENSOR_MAP_SWIZZLE_64B);
        else
            tma_out = make_tmap_2d(&tsh, e.split_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, e.ld_split, 32, TC_BM) &&
                      make_tmap_2d(&tsl, e.split_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, e.ld_split, 32, TC_BM);
    }
    a.tma_store = tma_out ? 1 : 0;
    const bool g = e.gather_a || e.gather_b, r = e.residual != nullptr;
#define VLSAT_TC_LAUNCH(BN_, ST_, PS_, KD_, G_, R_) launch_tc<BN_, ST_, PS_, KD_, G_, R_>(ta, tal, tb, tbl, ty, tsh, tsl, a, st)
    if (kind == 1 && passes == 1) return bn == 64 ? VLSAT_TC_LAUNCH(64, 6, 1, Kind::BF16, true, true) : VLSAT_TC_LAUNCH(128, 6, 1, Kind::BF16, true, true);
    if (kind == 1) {
        // BF16x3: the engine of the hot path gets epilogue instantiations without the unused prefetch registers
        if (bn == 64) return VLSAT_TC_LAUNCH(64, 4, 3, Kind::BF16, true, true);
        if (!g && !r) return VLSAT_TC_LAUNCH(128, 3, 3, Kind::BF16, false, false);
        if (g && !r) return VLSAT_TC_LAUNCH(128, 3, 3, Kind::BF16, true, false);
        if (!g && r) return VLSAT_TC_LAUNCH(128, 3, 3, Kind::BF16, false, true);
        return VLSAT_TC_LAUNCH(128, 3, 3, Kind::BF16, true, true);
    }
    if (passes == 3) return bn == 64 ? VLSAT_TC_LAUNCH(64, 4, 3, Kind::TF32, true, true) : VLSAT_TC_LAUNCH(128, 3, 3, Kind::TF32, true, true);
    return bn == 64 ? VLSAT_TC_LAUNCH(64, 6, 1, Kind::TF32, true, true) : VLSAT_TC_LAUNCH(128, 6, 1, Kind::TF32, true, true);
#undef VLSAT_TC_LAUNCH
}

// ------------------------------------------------------------------------------ backward GEMMs on stored operands
int sum_slabs(const float* slabs, int64_t slab_stride, int splits, float* out, int64_t ldo, int64_t rows, int64_t cols, cudaStream_t st);

// Reduction split of a [M, N] output with num_kb K blocks: fill the 148 SMs when there are few output tiles (weight
// gradients: small [N_w, K_w] outputs, reductions over thousands of rows). Every split gets at least one K block.
static void pick_reduction_split(int64_t M, int64_t N, int bn, int64_t num_kb, int* splits, int* kb_per) {
    const int64_t tiles = ceil_div(M, TC_BM) * ceil_div(N, bn);
    int best = 1; double best_cost = 1e30;
    for (int s = 1; s <= kNumSMs && s <= num_kb; ++s) {
        const int64_t per = ceil_div(num_kb, s);
        if (ceil_div(num_kb, per) != s) continue;                // same blocks per split as a smaller s: skip
        const double waves = (double)ceil_div(tiles * s, kNumSMs);
        // K blocks per wave + tile prologue / epilogue (~8 blocks' worth) + writing and summing the slabs (the sum is its own
        // launch: ~8.5 us in the ncu launch list of a training step = 14 K blocks; 95 of them per step cost 0.8 ms when the
        // penalty was 6, so split only where it pays for that)
        const double cost = waves * ((double)per + 8.0) + (s > 1 ? 14.0 + 0.1 * s : 0.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
    }
    *splits = best;
    *kb_per = (int)ceil_div(num_kb, best);
}

size_t gemm_pairs_workspace_bytes(int mode, int64_t M, int64_t N, int64_t K) {
    if (mode != 3) return 0;
    const int bn = (N <= 64) ? 64 : 128;
    int splits, kb_per;
    pick_reduction_split(M, N, bn, ceil_div(K, 64), &splits, &kb_per);
    return splits > 1 ? (size_t)splits * (size_t)(ceil_div(M, TC_BM) * TC_BM) * (size_t)N * sizeof(float) : 0;
}

// y [M, N] (fp32, row stride ldy) from bf16 (hi, lo) pair operands read AS STORED:
//   mode 2: y = a b      a stored [M, K] (row stride lda), b stored [K, N] (row stride ldb)
//   mode 3: y = a^T b    a stored [K, M], b stored [K, N]; the reduction runs over the stored rows and is split over CTAs
// Row strides in elements, multiples of 8; M, N multiples of 8 where they are a stored row length.
int gemm_pairs_tc(int mode, const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, const uint16_t* b_hi, const uint16_t* b_lo,
                  int64_t ldb, float* y, int64_t ldy, int64_t M, int64_t N, int64_t K, void* workspace, size_t workspace_bytes,
                  cudaStream_t st) {
    if (mode != 2 && mode != 3) return VLSAT_ERR_INVALID_ARG;
    if ((lda | ldb) % 8 || ldy % 4 || ((uintptr_t)y & 15) || M >= (1ll << 31) || N >= (1ll << 31) || K >= (1ll << 31)) return VLSAT_ERR_UNSUPPORTED;
    LinearArgs a;
    a.x = nullptr; a.ldx = lda; a.w = nullptr; a.ldw = ldb; a.y = y; a.ldy = ldy; a.M = M; a.N = N; a.K = K;
    a.epi = vlsat_epilogue{}; a.epi.alpha = 1.f;
    a.trace = nullptr;
    const int64_t tiles128 = ceil_div(M, TC_BM) * ceil_div(N, 128);
    const int bn = (N <= 64 || (mode == 2 && 2 * tiles128 <= kNumSMs)) ? 64 : 128;
    const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUtensorMap ta, tal, tb, tbl, ty;
    bool ok;
    if (mode == 2) ok = make_tmap_2d(&ta, a_hi, BF, 2, M, K, lda, 64, TC_BM) && make_tmap_2d(&tal, a_lo, BF, 2, M, K, lda, 64, TC_BM);
    else ok = make_tmap_2d(&ta, a_hi, BF, 2, K, M, lda, 64, 64) && make_tmap_2d(&tal, a_lo, BF, 2, K, M, lda, 64, 64);
    ok = ok && make_tmap_2d(&tb, b_hi, BF, 2, K, N, ldb, 64, 64) && make_tmap_2d(&tbl, b_lo, BF, 2, K, N, ldb, 64, 64);
    if (!ok) return VLSAT_ERR_UNSUPPORTED;
    int splits = 1, kb_per = 0;
    float* dst = y; int64_t ld_dst = ldy; int64_t rows_dst = M;
    if (mode == 3) {
        pick_reduction_split(M, N, bn, ceil_div(K, 64), &splits, &kb_per);
        if (splits > 1) {
            const size_t need = gemm_pairs_workspace_bytes(mode, M, N, K);
            if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 15) || N % 4) return VLSAT_ERR_WORKSPACE;
            a.splits = splits; a.kb_per_split = kb_per; a.slab_rows = ceil_div(M, TC_BM) * TC_BM;
            dst = (float*)workspace; ld_dst = N; rows_dst = a.slab_rows * splits;
            a.y = dst; a.ldy = ld_dst;
        }
    }
    if (!make_tmap_2d(&ty, dst, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, rows_dst, N, ld_dst, 32, TC_BM)) return VLSAT_ERR_UNSUPPORTED;
    a.tma_store = 1;
    int rc;
    if (tc_passes() == 1) {
        if (mode == 2) rc = bn == 64 ? launch_tc<64, 6, 1, Kind::BF16, false, false, 2>(ta, tal, tb, tbl, ty, ty, ty, a, st)
                                     : launch_tc<128, 6, 1, Kind::BF16, false, false, 2>(ta, tal, tb, tbl, ty, ty, ty, a, st);
        else rc = bn == 64 ? launch_tc<64, 6, 1, Kind::BF16, false, false, 3>(ta, tal, tb, tbl, ty, ty, ty, a, st)
                           : launch_tc<128, 6, 1, Kind::BF16, false, false, 3>(ta, tal, tb, tbl, ty, ty, ty, a, st);
    } else if (mode == 2) rc = bn == 64 ? launch_tc<64, 4, 3, Kind::BF16, false, false, 2>(ta, tal, tb, tbl, ty, ty, ty, a, st)
                                 : launch_tc<128, 3, 3, Kind::BF16, false, false, 2>(ta, tal, tb, tbl, ty, ty, ty, a, st);
    else rc = bn == 64 ? launch_tc<64, 4, 3, Kind::BF16, false, false, 3>(ta, tal, tb, tbl, ty, ty, ty, a, st)
                       : launch_tc<128, 3, 3, Kind::BF16, false, false, 3>(ta, tal, tb, tbl, ty, ty, ty, a, st);
    if (rc != VLSAT_OK || splits == 1) return rc;
    return sum_slabs(dst, a.slab_rows * N, splits, y, ldy, M, N, st);
}

}  // namespace vlsat

using namespace vlsat;

extern "C" size_t vlsat_gemm_pairs_workspace_bytes(int mode, int64_t M, int64_t N, int64_t K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    return gemm_pairs_workspace_bytes(mode, M, N, K);
}

extern "C" int vlsat_gemm_pairs(int mode, const void* a_hi, const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo,
                                int64_t ldb, float* y, int64_t ldy, int64_t M, int64_t N, int64_t K, void* workspace,
                                size_t workspace_bytes, void* stream) {
    VLSAT_REQUIRE(M >= 0 && N >= 0 && K >= 1 && (mode == VLSAT_GEMM_NN || mode == VLSAT_GEMM_TN));
    if (M == 0 || N == 0) return VLSAT_OK;
    VLSAT_REQUIRE(a_hi && a_lo && b_hi && b_lo && y && ldy >= N && ldb >= N && lda >= (mode == VLSAT_GEMM_NN ? K : M));
    VLSAT_SUPPORT(encode_fn() != nullptr && N >= 8);
    VLSAT_SUPPORT((((uintptr_t)a_hi | (uintptr_t)a_lo | (uintptr_t)b_hi | (uintptr_t)b_lo) & 15) == 0);
    return gemm_pairs_tc(mode, (const uint16_t*)a_hi, (const uint16_t*)a_lo, lda, (const uint16_t*)b_hi, (const uint16_t*)b_lo, ldb,
                         y, ldy, M, N, K, workspace, workspace_bytes, (cudaStream_t)stream);
}
