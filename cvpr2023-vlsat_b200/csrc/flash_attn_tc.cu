// A9 on the tensor cores: streaming softmax(Q K^T / sqrt(dk)) V for cross_attn_rel (network_MMG.py:231),
// fp32-accurate through the same 3xTF32 operand split as the dense projections.
//
// One CTA = 128 queries of one head (dk = 64). KV tiles of 64 keys stream through a 2-stage TMA ring.
//   warp 0     : TMA producer. Q (hi, lo) once; per tile K (hi, lo) as [64 keys x 64] K-major and
//                V^T (hi, lo) as [64 dims x 64 keys] K-major (the value projection is produced transposed
//                by the GEMM, so both MMAs use the plain K-major shared-memory descriptor).
//   warp 1     : tcgen05.mma issue. S = Q K^T into TMEM (128 lanes x 64 cols); after the softmax warps
//                have written P (hi, lo) back to TMEM, PV = P V^T^T with the A operand read from TMEM.
//   warps 2..5, 6..9 : two softmax warpgroups, one query row per thread each. Warpgroup g owns the tiles with
//                t & 1 == g and TMEM buffer g (S, P_hi, P_lo, PV double-buffered = 512 columns): tcgen05.ld S,
//                online softmax in the exp2 domain, split P into tf32 hi/lo, tcgen05.st to TMEM; the tile's P.V
//                is folded into a per-warpgroup fp32 row accumulator in registers (o = o * corr + pv). The two
//                partial (m, l, o) states are merged through shared memory at the end.
// The softmax is quarter-rate (ex2) bound, so two warpgroups keep the tensor pipe (S(t+2), PV(t+1), ...) busy.
#include "common.cuh"
#include "tc_common.cuh"
#include <float.h>

namespace vlsat {

using namespace tc;

constexpr int FT_BQ = 128, FT_BKV = 64, FT_DK = 64, FT_THREADS = 320, FT_STAGES = 2;
constexpr int FT_Q_BYTES = 4 * FT_BQ * 128;              // Q_hi, Q_lo x two 32-float halves
constexpr int FT_K_STAGE = 4 * FT_BKV * 128;             // K_hi h0,h1 | K_lo h0,h1   (64 rows x 128 B each)
constexpr int FT_V_STAGE = 4 * FT_DK * 128;              // Vt_hi k0,k1 | Vt_lo k0,k1
constexpr uint32_t FT_TMEM_COLS = 512;

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
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

// D[tmem] (+)= A[tmem] * B[smem]   (A operand read from tensor memory: one lane per row, one column per tf32)
__device__ __forceinline__ void mma_ts_tf32(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// TMEM columns (all double-buffered by tile parity b = t & 1):
//   S[b] = 64 b | P_hi[b] = 128 + 64 b | P_lo[b] = 256 + 64 b | PV[b] = 384 + 64 b
__global__ void __launch_bounds__(FT_THREADS, 1)
flash_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_qhi, const __grid_constant__ CUtensorMap tm_qlo,
                     const __grid_constant__ CUtensorMap tm_khi, const __grid_constant__ CUtensorMap tm_klo,
                     const __grid_constant__ CUtensorMap tm_vhi, const __grid_constant__ CUtensorMap tm_vlo,
                     float* __restrict__ out, int64_t ldo, float* __restrict__ lse, int nq, int nk, float scale_log2e) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array keeps LDS/STS
    uint8_t* q_smem = smem;                                  // [Q_hi h0 | Q_hi h1 | Q_lo h0 | Q_lo h1] x 16 KB
    uint8_t* k_smem = smem + FT_Q_BYTES;                     // stages x FT_K_STAGE
    uint8_t* v_smem = k_smem + FT_STAGES * FT_K_STAGE;       // stages x FT_V_STAGE
    uint64_t* bars = reinterpret_cast<uint64_t*>(v_smem + FT_STAGES * FT_V_STAGE);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;                             // [2]
    uint64_t* k_empty = bars + 3;                            // [2]
    uint64_t* v_full = bars + 5;                             // [2]
    uint64_t* v_empty = bars + 7;                            // [2]
    uint64_t* s_full = bars + 9;                             // [2]
    uint64_t* p_ready = bars + 11;                           // [2]
    uint64_t* pv_full = bars + 13;                           // [2]
    uint64_t* all_done = bars + 15;                          // every MMA of this CTA has retired
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    const int q0 = blockIdx.x * FT_BQ;
    const int n_tiles = (nk + FT_BKV - 1) / FT_BKV;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_qhi); prefetch_tmap(&tm_qlo); prefetch_tmap(&tm_khi);
        prefetch_tmap(&tm_klo); prefetch_tmap(&tm_vhi); prefetch_tmap(&tm_vlo);
        mbar_init(q_full, 1);
        mbar_init(all_done, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
            mbar_init(&s_full[s], 1); mbar_init(&p_ready[s], 128); mbar_init(&pv_full[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_holder, FT_TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, FT_Q_BYTES);
            for (int h = 0; h < 2; ++h) {
                tma_load_2d(q_smem + h * 16384, &tm_qhi, q_full, head * FT_DK + h * 32, q0);
                tma_load_2d(q_smem + 32768 + h * 16384, &tm_qlo, q_full, head * FT_DK + h * 32, q0);
            }
        }
        __syncwarp();
        for (int t = 0; t < n_tiles; ++t) {
            const int s = t & 1;
            const uint32_t ph = (t >> 1) & 1;
            const int k0 = t * FT_BKV;
            mbar_wait(&k_empty[s], ph ^ 1);
            if (elect_one()) {
                uint8_t* st = k_smem + s * FT_K_STAGE;
                mbar_arrive_expect_tx(&k_full[s], FT_K_STAGE);
                for (int h = 0; h < 2; ++h) {                   // K rows = keys, columns = this head's dims
                    tma_load_2d(st + h * 8192, &tm_khi, &k_full[s], head * FT_DK + h * 32, k0);
                    tma_load_2d(st + 16384 + h * 8192, &tm_klo, &k_full[s], head * FT_DK + h * 32, k0);
                }
            }
            __syncwarp();
            mbar_wait(&v_empty[s], ph ^ 1);
            if (elect_one()) {
                uint8_t* st = v_smem + s * FT_V_STAGE;
                mbar_arrive_expect_tx(&v_full[s], FT_V_STAGE);
                for (int h = 0; h < 2; ++h) {                   // V^T rows = dims, columns = keys
                    tma_load_2d(st + h * 8192, &tm_vhi, &v_full[s], k0 + h * 32, head * FT_DK);
                    tma_load_2d(st + 16384 + h * 8192, &tm_vlo, &v_full[s], k0 + h * 32, head * FT_DK);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc<Kind::TF32>(FT_BQ, 64);
        const uint64_t dq = make_sdesc_k128(smem_u32(q_smem));
        const uint64_t dk0 = make_sdesc_k128(smem_u32(k_smem));
        const uint64_t dv0 = make_sdesc_k128(smem_u32(v_smem));
        // S(t) = Q K(t)^T into S[t & 1]; releases the K stage when done
        auto issue_s = [&](int t) {
            const int s = t & 1;
            mbar_wait(&k_full[s], (t >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dk = dk0 + (uint64_t)(s * (FT_K_STAGE >> 4));
                const uint32_t ts = tmem_base + 64 * s;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {                 // 8 dims per MMA
                    const uint32_t oa = (kk >> 2) * (16384 >> 4) + (kk & 3) * 2, ob = (kk >> 2) * (8192 >> 4) + (kk & 3) * 2;
                    mma_ss<Kind::TF32>(ts, dq + (32768 >> 4) + oa, dk + ob, idesc, kk > 0);       // Q_lo K_hi
                    mma_ss<Kind::TF32>(ts, dq + oa, dk + (16384 >> 4) + ob, idesc, 1);            // Q_hi K_lo
                    mma_ss<Kind::TF32>(ts, dq + oa, dk + ob, idesc, 1);                           // Q_hi K_hi
                }
                tc_commit(&k_empty[s]);
                tc_commit(&s_full[s]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        issue_s(0);
        if (n_tiles > 1) issue_s(1);
        for (int t = 0; t < n_tiles; ++t) {
            const int s = t & 1;
            const uint32_t ph = (t >> 1) & 1;
            mbar_wait(&v_full[s], ph);
            mbar_wait(&p_ready[s], ph);                          // P(t) is in TMEM, S[s] and PV[s] are drained
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dv = dv0 + (uint64_t)(s * (FT_V_STAGE >> 4));
                const uint32_t tphi = tmem_base + 128 + 64 * s, tplo = tmem_base + 256 + 64 * s, tpv = tmem_base + 384 + 64 * s;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {                 // 8 keys per MMA
                    const uint32_t ob = (kk >> 2) * (8192 >> 4) + (kk & 3) * 2;
                    mma_ts_tf32(tpv, tplo + kk * 8, dv + ob, idesc, kk > 0);                     // P_lo V_hi
                    mma_ts_tf32(tpv, tphi + kk * 8, dv + (16384 >> 4) + ob, idesc, 1);          // P_hi V_lo
                    mma_ts_tf32(tpv, tphi + kk * 8, dv + ob, idesc, 1);                         // P_hi V_hi
                }
                tc_commit(&v_empty[s]);
                tc_commit(&pv_full[s]);
            }
            __syncwarp();
            if (t + 2 < n_tiles) issue_s(t + 2);                 // same buffer as tile t: its S has been consumed
        }
        if (elect_one()) tc_commit(all_done);
        __syncwarp();
    } else {
        const int wg = (warp - 2) >> 2;                          // softmax warpgroup 0 / 1 = TMEM buffer = tile parity
        const int qd = warp & 3;
        const int row_l = qd * 32 + lane;
        const int row = q0 + row_l;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const uint32_t t_s = tmem_base + 64 * wg, t_phi = tmem_base + 128 + 64 * wg, t_plo = tmem_base + 256 + 64 * wg,
                       t_pv = tmem_base + 384 + 64 * wg;
        float o[FT_DK];
#pragma unroll
        for (int d = 0; d < FT_DK; ++d) o[d] = 0.f;
        float m_run = -FLT_MAX, l_run = 0.f, corr_prev = 1.f;
        uint32_t r[32], r2[32];
        // o <- o * corr + P(t) V(t) for this warpgroup's previous tile (its PV buffer must be drained before P is rewritten)
        auto fold = [&](int t, float corr) {
            mbar_wait(&pv_full[wg], (t >> 1) & 1);
            tc_fence_after();
            uint32_t a[32];
            tmem_ld_32x32(t_pv + lane_off, a);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < 32; ++d) o[d] = fmaf(o[d], corr, __uint_as_float(a[d]));
            tmem_ld_32x32(t_pv + lane_off + 32, a);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < 32; ++d) o[32 + d] = fmaf(o[32 + d], corr, __uint_as_float(a[d]));
            tc_fence_before();
        };
        int last = -1;
        for (int t = wg; t < n_tiles; t += 2) {
            const int k0 = t * FT_BKV;
            mbar_wait(&s_full[wg], (t >> 1) & 1);
            tc_fence_after();
            tmem_ld_32x32(t_s + lane_off, r);
            tmem_ld_32x32(t_s + lane_off + 32, r2);
            tmem_ld_wait();
            const bool tail = k0 + FT_BKV > nk;                  // only the last tile can be ragged
            if (tail) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (k0 + j >= nk) r[j] = __float_as_uint(-FLT_MAX);
                    if (k0 + 32 + j >= nk) r2[j] = __float_as_uint(-FLT_MAX);
                }
            }
            float mx = -FLT_MAX;
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaxf(__uint_as_float(r[j]), __uint_as_float(r2[j])));
            const float m_new = fmaxf(m_run, mx * scale_log2e);  // scale > 0 commutes with max
            const float corr = ex2(m_run - m_new);
            const float neg_m = -m_new;
            if (last >= 0) fold(last, corr_prev);                // drains PV[wg] and guarantees P[wg] is no longer read
            float rs = 0.f;
            {
                uint32_t lo[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float p = ex2(fmaf(__uint_as_float(r[j]), scale_log2e, neg_m));
                    rs += p;
                    const uint32_t h = tc::tf32_rna_bits(p);
                    r[j] = h; lo[j] = __float_as_uint(p - __uint_as_float(h));
                }
                tmem_st_32x32(t_phi + lane_off, r);
                tmem_st_32x32(t_plo + lane_off, lo);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float p = ex2(fmaf(__uint_as_float(r2[j]), scale_log2e, neg_m));
                    rs += p;
                    const uint32_t h = tc::tf32_rna_bits(p);
                    r2[j] = h; lo[j] = __float_as_uint(p - __uint_as_float(h));
                }
                tmem_st_32x32(t_phi + lane_off + 32, r2);
                tmem_st_32x32(t_plo + lane_off + 32, lo);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_ready[wg]);
            l_run = l_run * corr + rs;
            m_run = m_new;
            corr_prev = corr;
            last = t;
        }
        if (last >= 0) {
            // the correction for the last tile's P.V is 1 relative to the final running max of this warpgroup
            mbar_wait(&pv_full[wg], (last >> 1) & 1);
            tc_fence_after();
            uint32_t a[32];
            tmem_ld_32x32(t_pv + lane_off, a);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < 32; ++d) o[d] = fmaf(o[d], corr_prev, __uint_as_float(a[d]));
            tmem_ld_32x32(t_pv + lane_off + 32, a);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < 32; ++d) o[32 + d] = fmaf(o[32 + d], corr_prev, __uint_as_float(a[d]));
            tc_fence_before();
        }
        // ---- merge the two warpgroups' partial softmax states (K staging memory is free by now)
        float* mrg = reinterpret_cast<float*>(k_smem);           // [128][66]: m, l, o[64]
        if (wg == 1) {
            mbar_wait(all_done, 0);                              // K staging is reused: no MMA may still read it
            mrg[row_l * 66 + 0] = m_run; mrg[row_l * 66 + 1] = l_run;
#pragma unroll
            for (int d = 0; d < FT_DK; ++d) mrg[row_l * 66 + 2 + d] = o[d];
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (wg == 0 && row < nq) {
            const float m1 = mrg[row_l * 66 + 0], l1 = mrg[row_l * 66 + 1];
            const float m = fmaxf(m_run, m1);
            const float c0 = ex2(m_run - m), c1 = ex2(m1 - m);
            const float l = l_run * c0 + l1 * c1;
            const float inv = 1.f / l;
            float* orow = out + (int64_t)row * ldo + head * FT_DK;
#pragma unroll
            for (int d = 0; d < FT_DK; d += 4) {
                float4 res;
                res.x = (o[d] * c0 + mrg[row_l * 66 + 2 + d] * c1) * inv;
                res.y = (o[d + 1] * c0 + mrg[row_l * 66 + 3 + d] * c1) * inv;
                res.z = (o[d + 2] * c0 + mrg[row_l * 66 + 4 + d] * c1) * inv;
                res.w = (o[d + 3] * c0 + mrg[row_l * 66 + 5 + d] * c1) * inv;
                *reinterpret_cast<float4*>(orow + d) = res;
            }
            if (lse) lse[(int64_t)head * nq + row] = (m + log2f(l)) * 0.6931471805599453f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, FT_TMEM_COLS);
}

// q_hi/q_lo [nq, >= H*64] (row stride ldq), k_hi/k_lo [nk, ...] (ldk), vt_hi/vt_lo [H*64, nk] (row stride ldvt).
int flash_attn_tc(const float* q_hi, const float* q_lo, int64_t ldq, const float* k_hi, const float* k_lo, int64_t ldk,
                  const float* vt_hi, const float* vt_lo, int64_t ldvt, float* out, int64_t ldo, float* lse,
                  int64_t nq, int64_t nk, int n_heads, int dk, cudaStream_t st) {
    if (dk != FT_DK || nq >= (1ll << 31) || nk >= (1ll << 31)) return VLSAT_ERR_UNSUPPORTED;
    if ((ldq | ldk | ldvt | ldo) % 4) return VLSAT_ERR_UNSUPPORTED;
    CUtensorMap tq, tql, tk, tkl, tv, tvl;
    const uint64_t d = (uint64_t)n_heads * dk;
    const auto F32 = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    bool ok = make_tmap_2d(&tq, q_hi, F32, 4, nq, d, ldq, 32, FT_BQ) && make_tmap_2d(&tql, q_lo, F32, 4, nq, d, ldq, 32, FT_BQ) &&
              make_tmap_2d(&tk, k_hi, F32, 4, nk, d, ldk, 32, FT_BKV) && make_tmap_2d(&tkl, k_lo, F32, 4, nk, d, ldk, 32, FT_BKV) &&
              make_tmap_2d(&tv, vt_hi, F32, 4, d, nk, ldvt, 32, FT_DK) && make_tmap_2d(&tvl, vt_lo, F32, 4, d, nk, ldvt, 32, FT_DK);
    if (!ok) return VLSAT_ERR_UNSUPPORTED;
    const size_t smem = FT_Q_BYTES + FT_STAGES * (FT_K_STAGE + FT_V_STAGE) + 1024 + 256;
    cudaFuncSetAttribute(flash_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((unsigned)ceil_div(nq, FT_BQ), (unsigned)n_heads);
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)dk);
    flash_attn_tc_kernel<<<grid, FT_THREADS, smem, st>>>(tq, tql, tk, tkl, tv, tvl, out, ldo, lse, (int)nq, (int)nk, scale_log2e);
    return finish_launch();
}

}  // namespace vlsat
