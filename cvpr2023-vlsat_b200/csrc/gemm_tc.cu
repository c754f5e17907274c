// Dense projection on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32 in / fp32 out.
//
// fp32 parity on tensor cores: every operand is split once into  hi = rna_tf32(x)  and  lo = x - hi
// (exact in fp32; |lo| <= 2^-12 |x|) and each K step issues three kind::tf32 MMAs
//     D += A_hi B_hi + A_hi B_lo + A_lo B_hi
// into the same fp32 TMEM accumulator ("3xTF32"). The dropped lo*lo term and the hardware's tf32
// conversion of lo are both <= 2^-23 relative, i.e. fp32-level, so results agree with the FFMA engine
// to ~1e-6 while running on the tensor pipe.
//
// Kernel shape: one 128 x BN output tile per CTA, K blocks of 32 fp32 (= one 128-byte swizzle row).
//   warp 0      : TMA producer (4 operand tiles per stage: A_hi, A_lo, B_hi, B_lo) on an mbarrier ring
//   warp 1      : TMEM allocation + single-thread tcgen05.mma issue, tcgen05.commit releases the stage
//   warps 2..5  : epilogue - tcgen05.ld the accumulator (one TMEM lane = one output row per thread),
//                 fused bias / row-gather / activation / residual / scale, vectorised global stores
// Tails in M, N and K are handled by TMA out-of-bounds zero fill plus masking in the epilogue.
#include "epilogue.cuh"
#include "tc_common.cuh"
#include <algorithm>

namespace vlsat {

using namespace tc;

__device__ __forceinline__ long long gtime() { unsigned long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); return (long long)g; }

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                 // fp32 elements per K block = 128 bytes
constexpr int TC_THREADS = 192;
constexpr int TC_A_TILE = TC_BM * 128;    // bytes

// hi = round-to-nearest tf32 (low 13 mantissa bits zero), lo = x - hi
__global__ void tf32_split_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int64_t cols,
                                  float* __restrict__ hi, float* __restrict__ lo) {
    const int64_t c4 = cols >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * c4) return;
    const int64_t r = idx / c4, c = (idx % c4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ldx + c));
    float in[4] = {v.x, v.y, v.z, v.w}, h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(in[i]));
        h[i] = __uint_as_float(u);
        l[i] = in[i] - h[i];
    }
    *reinterpret_cast<float4*>(hi + r * cols + c) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(lo + r * cols + c) = make_float4(l[0], l[1], l[2], l[3]);
}

// Persistent kernel: gridDim.x CTAs walk the output tiles round-robin (n fastest, so the CTAs that run
// together share A row-tiles in L2). Two TMEM accumulators: the epilogue of tile i overlaps the main loop
// of tile i+1.
template <int BN, int STAGES, int PASSES>
__global__ void __launch_bounds__(TC_THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tm_ahi, const __grid_constant__ CUtensorMap tm_alo,
                 const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo,
                 const LinearArgs a) {
    constexpr int B_TILE = BN * 128;
    constexpr int STAGE_BYTES = (PASSES == 3 ? 2 : 1) * (TC_A_TILE + B_TILE);
    constexpr int SCR_LD = 36;                              // floats per scratch row (32 + pad, 16 B aligned)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* scratch = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);       // [4 warps][32][SCR_LD]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(scratch + 4 * 32 * SCR_LD);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;                // [2]
    uint64_t* acc_empty = acc_full + 2;                     // [2]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (int)((a.K + TC_BK - 1) / TC_BK);
    const int tiles_n = (int)((a.N + BN - 1) / BN), tiles_m = (int)((a.M + TC_BM - 1) / TC_BM);
    const int n_tiles = tiles_m * tiles_n;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_ahi); prefetch_tmap(&tm_bhi);
        if (PASSES == 3) { prefetch_tmap(&tm_alo); prefetch_tmap(&tm_blo); }
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_holder, 2 * BN); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // warp-uniform loops; one elected lane issues (keeps TMA / MMA issue on the uniform datapath)
        int g = 0;                                          // k-block counter across tiles -> ring position
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int m0 = (t / tiles_n) * TC_BM, n0 = (t % tiles_n) * BN;
            for (int kb = 0; kb < num_kb; ++kb, ++g) {
                const int s = g % STAGES;
                mbar_wait(&empty_bar[s], ((g / STAGES) & 1) ^ 1);
                if (elect_one()) {
                    uint8_t* st = smem + s * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                    tma_load_2d(st, &tm_ahi, &full_bar[s], kb * TC_BK, m0);
                    tma_load_2d(st + TC_A_TILE, &tm_bhi, &full_bar[s], kb * TC_BK, n0);
                    if (PASSES == 3) {
                        tma_load_2d(st + TC_A_TILE + B_TILE, &tm_alo, &full_bar[s], kb * TC_BK, m0);
                        tma_load_2d(st + 2 * TC_A_TILE + B_TILE, &tm_blo, &full_bar[s], kb * TC_BK, n0);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc<Kind::TF32>(TC_BM, BN);
        // descriptors differ only in the 14-bit start-address field: build one, then add offsets (>>4)
        const uint64_t desc0 = make_sdesc_k128(smem_u32(smem));
        int g = 0, i = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
            const int buf = i & 1;
            mbar_wait(&acc_empty[buf], ((i >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tacc = tmem_base + buf * BN;
            for (int kb = 0; kb < num_kb; ++kb, ++g) {
                const int s = g % STAGES;
                mbar_wait(&full_bar[s], (g / STAGES) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t dah = desc0 + (uint64_t)(s * (STAGE_BYTES >> 4));
                    const uint64_t dbh = dah + (TC_A_TILE >> 4);
                    const uint64_t dal = dah + ((TC_A_TILE + B_TILE) >> 4);
                    const uint64_t dbl = dah + ((2 * TC_A_TILE + B_TILE) >> 4);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {       // UMMA_K = 8 for tf32 (32 bytes = 2 x 16 B)
                        const uint32_t acc = (k > 0) ? 1u : (kb > 0 ? 1u : 0u);
                        if (PASSES == 3) {
                            // small terms first, then the leading product
                            mma_ss<Kind::TF32>(tacc, dal + 2 * k, dbh + 2 * k, idesc, acc);
                            mma_ss<Kind::TF32>(tacc, dah + 2 * k, dbl + 2 * k, idesc, 1);
                            mma_ss<Kind::TF32>(tacc, dah + 2 * k, dbh + 2 * k, idesc, 1);
                        } else {
                            mma_ss<Kind::TF32>(tacc, dah + 2 * k, dbh + 2 * k, idesc, acc);
                        }
                    }
                    tc_commit(&empty_bar[s]);                    // stage reusable once these MMAs retire
                    if (kb == num_kb - 1) tc_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;                              // TMEM lane quarter owned by this warp
        float* scr = scratch + q * 32 * SCR_LD;
        const vlsat_epilogue& e = a.epi;
        auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
        // fully vectorised epilogue when every operand row is 16-byte addressable
        const bool vec_ok = (a.ldy % 4 == 0) && al16(a.y) && (BN % 4 == 0) &&
                            (!e.split_hi || (e.ld_split % 4 == 0 && al16(e.split_hi) && al16(e.split_lo))) &&
                            (!e.bias || e.bias_per_row || al16(e.bias)) &&
                            (!(e.gather_a || e.gather_b) || (e.ld_gather % 4 == 0 && al16(e.gather_a) && al16(e.gather_b))) &&
                            (!e.residual || (e.ld_res % 4 == 0 && al16(e.residual)));
        const float post_scale = e.scale_ptr ? expf(__ldg(e.scale_ptr)) : 1.f;
        const int sub = lane >> 3, c4 = (lane & 7) * 4;      // transposed domain: 4 rows x 8 float4 per instruction
        int i = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
            const int buf = i & 1;
            const int m0 = (t / tiles_n) * TC_BM, n0 = (t % tiles_n) * BN;
            const int64_t m_own = (int64_t)m0 + q * 32 + lane;
            const bool own_ok = m_own < a.M;
            const int64_t ia_own = (own_ok && e.gather_a) ? e.idx_a[m_own] : 0;
            const int64_t ib_own = (own_ok && e.gather_b) ? e.idx_b[m_own] : 0;
            mbar_wait(&acc_full[buf], (i >> 1) & 1);
            tc_fence_after();
            const uint32_t tacc = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                const int64_t nb = (int64_t)n0 + c0;
                if (nb >= a.N) break;                        // warp-uniform
                uint32_t r[32];
                tmem_ld_32x32(tacc + (uint32_t)c0, r);
                tmem_ld_wait();
                if (vec_ok && nb + 32 <= a.N) {
                    // stage the 32x32 block through shared memory so that every global access of the warp
                    // covers 4 complete 128-byte rows instead of 32 partial ones
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(scr + lane * SCR_LD + j) =
                            make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                    __syncwarp();
                    const int64_t n = nb + c4;
                    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e.bias && !e.bias_per_row) bias4 = __ldg(reinterpret_cast<const float4*>(e.bias + n));
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rr = it * 4 + sub;
                        const int64_t m = (int64_t)m0 + q * 32 + rr;
                        const int64_t ia = __shfl_sync(0xffffffffu, ia_own, rr), ib = __shfl_sync(0xffffffffu, ib_own, rr);
                        if (m >= a.M) continue;
                        float4 v = *reinterpret_cast<const float4*>(scr + rr * SCR_LD + c4);
                        if (e.bias) {
                            if (e.bias_per_row) { const float b = __ldg(e.bias + m); v.x += b; v.y += b; v.z += b; v.w += b; }
                            else { v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w; }
                        }
                        if (e.gather_a) { const float4 g4 = __ldg(reinterpret_cast<const float4*>(e.gather_a + ia * e.ld_gather + n)); v.x += g4.x; v.y += g4.y; v.z += g4.z; v.w += g4.w; }
                        if (e.gather_b) { const float4 g4 = __ldg(reinterpret_cast<const float4*>(e.gather_b + ib * e.ld_gather + n)); v.x += g4.x; v.y += g4.y; v.z += g4.z; v.w += g4.w; }
                        if (e.act == VLSAT_ACT_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                        else if (e.act == VLSAT_ACT_SIGMOID) { v.x = apply_act(v.x, e.act); v.y = apply_act(v.y, e.act); v.z = apply_act(v.z, e.act); v.w = apply_act(v.w, e.act); }
                        if (e.residual) {
                            const float4 g4 = __ldg(reinterpret_cast<const float4*>(e.residual + m * e.ld_res + n));
                            v.x = e.alpha * v.x + e.beta * g4.x; v.y = e.alpha * v.y + e.beta * g4.y;
                            v.z = e.alpha * v.z + e.beta * g4.z; v.w = e.alpha * v.w + e.beta * g4.w;
                        } else if (e.alpha != 1.f) { v.x *= e.alpha; v.y *= e.alpha; v.z *= e.alpha; v.w *= e.alpha; }
                        if (e.scale_ptr) { v.x *= post_scale; v.y *= post_scale; v.z *= post_scale; v.w *= post_scale; }
                        if (a.y) *reinterpret_cast<float4*>(a.y + m * a.ldy + n) = v;
                        if (e.split_hi) {
                            float4 hi, lo;
                            split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y);
                            split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
                            *reinterpret_cast<float4*>(e.split_hi + m * e.ld_split + n) = hi;
                            *reinterpret_cast<float4*>(e.split_lo + m * e.ld_split + n) = lo;
                        }
                    }
                    __syncwarp();
                } else if (own_ok) {
#pragma unroll 4
                    for (int j = 0; j < 32; ++j)
                        if (nb + j < a.N) {
                            const float v = epilogue_one(e, __uint_as_float(r[j]), m_own, nb + j, ia_own, ib_own, post_scale);
                            if (a.y) a.y[m_own * a.ldy + nb + j] = v;
                            if (e.split_hi) split_tf32(v, e.split_hi[m_own * e.ld_split + nb + j], e.split_lo[m_own * e.ld_split + nb + j]);
                        }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);                    // 128 arrivals free the accumulator
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

long long* g_trace = nullptr;

bool linear_tc_eligible(const float* x, int64_t ldx, const float* w, int64_t ldw, int64_t M, int64_t N, int64_t K) {
    return (K % 4 == 0) && K >= 32 && (ldx % 4 == 0) && (ldw % 4 == 0) && ((uintptr_t)x % 16 == 0) &&
           ((uintptr_t)w % 16 == 0) && M >= 1 && N >= 8 && M < (1ll << 31) && N < (1ll << 31) && encode_fn() != nullptr;
}

size_t linear_tc_workspace_bytes(int64_t M, int64_t N, int64_t K, bool need_x, bool need_w) {
    return (size_t)((need_x ? 2 * M * K : 0) + (need_w ? 2 * N * K : 0)) * sizeof(float);
}

int tf32_split(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* hi, float* lo, cudaStream_t st) {
    const int64_t n = rows * (cols / 4);
    if (n == 0) return VLSAT_OK;
    tf32_split_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(x, ldx, rows, cols, hi, lo);
    return finish_launch();
}

template <int BN, int STAGES, int PASSES>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tal, const CUtensorMap& tb, const CUtensorMap& tbl,
                     const LinearArgs& a, cudaStream_t st) {
    constexpr int STAGE_BYTES = (PASSES == 3 ? 2 : 1) * (TC_A_TILE + BN * 128);
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 4 * 32 * 36 * 4 /*epilogue scratch*/ + 1024 /*align*/ + 256 /*barriers*/;
    auto kern = linear_tc_kernel<BN, STAGES, PASSES>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t n_tiles = ceil_div(a.N, BN) * ceil_div(a.M, TC_BM);
    const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, kNumSMs);
    kern<<<grid, TC_THREADS, smem, st>>>(ta, tal, tb, tbl, a);
    return finish_launch();
}

// x_hi/x_lo and w_hi/w_lo: compact [rows, K] split operands (ld = K). passes = 3 (3xTF32) or 1 (plain TF32).
int linear_tc(const float* x_hi, const float* x_lo, const float* w_hi, const float* w_lo, float* y, int64_t ldy,
              int64_t M, int64_t N, int64_t K, const vlsat_epilogue* epi, int passes, cudaStream_t st) {
    LinearArgs a;
    a.x = x_hi; a.ldx = K; a.w = w_hi; a.ldw = K; a.y = y; a.ldy = ldy; a.M = M; a.N = N; a.K = K;
    if (epi) a.epi = *epi;
    else { a.epi = vlsat_epilogue{}; a.epi.alpha = 1.f; }
    a.trace = g_trace;
    const int bn = (N <= 64) ? 64 : 128;
    CUtensorMap ta, tal, tb, tbl;
    bool ok = make_tmap_2d(&ta, x_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, K, K, TC_BK, TC_BM) &&
              make_tmap_2d(&tb, w_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, N, K, K, TC_BK, bn);
    if (passes == 3)
        ok = ok && make_tmap_2d(&tal, x_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, K, K, TC_BK, TC_BM) &&
             make_tmap_2d(&tbl, w_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, N, K, K, TC_BK, bn);
    else { tal = ta; tbl = tb; }
    if (!ok) return VLSAT_ERR_UNSUPPORTED;
    if (passes == 3) return bn == 64 ? launch_tc<64, 4, 3>(ta, tal, tb, tbl, a, st) : launch_tc<128, 3, 3>(ta, tal, tb, tbl, a, st);
    return bn == 64 ? launch_tc<64, 6, 1>(ta, tal, tb, tbl, a, st) : launch_tc<128, 6, 1>(ta, tal, tb, tbl, a, st);
}

}  // namespace vlsat
