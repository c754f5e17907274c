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

template <int BN, int STAGES, int PASSES>
__global__ void __launch_bounds__(TC_THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tm_ahi, const __grid_constant__ CUtensorMap tm_alo,
                 const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo,
                 const LinearArgs a) {
    constexpr int B_TILE = BN * 128;
    constexpr int STAGE_BYTES = (PASSES == 3 ? 2 : 1) * (TC_A_TILE + B_TILE);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
    long long* trace = (blockIdx.x == 0 && blockIdx.y == 0) ? a.trace : nullptr;
    if (trace && threadIdx.x == 0) trace[0] = clock64();
    const int cta_lin = blockIdx.y * gridDim.x + blockIdx.x;
    if (a.trace && threadIdx.x == 0 && cta_lin < 600) {
        unsigned long long gt; unsigned smid;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        a.trace[128 + 3 * cta_lin] = (long long)gt; a.trace[128 + 3 * cta_lin + 2] = smid;
    }
    const int num_kb = (int)((a.K + TC_BK - 1) / TC_BK);

    auto a_hi = [&](int s) { return smem + s * STAGE_BYTES; };
    auto b_hi = [&](int s) { return smem + s * STAGE_BYTES + TC_A_TILE; };
    auto a_lo = [&](int s) { return smem + s * STAGE_BYTES + TC_A_TILE + B_TILE; };
    auto b_lo = [&](int s) { return smem + s * STAGE_BYTES + 2 * TC_A_TILE + B_TILE; };

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_ahi); prefetch_tmap(&tm_bhi);
        if (PASSES == 3) { prefetch_tmap(&tm_alo); prefetch_tmap(&tm_blo); }
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_holder, BN); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    if (trace && threadIdx.x == 0) { trace[1] = clock64(); trace[100] = gtime(); }

    if (warp == 0) {
        // warp-uniform loop; one elected lane issues (keeps the TMA / MMA issue on the uniform datapath
        // without per-instruction divergence handling)
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (trace && lane == 0 && kb < 40) trace[8 + 3 * kb] = clock64();
            if (elect_one()) {
                mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                tma_load_2d(a_hi(s), &tm_ahi, &full_bar[s], kb * TC_BK, m0);
                tma_load_2d(b_hi(s), &tm_bhi, &full_bar[s], kb * TC_BK, n0);
                if (PASSES == 3) {
                    tma_load_2d(a_lo(s), &tm_alo, &full_bar[s], kb * TC_BK, m0);
                    tma_load_2d(b_lo(s), &tm_blo, &full_bar[s], kb * TC_BK, n0);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc<Kind::TF32>(TC_BM, BN);
        // descriptors differ only in the 14-bit start-address field: build one, then add offsets (>>4)
        const uint64_t desc0 = make_sdesc_k128(smem_u32(smem));
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (trace && lane == 0 && kb < 40) { trace[8 + 3 * kb + 1] = clock64(); if (kb == 0) trace[101] = gtime(); if (kb == num_kb - 1) trace[102] = gtime(); }
            if (elect_one()) {
                const uint64_t dah = desc0 + (uint64_t)(s * (STAGE_BYTES >> 4));
                const uint64_t dbh = dah + (TC_A_TILE >> 4);
                const uint64_t dal = dah + ((TC_A_TILE + B_TILE) >> 4);
                const uint64_t dbl = dah + ((2 * TC_A_TILE + B_TILE) >> 4);
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {           // UMMA_K = 8 for tf32 (32 bytes = 2 x 16 B)
                    const uint32_t acc = (k > 0) ? 1u : (kb > 0 ? 1u : 0u);
                    if (PASSES == 3) {
                        // small terms first, then the leading product
                        mma_ss<Kind::TF32>(tmem_base, dal + 2 * k, dbh + 2 * k, idesc, acc);
                        mma_ss<Kind::TF32>(tmem_base, dah + 2 * k, dbl + 2 * k, idesc, 1);
                        mma_ss<Kind::TF32>(tmem_base, dah + 2 * k, dbh + 2 * k, idesc, 1);
                    } else {
                        mma_ss<Kind::TF32>(tmem_base, dah + 2 * k, dbh + 2 * k, idesc, acc);
                    }
                }
                tc_commit(&empty_bar[s]);                        // stage reusable once these MMAs retire
            }
            __syncwarp();
            if (trace && lane == 0 && kb < 40) trace[8 + 3 * kb + 2] = clock64();
        }
        if (elect_one()) tc_commit(accum_bar);                   // accumulator complete
        __syncwarp();
    } else {
        const int q = warp & 3;                                  // TMEM lane quarter owned by this warp
        const int64_t m = (int64_t)m0 + q * 32 + lane;
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        if (trace && threadIdx.x == 64) { trace[2] = clock64(); trace[103] = gtime(); }
        const bool row_ok = m < a.M;
        const int64_t ia = (row_ok && a.epi.gather_a) ? a.epi.idx_a[m] : 0;
        const int64_t ib = (row_ok && a.epi.gather_b) ? a.epi.idx_b[m] : 0;
        const float post_scale = a.epi.scale_ptr ? expf(__ldg(a.epi.scale_ptr)) : 1.f;
        const vlsat_epilogue& e = a.epi;
        auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
        // fully vectorised epilogue when every operand row is 16-byte addressable
        const bool vec_ok = (a.ldy % 4 == 0) && al16(a.y) && (n0 % 4 == 0) &&
                            (!e.bias || e.bias_per_row || al16(e.bias)) &&
                            (!(e.gather_a || e.gather_b) || (e.ld_gather % 4 == 0 && al16(e.gather_a) && al16(e.gather_b))) &&
                            (!e.residual || (e.ld_res % 4 == 0 && al16(e.residual)));
        const float row_bias = (row_ok && e.bias && e.bias_per_row) ? __ldg(e.bias + m) : 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            if (n0 + c0 >= a.N) break;                           // warp-uniform
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            if (!row_ok) continue;
            const int64_t nb = n0 + c0;
            float* yrow = a.y + m * a.ldy + nb;
            if (vec_ok && nb + 32 <= a.N) {
                // every branch below is warp-uniform and sits OUTSIDE the element loops, so the sigmoid's
                // exp/divide are never executed (or if-converted) on the ReLU / identity paths
                float t[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) t[j] = __uint_as_float(r[j]);
                if (e.bias) {
                    if (e.bias_per_row) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) t[j] += row_bias;
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + nb + j));
                            t[j] += b.x; t[j + 1] += b.y; t[j + 2] += b.z; t[j + 3] += b.w;
                        }
                    }
                }
                if (e.gather_a) {
                    const float4* ga = reinterpret_cast<const float4*>(e.gather_a + ia * e.ld_gather + nb);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const float4 g = __ldg(ga + j); t[4 * j] += g.x; t[4 * j + 1] += g.y; t[4 * j + 2] += g.z; t[4 * j + 3] += g.w; }
                }
                if (e.gather_b) {
                    const float4* gb = reinterpret_cast<const float4*>(e.gather_b + ib * e.ld_gather + nb);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const float4 g = __ldg(gb + j); t[4 * j] += g.x; t[4 * j + 1] += g.y; t[4 * j + 2] += g.z; t[4 * j + 3] += g.w; }
                }
                if (e.act == VLSAT_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) t[j] = fmaxf(t[j], 0.f);
                } else if (e.act == VLSAT_ACT_SIGMOID) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) t[j] = 1.f / (1.f + expf(-t[j]));
                }
                if (e.residual) {
                    const float4* rr = reinterpret_cast<const float4*>(e.residual + m * e.ld_res + nb);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 g = __ldg(rr + j);
                        t[4 * j] = e.alpha * t[4 * j] + e.beta * g.x; t[4 * j + 1] = e.alpha * t[4 * j + 1] + e.beta * g.y;
                        t[4 * j + 2] = e.alpha * t[4 * j + 2] + e.beta * g.z; t[4 * j + 3] = e.alpha * t[4 * j + 3] + e.beta * g.w;
                    }
                } else if (e.alpha != 1.f) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) t[j] *= e.alpha;
                }
                if (e.scale_ptr) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) t[j] *= post_scale;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(yrow + j) = make_float4(t[j], t[j + 1], t[j + 2], t[j + 3]);
            } else {
#pragma unroll 4
                for (int j = 0; j < 32; ++j)
                    if (nb + j < a.N)
                        yrow[j] = epilogue_one(e, __uint_as_float(r[j]), m, nb + j, ia, ib, post_scale);
            }
        }
    }
    if (trace && threadIdx.x == 64) trace[104] = gtime();
    tc_fence_before();
    __syncthreads();
    if (trace && threadIdx.x == 0) { trace[3] = clock64(); trace[105] = gtime(); }
    if (a.trace && threadIdx.x == 0 && cta_lin < 600) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        a.trace[128 + 3 * cta_lin + 1] = (long long)gt;
    }
    if (warp == 1) tmem_dealloc(tmem_base, BN);
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
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    auto kern = linear_tc_kernel<BN, STAGES, PASSES>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((unsigned)ceil_div(a.N, BN), (unsigned)ceil_div(a.M, TC_BM));
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
