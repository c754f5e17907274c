// A1 on the tensor cores: fused PointNet encoder (c_in -> 64 -> 128 -> c_out, ReLU after every layer, max
// over the points of an object), BF16x3 (bf16 (hi, lo) pairs, three kind::f16 MMAs per K step: ~1e-5 relative, see
// csrc/gemm_tc.cu), no TMA: every operand is produced on chip.
//
// Orientation: channels on the TMEM lanes, points on the columns, so the max over points is a per-thread
// running maximum (no cross-lane reduction):
//   D2[c2 = 128 lanes][64 pts] = W2 (A operand, TMEM resident, hi/lo)  x  h1^T (B operand, smem, K-major)
//   D3[ch = 128 lanes][64 pts] = W3 chunk (A operand, TMEM resident)   x  h2^T (B operand, smem, K-major)
// A CTA is stationary on one 128-channel chunk of W3 and walks over objects; per 64-point tile:
//   workers (4 warps)  layer 1 in FFMA -> h1 tile as bf16 hi/lo in the 128B-swizzled UMMA layout
//   MMA warp           MMA2 -> D2
//   workers            h2 = relu(D2 + b2) -> hi/lo -> swizzled smem (transposing store, conflict free)
//   MMA warp           MMA3 -> D3
//   workers            running max over the tile's valid points (+ arg max)
// Layers 1-2 are recomputed by the c_out/128 chunk CTAs of an object (8% of the FLOPs, on the tensor pipe).
// HBM traffic: the points once per chunk CTA (L2 hits after the first) and c_out floats per object.
#include "common.cuh"
#include "epilogue.cuh"
#include "tc_common.cuh"
#include <float.h>
#include <algorithm>

namespace vlsat {

using namespace tc;


constexpr int PT_WORKERS = 256;          // two worker groups: group h owns the points [32h, 32h + 32) of every 64-point tile
constexpr int PT_THREADS = 32 + PT_WORKERS;   // warp 0: MMA issue + TMEM; warps 1..8: workers (TMEM lane quarter = warp & 3)
constexpr int PT_TP = 64;                // points per tile
constexpr int PT_C1 = 64, PT_C2 = 128, PT_CIN_MAX = 16;
// TMEM columns (bf16 A operands: two K elements per 32-bit column):
//   W3 hi [0,64) lo [64,128) | W2 hi [128,160) lo [160,192) | D2 [192,256) | D3 [256,320)
constexpr uint32_t PT_W3HI = 0, PT_W3LO = 64, PT_W2HI = 128, PT_W2LO = 160, PT_D2 = 192, PT_D3 = 256;

__device__ __forceinline__ void pt_tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
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
__device__ __forceinline__ void pt_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// one value -> bf16 hi / lo halves (round half up on the magnitude, see split_bf16x2)
__device__ __forceinline__ void pt_split16(float v, uint16_t& hi, uint16_t& lo) {
    const uint32_t h = (__float_as_uint(v) + 0x8000u) & 0xffff0000u;
    hi = (uint16_t)(h >> 16);
    lo = (uint16_t)((__float_as_uint(v - __uint_as_float(h)) + 0x8000u) >> 16);
}
__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(PT_WORKERS) : "memory"); }

struct PointNetTcSmem {
    uint8_t b1_hi[PT_TP * 128];          // h1 tile  [64 pts][64 k] bf16: one k-block of 64 rows x 128 B (swizzled)
    uint8_t b1_lo[PT_TP * 128];
    uint8_t b2_hi[2 * PT_TP * 128];      // h2 tile  [64 pts][128 k] bf16: 2 k-blocks
    uint8_t b2_lo[2 * PT_TP * 128];
    float xs[PT_CIN_MAX][PT_TP];
    float w1[PT_C1][PT_CIN_MAX];
    float b1[PT_C1];
    float mx[128]; int mi[128];          // running max / arg max of the upper point half, handed to the lower half per object
    uint64_t bars[4];                    // b1_ready (workers), d2_full (1), b2_ready (workers), d3_full (1)
    uint32_t tmem_holder;
};

// CIN > 0: compile-time input width (3 / 6 / 9 / 11 on the path); CIN == 0: run-time width up to 16.
template <int CIN>
__global__ void __launch_bounds__(PT_THREADS, 1)
pointnet_tc_kernel(const float* __restrict__ x, int64_t n_obj, int c_in_rt, int64_t n_pts,
                   const float* __restrict__ w1, const float* __restrict__ b1,
                   const float* __restrict__ w2, const float* __restrict__ b2,
                   const float* __restrict__ w3, const float* __restrict__ b3, int c_out,
                   float* __restrict__ out, int32_t* __restrict__ argmax) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    PointNetTcSmem& s = *reinterpret_cast<PointNetTcSmem*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));   // keeps the shared address space (LDS/STS)
    uint64_t* b1_ready = &s.bars[0]; uint64_t* d2_full = &s.bars[1]; uint64_t* b2_ready = &s.bars[2]; uint64_t* d3_full = &s.bars[3];
    const int c_in = CIN > 0 ? CIN : c_in_rt;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks = c_out / 128;
    const int chunk = blockIdx.x % n_chunks;
    const int obj_stride = gridDim.x / n_chunks;
    const int obj0 = blockIdx.x / n_chunks;
    const int tiles_per_obj = (int)((n_pts + PT_TP - 1) / PT_TP);
    const int n_local_obj = obj0 < n_obj ? (int)((n_obj - obj0 + obj_stride - 1) / obj_stride) : 0;
    const int total_it = n_local_obj * tiles_per_obj;            // tiles this CTA walks, flattened (object-major)

    if (threadIdx.x == 0) {
        mbar_init(b1_ready, PT_WORKERS); mbar_init(d2_full, 1); mbar_init(b2_ready, PT_WORKERS); mbar_init(d3_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) { tmem_alloc(&s.tmem_holder, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = s.tmem_holder;
    pdl_wait();                                                  // everything above is local to the CTA

    if (warp == 0) {
        // ------------------------------------------------------------------ MMA issue (one elected lane)
        constexpr uint32_t idesc = make_idesc<Kind::BF16>(128, PT_TP);
        const uint64_t d_b1hi = make_sdesc_k128(smem_u32(s.b1_hi)), d_b1lo = make_sdesc_k128(smem_u32(s.b1_lo));
        const uint64_t d_b2hi = make_sdesc_k128(smem_u32(s.b2_hi)), d_b2lo = make_sdesc_k128(smem_u32(s.b2_lo));
        for (int it = 0; it < total_it; ++it) {
            const uint32_t ph = it & 1;
            mbar_wait(b1_ready, ph);                             // h1 tile in smem (and W2/W3 in TMEM on the first pass)
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < PT_C1 / 16; ++kk) {           // 16 channels per MMA = 8 packed TMEM columns = 32 smem bytes
                    const uint32_t ob = (kk >> 2) * ((PT_TP * 128) >> 4) + (kk & 3) * 2;
                    pt_mma_ts(tm + PT_D2, tm + PT_W2LO + kk * 8, d_b1hi + ob, idesc, kk > 0);
                    pt_mma_ts(tm + PT_D2, tm + PT_W2HI + kk * 8, d_b1lo + ob, idesc, 1);
                    pt_mma_ts(tm + PT_D2, tm + PT_W2HI + kk * 8, d_b1hi + ob, idesc, 1);
                }
                tc_commit(d2_full);
            }
            __syncwarp();
            mbar_wait(b2_ready, ph);                             // h2 tile in smem; D3 of the previous tile drained
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < PT_C2 / 16; ++kk) {
                    const uint32_t ob = (kk >> 2) * ((PT_TP * 128) >> 4) + (kk & 3) * 2;
                    pt_mma_ts(tm + PT_D3, tm + PT_W3LO + kk * 8, d_b2hi + ob, idesc, kk > 0);
                    pt_mma_ts(tm + PT_D3, tm + PT_W3HI + kk * 8, d_b2lo + ob, idesc, 1);
                    pt_mma_ts(tm + PT_D3, tm + PT_W3HI + kk * 8, d_b2hi + ob, idesc, 1);
                }
                tc_commit(d3_full);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------------------ workers
        const int wt = threadIdx.x - 32;                         // 0..255
        const int half = wt >> 7;                                // point half of the tile this thread reads back from TMEM
        const int qd = warp & 3;
        const int l = qd * 32 + lane;                            // TMEM lane = channel index inside W2 / the W3 chunk
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        // stage layer-1 weights; put W2 row l and W3 row (chunk*128 + l) into TMEM as tf32 hi / lo
        for (int i = wt; i < PT_C1 * c_in; i += PT_WORKERS) s.w1[i / c_in][i % c_in] = __ldg(w1 + i);
        for (int i = wt; i < PT_C1; i += PT_WORKERS) s.b1[i] = __ldg(b1 + i);
        if (half == 0) {                                         // one writer per TMEM lane
            uint32_t hi[32], lo[32];
            const float* w2row = w2 + (int64_t)l * PT_C1;
#pragma unroll
            for (int j = 0; j < 32; ++j) split_bf16x2(__ldg(w2row + 2 * j), __ldg(w2row + 2 * j + 1), hi[j], lo[j]);
            pt_tmem_st32(tm + PT_W2HI + lane_off, hi);
            pt_tmem_st32(tm + PT_W2LO + lane_off, lo);
            const float* w3row = w3 + ((int64_t)chunk * 128 + l) * PT_C2;
            for (int c0 = 0; c0 < PT_C2 / 2; c0 += 32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) split_bf16x2(__ldg(w3row + 2 * (c0 + j)), __ldg(w3row + 2 * (c0 + j) + 1), hi[j], lo[j]);
                pt_tmem_st32(tm + PT_W3HI + lane_off + c0, hi);
                pt_tmem_st32(tm + PT_W3LO + lane_off + c0, lo);
            }
            tmem_st_wait();
        }
        const float b2l = __ldg(b2 + l);
        const float b3l = __ldg(b3 + chunk * 128 + l);
        const int p_own = wt & 63, kq = wt >> 6;                 // layer 1: point p_own, channels kq*16 .. +15
        constexpr int XPT = (PT_CIN_MAX * PT_TP + PT_WORKERS - 1) / PT_WORKERS;    // x values each worker stages per tile (<= 4)
        float xr[XPT];
        // register prefetch of the x tile of flattened iteration `it` (global latency hidden behind the previous tile)
        auto prefetch_x = [&](int it) {
            const int64_t obj = obj0 + (int64_t)(it / tiles_per_obj) * obj_stride;
            const int64_t p0 = (int64_t)(it % tiles_per_obj) * PT_TP;
            const float* xo = x + obj * (int64_t)c_in * n_pts;
#pragma unroll
            for (int u = 0; u < XPT; ++u) {
                const int i = wt + u * PT_WORKERS;
                const int d = i / PT_TP, p = i % PT_TP;
                xr[u] = (it < total_it && d < c_in && p0 + p < n_pts) ? __ldg(xo + d * n_pts + p0 + p) : 0.f;
            }
        };
        // layer 1 (FFMA): h1[p][k] = relu(b1[k] + sum_d w1[k][d] x[d][p]) -> B1 tile as tf32 hi/lo, swizzled rows
        auto layer1 = [&](int it) {
#pragma unroll
            for (int u = 0; u < XPT; ++u) {
                const int i = wt + u * PT_WORKERS;
                if (i < c_in * PT_TP) s.xs[i / PT_TP][i % PT_TP] = xr[u];
            }
            worker_barrier();
            prefetch_x(it + 1);
            float xv[CIN > 0 ? CIN : PT_CIN_MAX];
#pragma unroll
            for (int d = 0; d < (CIN > 0 ? CIN : PT_CIN_MAX); ++d) xv[d] = (d < c_in) ? s.xs[d][p_own] : 0.f;
            uint8_t* rowh = s.b1_hi + p_own * 128;               // row = point, 64 channels x bf16 = 128 bytes
            uint8_t* rowl = s.b1_lo + p_own * 128;
#pragma unroll
            for (int c = 0; c < 2; ++c) {                        // 2 chunks of 8 channels = one 16-byte unit each
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float a2[2];
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const int k = kq * 16 + c * 8 + u * 2 + v;
                        float acc = s.b1[k];
#pragma unroll
                        for (int d = 0; d < (CIN > 0 ? CIN : PT_CIN_MAX); ++d)
                            if (CIN > 0 || d < c_in) acc = fmaf(s.w1[k][d], xv[d], acc);
                        a2[v] = fmaxf(acc, 0.f);
                    }
                    split_bf16x2(a2[0], a2[1], hi[u], lo[u]);
                }
                const int pos = ((kq * 2 + c) ^ (p_own & 7)) * 16;      // 128B swizzle: 16-byte chunk index XOR (row mod 8)
                *reinterpret_cast<uint4*>(rowh + pos) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(rowl + pos) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            fence_proxy_async();                                 // generic-proxy smem writes -> visible to the tensor core
            tc_fence_before();
            mbar_arrive(b1_ready);
            worker_barrier();                                    // xs may be overwritten by the next call
        };
        worker_barrier();                                        // w1 / b1 staged
        float run_max = -FLT_MAX;
        int run_idx = 0;
        if (total_it > 0) { prefetch_x(0); layer1(0); }
        for (int it = 0; it < total_it; ++it) {
            const uint32_t ph = it & 1;
            const int t = it % tiles_per_obj;
            const int64_t obj = obj0 + (int64_t)(it / tiles_per_obj) * obj_stride;
            const int64_t p0 = (int64_t)t * PT_TP;
            const int valid = (int)min((int64_t)PT_TP, n_pts - p0);
            // ---- layer 2 epilogue: h2 = relu(D2 + b2) -> B2 tile (transposing, swizzled store)
            mbar_wait(d2_full, ph);
            tc_fence_after();
            {
                uint8_t* bh = s.b2_hi + (l >> 6) * (PT_TP * 128);         // k-block of channel l (64 bf16 channels per 128-byte row)
                uint8_t* bl = s.b2_lo + (l >> 6) * (PT_TP * 128);
                const int chunk16 = (l & 63) >> 3, cb = (l & 7) * 2;      // 16-byte chunk of channel l inside the row, byte inside it
                uint32_t a0[32];
                tmem_ld_32x32(tm + PT_D2 + lane_off + 32 * half, a0);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    uint16_t hi, lo;
                    pt_split16(fmaxf(__uint_as_float(a0[j]) + b2l, 0.f), hi, lo);
                    const int off = (32 * half + j) * 128 + ((chunk16 ^ (j & 7)) << 4) + cb;
                    *reinterpret_cast<uint16_t*>(bh + off) = hi;
                    *reinterpret_cast<uint16_t*>(bl + off) = lo;
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(b2_ready);
            // ---- layer 1 of the NEXT tile while the tensor core runs MMA3 of this one
            if (it + 1 < total_it) layer1(it + 1);
            // ---- layer 3 epilogue: running max over the valid points of the tile
            mbar_wait(d3_full, ph);
            tc_fence_after();
            {
                uint32_t a0[32];
                tmem_ld_32x32(tm + PT_D3 + lane_off + 32 * half, a0);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float v = __uint_as_float(a0[j]);
                    if (32 * half + j < valid && v > run_max) { run_max = v; run_idx = (int)p0 + 32 * half + j; }
                }
            }
            tc_fence_before();
            if (t == tiles_per_obj - 1) {
                // the two point halves of a channel meet in smem; ties keep the lower point index (the order a single
                // scan over the points would have produced)
                if (half == 1) { s.mx[l] = run_max; s.mi[l] = run_idx; }
                worker_barrier();
                if (half == 0) {
                    const float om = s.mx[l]; const int oi = s.mi[l];
                    if (om > run_max || (om == run_max && oi < run_idx)) { run_max = om; run_idx = oi; }
                    const int64_t o = obj * c_out + chunk * 128 + l;
                    out[o] = fmaxf(run_max + b3l, 0.f);          // bias and ReLU commute with the max
                    if (argmax) argmax[o] = run_idx;
                }
                worker_barrier();                                // s.mx / s.mi are rewritten at the next object
                run_max = -FLT_MAX; run_idx = 0;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

bool pointnet_tc_eligible(int c_in, int c1, int c2, int c_out, int64_t n_pts) {
    return c1 == PT_C1 && c2 == PT_C2 && c_in <= PT_CIN_MAX && c_out % 128 == 0 && c_out / 128 <= kNumSMs && n_pts >= 1;
}

int pointnet_tc(const float* x, int64_t n_obj, int c_in, int64_t n_pts, const float* w1, const float* b1,
                const float* w2, const float* b2, const float* w3, const float* b3, int c_out,
                float* out, int32_t* argmax, cudaStream_t st) {
    const int n_chunks = c_out / 128;
    const int64_t per_chunk = std::max<int64_t>(1, std::min<int64_t>(n_obj, kNumSMs / n_chunks));
    const unsigned grid = (unsigned)(per_chunk * n_chunks);
    const size_t smem = sizeof(PointNetTcSmem) + 1024;
#define PT_LAUNCH(CIN_)                                                                                          \
    do {                                                                                                          \
        cudaFuncSetAttribute(pointnet_tc_kernel<CIN_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        launch_k(pointnet_tc_kernel<CIN_>, dim3(grid), dim3(PT_THREADS), smem, st, x, n_obj, c_in, n_pts, w1, b1, w2, b2, w3, b3, \
                 c_out, out, argmax);                                                                             \
    } while (0)
    switch (c_in) {
        case 3: PT_LAUNCH(3); break;
        case 6: PT_LAUNCH(6); break;
        case 9: PT_LAUNCH(9); break;
        case 11: PT_LAUNCH(11); break;
        default: PT_LAUNCH(0); break;
    }
#undef PT_LAUNCH
    return finish_launch();
}

}  // namespace vlsat
