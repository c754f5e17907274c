// FP32 FFMA GEMM with the fused epilogue of vlsat_linear_fwd.
//   y[m,n] = post(act(sum_k x[m,k] w[n,k] + bias[n] + ga[ia[m],n] + gb[ib[m],n]))
// Both operands are K-major (nn.Linear layout), so one kernel covers every projection on the path.
// This is the exact-fp32 engine: it serves the shapes the tensor-core engine does not take (tiny K/N,
// unaligned rows) and is the numerical cross-check for it (tests/test_linear.py).
#include "epilogue.cuh"

namespace vlsat {

// BMxBN tile, BK=16, (BM/T)x(BN/T) threads each owning a TxT register tile. For T=8 the tile is split
// into two 4-wide halves BM/2 (BN/2) apart so every shared-memory read is a conflict-free LDS.128.
template <int BM, int BN, int T, bool VEC>
__global__ void __launch_bounds__((BM / T) * (BN / T))
linear_simt_kernel(const LinearArgs a) {
    pdl_entry();
    constexpr int BK = 16;
    constexpr int NT = (BM / T) * (BN / T);
    constexpr int H = T / 4;                       // number of 4-wide halves per dimension
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / T), ty = tid / (BN / T);
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;

    constexpr int A_LD = BM * BK / 4 / NT;         // float4 loads per thread per tile
    constexpr int B_LD = BN * BK / 4 / NT;
    static_assert(A_LD >= 1 && B_LD >= 1, "tile too small for the thread count");
    float4 ra[A_LD], rb[B_LD];

    auto load_tile = [&](int64_t k0) {
#pragma unroll
        for (int i = 0; i < A_LD; ++i) {
            int idx = tid + i * NT, row = idx / (BK / 4), kq = idx % (BK / 4);
            int64_t m = m0 + row, k = k0 + kq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < a.M) {
                const float* p = a.x + m * a.ldx + k;
                if (VEC) { if (k < a.K) v = __ldg(reinterpret_cast<const float4*>(p)); }
                else {
                    if (k + 0 < a.K) v.x = __ldg(p + 0);
                    if (k + 1 < a.K) v.y = __ldg(p + 1);
                    if (k + 2 < a.K) v.z = __ldg(p + 2);
                    if (k + 3 < a.K) v.w = __ldg(p + 3);
                }
            }
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < B_LD; ++i) {
            int idx = tid + i * NT, row = idx / (BK / 4), kq = idx % (BK / 4);
            int64_t n = n0 + row, k = k0 + kq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < a.N) {
                const float* p = a.w + n * a.ldw + k;
                if (VEC) { if (k < a.K) v = __ldg(reinterpret_cast<const float4*>(p)); }
                else {
                    if (k + 0 < a.K) v.x = __ldg(p + 0);
                    if (k + 1 < a.K) v.y = __ldg(p + 1);
                    if (k + 2 < a.K) v.z = __ldg(p + 2);
                    if (k + 3 < a.K) v.w = __ldg(p + 3);
                }
            }
            rb[i] = v;
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_LD; ++i) {
            int idx = tid + i * NT, row = idx / (BK / 4), kq = idx % (BK / 4);
            As[buf][kq * 4 + 0][row] = ra[i].x; As[buf][kq * 4 + 1][row] = ra[i].y;
            As[buf][kq * 4 + 2][row] = ra[i].z; As[buf][kq * 4 + 3][row] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < B_LD; ++i) {
            int idx = tid + i * NT, row = idx / (BK / 4), kq = idx % (BK / 4);
            Bs[buf][kq * 4 + 0][row] = rb[i].x; Bs[buf][kq * 4 + 1][row] = rb[i].y;
            Bs[buf][kq * 4 + 2][row] = rb[i].z; Bs[buf][kq * 4 + 3][row] = rb[i].w;
        }
    };

    float acc[T][T];
#pragma unroll
    for (int i = 0; i < T; ++i)
#pragma unroll
        for (int j = 0; j < T; ++j) acc[i][j] = 0.f;

    const int64_t n_tiles = ceil_div(a.K, BK);
    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int64_t t = 0; t < n_tiles; ++t) {
        const int buf = (int)(t & 1);
        if (t + 1 < n_tiles) load_tile((t + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[T], bv[T];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float4 v = *reinterpret_cast<const float4*>(&As[buf][k][h * (BM / H) + ty * 4]);
                av[h * 4 + 0] = v.x; av[h * 4 + 1] = v.y; av[h * 4 + 2] = v.z; av[h * 4 + 3] = v.w;
                float4 u = *reinterpret_cast<const float4*>(&Bs[buf][k][h * (BN / H) + tx * 4]);
                bv[h * 4 + 0] = u.x; bv[h * 4 + 1] = u.y; bv[h * 4 + 2] = u.z; bv[h * 4 + 3] = u.w;
            }
#pragma unroll
            for (int i = 0; i < T; ++i)
#pragma unroll
                for (int j = 0; j < T; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (t + 1 < n_tiles) store_tile(buf ^ 1);
        __syncthreads();
    }

    const float post_scale = a.epi.scale_ptr ? expf(__ldg(a.epi.scale_ptr)) : 1.f;
#pragma unroll
    for (int i = 0; i < T; ++i) {
        const int64_t m = m0 + (i / 4) * (BM / H) + ty * 4 + (i % 4);
        if (m >= a.M) continue;
        const int64_t ia = a.epi.gather_a ? a.epi.idx_a[m] : 0;
        const int64_t ib = a.epi.gather_b ? a.epi.idx_b[m] : 0;
#pragma unroll
        for (int j = 0; j < T; ++j) {
            const int64_t n = n0 + (j / 4) * (BN / H) + tx * 4 + (j % 4);
            if (n < a.N) {
                const float v = epilogue_one(a.epi, acc[i][j], m, n, ia, ib, post_scale);
                if (a.y) a.y[m * a.ldy + n] = v;
                if (a.epi.split_hi) store_split(a.epi, v, m, n);
            }
        }
    }
}

// Tiny reduction (K <= 16: the first layers of the encoders read 3-, 11- and 4-column inputs): the tiled kernel above spends
// its time on mostly empty K tiles and scalar stores (207 us for the [163840, 64] x K = 3 recompute of the PointNet backward,
// which is 42 MB of output). Here a thread holds its row of x in registers and writes four consecutive outputs as one float4;
// the weights ([N, K], a few KB) and the bias sit in shared memory. Exact fp32 FFMA, same summation order as a plain dot product.
template <int KMAX>
__global__ void __launch_bounds__(256)
linear_smallk_kernel(const LinearArgs a) {
    pdl_entry();
    extern __shared__ float sw[];                       // [N][K] weights, then [N] bias
    const int K = (int)a.K, N = (int)a.N;
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) sw[i] = __ldg(a.w + (int64_t)(i / K) * a.ldw + (i % K));
    float* sb = sw + N * K;
    for (int i = threadIdx.x; i < N; i += blockDim.x) sb[i] = (a.epi.bias && !a.epi.bias_per_row) ? __ldg(a.epi.bias + i) : 0.f;
    __syncthreads();
    const int n4 = N / 4;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.M * n4) return;
    const int64_t m = t / n4; const int n = (int)(t - m * n4) * 4;
    float xv[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) xv[k] = k < K ? __ldg(a.x + m * a.ldx + k) : 0.f;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) if (k < K) acc = fmaf(xv[k], sw[(n + j) * K + k], acc);
        acc += sb[n + j];
        o[j] = apply_act(acc, a.epi.act) * a.epi.alpha;
    }
    *reinterpret_cast<float4*>(a.y + m * a.ldy + n) = make_float4(o[0], o[1], o[2], o[3]);
    if (a.epi.split_hi) {                               // bf16 (hi, lo) pair of the same four outputs (vectorisable layout checked by the host)
        uint32_t h0, l0, h1, l1;
        split_bf16x2(o[0], o[1], h0, l0); split_bf16x2(o[2], o[3], h1, l1);
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(a.epi.split_hi) + m * a.epi.ld_split + n) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(a.epi.split_lo) + m * a.epi.ld_split + n) = make_uint2(l0, l1);
    }
}

template <int BM, int BN, int T>
static void launch_simt(const LinearArgs& a, bool vec, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(a.N, BN), (unsigned)ceil_div(a.M, BM));
    constexpr int NT = (BM / T) * (BN / T);
    if (vec) launch_k(linear_simt_kernel<BM, BN, T, true>, grid, dim3(NT), 0, st, a);
    else     launch_k(linear_simt_kernel<BM, BN, T, false>, grid, dim3(NT), 0, st, a);
}

int linear_simt(const float* x, int64_t ldx, const float* w, int64_t ldw, float* y, int64_t ldy,
                int64_t M, int64_t N, int64_t K, const vlsat_epilogue* epi, cudaStream_t st) {
    LinearArgs a;
    a.x = x; a.ldx = ldx; a.w = w; a.ldw = ldw; a.y = y; a.ldy = ldy; a.M = M; a.N = N; a.K = K;
    if (epi) a.epi = *epi;
    else { a.epi = vlsat_epilogue{}; a.epi.alpha = 1.f; }
    a.trace = nullptr; a.tma_store = 0;
    // plain bias + activation epilogue on a tiny reduction: the register-row kernel
    const vlsat_epilogue& e = a.epi;
    if (K <= 16 && N % 4 == 0 && N <= 1024 && y && ldy % 4 == 0 && ((uintptr_t)y % 16 == 0) && !e.gather_a && !e.gather_b && !e.residual &&
        !e.scale_ptr && !e.bias_per_row && M * (N / 4) < (1ll << 40) && (N * K + N) * 4 <= 48 * 1024 &&
        (!e.split_hi || (e.split_fmt == VLSAT_SPLIT_BF16 && e.ld_split % 4 == 0 && ((((uintptr_t)e.split_hi | (uintptr_t)e.split_lo) & 7) == 0)))) {
        const size_t smem = (size_t)(N * K + N) * sizeof(float);
        launch_k(linear_smallk_kernel<16>, dim3((unsigned)ceil_div(M * (N / 4), 256)), dim3(256), smem, st, a);
        return finish_launch();
    }
    const bool vec = (K % 4 == 0) && (ldx % 4 == 0) && (ldw % 4 == 0) &&
                     ((uintptr_t)x % 16 == 0) && ((uintptr_t)w % 16 == 0);
    // Big tiles only when they still fill the machine; otherwise 64x64 tiles for more CTAs.
    const int64_t big_ctas = ceil_div(M, 128) * ceil_div(N, 128);
    if (big_ctas >= 2 * kNumSMs) launch_simt<128, 128, 8>(a, vec, st);
    else                         launch_simt<64, 64, 4>(a, vec, st);
    return finish_launch();
}

}  // namespace vlsat
