// A1: fused PointNet encoder - per-point MLP (c_in -> 64 -> 128 -> c_out, ReLU after each layer) and the
// symmetric max over the points of an object, in ONE kernel: the [n_obj, C, P] activations of the
// reference (network_PointNet.py:141-164) never exist in HBM. Algorithmic HBM traffic is the input
// points (c_in*4 B per point) plus c_out*4 B per object; the 1x1-conv weights are staged in shared memory.
//
// One CTA = one object x one group of 64-channel output chunks. Per tile of 64 points:
//   layer 1, 2 -> shared memory (K-major rows, padded so float4 reads along K are conflict-free),
//   layer 3    -> 64x64 register-tiled FFMA product per output chunk, bias+ReLU, max over the tile's
//                 points folded into a running per-channel maximum.
#include "common.cuh"
#include <float.h>

namespace vlsat {

constexpr int PN_TP = 64;     // points per tile
constexpr int PN_CC = 64;     // output channels per chunk
constexpr int PN_C1 = 64;
constexpr int PN_C2 = 128;
constexpr int PN_CIN_MAX = 16;
constexpr int PN_THREADS = 256;
constexpr int PN_S1 = PN_C1 + 4;   // row stride of K-major smem operands with K = c1
constexpr int PN_S2 = PN_C2 + 4;   // ... with K = c2 (132: rows 1 apart are 16 B apart mod 128 B)

struct PointNetSmem {
    float w2[PN_C2][PN_S1];
    float w3[PN_CC][PN_S2];
    float h2[PN_TP][PN_S2];
    float h1[PN_TP][PN_S1];
    float xs[PN_CIN_MAX][PN_TP];
    float w1[PN_C1][PN_CIN_MAX];
    float b1[PN_C1];
    float b2[PN_C2];
    float red[16][PN_CC];
    int red_idx[16][PN_CC];
};

__global__ void __launch_bounds__(PN_THREADS, 1)
pointnet_fwd_kernel(const float* __restrict__ x, int c_in, int64_t n_pts,
                    const float* __restrict__ w1, const float* __restrict__ b1,
                    const float* __restrict__ w2, const float* __restrict__ b2,
                    const float* __restrict__ w3, const float* __restrict__ b3, int c_out,
                    int chunks_per_cta, float* __restrict__ out, int32_t* __restrict__ argmax) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PointNetSmem& s = *reinterpret_cast<PointNetSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t obj = blockIdx.x;
    const int n_chunks = (c_out + PN_CC - 1) / PN_CC;
    const int chunk_lo = blockIdx.y * chunks_per_cta;
    const int chunk_hi = min(n_chunks, chunk_lo + chunks_per_cta);
    const float* xo = x + obj * (int64_t)c_in * n_pts;

    // stage layer-1/2 weights once per CTA
    for (int i = tid; i < PN_C1 * c_in; i += PN_THREADS) s.w1[i / c_in][i % c_in] = __ldg(w1 + i);
    for (int i = tid; i < PN_C1; i += PN_THREADS) s.b1[i] = __ldg(b1 + i);
    for (int i = tid; i < PN_C2; i += PN_THREADS) s.b2[i] = __ldg(b2 + i);
    for (int i = tid; i < PN_C2 * PN_C1 / 4; i += PN_THREADS) {
        int r = i / (PN_C1 / 4), q = i % (PN_C1 / 4);
        *reinterpret_cast<float4*>(&s.w2[r][q * 4]) = __ldg(reinterpret_cast<const float4*>(w2 + r * PN_C1) + q);
    }

    // running maximum of the channels this thread finalises (thread t < 64 owns channel t of each chunk)
    constexpr int MAX_CHUNKS_PER_CTA = 16;
    float run_max[MAX_CHUNKS_PER_CTA];
    int run_idx[MAX_CHUNKS_PER_CTA];
#pragma unroll
    for (int i = 0; i < MAX_CHUNKS_PER_CTA; ++i) { run_max[i] = -FLT_MAX; run_idx[i] = 0; }

    for (int64_t p0 = 0; p0 < n_pts; p0 += PN_TP) {
        const int valid = (int)min((int64_t)PN_TP, n_pts - p0);
        __syncthreads();   // previous tile fully consumed (h2, xs) / weights staged
        for (int i = tid; i < c_in * PN_TP; i += PN_THREADS) {
            int d = i / PN_TP, p = i % PN_TP;
            s.xs[d][p] = (p < valid) ? __ldg(xo + d * n_pts + p0 + p) : 0.f;
        }
        __syncthreads();
        // layer 1: h1[p][j] = relu(b1[j] + sum_d w1[j][d] x[d][p]);  lane -> j, 4 point groups
        {
            const int j = tid & 63, pg = tid >> 6;
            float wj[PN_CIN_MAX];
#pragma unroll
            for (int d = 0; d < PN_CIN_MAX; ++d) wj[d] = (d < c_in) ? s.w1[j][d] : 0.f;
            const float bj = s.b1[j];
            for (int p = pg; p < PN_TP; p += 4) {
                float acc = bj;
#pragma unroll
                for (int d = 0; d < PN_CIN_MAX; ++d) if (d < c_in) acc = fmaf(wj[d], s.xs[d][p], acc);
                s.h1[p][j] = fmaxf(acc, 0.f);
            }
        }
        __syncthreads();
        // layer 2: h2[p][c] = relu(b2[c] + sum_k h1[p][k] w2[c][k]); thread: p = ty+16i (4), c = tx+16j (8)
        {
            float acc[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 4
            for (int k = 0; k < PN_C1; k += 4) {
                float4 a[4], b[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(&s.h1[ty + 16 * i][k]);
#pragma unroll
                for (int j = 0; j < 8; ++j) b[j] = *reinterpret_cast<const float4*>(&s.w2[tx + 16 * j][k]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                        acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                        acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                        acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
                    }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    s.h2[ty + 16 * i][tx + 16 * j] = fmaxf(acc[i][j] + s.b2[tx + 16 * j], 0.f);
        }
        // layer 3 + max, one 64-channel chunk at a time
        for (int ch = chunk_lo; ch < chunk_hi; ++ch) {
            const int c0 = ch * PN_CC;
            __syncthreads();   // h2 ready (first chunk) / previous chunk's w3 + red consumed
            for (int i = tid; i < PN_CC * PN_C2 / 4; i += PN_THREADS) {
                int r = i / (PN_C2 / 4), q = i % (PN_C2 / 4);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + r < c_out) v = __ldg(reinterpret_cast<const float4*>(w3 + (int64_t)(c0 + r) * PN_C2) + q);
                *reinterpret_cast<float4*>(&s.w3[r][q * 4]) = v;
            }
            __syncthreads();
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
            for (int k = 0; k < PN_C2; k += 4) {
                float4 a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(&s.h2[ty + 16 * i][k]);
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(&s.w3[tx + 16 * j][k]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                        acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                        acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                        acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
                    }
            }
            // max over this thread's 4 points (pre-activation: relu and bias commute with max)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float m = -FLT_MAX; int mi = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int p = ty + 16 * i;
                    if (p < valid && acc[i][j] > m) { m = acc[i][j]; mi = p; }
                }
                s.red[ty][tx + 16 * j] = m;
                s.red_idx[ty][tx + 16 * j] = mi;
            }
            __syncthreads();
            if (tid < PN_CC) {
                float m = run_max[0]; int mi = run_idx[0];
                // (compile-time indexing of the per-chunk registers)
#pragma unroll
                for (int q = 0; q < MAX_CHUNKS_PER_CTA; ++q) if (q == ch - chunk_lo) { m = run_max[q]; mi = run_idx[q]; }
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const float v = s.red[r][tid];
                    if (v > m) { m = v; mi = (int)p0 + s.red_idx[r][tid]; }
                }
#pragma unroll
                for (int q = 0; q < MAX_CHUNKS_PER_CTA; ++q) if (q == ch - chunk_lo) { run_max[q] = m; run_idx[q] = mi; }
            }
        }
    }
    if (tid < PN_CC) {
#pragma unroll
        for (int q = 0; q < MAX_CHUNKS_PER_CTA; ++q) {
            const int ch = chunk_lo + q;
            const int c = ch * PN_CC + tid;
            if (ch < chunk_hi && c < c_out) {
                out[obj * c_out + c] = fmaxf(run_max[q] + __ldg(b3 + c), 0.f);
                if (argmax) argmax[obj * c_out + c] = run_idx[q];
            }
        }
    }
}

}  // namespace vlsat

using namespace vlsat;

extern "C" int vlsat_pointnet_fwd(const float* x, int64_t n_obj, int c_in, int64_t n_pts,
                                  const float* w1, const float* b1, int c1,
                                  const float* w2, const float* b2, int c2,
                                  const float* w3, const float* b3, int c_out,
                                  float* out, int32_t* argmax, void* stream) {
    VLSAT_REQUIRE(x && w1 && b1 && w2 && b2 && w3 && b3 && out);
    VLSAT_REQUIRE(n_obj >= 0 && n_pts >= 1 && c_in >= 1 && c_out >= 1);
    VLSAT_SUPPORT(c1 == PN_C1 && c2 == PN_C2 && c_in <= PN_CIN_MAX);
    if (n_obj == 0) return VLSAT_OK;
    const int n_chunks = (c_out + PN_CC - 1) / PN_CC;
    // Fill the machine when there are few objects by splitting the output chunks across CTAs
    // (layers 1-2 are recomputed per split: 8% of the FLOPs).
    int splits = (int)min((int64_t)n_chunks, max((int64_t)1, ceil_div(2 * kNumSMs, n_obj)));
    int chunks_per_cta = (int)ceil_div(n_chunks, splits);
    VLSAT_SUPPORT(chunks_per_cta <= 16);
    splits = (int)ceil_div(n_chunks, chunks_per_cta);
    VLSAT_SUPPORT(n_obj <= 0x7fffffff);
    const size_t smem = sizeof(PointNetSmem);
    cudaFuncSetAttribute(pointnet_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((unsigned)n_obj, (unsigned)splits);
    pointnet_fwd_kernel<<<grid, PN_THREADS, smem, (cudaStream_t)stream>>>(
        x, c_in, n_pts, w1, b1, w2, b2, w3, b3, c_out, chunks_per_cta, out, argmax);
    return finish_launch();
}
