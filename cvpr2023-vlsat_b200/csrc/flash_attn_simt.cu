// A9 (exact-fp32 engine): streaming softmax(Q K^T / sqrt(dk)) V over all keys, FFMA only.
// The reference materialises [1,H,nq,nk] scores three times over (attention.py:60-75); here a CTA owns
// 64 queries of one head, streams 64-key tiles through shared memory and keeps an online softmax, so
// memory is O(n*d). This engine is the numerical cross-check of the tensor-core engine and serves
// head sizes it does not take.
#include "common.cuh"
#include <float.h>

namespace vlsat {

constexpr int FA_BQ = 64, FA_BK = 64, FA_THREADS = 256;

template <int DK>
__global__ void __launch_bounds__(FA_THREADS)
flash_attn_simt_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                       const float* __restrict__ v, int64_t ldv, float* __restrict__ out, int64_t ldo,
                       float* __restrict__ lse, int64_t nq, int64_t nk, float scale_log2e) {
    constexpr int SD = DK + 4;          // padded row stride (floats): rows 1 apart are 16 B apart mod 128 B
    constexpr int SP = FA_BK + 4;
    constexpr int DV = DK / 16;         // output dims per thread (4 for dk 64)
    extern __shared__ __align__(16) float sm[];
    float* Qs = sm;                     // [64][SD]
    float* Ks = Qs + FA_BQ * SD;        // [64][SD]
    float* Vs = Ks + FA_BK * SD;        // [64][SD]
    float* Ps = Vs + FA_BK * SD;        // [64][SP]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int head = blockIdx.y;
    const int64_t q0 = (int64_t)blockIdx.x * FA_BQ;

    for (int i = tid; i < FA_BQ * DK / 4; i += FA_THREADS) {
        const int r = i / (DK / 4), c = i % (DK / 4);
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < nq) val = __ldg(reinterpret_cast<const float4*>(q + (q0 + r) * ldq + head * DK) + c);
        *reinterpret_cast<float4*>(Qs + r * SD + c * 4) = val;
    }
    float m_run[4], l_run[4], o[4][DV];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m_run[i] = -FLT_MAX; l_run[i] = 0.f;
#pragma unroll
        for (int d = 0; d < DV; ++d) o[i][d] = 0.f;
    }

    for (int64_t k0 = 0; k0 < nk; k0 += FA_BK) {
        __syncthreads();                // previous tile's Ks/Vs/Ps fully consumed (and Qs staged)
        for (int i = tid; i < FA_BK * DK / 4; i += FA_THREADS) {
            const int r = i / (DK / 4), c = i % (DK / 4);
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (k0 + r < nk) {
                kv = __ldg(reinterpret_cast<const float4*>(k + (k0 + r) * ldk + head * DK) + c);
                vv = __ldg(reinterpret_cast<const float4*>(v + (k0 + r) * ldv + head * DK) + c);
            }
            *reinterpret_cast<float4*>(Ks + r * SD + c * 4) = kv;
            *reinterpret_cast<float4*>(Vs + r * SD + c * 4) = vv;
        }
        __syncthreads();
        // S = Q K^T : rows ty+16i, cols tx+16j
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
        for (int d = 0; d < DK; d += 4) {
            float4 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * i) * SD + d);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * SD + d);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s[i][j] = fmaf(a[i].x, b[j].x, s[i][j]); s[i][j] = fmaf(a[i].y, b[j].y, s[i][j]);
                    s[i][j] = fmaf(a[i].z, b[j].z, s[i][j]); s[i][j] = fmaf(a[i].w, b[j].w, s[i][j]);
                }
        }
        // online softmax (base-2 domain): a row's 64 scores live in the 16 tx lanes of one half-warp
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -FLT_MAX;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s[i][j] = (k0 + tx + 16 * j < nk) ? s[i][j] * scale_log2e : -FLT_MAX;
                mx = fmaxf(mx, s[i][j]);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float m_new = fmaxf(m_run[i], mx);
            const float corr = exp2f(m_run[i] - m_new);
            float rs = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float p = (k0 + tx + 16 * j < nk) ? exp2f(s[i][j] - m_new) : 0.f;
                Ps[(ty + 16 * i) * SP + tx + 16 * j] = p;
                rs += p;
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
            l_run[i] = l_run[i] * corr + rs;
            m_run[i] = m_new;
#pragma unroll
            for (int d = 0; d < DV; ++d) o[i][d] *= corr;
        }
        __syncthreads();
        // O += P V : rows ty+16i, dims tx*4 + 64*g
#pragma unroll 2
        for (int c = 0; c < FA_BK; c += 4) {
            float4 p[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) p[i] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * i) * SP + c);
#pragma unroll
            for (int g = 0; g < DV / 4; ++g) {
                float4 vv[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) vv[cc] = *reinterpret_cast<const float4*>(Vs + (c + cc) * SD + g * 64 + tx * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float pc[4] = {p[i].x, p[i].y, p[i].z, p[i].w};
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        o[i][g * 4 + 0] = fmaf(pc[cc], vv[cc].x, o[i][g * 4 + 0]);
                        o[i][g * 4 + 1] = fmaf(pc[cc], vv[cc].y, o[i][g * 4 + 1]);
                        o[i][g * 4 + 2] = fmaf(pc[cc], vv[cc].z, o[i][g * 4 + 2]);
                        o[i][g * 4 + 3] = fmaf(pc[cc], vv[cc].w, o[i][g * 4 + 3]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = q0 + ty + 16 * i;
        if (r >= nq) continue;
        const float inv = 1.f / l_run[i];
#pragma unroll
        for (int g = 0; g < DV / 4; ++g) {
            float4 res = make_float4(o[i][g * 4 + 0] * inv, o[i][g * 4 + 1] * inv, o[i][g * 4 + 2] * inv, o[i][g * 4 + 3] * inv);
            *reinterpret_cast<float4*>(out + r * ldo + head * DK + g * 64 + tx * 4) = res;
        }
        if (lse && tx == 0) lse[(int64_t)head * nq + r] = (m_run[i] + log2f(l_run[i])) * 0.6931471805599453f;
    }
}

int flash_attn_simt(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                    float* out, int64_t ldo, float* lse, int64_t nq, int64_t nk, int n_heads, int dk, cudaStream_t st) {
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)dk);
    dim3 grid((unsigned)ceil_div(nq, FA_BQ), (unsigned)n_heads);
    if (dk == 64) {
        const size_t smem = sizeof(float) * (3 * 64 * (64 + 4) + 64 * (FA_BK + 4));
        cudaFuncSetAttribute(flash_attn_simt_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        flash_attn_simt_kernel<64><<<grid, FA_THREADS, smem, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, lse, nq, nk, scale_log2e);
    } else if (dk == 128) {
        const size_t smem = sizeof(float) * (3 * 64 * (128 + 4) + 64 * (FA_BK + 4));
        cudaFuncSetAttribute(flash_attn_simt_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        flash_attn_simt_kernel<128><<<grid, FA_THREADS, smem, st>>>(q, ldq, k, ldk, v, ldv, out, ldo, lse, nq, nk, scale_log2e);
    } else {
        return VLSAT_ERR_UNSUPPORTED;
    }
    return finish_launch();
}

}  // namespace vlsat
