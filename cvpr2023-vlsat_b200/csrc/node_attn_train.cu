// A6 + A7, differentiable form: attention between the nodes of one scene with the distance bias given as an
// explicit per-pair tensor (so that autograd can sum its gradient over the 2L attention calls that share it and
// run the bias MLP's backward once). Reference: network_MMG.py:181-205,217-218; attention.py:54-77.
//
//   bias row of the pair (query a, key b) = pair_off[a] + (b - seg_start[a]),  [*, H]
//   S[h, b] = q_a[h] . k_b[h] / sqrt(dk) + bias[pair, h];  P = softmax_b S;  out_a[h] = sum_b P[h, b] v_b[h]
// backward (one CTA per query, P recomputed):
//   dP[h,b] = dO_a[h] . v_b[h];  dS = P (dP - sum_b P dP);  dbias[pair, h] = dS
//   dq_a[h] = sum_b dS k_b[h] / sqrt(dk);  dk_b[h] += dS q_a[h] / sqrt(dk);  dv_b[h] += P dO_a[h]   (atomic over queries)
#include "common.cuh"
#include <float.h>

namespace vlsat {

constexpr int NAT_THREADS = 256;

// scores + softmax of query a into shared memory: sp[h * ns + j], j = key index inside the scene
__device__ __forceinline__ void nat_probs(const float* __restrict__ qs, const float* __restrict__ k, int64_t ldk,
                                          const float* __restrict__ bias, int64_t pair0, int s0, int ns, int H, int dk,
                                          float scale, float* sp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = NAT_THREADS / 32;
    for (int t = tid; t < H * ns; t += NAT_THREADS) {
        const int j = t / H, h = t % H;                 // consecutive threads: heads of one key row (contiguous in memory)
        const float4* kp = reinterpret_cast<const float4*>(k + (int64_t)(s0 + j) * ldk + h * dk);
        const float4* qp = reinterpret_cast<const float4*>(qs + h * dk);
        float dot = 0.f;
        for (int d = 0; d < dk / 4; ++d) {
            const float4 kv = __ldg(kp + d), qv = qp[d];
            dot = fmaf(kv.x, qv.x, dot); dot = fmaf(kv.y, qv.y, dot); dot = fmaf(kv.z, qv.z, dot); dot = fmaf(kv.w, qv.w, dot);
        }
        sp[h * ns + j] = dot * scale + __ldg(bias + (pair0 + j) * H + h);
    }
    __syncthreads();
    for (int h = warp; h < H; h += nwarp) {
        float mx = -FLT_MAX;
        for (int j = lane; j < ns; j += 32) mx = fmaxf(mx, sp[h * ns + j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < ns; j += 32) { const float e = expf(sp[h * ns + j] - mx); sp[h * ns + j] = e; sum += e; }
        const float inv = 1.f / warp_sum(sum);
        for (int j = lane; j < ns; j += 32) sp[h * ns + j] *= inv;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NAT_THREADS)
node_attn_bias_fwd_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                          const float* __restrict__ v, int64_t ldv, const float* __restrict__ bias,
                          const int64_t* __restrict__ pair_off, const int32_t* __restrict__ seg_start,
                          const int32_t* __restrict__ seg_end, int H, int dk, float* __restrict__ out, int64_t ldo) {
    pdl_entry();
    extern __shared__ __align__(16) float sm[];
    const int64_t a = blockIdx.x;
    const int s0 = seg_start[a], ns = seg_end[a] - s0;
    float* qs = sm;                 // H * dk
    float* sp = sm + H * dk;        // H * ns
    for (int i = threadIdx.x; i < H * dk; i += NAT_THREADS) qs[i] = __ldg(q + a * ldq + i);
    __syncthreads();
    nat_probs(qs, k, ldk, bias, pair_off[a], s0, ns, H, dk, rsqrtf((float)dk), sp);
    for (int i = threadIdx.x; i < H * dk; i += NAT_THREADS) {
        const int h = i / dk;
        float acc = 0.f;
        for (int j = 0; j < ns; ++j) acc = fmaf(sp[h * ns + j], __ldg(v + (int64_t)(s0 + j) * ldv + i), acc);
        out[a * ldo + i] = acc;
    }
}

__global__ void __launch_bounds__(NAT_THREADS)
node_attn_bias_bwd_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                          const float* __restrict__ v, int64_t ldv, const float* __restrict__ bias,
                          const int64_t* __restrict__ pair_off, const int32_t* __restrict__ seg_start,
                          const int32_t* __restrict__ seg_end, const float* __restrict__ dout, int64_t lddo, int H, int dk,
                          float* __restrict__ dq, int64_t lddq, float* __restrict__ dkk, int64_t lddk, float* __restrict__ dv,
                          int64_t lddv, float* __restrict__ dbias) {
    pdl_entry();
    extern __shared__ __align__(16) float sm[];
    const int64_t a = blockIdx.x;
    const int s0 = seg_start[a], ns = seg_end[a] - s0;
    const int64_t pair0 = pair_off[a];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = NAT_THREADS / 32;
    float* qs = sm;                    // H * dk
    float* dos = qs + H * dk;          // H * dk
    float* sp = dos + H * dk;          // H * ns   probabilities
    float* sd = sp + H * ns;           // H * ns   dP, then dS
    const float scale = rsqrtf((float)dk);
    for (int i = tid; i < H * dk; i += NAT_THREADS) { qs[i] = __ldg(q + a * ldq + i); dos[i] = __ldg(dout + a * lddo + i); }
    __syncthreads();
    nat_probs(qs, k, ldk, bias, pair0, s0, ns, H, dk, scale, sp);
    for (int t = tid; t < H * ns; t += NAT_THREADS) {
        const int j = t / H, h = t % H;
        const float4* vp = reinterpret_cast<const float4*>(v + (int64_t)(s0 + j) * ldv + h * dk);
        const float4* gp = reinterpret_cast<const float4*>(dos + h * dk);
        float dot = 0.f;
        for (int d = 0; d < dk / 4; ++d) {
            const float4 vv = __ldg(vp + d), gv = gp[d];
            dot = fmaf(vv.x, gv.x, dot); dot = fmaf(vv.y, gv.y, dot); dot = fmaf(vv.z, gv.z, dot); dot = fmaf(vv.w, gv.w, dot);
        }
        sd[h * ns + j] = dot;
    }
    __syncthreads();
    for (int h = warp; h < H; h += nwarp) {
        float acc = 0.f;
        for (int j = lane; j < ns; j += 32) acc += sp[h * ns + j] * sd[h * ns + j];
        acc = warp_sum(acc);
        for (int j = lane; j < ns; j += 32) {
            const float ds = sp[h * ns + j] * (sd[h * ns + j] - acc);
            sd[h * ns + j] = ds;
            dbias[(pair0 + j) * H + h] = ds;
        }
    }
    __syncthreads();
    for (int i = tid; i < H * dk; i += NAT_THREADS) {
        const int h = i / dk;
        const float qv = qs[i] * scale, gv = dos[i];
        float acc = 0.f;
        for (int j = 0; j < ns; ++j) {
            const float ds = sd[h * ns + j], p = sp[h * ns + j];
            acc = fmaf(ds, __ldg(k + (int64_t)(s0 + j) * ldk + i), acc);
            atomicAdd(dkk + (int64_t)(s0 + j) * lddk + i, ds * qv);
            atomicAdd(dv + (int64_t)(s0 + j) * lddv + i, p * gv);
        }
        dq[a * lddq + i] = acc * scale;
    }
}

}  // namespace vlsat

using namespace vlsat;

static int nat_check(int n_heads, int dk, int max_scene, size_t smem) {
    if (n_heads < 1 || dk < 4 || dk % 4 != 0 || max_scene < 1) return VLSAT_ERR_UNSUPPORTED;
    if (smem > 200 * 1024) return VLSAT_ERR_UNSUPPORTED;
    return VLSAT_OK;
}

/* max_scene = largest scene size in the batch (host-known upper bound; sizes the shared-memory score buffer). */
extern "C" int vlsat_node_attn_bias_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                        const float* bias, const int64_t* pair_off, const int32_t* seg_start,
                                        const int32_t* seg_end, int n_heads, int dk, int max_scene, float* out, int64_t ldo,
                                        int64_t n_nodes, void* stream) {
    VLSAT_REQUIRE(n_nodes >= 0);
    if (n_nodes == 0) return VLSAT_OK;
    VLSAT_REQUIRE(q && k && v && bias && pair_off && seg_start && seg_end && out);
    const size_t smem = sizeof(float) * ((size_t)n_heads * dk + (size_t)n_heads * max_scene);
    int rc = nat_check(n_heads, dk, max_scene, smem);
    if (rc) return rc;
    VLSAT_SUPPORT(ldk % 4 == 0 && ((uintptr_t)k % 16 == 0));
    cudaFuncSetAttribute(node_attn_bias_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(node_attn_bias_fwd_kernel, dim3((unsigned)n_nodes), dim3(NAT_THREADS), smem, (cudaStream_t)stream, q, ldq, k, ldk, v, ldv, bias, pair_off, seg_start,
                                                                                              seg_end, n_heads, dk, out, ldo);
    return finish_launch();
}

/* dk_out / dv_out must be zero-filled by the caller (accumulated atomically over the queries of a scene). */
extern "C" int vlsat_node_attn_bias_bwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                        const float* bias, const int64_t* pair_off, const int32_t* seg_start,
                                        const int32_t* seg_end, const float* dout, int64_t lddo, int n_heads, int dk,
                                        int max_scene, float* dq, int64_t lddq, float* dk_out, int64_t lddk, float* dv_out,
                                        int64_t lddv, float* dbias, int64_t n_nodes, void* stream) {
    VLSAT_REQUIRE(n_nodes >= 0);
    if (n_nodes == 0) return VLSAT_OK;
    VLSAT_REQUIRE(q && k && v && bias && pair_off && seg_start && seg_end && dout && dq && dk_out && dv_out && dbias);
    const size_t smem = sizeof(float) * (2 * (size_t)n_heads * dk + 2 * (size_t)n_heads * max_scene);
    int rc = nat_check(n_heads, dk, max_scene, smem);
    if (rc) return rc;
    VLSAT_SUPPORT(ldk % 4 == 0 && ldv % 4 == 0 && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0));
    cudaFuncSetAttribute(node_attn_bias_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(node_attn_bias_bwd_kernel, dim3((unsigned)n_nodes), dim3(NAT_THREADS), smem, (cudaStream_t)stream, 
        q, ldq, k, ldk, v, ldv, bias, pair_off, seg_start, seg_end, dout, lddo, n_heads, dk, dq, lddq, dk_out, lddk, dv_out, lddv, dbias);
    return finish_launch();
}
