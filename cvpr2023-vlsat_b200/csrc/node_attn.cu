// A6 + A7: attention between the nodes of one scene, with the distance-bias MLP evaluated on the fly.
//
// The reference (network_MMG.py:181-205) loops over scenes in Python, materialises a dense
// [1,H,N,N] bias and a [1,1,N,N] block-diagonal mask and runs dense N x N attention. Masked scores are
// -inf, i.e. exactly zero weight, so attention restricted to the scene of the query is the same maths.
// One CTA per query node; keys stream in tiles of 64 with an online softmax, so a scene may be any size.
#include "common.cuh"
#include <float.h>

namespace vlsat {

// packed self_attn_fc weights (floats): see ops.py::pack_attn_fc
//   w0 [32,4] | b0 [32] | g0 [32] | be0 [32] | w1 [32,32] | b1 [32] | g1 [32] | be1 [32] | w2 [H,32] | b2 [H]
struct FcOffsets {
    static constexpr int W0 = 0, B0 = 128, G0 = 160, BE0 = 192, W1 = 224, B1 = 1248, G1 = 1280, BE1 = 1312, W2 = 1344;
    __host__ __device__ static int b2(int H) { return W2 + 32 * H; }
    __host__ __device__ static int total(int H) { return W2 + 33 * H; }
};

__global__ void scene_ranges_kernel(const int64_t* __restrict__ bid, int64_t n, int32_t* seg_start,
                                    int32_t* seg_end, int32_t* err) {
    pdl_entry();
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int64_t me = bid[a];
    if (a > 0 && bid[a - 1] > me && err) *err = 1;
    int64_t lo = 0, hi = a;              // first index with bid == me (sorted)
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (bid[mid] < me) lo = mid + 1; else hi = mid; }
    seg_start[a] = (int32_t)lo;
    lo = a; hi = n;                      // first index with bid > me
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (bid[mid] <= me) lo = mid + 1; else hi = mid; }
    seg_end[a] = (int32_t)lo;
}

constexpr int NA_TK = 64;        // keys per tile
constexpr int NA_THREADS = 256;
constexpr int NA_MAXH = 16;

template <int DK>
__global__ void __launch_bounds__(NA_THREADS)
node_attn_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                 const float* __restrict__ v, int64_t ldv, const float* __restrict__ centres, int64_t ldc,
                 const int32_t* __restrict__ seg_start, const int32_t* __restrict__ seg_end,
                 const float* __restrict__ fc, int H, float* __restrict__ out, int64_t ldo, int skip_upto) {
    pdl_entry();
    extern __shared__ __align__(16) float sm[];
    if (seg_end[blockIdx.x] - seg_start[blockIdx.x] <= skip_upto) return;    // served by node_attn_scene_kernel
    float* fcs = sm;                                  // FcOffsets::total(H)
    float* qs = fcs + ((FcOffsets::total(H) + 3) & ~3); // H*DK
    float* hbuf = qs + H * DK;                        // NA_TK * 33
    float* bias_s = hbuf + NA_TK * 33;                // NA_TK * H
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t a = blockIdx.x;
    const int s0 = seg_start[a], s1 = seg_end[a];

    for (int i = tid; i < FcOffsets::total(H); i += NA_THREADS) fcs[i] = __ldg(fc + i);
    for (int i = tid; i < H * DK; i += NA_THREADS) qs[i] = __ldg(q + a * ldq + i);
    const float cax = centres[a * ldc + 0], cay = centres[a * ldc + 1], caz = centres[a * ldc + 2];
    const float scale = rsqrtf((float)DK);

    constexpr int DPL = DK / 32;                      // output dims per lane
    constexpr int NHW = NA_MAXH / 8;                  // heads per warp (upper bound)
    float m_run[NHW], l_run[NHW], acc[NHW][DPL];
#pragma unroll
    for (int i = 0; i < NHW; ++i) {
        m_run[i] = -FLT_MAX; l_run[i] = 0.f;
#pragma unroll
        for (int d = 0; d < DPL; ++d) acc[i][d] = 0.f;
    }
    __syncthreads();

    for (int b0 = s0; b0 < s1; b0 += NA_TK) {
        const int nb = min(NA_TK, s1 - b0);
        // ---- distance-bias MLP for the pairs (a, b0 + pr): 4 threads per pair, 8 hidden units each
        {
            const int pr = tid >> 2, part = tid & 3;
            const bool live = pr < nb;
            const int64_t b = b0 + (live ? pr : 0);
            const float dx = centres[b * ldc + 0] - cax, dy = centres[b * ldc + 1] - cay, dz = centres[b * ldc + 2] - caz;
            const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
            float h[8];
            float s1_ = 0.f, s2_ = 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = part * 8 + u;
                const float* w = fcs + FcOffsets::W0 + j * 4;
                float t = fcs[FcOffsets::B0 + j] + w[0] * dx + w[1] * dy + w[2] * dz + w[3] * dist;
                t = fmaxf(t, 0.f);
                h[u] = t; s1_ += t;
            }
            s1_ += __shfl_xor_sync(0xffffffffu, s1_, 1); s1_ += __shfl_xor_sync(0xffffffffu, s1_, 2);
            float mean = s1_ * (1.f / 32.f);
#pragma unroll
            for (int u = 0; u < 8; ++u) { float d = h[u] - mean; s2_ += d * d; }
            s2_ += __shfl_xor_sync(0xffffffffu, s2_, 1); s2_ += __shfl_xor_sync(0xffffffffu, s2_, 2);
            float rstd = rsqrtf(s2_ * (1.f / 32.f) + 1e-5f);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = part * 8 + u;
                hbuf[pr * 33 + j] = (h[u] - mean) * rstd * fcs[FcOffsets::G0 + j] + fcs[FcOffsets::BE0 + j];
            }
            __syncwarp();
            s1_ = 0.f; s2_ = 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = part * 8 + u;
                const float* w = fcs + FcOffsets::W1 + j * 32;
                float t = fcs[FcOffsets::B1 + j];
#pragma unroll
                for (int i = 0; i < 32; ++i) t = fmaf(w[i], hbuf[pr * 33 + i], t);
                t = fmaxf(t, 0.f);
                h[u] = t; s1_ += t;
            }
            s1_ += __shfl_xor_sync(0xffffffffu, s1_, 1); s1_ += __shfl_xor_sync(0xffffffffu, s1_, 2);
            mean = s1_ * (1.f / 32.f);
#pragma unroll
            for (int u = 0; u < 8; ++u) { float d = h[u] - mean; s2_ += d * d; }
            s2_ += __shfl_xor_sync(0xffffffffu, s2_, 1); s2_ += __shfl_xor_sync(0xffffffffu, s2_, 2);
            rstd = rsqrtf(s2_ * (1.f / 32.f) + 1e-5f);
            __syncwarp();                       // everyone has finished reading layer-1 values
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = part * 8 + u;
                hbuf[pr * 33 + j] = (h[u] - mean) * rstd * fcs[FcOffsets::G1 + j] + fcs[FcOffsets::BE1 + j];
            }
            __syncwarp();
            for (int hh = part; hh < H; hh += 4) {
                const float* w = fcs + FcOffsets::W2 + hh * 32;
                float t = fcs[FcOffsets::b2(H) + hh];
#pragma unroll
                for (int i = 0; i < 32; ++i) t = fmaf(w[i], hbuf[pr * 33 + i], t);
                bias_s[pr * H + hh] = t;
            }
        }
        __syncthreads();
        // ---- scores + online softmax + PV; warp w owns heads w, w+8, ...
#pragma unroll
        for (int hi = 0; hi < NHW; ++hi) {
            const int hh = warp + 8 * hi;
            if (hh >= H) break;
            float sc[NA_TK / 32];
            float tile_max = -FLT_MAX;
#pragma unroll
            for (int r = 0; r < NA_TK / 32; ++r) {
                const int kb = lane + 32 * r;
                float s = -FLT_MAX;
                if (kb < nb) {
                    const float4* kp = reinterpret_cast<const float4*>(k + (int64_t)(b0 + kb) * ldk + hh * DK);
                    const float4* qp = reinterpret_cast<const float4*>(qs + hh * DK);
                    float dot = 0.f;
#pragma unroll
                    for (int d = 0; d < DK / 4; ++d) {
                        const float4 kv = __ldg(kp + d), qv = qp[d];
                        dot = fmaf(kv.x, qv.x, dot); dot = fmaf(kv.y, qv.y, dot);
                        dot = fmaf(kv.z, qv.z, dot); dot = fmaf(kv.w, qv.w, dot);
                    }
                    s = dot * scale + bias_s[kb * H + hh];
                }
                sc[r] = s; tile_max = fmaxf(tile_max, s);
            }
            tile_max = warp_max(tile_max);
            const float m_new = fmaxf(m_run[hi], tile_max);
            const float corr = __expf(m_run[hi] - m_new);
            float psum = 0.f;
#pragma unroll
            for (int r = 0; r < NA_TK / 32; ++r) {
                const int kb = lane + 32 * r;
                sc[r] = (kb < nb) ? __expf(sc[r] - m_new) : 0.f;
                psum += sc[r];
            }
            psum = warp_sum(psum);
            l_run[hi] = l_run[hi] * corr + psum;
            m_run[hi] = m_new;
#pragma unroll
            for (int d = 0; d < DPL; ++d) acc[hi][d] *= corr;
#pragma unroll
            for (int r = 0; r < NA_TK / 32; ++r) {
                // keys in groups of 8: all value loads of a group are issued before the FMAs consume them, so
                // their L2 latencies overlap instead of adding up (a scene has tens of keys, the loop is latency bound)
                for (int j0 = 0; j0 < 32; j0 += 8) {
                    if (j0 + 32 * r >= nb) break;                // warp-uniform
                    float vv[8][DPL], pp[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int kb = j0 + u + 32 * r;
                        pp[u] = __shfl_sync(0xffffffffu, sc[r], j0 + u);
                        const float* vp = v + (int64_t)(b0 + min(kb, nb - 1)) * ldv + hh * DK;
#pragma unroll
                        for (int d = 0; d < DPL; ++d) vv[u][d] = __ldg(vp + lane + 32 * d);
                        if (kb >= nb) pp[u] = 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
#pragma unroll
                        for (int d = 0; d < DPL; ++d) acc[hi][d] = fmaf(pp[u], vv[u][d], acc[hi][d]);
                }
            }
        }
        __syncthreads();   // bias_s / hbuf reused by the next tile
    }
#pragma unroll
    for (int hi = 0; hi < NHW; ++hi) {
        const int hh = warp + 8 * hi;
        if (hh >= H) break;
        const float inv = 1.f / l_run[hi];
#pragma unroll
        for (int d = 0; d < DPL; ++d) out[a * ldo + hh * DK + lane + 32 * d] = acc[hi][d] * inv;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Scene-resident form for scenes of up to NS_MAX nodes (every scene of the 3RScan / BASELINE shapes). The bias
// depends only on the object centres, and MMG.forward runs 2 * depth attentions over the same scenes, so it is
// evaluated ONCE per forward into a head-major table  tab[(h * N + a) * NS_MAX + j]  (query node a, j-th key of its
// scene: the row a (scene, head) CTA reads for one query is 64 contiguous floats);
// each attention call is then one CTA per (scene, head) with that head's Q, K, V slices of the scene in shared
// memory: every K / V row is read once per scene instead of once per query.
constexpr int NS_MAX = 64;
constexpr int NS_THREADS = 128;

__global__ void __launch_bounds__(NA_THREADS)
node_bias_table_kernel(const float* __restrict__ centres, int64_t ldc, const int32_t* __restrict__ seg_start,
                       const int32_t* __restrict__ seg_end, const float* __restrict__ fc, int H, float* __restrict__ tab,
                       int64_t n_nodes) {
    pdl_entry();
    extern __shared__ __align__(16) float sm[];
    float* fcs = sm;                                  // FcOffsets::total(H)
    float* hbuf = fcs + ((FcOffsets::total(H) + 3) & ~3);   // NS_MAX * 33
    const int tid = threadIdx.x;
    const int64_t a = blockIdx.x;
    const int s0 = seg_start[a], ns = seg_end[a] - s0;
    if (ns > NS_MAX) return;                          // large scenes keep the streaming kernel (bias evaluated in place)
    for (int i = tid; i < FcOffsets::total(H); i += NA_THREADS) fcs[i] = __ldg(fc + i);
    const float cax = centres[a * ldc + 0], cay = centres[a * ldc + 1], caz = centres[a * ldc + 2];
    __syncthreads();
    // 4 threads per pair, 8 hidden units each (same arithmetic, in the same order, as node_attn_kernel)
    const int pr = tid >> 2, part = tid & 3;
    const bool live = pr < ns;
    const int64_t b = s0 + (live ? pr : 0);
    const float dx = centres[b * ldc + 0] - cax, dy = centres[b * ldc + 1] - cay, dz = centres[b * ldc + 2] - caz;
    const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
    float h[8];
    float s1_ = 0.f, s2_ = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int j = part * 8 + u;
        const float* w = fcs + FcOffsets::W0 + j * 4;
        float t = fcs[FcOffsets::B0 + j] + w[0] * dx + w[1] * dy + w[2] * dz + w[3] * dist;
        t = fmaxf(t, 0.f);
        h[u] = t; s1_ += t;
    }
    s1_ += __shfl_xor_sync(0xffffffffu, s1_, 1); s1_ += __shfl_xor_sync(0xffffffffu, s1_, 2);
    float mean = s1_ * (1.f / 32.f);
#pragma unroll
    for (int u = 0; u < 8; ++u) { float d = h[u] - mean; s2_ += d * d; }
    s2_ += __shfl_xor_sync(0xffffffffu, s2_, 1); s2_ += __shfl_xor_sync(0xffffffffu, s2_, 2);
    float rstd = rsqrtf(s2_ * (1.f / 32.f) + 1e-5f);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int j = part * 8 + u;
        hbuf[pr * 33 + j] = (h[u] - mean) * rstd * fcs[FcOffsets::G0 + j] + fcs[FcOffsets::BE0 + j];
    }
    __syncwarp();
    s1_ = 0.f; s2_ = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int j = part * 8 + u;
        const float* w = fcs + FcOffsets::W1 + j * 32;
        float t = fcs[FcOffsets::B1 + j];
#pragma unroll
        for (int i = 0; i < 32; ++i) t = fmaf(w[i], hbuf[pr * 33 + i], t);
        t = fmaxf(t, 0.f);
        h[u] = t; s1_ += t;
    }
    s1_ += __shfl_xor_sync(0xffffffffu, s1_, 1); s1_ += __shfl_xor_sync(0xffffffffu, s1_, 2);
    mean = s1_ * (1.f / 32.f);
#pragma unroll
    for (int u = 0; u < 8; ++u) { float d = h[u] - mean; s2_ += d * d; }
    s2_ += __shfl_xor_sync(0xffffffffu, s2_, 1); s2_ += __shfl_xor_sync(0xffffffffu, s2_, 2);
    rstd = rsqrtf(s2_ * (1.f / 32.f) + 1e-5f);
    __syncwarp();                       // everyone has finished reading layer-1 values
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int j = part * 8 + u;
        hbuf[pr * 33 + j] = (h[u] - mean) * rstd * fcs[FcOffsets::G1 + j] + fcs[FcOffsets::BE1 + j];
    }
    __syncwarp();
    if (live) {
        for (int hh = part; hh < H; hh += 4) {
            const float* w = fcs + FcOffsets::W2 + hh * 32;
            float t = fcs[FcOffsets::b2(H) + hh];
#pragma unroll
            for (int i = 0; i < 32; ++i) t = fmaf(w[i], hbuf[pr * 33 + i], t);
            tab[((int64_t)hh * n_nodes + a) * NS_MAX + pr] = t;
        }
    }
}

// grid (ceil(n_nodes / NS_CAND), H): a CTA looks at NS_CAND consecutive nodes and serves, for its head, the scene of
// every one of them that is the first node of its scene (no host-side scene list: the scene count never leaves the GPU).
constexpr int NS_CAND = 16;
template <int DK>
__global__ void __launch_bounds__(NS_THREADS)
node_attn_scene_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                       const float* __restrict__ v, int64_t ldv, const float* __restrict__ tab,
                       const int32_t* __restrict__ seg_start, const int32_t* __restrict__ seg_end, int H,
                       float* __restrict__ out, int64_t ldo, int64_t n_nodes) {
    pdl_entry();
    const int hh = blockIdx.y;
    constexpr int LD = DK + 4;                        // 16-byte aligned rows whose stride is 4 banks off a multiple of 32:
                                                      // eight lanes reading float4s of eight consecutive rows cover all 32 banks
    constexpr int QB = 4;                             // queries a warp carries at once: four independent accumulator chains
                                                      // per lane (the loops are latency-bound, not throughput-bound)
    extern __shared__ __align__(16) float sm[];
    float* qs = sm; float* ks = qs + NS_MAX * LD; float* vs = ks + NS_MAX * LD;
    float* ps = vs + NS_MAX * LD;                     // [warps][QB][NS_MAX] probabilities
    __shared__ int lead[NS_CAND];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // candidate nodes of this CTA that start a scene of at most NS_MAX nodes (all lookups in flight at once)
    if (tid < NS_CAND) {
        const int a = blockIdx.x * NS_CAND + tid;
        int ok = 0;
        if (a < n_nodes) { const int s0 = seg_start[a]; ok = (a == s0 && seg_end[a] - s0 <= NS_MAX) ? 1 : 0; }
        lead[tid] = ok;
    }
    __syncthreads();
    const float scale = rsqrtf((float)DK);
    for (int c = 0; c < NS_CAND; ++c) {
        if (!lead[c]) continue;                       // block-uniform
        const int s0 = blockIdx.x * NS_CAND + c, ns = seg_end[s0] - s0;
        __syncthreads();                              // the previous scene's rows are no longer read
        for (int i = tid; i < ns * (DK / 4); i += NS_THREADS) {
            const int row = i / (DK / 4), c4 = i % (DK / 4);
            *reinterpret_cast<float4*>(qs + row * LD + c4 * 4) = __ldg(reinterpret_cast<const float4*>(q + (int64_t)(s0 + row) * ldq + hh * DK) + c4);
            *reinterpret_cast<float4*>(ks + row * LD + c4 * 4) = __ldg(reinterpret_cast<const float4*>(k + (int64_t)(s0 + row) * ldk + hh * DK) + c4);
            *reinterpret_cast<float4*>(vs + row * LD + c4 * 4) = __ldg(reinterpret_cast<const float4*>(v + (int64_t)(s0 + row) * ldv + hh * DK) + c4);
        }
        __syncthreads();
        float* pw = ps + warp * QB * NS_MAX;
        for (int i0 = warp * QB; i0 < ns; i0 += (NS_THREADS / 32) * QB) {        // QB queries per warp at a time
            float sc[QB][NS_MAX / 32];
            // bias rows (64 contiguous floats per query): in flight during the dot products
#pragma unroll
            for (int u = 0; u < QB; ++u) {
                const float* bias = tab + ((int64_t)hh * n_nodes + (s0 + min(i0 + u, ns - 1))) * NS_MAX;
#pragma unroll
                for (int r = 0; r < NS_MAX / 32; ++r) sc[u][r] = (lane + 32 * r < ns) ? __ldg(bias + lane + 32 * r) : 0.f;
            }
#pragma unroll
            for (int r = 0; r < NS_MAX / 32; ++r) {
                const int j = lane + 32 * r;
                if (r * 32 >= ns) {                              // warp-uniform: no key in this pass
#pragma unroll
                    for (int u = 0; u < QB; ++u) sc[u][r] = -FLT_MAX;
                    continue;
                }
                const float4* kj = reinterpret_cast<const float4*>(ks + min(j, ns - 1) * LD);
                float dot[QB];
#pragma unroll
                for (int u = 0; u < QB; ++u) dot[u] = 0.f;
#pragma unroll
                for (int d = 0; d < DK / 4; ++d) {
                    const float4 kv = kj[d];
#pragma unroll
                    for (int u = 0; u < QB; ++u) {
                        const float4 qv = *reinterpret_cast<const float4*>(qs + min(i0 + u, ns - 1) * LD + 4 * d);    // broadcast
                        dot[u] = fmaf(qv.x, kv.x, dot[u]); dot[u] = fmaf(qv.y, kv.y, dot[u]);
                        dot[u] = fmaf(qv.z, kv.z, dot[u]); dot[u] = fmaf(qv.w, kv.w, dot[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < QB; ++u) sc[u][r] = (j < ns) ? dot[u] * scale + sc[u][r] : -FLT_MAX;
            }
#pragma unroll
            for (int u = 0; u < QB; ++u) {
                float mx = -FLT_MAX;
#pragma unroll
                for (int r = 0; r < NS_MAX / 32; ++r) mx = fmaxf(mx, sc[u][r]);
                mx = warp_max(mx);
                float sum = 0.f;
#pragma unroll
                for (int r = 0; r < NS_MAX / 32; ++r) {
                    sc[u][r] = (lane + 32 * r < ns) ? __expf(sc[u][r] - mx) : 0.f;
                    sum += sc[u][r];
                }
                const float inv = 1.f / warp_sum(sum);
#pragma unroll
                for (int r = 0; r < NS_MAX / 32; ++r) pw[u * NS_MAX + lane + 32 * r] = sc[u][r] * inv;
            }
            __syncwarp();
            float acc[QB][DK / 32];
#pragma unroll
            for (int u = 0; u < QB; ++u)
#pragma unroll
                for (int d = 0; d < DK / 32; ++d) acc[u][d] = 0.f;
            for (int j = 0; j < ns; ++j) {
                float vv[DK / 32];
#pragma unroll
                for (int d = 0; d < DK / 32; ++d) vv[d] = vs[j * LD + lane + 32 * d];
#pragma unroll
                for (int u = 0; u < QB; ++u) {
                    const float pj = pw[u * NS_MAX + j];         // broadcast
#pragma unroll
                    for (int d = 0; d < DK / 32; ++d) acc[u][d] = fmaf(pj, vv[d], acc[u][d]);
                }
            }
#pragma unroll
            for (int u = 0; u < QB; ++u) {
                if (i0 + u < ns) {
#pragma unroll
                    for (int d = 0; d < DK / 32; ++d) out[(int64_t)(s0 + i0 + u) * ldo + hh * DK + lane + 32 * d] = acc[u][d];
                }
            }
            __syncwarp();                                    // pw is rewritten for this warp's next queries
        }
    }
}

}  // namespace vlsat

using namespace vlsat;

extern "C" int vlsat_scene_ranges(const int64_t* batch_ids, int64_t n_nodes, int32_t* seg_start,
                                  int32_t* seg_end, int32_t* err_flag, void* stream) {
    VLSAT_REQUIRE(batch_ids && seg_start && seg_end && n_nodes >= 0);
    VLSAT_SUPPORT(n_nodes < 0x7fffffff);
    if (n_nodes == 0) return VLSAT_OK;
    launch_k(scene_ranges_kernel, dim3((unsigned)ceil_div(n_nodes, 256)), dim3(256), 0, (cudaStream_t)stream, 
        batch_ids, n_nodes, seg_start, seg_end, err_flag);
    return finish_launch();
}

extern "C" int vlsat_node_attn_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v,
                                   int64_t ldv, const float* centres, int64_t ld_centres,
                                   const int32_t* seg_start, const int32_t* seg_end, const float* fc_w,
                                   int n_heads, int dk, float* out, int64_t ldo, int64_t n_nodes, int skip_scenes_upto,
                                   void* stream) {
    VLSAT_REQUIRE(q && k && v && centres && seg_start && seg_end && fc_w && out && n_nodes >= 0);
    VLSAT_SUPPORT(n_heads >= 1 && n_heads <= NA_MAXH && (dk == 64 || dk == 128 || dk == 32));
    VLSAT_SUPPORT(ldq % 4 == 0 && ldk % 4 == 0 && ((uintptr_t)k % 16 == 0) && n_nodes < 0x7fffffff);
    if (n_nodes == 0) return VLSAT_OK;
    const size_t smem = sizeof(float) * (((FcOffsets::total(n_heads) + 3) & ~3) + n_heads * dk + NA_TK * 33 + NA_TK * n_heads);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)n_nodes;
#define LAUNCH(DK_) launch_k(node_attn_kernel<DK_>, grid, dim3(NA_THREADS), smem, st, q, ldq, k, ldk, v, ldv, centres, ld_centres, \
        seg_start, seg_end, fc_w, n_heads, out, ldo, skip_scenes_upto)
    if (dk == 64) LAUNCH(64); else if (dk == 128) LAUNCH(128); else LAUNCH(32);
#undef LAUNCH
    return finish_launch();
}

extern "C" int vlsat_node_bias_table_max_scene(void) { return NS_MAX; }

extern "C" int vlsat_node_bias_table(const float* centres, int64_t ld_centres, const int32_t* seg_start,
                                     const int32_t* seg_end, const float* fc_w, int n_heads, float* table,
                                     int64_t n_nodes, void* stream) {
    VLSAT_REQUIRE(centres && seg_start && seg_end && fc_w && table && n_nodes >= 0);
    VLSAT_SUPPORT(n_heads >= 1 && n_heads <= NA_MAXH && n_nodes < 0x7fffffff);
    if (n_nodes == 0) return VLSAT_OK;
    const size_t smem = sizeof(float) * (((FcOffsets::total(n_heads) + 3) & ~3) + NS_MAX * 33);
    launch_k(node_bias_table_kernel, dim3((unsigned)n_nodes), dim3(NA_THREADS), smem, (cudaStream_t)stream, centres, ld_centres, seg_start, seg_end,
                                                                                         fc_w, n_heads, table, n_nodes);
    return finish_launch();
}

extern "C" int vlsat_node_attn_scene_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v,
                                         int64_t ldv, const float* table, const int32_t* seg_start,
                                         const int32_t* seg_end, int n_heads, int dk, float* out, int64_t ldo,
                                         int64_t n_nodes, void* stream) {
    VLSAT_REQUIRE(q && k && v && table && seg_start && seg_end && out && n_nodes >= 0);
    VLSAT_SUPPORT(n_heads >= 1 && n_heads <= NA_MAXH && (dk == 64 || dk == 32));
    VLSAT_SUPPORT(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) % 16 == 0) &&
                  n_nodes < 0x7fffffff && n_heads <= 65535);
    if (n_nodes == 0) return VLSAT_OK;
    const dim3 grid((unsigned)ceil_div(n_nodes, NS_CAND), (unsigned)n_heads);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = sizeof(float) * (3 * NS_MAX * (dk + 4) + (NS_THREADS / 32) * 4 * NS_MAX);
    if (dk == 64) {
        cudaFuncSetAttribute(node_attn_scene_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        launch_k(node_attn_scene_kernel<64>, grid, dim3(NS_THREADS), smem, st, q, ldq, k, ldk, v, ldv, table, seg_start, seg_end, n_heads, out, ldo, n_nodes);
    } else {
        launch_k(node_attn_scene_kernel<32>, grid, dim3(NS_THREADS), smem, st, q, ldq, k, ldk, v, ldv, table, seg_start, seg_end, n_heads, out, ldo, n_nodes);
    }
    return finish_launch();
}
