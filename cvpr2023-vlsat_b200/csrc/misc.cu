// Small memory-bound kernels of the path: edge descriptor (A2), residual LayerNorm (A7 tail),
// inter-layer ReLU (A10), row L2 normalisation (A13) and the spatial tail of the 3-D node feature (A4).
#include "common.cuh"

namespace vlsat {

__global__ void edge_descriptor_kernel(const float* __restrict__ desc, const int64_t* __restrict__ ei,
                                       int64_t n_edges, float* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_edges * 11) return;
    const int64_t e = idx / 11; const int c = (int)(idx % 11);
    const float a = __ldg(desc + ei[e] * 11 + c), b = __ldg(desc + ei[n_edges + e] * 11 + c);
    out[idx] = (c < 6) ? (a - b) : logf(a / b);
}

// one warp per row; D <= 32 * LN_MAX_PER_LANE
constexpr int LN_MAX_PER_LANE = 32;
__global__ void add_layernorm_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ res, int64_t ldr,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float* __restrict__ y, int64_t ldy, int64_t M, int D, float eps, int relu) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    float vals[LN_MAX_PER_LANE];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
        const int c = lane + 32 * i;
        float t = 0.f;
        if (c < D) { t = x[row * ldx + c]; if (res) t += res[row * ldr + c]; }
        vals[i] = t; sum += t;
    }
    sum = warp_sum(sum);
    const float mean = sum / (float)D;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
        const int c = lane + 32 * i;
        if (c < D) { const float d = vals[i] - mean; sq += d * d; }
    }
    sq = warp_sum(sq);
    const float rstd = rsqrtf(sq / (float)D + eps);
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
        const int c = lane + 32 * i;
        if (c < D) {
            float t = (vals[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
            if (relu) t = fmaxf(t, 0.f);
            y[row * ldy + c] = t;
        }
    }
}

__global__ void relu_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n4, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        float4 v = reinterpret_cast<const float4*>(x)[i];
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        reinterpret_cast<float4*>(y)[i] = v;
    }
    if (i == 0) for (int64_t j = n4 * 4; j < n; ++j) y[j] = fmaxf(x[j], 0.f);
}
__global__ void relu_scalar_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = fmaxf(x[i], 0.f);
}

__global__ void row_l2norm_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t M, int D) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    float sq = 0.f;
    for (int c = lane; c < D; c += 32) { const float t = x[row * D + c]; sq += t * t; }
    sq = warp_sum(sq);
    const float inv = 1.f / sqrtf(sq);
    for (int c = lane; c < D; c += 32) y[row * D + c] = x[row * D + c] * inv;
}

__global__ void spatial_tail_kernel(const float* __restrict__ desc, float* __restrict__ out, int64_t ld, int col0, int64_t n) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * 8) return;
    const int64_t r = idx / 8; const int c = (int)(idx % 8);
    const float d = desc[r * 11 + 3 + c];
    out[r * ld + col0 + c] = (c < 6) ? d : logf(d);
}

}  // namespace vlsat

using namespace vlsat;

extern "C" int vlsat_edge_descriptor_fwd(const float* desc, int64_t n_nodes, const int64_t* edge_index,
                                         int64_t n_edges, float* out, void* stream) {
    VLSAT_REQUIRE(n_edges >= 0 && n_nodes >= 0);
    if (n_edges == 0) return VLSAT_OK;
    VLSAT_REQUIRE(desc && edge_index && out);
    edge_descriptor_kernel<<<(unsigned)ceil_div(n_edges * 11, 256), 256, 0, (cudaStream_t)stream>>>(desc, edge_index, n_edges, out);
    return finish_launch();
}

extern "C" int vlsat_add_layernorm_fwd(const float* x, int64_t ldx, const float* res, int64_t ld_res,
                                       const float* gamma, const float* beta, float* y, int64_t ldy,
                                       int64_t M, int D, float eps, int relu, void* stream) {
    VLSAT_REQUIRE(M >= 0 && D >= 1);
    if (M == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && gamma && beta && y && ldx >= D && ldy >= D && (!res || ld_res >= D));
    VLSAT_SUPPORT(D <= 32 * LN_MAX_PER_LANE);
    add_layernorm_kernel<<<(unsigned)ceil_div(M * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, res, ld_res, gamma, beta, y, ldy, M, D, eps, relu);
    return finish_launch();
}

extern "C" int vlsat_relu_fwd(const float* x, float* y, int64_t numel, void* stream) {
    VLSAT_REQUIRE(numel >= 0);
    if (numel == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && y);
    if (((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0)) {
        const int64_t n4 = numel / 4;
        relu_kernel<<<(unsigned)ceil_div(n4 > 0 ? n4 : 1, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n4, numel);
    } else {
        relu_scalar_kernel<<<(unsigned)ceil_div(numel, 256), 256, 0, (cudaStream_t)stream>>>(x, y, numel);
    }
    return finish_launch();
}

extern "C" int vlsat_row_l2norm_fwd(const float* x, float* y, int64_t M, int D, void* stream) {
    VLSAT_REQUIRE(M >= 0 && D >= 1);
    if (M == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && y);
    row_l2norm_kernel<<<(unsigned)ceil_div(M * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, y, M, D);
    return finish_launch();
}

extern "C" int vlsat_spatial_tail_fwd(const float* desc, float* out, int64_t ld_out, int col0, int64_t n_nodes, void* stream) {
    VLSAT_REQUIRE(n_nodes >= 0);
    if (n_nodes == 0) return VLSAT_OK;
    VLSAT_REQUIRE(desc && out && col0 >= 0 && col0 + 8 <= ld_out);
    spatial_tail_kernel<<<(unsigned)ceil_div(n_nodes * 8, 256), 256, 0, (cudaStream_t)stream>>>(desc, out, ld_out, col0, n_nodes);
    return finish_launch();
}
