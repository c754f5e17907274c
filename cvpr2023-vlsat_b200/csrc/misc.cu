// Small memory-bound kernels of the path: edge descriptor (A2), residual LayerNorm (A7 tail),
// inter-layer ReLU (A10), row L2 normalisation (A13) and the spatial tail of the 3-D node feature (A4).
#include "common.cuh"
#include "epilogue.cuh"

namespace vlsat {

__global__ void edge_descriptor_kernel(const float* __restrict__ desc, const int64_t* __restrict__ ei,
                                       int64_t n_edges, float* __restrict__ out) {
    pdl_entry();
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
    pdl_entry();
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

// D = 128 * NV: each lane owns NV float4 groups (columns 4*lane + 128*i ..), every access is a full 512-byte warp row.
// Optionally also writes the bf16 (hi, lo) pair of the result (the operand format of a following projection).
template <int NV>
__global__ void add_layernorm_vec_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ res, int64_t ldr,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float* __restrict__ y, int64_t ldy, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                         int64_t ld_split, int64_t M, float eps, int relu) {
    pdl_entry();
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    constexpr int D = 128 * NV;
    float4 v[NV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = __ldg(reinterpret_cast<const float4*>(x + row * ldx + 128 * i) + lane);
        if (res) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(res + row * ldr + 128 * i) + lane);
            v[i].x += r.x; v[i].y += r.y; v[i].z += r.z; v[i].w += r.w;
        }
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(sum) / (float)D;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + 128 * i) + lane), b = __ldg(reinterpret_cast<const float4*>(beta + 128 * i) + lane);
        float4 t = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                               (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
        if (relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
        if (y) *(reinterpret_cast<float4*>(y + row * ldy + 128 * i) + lane) = t;
        if (hi) {
            uint32_t h0, l0, h1, l1;
            split_bf16x2(t.x, t.y, h0, l0); split_bf16x2(t.z, t.w, h1, l1);
            *(reinterpret_cast<uint2*>(hi + row * ld_split + 128 * i) + lane) = make_uint2(h0, h1);
            *(reinterpret_cast<uint2*>(lo + row * ld_split + 128 * i) + lane) = make_uint2(l0, l1);
        }
    }
}

// y = relu(x) with the bf16 (hi, lo) pair of y as an optional second output (flat, numel % 4 == 0)
__global__ void relu_pair_kernel(const float* __restrict__ x, float* __restrict__ y, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int64_t n4) {
    pdl_entry();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    if (y) reinterpret_cast<float4*>(y)[i] = v;
    uint32_t h0, l0, h1, l1;
    split_bf16x2(v.x, v.y, h0, l0); split_bf16x2(v.z, v.w, h1, l1);
    reinterpret_cast<uint2*>(hi)[i] = make_uint2(h0, h1);
    reinterpret_cast<uint2*>(lo)[i] = make_uint2(l0, l1);
}

__global__ void relu_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n4, int64_t n) {
    pdl_entry();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        float4 v = reinterpret_cast<const float4*>(x)[i];
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        reinterpret_cast<float4*>(y)[i] = v;
    }
    if (i == 0) for (int64_t j = n4 * 4; j < n; ++j) y[j] = fmaxf(x[j], 0.f);
}
__global__ void relu_scalar_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
    pdl_entry();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = fmaxf(x[i], 0.f);
}

__global__ void row_l2norm_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t M, int D) {
    pdl_entry();
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
    pdl_entry();
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
    launch_k(edge_descriptor_kernel, dim3((unsigned)ceil_div(n_edges * 11, 256)), dim3(256), 0, (cudaStream_t)stream, desc, edge_index, n_edges, out);
    return finish_launch();
}

extern "C" int vlsat_add_layernorm_fwd(const float* x, int64_t ldx, const float* res, int64_t ld_res,
                                       const float* gamma, const float* beta, float* y, int64_t ldy,
                                       int64_t M, int D, float eps, int relu, void* split_hi, void* split_lo,
                                       int64_t ld_split, void* stream) {
    VLSAT_REQUIRE(M >= 0 && D >= 1);
    if (M == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && gamma && beta && ldx >= D && (!y || ldy >= D) && (!res || ld_res >= D));
    VLSAT_SUPPORT(D <= 32 * LN_MAX_PER_LANE);
    VLSAT_REQUIRE((split_hi == nullptr) == (split_lo == nullptr) && (!split_hi || ld_split >= D));
    VLSAT_REQUIRE(y || split_hi);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)ceil_div(M * 32, 256);
    auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
    const bool vec = D % 128 == 0 && D <= 1024 && ldx % 4 == 0 && ldy % 4 == 0 && (!res || (ld_res % 4 == 0 && al16(res))) && al16(x) &&
                     (!y || al16(y)) && al16(gamma) && al16(beta) &&
                     (!split_hi || (ld_split % 4 == 0 && ((uintptr_t)split_hi & 7) == 0 && ((uintptr_t)split_lo & 7) == 0));
    uint16_t* hi = (uint16_t*)split_hi; uint16_t* lo = (uint16_t*)split_lo;
#define VLSAT_LN(NV_) launch_k(add_layernorm_vec_kernel<NV_>, grid, dim3(256), 0, st, x, ldx, res, ld_res, gamma, beta, y, ldy, hi, lo, ld_split, M, eps, relu)
    if (vec) {
        switch (D / 128) {
            case 1: VLSAT_LN(1); break; case 2: VLSAT_LN(2); break; case 3: VLSAT_LN(3); break; case 4: VLSAT_LN(4); break;
            case 5: VLSAT_LN(5); break; case 6: VLSAT_LN(6); break; case 7: VLSAT_LN(7); break; default: VLSAT_LN(8); break;
        }
        return finish_launch();
    }
#undef VLSAT_LN
    VLSAT_SUPPORT(!split_hi && y);            // the pair output needs the vectorised layout
    launch_k(add_layernorm_kernel, grid, dim3(256), 0, st, x, ldx, res, ld_res, gamma, beta, y, ldy, M, D, eps, relu);
    return finish_launch();
}

extern "C" int vlsat_relu_fwd(const float* x, float* y, int64_t numel, void* stream) {
    VLSAT_REQUIRE(numel >= 0);
    if (numel == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && y);
    if (((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0)) {
        const int64_t n4 = numel / 4;
        launch_k(relu_kernel, dim3((unsigned)ceil_div(n4 > 0 ? n4 : 1, 256)), dim3(256), 0, (cudaStream_t)stream, x, y, n4, numel);
    } else {
        launch_k(relu_scalar_kernel, dim3((unsigned)ceil_div(numel, 256)), dim3(256), 0, (cudaStream_t)stream, x, y, numel);
    }
    return finish_launch();
}

extern "C" int vlsat_relu_pair_fwd(const float* x, float* y, void* split_hi, void* split_lo, int64_t numel, void* stream) {
    VLSAT_REQUIRE(numel >= 0);
    if (numel == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && split_hi && split_lo);
    VLSAT_SUPPORT(numel % 4 == 0 && ((uintptr_t)x % 16 == 0) && (!y || (uintptr_t)y % 16 == 0) && ((uintptr_t)split_hi % 8 == 0) &&
                  ((uintptr_t)split_lo % 8 == 0));
    launch_k(relu_pair_kernel, dim3((unsigned)ceil_div(numel / 4, 256)), dim3(256), 0, (cudaStream_t)stream, x, y, (uint16_t*)split_hi, (uint16_t*)split_lo, numel / 4);
    return finish_launch();
}

extern "C" int vlsat_row_l2norm_fwd(const float* x, float* y, int64_t M, int D, void* stream) {
    VLSAT_REQUIRE(M >= 0 && D >= 1);
    if (M == 0) return VLSAT_OK;
    VLSAT_REQUIRE(x && y);
    launch_k(row_l2norm_kernel, dim3((unsigned)ceil_div(M * 32, 256)), dim3(256), 0, (cudaStream_t)stream, x, y, M, D);
    return finish_launch();
}

extern "C" int vlsat_spatial_tail_fwd(const float* desc, float* out, int64_t ld_out, int col0, int64_t n_nodes, void* stream) {
    VLSAT_REQUIRE(n_nodes >= 0);
    if (n_nodes == 0) return VLSAT_OK;
    VLSAT_REQUIRE(desc && out && col0 >= 0 && col0 + 8 <= ld_out);
    launch_k(spatial_tail_kernel, dim3((unsigned)ceil_div(n_nodes * 8, 256)), dim3(256), 0, (cudaStream_t)stream, desc, out, ld_out, col0, n_nodes);
    return finish_launch();
}
