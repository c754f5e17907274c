// SURVEY 8(f) N2, the device-side part of the input pipeline: per-object point gathering, descriptor and centring.
// Replaces, per object, src/dataset/dataset_3dssg.py:288-293 (obj_pointset = points[choice]; gen_descriptor; zero_mean),
// src/utils/op_utils.py:47-64 (gen_descriptor: centroid, unbiased std, extent, volume, longest side) and the
// permute(0, 2, 1).contiguous() of src/model/model.py:71 - one kernel from the cached scan cloud straight to the
// channels-first [N, C, P] tensor PointNetfeat reads and the [N, 11] descriptor, instead of a python loop over objects on
// the loader's CPU. One CTA per object; reductions in a fixed order (bitwise reproducible); the variance is a second
// pass around the mean (scan coordinates sit metres away from the origin, a one-pass E[x^2] - E[x]^2 would cancel).
// HBM: 8 B of index + C * 4 B gathered + C * 4 B written per sampled point.
#include "common.cuh"
#include <float.h>

namespace vlsat {

constexpr int OP_THREADS = 128;

__device__ __forceinline__ float block_reduce(float v, int op, float* red) {      // op 0 sum, 1 min, 2 max; all threads get the result
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float u = __shfl_xor_sync(0xffffffffu, v, o);
        v = op == 0 ? v + u : (op == 1 ? fminf(v, u) : fmaxf(v, u));
    }
    __syncthreads();                                     // red[] of the previous reduction has been read
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < OP_THREADS / 32; ++w) r = op == 0 ? r + red[w] : (op == 1 ? fminf(r, red[w]) : fmaxf(r, red[w]));
    return r;
}

__global__ void __launch_bounds__(OP_THREADS)
object_prep_kernel(const float* __restrict__ cloud, int64_t ld, int64_t n_cloud, const int64_t* __restrict__ choice,
                   int64_t P, int C, float* __restrict__ obj_points, float* __restrict__ descriptor) {
    pdl_entry();
    __shared__ float red[OP_THREADS / 32];
    const int64_t o = blockIdx.x;
    const int64_t* ch = choice + o * P;
    float s[3] = {0.f, 0.f, 0.f}, lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int64_t p = threadIdx.x; p < P; p += OP_THREADS) {
        int64_t i = ch[p];
        i = i < 0 ? 0 : (i >= n_cloud ? n_cloud - 1 : i);          // out-of-range indices are clamped, never read out of bounds
        const float* row = cloud + i * ld;
#pragma unroll
        for (int c = 0; c < 3; ++c) { const float v = row[c]; s[c] += v; lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v); }
    }
    float mean[3], dims[3], sd[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        mean[c] = block_reduce(s[c], 0, red) / (float)P;
        dims[c] = block_reduce(hi[c], 2, red) - block_reduce(lo[c], 1, red);
    }
    float q[3] = {0.f, 0.f, 0.f};
    for (int64_t p = threadIdx.x; p < P; p += OP_THREADS) {
        int64_t i = ch[p];
        i = i < 0 ? 0 : (i >= n_cloud ? n_cloud - 1 : i);
        const float* row = cloud + i * ld;
#pragma unroll
        for (int c = 0; c < 3; ++c) { const float d = row[c] - mean[c]; q[c] += d * d; }
        // channels-first output, threads along the point axis: coalesced stores
        for (int c = 0; c < C; ++c) obj_points[(o * C + c) * P + p] = row[c] - (c < 3 ? mean[c] : 0.f);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) sd[c] = sqrtf(block_reduce(q[c], 0, red) / (float)(P - 1));      // torch.std: unbiased; P == 1 -> NaN
    if (threadIdx.x == 0) {
        float* d = descriptor + o * 11;
        d[0] = mean[0]; d[1] = mean[1]; d[2] = mean[2];
        d[3] = sd[0]; d[4] = sd[1]; d[5] = sd[2];
        d[6] = dims[0]; d[7] = dims[1]; d[8] = dims[2];
        d[9] = dims[0] * dims[1] * dims[2];
        d[10] = fmaxf(dims[0], fmaxf(dims[1], dims[2]));
    }
}

}  // namespace vlsat

using namespace vlsat;

extern "C" int vlsat_object_prep_fwd(const float* cloud, int64_t ld_cloud, int64_t n_cloud, int n_channels, const int64_t* choice,
                                     int64_t n_obj, int64_t n_pts, float* obj_points, float* descriptor, void* stream) {
    VLSAT_REQUIRE(n_obj >= 0 && n_pts >= 1 && n_cloud >= 1 && n_channels >= 3 && ld_cloud >= n_channels);
    if (n_obj == 0) return VLSAT_OK;
    VLSAT_REQUIRE(cloud && choice && obj_points && descriptor);
    VLSAT_SUPPORT(n_obj < (1ll << 31));
    launch_k(object_prep_kernel, dim3((unsigned)n_obj), dim3(OP_THREADS), 0, (cudaStream_t)stream, cloud, ld_cloud, n_cloud, choice,
             n_pts, n_channels, obj_points, descriptor);
    return finish_launch();
}
