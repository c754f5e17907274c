// Shared helpers for the vlsat_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/vlsat_b200.h"

namespace vlsat {

extern long long g_launch_count;   // bench bookkeeping: kernels launched through the C ABI

inline int finish_launch(int n_kernels = 1) {
    g_launch_count += n_kernels;
    return cudaGetLastError() == cudaSuccess ? VLSAT_OK : VLSAT_ERR_LAUNCH;
}

constexpr int kNumSMs = 148;   // B200

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace vlsat

#define VLSAT_REQUIRE(cond) do { if (!(cond)) return VLSAT_ERR_INVALID_ARG; } while (0)
#define VLSAT_SUPPORT(cond) do { if (!(cond)) return VLSAT_ERR_UNSUPPORTED; } while (0)
