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
int tc_passes();               // 3 = BF16x3 (fp32 parity), 1 = single-pass bf16 (vlsat_set_precision, csrc/capi.cu)

// Programmatic dependent launch. Every kernel of the forward path is launched with the programmatic-stream-serialisation
// attribute and starts with pdl_launch_dependents() (the next kernel may be scheduled as soon as all CTAs of this one have
// started) and, after its purely local prologue (barrier init, TMEM allocation, descriptor prefetch), pdl_wait(), which
// blocks until the preceding grid has completed and its memory is visible. Launch latency and prologues of the ~120
// kernels of a forward overlap the tail of their predecessors. The wait precedes every global access and every early
// exit, so completion stays transitive along the chain. VLSAT_PDL=0 switches the attribute off (the instructions are then
// no-ops).
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() { pdl_launch_dependents(); pdl_wait(); }      // kernels without a local prologue

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);       // errors surface in finish_launch()
}

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
