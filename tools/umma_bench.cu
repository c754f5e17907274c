// Micro-benchmark: throughput of dependent / independent tcgen05.mma chains with minimal issue overhead
// (descriptors precomputed, fully unrolled, warp-uniform control flow, elect.sync around the issue).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_bench umma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../cvpr2023-vlsat_b200/csrc/tc_common.cuh"
using namespace vlsat::tc;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// MODE: 0 tf32 SS, 1 bf16 SS, 2 tf32 TS.  NACC accumulators used round-robin. 16 MMAs per loop iteration.
template <int MODE, int N, int NACC>
__global__ void __launch_bounds__(128, 1) bench(int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    for (int i = threadIdx.x; i < 6 * 32768 / 4; i += 128) ((float*)smem)[i] = 0.001f * (i % 97);
    fence_proxy_async();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&holder, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = holder;
    if (threadIdx.x < 32) {
        constexpr uint32_t idesc = MODE == 1 ? make_idesc<Kind::BF16>(128, N) : make_idesc<Kind::TF32>(128, N);
        const uint64_t da0 = make_sdesc_k128(smem_u32(smem)), db0 = make_sdesc_k128(smem_u32(smem + 3 * 32768));
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint64_t da = da0 + (uint64_t)((i & 3) * 2 + ((i >> 2) % 3) * 2048);
                    const uint64_t db = db0 + (uint64_t)((i & 3) * 2 + ((i >> 2) % 3) * 2048);
                    const uint32_t d = tm + (i % NACC) * N;
                    if (MODE == 0) mma_ss<Kind::TF32>(d, da, db, idesc, 1);
                    else if (MODE == 1) mma_ss<Kind::BF16>(d, da, db, idesc, 1);
                    else mma_ts(d, tm + 448 + (i & 3) * 8, db, idesc, 1);
                }
            }
            __syncwarp();
        }
        long long t1 = clock64();
        if (elect_one()) tc_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int MODE, int N, int NACC>
void run(const char* name, long long* d) {
    auto k = bench<MODE, N, NACC>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 32768 + 1024);
    const int iters = 64;
    long long h[2];
    for (int rep = 0; rep < 2; ++rep) {
        k<<<1, 128, 6 * 32768 + 1024>>>(iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    }
    printf("%-8s N=%3d acc=%d : issue %.1f clk/MMA, complete %.1f clk/MMA\n", name, N, NACC, (double)h[0] / (iters * 16), (double)h[1] / (iters * 16));
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    run<0, 64, 1>("tf32 SS", d);  run<0, 64, 2>("tf32 SS", d);  run<0, 64, 4>("tf32 SS", d);
    run<0, 128, 1>("tf32 SS", d); run<0, 128, 2>("tf32 SS", d); run<0, 128, 3>("tf32 SS", d);
    run<0, 256, 1>("tf32 SS", d);
    run<1, 64, 1>("bf16 SS", d);  run<1, 128, 1>("bf16 SS", d); run<1, 128, 3>("bf16 SS", d); run<1, 256, 1>("bf16 SS", d);
    run<2, 64, 1>("tf32 TS", d);  run<2, 64, 4>("tf32 TS", d);  run<2, 128, 1>("tf32 TS", d); run<2, 128, 3>("tf32 TS", d);
    run<2, 256, 1>("tf32 TS", d);
    return 0;
}
