// Microbenchmark: tcgen05.ld throughput of one SM (bytes per clock), for the roofline of the A9 kernels' softmax / dS passes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tmem_ld_bench tools/tmem_ld_bench.cu ; run on a B200.
// Each warp reads 32 lanes x 32 columns x 4 B = 4 KB per tcgen05.ld.32x32b.x32; W warps (W = 4, 8, 16) loop ITER times.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) k(int iters, int x16, long long* out, uint32_t* sink) {
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&holder)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = holder + ((uint32_t)((warp & 3) * 32) << 16) + 64u * ((warp >> 2) & 3);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t r[32];
        if (x16) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                           "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(base + 32u * (i & 1)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += r[0] ^ r[15];
        } else {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                           "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                           "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                           "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(base + 32u * (i & 1)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += r[0] ^ r[31];
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(holder), "r"(512u) : "memory");
}

int main() {
    long long* out; uint32_t* sink;
    cudaMalloc(&out, 148 * sizeof(long long)); cudaMalloc(&sink, 148 * 512 * 4);
    const int iters = 20000;
    for (int x16 = 0; x16 < 2; ++x16)
        for (int warps : {4, 8, 16}) {
            k<<<148, warps * 32>>>(iters, x16, out, sink);
            k<<<148, warps * 32>>>(iters, x16, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            const double bytes = (double)iters * warps * 32 * (x16 ? 16 : 32) * 4;
            printf("{\"shape\": \"32x32b.x%d\", \"warps\": %d, \"clk\": %lld, \"bytes_per_clk_per_sm\": %.1f, \"err\": \"%s\"}\n", x16 ? 16 : 32, warps, h[0],
                   bytes / (double)h[0], cudaGetErrorString(e));
        }
    return 0;
}
