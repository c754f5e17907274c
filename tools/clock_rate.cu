// What does clock64() tick at, right after idle and after sustained load? (interprets in-kernel traces)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin(long long ticks, long long* out) {
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
    long long c0 = clock64();
    while (clock64() - c0 < ticks) {}
    long long c1 = clock64();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = c1 - c0; out[1] = (long long)(g1 - g0); }
}
__global__ void burn(float* p, int iters) {
    float a = p[threadIdx.x], b = 1.0001f;
    for (int i = 0; i < iters; ++i) { a = fmaf(a, b, 0.5f); b = fmaf(b, a, 0.25f); }
    p[threadIdx.x + blockIdx.x * blockDim.x] = a + b;
}
int main() {
    long long *d, h[2]; cudaMalloc(&d, 16);
    float* p; cudaMalloc(&p, 148 * 8 * 256 * 4); cudaMemset(p, 0, 148 * 8 * 256 * 4);
    for (int rep = 0; rep < 3; ++rep) {
        spin<<<1, 32>>>(20000, d); cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("after idle : %lld ticks in %lld ns -> %.3f GHz\n", h[0], h[1], (double)h[0] / h[1]);
        spin<<<1, 32>>>(2000000, d); cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("1 ms spin  : %lld ticks in %lld ns -> %.3f GHz\n", h[0], h[1], (double)h[0] / h[1]);
        burn<<<148 * 8, 256>>>(p, 4000000); spin<<<1, 32>>>(20000, d);
        cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("after burn : %lld ticks in %lld ns -> %.3f GHz\n", h[0], h[1], (double)h[0] / h[1]);
    }
    return 0;
}
