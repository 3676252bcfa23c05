// Microbenchmark: MUFU.EX2 throughput per SM on B200 as a function of resident warps (test infrastructure).
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int ILP>
__global__ void mufu_kernel(float* out, int iters, float seed) {
    float v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = ex2(v[i]) - 1.0f;   // 1 MUFU + 1 FADD per element
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 12345.678f) out[0] = s;
}

template <int ILP>
__global__ void fma_kernel(float* out, int iters, float seed) {
    float v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = seed + i * 0.001f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fmaf(v[i], 0.999f, 0.001f);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 12345.678f) out[0] = s;
}

int main() {
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int iters = 4096;
    for (int warps_per_sm : {4, 8, 16, 32, 64}) {
        const int threads = warps_per_sm * 32 > 1024 ? 1024 : warps_per_sm * 32;
        const int blocks_per_sm = warps_per_sm * 32 / threads;
        const int grid = 148 * blocks_per_sm;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            mufu_kernel<16><<<grid, threads>>>(d, iters, 0.5f);
            cudaEventRecord(b); cudaEventSynchronize(b);
        }
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double ops = (double)grid * threads * iters * 16;
        printf("mufu warps/SM=%2d  %.3f ms  %.2f G ex2/s  -> %.2f ex2/clk/SM at max clock %d MHz\n", warps_per_sm, ms,
               ops / ms / 1e6, ops / (ms * 1e-3) / 148 / (clk_khz * 1e3), clk_khz / 1000);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            fma_kernel<16><<<grid, threads>>>(d, iters, 0.5f);
            cudaEventRecord(b); cudaEventSynchronize(b);
        }
        cudaEventElapsedTime(&ms, a, b);
        printf("fma  warps/SM=%2d  %.3f ms  -> %.2f fma/clk/SM\n", warps_per_sm, ms, ops / (ms * 1e-3) / 148 / (clk_khz * 1e3));
    }
    return 0;
}
