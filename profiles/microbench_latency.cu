// Dependent-issue latencies on sm_100a (one warp, clock64 around a chain of N dependent instructions).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_latency.bin microbench_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k(double* out, long long* cyc, double a, double b)
{
    __shared__ int ism[64];
    ism[threadIdx.x & 63] = (threadIdx.x + 1) & 31;
    __syncthreads();
    double x = a; long long t0, t1; int r = 0;
    // DFMA
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(b), "d"(a));
    t1 = clock64(); if (threadIdx.x == 0) cyc[r] = t1 - t0; ++r;
    // DADD
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(b));
    t1 = clock64(); if (threadIdx.x == 0) cyc[r] = t1 - t0; ++r;
    // DMUL
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(b));
    t1 = clock64(); if (threadIdx.x == 0) cyc[r] = t1 - t0; ++r;
    // SHFL.64 (two 32-bit shuffles) + dependent
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1);
    t1 = clock64(); if (threadIdx.x == 0) cyc[r] = t1 - t0; ++r;
    // LDS pointer chase (32-bit)
    int p = threadIdx.x & 31;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) p = ism[p];
    t1 = clock64(); if (threadIdx.x == 0) cyc[r] = t1 - t0; ++r;
    // DSETP + select chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("{ .reg .pred q; setp.gt.f64 q, %0, %1; selp.f64 %0, %2, %0, q; }" : "+d"(x) : "d"(b), "d"(a));
    t1 = clock64(); if (threadIdx.x == 0) cyc[r] = t1 - t0; ++r;
    // FFMA for reference
    float f = (float)a, fb = (float)b;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) f = fmaf(f, fb, fb);
    t1 = clock64(); if (threadIdx.x == 0) cyc[r] = t1 - t0; ++r;
    // __syncthreads round trip (blockDim warps)
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) __syncthreads();
    t1 = clock64(); if (threadIdx.x == 0) cyc[r] = t1 - t0; ++r;
    // ballot + ffs chain
    unsigned m = 0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) { m = __ballot_sync(0xffffffffu, (p + m) & 1); p += __ffs(m); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[r] = t1 - t0; ++r;
    out[threadIdx.x] = x + p + f + m;
}
int main()
{
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 16);
    const char* nm[] = {"DFMA", "DADD", "DMUL", "SHFL.64", "LDS chase", "DSETP+SELP", "FFMA", "BAR.SYNC x64 (/64)", "ballot+ffs"};
    for (int warps : {1, 2, 8}) {
        for (int rep = 0; rep < 2; ++rep) k<<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999);
        long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("warps/CTA = %d:", warps);
        for (int i = 0; i < 9; ++i) printf("  %s %.1f", nm[i], (double)h[i] / (i == 7 ? 64 : N));
        printf("\n");
    }
    return 0;
}
