// FP64 issue-rate microbenchmark for sm_100a: DMMA.8x8x4 vs DFMA, register-resident, no memory traffic.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_fp64 microbench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double a0, double b0)
{
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a0, double b0)
{
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float timeit(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main()
{
    double* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        for (int ctas : {1, 2}) {
            if (warps * ctas > 64) continue;
            int grid = 148 * ctas, block = warps * 32;
            float ms = timeit([&] { k_dmma<16><<<grid, block>>>(out, iters, 1.0, 1.0); });
            double fl = 2.0 * 256 * 16 * (double)iters * warps * grid;
            printf("DMMA  nacc=16 warps/cta=%2d ctas/sm=%d : %7.2f TFLOP/s\n", warps, ctas, fl / ms / 1e9);
            ms = timeit([&] { k_dmma<4><<<grid, block>>>(out, iters, 1.0, 1.0); });
            fl = 2.0 * 256 * 4 * (double)iters * warps * grid;
            printf("DMMA  nacc= 4 warps/cta=%2d ctas/sm=%d : %7.2f TFLOP/s\n", warps, ctas, fl / ms / 1e9);
            ms = timeit([&] { k_dfma<16><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
            fl = 2.0 * 32 * 16 * (double)iters * warps * grid;
            printf("DFMA  nacc=16 warps/cta=%2d ctas/sm=%d : %7.2f TFLOP/s\n", warps, ctas, fl / ms / 1e9);
        }
    }
    return 0;
}
