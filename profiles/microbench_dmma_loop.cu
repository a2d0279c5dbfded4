// Inner-loop ceiling of the DMMA GEMM on sm_100a: fragments from shared memory exactly like gemm.cu,
// no global traffic.  Variants: MI x NJ warp tiles, fragment double buffering.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MI, int NJ, bool DBUF, bool REGONLY>
__global__ void __launch_bounds__(128) k_loop(double* out, int iters)
{
    __shared__ double As[16 * 68], Bs[64 * 20];
    for (int i = threadIdx.x; i < 16 * 68; i += 128) As[i] = 1.0 + i * 1e-6;
    for (int i = threadIdx.x; i < 64 * 20; i += 128) Bs[i] = 1.0 - i * 1e-6;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int wm0 = (warp & 1) * 32, wn0 = (warp >> 1) * 32;
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double af[2][MI], bf[2][NJ];
    auto ld = [&](int buf, int kk) {
#pragma unroll
        for (int i = 0; i < MI; ++i) af[buf][i] = As[(kk * 4 + t) * 68 + ((wm0 + i * 8 + g) & 63)];
#pragma unroll
        for (int j = 0; j < NJ; ++j) bf[buf][j] = Bs[((wn0 + j * 8 + g) & 63) * 20 + kk * 4 + t];
    };
    if (REGONLY) { ld(0, 0); ld(1, 1); }
    for (int it = 0; it < iters; ++it) {
        if (!REGONLY && DBUF) ld(0, 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int cur = DBUF ? (kk & 1) : 0;
            if (!REGONLY) {
                if (DBUF) { if (kk < 3) ld(cur ^ 1, kk + 1); }
                else ld(0, kk);
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * 128 + threadIdx.x] = s;
}

template <class F> float timeit(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

template <int MI, int NJ, bool DBUF, bool REGONLY> void run(const char* name, double* out)
{
    const int iters = 4000;
    for (int ctas : {1, 2, 3, 4}) {
        float ms = timeit([&] { k_loop<MI, NJ, DBUF, REGONLY><<<148 * ctas, 128>>>(out, iters); });
        double fl = 2.0 * 256 * MI * NJ * 4.0 * iters * 4 * 148 * ctas;
        printf("%-28s MIxNJ=%dx%d ctas/sm=%d : %6.2f TFLOP/s\n", name, MI, NJ, ctas, fl / ms / 1e9);
    }
}

int main()
{
    double* out; cudaMalloc(&out, 148 * 8 * 128 * sizeof(double));
    run<4, 4, false, true>("registers only", out);
    run<4, 4, false, false>("smem, single-buffered frags", out);
    run<4, 4, true, false>("smem, double-buffered frags", out);
    run<2, 8, false, false>("smem, single", out);
    run<8, 2, false, false>("smem, single", out);
    run<2, 4, false, false>("smem, single", out);
    run<4, 2, true, false>("smem, double", out);
    run<4, 8, false, true>("registers only", out);
    run<4, 8, false, false>("smem, single", out);
    run<4, 8, true, false>("smem, double", out);
    return 0;
}
