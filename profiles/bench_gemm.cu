// Standalone timing of the batched DMMA GEMM (montecarlo.jl_b200/csrc/gemm.cu) on the cfg-4 launch shape
// (296 matrices of 256^3, A = the shared hopping exponential), against cuBLAS strided-batched DGEMM on the same shape.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/bench_gemm profiles/bench_gemm.cu -lcublas -lcuda
#include <cublas_v2.h>
#include <cstdio>
#include <vector>

#include "../montecarlo.jl_b200/csrc/gemm.cu"

namespace dqmc { thread_local long long* t_launch_counter = nullptr; }
using namespace dqmc;

template <class F> static float timeit(F f, int reps = 20)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 256, batch = argc > 2 ? atoi(argv[2]) : 296;
    const int ld = n + (n & 1);
    const long long ms_ = (long long)ld * n;
    // several independent operand sets so that consecutive launches do not hit L2 (inputs larger than L2)
    const int NSET = 4;
    double *A, *B, *C, *v;
    cudaMalloc(&A, ms_ * 8); cudaMalloc(&B, ms_ * 8 * batch * NSET); cudaMalloc(&C, ms_ * 8 * batch * NSET);
    cudaMalloc(&v, (size_t)n * batch * 8);
    std::vector<double> h(ms_ * batch);
    for (size_t i = 0; i < h.size(); ++i) h[i] = 1e-3 * (double)((i * 2654435761u) % 1000) - 0.5;
    cudaMemcpy(A, h.data(), ms_ * 8, cudaMemcpyHostToDevice);
    for (int s = 0; s < NSET; ++s) cudaMemcpy(B + s * ms_ * batch, h.data(), ms_ * 8 * batch, cudaMemcpyHostToDevice);
    for (size_t i = 0; i < (size_t)n * batch; ++i) h[i] = 1.0 + 1e-3 * (double)(i % 7);
    cudaMemcpy(v, h.data(), (size_t)n * batch * 8, cudaMemcpyHostToDevice);
    const double flop = 2.0 * n * n * (double)n * batch;

    GemmParams g{};
    g.M = g.N = g.K = n; g.lda = g.ldb = g.ldc = ld;
    g.A = A; g.strideA = 0; g.transA = 0;
    g.B = B; g.strideB = ms_; g.transB = 0;
    g.C = C; g.strideC = ms_;
    g.alpha = 1.0; g.beta = 0.0; g.rs = no_scale(); g.ks = no_scale(); g.cs = no_scale(); g.batch = batch;
    Scale vs{}; vs.mode = 1; vs.vec = v; vs.stride = n; vs.nb = 1;

    int set = 0;
    auto run = [&](const char* name, GemmParams q) {
        float ms = timeit([&] {
            GemmParams r = q; r.B = B + (long long)set * ms_ * batch; r.C = C + (long long)set * ms_ * batch;
            set = (set + 1) % NSET;
            launch_gemm(r, 0);
        });
        printf("%-28s %7.4f ms  %6.2f TFLOP/s\n", name, ms, flop / ms / 1e9);
    };
    run("plain (A shared)", g);
    { GemmParams q = g; q.ks = vs; run("ks scale", q); }
    { GemmParams q = g; q.rs = vs; run("rs scale", q); }
    { GemmParams q = g; q.cs = vs; run("cs scale", q); }
    { GemmParams q = g; q.A = B + 3 * ms_ * batch; q.strideA = ms_; q.transB = 1; run("A per chain, B^T", q); }

    cublasHandle_t hd; cublasCreate(&hd);
    const double one = 1.0, zero = 0.0;
    float ms = timeit([&] {
        const double* Bp = B + (long long)set * ms_ * batch; double* Cp = C + (long long)set * ms_ * batch;
        set = (set + 1) % NSET;
        cublasDgemmStridedBatched(hd, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, ld, 0, Bp, ld, ms_, &zero, Cp, ld, ms_, batch);
    });
    printf("%-28s %7.4f ms  %6.2f TFLOP/s\n", "cuBLAS strided batched", ms, flop / ms / 1e9);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
