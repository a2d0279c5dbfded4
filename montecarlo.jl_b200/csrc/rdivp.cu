// rdivp.cu -- batched pivoted right-division by an upper-triangular matrix.
//
// Replaces `rdivp!(A, T, O, pivot)` (reference src/flavors/DQMC/linalg/real.jl:198-226,
// per block blockdiagonal.jl:286-324):  A <- A[:, pivot] * inv(triu(T)).
//
// Blocked forward substitution over 32-column panels: the off-diagonal part of every
// panel is a DMMA GEMM (O_J -= X_{<J} * T[<J, J], the level-3 bulk, ~n^3 flops), only
// the 32 x 32 diagonal blocks are solved by substitution, one thread per row with the
// row of the panel held in registers and the diagonal block in shared memory.
// Same operation order per row as the reference (subtract, then divide by T[j,j]).
//
// Measured and not kept (round 2): ONE kernel per call -- X = A[:, pivot] T^-1 is independent per row of A, so a CTA of
// four warps owned 32 rows for the whole substitution (row block gathered through the pivot into shared memory, DMMA for
// the off-diagonal part of each panel with T read through L1 / L2, the diagonal block solved by one warp).  1.08 ms against
// 0.47 ms for the 17 launches here: the eight panels become one dependent chain per CTA (gather -> 8 x (DMMA k loop on
// L2 latency -> 32 divisions in sequence)) at two CTAs per SM, while the per-panel launches run every row of every matrix
// of the batch at once.
#include "common.cuh"

namespace dqmc {

constexpr int JB = 32;

__global__ void __launch_bounds__(128)
trsm_diag_kernel(double* X, const double* O, const double* T, int n, int ld, int j0, int jb,
                 long long strideA, long long strideT, long long strideW)
{
    __shared__ double Ts[JB][JB + 1];
    const int mat = blockIdx.y;
    const double* t = T + (long long)mat * strideT;
    const double* o = O + (long long)mat * strideW;
    double* x = X + (long long)mat * strideA;
    for (int e = threadIdx.x; e < JB * JB; e += blockDim.x) {
        const int k = e % JB, c = e / JB;
        Ts[k][c] = (k < jb && c < jb) ? t[(j0 + k) + (long long)(j0 + c) * ld] : ((k == c) ? 1.0 : 0.0);
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r[JB];
#pragma unroll
    for (int c = 0; c < JB; ++c) r[c] = (c < jb) ? o[i + (long long)(j0 + c) * ld] : 0.0;
#pragma unroll
    for (int c = 0; c < JB; ++c) {
        double v = r[c];
#pragma unroll
        for (int k = 0; k < c; ++k) v -= r[k] * Ts[k][c];
        r[c] = v / Ts[c][c];
    }
#pragma unroll
    for (int c = 0; c < JB; ++c)
        if (c < jb) x[i + (long long)(j0 + c) * ld] = r[c];
}

cudaError_t launch_rdivp(const RdivpParams& p, cudaStream_t st)
{
    if (p.batch <= 0) return cudaSuccess;
    cudaError_t e = launch_permute_cols(p.A, p.work, p.pivot, p.n, p.ld, p.strideA, p.stridePivot, p.batch, st);
    if (e != cudaSuccess) return e;
    if (p.strideW != p.strideA) return cudaErrorInvalidValue;
    for (int j0 = 0; j0 < p.n; j0 += JB) {
        const int jb = (p.n - j0 < JB) ? (p.n - j0) : JB;
        if (j0 > 0) {
            GemmParams g{};
            g.M = p.n; g.N = jb; g.K = j0;
            g.A = p.A; g.lda = p.ld; g.strideA = p.strideA; g.transA = 0;
            g.B = p.T + (long long)j0 * p.ld; g.ldb = p.ld; g.strideB = p.strideT; g.transB = 0;
            g.C = p.work + (long long)j0 * p.ld; g.ldc = p.ld; g.strideC = p.strideW;
            g.alpha = -1.0; g.beta = 1.0;
            g.rs = no_scale(); g.ks = no_scale(); g.cs = no_scale();
            g.add_diag = nullptr; g.add_stride = 0; g.batch = p.batch;
            e = launch_gemm(g, st);
            if (e != cudaSuccess) return e;
        }
        dim3 grid((unsigned)((p.n + 127) / 128), (unsigned)p.batch);
        trsm_diag_kernel<<<grid, 128, 0, st>>>(p.A, p.work, p.T, p.n, p.ld, j0, jb, p.strideA, p.strideT,
                                               p.strideW);
        count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace dqmc
