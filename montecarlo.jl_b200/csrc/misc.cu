// misc.cu -- small elementwise / reduction kernels around the hot kernels.
#include "common.cuh"
#include <math.h>

namespace dqmc {

__global__ void set_identity_kernel(double* A, int n, int ld, long long stride)
{
    double* a = A + (long long)blockIdx.y * stride;
    const long long tot = (long long)ld * n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot;
         e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % ld), j = (int)(e / ld);
        a[e] = (i == j && i < n) ? 1.0 : 0.0;
    }
}

cudaError_t launch_set_identity(double* A, int n, int ld, long long stride, int batch, cudaStream_t st)
{
    if (batch <= 0) return cudaSuccess;
    const long long tot = (long long)ld * n;
    dim3 grid((unsigned)((tot + 255) / 256 > 64 ? 64 : (tot + 255) / 256), (unsigned)batch);
    set_identity_kernel<<<grid, 256, 0, st>>>(A, n, ld, stride);
    count_launch();
    return cudaGetLastError();
}

__global__ void fill_kernel(double* v, double val, long long count)
{
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < count;
         e += (long long)gridDim.x * blockDim.x) v[e] = val;
}

cudaError_t launch_fill(double* v, double val, long long count, cudaStream_t st)
{
    if (count <= 0) return cudaSuccess;
    long long b = (count + 255) / 256; if (b > 1184) b = 1184;
    fill_kernel<<<(unsigned)b, 256, 0, st>>>(v, val, count);
    count_launch();
    return cudaGetLastError();
}

// O[:, j] = A[:, pivot[j]]   (first half of rdivp!, reference real.jl:204-209)
__global__ void permute_cols_kernel(const double* A, double* O, const int* pivot, int n, int ld,
                                    long long stride, long long pstride)
{
    const int mat = blockIdx.y;
    const double* a = A + (long long)mat * stride;
    double* o = O + (long long)mat * stride;
    const int* pv = pivot + (long long)mat * pstride;
    for (int j = blockIdx.x; j < n; j += gridDim.x) {
        const int pj = pv[j];
        for (int i = threadIdx.x; i < n; i += blockDim.x) o[i + (long long)j * ld] = a[i + (long long)pj * ld];
    }
}

cudaError_t launch_permute_cols(const double* A, double* O, const int* pivot, int n, int ld,
                                long long stride, long long pstride, int batch, cudaStream_t st)
{
    if (batch <= 0) return cudaSuccess;
    dim3 grid((unsigned)(n < 32 ? n : 32), (unsigned)batch);
    permute_cols_kernel<<<grid, 128, 0, st>>>(A, O, pivot, n, ld, stride, pstride);
    count_launch();
    return cudaGetLastError();
}

// Propagation-error check (reference stack.jl:644-654, 699-709): d = max|A - B| over the
// chain's matrices; if d > thresh push into MagnitudeStats [count, sum log10, min, max].
__global__ void prop_error_kernel(const double* A, const double* B, int n, int ld, long long stride_chain,
                                  int nb, double thresh, double* stats, int s)
{
    const int chain = blockIdx.x;
    const double* a = A + (long long)chain * stride_chain;
    const double* b = B + (long long)chain * stride_chain;
    // one CTA per chain streams 2 x nb x n x n doubles: four independent column loads per thread keep enough bytes in
    // flight for one SM's share of HBM bandwidth
    double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
    const int cols = nb * n;
    // s * ld threads own (row, column group); the block is rounded up to whole warps and the extra threads only take
    // part in the reduction (with m = 0)
    const bool active = (int)threadIdx.x < s * ld;
    for (int i = active ? (int)threadIdx.x % ld : n; i < n; i += ld)   // at most one pass: threads own a row, walk columns
        for (int j = threadIdx.x / ld; j < cols; j += 4 * s) {
            const long long e0 = (long long)j * ld + i;
            const bool p1 = j + s < cols, p2 = j + 2 * s < cols, p3 = j + 3 * s < cols;
            const double a0 = a[e0], b0 = b[e0];
            const double a1 = p1 ? a[e0 + (long long)s * ld] : 0.0, b1 = p1 ? b[e0 + (long long)s * ld] : 0.0;
            const double a2 = p2 ? a[e0 + 2LL * s * ld] : 0.0, b2 = p2 ? b[e0 + 2LL * s * ld] : 0.0;
            const double a3 = p3 ? a[e0 + 3LL * s * ld] : 0.0, b3 = p3 ? b[e0 + 3LL * s * ld] : 0.0;
            m0 = fmax(m0, fabs(a0 - b0)); m1 = fmax(m1, fabs(a1 - b1));
            m2 = fmax(m2, fabs(a2 - b2)); m3 = fmax(m3, fabs(a3 - b3));
        }
    double m = fmax(fmax(m0, m1), fmax(m2, m3));
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
        if (m > thresh) {
            double* s = stats + (long long)chain * 4;
            s[0] += 1.0; s[1] += log10(m);
            s[2] = fmin(s[2], m); s[3] = fmax(s[3], m);
        }
    }
}

cudaError_t launch_prop_error(const double* A, const double* B, int n, int ld, long long stride_chain,
                              int nb, int n_chains, double thresh, double* stats, cudaStream_t st)
{
    if (n_chains <= 0) return cudaSuccess;
    // threads = a multiple of ld so that every thread owns one row (rows >= n are padding and idle)
    const int s = 1024 / ld;
    if (s == 0) {   // ld > 1024 does not occur (n <= 512); keep the kernel's ownership rule valid anyway
        return cudaErrorInvalidValue;
    }
    const int threads = ((s * ld + 31) / 32) * 32;       // whole warps: the shuffle reduction needs full masks
    prop_error_kernel<<<(unsigned)n_chains, threads, 0, st>>>(A, B, n, ld, stride_chain, nb, thresh, stats, s);
    count_launch();
    return cudaGetLastError();
}

// per-element accumulators for the observable reduction: sum += x, sumsq += x^2
__global__ void accumulate_kernel(const double* G, double* sum, double* sumsq, long long count)
{
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < count;
         e += (long long)gridDim.x * blockDim.x) {
        const double x = G[e];
        sum[e] += x; sumsq[e] += x * x;
    }
}

cudaError_t launch_accumulate(const double* G, double* sum, double* sumsq, long long count, cudaStream_t st)
{
    if (count <= 0) return cudaSuccess;
    long long b = (count + 255) / 256; if (b > 1184) b = 1184;
    accumulate_kernel<<<(unsigned)b, 256, 0, st>>>(G, sum, sumsq, count);
    count_launch();
    return cudaGetLastError();
}

// O = diag(rs) * A * diag(cs) + add + add_diag * I.  A == nullptr stands for the identity, which gives
// copyto!(O, Diagonal(d)) (unequal_time_stack.jl:597, 679); add_diag = -1 with unit scales is
// vsub!(O, A, I) (linalg/real.jl:122-132); add != nullptr is rvadd! (:117-121).
__global__ void scale_add_kernel(double* O, const double* A, Scale rs, Scale cs, const double* add, double add_diag,
                                 int n, int ld, long long stride)
{
    const int mat = blockIdx.y;
    const long long off = (long long)mat * stride;
    const long long tot = (long long)ld * n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < tot;
         e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % ld), j = (int)(e / ld);
        if (i >= n) continue;
        double v = A ? A[off + e] : ((i == j) ? 1.0 : 0.0);
        if (v != 0.0 || A) {
            if (rs.mode) v *= scale_at(rs, mat, i);
            if (cs.mode) v *= scale_at(cs, mat, j);
        }
        if (add) v += add[off + e];
        if (i == j) v += add_diag;
        O[off + e] = v;
    }
}

cudaError_t launch_scale_add(double* O, const double* A, Scale rs, Scale cs, const double* add, double add_diag,
                             int n, int ld, long long stride, int batch, cudaStream_t st)
{
    if (batch <= 0) return cudaSuccess;
    const long long tot = (long long)ld * n;
    dim3 grid((unsigned)((tot + 255) / 256 > 64 ? 64 : (tot + 255) / 256), (unsigned)batch);
    scale_add_kernel<<<grid, 256, 0, st>>>(O, A, rs, cs, add, add_diag, n, ld, stride);
    count_launch();
    return cudaGetLastError();
}

// compress(field) = BitArray(conf .== 1) (fields.jl:331): Julia's BitArray keeps bit i of the column-major array
// in chunks[i >> 6] at bit position i & 63.  One thread per 64-bit chunk; pack != 0: conf -> chunks, else the
// inverse decompress! (fields.jl:334, conf = 2 bit - 1).
__global__ void conf_bits_kernel(int8_t* conf, unsigned long long* chunks, long long nvalues, long long words_per_chain,
                                 int n_chains, int pack, int ghq)
{
    const long long tot = words_per_chain * n_chains;
    const int vpw = ghq ? 32 : 64;                       // values per 64-bit word
    for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < tot; w += (long long)gridDim.x * blockDim.x) {
        const long long chain = w / words_per_chain, wi = w - chain * words_per_chain;
        int8_t* c = conf + chain * nvalues + wi * vpw;
        const int nv = (int)((nvalues - wi * vpw < vpw) ? (nvalues - wi * vpw) : vpw);
        if (pack) {
            unsigned long long v = 0ull;
            if (!ghq) for (int b = 0; b < nv; ++b) v |= (unsigned long long)(c[b] == 1) << b;
            else                                         // fields.jl:476-480: (1, 2, 3, 4) -> (00, 01, 10, 11), high bit first
                for (int b = 0; b < nv; ++b) {
                    const unsigned long long x = (unsigned long long)((c[b] - 1) & 3);
                    v |= ((x >> 1) << (2 * b)) | ((x & 1ull) << (2 * b + 1));
                }
            chunks[w] = v;
        } else {
            const unsigned long long v = chunks[w];
            if (!ghq) for (int b = 0; b < nv; ++b) c[b] = (int8_t)(2 * (int)((v >> b) & 1ull) - 1);
            else                                         // fields.jl:481-489: 1 + 2 bit1 + bit2
                for (int b = 0; b < nv; ++b)
                    c[b] = (int8_t)(1 + 2 * (int)((v >> (2 * b)) & 1ull) + (int)((v >> (2 * b + 1)) & 1ull));
        }
    }
}

cudaError_t launch_conf_bits(int8_t* conf, unsigned long long* chunks, long long nvalues, int n_chains, int pack, int ghq,
                             cudaStream_t st)
{
    if (n_chains <= 0) return cudaSuccess;
    const long long nbits = nvalues * (ghq ? 2 : 1);
    const long long wpc = (nbits + 63) / 64, tot = wpc * n_chains;
    long long b = (tot + 255) / 256; if (b > 1184) b = 1184;
    conf_bits_kernel<<<(unsigned)b, 256, 0, st>>>(conf, chunks, nvalues, wpc, n_chains, pack, ghq);
    count_launch();
    return cudaGetLastError();
}

}  // namespace dqmc
