// gemm.cu -- batched FP64 tensor-core GEMM with fused diagonal scalings.
//
// Replaces the reference's `vmul!` family (src/flavors/DQMC/linalg/real.jl:7-102)
// and the slice-matrix products built from it (src/flavors/DQMC/stack.jl:319-367):
//     C = beta*C + alpha * diag(rs) * op(A) * diag(ks) * op(B) * diag(cs) + diag(add)
// for a batch of independent column-major matrices; A or B may be shared by the
// whole batch (stride 0), which is how the hopping exponentials are applied.
//
// sm_100a: tcgen05.mma has no f64 kind, so FP64 tensor math is the warp-level
// DMMA (PTX mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).  One DMMA is 256 FMA and the
// SM retires 64 FP64 FMA/clk, so operand delivery is never the limit: tiles are
// staged with a 3-stage cp.async (LDGSTS) ring into padded shared memory laid out
// so that every fragment load is bank-conflict free (leading dimension = 4 mod 16
// doubles).  Roofline: FP64 pipe.
#include "common.cuh"

namespace dqmc {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }

// K-major tiles: (x, k) at x * (BK + 4) + k ; BK + 4 = 4 mod 16 for BK = 16 and 32

// Per-thread copy descriptors of one operand tile (XT x BK in "x,k" terms), computed ONCE per CTA:
// every k-tile then costs one pointer bump and NCH cp.async with precomputed offsets (the div/mod
// address arithmetic per chunk was ~20 % of the main loop).
//  KMAJOR == false : global element (x, k) at g[x + k * ld]  (x contiguous) -> smem (x,k) at k*(XT+4) + x
//  KMAJOR == true  : global element (x, k) at g[k + x * ld]  (k contiguous) -> smem (x,k) at x*(BK+4) + k
template <int XT, bool KMAJOR, int NTHREADS, int BK>
struct TileLoader {
    static constexpr int NCHUNK = XT * BK / 2;                       // 16-byte chunks per tile
    static constexpr int NCH = (NCHUNK + NTHREADS - 1) / NTHREADS;   // per thread
    int info[NCH];      // bits 0-4: bytes allowed by the x edge (0 / 8 / 16), bits 8-15: k, bits 16-23: x inside the tile
    int ld_;
    const double* g;    // points at (x0, k0) of the current k-tile
    long long kstep;    // elements to advance per k-tile

    __device__ __forceinline__ void init(const double* base, int ld, int x0, int X, int tid)
    {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = tid + i * NTHREADS;
            int x, k, xb;
            if constexpr (!KMAJOR) {
                constexpr int CH = XT / 2;
                k = c / CH; x = (c - k * CH) * 2;
                xb = (x0 + x + 1 < X) ? 16 : ((x0 + x < X) ? 8 : 0);
            } else {
                constexpr int CH = BK / 2;
                x = c / CH; k = (c - x * CH) * 2;
                xb = (x0 + x < X) ? 16 : 0;
            }
            if (c >= NCHUNK) { xb = 0; x = 0; k = 0; }
            info[i] = xb | (k << 8) | (x << 16);
        }
        ld_ = ld;
        if constexpr (!KMAJOR) { g = base + x0; kstep = (long long)BK * ld; }
        else { g = base + (long long)x0 * ld; kstep = BK; }
    }

    // issue the copies of the k-tile `g` currently points at; krem = K - k0 (> 0)
    __device__ __forceinline__ void issue(double* s, int krem, int tid)
    {
        if (krem >= BK) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                if (NCHUNK % NTHREADS != 0 && tid + i * NTHREADS >= NCHUNK) continue;
                const int xb = info[i] & 31, k = (info[i] >> 8) & 255, x = info[i] >> 16;
                const int so = KMAJOR ? (k + x * ld_) : (x + k * ld_);
                const int dof = KMAJOR ? (x * (BK + 4) + k) : (k * (XT + 4) + x);
                cp_async16(s + dof, xb ? (g + so) : g, xb);
            }
        } else {
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                if (NCHUNK % NTHREADS != 0 && tid + i * NTHREADS >= NCHUNK) continue;
                const int xb = info[i] & 31, k = (info[i] >> 8) & 255, x = info[i] >> 16;
                const int so = KMAJOR ? (k + x * ld_) : (x + k * ld_);
                const int dof = KMAJOR ? (x * (BK + 4) + k) : (k * (XT + 4) + x);
                int bytes;
                if constexpr (!KMAJOR) bytes = (k < krem) ? xb : 0;
                else bytes = xb ? ((k + 1 < krem) ? 16 : ((k < krem) ? 8 : 0)) : 0;
                cp_async16(s + dof, bytes ? (g + so) : g, bytes);
            }
        }
    }
    __device__ __forceinline__ void advance() { g += kstep; }
};

template <int XT, bool KMAJOR, int BK>
__device__ __forceinline__ double tile_at(const double* s, int x, int k)
{
    if constexpr (!KMAJOR) return s[k * (XT + 4) + x];
    else return s[x * (BK + 4) + k];
}

template <int XT, bool KMAJOR, int BK> __host__ __device__ constexpr int tile_elems() { return KMAJOR ? XT * (BK + 4) : BK * (XT + 4); }

template <int BM, int BN, int WM, int WN, bool TA, bool TB, int BK, int STAGES, int MINB>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, MINB)
gemm_kernel(const GemmParams p)
{
    constexpr int NWM = BM / WM, NWN = BN / WN, NT = NWM * NWN * 32;
    constexpr int MI = WM / 8, NJ = WN / 8;
    // A tile is K-major in global iff transA (A stored K x M, k contiguous)
    constexpr bool AK = TA;
    // B tile (n, k): global B is K x N col-major (k contiguous) unless transB
    constexpr bool BKM = !TB;
    constexpr int AE = tile_elems<BM, AK, BK>(), BE = tile_elems<BN, BKM, BK>();

    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * AE;
    double* Ks = Bs + STAGES * BE;          // [STAGES][BK] inner scale

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp % NWM) * WM, wn0 = (warp / NWM) * WN;
    const bool has_ks = p.ks.mode != 0;
    const bool has_rs = p.rs.mode != 0, has_cs = p.cs.mode != 0;

    // one CTA per output tile (a persistent variant with a cross-tile cp.async ring was measured
    // slower: 25.8 vs 26.8 TFLOP/s at K = 256 -- the block scheduler already staggers the 8 waves)
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN, mat = blockIdx.z;
    const double* A = p.A + (long long)mat * p.strideA;
    const double* B = p.B + (long long)mat * p.strideB;
    double* C = p.C + (long long)mat * p.strideC;
    const int KT = (p.K + BK - 1) / BK;

    // C read-modify-write without scalings (the panel updates of rdivp.cu):
    // the accumulators start from beta / alpha * C, so the loads of C overlap the main loop instead of sitting
    // behind it in the epilogue.  Exact for alpha = +-1.
    const bool preload = (p.beta != 0.0) && !has_rs && !has_cs && !p.add_diag && (p.alpha == 1.0 || p.alpha == -1.0);
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    if (preload) {
        const double f = p.beta * p.alpha;               // beta / alpha for alpha = +-1
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int row = m0 + wm0 + i * 8 + g;
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = n0 + wn0 + j * 8 + 2 * t + e;
                    if (row < p.M && col < p.N) acc[i][j][e] = f * C[row + (long long)col * p.ldc];
                }
        }
    }

    TileLoader<BM, AK, NT, BK> la;
    TileLoader<BN, BKM, NT, BK> lb;
    la.init(A, p.lda, m0, p.M, tid);
    lb.init(B, p.ldb, n0, p.N, tid);
    int kr_i = 0;                            // k-tile the loaders point at
    auto issue = [&](int stage) {
        const int k0 = kr_i * BK;
        la.issue(As + stage * AE, p.K - k0, tid);
        lb.issue(Bs + stage * BE, p.K - k0, tid);
        la.advance(); lb.advance();
        if (has_ks && tid < BK) {
            const int gk = k0 + tid;
            Ks[stage * BK + tid] = (gk < p.K) ? scale_at(p.ks, mat, gk) : 0.0;
        }
        ++kr_i;
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) issue(s);
        cp_async_commit();
    }

    int stage = 0, stage_n = STAGES - 1;     // stage being computed / stage the next issue goes to
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (kt + STAGES - 1 < KT) issue(stage_n);
        cp_async_commit();
        if (++stage_n == STAGES) stage_n = 0;

        const double* as = As + stage * AE;
        const double* bs = Bs + stage * BE;
        const double* ks = Ks + stage * BK;
        if (++stage == STAGES) stage = 0;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            const int k = kk * 4 + t;
            double af[MI], bf[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) af[i] = tile_at<BM, AK, BK>(as, wm0 + i * 8 + g, k);
#pragma unroll
            for (int j = 0; j < NJ; ++j) bf[j] = tile_at<BN, BKM, BK>(bs, wn0 + j * 8 + g, k);
            if (has_ks) {
                const double sk = ks[k];
#pragma unroll
                for (int j = 0; j < NJ; ++j) bf[j] *= sk;
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread owns C[row = g, cols = 2t, 2t+1] of every 8x8 block
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int row = m0 + wm0 + i * 8 + g;
        if (row >= p.M) continue;
        const double r = has_rs ? scale_at(p.rs, mat, row) : 1.0;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = n0 + wn0 + j * 8 + 2 * t + e;
                if (col >= p.N) continue;
                double v = p.alpha * acc[i][j][e] * r;
                if (has_cs) v *= scale_at(p.cs, mat, col);
                if (p.add_diag && row == col) v += p.add_diag[(long long)mat * p.add_stride + row];
                double* dst = C + row + (long long)col * p.ldc;
                if (p.beta != 0.0 && !preload) v += p.beta * (*dst);
                *dst = v;
            }
        }
    }
}

template <int BM, int BN, int WM, int WN, bool TA, bool TB, int BK = 16, int STAGES = 3, int MINB = 4>
static cudaError_t launch_cfg(const GemmParams& p, cudaStream_t st)
{
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    constexpr int AE = tile_elems<BM, TA, BK>(), BE = tile_elems<BN, !TB, BK>();
    constexpr int smem = (STAGES * (AE + BE) + STAGES * BK) * (int)sizeof(double);
    auto kern = gemm_kernel<BM, BN, WM, WN, TA, TB, BK, STAGES, MINB>;
    static SmemAttr attr;
    cudaError_t e = attr.ensure(kern, smem);
    if (e != cudaSuccess) return e;
    dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN, p.batch);
    kern<<<grid, NT, smem, st>>>(p);
    count_launch();
    return cudaGetLastError();
}

template <bool TA, bool TB>
static cudaError_t launch_tiles(const GemmParams& p, cudaStream_t st)
{
    // pick the tile that wastes the least padded work; ties go to the larger tile
    auto waste = [&](int bm, int bn) {
        const double mm = (double)((p.M + bm - 1) / bm * bm), nn = (double)((p.N + bn - 1) / bn * bn);
        return mm * nn / ((double)p.M * (double)p.N);
    };
    const double w64 = waste(64, 64), w48 = waste(48, 48), w32 = waste(32, 32);
    // Measured on 296 x 256^3 (profiles/r1_summary.md): 64 x 64 tiles, 2 stages, 5 CTAs/SM is the best of 22 variants
    // (128 x 64 / 128 x 128 tiles, BK = 8 / 32, 3-4 stages were all slower: more co-resident CTAs at independent phases
    // beat larger tiles); n = 288: a 96 x 96 CTA tile covers it exactly like 48 x 48 does but measured slower.
    if (w64 <= w48 + 1e-9 && w64 <= w32 + 1e-9)
        return launch_cfg<64, 64, 32, 32, TA, TB, 16, 2, 5>(p, st);
    if (w48 <= w32 + 1e-9) return launch_cfg<48, 48, 24, 24, TA, TB>(p, st);
    return launch_cfg<32, 32, 16, 16, TA, TB>(p, st);
}

cudaError_t launch_gemm(const GemmParams& p, cudaStream_t st)
{
    if (p.batch <= 0 || p.M <= 0 || p.N <= 0) return cudaSuccess;
    if ((p.lda & 1) || (p.ldb & 1)) return cudaErrorInvalidValue;   // 16-byte cp.async columns
    if (!p.transA && !p.transB) return launch_tiles<false, false>(p, st);
    if (!p.transA && p.transB) return launch_tiles<false, true>(p, st);
    if (p.transA && !p.transB) return launch_tiles<true, false>(p, st);
    return launch_tiles<true, true>(p, st);
}

}  // namespace dqmc
