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
// SM retires 64 FP64 FMA/clk, so the FP64 pipe is the roofline.
//
// Data path: operand tiles are fetched by the TMA engine (cp.async.bulk.tensor.3d, SASS UTMALDG) through
// tensor maps over (rows, columns, batch), into a multi-stage shared-memory ring whose slots are handed over
// with transaction mbarriers: one producer warp issues the loads (and evaluates the inner diagonal factor of
// the k-tile), the consumer warps run DMMA and release the slot -- no CTA barrier in the k loop, no
// per-thread address arithmetic.  Bank conflicts: the box the TMA engine copies is 4 doubles WIDER than the
// tile, which makes the leading dimension of the shared tile = 4 (mod 16) doubles -- the conflict-free padded
// layout -- without any swizzle (the extra columns are never read).  Out-of-bounds parts of a box are
// zero-filled by the engine, which is all the ragged-edge handling the loads need.
#include <cuda.h>

#include "common.cuh"

namespace dqmc {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- mbarrier / TMA primitives ------------------------------------------------------------------------------
__device__ __forceinline__ unsigned g_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void g_mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void g_mbar_arrive_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "GMW_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra GMD_%=;\n"
                 "bra GMW_%=;\n"
                 "GMD_%=:\n"
                 "}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

// Shared tiles, in "x, k" terms (x = m for A, n for B):
//  KMAJOR == false : global element (x, k) at g[x + k * ld]  (x contiguous) -> smem (x, k) at k * (XT + 4) + x
//  KMAJOR == true  : global element (x, k) at g[k + x * ld]  (k contiguous) -> smem (x, k) at x * (BK + 4) + k
// both leading dimensions are = 4 (mod 16) doubles for XT in {32, 48, 64} and BK = 16: conflict-free fragment loads
template <int XT, bool KMAJOR, int BK>
__device__ __forceinline__ double tile_at(const double* s, int x, int k)
{
    if constexpr (!KMAJOR) return s[k * (XT + 4) + x];
    else return s[x * (BK + 4) + k];
}
template <int XT, bool KMAJOR, int BK> __host__ __device__ constexpr int tile_elems() { return KMAJOR ? XT * (BK + 4) : BK * (XT + 4); }

template <int BM, int BN, int WM, int WN, bool TA, bool TB, int BK, int STAGES, int MINB>
__global__ void __launch_bounds__(((BM / WM) * (BN / WN) + 1) * 32, MINB)
gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GemmParams p)
{
    constexpr int NWM = BM / WM, NWN = BN / WN, NCW = NWM * NWN;      // consumer warps; warp NCW is the producer
    constexpr int MI = WM / 8, NJ = WN / 8;
    // A tile is K-major in global iff transA (A stored K x M, k contiguous)
    constexpr bool AK = TA;
    // B tile (n, k): global B is K x N col-major (k contiguous) unless transB
    constexpr bool BKM = !TB;
    constexpr int AE = tile_elems<BM, AK, BK>(), BE = tile_elems<BN, BKM, BK>();
    static_assert((AE * 8) % 128 == 0 && (BE * 8) % 128 == 0, "TMA destinations stay 128-byte aligned");

    // 128-byte aligned for the TMA destinations.  Declared with the alignment instead of rounding a generic pointer up by
    // hand: through the integer round trip the compiler loses the address space and every fragment load becomes a generic
    // LD.E instead of LDS (standalone 30.6 -> 31.0 TFLOP/s).
    extern __shared__ __align__(128) double smem[];
    double* As = smem;                                   // [STAGES][AE]
    double* Bs = smem + STAGES * AE;                     // [STAGES][BE]
    double* Ks = Bs + STAGES * BE;                       // [STAGES][BK] inner scale of the k-tile
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(Ks + STAGES * BK);   // full[STAGES], empty[STAGES]
    const unsigned full0 = g_smem_u32(bars), empty0 = g_smem_u32(bars + STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN, mat = blockIdx.z;
    const int KT = (p.K + BK - 1) / BK;
    const bool has_ks = p.ks.mode != 0;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { g_mbar_init(full0 + 8 * s, 1); g_mbar_init(empty0 + 8 * s, NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp == NCW) {
        // ================================ producer warp ================================
        const int za = p.strideA ? mat : 0, zb = p.strideB ? mat : 0;
        for (int kt = 0; kt < KT; ++kt) {
            const int st = kt % STAGES, k0 = kt * BK;
            if (kt >= STAGES) g_mbar_wait(empty0 + 8 * st, (unsigned)((kt / STAGES) - 1) & 1u);
            if (has_ks && lane < BK) {
                const int gk = k0 + lane;
                Ks[st * BK + lane] = (gk < p.K) ? scale_at(p.ks, mat, gk) : 0.0;
            }
            __syncwarp();
            if (lane == 0) {
                g_mbar_arrive_expect_tx(full0 + 8 * st, (unsigned)((AE + BE) * sizeof(double)));
                if constexpr (!AK) tma_load_3d(g_smem_u32(As + st * AE), &mapA, m0, k0, za, full0 + 8 * st);
                else tma_load_3d(g_smem_u32(As + st * AE), &mapA, k0, m0, za, full0 + 8 * st);
                if constexpr (!BKM) tma_load_3d(g_smem_u32(Bs + st * BE), &mapB, n0, k0, zb, full0 + 8 * st);
                else tma_load_3d(g_smem_u32(Bs + st * BE), &mapB, k0, n0, zb, full0 + 8 * st);
            }
        }
        return;
    }

    // ================================ consumer warps ================================
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp % NWM) * WM, wn0 = (warp / NWM) * WN;
    const bool has_rs = p.rs.mode != 0, has_cs = p.cs.mode != 0;
    double* C = p.C + (long long)mat * p.strideC;

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

    for (int kt = 0; kt < KT; ++kt) {
        const int st = kt % STAGES;
        g_mbar_wait(full0 + 8 * st, (unsigned)(kt / STAGES) & 1u);
        const double* as = As + st * AE;
        const double* bs = Bs + st * BE;
        const double* ks = Ks + st * BK;
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
        __syncwarp();
        if (lane == 0) g_mbar_arrive(empty0 + 8 * st);   // this warp is done with the slot
    }

    // epilogue: thread owns C[row = g, cols = 2t, 2t+1] of every 8x8 block.  The column factors are read before the
    // stores (the compiler cannot hoist them itself: the stores to C might alias the factor vectors).
    double csv[NJ][2];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int col = n0 + wn0 + j * 8 + 2 * t + e;
            csv[j][e] = (has_cs && col < p.N) ? scale_at(p.cs, mat, col) : 1.0;
        }
    // Interior tiles without a diagonal term or a read of C take a branch-free path: the general path below is one
    // dependent chain of bounds checks, 64-bit index arithmetic and branches PER ELEMENT (measured with clock64: ~340
    // cycles per element, 11 000 cycles = 20 % of the life of a 64 x 64 x 256 CTA), this one a handful of independent
    // multiplies and stores per element.
    const bool fast = (m0 + BM <= p.M) && (n0 + BN <= p.N) && !p.add_diag && (p.beta == 0.0 || preload);
    if (fast) {
        double rsv[MI];
#pragma unroll
        for (int i = 0; i < MI; ++i) rsv[i] = has_rs ? scale_at(p.rs, mat, m0 + wm0 + i * 8 + g) : 1.0;
        const double alpha = p.alpha;
        double* cb = C + (m0 + wm0 + g) + (long long)(n0 + wn0 + 2 * t) * p.ldc;
        const long long ldc = p.ldc;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double* cc = cb + (long long)(j * 8 + e) * ldc;
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    double v = alpha * acc[i][j][e] * rsv[i];        // same order of the factors as the general path
                    if (has_cs) v *= csv[j][e];
                    cc[i * 8] = v;
                }
            }
        return;
    }
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
                if (has_cs) v *= csv[j][e];
                if (p.add_diag && row == col) v += p.add_diag[(long long)mat * p.add_stride + row];
                double* dst = C + row + (long long)col * p.ldc;
                if (p.beta != 0.0 && !preload) v += p.beta * (*dst);
                *dst = v;
            }
        }
    }
}

// ---- host: tensor maps ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) f = nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}

// (dim0 = contiguous rows of the stored matrix, dim1 = its columns, dim2 = batch); box0 x box1 x 1 doubles per load
static cudaError_t make_map(CUtensorMap* map, const double* base, int rows, int cols, int ld, long long stride, int batch,
                            int box0, int box1)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return cudaErrorNotSupported;
    const bool shared = stride == 0;
    cuuint64_t dims[3] = {(cuuint64_t)rows, (cuuint64_t)cols, (cuuint64_t)(shared ? 1 : batch)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 8ull, (cuuint64_t)(shared ? (long long)ld * cols : stride) * 8ull};
    cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int BM, int BN, int WM, int WN, bool TA, bool TB, int BK = 16, int STAGES = 3, int MINB = 4>
static cudaError_t launch_cfg(const GemmParams& p, cudaStream_t st)
{
    constexpr int NT = ((BM / WM) * (BN / WN) + 1) * 32;
    constexpr int AE = tile_elems<BM, TA, BK>(), BE = tile_elems<BN, !TB, BK>();
    constexpr int smem = (STAGES * (AE + BE) + STAGES * BK) * (int)sizeof(double) + 2 * STAGES * 8 + 128;
    auto kern = gemm_kernel<BM, BN, WM, WN, TA, TB, BK, STAGES, MINB>;
    static SmemAttr attr;
    cudaError_t e = attr.ensure(kern, smem);
    if (e != cudaSuccess) return e;
    CUtensorMap mapA, mapB;
    // A: M x K (or K x M when transposed), B: K x N (or N x K when transposed), as stored
    if (!TA) e = make_map(&mapA, p.A, p.M, p.K, p.lda, p.strideA, p.batch, BM + 4, BK);
    else e = make_map(&mapA, p.A, p.K, p.M, p.lda, p.strideA, p.batch, BK + 4, BM);
    if (e != cudaSuccess) return e;
    if (TB) e = make_map(&mapB, p.B, p.N, p.K, p.ldb, p.strideB, p.batch, BN + 4, BK);
    else e = make_map(&mapB, p.B, p.K, p.N, p.ldb, p.strideB, p.batch, BK + 4, BN);
    if (e != cudaSuccess) return e;
    dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN, p.batch);
    kern<<<grid, NT, smem, st>>>(mapA, mapB, p);
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
    // Measured on 296 x 256^3: BK = 16 with 4 stages at 3 CTAs/SM 28.98 TFLOP/s, 3 stages 28.44, BK = 32 with 2 stages 28.99,
    // 2 CTAs/SM 25.2; the cp.async ring this replaces reached 28.4 with 5 CTAs/SM (DMMA pipe 81 % busy either way).
    // profiles/bench_gemm.cu (standalone, 30.5-31.2 on its box; cuBLAS batched = cutlass_80 d884gemm 64x128_16x3, 32 x 64 warp
    // tiles, 220 registers, 2 CTAs/SM: 33.0): persistent CTAs with a static tile walk 28.3 (lock-step CTAs, no dynamic
    // balancing); 64 x 128 / 128 x 64 tiles with 32 x 64 warp tiles, with or without the producer warp, 22-26 (two warps per
    // scheduler do not hide the fragment loads under ptxas' DMMA + NOP schedule); no epilogue stores 29.6 (= no change); an
    // 8 x longer k loop 33.4 -- the remaining loss is half pipeline fill per 16-k-tile output tile, half the main loop.
    // With LDS fragment loads and explicitly double-buffered fragments the 64 x 128 tile reaches 26.6, the 64 x 64 tile
    // without a producer warp 30.0 against 31.0 with it.  With the TMA loads removed (consumers run on stale tiles) 31.45 and,
    // with the 8 x longer k loop, 33.4: operand delivery is not the limit.  profiles/microbench_dmma_loop.cu: the bare
    // shared-memory-fed DMMA loop reaches 31.7 with ONE warp per scheduler and 36.9 with two.  Without any mbarrier traffic
    // either: 32.1 (34.0 with the 8 x longer k loop).  A 64 x 128 tile without producer warp whose slots are re-armed by
    // the LAST consumer warp to finish them (shared-memory counter, nobody waits): 26.5 as well; compiling the inner
    // diagonal factor out, constant instead of random operands: no change.  What separates the 64 x 128 kernel (two warps
    // per scheduler, 72 % of the pipe) from the microbenchmark (99 % with two) was found with clock64 stamps inside the kernel:
    // the EPILOGUE.  Its per-element chain of bounds checks, 64-bit index arithmetic and branches took ~340 cycles per element
    // -- 11 000 cycles (20 %) of the life of a 64 x 64 CTA, 21 000-28 000 (23-28 %) of a 64 x 128 one -- and predicating only
    // the store (the "no epilogue" experiment above) left all of it in place.  With the branch-free path for interior tiles:
    // 32.1 standalone (cuBLAS batched 32.95), 30.7 in the sweep; consumer-only upper bound at K = 256 (no loads, no
    // barriers) 32.7 for both the 64 x 64 and the 64 x 128 tile, so the large tile has nothing left to give.
    if (w64 <= w48 + 1e-9 && w64 <= w32 + 1e-9) {
        // transposed A: BOTH operand tiles are k-major, i.e. boxes of 64 rows of BK + 4 doubles; at BK = 16 the 160-byte rows
        // cost the TMA engine more requests per byte (0.360 against 0.325 ms per 296 x 256^3 launch for the plain variant); with
        // BK = 32 and two stages (same shared memory, 288-byte rows) the transposed variant runs at 0.325 ms too
        if constexpr (TA) return launch_cfg<64, 64, 32, 32, TA, TB, 32, 2, 3>(p, st);
        else return launch_cfg<64, 64, 32, 32, TA, TB, 16, 4, 3>(p, st);
    }
    if (w48 <= w32 + 1e-9) return launch_cfg<48, 48, 24, 24, TA, TB>(p, st);
    return launch_cfg<32, 32, 16, 16, TA, TB>(p, st);
}

cudaError_t launch_gemm(const GemmParams& p, cudaStream_t st)
{
    if (p.batch <= 0 || p.M <= 0 || p.N <= 0) return cudaSuccess;
    // TMA: 16-byte aligned base addresses and strides (even leading dimensions, even offsets)
    if ((p.lda & 1) || (p.ldb & 1) || (p.strideA & 1) || (p.strideB & 1) ||
        ((uintptr_t)p.A & 15) || ((uintptr_t)p.B & 15)) return cudaErrorInvalidValue;
    if (!p.transA && !p.transB) return launch_tiles<false, false>(p, st);
    if (!p.transA && p.transB) return launch_tiles<false, true>(p, st);
    if (p.transA && !p.transB) return launch_tiles<true, false>(p, st);
    return launch_tiles<true, true>(p, st);
}

}  // namespace dqmc
