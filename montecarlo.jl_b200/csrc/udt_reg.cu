// udt_reg.cu -- register-resident batched column-pivoted Householder QR -> UDT (v2 of udt.cu).
//
// Same mathematics and outputs as udt.cu (reference src/flavors/DQMC/linalg/UDT.jl:216-334),
// different residence: the n x n/CS column panel of a CTA lives in REGISTERS.
//   * cluster of CS CTAs per matrix, columns dealt cyclically (column c -> CTA c % CS),
//   * inside a CTA warp w owns 8 consecutive local columns, lane l owns rows l, l+32, ...
//     => thread holds a[8][RPL] doubles, RPL = ceil(n/32) (n = 256: 64 doubles = 128 registers),
//   * per Householder step every lane needs only its RPL entries of v (one conflict-free
//     shared-memory read each), the 8 column dot products are reduced with a
//     recursive-halving shuffle tree (17 64-bit shuffles instead of 8 x 5 butterflies),
//   * the squared norms of the remaining columns are recomputed in the same pass
//     (UDT.jl:175-192 recomputes them every step as well),
//   * ONE cluster barrier per step: each CTA publishes its best remaining column (norm, index,
//     column tail) into every peer's shared memory through DSMEM; after the barrier every CTA
//     builds the identical reflector redundantly.  Columns are never swapped (free un-pivoting).
//   * forming Q (UDT.jl:272-288) runs backwards with NO block-level synchronisation at all: each
//     warp streams the Householder vectors from L2 into registers (prefetched one step ahead).
// Bound: latency of the per-step critical path (barrier + shuffle trees), then FP64 FMA issue.
#include <cooperative_groups.h>
#include <stdlib.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace dqmc {

struct UdtRegGeom { int cs, nwarps, nloc, nv, rpl, skip_q; size_t smem; long long* dbg; };

static bool udt_reg_geometry(int n, UdtRegGeom& g)
{
    const int rpl = (n + 31) / 32;
    if (rpl > 9) return false;
    const int maxw = (rpl <= 4) ? 16 : ((rpl <= 6) ? 9 : 8);   // matches the __launch_bounds__ below
    static const int min_cs = getenv("DQMC_UDT_CS") ? atoi(getenv("DQMC_UDT_CS")) : 1;   // experiment knob
    for (int cs = min_cs; cs <= 8; cs *= 2) {
        const int nloc = (n + cs - 1) / cs;
        const int w = (nloc + 7) / 8;
        if (w <= maxw) {
            g.cs = cs; g.nwarps = w; g.nloc = nloc; g.rpl = rpl; g.nv = rpl * 32;
            g.dbg = nullptr;
            g.skip_q = getenv("DQMC_UDT_SKIPQ") ? 1 : 0;   // timing experiments only (results are wrong)
            g.smem = ((size_t)2 * cs * g.nv + 2 * n + 16 + 2 * 32) * sizeof(double) +
                     ((size_t)w * 8 + n + 16 + 2 * 32 + 8) * sizeof(int);
            return true;
        }
    }
    return false;
}

// sum over the 32 lanes of 8 values per lane; afterwards EVERY lane holds all 8 totals.
__device__ __forceinline__ void warp_allreduce8(double (&x)[8], int lane)
{
    // recursive halving: after the three exchange rounds lane holds the partial of column
    // ((lane>>4)&1)*4 + ((lane>>3)&1)*2 + ((lane>>2)&1) summed over 8 lanes' worth of data
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
    double y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double send = h16 ? x[i] : x[i + 4];
        const double keep = h16 ? x[i + 4] : x[i];
        y[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    double z[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double send = h8 ? y[i] : y[i + 2];
        const double keep = h8 ? y[i + 2] : y[i];
        z[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    double w = (h4 ? z[1] : z[0]) + __shfl_xor_sync(0xffffffffu, h4 ? z[0] : z[1], 4);
    w += __shfl_xor_sync(0xffffffffu, w, 2);
    w += __shfl_xor_sync(0xffffffffu, w, 1);
    // lane now holds the total of column cidx = 4*h16 + 2*h8 + h4 ; gather all eight
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int src = ((c & 4) ? 16 : 0) | ((c & 2) ? 8 : 0) | ((c & 1) ? 4 : 0);
        x[c] = __shfl_sync(0xffffffffu, w, src);
    }
}

// same tree without the final gather: lane ends with the total of column col_of_lane(lane)
__device__ __forceinline__ double warp_reduce8(const double (&x)[8], int lane)
{
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
    double y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double send = h16 ? x[i] : x[i + 4];
        const double keep = h16 ? x[i + 4] : x[i];
        y[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    double z[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double send = h8 ? y[i] : y[i + 2];
        const double keep = h8 ? y[i + 2] : y[i];
        z[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    double w = (h4 ? z[1] : z[0]) + __shfl_xor_sync(0xffffffffu, h4 ? z[0] : z[1], 4);
    w += __shfl_xor_sync(0xffffffffu, w, 2);
    w += __shfl_xor_sync(0xffffffffu, w, 1);
    return w;
}
__device__ __forceinline__ int col_of_lane(int lane) { return ((lane & 16) ? 4 : 0) | ((lane & 8) ? 2 : 0) | ((lane & 4) ? 1 : 0); }

// cluster barrier with release/acquire at cluster scope (cg::cluster_group::sync() adds a
// GPU-scope MEMBAR in front of the same barrier; DSMEM + cluster-scope ordering is all we need)
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_barrier() { cluster_arrive(); cluster_wait(); }

// dot products of v with all 8 columns over register rows R0..RPL-1 (v is 0 on rows < j).  The 8
// accumulation chains are interleaved explicitly (r outer, c inner): FP64 FMA has ~10 cycles of
// dependent latency and ptxas keeps source order under this register pressure.  Inactive columns are
// computed too and masked afterwards (their dot is forced to 0 so the update leaves them untouched).
template <int RPL>
__device__ __forceinline__ void col_dots(const double (&a)[8][RPL], const double (&v)[RPL], unsigned act, double (&part)[8], int r0)
{
    double d[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) d[c] = 0.0;
#pragma unroll
    for (int r = 0; r < RPL; ++r)
        if (r >= r0) {                                   // warp-uniform branch around 8 independent FMAs
#pragma unroll
            for (int c = 0; c < 8; ++c) d[c] = fma(v[r], a[c][r], d[c]);
        }
#pragma unroll
    for (int c = 0; c < 8; ++c) part[c] = ((act >> c) & 1u) ? d[c] : 0.0;
}

// a[:, c] -= v * (tau * dot_c); part[c] <- sum of squares of the rows > j (only register row r0
// can contain rows <= j, it is masked with an integer AND instead of FP64 selects)
template <int RPL, bool NORMS>
__device__ __forceinline__ void col_update(double (&a)[8][RPL], const double (&v)[RPL], unsigned act, double (&part)[8],
                                           double tau, unsigned long long m0, int r0)
{
    double sd[8], nr[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { sd[c] = part[c] * tau; nr[c] = 0.0; }
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        if (r > r0) {
            const double nv = -v[r];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const double x = fma(nv, sd[c], a[c][r]);
                a[c][r] = x;
                if (NORMS) nr[c] = fma(x, x, nr[c]);
            }
        } else if (r == r0) {
            const double nv = -v[r];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const double x = fma(nv, sd[c], a[c][r]);
                a[c][r] = x;
                if (NORMS) nr[c] = fma(__longlong_as_double(__double_as_longlong(x) & (long long)m0), x, nr[c]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) part[c] = nr[c];
}

// 1/sqrt(x) and sqrt(x) to ~1 ulp: hardware seed (MUFU.RSQ64H) + two coupled Newton steps.  The
// library sqrt()/division are ~15-deep dependent FP64 chains each; they sit on the critical path
// of every Householder step.  x must be a positive normal number.
__device__ __forceinline__ void fast_rsqrt_sqrt(double x, double& rs, double& sq)
{
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double g = x * r, h = 0.5 * r;
    double e = fma(-g, h, 0.5);
    g = fma(g, e, g); h = fma(h, e, h);
    e = fma(-g, h, 0.5);
    g = fma(g, e, g); h = fma(h, e, h);
    sq = g; rs = h + h;
}
__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

#define DQMC_TICK(slot) do { if (dbg) { const long long t__ = clock64(); if (tid == 0) dbg[slot] += t__ - tprev; tprev = t__; } } while (0)

template <int RPL>
__global__ void __launch_bounds__((RPL <= 4) ? 512 : ((RPL <= 6) ? 288 : 256))
udt_reg_kernel(const UdtParams p, const UdtRegGeom gm)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = gm.cs, nv = gm.nv, n = p.n;
    const int csh = 31 - __clz(CS);                     // CS is 1, 2, 4 or 8
    const int rank = (int)cluster.block_rank();
    const int mat = blockIdx.x / CS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = gm.nwarps;
    const int nloc = (n - rank + CS - 1) / CS;           // local columns: slot s <-> column s * CS + rank

    extern __shared__ __align__(16) double sm[];
    double* vbuf = sm;                                   // [2][CS][nv]
    double* dvec = vbuf + (size_t)2 * CS * nv;           // [n]
    double* taus = dvec + n;                             // [n]
    double* candval = taus + n;                          // [2][8]
    double* wbval = candval + 16;                        // [2][32] per-warp candidates (double buffered)
    int* colstep = (int*)(wbval + 64);                   // [nwarps * 8]
    int* perm = colstep + nwarps * 8;                    // [n]
    int* candcol = perm + n;                             // [2][8]
    int* wbcol = candcol + 16;                           // [2][32]

    const double* Ag = p.A + (long long)mat * p.strideA;
    double* Vg = p.Vwork + (long long)mat * p.strideV;
    const int ld = p.ld, ldv = p.ldv;
    long long* dbg = (blockIdx.x == 0) ? gm.dbg : nullptr;   // per-phase cycle counters of CTA 0 (debug)
    long long tprev = dbg ? clock64() : 0;

    // ---- load the panel into registers ------------------------------------------------------
    double a[8][RPL];
    unsigned act = 0;                                    // warp-uniform: bit c set <=> column still active
    double part[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int s = warp * 8 + c;
        const bool have = s < nloc;
        const int col = s * CS + rank;
        const double sc = (have && p.colscale.mode) ? scale_at(p.colscale, mat, col) : 1.0;
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int row = lane + 32 * r;
            const double v = (have && row < n) ? Ag[row + (long long)col * ld] * sc : 0.0;
            a[c][r] = v;
            acc += v * v;
        }
        part[c] = acc;
        if (have) act |= 1u << c;
    }
    double mynorm = warp_reduce8(part, lane);            // norm of column col_of_lane(lane)
    if (lane < 8) colstep[warp * 8 + lane] = -1;

    // candidate of this warp -> shared (slot parity q)
    auto warp_candidate = [&](int q) {
        const int c = col_of_lane(lane);
        double bv = ((act >> c) & 1u) ? mynorm : -1.0;
        int bc = (warp * 8 + c) * CS + rank;
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
            if (ov > bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
        }
        if (lane == 0) { wbval[q * 32 + warp] = bv; wbcol[q * 32 + warp] = bc; }
    };

    // CTA winner -> publish (norm, column index, column tail from row j) into every peer
    auto publish = [&](int j) {
        const int q = j & 1;
        double bv = wbval[q * 32]; int bc = wbcol[q * 32];
        for (int w = 1; w < nwarps; ++w) {
            const double ov = wbval[q * 32 + w]; const int oc = wbcol[q * 32 + w];
            if (ov > bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
        }
        const int s = (bv >= 0.0) ? ((bc - rank) >> csh) : -1;
        if (s >= 0 && (s >> 3) == warp) {                // the warp that owns the winning column
            const int cc = s & 7;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c == cc) {                           // warp-uniform: stores straight from registers
                    for (int rk = 0; rk < CS; ++rk) {
                        double* rv = cluster.map_shared_rank(vbuf, rk) + ((size_t)q * CS + rank) * nv;
#pragma unroll
                        for (int r = 0; r < RPL; ++r) {
                            const int row = lane + 32 * r;
                            if (row >= j) rv[row] = a[c][r];
                        }
                    }
                }
        }
        if (tid == 0) {
            for (int rk = 0; rk < CS; ++rk) {
                cluster.map_shared_rank(candval, rk)[q * 8 + rank] = bv;
                cluster.map_shared_rank(candcol, rk)[q * 8 + rank] = bc;
            }
        }
    };

    cluster_barrier();                                   // peers resident before any DSMEM store

    DQMC_TICK(0);
    double vprev[RPL];                                   // Householder vector of the previous step (stored late)
    bool store_prev = false;
#pragma unroll
    for (int r = 0; r < RPL; ++r) vprev[r] = 0.0;
    for (int j = 0; j < n; ++j) {
        const int q = j & 1;
        // ---- pick and publish this CTA's best remaining column, then ONE cluster barrier ------
        warp_candidate(q);
        DQMC_TICK(7);
        __syncthreads();
        DQMC_TICK(8);
        publish(j);
        DQMC_TICK(9);
        cluster_arrive();
        // the previous Householder vector (0 .. 0 1 v) goes to global memory (for Q) between arrive and
        // wait, so that the release fence of the barrier never has to wait for these stores
        if (store_prev) {
#pragma unroll
            for (int r = 0; r < RPL; ++r) Vg[lane + 32 * r + (long long)(j - 1) * ldv] = vprev[r];
        }
        cluster_wait();
        DQMC_TICK(10);
        // ---- global winner, identical in every CTA ------------------------------------------
        double bv = candval[q * 8]; int bc = candcol[q * 8], br = 0;
        for (int r = 1; r < CS; ++r) {
            const double ov = candval[q * 8 + r]; const int oc = candcol[q * 8 + r];
            if (ov > bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; br = r; }
        }
        const double* raw = vbuf + ((size_t)q * CS + br) * nv;
        // ---- reflector (UDT.jl:157-172) ------------------------------------------------------
        double xi1 = raw[j], tau, rjj, inv;
        if (bv < 1e-290 || bv > 1e290) {                 // exact zero / out of the fast path's range: library math
            if (bv == 0.0) { tau = 0.0; rjj = xi1; inv = 0.0; }
            else {
                const double nu = copysign(sqrt(bv), xi1);
                xi1 += nu;
                rjj = -nu; tau = xi1 / nu; inv = 1.0 / xi1;
            }
        } else {
            double rs, sq;
            fast_rsqrt_sqrt(bv, rs, sq);
            const double nu = copysign(sq, xi1);
            xi1 += nu;                                   // |xi1| >= sqrt(bv): never cancels
            rjj = -nu; tau = xi1 * copysign(rs, nu); inv = fast_rcp(xi1);
        }
        DQMC_TICK(1);                                    // winner + reflector scalars
        const int r0 = j >> 5;                           // first register row that can be >= j
        double v[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int row = lane + 32 * r;
            double x = (row > j) ? raw[row] * inv : ((row == j) ? 1.0 : 0.0);
            v[r] = (row < n) ? x : 0.0;
        }
        const unsigned long long m0 = (lane + 32 * r0 > j) ? ~0ull : 0ull;   // rows of register row r0 that are > j
        const bool store_v = (rank == br && warp == 0);  // done after the barrier arrive, see below
        if (tid == 0) {
            const double ad = fabs(rjj);
            dvec[j] = (ad == 0.0) ? 1.0 : ad;
            taus[j] = tau;
            perm[j] = bc;
        }
        if (rank == br) {
            const int s = (bc - rank) >> csh;
            if ((s >> 3) == warp) {                      // retire the pivot column, store R_jj
                const int cc = s & 7;
                act &= ~(1u << cc);
                if (lane == 0) colstep[s] = j;
#pragma unroll
                for (int r = 0; r < RPL; ++r)
                    if (r == r0) {                       // warp-uniform
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (c == cc && lane == (j & 31)) a[c][r] = rjj;
                    }
            }
        }

        DQMC_TICK(2);                                    // v, bookkeeping, retire
        // ---- apply H_j to the active columns of this warp, fused norm recompute -------------
        if (act != 0u) {                                 // warp-uniform
            col_dots<RPL>(a, v, act, part, r0);
            DQMC_TICK(3);
            warp_allreduce8(part, lane);
            DQMC_TICK(4);
            col_update<RPL, true>(a, v, act, part, tau, m0, r0);
            DQMC_TICK(5);
            mynorm = warp_reduce8(part, lane);
            DQMC_TICK(6);
        }
        store_prev = store_v;
#pragma unroll
        for (int r = 0; r < RPL; ++r) vprev[r] = v[r];
    }
    if (store_prev) {
#pragma unroll
        for (int r = 0; r < RPL; ++r) Vg[lane + 32 * r + (long long)(n - 1) * ldv] = vprev[r];
    }
    __threadfence();
    cluster_barrier();       // V (global) of every step owner is visible cluster-wide; dvec/perm final

    // ---- D, pivot, T ----------------------------------------------------------------------------
    if (rank == 0) {
        double* Dg = p.D + (long long)mat * p.strideD;
        int* pg = p.pivot ? p.pivot + (long long)mat * p.stridePivot : nullptr;
        for (int i = tid; i < n; i += nwarps * 32) { Dg[i] = dvec[i]; if (pg) pg[i] = perm[i]; }
    }
    {
        double* Tg = p.T + (long long)mat * p.strideT;
        double dinv[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) { const int row = lane + 32 * r; dinv[r] = (row < n) ? 1.0 / dvec[row] : 0.0; }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int s = warp * 8 + c;
            if (s < nloc) {
                const int js = colstep[s];
                const int oc = p.pivot_applied ? (s * CS + rank) : js;
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    const int row = lane + 32 * r;
                    if (row < n) Tg[row + (long long)oc * ld] = (row <= js) ? a[c][r] * dinv[r] : 0.0;
                }
            }
        }
    }

    DQMC_TICK(11);
    // ---- explicit Q, backwards (UDT.jl:272-288); warps are independent from here on --------------
    int cmax = -1;                                       // largest column owned by this warp
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int s = warp * 8 + c;
        const int col = s * CS + rank;
        if (s < nloc) cmax = col;
#pragma unroll
        for (int r = 0; r < RPL; ++r) a[c][r] = (s < nloc && lane + 32 * r == col) ? 1.0 : 0.0;
    }
    if (cmax >= 0 && !gm.skip_q) {
        // reflector k only touches columns >= k, so this warp starts at k = cmax
        double vn[RPL];
        auto load_v = [&](int k, double (&dst)[RPL]) {   // V columns are stored complete (0 .. 0 1 v), ldv = 32 * RPL
            const double* src = Vg + (long long)(k < 0 ? 0 : k) * ldv + lane;
#pragma unroll
            for (int r = 0; r < RPL; ++r) dst[r] = src[32 * r];
        };
        load_v(cmax, vn);
        for (int k = cmax; k >= 0; --k) {
            double v[RPL];
#pragma unroll
            for (int r = 0; r < RPL; ++r) v[r] = vn[r];
            load_v(k - 1, vn);                           // prefetch the next vector from L2
            const double tau = taus[k];
            const int r0 = k >> 5;
            unsigned m = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int s = warp * 8 + c;
                if (s < nloc && s * CS + rank >= k) m |= 1u << c;
            }
            col_dots<RPL>(a, v, m, part, r0);
            warp_allreduce8(part, lane);
            col_update<RPL, false>(a, v, m, part, tau, 0ull, r0);
        }
    }
    DQMC_TICK(12);
    {
        double* Ug = p.U + (long long)mat * p.strideU;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int s = warp * 8 + c;
            if (s < nloc) {
                const int col = s * CS + rank;
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    const int row = lane + 32 * r;
                    if (row < n) Ug[row + (long long)col * ld] = a[c][r];
                }
            }
        }
    }
}

template <int RPL>
static cudaError_t launch_reg(const UdtParams& p, const UdtRegGeom& g, cudaStream_t st)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(p.batch * g.cs));
    cfg.blockDim = dim3((unsigned)(g.nwarps * 32));
    cfg.dynamicSmemBytes = g.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)g.cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    ++g_kernel_launches;
    return cudaLaunchKernelEx(&cfg, udt_reg_kernel<RPL>, p, g);
}

bool udt_reg_supported(int n) { UdtRegGeom g; return udt_reg_geometry(n, g); }

cudaError_t launch_udt_reg(const UdtParams& p, cudaStream_t st)
{
    if (p.batch <= 0) return cudaSuccess;
    UdtRegGeom g;
    if (!udt_reg_geometry(p.n, g)) return cudaErrorInvalidConfiguration;
    static const bool want_dbg = getenv("DQMC_UDT_DBG") != nullptr;
    static long long* dbg_buf = nullptr;
    if (want_dbg) {
        if (!dbg_buf) cudaMalloc(&dbg_buf, 16 * sizeof(long long));
        cudaMemsetAsync(dbg_buf, 0, 16 * sizeof(long long), st);
        g.dbg = dbg_buf;
    }
    cudaError_t err;
    switch (g.rpl) {
    case 1: err = launch_reg<1>(p, g, st); break;
    case 2: err = launch_reg<2>(p, g, st); break;
    case 3: err = launch_reg<3>(p, g, st); break;
    case 4: err = launch_reg<4>(p, g, st); break;
    case 5: err = launch_reg<5>(p, g, st); break;
    case 6: err = launch_reg<6>(p, g, st); break;
    case 7: err = launch_reg<7>(p, g, st); break;
    case 8: err = launch_reg<8>(p, g, st); break;
    default: err = launch_reg<9>(p, g, st); break;
    }
    if (want_dbg && err == cudaSuccess) {
        long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
        static const char* nm[13] = {"load+first publish", "winner+reflector", "v+retire", "dots", "allreduce8", "update+norms",
                                     "reduce8", "warp candidate", "syncthreads", "publish", "vstore+cluster wait", "outputs D/T", "form Q"};
        fprintf(stderr, "[udt dbg n=%d cs=%d] cycles of CTA 0 thread 0:", p.n, g.cs);
        for (int i = 0; i < 13; ++i) fprintf(stderr, " %s=%lld", nm[i], h[i]);
        fprintf(stderr, "\n");
    }
    return err;
}

}  // namespace dqmc
