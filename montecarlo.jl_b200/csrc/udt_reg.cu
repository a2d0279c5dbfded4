// udt_reg.cu -- register-resident, multi-level batched column-pivoted Householder QR -> UDT.
//
// Same mathematics and outputs as the reference's udt_AVX_pivot! (src/flavors/DQMC/linalg/UDT.jl:216-334:
// indmaxcolumn :175-192, reflector! :157-172, reflectorApply! :53-70, Q accumulation :272-288,
// D = |diag R| with 0 -> 1 :293-301, T = D^-1 R P^T or the unpivoted upper triangle :311-334).
//
// Residence.  The n x n/CS column panel of a CTA lives in REGISTERS:
//   * cluster of CS CTAs per matrix, columns dealt cyclically (column c -> CTA c % CS),
//   * inside a CTA warp w owns 8 consecutive local columns, lane l owns rows l, l+32, ...
//     => thread holds a[8][RPL] doubles, RPL = ceil(n/32) (n = 256: 64 doubles = 128 registers),
//   * per Householder step every lane needs only its RPL entries of v, the 8 column dot products are
//     reduced with a recursive-halving shuffle tree, the squared norms of the remaining columns are
//     recomputed in the same pass (the reference recomputes them every step as well),
//   * ONE cluster barrier per step: each CTA publishes its best remaining column (norm, index, column
//     tail) into every peer's shared memory through DSMEM; every CTA then builds the identical
//     reflector redundantly.  Columns are never swapped (un-pivoting is free).
//
// Levels.  A step costs ~3 us of latency (barrier, selection, shuffle trees) whatever the size of the
// trailing matrix, and a 512 KB matrix needs 4 SMs, so only 33 of the 296 matrices of a cfg-4 launch
// are in flight.  The factorisation is therefore cut into levels: level 0 does steps 0 .. n/2 on the
// cluster, writes the rows of R it has finished and exports the (compacted) trailing block; level 1
// factors that (n/2) x (n/2) block -- which fits ONE SM, so 148 matrices are in flight and there is no
// cluster barrier -- and so on down to 64 columns.  Every level recomputes the column norms from
// scratch (as the reference does at every step), so the arithmetic is unchanged.
//
// Q.  Formed by a separate full-grid kernel, backwards (UDT.jl:272-288), with NO block-level
// synchronisation: each warp streams the Householder vectors from L2 into registers one step ahead.
//
// Bound: latency of the per-step critical path; FP64 FMA pipe ~13 % busy (profiles/r1_summary.md).
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace dqmc {

struct UdtLevel {
    // geometry
    int cs, nwarps, nloc, nv, rpl; size_t smem;
    // problem
    int n;          // size of this level's (sub)matrix
    int jstop;      // Householder steps done at this level (== n at the last level)
    int joff;       // steps done by the previous levels == row/step offset of all outputs
    int ld;         // leading dimension of the input
    const double* A; long long strideA;          // level 0: the caller's matrix; level > 0: trailing block
    const int* cmap; long long strideCmap;       // physical column of local column k (nullptr: identity)
    double* S; int ldS; long long strideS;       // trailing block out ((n - jstop)^2), if jstop < n
    int* cmap_out; long long strideCmapOut;
    double* Tphys; long long strideTp;           // T in physical column order (ld = p.ld)
};

static bool udt_level_geometry(int n, UdtLevel& g)
{
    const int rpl = (n + 31) / 32;
    if (rpl > 9) return false;
    const int maxw = (rpl <= 4) ? 16 : ((rpl <= 6) ? 12 : 8);  // matches the __launch_bounds__ below
    for (int cs = 1; cs <= 8; cs *= 2) {
        const int nloc = (n + cs - 1) / cs;
        const int w = (nloc + 7) / 8;
        if (w <= maxw) {
            g.cs = cs; g.nwarps = w; g.nloc = nloc; g.rpl = rpl; g.nv = rpl * 32;
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
    // 4*bit4 + 2*bit3 + bit2 of its lane id, summed over 8 lanes' worth of data
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

// cluster barrier with release/acquire at cluster scope (cg::cluster_group::sync() adds a GPU-scope
// MEMBAR in front of the same barrier; DSMEM + cluster-scope ordering is all we need)
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_barrier() { cluster_arrive(); cluster_wait(); }

// dot products of v with all 8 columns over register rows r0..RPL-1 (v is 0 on rows < j).  The 8
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
__device__ __forceinline__ void col_update(double (&a)[8][RPL], const double (&v)[RPL], double (&part)[8],
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

// ================================================================================================
// QR steps [0, jstop) of one level
// ================================================================================================
template <int RPL>
__global__ void __launch_bounds__((RPL <= 4) ? 512 : ((RPL <= 6) ? 384 : 256))
udt_steps_kernel(const UdtParams p, const UdtLevel L)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = L.cs, nv = L.nv, n = L.n, jstop = L.jstop, joff = L.joff;
    const int csh = 31 - __clz(CS);                      // CS is 1, 2, 4 or 8
    const int rank = (int)cluster.block_rank();
    const int mat = blockIdx.x / CS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = L.nwarps;
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

    const double* Ag = L.A + (long long)mat * L.strideA;
    const int* cmap = L.cmap ? L.cmap + (long long)mat * L.strideCmap : nullptr;
    double* Vg = p.Vwork + (long long)mat * p.strideV;
    const int ld = L.ld, ldv = p.ldv;

    // ---- load the panel into registers ------------------------------------------------------
    double a[8][RPL];
    unsigned act = 0;                                    // warp-uniform: bit c set <=> column still active
    double part[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int s = warp * 8 + c;
        const bool have = s < nloc;
        const int col = s * CS + rank;
        const double sc = (have && joff == 0 && p.colscale.mode) ? scale_at(p.colscale, mat, col) : 1.0;
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
        // CTA winner: every warp reduces the <= 16 warp candidates with a shuffle butterfly (the order of the
        // comparisons does not matter: (norm, -column) is a total order), instead of every thread scanning them
        const int wl = lane & 15;
        double bv = (wl < nwarps) ? wbval[q * 32 + wl] : -2.0;
        int bc = (wl < nwarps) ? wbcol[q * 32 + wl] : 0x7fffffff;
#pragma unroll
        for (int o = 1; o <= 8; o <<= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
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

    if (CS > 1) cluster_barrier();                       // peers resident before any DSMEM store
    else __syncthreads();

    double vprev[RPL];                                   // Householder vector of the previous step (stored late)
    bool store_prev = false;
#pragma unroll
    for (int r = 0; r < RPL; ++r) vprev[r] = 0.0;
    // Householder vector (0 .. 0 1 v) of step jj -> column joff + jj of V, rows joff .. ; rows < joff are zero
    auto store_v = [&](int jj, const double (&vv)[RPL]) {
        double* col = Vg + (long long)(joff + jj) * ldv;
        for (int i = lane; i < joff; i += 32) col[i] = 0.0;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int row = lane + 32 * r;
            if (joff + row < ldv) col[joff + row] = vv[r];
        }
    };

    for (int j = 0; j < jstop; ++j) {
        const int q = j & 1;
        // ---- pick and publish this CTA's best remaining column, then ONE cluster barrier ------
        warp_candidate(q);
        __syncthreads();
        publish(j);
        if (CS > 1) cluster_arrive(); else __syncthreads();
        // the previous Householder vector goes to global memory (for Q) between arrive and wait, so
        // that the release fence of the barrier never has to wait for these stores
        if (store_prev) store_v(j - 1, vprev);
        if (CS > 1) cluster_wait();
        // ---- global winner, identical in every CTA ------------------------------------------
        // cluster winner, identical in every CTA and warp: butterfly over the <= 8 CTA candidates; the owner of
        // column c is CTA c % CS (cyclic dealing)
        const int cl = lane & 7;
        double bv = (cl < CS) ? candval[q * 8 + cl] : -2.0;
        int bc = (cl < CS) ? candcol[q * 8 + cl] : 0x7fffffff;
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
            if (ov > bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
        }
        const int br = (bv >= 0.0) ? (bc & (CS - 1)) : 0;
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
        const int r0 = j >> 5;                           // first register row that can be >= j
        double v[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int row = lane + 32 * r;
            double x = (row > j) ? raw[row] * inv : ((row == j) ? 1.0 : 0.0);
            v[r] = (row < n) ? x : 0.0;
        }
        const unsigned long long m0 = (lane + 32 * r0 > j) ? ~0ull : 0ull;   // rows of register row r0 that are > j
        const bool i_store_v = (rank == br && warp == 0);
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

        // ---- apply H_j to the active columns of this warp, fused norm recompute -------------
        if (act != 0u) {                                 // warp-uniform
            col_dots<RPL>(a, v, act, part, r0);
            warp_allreduce8(part, lane);
            col_update<RPL, true>(a, v, part, tau, m0, r0);
            mynorm = warp_reduce8(part, lane);
        }
        store_prev = i_store_v;
#pragma unroll
        for (int r = 0; r < RPL; ++r) vprev[r] = v[r];
    }
    if (store_prev) store_v(jstop - 1, vprev);
    __syncthreads();                                     // dvec / taus / perm / colstep of the last step visible

    // ---- D, tau, pivot of this level ---------------------------------------------------------------
    if (rank == 0) {
        double* Dg = p.D + (long long)mat * p.strideD + joff;
        double* tg = p.tau + (long long)mat * p.strideTau + joff;
        int* pg = p.pivot ? p.pivot + (long long)mat * p.stridePivot + joff : nullptr;
        for (int i = tid; i < jstop; i += nwarps * 32) {
            Dg[i] = dvec[i]; tg[i] = taus[i];
            if (pg) { const int pc = perm[i]; pg[i] = cmap ? cmap[pc] : pc; }
        }
    }
    // ---- rows joff .. of T (physical column order): finished columns completely, active ones up to jstop
    {
        double* Tg = L.Tphys + (long long)mat * L.strideTp;
        const int n_tot = p.n;
        double dinv[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) { const int row = lane + 32 * r; dinv[r] = (row < jstop) ? 1.0 / dvec[row] : 0.0; }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int s = warp * 8 + c;
            if (s < nloc) {
                const int col = s * CS + rank;
                const int pc = cmap ? cmap[col] : col;
                const int js = colstep[s];
                double* tc = Tg + joff + (long long)pc * p.ld;
                if (js >= 0) {                           // pivoted at this level: rows <= js are R, the rest 0
#pragma unroll
                    for (int r = 0; r < RPL; ++r) {
                        const int row = lane + 32 * r;
                        if (joff + row < n_tot) tc[row] = (row <= js) ? a[c][r] * dinv[r] : 0.0;
                    }
                } else {                                 // still active: rows < jstop are final (R12)
#pragma unroll
                    for (int r = 0; r < RPL; ++r) {
                        const int row = lane + 32 * r;
                        if (row < jstop) tc[row] = a[c][r] * dinv[r];
                    }
                }
            }
        }
    }
    // ---- export the compacted trailing block for the next level --------------------------------------
    if (jstop < n) {
        double* Sg = L.S + (long long)mat * L.strideS;
        int* cmo = L.cmap_out + (long long)mat * L.strideCmapOut;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int s = warp * 8 + c;
            if (s < nloc && colstep[s] < 0) {
                const int col = s * CS + rank;
                // compact index = number of still-active columns with a smaller index
                //               = col - #(pivoted columns < col); the pivoted set is perm[0 .. jstop)
                int cnt = 0;
                for (int i = lane; i < jstop; i += 32) cnt += (perm[i] < col) ? 1 : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                const int k = col - cnt;
                if (lane == 0) cmo[k] = cmap ? cmap[col] : col;
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    const int row = lane + 32 * r;
                    if (row >= jstop && row < n) Sg[(row - jstop) + (long long)k * L.ldS] = a[c][r];
                }
            }
        }
    }
}

// ================================================================================================
// explicit Q, blocked: four reflectors at a time in compact WY form,
//     H_k H_{k+1} H_{k+2} H_{k+3} = I - V T V^T   (T 4 x 4 upper triangular, LAPACK dlarft "forward, columnwise").
// The per-reflector kernel above is bound by its dependent chain (dots -> shuffle tree -> update) once per
// reflector; here one tree serves four reflectors and the 32 dot-product chains of a warp are independent,
// so the kernel runs at the FP64 pipe instead of at shuffle latency.  The four Householder vectors of a
// block are staged once per CTA in shared memory (cp.async, double buffered) instead of once per warp from L2.
// ================================================================================================

// T factors of every block of four reflectors: T4[mat][g][j + 4 * i] = T[j][i]
__global__ void __launch_bounds__(128)
udt_wy_t_kernel(const UdtParams p, double* __restrict__ T4, int ngroups)
{
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= p.batch * ngroups) return;
    const int mat = gw / ngroups, g = gw - mat * ngroups;
    const int n = p.n, ldv = p.ldv;
    const double* Vg = p.Vwork + (long long)mat * p.strideV;
    const double* tg = p.tau + (long long)mat * p.strideTau;
    double s01 = 0, s02 = 0, s03 = 0, s12 = 0, s13 = 0, s23 = 0;
    const int k0 = 4 * g;
    for (int row = k0 + lane; row < ldv; row += 32) {           // vectors are zero above their own index
        double v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (k0 + j < n) ? Vg[(long long)(k0 + j) * ldv + row] : 0.0;
        s01 = fma(v[0], v[1], s01); s02 = fma(v[0], v[2], s02); s03 = fma(v[0], v[3], s03);
        s12 = fma(v[1], v[2], s12); s13 = fma(v[1], v[3], s13); s23 = fma(v[2], v[3], s23);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s01 += __shfl_xor_sync(0xffffffffu, s01, o); s02 += __shfl_xor_sync(0xffffffffu, s02, o);
        s03 += __shfl_xor_sync(0xffffffffu, s03, o); s12 += __shfl_xor_sync(0xffffffffu, s12, o);
        s13 += __shfl_xor_sync(0xffffffffu, s13, o); s23 += __shfl_xor_sync(0xffffffffu, s23, o);
    }
    if (lane == 0) {
        double t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = (k0 + j < n) ? tg[k0 + j] : 0.0;
        double T[4][4] = {};
        T[0][0] = t[0]; T[1][1] = t[1]; T[2][2] = t[2]; T[3][3] = t[3];
        T[0][1] = -t[1] * (T[0][0] * s01);
        T[0][2] = -t[2] * (T[0][0] * s02 + T[0][1] * s12);
        T[1][2] = -t[2] * (T[1][1] * s12);
        T[0][3] = -t[3] * (T[0][0] * s03 + T[0][1] * s13 + T[0][2] * s23);
        T[1][3] = -t[3] * (T[1][1] * s13 + T[1][2] * s23);
        T[2][3] = -t[3] * (T[2][2] * s23);
        double* out = T4 + ((long long)mat * ngroups + g) * 16;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) out[j + 4 * i] = T[j][i];
    }
}

__device__ __forceinline__ void cp_async16_q(void* smem, const void* gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(sa), "l"(gmem));
}

template <int RPL>
__global__ void __launch_bounds__(256)
udt_formq4_kernel(const UdtParams p, const double* __restrict__ T4, int ngroups)
{
    constexpr int NV = RPL * 32;
    const int n = p.n, ld = p.ld, ldv = p.ldv;
    const int ctas_per_mat = (n + 63) / 64;
    const int mat = blockIdx.x / ctas_per_mat, part_i = blockIdx.x - mat * ctas_per_mat;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col0 = part_i * 64 + warp * 8;             // this warp owns columns col0 .. col0 + 7
    const double* Vg = p.Vwork + (long long)mat * p.strideV;
    const double* Tg = T4 + (long long)mat * ngroups * 16;
    double* Ug = p.U + (long long)mat * p.strideU;

    // CH blocks of four reflectors per stage: one CTA barrier per 16 reflectors, so the warps drift apart and one
    // warp's shuffle tree overlaps another's FMAs
    constexpr int CH = 4;
    extern __shared__ __align__(16) double fq_sm[];
    double* vsb = fq_sm;                                 // [2][CH * 4][NV]
    double* tsb = fq_sm + (size_t)2 * CH * 4 * NV;       // [2][CH][16]
#define VS(s_, jj_, r_) vsb[((size_t)(s_) * CH * 4 + (jj_)) * NV + (r_)]
#define TS(s_, gb_, i_) tsb[((s_) * CH + (gb_)) * 16 + (i_)]

    double a[8][RPL];
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int r = 0; r < RPL; ++r) a[c][r] = (col0 + c < n && lane + 32 * r == col0 + c) ? 1.0 : 0.0;

    const int ctop = min(n, part_i * 64 + 64) - 1;       // highest column of this CTA
    const int gtop = ctop >> 2;
    const int wtop = (col0 < n) ? (min(n - 1, col0 + 7) >> 2) : -1;   // highest block that touches this warp
    const bool dense = (ldv == NV);
    auto stage = [&](int cidx, int s) {                  // blocks CH cidx .. CH cidx + CH - 1 -> stage s
        for (int e = tid; e < CH * 4 * (NV / 2); e += 256) {
            const int jj = e / (NV / 2), r2 = (e - jj * (NV / 2)) * 2;
            const int k = CH * 4 * cidx + jj;
            if (k < n && (dense || r2 + 1 < ldv)) cp_async16_q(&VS(s, jj, r2), Vg + (long long)k * ldv + r2);
            else { VS(s, jj, r2) = 0.0; VS(s, jj, r2 + 1) = 0.0; }
        }
        if (tid < CH * 16) {
            const int g = CH * cidx + (tid >> 4);
            TS(s, tid >> 4, tid & 15) = (g < ngroups) ? Tg[(long long)g * 16 + (tid & 15)] : 0.0;
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };

    const int ctopc = gtop / CH;
    stage(ctopc, 0);
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
    for (int cidx = ctopc; cidx >= 0; --cidx) {
        const int s = (ctopc - cidx) & 1;
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();                                 // chunk visible; everybody is done with the previous one
        if (cidx > 0) stage(cidx - 1, s ^ 1);
      for (int gb = CH - 1; gb >= 0; --gb) {
        const int g = CH * cidx + gb;
        if (g > gtop || g > wtop) continue;              // warp-uniform
        const int r0 = (4 * g) >> 5;                     // first register row a reflector of this block touches
        // ---- W = V^T A: 32 independent chains ---------------------------------------------------
        double d[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 8; ++c) d[j][c] = 0.0;
#pragma unroll
        for (int r = 0; r < RPL; ++r)
            if (r >= r0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double v = VS(s, gb * 4 + j, lane + 32 * r);
#pragma unroll
                    for (int c = 0; c < 8; ++c) d[j][c] = fma(v, a[c][r], d[j][c]);
                }
            }
        // ---- one reduction tree for the whole block: halve over the columns, butterfly over the rest --
        double y1[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double send = h16 ? d[j][c] : d[j][c + 4];
                const double keep = h16 ? d[j][c + 4] : d[j][c];
                y1[j][c] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
        double y2[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const double send = h8 ? y1[j][c] : y1[j][c + 2];
                const double keep = h8 ? y1[j][c + 2] : y1[j][c];
                y2[j][c] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
        double w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double send = h4 ? y2[j][0] : y2[j][1];
            const double keep = h4 ? y2[j][1] : y2[j][0];
            w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] += __shfl_xor_sync(0xffffffffu, w[j], 2);
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] += __shfl_xor_sync(0xffffffffu, w[j], 1);
        // ---- Y = T W for the column this lane ended up with (col_of_lane) ------------------------
        double y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int i = j; i < 4; ++i) acc = fma(TS(s, gb, j + 4 * i), w[i], acc);
            y[j] = acc;
        }
        // ---- A -= V Y, four columns at a time (register budget) ------------------------------------
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            double yy[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int cc = half * 4 + c;
                const int src = ((cc & 4) ? 16 : 0) | ((cc & 2) ? 8 : 0) | ((cc & 1) ? 4 : 0);
#pragma unroll
                for (int j = 0; j < 4; ++j) yy[j][c] = -__shfl_sync(0xffffffffu, y[j], src);
            }
#pragma unroll
            for (int r = 0; r < RPL; ++r)
                if (r >= r0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double v = VS(s, gb * 4 + j, lane + 32 * r);
#pragma unroll
                        for (int c = 0; c < 4; ++c) a[half * 4 + c][r] = fma(v, yy[j][c], a[half * 4 + c][r]);
                    }
                }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int col = col0 + c;
        if (col < n) {
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const int row = lane + 32 * r;
                if (row < n) Ug[row + (long long)col * ld] = a[c][r];
            }
        }
    }
}
#undef VS
#undef TS

// ================================================================================================
// host side
// ================================================================================================
template <int RPL>
static cudaError_t launch_steps(const UdtParams& p, const UdtLevel& g, cudaStream_t st)
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
    count_launch();
    return cudaLaunchKernelEx(&cfg, udt_steps_kernel<RPL>, p, g);
}

template <int RPL>
static cudaError_t launch_formq4(const UdtParams& p, double* T4, cudaStream_t st)
{
    const int ngroups = (p.n + 3) / 4;
    const int warps = p.batch * ngroups;
    count_launch();
    udt_wy_t_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, st>>>(p, T4, ngroups);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int ctas = p.batch * ((p.n + 63) / 64);
    count_launch();
    constexpr int smem = (2 * 4 * 4 * RPL * 32 + 2 * 4 * 16) * (int)sizeof(double);
    // (Q held in the DMMA accumulator layout with the update as one DMMA per 8 rows was measured SLOWER: 1.38 vs
    //  0.96 ms per 296 x 256^2 -- see DESIGN.md; it is not part of the build)
    static SmemAttr attr;
    e = attr.ensure(udt_formq4_kernel<RPL>, smem);
    if (e != cudaSuccess) return e;
    udt_formq4_kernel<RPL><<<(unsigned)ctas, 256, smem, st>>>(p, T4, ngroups);
    return cudaGetLastError();
}

#define DQMC_RPL_SWITCH(rpl, CALL) \
    switch (rpl) { \
    case 1: { constexpr int R = 1; err = CALL; } break; case 2: { constexpr int R = 2; err = CALL; } break; \
    case 3: { constexpr int R = 3; err = CALL; } break; case 4: { constexpr int R = 4; err = CALL; } break; \
    case 5: { constexpr int R = 5; err = CALL; } break; case 6: { constexpr int R = 6; err = CALL; } break; \
    case 7: { constexpr int R = 7; err = CALL; } break; case 8: { constexpr int R = 8; err = CALL; } break; \
    default: { constexpr int R = 9; err = CALL; } break; }

bool udt_reg_supported(int n) { UdtLevel g{}; return udt_level_geometry(n, g); }

// Level sizes are the sizes at which the geometry gets cheaper: <= 256 needs 4 SMs per matrix (37 matrices
// in flight), <= 192 two (74: 12 warps of 8 columns x 6 register rows), <= 128 one (148).  Returns the size
// of the trailing block the level that starts with nk columns hands on (0: it finishes the factorisation).
static int udt_next_level_size(int nk)
{
    if (nk <= 64) return 0;
    static const int sizes[4] = {256, 192, 128, 64};
    for (int k = 0; k < 4; ++k)
        if (sizes[k] < nk) return sizes[k];
    return 0;
}

// scratch needed per matrix besides Vwork: Tphys (ld * n doubles), one trailing-block buffer per level
// (rounded-up leading dimensions) and two column maps (2 n ints)
size_t udt_reg_scratch_doubles(int n, int ld)
{
    size_t tot = (size_t)ld * n + 64;
    for (int nk = udt_next_level_size(n); nk > 0; nk = udt_next_level_size(nk)) tot += (size_t)((nk + 1) & ~1) * nk;
    tot += (size_t)((n + 3) / 4) * 16;                   // T factors of the blocked form-Q
    return tot;
}
size_t udt_reg_scratch_ints(int n) { return (size_t)2 * n + 16; }

cudaError_t launch_udt_reg(const UdtParams& p, cudaStream_t st)
{
    if (p.batch <= 0) return cudaSuccess;
    if (!p.scratch || !p.iscratch) return cudaErrorInvalidValue;
    const int n = p.n;
    // scratch layout: [Tphys: batch x ld*n] [S level 0: batch x ...] [S level 1: batch x ...] ...
    double* Tphys_base = p.scratch;
    const long long strideTp = (long long)p.ld * n;
    double* S_base = p.scratch + (size_t)p.batch * strideTp;
    const bool direct_T = p.pivot_applied != 0;          // physical order IS the requested output

    int nk = n, joff = 0, level = 0;
    const double* Ain = p.A; long long strideIn = p.strideA; int ldin = p.ld;
    const int* cmap_in = nullptr; long long strideCm = 0;
    size_t s_off = 0;
    cudaError_t err = cudaSuccess;
    while (nk > 0) {
        UdtLevel g{};
        if (!udt_level_geometry(nk, g)) return cudaErrorInvalidConfiguration;
        const int jstop = nk - udt_next_level_size(nk);
        g.n = nk; g.jstop = jstop; g.joff = joff; g.ld = ldin;
        g.A = Ain; g.strideA = strideIn; g.cmap = cmap_in; g.strideCmap = strideCm;
        g.Tphys = direct_T ? p.T : Tphys_base; g.strideTp = direct_T ? p.strideT : strideTp;
        const int n2 = nk - jstop;
        if (n2 > 0) {
            g.ldS = (n2 + 1) & ~1; g.strideS = (long long)g.ldS * n2;
            g.S = S_base + s_off; s_off += (size_t)p.batch * g.strideS;
            g.cmap_out = p.iscratch + (size_t)(level & 1) * p.batch * n; g.strideCmapOut = n;
        }
        DQMC_RPL_SWITCH(g.rpl, (launch_steps<R>(p, g, st)))
        if (err != cudaSuccess) return err;
        if (n2 > 0) {
            Ain = g.S; strideIn = g.strideS; ldin = g.ldS;
            cmap_in = g.cmap_out; strideCm = n;
        }
        joff += jstop; nk = n2; ++level;
    }
    // Q
    const int rpl = (n + 31) / 32;
    double* T4 = S_base + s_off;                         // behind the trailing-block buffers
    DQMC_RPL_SWITCH(rpl, (launch_formq4<R>(p, T4, st)))
    if (err != cudaSuccess) return err;
    // the Val(false) form wants the columns of D^-1 R in pivot (logical) order
    if (!direct_T)
        err = launch_permute_cols(Tphys_base, p.T, p.pivot, n, p.ld, strideTp, p.stridePivot, p.batch, st);
    return err;
}

}  // namespace dqmc
