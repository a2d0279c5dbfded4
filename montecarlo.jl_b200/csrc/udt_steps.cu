// udt_steps.cu -- the Householder steps of the batched column-pivoted QR behind udt_AVX_pivot!
// (reference src/flavors/DQMC/linalg/UDT.jl:216-334: indmaxcolumn :175-192, reflector! :157-172,
// reflectorApply! :53-70; D = |diag R| with 0 -> 1 :293-301).
//
// Residence.  A thread-block cluster of CS CTAs holds one matrix in REGISTERS, columns dealt cyclically over the
// CTAs.  Inside a warp the 32 lanes form an 8 x 4 grid (g = lane >> 2, t = lane & 3): a warp owns 4 * CPT local
// columns, lane (g, t) holds CPT of them (slot (warp * CPT + e) * 4 + t) and, of each, the rows 8 i + g, i < RPT
// -- the m8n8 accumulator tiling of the FP64 tensor-core instruction, chosen here because it makes every
// reduction of a Householder step SHORT: a column dot product is thread-local over RPT rows and then a 3-round
// xor butterfly over g that leaves the total in all 8 lanes (the previous layout -- lane = row, 8 columns per
// thread -- needed a 5-round recursive-halving tree plus a gather, 17 64-bit shuffles against 3 per column).
// The Householder vector is read from shared memory (two passes of RPT / 2 LDS.128, conflict-free layout
// [g][RPT + 2]) instead of living in registers, which frees the registers for 16 warps per CTA at n = 256
// (4 warps per scheduler to hide the dependent-latency chains; the old kernel had 2).
//
// One Householder step (no cluster barrier, no fence: data moves with st.async + transaction mbarriers):
//   A  every thread knows the squared norms of its CPT columns (identical in the 8 lanes of a column); the warp's best
//      is three REDUX instructions on the order-preserving bit pattern of the norm; lane 0 posts it; __syncthreads.
//      (The sqrt / rsqrt of the warp's best norm is started speculatively before the barrier.)
//   B  every warp checks with one vote whether any posted candidate beats its own; the warp that holds the CTA's best
//      stages the column in shared memory, finishes the reflector scalars and writes the FINISHED vector
//      (0 .. 0 1 v) and (norm, column, tau, R_jj) into its send buffer; ONE bulk copy per peer
//      (cp.async.bulk.shared::cluster.shared::cta ... mbarrier::complete_tx::bytes) delivers it to every CTA of the
//      cluster -- speculatively: all CS CTAs send their own candidate, only the cluster-wide winner is used.  The
//      receiving mbarrier counts bytes, so there is no barrier.cluster (which costs MEMBAR.ALL.GPU + UCGABAR +
//      CCTL.IVALL per step) and no release fence.
//   C  warp 0 waits for the phase of the CTA's exchange mbarrier (CS entries), __syncthreads releases the others
//      (sixteen warps polling try_wait cost 90 issue slots per warp and step).
//   D  every thread picks the winner (three REDUX over the <= 8 records), then
//        dots:    d_c = v . a_c           RPT FMAs per column + 3 shuffle rounds
//        update:  a_c -= v (tau d_c)      RPT FMAs per column, fused with the recomputation of the remaining squared
//                                         norms (the reference recomputes them from scratch every step too,
//                                         UDT.jl:175-192) + 3 shuffle rounds.
//      Columns are never swapped (un-pivoting of T = D^-1 R P^T is free); the pivot column is simply retired.
//   The Householder vector goes to global memory (for the Q kernel) from the winner's copy, one element per thread.
//   Buffers are double buffered by step parity; the data dependences of the algorithm order their reuse (a CTA can
//   only send step j + 2 after every CTA has sent step j + 1, i.e. has finished reading step j).
//
// Where a step goes (clock64 stamps inside the kernel, 128-column level, one CTA per matrix, ~4600 cycles): selection 510,
// vote 520, the winning warp's reflector (stage 250, scalars 170, scaled vector 590) ~1000 during which the other fifteen
// warps wait, winner record 250, dots 510, update + norms 930, barriers ~100.  Measured and not kept: the eight lanes that
// hold the column write the scaled vector from their registers instead of all lanes from the staged copy (no change);
// every warp finishing the reflector of its OWN candidate before the barrier, speculatively, so that nobody waits for the
// winner (3.35 instead of 3.04 ms per 296 x 256^2 call: sixteen warps each issuing the ~250 extra instructions cost more
// than fifteen warps waiting for one).
//
// Levels: see udt_reg.cu (the factorisation is cut where the geometry gets cheaper).
// Bound: dependent-instruction latency of the step (FP64 pipe < 20 % busy); DESIGN.md section 3.2 has the numbers.
#include <cooperative_groups.h>

#include <algorithm>

#include "udt_level.cuh"

namespace cg = cooperative_groups;

namespace dqmc {

// Launch bound of every (rows per thread, columns per thread) instantiation.  ptxas allocates registers for thread counts
// in steps of 128 (512 threads -> 128 registers, 384 -> 168, 256 -> 255), so the bound is the largest of those three
// under which the panel (2 RPT CPT registers) plus the ~64-register working set compiles without spilling
// (checked with -Xptxas -v: CPT = 1 fits 128 registers up to RPT = 36).
__host__ __device__ constexpr int udt_max_threads(int rpt, int cpt)
{
    const int panel = 2 * rpt * cpt;
    return (panel <= 72) ? 512 : ((panel <= 112) ? 384 : 256);
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }

// ---- mbarrier / DSMEM primitives (PTX ISA 8.x, sm_90+) ---------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "MBW_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra MBD_%=;\n"
                 "bra MBW_%=;\n"
                 "MBD_%=:\n"
                 "}" :: "r"(bar), "r"(parity) : "memory");
}
// bulk copy (DMA) from this CTA's shared memory into the shared memory of a cluster peer (or of this CTA); the peer's
// mbarrier is credited `bytes` in ONE transaction (per-element st.async costs one mbarrier update per 16 bytes:
// measured 2.8 vs 2.3 ms on the 256-column level)
__device__ __forceinline__ void bulk_copy_to_peer(unsigned rdst, unsigned src, unsigned bytes, unsigned rbar)
{
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(rdst), "r"(src), "r"(bytes), "r"(rbar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// reflector (UDT.jl:157-172) from the squared norm of the column tail and its leading element:
// x_j += nu, tau = x_j / nu, v = x / x_j, R_jj = -nu with nu = sign(x_j) |x|
__device__ __forceinline__ void reflector_scalars(double bv, double xi1, double& tau, double& rjj, double& inv);

// 1/sqrt(x) and sqrt(x) to ~1 ulp: hardware seed (MUFU.RSQ64H) + two coupled Newton steps.  x: positive normal.
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

__device__ __forceinline__ void reflector_scalars(double bv, double xi1, double& tau, double& rjj, double& inv)
{
    if (bv < 1e-290 || bv > 1e290) {                     // exact zero / out of the fast path's range: library math
        if (!(bv > 0.0)) { tau = 0.0; rjj = xi1; inv = 0.0; }
        else {
            const double nu = copysign(sqrt(bv), xi1);
            xi1 += nu;
            rjj = -nu; tau = xi1 / nu; inv = 1.0 / xi1;
        }
    } else {
        double rs, sq;
        fast_rsqrt_sqrt(bv, rs, sq);
        const double nu = copysign(sq, xi1);
        xi1 += nu;                                       // |xi1| >= sqrt(bv): never cancels
        rjj = -nu; tau = xi1 * copysign(rs, nu); inv = fast_rcp(xi1);
    }
}

// (value, column) candidates: larger value wins, ties go to the smaller column -- a total order, so any
// reduction order gives the same winner
__device__ __forceinline__ void cand_merge(double& bv, int& bc, double ov, int oc)
{
    if (ov > bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
}

// Candidates as integers: a squared norm v >= 0 maps to bits(v) + 1 (order preserving, 0 is kept for "no candidate"),
// so the best of a warp is three REDUX instructions (max of the high word, max of the low word among the lanes that
// hold that high word, min of the column among the lanes that hold the maximum) instead of a shuffle butterfly.
__device__ __forceinline__ unsigned long long cand_key(double v)
{
    return (v >= 0.0) ? (unsigned long long)__double_as_longlong(v) + 1ull : 0ull;
}
__device__ __forceinline__ double key_value(unsigned long long k)
{
    return k ? __longlong_as_double((long long)(k - 1ull)) : -1.0;
}
// -> key of the best candidate of the warp; col <- its column (ties: smallest column)
__device__ __forceinline__ unsigned long long warp_best(unsigned long long key, int& col)
{
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned bhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned blo = __reduce_max_sync(0xffffffffu, (hi == bhi) ? lo : 0u);
    const bool mine = (hi == bhi) && (lo == blo);
    col = (int)__reduce_min_sync(0xffffffffu, mine ? (unsigned)col : 0xffffffffu);
    return ((unsigned long long)bhi << 32) | blo;
}

// CPS > 0: HYBRID panel.  Besides its CPT register columns a thread owns CPS columns in a private strip of shared memory
// (row pairs as double2, element (es, pair) of thread tid at ((es * RPT / 2 + pair) * NT + tid): consecutive lanes ->
// conflict-free LDS.128 / STS.128).  The strips double the panel an SM holds, so a matrix needs half the SMs and twice as
// many matrices are in flight; a step pays ~3 shared-memory passes over the strips (LSU bound) on top of its latency chain.
template <int RPT, int CPT, int CPS>
__global__ void __launch_bounds__(udt_max_threads(RPT, CPT), 1)
udt_steps_kernel(const UdtParams p, const UdtLevel L)
{
    static_assert(RPT % 2 == 0, "rows per thread come in LDS.128 pairs");
    constexpr int CT = CPT + CPS;                        // columns per thread: e < CPT in registers, the rest in the strip
    constexpr int VP = RPT + 2;                          // [g][VP]: 2 VP = 4 (mod 8) words -> conflict-free LDS.128 over g
    constexpr int VB = 8 * VP;                           // doubles per published vector
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = L.cs, n = L.n, jstop = L.jstop, joff = L.joff;
    const int rank = (CS > 1) ? (int)cluster.block_rank() : 0;
    const int mat = blockIdx.x / CS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = L.nwarps;
    const int g = lane >> 2, t = lane & 3;
    const int nloc = (n - rank + CS - 1) / CS;           // local columns: slot s <-> column s * CS + rank

    extern __shared__ __align__(16) double sm[];
    constexpr int VE = VB + 4;                           // published entry: vector [g][VP] + record (norm, column, tau, R_jj)
    double* vbuf = sm;                                   // [2][CS][VE]  received entries (parity of the step)
    double* sendbuf = vbuf + (size_t)2 * CS * VE;        // [2][VE]      this CTA's entry (source of the bulk copies)
    double* stage = sendbuf + 2 * VE;                    // [VB]         raw best column of the sending warp
    double* dvec = stage + VB;                           // [n]  |R_jj| (0 -> 1)
    double* taus = dvec + n;                             // [n]
    double* rdia = taus + n;                             // [n]  R_jj
    unsigned long long* wbkey = reinterpret_cast<unsigned long long*>(rdia + n);     // [2][16] per-warp candidates
    unsigned long long* bars = wbkey + 32;               // [0..1] exchange mbarriers
    int* colstep = (int*)(bars + 2);                     // [nwarps * 4 * CT]  step at which a local slot was retired (-1: active)
    int* perm = colstep + nwarps * 4 * CT;               // [n]
    int* wbcol = perm + n;                               // [2][16]
    const int NT = nwarps * 32;
    // strips: behind the int arrays, 16-byte aligned
    // (offset arithmetic on the shared base pointer: an integer round trip would turn the strip accesses into generic loads)
    const size_t strip_off = ((size_t)(reinterpret_cast<const char*>(wbcol + 32) - reinterpret_cast<const char*>(sm)) + 15) & ~(size_t)15;
    double2* strip = reinterpret_cast<double2*>(reinterpret_cast<char*>(sm) + strip_off) + tid;
#define STRIP(es_, pair_) strip[(size_t)((es_) * (RPT / 2) + (pair_)) * NT]
    const unsigned xbar0 = smem_u32(bars);
    const unsigned tx_bytes = (unsigned)CS * (unsigned)(VE * 8);
    const float inv_cs = 1.0f / (float)CS;               // exact small-integer division by the cluster size

    const double* Ag = L.A + (long long)mat * L.strideA;
    const int* cmap = L.cmap ? L.cmap + (long long)mat * L.strideCmap : nullptr;
    double* Vg = p.Vwork + (long long)mat * p.strideV;
    const int ld = L.ld, ldv = p.ldv;

    // ---- load the panel into registers ------------------------------------------------------
    double a[CPT][RPT];
    double nrm[CT];
    unsigned act = 0;                                    // bit e set <=> column e of this thread is still active
#pragma unroll
    for (int e = 0; e < CT; ++e) {
        const int s = (warp * CT + e) * 4 + t;
        const bool have = s < nloc;
        const int col = s * CS + rank;
        const double sc = (have && joff == 0 && p.colscale.mode) ? scale_at(p.colscale, mat, col) : 1.0;
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int i = 0; i < RPT; i += 2) {
            const int r0 = 8 * i + g, r1 = r0 + 8;
            const double x0 = (have && r0 < n) ? Ag[r0 + (long long)col * ld] * sc : 0.0;
            const double x1 = (have && r1 < n) ? Ag[r1 + (long long)col * ld] * sc : 0.0;
            if (e < CPT) { a[e < CPT ? e : 0][i] = x0; a[e < CPT ? e : 0][i + 1] = x1; }
            else STRIP(e - CPT, i >> 1) = make_double2(x0, x1);
            acc0 = fma(x0, x0, acc0); acc1 = fma(x1, x1, acc1);
        }
        double acc = acc0 + acc1;
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        nrm[e] = acc;
        if (have) act |= 1u << e;
        if (g == 0) colstep[s] = -1;
    }

    if (tid == 0) {
        mbar_init(xbar0, 1); mbar_init(xbar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    cluster_arrive(); cluster_wait();                    // once: peers resident and their mbarriers initialised

    for (int j = 0; j < jstop; ++j) {
        const int q = j & 1;
        if (CS > 1 && tid == 0) mbar_arrive_expect_tx(xbar0 + 8 * q, tx_bytes);
        // ---- A: best remaining column of the thread -> warp (three REDUX on the order-preserving bit pattern) -----
        double bv = -1.0; int bc = 0x7fffffff;
#pragma unroll
        for (int e = 0; e < CT; ++e)
            if ((act >> e) & 1u) cand_merge(bv, bc, nrm[e], ((warp * CT + e) * 4 + t) * CS + rank);
        const unsigned long long key_w = warp_best(cand_key(bv), bc);     // bc <- column of the warp's best
        const bool live_w = key_w != 0ull;
        const double bv_w = key_value(key_w);
        double rs_w = 0.0, sq_w = 0.0;
        const bool fast_w = live_w && bv_w >= 1e-290 && bv_w <= 1e290;
        if (fast_w) fast_rsqrt_sqrt(bv_w, rs_w, sq_w);   // speculative: off the owner's critical path
        if (lane == 0) { wbkey[q * 16 + warp] = key_w; wbcol[q * 16 + warp] = bc; }
        __syncthreads();
        // ---- B: the warp that holds the CTA's best sends the finished reflector to every CTA of the cluster -------
        bool sender;
        {
            const unsigned long long ok = (lane < nwarps) ? wbkey[q * 16 + lane] : 0ull;
            const int oc = (lane < nwarps) ? wbcol[q * 16 + lane] : 0x7fffffff;
            const bool beaten = ok > key_w || (ok == key_w && oc < bc);
            const bool any_live = __any_sync(0xffffffffu, ok != 0ull);
            sender = any_live ? (live_w && !__any_sync(0xffffffffu, beaten)) : (warp == 0);
        }
        if (sender) {
            double* wst = stage;
            double tau_w = 0.0, rjj_w = 0.0, inv_w = 0.0;
            if (live_w) {
                const int sw = (int)((float)(bc - rank) * inv_cs + 0.5f);
                const int es = (sw >> 2) % CT, ts = sw & 3;
#pragma unroll
                for (int e = 0; e < CT; ++e)
                    if (e == es && t == ts) {
#pragma unroll
                        for (int i = 0; i < RPT; i += 2)
                            *reinterpret_cast<double2*>(wst + g * VP + i) =
                                (e < CPT) ? make_double2(a[e < CPT ? e : 0][i], a[e < CPT ? e : 0][i + 1]) : STRIP(e - CPT, i >> 1);
                    }
                __syncwarp();
                double xi1 = wst[(j & 7) * VP + (j >> 3)];
                if (fast_w) {                            // reflector (UDT.jl:157-172)
                    const double nu = copysign(sq_w, xi1);
                    xi1 += nu;                           // |xi1| >= sqrt(bv): never cancels
                    rjj_w = -nu; tau_w = xi1 * copysign(rs_w, nu); inv_w = fast_rcp(xi1);
                } else reflector_scalars(bv_w, xi1, tau_w, rjj_w, inv_w);
            }
            // one CTA per matrix: the entry is written in place and the CTA barrier below publishes it -- no copy engine,
            // no proxy fence, no mbarrier round trip on the critical path of the step
            double* sb = (CS > 1) ? sendbuf + q * VE : vbuf + (size_t)q * VE;
#pragma unroll
            for (int it = 0; it < (VB + 63) / 64; ++it) {          // unrolled: the LDS -> DMUL -> STS chains of the rounds overlap
                const int idx = 2 * lane + 64 * it;
                if (idx >= VB) break;
                const int gg = idx / VP, ii = idx - gg * VP;
                double2 y = make_double2(0.0, 0.0);
                if (live_w && ii < RPT) {
                    const double2 x = *reinterpret_cast<const double2*>(wst + idx);
                    const int r0 = 8 * ii + gg, r1 = r0 + 8;
                    y.x = (r0 > j) ? x.x * inv_w : ((r0 == j) ? 1.0 : 0.0);
                    y.y = (r1 > j) ? x.y * inv_w : ((r1 == j) ? 1.0 : 0.0);
                }
                *reinterpret_cast<double2*>(sb + idx) = y;
            }
            if (lane == 0) {
                *reinterpret_cast<double2*>(sb + VB) = make_double2(live_w ? bv_w : -1.0, (double)bc);
                *reinterpret_cast<double2*>(sb + VB + 2) = make_double2(tau_w, rjj_w);
            }
            if (CS > 1) {
                fence_proxy_async();                     // generic-proxy writes -> visible to the bulk-copy engine
                __syncwarp();
                if (lane < CS)
                    bulk_copy_to_peer(mapa_u32(smem_u32(vbuf + ((size_t)q * CS + rank) * VE), (unsigned)lane), smem_u32(sb),
                                      (unsigned)(VE * 8), mapa_u32(xbar0 + 8 * q, (unsigned)lane));
            }
        }
        // ---- C: one warp waits for the CS entries of this step, the hardware barrier releases the others ----------
        if (CS > 1 && warp == 0) mbar_wait(xbar0 + 8 * q, (unsigned)(j >> 1) & 1u);
        __syncthreads();
        // ---- D: cluster winner, identical in every CTA and thread -----------------------------------------
        {
            const int cl = lane & 7;
            const double2 rec = (cl < CS) ? *reinterpret_cast<const double2*>(vbuf + ((size_t)q * CS + cl) * VE + VB)
                                          : make_double2(-1.0, 2147483647.0);
            bc = (int)rec.y;
            bv = key_value(warp_best(cand_key(rec.x), bc));
        }
        const int sglob = (int)(((float)bc + 0.5f) * inv_cs);     // floor(bc / CS): the half keeps float rounding away from the integers
        const int br = (bv >= 0.0) ? (bc - sglob * CS) : 0;
        const double2 trec = *reinterpret_cast<const double2*>(vbuf + ((size_t)q * CS + br) * VE + VB + 2);
        const double tau = trec.x;
        const double* vw = vbuf + ((size_t)q * CS + br) * VE;     // the winner's vector, [g][VP]
        if (tid == 0) {
            const double rjj = trec.y;
            const double ad = fabs(rjj);
            dvec[j] = (ad == 0.0) ? 1.0 : ad;
            taus[j] = tau; rdia[j] = rjj; perm[j] = bc;
        }
        if (rank == br) {
            // retire the pivot column; its Householder vector goes to global memory for the Q kernel
            const int s = sglob;                         // bc = s * CS + rank
            if (warp == s / (4 * CT) && t == (s & 3)) {
                act &= ~(1u << ((s >> 2) % CT));
                if (g == 0) colstep[s] = j;
            }
            double* vcol = Vg + (long long)(joff + j) * ldv;
#pragma unroll 1                                         // (unrolled, the run-time stride costs an integer division for the trip count every step)
            for (int r = tid; r < ldv; r += nwarps * 32) {
                const int lr = r - joff;
                vcol[r] = (lr >= 0 && lr < 8 * RPT) ? vw[(lr & 7) * VP + (lr >> 3)] : 0.0;
            }
        }
        // ---- apply H_j to the active columns, fused recomputation of the remaining norms -------------------
        // (register columns e < CPT and strip columns e >= CPT in the same passes: one read of v, one butterfly phase)
        if (__any_sync(0xffffffffu, act != 0u)) {        // warp-uniform: the butterflies below need all 32 lanes
            constexpr int GS = (RPT % 8 == 0) ? 8 : ((RPT % 4 == 0) ? 4 : 2);   // rows-of-8 per group: one skip test per group
            const int gi0 = (j >> 3) / GS;               // groups below hold only finished rows (v = 0 there)
            const double* vb = vw + g * VP;
            double d0[CT], d1[CT];
#pragma unroll
            for (int e = 0; e < CT; ++e) d0[e] = d1[e] = 0.0;
#pragma unroll
            for (int gi = 0; gi < RPT / GS; ++gi)
                if (gi >= gi0) {                         // uniform over the CTA
#pragma unroll
                    for (int i = gi * GS; i < gi * GS + GS; i += 2) {
                        const double2 v = *reinterpret_cast<const double2*>(vb + i);
#pragma unroll
                        for (int e = 0; e < CT; ++e) {
                            double2 x;
                            if (e < CPT) x = make_double2(a[e < CPT ? e : 0][i], a[e < CPT ? e : 0][i + 1]);
                            else x = STRIP(e - CPT, i >> 1);
                            d0[e] = fma(v.x, x.x, d0[e]); d1[e] = fma(v.y, x.y, d1[e]);
                        }
                    }
                }
            double sd[CT];
#pragma unroll
            for (int e = 0; e < CT; ++e) {
                double d = d0[e] + d1[e];
                d += __shfl_xor_sync(0xffffffffu, d, 4);
                d += __shfl_xor_sync(0xffffffffu, d, 8);
                d += __shfl_xor_sync(0xffffffffu, d, 16);
                sd[e] = ((act >> e) & 1u) ? -tau * d : 0.0;
            }
#pragma unroll
            for (int e = 0; e < CT; ++e) d0[e] = d1[e] = 0.0;      // now the norm accumulators
#pragma unroll
            for (int gi = 0; gi < RPT / GS; ++gi) {
                if (gi > gi0) {
#pragma unroll
                    for (int i = gi * GS; i < gi * GS + GS; i += 2) {
                        const double2 v = *reinterpret_cast<const double2*>(vb + i);
#pragma unroll
                        for (int e = 0; e < CT; ++e) {
                            double x0, x1;
                            if (e < CPT) {
                                x0 = fma(v.x, sd[e], a[e < CPT ? e : 0][i]); x1 = fma(v.y, sd[e], a[e < CPT ? e : 0][i + 1]);
                                a[e < CPT ? e : 0][i] = x0; a[e < CPT ? e : 0][i + 1] = x1;
                            } else {
                                const double2 x = STRIP(e - CPT, i >> 1);
                                x0 = fma(v.x, sd[e], x.x); x1 = fma(v.y, sd[e], x.y);
                                STRIP(e - CPT, i >> 1) = make_double2(x0, x1);
                            }
                            d0[e] = fma(x0, x0, d0[e]); d1[e] = fma(x1, x1, d1[e]);
                        }
                    }
                } else if (gi == gi0) {                  // the only group that can hold rows <= j: they do not count
#pragma unroll
                    for (int i = gi * GS; i < gi * GS + GS; i += 2) {
                        const double2 v = *reinterpret_cast<const double2*>(vb + i);
                        const bool k0 = 8 * i + g > j, k1 = 8 * i + 8 + g > j;
#pragma unroll
                        for (int e = 0; e < CT; ++e) {
                            double x0, x1;
                            if (e < CPT) {
                                x0 = fma(v.x, sd[e], a[e < CPT ? e : 0][i]); x1 = fma(v.y, sd[e], a[e < CPT ? e : 0][i + 1]);
                                a[e < CPT ? e : 0][i] = x0; a[e < CPT ? e : 0][i + 1] = x1;
                            } else {
                                const double2 x = STRIP(e - CPT, i >> 1);
                                x0 = fma(v.x, sd[e], x.x); x1 = fma(v.y, sd[e], x.y);
                                STRIP(e - CPT, i >> 1) = make_double2(x0, x1);
                            }
                            const double m0 = k0 ? x0 : 0.0, m1 = k1 ? x1 : 0.0;
                            d0[e] = fma(m0, m0, d0[e]); d1[e] = fma(m1, m1, d1[e]);
                        }
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < CT; ++e) {
                double d = d0[e] + d1[e];
                d += __shfl_xor_sync(0xffffffffu, d, 4);
                d += __shfl_xor_sync(0xffffffffu, d, 8);
                d += __shfl_xor_sync(0xffffffffu, d, 16);
                nrm[e] = d;
            }
        }
    }
    __syncthreads();                                     // dvec / taus / rdia / perm / colstep of the last step visible

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
#pragma unroll
        for (int e = 0; e < CT; ++e) {
            const int s = (warp * CT + e) * 4 + t;
            if (s < nloc) {
                const int col = s * CS + rank;
                const int pc = cmap ? cmap[col] : col;
                const int js = colstep[s];
                double* tc = Tg + joff + (long long)pc * p.ld;
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    const int row = 8 * i + g;
                    double x;
                    if (e < CPT) x = a[e < CPT ? e : 0][i];
                    else { const double2 xx = STRIP(e - CPT, i >> 1); x = (i & 1) ? xx.y : xx.x; }
                    if (js >= 0) {                       // pivoted at this level: rows < js are R, row js is R_jj, the rest 0
                        if (joff + row < n_tot)
                            tc[row] = (row < js) ? x / dvec[row] : ((row == js) ? rdia[js] / dvec[js] : 0.0);
                    } else {                             // still active: rows < jstop are final (R12)
                        if (row < jstop) tc[row] = x / dvec[row];
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
        for (int e = 0; e < CT; ++e) {
            const int s = (warp * CT + e) * 4 + t;
            const bool live = s < nloc && colstep[s] < 0;
            const int col = s * CS + rank;
            // compact index = number of still-active columns with a smaller index
            //               = col - #(pivoted columns < col); the pivoted set is perm[0 .. jstop); the 8 lanes of a
            // column split the count
            int cnt = 0;
            if (live)
                for (int i = g; i < jstop; i += 8) cnt += (perm[i] < col) ? 1 : 0;
            cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, 8);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, 16);
            if (live) {
                const int k = col - cnt;
                if (g == 0) cmo[k] = cmap ? cmap[col] : col;
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    const int row = 8 * i + g;
                    double x;
                    if (e < CPT) x = a[e < CPT ? e : 0][i];
                    else { const double2 xx = STRIP(e - CPT, i >> 1); x = (i & 1) ? xx.y : xx.x; }
                    if (row >= jstop && row < n) Sg[(row - jstop) + (long long)k * L.ldS] = x;
                }
            }
        }
    }
}
#undef STRIP

// ================================================================================================
// host side: geometry table and launch
// ================================================================================================
// (rows per thread, register columns per thread, strip columns per thread)
#define UDT_FOR_EACH_GEOM(X) \
    X(2, 1, 0) X(2, 2, 0) X(4, 1, 0) X(4, 2, 0) X(8, 1, 0) X(8, 2, 0) X(12, 1, 0) X(12, 2, 0) X(12, 4, 0) X(16, 1, 0) X(16, 2, 0) \
    X(16, 4, 0) X(20, 1, 0) X(20, 2, 0) X(20, 4, 0) X(24, 1, 0) X(24, 2, 0) X(28, 1, 0) X(28, 2, 0) X(32, 1, 0) X(32, 2, 0) \
    X(36, 1, 0) X(36, 2, 0) X(32, 1, 1)

struct UdtKernelEntry { int rpt, cpt, cps, max_threads; const void* fn; };
static const UdtKernelEntry* udt_kernel_table(int& count)
{
#define UDT_ENTRY(R, C, S) {R, C, S, udt_max_threads(R, C), (const void*)udt_steps_kernel<R, C, S>},
    static const UdtKernelEntry table[] = {UDT_FOR_EACH_GEOM(UDT_ENTRY)};
#undef UDT_ENTRY
    count = (int)(sizeof(table) / sizeof(table[0]));
    return table;
}

static size_t udt_steps_smem(int rpt, int cpt, int cps, int cs, int nwarps, int n)
{
    const int vb = 8 * (rpt + 2), ve = vb + 4;
    size_t b = ((size_t)2 * cs * ve + 2 * ve + vb + 3 * (size_t)n + 32 + 2) * sizeof(double) +
               ((size_t)nwarps * 4 * (cpt + cps) + n + 32) * sizeof(int);
    if (cps > 0) b += 16 + (size_t)cps * (rpt / 2) * nwarps * 32 * 16;   // strips (16-byte aligned)
    return b;
}

// Chooses the geometry that keeps the most matrices in flight per SM (registers: what the launch bound of the
// instantiation allows); ties go to more warps per matrix (shorter dependent chains per thread).  Cluster sizes need
// not be powers of two (n = 288: 5 CTAs of 58 columns).
bool udt_steps_geometry(int nk, UdtLevel& g)
{
    int count = 0;
    const UdtKernelEntry* tab = udt_kernel_table(count);
    const int need = (nk + 7) / 8;
    int rpt = 0;
    for (int k = 0; k < count; ++k)
        if (tab[k].rpt >= need && (rpt == 0 || tab[k].rpt < rpt)) rpt = tab[k].rpt;
    if (rpt == 0) return false;
    double best_score = -1.0;
    for (int k = 0; k < count; ++k) {
        if (tab[k].rpt != rpt) continue;
        const int cpt = tab[k].cpt, cps = tab[k].cps, maxw = tab[k].max_threads / 32;
        const int regs = (tab[k].max_threads == 512) ? 128 : ((tab[k].max_threads == 384) ? 168 : 255);
        for (int cs = 1; cs <= 8; ++cs) {
            const int nloc = (nk + cs - 1) / cs;
            const int w = (nloc + 4 * (cpt + cps) - 1) / (4 * (cpt + cps));
            if (w > maxw || w > 16) continue;
            const int threads = w * 32;
            const size_t smem = udt_steps_smem(rpt, cpt, cps, cs, w, nk);
            int per_sm = 65536 / (regs * threads);
            per_sm = std::min(per_sm, 2048 / threads);
            per_sm = std::min(per_sm, (int)((227 * 1024) / (smem + 1024)));
            per_sm = std::min(per_sm, 16);
            if (per_sm < 1) continue;
            // (measured at n = 256: 16 warps x 1 column per thread 1.51 ms, 8 warps x 2 columns 1.57 ms -- latency bound;
            //  preferring FEWER warps at every level: 256 -> 1.58 vs 1.62, 128 -> 0.180 vs 0.172, n = 144 0.80 vs 0.74 ms per call)
            // A hybrid step pays three shared-memory passes over the strips.  Measured on 296 x 256^2: 5.1 us per step on
            // clusters of 2 (74 matrices in flight) against 2.6 us on clusters of 4 (33): 1.29 vs 1.63 ms for the 256-column
            // level (1.48 while the strip columns had their own dot / update passes behind the register columns' instead
            // of sharing them); the 192-column level on ONE hybrid CTA per matrix: 0.69 vs 0.69 ms; n = 288 on hybrid
            // clusters of 3 instead of 5: 2.45 vs 2.39 ms per call.  The factor below keeps exactly the first case.
            const double score = (double)per_sm / cs / (cps > 0 ? 1.9 : 1.0) + 1e-3 * w;
            if (score > best_score) {
                best_score = score;
                g.cs = cs; g.cpt = cpt; g.cps = cps; g.rpt = rpt; g.nwarps = w; g.smem = smem;
            }
            break;                                       // larger clusters only lose from here
        }
    }
    return best_score > 0.0;
}

cudaError_t launch_udt_steps(const UdtParams& p, const UdtLevel& L, cudaStream_t st)
{
    int count = 0;
    const UdtKernelEntry* tab = udt_kernel_table(count);
    const void* fn = nullptr;
    for (int k = 0; k < count; ++k)
        if (tab[k].rpt == L.rpt && tab[k].cpt == L.cpt && tab[k].cps == L.cps) fn = tab[k].fn;
    if (!fn) return cudaErrorInvalidConfiguration;
    if (L.smem > 48 * 1024) {
        static SmemAttr attr[64];                        // one per table entry
        int k = 0;
        for (; k < count; ++k) if (tab[k].fn == fn) break;
        cudaError_t e = attr[k].ensure(fn, L.smem);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(p.batch * L.cs));
    cfg.blockDim = dim3((unsigned)(L.nwarps * 32));
    cfg.dynamicSmemBytes = L.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)L.cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    count_launch();
    UdtParams pp = p; UdtLevel ll = L;
    void* args[2] = {&pp, &ll};
    return cudaLaunchKernelExC(&cfg, fn, args);
}

}  // namespace dqmc
