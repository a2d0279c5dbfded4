// update3.cu -- sweep_spatial in "submatrix" form: block-restricted proposals, G0 + Bc X Br flush.
//
// Same mathematics and decisions as update.cu / the reference (`sweep_spatial`, local_updates.jl:23-60;
// `propose_local` / `calculate_detratio!`, fields.jl:388-393, 440-449, 63-84; `accept_local!` ->
// `update_greens!`, fields.jl:340-344, 271-286).  One CTA per Markov chain, the sites in blocks I of kb:
//
//   1. serial phase   Only G[I, I] (per flavor) is touched, and it lives in REGISTERS (one 4 x 4 patch per
//      thread).  All lanes of warp 0 evaluate the next 32 proposals at once against the current diagonal; up
//      to the first accepted one these are exactly the sequential decisions (a rejected proposal changes
//      nothing).  An accepted flip applies the reference's rank-1 update to the kb x kb patches and records
//      the block-restricted column / row of its factors.
//   2. X              After the block, all k accepted flips together are  G' = G0 + Bc X Br  with
//      Bc = G0[:, I] - E_I, Br = G0[I, :] and a kb x kb matrix X = MU MW that follows from the recorded
//      restrictions alone (two triangular recurrences on k x k unit matrices: k^3 flops, no n in sight).
//   3. flush          Per flavor: the accepted columns of G0 (Bc) and the row panel G0[I, :] are staged in
//      shared memory with cp.async, T = X Br is formed in place (one thread per column), and G += Bc T
//      streams over G with DMMA from shared memory at HBM rate (the flush of update.cu).
//
// The serial part no longer scales with n, and shared memory holds the factors of ONE flavor at a time, so
// kb = 44 instead of 24 at n = 256 with two flavors: 6 instead of 11 read-modify-write passes over G.
//
// Measured and not kept (round 2): blocks delimited by ACCEPTED flips instead of sites -- the panels only ever hold the
// accepted columns / rows, so a block can take proposals from a window of up to 64 sites (two diagonal entries per lane)
// until kb flips are accepted: 5 instead of 8 passes over G at cfg 5, 4-5 instead of 6 at cfg 4.  The serial phase and the
// X build grow faster than the flush shrinks: cfg 5 (one flavor, 4 x 4 patches on all 256 threads) window 40 / 48 / 56 / 64
// = 0.624 / 0.640 / 0.640 / 0.680 ms per slice visit; cfg 4 (two flavors: 4 x 8 patches) 0.97 against 0.83 ms.  Smaller
// blocks lose as well (cfg 4: 36 -> 0.850, 32 -> 0.850, 24 -> 1.04): the kb the host picks is the optimum.
// Measured in the last session of round 2 and not kept: (a) taking the column / row of an accepted site from the register patches
// with one uniform switch on j & 3 instead of the selects (what gave slice_steps_kernel +7 % at cfg 2): cfg 4 252.8 -> 253.6 ms
// of update per sweep, cfg 5 124.9 -> 126.1; (b) on top of it, compile-time thread count and division-free index arithmetic in
// the X build (the run-time `e / (k * k)`, `r % k` cost ~25 instructions each): cfg 4 250.9, cfg 3 49.6 -> 49.3, but cfg 5
// 128.4-128.9 -- ptxas' spill pattern of this 255-register kernel moves with every change (NB = 2: 96 -> 168 bytes of stack,
// NB = 1: 120 -> 48) and decides more than the instructions saved.
// (c) evaluating only the half of the sites that holds the next proposal (what slice_steps_kernel does now): cfg 4 252.0, cfg 5
// 125.3, cfg 3 49.3 -- within noise of the kept kernel.  With two warps per scheduler the serial phase follows the dependent chain
// of an accept (decision -> ballot -> shuffle -> extraction -> barrier -> rank-1 update of the diagonal), not the length of the
// instruction stream around it.
// (d) 128-bit loads of the exchanged column / row after the barrier (the 64-bit loads of the column are 4-way bank conflicted;
// +4.5 % at cfg 2 in slice_steps_kernel): cfg 4 252.5, cfg 5 125.5, cfg 3 49.3 -- no change.
// (e) explicitly double-buffered fragments in the T = X Br loop (7 % of the samples, DMMA issue separated by NOPs): cfg 4 245.4 ->
// 249.0, cfg 3 47.7 -> 48.9, cfg 5 122.5 -> 120.8 (stack frame 96 -> 176 bytes at NB = 2): dropped for the headline configuration.
// Also without effect: walking the tiles of the flush in reverse order on every other block, so that a flush starts with the
// tiles the previous flush of the same flavor wrote last (L2 reuse: 155 MB of G against 126 MB of L2): 0.800 vs 0.799 ms.
#include "common.cuh"
#include "../../include/dqmc_rng.h"
#include <math.h>
#include <algorithm>

namespace dqmc {

constexpr int U3_KBT = 48;              // compile-time bound of kb (unrolled accept loops)
constexpr int U3_RP = U3_KBT + 2;       // row stride of the kb x kb work matrices
constexpr int PF_DIST = 5;              // flush: L2 prefetch distance in tile rounds

__device__ __forceinline__ void dmma884v(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void u3_cp_async8(void* smem, const void* gmem)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(a), "l"(gmem));
}
__device__ __forceinline__ void u3_cp_async16(void* smem, const void* gmem)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(a), "l"(gmem));
}
__device__ __forceinline__ void u3_cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
}
__device__ __forceinline__ double u3_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

static inline int update3_ldu(int n) { return n + (((4 - n) % 16) + 16) % 16; }   // == 4 mod 16

struct Upd3Shared {
    int k, next, acc_site;
    double coef[2];
    int xs[U3_KBT];
    double coefs[2][U3_KBT];
};

// Row `r` of the k x k matrix M with M[:, a] = e_a + sum_{a' < a} M[:, a'] S[a'][a] (times coefs[a] if COEF):
// the triangular recurrences behind U = Bc_acc MU and W = MW Br_acc.  One thread, its row in registers.
template <bool COEF>
__device__ __forceinline__ void tri_row(int r, int k, const double* __restrict__ S, const double* __restrict__ coefs,
                                        double* __restrict__ out)
{
    double hist[U3_KBT];
#pragma unroll
    for (int t = 0; t < U3_KBT / 8; ++t) {
        const int a0 = 8 * t;
        double acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = (a0 + q == r) ? 1.0 : 0.0;
        if (a0 < k) {
#pragma unroll
            for (int ap = 0; ap < a0; ++ap) {
                const double h = hist[ap];
                const double2* s2 = reinterpret_cast<const double2*>(S + (size_t)ap * U3_RP + a0);
#pragma unroll
                for (int q2 = 0; q2 < 4; ++q2) {
                    const double2 sv = s2[q2];
                    acc[2 * q2] = fma(h, sv.x, acc[2 * q2]);
                    acc[2 * q2 + 1] = fma(h, sv.y, acc[2 * q2 + 1]);
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
#pragma unroll
                for (int qp = 0; qp < q; ++qp) acc[q] = fma(acc[qp], S[(size_t)(a0 + qp) * U3_RP + a0 + q], acc[q]);
                if (a0 + q >= k) acc[q] = 0.0;              // (also keeps uninitialised S entries out)
                else if (COEF) acc[q] *= coefs[a0 + q];
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            hist[a0 + q] = acc[q];
            out[a0 + q] = acc[q];
        }
    }
}

template <int NB>
__global__ void __launch_bounds__(256)
update3_kernel(const UpdateParams p, const int ldu, const double em2a, const double ep2a)
{
    extern __shared__ __align__(16) double sm[];
    constexpr int nb = NB, RP = U3_RP;
    const int n = p.n, kb = p.kb, ld = p.ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x, nwarps = NT >> 5;
    const int chain = blockIdx.x;

    // ---- shared memory -----------------------------------------------------------------------------------
    const size_t RS = max((size_t)kb * ldu, (size_t)2 * nb * kb * RP);   // a panel or the work matrices aliasing it
    double* P1 = sm;                                        // [kb][ldu]  Bc_acc: accepted columns of G0 (minus e)
    double* P2 = P1 + RS;                                   // [kb][ldu]  row panel G0[I, :] -> T = X Br
    double* XR = P2 + RS;                                   // [nb][kb][RP]  X restricted to accepted rows: Xr[a][x]
    double* colv = XR + (size_t)nb * kb * RP;
    double* rowv = colv + 2 * nb * kb;                      // colv, rowv: [2 (parity of the accept)][nb][kb]
    double* sunif = rowv + 2 * nb * kb;                     // [n]
    Upd3Shared* sh = (Upd3Shared*)(sunif + n);
    int8_t* sconf = (int8_t*)(sh + 1);                      // [n]
    int8_t* sxnew = sconf + n;                              // [n] proposed value of every site (GHQ: drawn up front)
    // work matrices of the serial phase / X build alias the panels (each nb * kb * RP doubles)
    double* Ub = P1;                                        // [nb][kb (site x)][RP (accept a)]; later MU
    double* Wb = P1 + (size_t)nb * kb * RP;                 // [nb][kb (accept a)][RP (site y)]; later MWt
    double* SU = P2;                                        // [nb][kb][RP]
    double* SWt = P2 + (size_t)nb * kb * RP;

    double* G = p.G + (long long)chain * nb * p.strideG;
    int8_t* conf = p.conf_slice + (long long)chain * p.cstride;
    const double* utab = p.uniforms ? p.uniforms + (long long)chain * p.ustride : nullptr;
    const unsigned char* forced = p.forced ? p.forced + (long long)chain * p.tstride : nullptr;

    const unsigned long long sweep_now = (unsigned long long)(p.sweep_ptr ? *p.sweep_ptr : p.sweep);
    for (int i = tid; i < n; i += NT) {
        const int8_t x = conf[i];
        sconf[i] = x;
        sunif[i] = utab ? utab[i]
                        : dqmc_uniform(p.seed, (uint64_t)(p.chain0 + chain), sweep_now, (uint32_t)p.step, (uint32_t)i);
        if (p.kind >= 2) {                                  // x_new = choices[x_old, rand(1:3)] (fields.jl:528, 590)
            const double u2 = utab ? utab[n + i]
                                   : dqmc_uniform_choice(p.seed, (uint64_t)(p.chain0 + chain), sweep_now, (uint32_t)p.step, (uint32_t)i);
            sxnew[i] = (int8_t)dqmc_ghq_choice((int)x, u2);
        } else sxnew[i] = (int8_t)(-x);
    }

    int accepted = 0;
    double neg_cnt = 0.0, neg_sum = 0.0, neg_min = INFINITY, neg_max = -INFINITY;   // per lane of warp 0

    // register patches of G[I, I]: thread -> (flavor ob, patch (pbx, pby)) of 4 x 4 entries
    const int NPB = (kb + 3) >> 2;
    const bool owner = tid < nb * NPB * NPB;
    const int ob = tid / (NPB * NPB), ot = tid % (NPB * NPB);
    const int oxb = (ot % NPB) * 4, oyb = (ot / NPB) * 4;

    for (int i0 = 0; i0 < n; i0 += kb) {
        const int kbc = (n - i0 < kb) ? (n - i0) : kb;
        __syncthreads();                                    // previous flush (global G), sconf / sunif visible
        // ---- load G[I, I] into registers, its diagonal into shared memory ---------------------------------
        double g[4][4];
#pragma unroll
        for (int iy = 0; iy < 4; ++iy)
#pragma unroll
            for (int ix = 0; ix < 4; ++ix) {
                const int x = oxb + ix, y = oyb + iy;
                g[ix][iy] = (owner && x < kbc && y < kbc) ? G[(long long)ob * p.strideG + (i0 + x) + (long long)(i0 + y) * ld] : 0.0;
            }
        // Every lane of EVERY warp tracks the diagonal entries of the sites lane and lane + 32 in registers and takes
        // the decisions redundantly: (site, Delta / R) of an accepted flip are known to all threads without a
        // broadcast, so an accept costs one CTA barrier (between the extraction of its restricted column / row and
        // their use; colv / rowv are double buffered).
        double gd[2][2], un[2];
        Proposal pr[2];
        int fc[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = lane + 32 * h;
            const bool in = j < kbc;
            pr[h] = make_proposal(p.kind, in ? (int)sconf[i0 + j] : 1, in ? (int)sxnew[i0 + j] : 1, p.ghq, em2a, ep2a);
            un[h] = in ? sunif[i0 + j] : 2.0;
            fc[h] = (in && forced) ? (int)forced[i0 + j] : -1;
#pragma unroll
            for (int b = 0; b < nb; ++b)
                gd[b][h] = in ? G[(long long)b * p.strideG + (i0 + j) + (long long)(i0 + j) * ld] : 0.0;
        }
        __syncthreads();                                    // panels of the previous block are no longer read

        // ---- serial phase --------------------------------------------------------------------------------
        int k = 0, next = 0;
        for (;;) {
            double prob[2], cf[2][2];
            int acc[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = lane + 32 * h;
                const bool elig = (j >= next) && (j < kbc);
                double Rv[2];
#pragma unroll
                for (int b = 0; b < nb; ++b) {
                    Rv[b] = fma(pr[h].Dl[b], 1.0 - gd[b][h], 1.0);
                    cf[b][h] = pr[h].Dl[b] * u3_rcp(Rv[b]);            // Delta / R (vldiv22!), speculative
                }
                if (nb == 1) cf[1][h] = cf[0][h];
                prob[h] = proposal_prob(p.kind, pr[h], (nb == 1) ? Rv[0] * Rv[0] : Rv[0] * Rv[1]);
                int a_ = (fc[h] >= 0) ? (fc[h] != 0) : ((prob[h] > 1.0) ? 1 : (un[h] < prob[h]));
                acc[h] = elig ? a_ : 0;
            }
            const unsigned b0 = __ballot_sync(0xffffffffu, acc[0]);
            const unsigned b1 = __ballot_sync(0xffffffffu, acc[1]);
            const int jacc = b0 ? (__ffs(b0) - 1) : (b1 ? 32 + __ffs(b1) - 1 : -1);
            if (warp == 0) {                                // traces and statistics of the real decisions
                const int jlast = (jacc >= 0) ? jacc : kbc - 1;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = lane + 32 * h;
                    if (j >= next && j <= jlast) {
                        if (p.check_sign && prob[h] < 0.0) {
                            neg_cnt += 1.0; neg_sum += log10(fabs(prob[h]));
                            neg_min = fmin(neg_min, prob[h]); neg_max = fmax(neg_max, prob[h]);
                        }
                        if (p.probs) p.probs[(long long)chain * p.tstride + i0 + j] = prob[h];
                        if (p.decisions) p.decisions[(long long)chain * p.tstride + i0 + j] = (unsigned char)(j == jacc);
                    }
                }
            }
            if (jacc < 0) break;
            const int src = jacc & 31, hh = jacc >> 5;
            const double c0 = __shfl_sync(0xffffffffu, hh ? cf[0][1] : cf[0][0], src);
            const double c1 = __shfl_sync(0xffffffffu, hh ? cf[1][1] : cf[1][0], src);
            const int a = k, j = jacc;
            if (warp == 0 && lane == src) {
                sconf[i0 + j] = sxnew[i0 + j]; conf[i0 + j] = sconf[i0 + j];
                sh->xs[a] = j; sh->coefs[0][a] = c0; sh->coefs[1][a] = c1;
            }
            double* cvb = colv + (a & 1) * nb * kb;
            double* rvb = rowv + (a & 1) * nb * kb;
            if (owner) {
                if (j >= oyb && j < oyb + 4) {              // this patch holds part of column j
#pragma unroll
                    for (int ix = 0; ix < 4; ++ix) {
                        double v = 0.0;
#pragma unroll
                        for (int iy = 0; iy < 4; ++iy) v = (oyb + iy == j) ? g[ix][iy] : v;
                        const int x = oxb + ix;
                        const double cv = v - ((x == j) ? 1.0 : 0.0);
                        cvb[ob * kb + x] = cv;
                        Ub[((size_t)ob * kb + x) * RP + a] = cv;
                    }
                }
                if (j >= oxb && j < oxb + 4) {              // ... part of row j
                    const double cc = ob ? c1 : c0;
#pragma unroll
                    for (int iy = 0; iy < 4; ++iy) {
                        double v = 0.0;
#pragma unroll
                        for (int ix = 0; ix < 4; ++ix) v = (oxb + ix == j) ? g[ix][iy] : v;
                        const int y = oyb + iy;
                        const double rv = cc * v;
                        rvb[ob * kb + y] = rv;
                        Wb[((size_t)ob * kb + a) * RP + y] = rv;
                    }
                }
            }
            __syncthreads();
            if (owner) {
                double cv[4], rv[4];
#pragma unroll
                for (int ix = 0; ix < 4; ++ix) cv[ix] = cvb[ob * kb + oxb + ix];
#pragma unroll
                for (int iy = 0; iy < 4; ++iy) rv[iy] = rvb[ob * kb + oyb + iy];
#pragma unroll
                for (int iy = 0; iy < 4; ++iy)
#pragma unroll
                    for (int ix = 0; ix < 4; ++ix) g[ix][iy] = fma(cv[ix], rv[iy], g[ix][iy]);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int jj = lane + 32 * h;
                if (jj < kbc) {
#pragma unroll
                    for (int b = 0; b < nb; ++b) gd[b][h] = fma(cvb[b * kb + jj], rvb[b * kb + jj], gd[b][h]);
                }
            }
            k = a + 1; next = j + 1;
        }
        if (tid == 0) sh->k = k;
        accepted += k;
        __syncthreads();
        if (k == 0) continue;                               // uniform over the CTA: nothing to flush

        // ---- X = MU MW restricted to the accepted sites ---------------------------------------------------------
        // SU[a'][a] = Wb[a'][x_a], SWt[a'][a] = Ub[x_a][a']
        for (int e = tid; e < nb * k * U3_KBT; e += NT) {
            const int b = e / (k * U3_KBT), r = e - b * (k * U3_KBT);
            const int a = r % U3_KBT, ap = r / U3_KBT;
            if (a < k) {
                const int xa = sh->xs[a];
                SU[((size_t)b * kb + ap) * RP + a] = Wb[((size_t)b * kb + ap) * RP + xa];
                SWt[((size_t)b * kb + ap) * RP + a] = Ub[((size_t)b * kb + xa) * RP + ap];
            }
        }
        __syncthreads();
        {
            // group 0 .. nb-1: rows of MU (over Ub), group nb .. 2nb-1: rows of MWt (over Wb)
            const int grp = tid / U3_KBT, r = tid - grp * U3_KBT;
            if (grp < 2 * nb && r < k) {
                const int b = grp % nb;
                if (grp < nb) tri_row<false>(r, k, SU + (size_t)b * kb * RP, sh->coefs[b], Ub + ((size_t)b * kb + r) * RP);
                else tri_row<true>(r, k, SWt + (size_t)b * kb * RP, sh->coefs[b], Wb + ((size_t)b * kb + r) * RP);
            }
        }
        __syncthreads();
        // X_acc[b][a1][a2] = sum_a MU[a1][a] MWt[a2][a] (compact: accepted flips only), zero padded
        for (int e = tid; e < nb * kb * RP; e += NT) XR[e] = 0.0;
        __syncthreads();
        {
            // on the tensor pipe: 8 x 8 tiles of X = MU MWt^T, one tile per warp at a time (as one scalar dot product per thread
            // with two run-time divisions for its indices this loop took 10 % of the kernel's samples).  MU[a1][a] = 0 for
            // a < a1 and MWt[a2][a] = 0 for a < a2, so the k loop of tile (ti, tj) starts at 8 max(ti, tj); rows >= k of the
            // work matrices hold stale data that only reaches entries which are not stored.
            const int nt = (k + 7) >> 3, k4x = (k + 3) >> 2;
            const float inv_nt = 1.0f / (float)nt;
            const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
            for (int b = 0; b < nb; ++b)
                for (int tl = warp; tl < nt * nt; tl += nwarps) {
                    const int ti = (int)(((float)tl + 0.5f) * inv_nt), tj = tl - ti * nt;
                    const double* mu = Ub + ((size_t)b * kb + 8 * ti + gq) * RP + tq;
                    const double* mw = Wb + ((size_t)b * kb + 8 * tj + gq) * RP + tq;
                    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;        // two accumulator pairs: half the dependent chain
                    int s4 = 2 * (ti > tj ? ti : tj);
                    for (; s4 + 1 < k4x; s4 += 2) {
                        dmma884v(c0, c1, mu[4 * s4], mw[4 * s4]);
                        dmma884v(d0, d1, mu[4 * s4 + 4], mw[4 * s4 + 4]);
                    }
                    if (s4 < k4x) dmma884v(c0, c1, mu[4 * s4], mw[4 * s4]);
                    const int a1 = 8 * ti + gq, a2 = 8 * tj + 2 * tq;
                    if (a1 < k) {
                        if (a2 < k) XR[((size_t)b * kb + a1) * RP + a2] = c0 + d0;
                        if (a2 + 1 < k) XR[((size_t)b * kb + a1) * RP + a2 + 1] = c1 + d1;
                    }
                }
        }
        __syncthreads();

        // ---- flush, one flavor at a time: G_b += Bc_acc (Xr Br) -----------------------------------------------------
        const int g8 = lane >> 2, t4 = lane & 3;
        const int tiles = (n + 31) / 32;
        const int k4 = (k + 3) / 4;
        for (int b = 0; b < nb; ++b) {
            double* Gb = G + (long long)b * p.strideG;
            // accepted columns of G0 -> P1[a][r]; row panel G0[I, :] -> P2[x][c]
            for (int a = warp; a < k; a += nwarps) {
                const double* src = Gb + (long long)(i0 + sh->xs[a]) * ld;
                for (int r2 = lane * 2; r2 < n; r2 += 64) {
                    if (r2 + 1 < n) u3_cp_async16(P1 + (size_t)a * ldu + r2, src + r2);
                    else u3_cp_async8(P1 + (size_t)a * ldu + r2, src + r2);
                }
            }
            for (int c = warp; c < n; c += nwarps)
                for (int x = lane; x < kbc; x += 32) u3_cp_async8(P2 + (size_t)x * ldu + c, Gb + (i0 + x) + (long long)c * ld);
            u3_cp_async_wait_all();
            __syncthreads();
            if (tid < k) P1[(size_t)tid * ldu + i0 + sh->xs[tid]] -= 1.0;       // Bc = G0[:, I] - E_I
            const int ntl = tiles * tiles;
            const bool full = (n & 31) == 0;
            auto tile_ptr = [&](int tl, int& tm, int& tn) -> double* {
                tm = (tl % tiles) * 32; tn = (tl / tiles) * 32;
                return Gb + (tm + g8) + (long long)(tn + 2 * t4) * ld;      // element (mi, nj, e) at + mi * 8 + (nj * 8 + e) * ld
            };
            auto load_tile = [&](int tl, double (&dst)[4][4][2]) {
                int tm, tn;
                const double* gp = tile_ptr(tl, tm, tn);
#pragma unroll
                for (int nj = 0; nj < 4; ++nj)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double* gc = gp + (long long)(nj * 8 + e) * ld;
                        const bool cok = full || (tn + nj * 8 + 2 * t4 + e < n);
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi)
                            dst[mi][nj][e] = (cok && (full || tm + mi * 8 + g8 < n)) ? gc[mi * 8] : 0.0;
                    }
            };
            // the first G tile of this warp travels while T is formed (it does not depend on T)
            double cur[4][4][2];
            if (warp < ntl) load_tile(warp, cur);
            // T = X_acc Br_acc in place with DMMA: a warp owns 32 columns of the panel, reads its accepted rows
            // (gathered through xs) for all k before it writes rows 0 .. 4 k4 - 1 of the same columns.  Rows
            // k .. 4 k4 - 1 of X are zero, so the pad rows of the panel come out as zeros.
            const double* xr = XR + (size_t)b * kb * RP;
            for (int c0 = warp * 32; c0 < n; c0 += 32 * nwarps) {
                double tacc[U3_KBT / 8][4][2];
#pragma unroll
                for (int mi = 0; mi < U3_KBT / 8; ++mi)
#pragma unroll
                    for (int nj = 0; nj < 4; ++nj) tacc[mi][nj][0] = tacc[mi][nj][1] = 0.0;
                for (int kk = 0; kk < k4; ++kk) {
                    const int ap = kk * 4 + t4;
                    const int xa = (ap < k) ? sh->xs[ap] : 0;               // X column ap is zero for ap >= k
                    double bf[4];
#pragma unroll
                    for (int nj = 0; nj < 4; ++nj) {         // columns >= n are never stored: do not read past the row
                        const int cc = c0 + nj * 8 + g8;    // (another warp writes what follows it)
                        bf[nj] = (cc < n) ? P2[(size_t)xa * ldu + cc] : 0.0;
                    }
#pragma unroll
                    for (int mi = 0; mi < U3_KBT / 8; ++mi)
                        if (mi * 8 < 4 * k4) {              // uniform
                            const double af = (mi * 8 + g8 < kb) ? xr[(size_t)(mi * 8 + g8) * RP + ap] : 0.0;
#pragma unroll
                            for (int nj = 0; nj < 4; ++nj) dmma884v(tacc[mi][nj][0], tacc[mi][nj][1], af, bf[nj]);
                        }
                }
                __syncwarp();
#pragma unroll
                for (int mi = 0; mi < U3_KBT / 8; ++mi)
                    if (mi * 8 < 4 * k4) {
                        const int row = mi * 8 + g8;
                        if (row < 4 * k4) {
#pragma unroll
                            for (int nj = 0; nj < 4; ++nj)
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int c = c0 + nj * 8 + 2 * t4 + e;
                                    if (c < n) P2[(size_t)row * ldu + c] = tacc[mi][nj][e];
                                }
                        }
                    }
            }
            for (int e = tid; e < (4 * k4 - k) * ldu; e += NT) P1[(size_t)k * ldu + e] = 0.0;   // pad rows (P2's: per column above)
            __syncthreads();
            // G_b += sum_{a < k} P1[a][:] P2[a][:]^T, 32 x 32 tiles per warp with a one-tile look-ahead.  The pad rows
            // k .. 4 k4 - 1 of both panels are zero, so the k loop needs no predicates; rows / columns >= n of an edge
            // tile read whatever follows in shared memory and are never stored.
            {
                double nxt[4][4][2];
                for (int tl = warp; tl < ntl; tl += nwarps) {
                    {   // L2 prefetch of the tile PF_DIST rounds ahead: one column segment (256 B) per lane.  (Measured: 0.824 ->
                        // 0.788 ms per slice visit at cfg 4 with 3 rounds; prefetching a whole flavor during the serial
                        // phase thrashes L2 -- 155 MB of G for 148 chains -- and is slower: 0.871 ms.)
                        const int tp = tl + PF_DIST * nwarps;
                        if (tp < ntl) {
                            const int pm = (tp % tiles) * 32, pn = (tp / tiles) * 32 + lane;
                            if (pn < n) {
                                const double* pp = Gb + pm + (long long)pn * ld;
                                asm volatile("prefetch.global.L2 [%0];" :: "l"(pp));
                                asm volatile("prefetch.global.L2 [%0];" :: "l"(pp + 16));
                            }
                        }
                    }
                    if (tl + nwarps < ntl) load_tile(tl + nwarps, nxt);
                    int tm, tn;
                    double* gp = tile_ptr(tl, tm, tn);
                    const double* pa = P1 + (size_t)t4 * ldu + tm + g8;
                    const double* pb = P2 + (size_t)t4 * ldu + tn + g8;
                    for (int kk = 0; kk < k4; ++kk) {
                        double af[4], bf[4];
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi) af[mi] = pa[mi * 8];
#pragma unroll
                        for (int nj = 0; nj < 4; ++nj) bf[nj] = pb[nj * 8];
                        pa += 4 * ldu; pb += 4 * ldu;
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                            for (int nj = 0; nj < 4; ++nj) dmma884v(cur[mi][nj][0], cur[mi][nj][1], af[mi], bf[nj]);
                    }
#pragma unroll
                    for (int nj = 0; nj < 4; ++nj)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            double* gc = gp + (long long)(nj * 8 + e) * ld;
                            const bool cok = full || (tn + nj * 8 + 2 * t4 + e < n);
#pragma unroll
                            for (int mi = 0; mi < 4; ++mi) {
                                if (cok && (full || tm + mi * 8 + g8 < n)) gc[mi * 8] = cur[mi][nj][e];
                                cur[mi][nj][e] = nxt[mi][nj][e];
                            }
                        }
                }
            }
            __syncthreads();                                // P1 / P2 are restaged for the next flavor / block
        }
    }

    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            neg_cnt += __shfl_xor_sync(0xffffffffu, neg_cnt, o); neg_sum += __shfl_xor_sync(0xffffffffu, neg_sum, o);
            neg_min = fmin(neg_min, __shfl_xor_sync(0xffffffffu, neg_min, o));
            neg_max = fmax(neg_max, __shfl_xor_sync(0xffffffffu, neg_max, o));
        }
    }
    if (tid == 0) {
        if (p.accepted) p.accepted[chain] += accepted;
        if (p.stats && neg_cnt > 0.0) {
            double* s = p.stats + (long long)chain * 4;
            s[0] += neg_cnt; s[1] += neg_sum; s[2] = fmin(s[2], neg_min); s[3] = fmax(s[3], neg_max);
        }
    }
}

static size_t update3_smem(int n, int nb, int kb)
{
    const int ldu = update3_ldu(n);
    const size_t rs = std::max((size_t)kb * ldu, (size_t)2 * nb * kb * U3_RP);
    return (2 * rs + (size_t)nb * kb * U3_RP + 4 * nb * kb + n) * sizeof(double) + sizeof(Upd3Shared) + 2 * n + 16;
}

int update3_pick_kb(int n, int nb)
{
    // constraints: kb <= 48 (unrolled accept loops); (kb / 4)^2 * nb register patches <= 256 threads; 227 KB
    int best = 4;
    for (int kb = 4; kb <= U3_KBT; kb += 4) {
        const int npb = kb / 4;
        if (npb * npb * nb > 256) break;
        if (update3_smem(n, nb, kb) > 225 * 1024) break;
        best = kb;
    }
    // no point in a block larger than the lattice; prefer the smallest kb with the same number of passes
    const int nn = (n + 3) & ~3;
    if (best > nn) best = nn;
    const int passes = (n + best - 1) / best;
    while (best > 4 && (n + (best - 4) - 1) / (best - 4) == passes) best -= 4;
    return best;
}

cudaError_t launch_update3(const UpdateParams& p, cudaStream_t st)
{
    if (p.n_chains <= 0) return cudaSuccess;
    if (p.kb < 4 || p.kb > U3_KBT || (p.kb & 3) || p.nb < 1 || p.nb > 2) return cudaErrorInvalidValue;
    const int ldu = update3_ldu(p.n);
    const size_t smem = update3_smem(p.n, p.nb, p.kb);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    const double em2a = exp(-2.0 * p.alpha), ep2a = exp(2.0 * p.alpha);
    static SmemAttr attr[2];
    cudaError_t e = (p.nb == 1) ? attr[0].ensure(update3_kernel<1>, smem) : attr[1].ensure(update3_kernel<2>, smem);
    if (e != cudaSuccess) return e;
    if (p.nb == 1) update3_kernel<1><<<(unsigned)p.n_chains, 256, smem, st>>>(p, ldu, em2a, ep2a);
    else update3_kernel<2><<<(unsigned)p.n_chains, 256, smem, st>>>(p, ldu, em2a, ep2a);
    count_launch();
    return cudaGetLastError();
}

}  // namespace dqmc
