// slicestep.cu -- small lattices (n <= 64): a run of consecutive slice steps in ONE kernel, G resident on chip.
//
// For small matrices every kernel of the generic path is latency bound and a sweep is ~2M x 6 launches.  Between two
// stabilisations the reference's hot loop is, per time slice (local_updates.jl:7-14, stack.jl:594-603, 605-730):
//     sweep_spatial(mc)                    N proposals, rank-1 Green's function updates (fields.jl:271-286)
//     propagate(mc) -> wrap_greens!        G <- B_l G B_l^-1  (up)   /   G <- B_{l-1}^-1 G B_{l-1}  (down)
// This kernel runs up to safe_mult - 1 of these steps back to back with one CTA per Markov chain and G (all flavor
// blocks) in shared memory:
//   * proposals: every thread evaluates the decision redundantly from the shared diagonal (no broadcast); an accepted
//     flip is the reference's immediate rank-1 update G -= (e_i - G[:, i]) (Delta / R) G[i, :] applied to the shared
//     copy (n^2 FMAs over 256 threads, two CTA barriers) -- at n <= 64 that is cheaper than any delayed scheme;
//   * wrap: two n^3 products on the FP64 tensor pipe (DMMA m8n8k4) straight from shared memory, the hopping exponential
//     read through L1 (it is shared by all chains and stays cached), the diagonal e^{+-V} factors fused as k / row /
//     column scales.
// One flavor block (the register-patch path below), where an accept's ~2200 cycles went according to the ncu source page
// (profiles/r2i_slice_steps_source_regions.txt, before the last three changes): evaluation + extraction 42 % of the kernel's
// samples, the barrier 17 %, the shared-memory loads behind it 22 %, the two wrap products 17 %.  In the build: uniform switch for the
// patch column / row, half evaluation, padded 128-bit exchange, rank-1 update rotated into the next iteration (cfg 2: 9493 -> 10 912
// sweeps/s).  Measured and not kept: a cheaper trace path for warp 0 (one vote in front of the statistics / trace stores): 10 922 vs
// 10 912 -- warp 0 is not what the barrier waits for.
// The host state machine (capi.cu) replaces each run of "sweep_spatial + plain wrap" by one launch and keeps the
// separate kernels for the stabilisation steps.  Decisions are identical to the generic path and to the reference
// for the same uniforms (tests/test_gpu_parity.py covers both via dqmc_desc.update_variant).
#include "common.cuh"
#include "../../include/dqmc_rng.h"
#include <math.h>

namespace dqmc {

__device__ __forceinline__ void ss_dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double ss_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

static inline int ss_ld(int n) { return n + (((4 - n) % 16) + 16) % 16; }   // == 4 mod 16: conflict-free DMMA fragments

// C (n x n, column-major, ldc) = diag(rs) * A * diag(ks) * B * diag(cs); A, B column-major in shared or global memory.
// 8 warps: warp w owns the 16 x 32 block at rows (w & 3) * 16, columns (w >> 2) * 32; out-of-range parts are skipped.
template <bool TA = false>
__device__ __forceinline__ void ss_mm(double* __restrict__ C, int ldc, const double* __restrict__ A, int lda,
                                      const double* __restrict__ B, int ldb, int n, const double* __restrict__ ks,
                                      const double* __restrict__ rs, const double* __restrict__ cs)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
    if (wm >= n || wn >= n) return;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int k0 = 0; k0 < n; k0 += 4) {
        const int k = k0 + t;
        const bool kok = k < n;
        const double sk = (ks && kok) ? ks[k] : 1.0;
        double af[2], bf[4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = wm + 8 * i + g;
            af[i] = (kok && r < n) ? (TA ? A[k + (size_t)r * lda] : A[r + (size_t)k * lda]) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = wn + 8 * j + g;
            bf[j] = (kok && c < n) ? B[k + (size_t)c * ldb] * sk : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) ss_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int r = wm + 8 * i + g;
        if (r >= n) continue;
        const double rv = rs ? rs[r] : 1.0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = wn + 8 * j + 2 * t + e;
                if (c < n) C[r + (size_t)c * ldc] = acc[i][j][e] * rv * (cs ? cs[c] : 1.0);
            }
    }
}

template <int NB>
__global__ void __launch_bounds__(256)
slice_steps_kernel(const SliceStepParams p)
{
    extern __shared__ __align__(16) double sm[];
    constexpr int NT = 256;                                  // the launches below (compile-time strides: no trip-count divisions)
    const int n = p.n, ldg = p.ldg, tid = threadIdx.x;
    const int chain = blockIdx.x;
    double* Gs = sm;                                     // [NB][n][ldg]
    double* Ts = Gs + (size_t)NB * n * ldg;              // [n][ldg] wrap intermediate (one flavor at a time)
    double* xch = Ts + (size_t)n * ldg;                  // [2][2][64] one flavor: column / row of an accept, padded to whole patches,
                                                         //            double buffered by the parity of the accept (128-bit accesses)
    double* IG = xch + 256;                              // [2][n]  e_i - G[:, i]      (two flavors: per flavor)
    double* gr = IG + 2 * n;                             // [2][n]  (Delta / R) G[i, :]
    double* dpos = gr + 2 * n;                           // [NB][n]  e^{+V} of the wrap slice
    double* dneg = dpos + NB * n;                        // [NB][n]  e^{-V}
    double* su = dneg + NB * n;                          // [n] Metropolis uniforms of the step
    int8_t* sconf = (int8_t*)(su + n);                   // [n]
    int8_t* sxnew = sconf + n;                           // [n]

    double* G = p.G + (long long)chain * NB * p.strideG;
    int8_t* conf = p.conf + (long long)chain * p.cstride;
    const unsigned long long sweep_now = (unsigned long long)(p.sweep_ptr ? *p.sweep_ptr : p.sweep);
    const double em2a = p.em2a, ep2a = p.ep2a;

    // element ownership without integer divisions: thread -> row (tid & (rp - 1)) and every cstep-th column from tid / rp,
    // rp = the power of two that covers n (16, 32 or 64)
    const int rp = (n <= 16) ? 16 : ((n <= 32) ? 32 : 64), rsh = (n <= 16) ? 4 : ((n <= 32) ? 5 : 6);
    const int orow = tid & (rp - 1), ocol0 = tid >> rsh, cstep = NT >> rsh;
    const bool rok = orow < n;
    if (rok)
#pragma unroll
        for (int b = 0; b < NB; ++b)
            for (int col = ocol0; col < n; col += cstep)
                Gs[((size_t)b * n + col) * ldg + orow] = G[(long long)b * p.strideG + orow + (long long)col * p.ld];
    int accepted = 0;
    double neg_cnt = 0.0, neg_sum = 0.0, neg_min = INFINITY, neg_max = -INFINITY;   // thread 0 (one flavor: lanes of warp 0)

    for (int s = 0; s < p.nsteps; ++s) {
        const int l = p.slice0 + s * p.dir;              // 1-based slice of this step
        const int step = p.step0 + s;
        int8_t* cl = conf + (long long)(l - 1) * n;
        const double* utab = p.uniforms ? p.uniforms + (long long)chain * p.ustride + (long long)s * p.uf * n : nullptr;
        const long long toff = (long long)chain * p.tstride + (long long)s * n;
        __syncthreads();                                 // previous wrap done (Gs), previous step's su / sconf no longer read
        for (int i = tid; i < n; i += NT) {
            const int8_t x = cl[i];
            sconf[i] = x;
            su[i] = utab ? utab[i] : dqmc_uniform(p.seed, (uint64_t)(p.chain0 + chain), sweep_now, (uint32_t)step, (uint32_t)i);
            if (p.kind >= 2) {
                const double u2 = utab ? utab[n + i]
                                       : dqmc_uniform_choice(p.seed, (uint64_t)(p.chain0 + chain), sweep_now, (uint32_t)step, (uint32_t)i);
                sxnew[i] = (int8_t)dqmc_ghq_choice((int)x, u2);
            } else sxnew[i] = (int8_t)(-x);
        }
        __syncthreads();
        if constexpr (NB == 1) {
        // ---- sweep_spatial, one flavor block: G in REGISTERS (one 4 x 4 patch per thread, n <= 64) --------------------
        // Every lane of every warp tracks the diagonal entries of the sites lane and lane + 32 and evaluates all remaining
        // proposals at once against the current diagonal: up to the first accepted one these are exactly the sequential
        // decisions (a rejected proposal changes nothing), so (site, Delta / R) of an accept are known to all threads
        // without a broadcast and an accept costs ONE CTA barrier and 16 FMAs per thread (the shared-memory form above:
        // two barriers and 48 shared-memory accesses per thread).  The patches go back to Gs for the wrap.
        const int NPB = (n + 3) >> 2;
        const bool owner = tid < NPB * NPB;
        const int oxb = (tid % NPB) * 4, oyb = (tid / NPB) * 4;
        const int lane = tid & 31, warp = tid >> 5;
        double g[4][4];
#pragma unroll
        for (int iy = 0; iy < 4; ++iy)
#pragma unroll
            for (int ix = 0; ix < 4; ++ix) {
                const int x = oxb + ix, y = oyb + iy;
                g[ix][iy] = (owner && x < n && y < n) ? Gs[(size_t)y * ldg + x] : 0.0;
            }
        double gd[2], un[2];
        Proposal pr[2];
        int fc[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = lane + 32 * h;
            const bool in = j < n;
            pr[h] = make_proposal(p.kind, in ? (int)sconf[j] : 1, in ? (int)sxnew[j] : 1, p.ghq, em2a, ep2a);
            un[h] = in ? su[j] : 2.0;
            fc[h] = (in && p.forced) ? (int)p.forced[toff + j] : -1;
            gd[h] = in ? Gs[(size_t)j * ldg + j] : 0.0;
        }
        // The rank-1 update of an accept is applied at the TOP of the next iteration, in the same basic block as the next
        // evaluation: its shared-memory loads and 16 FMAs overlap the evaluation's dependent chain (reciprocal, decision,
        // ballot) instead of sitting alone behind the barrier.  Iteration 0 applies zeros.
        for (int e = tid; e < 256; e += NT) xch[e] = 0.0;
        __syncthreads();
        const int lox = owner ? oxb : 0, loy = owner ? oyb : 0;   // (non-owners update a patch nobody stores)
        int k = 0, next = 0;
        for (;;) {
            {
                const double* pcv = xch + ((k + 1) & 1) * 64;
                const double* prv = xch + 128 + ((k + 1) & 1) * 64;
                const double2 c01 = *reinterpret_cast<const double2*>(pcv + lox), c23 = *reinterpret_cast<const double2*>(pcv + lox + 2);
                const double2 r01 = *reinterpret_cast<const double2*>(prv + loy), r23 = *reinterpret_cast<const double2*>(prv + loy + 2);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int jj = lane + 32 * h;
                    if (jj < n) gd[h] = fma(pcv[jj], prv[jj], gd[h]);
                }
                const double cv[4] = {c01.x, c01.y, c23.x, c23.y}, rv[4] = {r01.x, r01.y, r23.x, r23.y};
#pragma unroll
                for (int iy = 0; iy < 4; ++iy)
#pragma unroll
                    for (int ix = 0; ix < 4; ++ix) g[ix][iy] = fma(cv[ix], rv[iy], g[ix][iy]);
            }
            // the half of the sites that holds `next` first; the other half only if nothing was accepted there (uniform
            // branches: the evaluation of a half is ~45 of the ~270 instructions of an accept)
            double prob[2] = {0.0, 0.0}, cf[2] = {0.0, 0.0};
            unsigned b0 = 0u, b1 = 0u;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if ((h == 0) ? (next < 32) : (b0 == 0u)) {
                    const int j = lane + 32 * h;
                    const bool elig = (j >= next) && (j < n);
                    const double Rv = 1.0 + pr[h].Dl[0] * (1.0 - gd[h]);
                    cf[h] = pr[h].Dl[0] * ss_rcp(Rv);                    // Delta / R, speculative
                    prob[h] = proposal_prob(p.kind, pr[h], Rv * Rv);
                    const int a_ = (fc[h] >= 0) ? (fc[h] != 0) : ((prob[h] > 1.0) || (un[h] < prob[h]));
                    const unsigned bb = __ballot_sync(0xffffffffu, elig ? a_ : 0);
                    if (h == 0) b0 = bb; else b1 = bb;
                }
            }
            const int jacc = b0 ? (__ffs(b0) - 1) : (b1 ? 32 + __ffs(b1) - 1 : -1);
            if (warp == 0) {                                 // traces and statistics of the real decisions
                const int jlast = (jacc >= 0) ? jacc : n - 1;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = lane + 32 * h;
                    if (j >= next && j <= jlast) {
                        if (p.check_sign && prob[h] < 0.0) {
                            neg_cnt += 1.0; neg_sum += log10(fabs(prob[h]));
                            neg_min = fmin(neg_min, prob[h]); neg_max = fmax(neg_max, prob[h]);
                        }
                        if (p.probs) p.probs[toff + j] = prob[h];
                        if (p.decisions) p.decisions[toff + j] = (unsigned char)(j == jacc);
                    }
                }
            }
            if (jacc < 0) break;
            const int src = jacc & 31, hh = jacc >> 5;
            const double c0 = __shfl_sync(0xffffffffu, hh ? cf[1] : cf[0], src);
            const int j = jacc;
            if (tid == src) { sconf[j] = sxnew[j]; cl[j] = sxnew[j]; }
            double* cvb = xch + (k & 1) * 64;                // G[:, j] - e_j   (the reference's -(e_j - G[:, j]))
            double* rvb = xch + 128 + (k & 1) * 64;          // (Delta / R) G[j, :]
            if (owner) {
                // column j & 3 / row j & 3 of the 4 x 4 patch (patches are aligned to multiples of 4: the index inside the patch
                // is uniform over the CTA -- one uniform switch instead of 32 selects on the register patch)
                double pc[4], prw[4];
                switch (j & 3) {
                case 0:
#pragma unroll
                    for (int q = 0; q < 4; ++q) { pc[q] = g[q][0]; prw[q] = g[0][q]; }
                    break;
                case 1:
#pragma unroll
                    for (int q = 0; q < 4; ++q) { pc[q] = g[q][1]; prw[q] = g[1][q]; }
                    break;
                case 2:
#pragma unroll
                    for (int q = 0; q < 4; ++q) { pc[q] = g[q][2]; prw[q] = g[2][q]; }
                    break;
                default:
#pragma unroll
                    for (int q = 0; q < 4; ++q) { pc[q] = g[q][3]; prw[q] = g[3][q]; }
                    break;
                }
                // (entries >= n of a patch are zero and stay zero: whole patches are exchanged, no bounds tests)
                if (j >= oyb && j < oyb + 4) {               // this patch holds part of column j
#pragma unroll
                    for (int ix = 0; ix < 4; ++ix) pc[ix] -= (oxb + ix == j) ? 1.0 : 0.0;
                    *reinterpret_cast<double2*>(cvb + oxb) = make_double2(pc[0], pc[1]);
                    *reinterpret_cast<double2*>(cvb + oxb + 2) = make_double2(pc[2], pc[3]);
                }
                if (j >= oxb && j < oxb + 4) {               // ... part of row j
                    *reinterpret_cast<double2*>(rvb + oyb) = make_double2(c0 * prw[0], c0 * prw[1]);
                    *reinterpret_cast<double2*>(rvb + oyb + 2) = make_double2(c0 * prw[2], c0 * prw[3]);
                }
            }
            __syncthreads();
            ++k; next = j + 1;
        }
        accepted += k;
        if (owner)
#pragma unroll
            for (int iy = 0; iy < 4; ++iy)
#pragma unroll
                for (int ix = 0; ix < 4; ++ix) {
                    const int x = oxb + ix, y = oyb + iy;
                    if (x < n && y < n) Gs[(size_t)y * ldg + x] = g[ix][iy];
                }
        __syncthreads();
        } else {
        // ---- sweep_spatial (local_updates.jl:23-60): decisions taken redundantly by every thread --------------
        for (int i = 0; i < n; ++i) {
            const Proposal pr = make_proposal(p.kind, (int)sconf[i], (int)sxnew[i], p.ghq, em2a, ep2a);
            double Rv[2];
#pragma unroll
            for (int b = 0; b < NB; ++b) Rv[b] = 1.0 + pr.Dl[b] * (1.0 - Gs[((size_t)b * n + i) * ldg + i]);
            const double prob = proposal_prob(p.kind, pr, (NB == 1) ? Rv[0] * Rv[0] : Rv[0] * Rv[1]);
            int acc;
            if (p.forced) acc = p.forced[toff + i] != 0;
            else acc = (prob > 1.0) || (su[i] < prob);
            if (tid == 0) {
                if (p.check_sign && prob < 0.0) {
                    neg_cnt += 1.0; neg_sum += log10(fabs(prob));
                    neg_min = fmin(neg_min, prob); neg_max = fmax(neg_max, prob);
                }
                if (p.probs) p.probs[toff + i] = prob;
                if (p.decisions) p.decisions[toff + i] = (unsigned char)acc;
            }
            if (acc) {                                   // uniform over the CTA
                // update_greens! (fields.jl:271-286): IG = e_i - G[:, i]; g = (Delta / R) G[i, :]; G -= IG g^T
                if (tid < n) {
                    const int k = tid;
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        const double* Gb = Gs + (size_t)b * n * ldg;
                        IG[b * n + k] = ((k == i) ? 1.0 : 0.0) - Gb[(size_t)i * ldg + k];
                        gr[b * n + k] = (pr.Dl[b] * ss_rcp(Rv[b])) * Gb[(size_t)k * ldg + i];
                    }
                }
                __syncthreads();
                if (rok)
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        const double mig = -IG[b * n + orow];
                        double* gcol = Gs + (size_t)b * n * ldg + orow;
                        const double* grb = gr + b * n;
                        // (measured at cfg 2: this plain loop 8171 sweeps/s; loads batched + 16-way unrolled 7684; pointer
                        //  increments + unroll 4: 7031)
                        for (int col = ocol0; col < n; col += cstep) gcol[(size_t)col * ldg] = fma(mig, grb[col], gcol[(size_t)col * ldg]);
                    }
                if (tid == 0) { sconf[i] = sxnew[i]; cl[i] = sxnew[i]; }
                ++accepted;
                __syncthreads();
            }
        }
        }
        // ---- wrap_greens! to the next slice (stack.jl:594-603) -------------------------------------------------------
        const int lw = (p.dir == 1) ? l : l - 1;         // slice whose B matrix wraps: B_l going up, B_{l-1} going down
        const int8_t* cw = conf + (long long)(lw - 1) * n;
        if (tid < n) {
            const int k = tid;
            const int x = (p.dir == 1) ? (int)sconf[k] : (int)cw[k];
            const int code = (p.kind >= 2) ? ((x - 1) & 3) : ((x > 0) ? 0 : 1);
#pragma unroll
            for (int b = 0; b < NB; ++b) { dpos[b * n + k] = p.lut[0][b][code]; dneg[b * n + k] = p.lut[1][b][code]; }
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            double* Gb = Gs + (size_t)b * n * ldg;
            if (p.dir == 1) {
                // up: tmp = eT2 (e^V G);  G = tmp e^-V eT2^-1          (multiply_slice_matrix_left!, ..._inv_right!)
                ss_mm(Ts, ldg, p.eT2, p.ld, Gb, ldg, n, dpos + b * n, nullptr, nullptr);
                __syncthreads();
                ss_mm(Gb, ldg, Ts, ldg, p.eT2i, p.ld, n, dneg + b * n, nullptr, nullptr);
            } else {
                // down: tmp = e^-V eT2^-1 G;  G = tmp eT2 e^V           (multiply_slice_matrix_inv_left!, ..._right!)
                ss_mm(Ts, ldg, p.eT2i, p.ld, Gb, ldg, n, nullptr, dneg + b * n, nullptr);
                __syncthreads();
                ss_mm(Gb, ldg, Ts, ldg, p.eT2, p.ld, n, nullptr, nullptr, dpos + b * n);
            }
            __syncthreads();
        }
    }
    if (rok)
#pragma unroll
        for (int b = 0; b < NB; ++b)
            for (int col = ocol0; col < n; col += cstep)
                G[(long long)b * p.strideG + orow + (long long)col * p.ld] = Gs[((size_t)b * n + col) * ldg + orow];
    if (NB == 1 && tid < 32) {                           // the lanes of warp 0 hold the statistics of their own sites
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
            double* st = p.stats + (long long)chain * 4;
            st[0] += neg_cnt; st[1] += neg_sum; st[2] = fmin(st[2], neg_min); st[3] = fmax(st[3], neg_max);
        }
    }
}

// ---- a run of slice-matrix products on one operand (small lattices) ---------------------------------------------
// add_slice_sequence_left / _right (stack.jl:377-416) and the builds of the unequal-time stack multiply one matrix by
// the slice matrices of a whole range, one GEMM launch per slice; for n <= 64 a launch is a few microseconds of latency
// for ~0.1 us of tensor work.  Here one CTA per matrix keeps the operand in shared memory and applies the `count` slice
// matrices back to back (same DMMA order over k as gemm.cu, so the products are bit-identical to the separate launches):
//   op 0  M <- eT2 e^{V_s} M               multiply_slice_matrix_left!          (slices first, first + 1, ...)
//   op 1  M <- e^{V_s} eT2^T M             multiply_daggered_slice_matrix_left! (slices first, first - 1, ...)
//   op 2  M <- e^{-V_s} eT2^-1 M           multiply_slice_matrix_inv_left!      (slices first, first - 1, ...)
__global__ void __launch_bounds__(256)
slice_chain_kernel(const SliceChainParams p)
{
    extern __shared__ __align__(16) double sm[];
    constexpr int NT = 256;                                  // the launches below (compile-time strides: no trip-count divisions)
    const int n = p.n, ldg = p.ldg, tid = threadIdx.x;
    const int mat = blockIdx.x, chain = mat / p.nb, blk = mat - chain * p.nb;
    double* Ma = sm;                                     // [n][ldg]
    double* Mb = Ma + (size_t)n * ldg;                   // [n][ldg]
    double* dv = Mb + (size_t)n * ldg;                   // [n] diagonal factor of the current slice
    const int rp = (n <= 16) ? 16 : ((n <= 32) ? 32 : 64), rsh = (n <= 16) ? 4 : ((n <= 32) ? 5 : 6);
    const int orow = tid & (rp - 1), ocol0 = tid >> rsh, cstep = NT >> rsh;
    const bool rok = orow < n;
    const int8_t* conf = p.conf + (long long)chain * p.cstride;
    if (p.src) {
        const double* S = p.src + (long long)mat * p.stride;
        if (rok) for (int col = ocol0; col < n; col += cstep) Ma[(size_t)col * ldg + orow] = S[orow + (long long)col * p.ld];
    } else {
        if (rok) for (int col = ocol0; col < n; col += cstep) Ma[(size_t)col * ldg + orow] = (col == orow) ? 1.0 : 0.0;
    }
    double* cur = Ma; double* oth = Mb;
    for (int i = 0; i < p.count; ++i) {
        const int s = (p.op == 0) ? p.first + i : p.first - i;      // 1-based slice
        __syncthreads();                                 // operand complete; dv of the previous slice no longer read
        if (tid < n) {
            const int x = conf[(long long)(s - 1) * n + tid];
            const int code = p.ghq ? ((x - 1) & 3) : ((x > 0) ? 0 : 1);
            dv[tid] = p.lut[blk][code];
        }
        __syncthreads();
        if (p.op == 0) ss_mm<false>(oth, ldg, p.E, p.ld, cur, ldg, n, dv, nullptr, nullptr);
        else if (p.op == 1) ss_mm<true>(oth, ldg, p.E, p.ld, cur, ldg, n, nullptr, dv, nullptr);
        else ss_mm<false>(oth, ldg, p.E, p.ld, cur, ldg, n, nullptr, dv, nullptr);
        double* t = cur; cur = oth; oth = t;
    }
    __syncthreads();
    double* D = p.dst + (long long)mat * p.stride;
    if (rok) for (int col = ocol0; col < n; col += cstep) D[orow + (long long)col * p.ld] = cur[(size_t)col * ldg + orow];
}

cudaError_t launch_slice_chain(SliceChainParams p, cudaStream_t st)
{
    if (p.n_mats <= 0) return cudaSuccess;
    p.ldg = ss_ld(p.n);
    const size_t smem = ((size_t)2 * p.n * p.ldg + p.n) * sizeof(double);
    static SmemAttr attr;
    cudaError_t e = attr.ensure(slice_chain_kernel, smem);
    if (e != cudaSuccess) return e;
    slice_chain_kernel<<<(unsigned)p.n_mats, 256, smem, st>>>(p);
    count_launch();
    return cudaGetLastError();
}

static size_t slice_steps_smem(int n, int nb)
{
    const int ldg = ss_ld(n);
    return ((size_t)(nb + 1) * n * ldg + 256 + (size_t)(4 + 2 * nb) * n + n) * sizeof(double) + 2 * (size_t)n + 16;
}

bool slice_steps_supported(int n, int nb) { return n <= 64 && slice_steps_smem(n, nb) <= 110 * 1024; }

cudaError_t launch_slice_steps(SliceStepParams p, cudaStream_t st)
{
    if (p.n_chains <= 0 || p.nsteps <= 0) return cudaSuccess;
    p.ldg = ss_ld(p.n);
    p.em2a = exp(-2.0 * p.alpha); p.ep2a = exp(2.0 * p.alpha);
    const size_t smem = slice_steps_smem(p.n, p.nb);
    static SmemAttr attr[2];
    cudaError_t e = (p.nb == 1) ? attr[0].ensure(slice_steps_kernel<1>, smem) : attr[1].ensure(slice_steps_kernel<2>, smem);
    if (e != cudaSuccess) return e;
    if (p.nb == 1) slice_steps_kernel<1><<<(unsigned)p.n_chains, 256, smem, st>>>(p);
    else slice_steps_kernel<2><<<(unsigned)p.n_chains, 256, smem, st>>>(p);
    count_launch();
    return cudaGetLastError();
}

}  // namespace dqmc
