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
            af[i] = (kok && r < n) ? A[r + (size_t)k * lda] : 0.0;
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
    const int n = p.n, ldg = p.ldg, tid = threadIdx.x, NT = blockDim.x;
    const int chain = blockIdx.x;
    double* Gs = sm;                                     // [NB][n][ldg]
    double* Ts = Gs + (size_t)NB * n * ldg;              // [n][ldg] wrap intermediate (one flavor at a time)
    double* IG = Ts + (size_t)n * ldg;                   // [NB][n]  e_i - G[:, i]
    double* gr = IG + NB * n;                            // [NB][n]  (Delta / R) G[i, :]
    double* dpos = gr + NB * n;                          // [NB][n]  e^{+V} of the wrap slice
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
    double neg_cnt = 0.0, neg_sum = 0.0, neg_min = INFINITY, neg_max = -INFINITY;   // thread 0

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
    if (tid == 0) {
        if (p.accepted) p.accepted[chain] += accepted;
        if (p.stats && neg_cnt > 0.0) {
            double* st = p.stats + (long long)chain * 4;
            st[0] += neg_cnt; st[1] += neg_sum; st[2] = fmin(st[2], neg_min); st[3] = fmax(st[3], neg_max);
        }
    }
}

static size_t slice_steps_smem(int n, int nb)
{
    const int ldg = ss_ld(n);
    return ((size_t)(nb + 1) * n * ldg + (size_t)4 * nb * n + n) * sizeof(double) + 2 * (size_t)n + 16;
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
