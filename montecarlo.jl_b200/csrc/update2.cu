// update2.cu -- sweep_spatial with block-restricted proposals and an out-of-kernel rank-kb flush.
//
// Same mathematics and decisions as update.cu / the reference (`sweep_spatial`, local_updates.jl:23-60;
// `propose_local` / `calculate_detratio!`, fields.jl:388-393, 440-449, 63-84; `accept_local!` ->
// `update_greens!`, fields.jl:340-344, 271-286), reorganised so that the strictly serial part of a block of
// kb sites touches kb x kb numbers instead of 2 kb n:
//
//   1. serial phase   Only the restriction G[I, I] of the Green's function to the block's sites I lives in
//      shared memory.  A proposal reads its diagonal element; an accepted flip applies the reference's rank-1
//      update to that kb x kb matrix at once and records the block-restricted column / row of its factors
//      (Ub[:, a] = G[I, i_a] - e, Wb[a, :] = coef_a G[i_a, I]).  Warp 0 runs ahead over rejected proposals.
//   2. factor build   The full-length delayed factors follow from G0 and the recorded restrictions,
//          u_a = (G0[:, i_a] - e_{i_a}) + sum_{a' < a} u_{a'} Wb[a'][x_a],
//          w_a = coef_a (G0[i_a, :] + sum_{a' < a} Ub[x_a][a'] w_{a'}),
//      one thread per row / column with its own earlier entries kept in registers (fully unrolled over the
//      accepts): n k^2 flops per block, off the serial path, no barriers, nothing re-read from memory.
//   3. flush          G += U W^T is a separate batched DMMA GEMM over the whole grid (gemm.cu, K = kb).
//
// kb no longer has to fit 2 kb n doubles per flavor in shared memory (24 at n = 256): kb = 64 needs a third of
// the flush passes over G, which is what bounded update.cu (HBM traffic 3.3 GB -> 1.3 GB per slice visit at cfg 4).
#include "common.cuh"
#include "../../include/dqmc_rng.h"
#include <math.h>
#include <stdlib.h>

namespace dqmc {

__device__ __forceinline__ double upd2_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

struct Upd2Shared {
    int k;                 // accepted flips of this block so far
    int next;              // first proposal warp 0 has not decided yet
    int acc_site;          // block-local site of the accept being applied, -1: block finished
    double coef[2];
    int xs[64];            // block-local site of accept a
    double coefs[2][64];
};

// Full-length factors of one flavor from G0 and the block-restricted records, one thread per row (U) or column
// (W): acc <- G0 entries; earlier accepts enter through S[a'][a] (SU[a'][a] = Wb[a'][x_a] for U, SWt[a'][a] =
// Ub[x_a][a'] for W); the history of this thread's own factor entries stays in registers (static indexing:
// the loops over accepts are fully unrolled), so nothing is re-read from memory.
template <int KB, bool IS_W>
__device__ __forceinline__ void build_factor_row(const double* __restrict__ Gb, int ld, int r, int i0, int k,
                                                 const int* __restrict__ xs, const double* __restrict__ S,
                                                 const double* __restrict__ coefs, double* __restrict__ Fg, int ldf)
{
    constexpr int KP = KB + 2;
    double hist[KB];
    // all G0 entries of this row / column first: one exposed memory latency instead of one per tile
#pragma unroll
    for (int a = 0; a < KB; ++a) {
        hist[a] = 0.0;
        if (a < k) {
            const int ia = i0 + xs[a];
            hist[a] = IS_W ? Gb[ia + (long long)r * ld] : (Gb[r + (long long)ia * ld] - ((r == ia) ? 1.0 : 0.0));
        }
    }
#pragma unroll
    for (int t = 0; t < KB / 8; ++t) {
        const int a0 = 8 * t;
        double acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = hist[a0 + q];
        if (a0 < k) {                                       // uniform over the CTA
#pragma unroll
            for (int ap = 0; ap < a0; ++ap) {
                const double h = hist[ap];
                const double2* s2 = reinterpret_cast<const double2*>(S + (size_t)ap * KP + a0);
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
                for (int qp = 0; qp < q; ++qp) acc[q] = fma(acc[qp], S[(size_t)(a0 + qp) * KP + a0 + q], acc[q]);
                if (IS_W) acc[q] *= (a0 + q < k) ? coefs[a0 + q] : 0.0;
                else if (a0 + q >= k) acc[q] = 0.0;
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            hist[a0 + q] = acc[q];
            Fg[r + (long long)(a0 + q) * ldf] = acc[q];
        }
    }
}

// One CTA per chain, one block of kbc sites starting at i0.
template <int KB, int NB>
__global__ void __launch_bounds__(256)
update_block_kernel(const UpdateParams p, const int i0, const int kbc, const double em2a, const double ep2a)
{
    extern __shared__ __align__(16) double sm[];
    constexpr int kb = KB, KP = KB + 2, nb = NB;
    // G[I, I] lives in registers during the serial phase: one PX x PY patch (rows x, columns y) per thread
    constexpr int XT = (KB * KB * NB >= 256) ? (KB * KB * NB / 256) : 1;
    constexpr int PX = (XT >= 32) ? 8 : ((XT >= 8) ? 4 : ((XT >= 2) ? 2 : 1));
    constexpr int PY = XT / PX;
    constexpr int TPF = KB * KB / XT;                       // threads per flavor
    const int n = p.n, ld = p.ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x;
    const int chain = blockIdx.x;

    double* SU = sm;                                        // [nb][kb][KP]  factor-build coefficients (after the serial phase)
    double* Ub = SU + (size_t)nb * kb * KP;                 // [nb][kb (site x)][KP (accept a)]
    double* Wb = Ub + (size_t)nb * kb * KP;                 // [nb][kb (accept a)][KP (site y)]; later SWt
    double* colv = Wb + (size_t)nb * kb * KP;               // [nb][kb]
    double* rowv = colv + nb * kb;                          // [nb][kb]
    double* gdiag = rowv + nb * kb;                         // [nb][kb]  current G_ii of the block's sites
    double* sunif = gdiag + nb * kb;                        // [kb]
    Upd2Shared* sh = (Upd2Shared*)(sunif + kb);
    int8_t* sconf = (int8_t*)(sh + 1);                      // [kb]

    double* G = p.G + (long long)chain * nb * p.strideG;
    int8_t* conf = p.conf_slice + (long long)chain * p.cstride;
    const double* utab = p.uniforms ? p.uniforms + (long long)chain * p.ustride : nullptr;
    const unsigned char* forced = p.forced ? p.forced + (long long)chain * p.tstride : nullptr;

    // ---- load G[I, I] (registers), its diagonal, conf and uniforms of the block -----------------------
    const bool owner = tid < NB * TPF;
    const int ob = tid / TPF, ot = tid % TPF;
    const int oxb = (ot % (KB / PX)) * PX, oyb = (ot / (KB / PX)) * PY;   // patch origin of this thread in flavor ob
    double g[PX][PY];
#pragma unroll
    for (int iy = 0; iy < PY; ++iy)
#pragma unroll
        for (int ix = 0; ix < PX; ++ix) {
            const int x = oxb + ix, y = oyb + iy;
            g[ix][iy] = (owner && x < kbc && y < kbc) ? G[(long long)ob * p.strideG + (i0 + x) + (long long)(i0 + y) * ld] : 0.0;
        }
    for (int e = tid; e < nb * KB; e += NT) {
        const int b = e / KB, x = e % KB;
        gdiag[e] = (x < kbc) ? G[(long long)b * p.strideG + (i0 + x) + (long long)(i0 + x) * ld] : 0.0;
        colv[e] = 0.0; rowv[e] = 0.0;
    }
    for (int x = tid; x < kbc; x += NT) {
        sconf[x] = conf[i0 + x];
        sunif[x] = utab ? utab[i0 + x]
                        : dqmc_uniform(p.seed, (uint64_t)(p.chain0 + chain), (uint64_t)p.sweep, (uint32_t)p.step, (uint32_t)(i0 + x));
    }
    if (tid == 0) { sh->k = 0; sh->next = 0; sh->acc_site = -1; }
    __syncthreads();

    double neg_cnt = 0.0, neg_sum = 0.0, neg_min = INFINITY, neg_max = -INFINITY;   // per lane of warp 0

    // ---- serial phase ------------------------------------------------------------------------------------
    for (;;) {
        if (warp == 0) {
            // Every lane evaluates one of the next 32 proposals against the current G_ii.  Up to the first accepted
            // one they are exactly the sequential decisions (a rejected proposal changes nothing); the rest is
            // discarded and re-evaluated after the update.
            int found = -1;
            double c0 = 0.0, c1 = 0.0;
            for (int base = sh->next; base < kbc && found < 0; base += 32) {
                const int j = base + lane;
                int acc = 0;
                double prob = 0.0, Rv[2] = {1.0, 1.0}, Dl[2] = {0.0, 0.0};
                if (j < kbc) {
                    const double x = (double)sconf[j];
                    const double e_dE = (x > 0.0) ? em2a : ep2a;        // exp(dE), dE = -2 alpha x
                    const double e_mdE = (x > 0.0) ? ep2a : em2a;
#pragma unroll
                    for (int b = 0; b < nb; ++b) {
                        const double gii = gdiag[b * KB + j];
                        Dl[b] = ((p.kind == 1 && b == 1) ? e_mdE : e_dE) - 1.0;
                        Rv[b] = 1.0 + Dl[b] * (1.0 - gii);
                    }
                    if (p.kind == 0) prob = e_mdE * ((nb == 1) ? Rv[0] * Rv[0] : Rv[0] * Rv[1]);
                    else prob = Rv[0] * Rv[1];
                    if (forced) acc = forced[i0 + j] != 0;
                    else if (prob > 1.0) acc = 1;
                    else acc = sunif[j] < prob;
                }
                const unsigned ballot = __ballot_sync(0xffffffffu, acc);
                const int first = ballot ? (__ffs(ballot) - 1) : 32;          // lanes <= first are real decisions
                if (j < kbc && lane <= first) {
                    if (p.check_sign && prob < 0.0) {
                        neg_cnt += 1.0; neg_sum += log10(fabs(prob));
                        neg_min = fmin(neg_min, prob); neg_max = fmax(neg_max, prob);
                    }
                    if (p.probs) p.probs[(long long)chain * p.tstride + i0 + j] = prob;
                    if (p.decisions) p.decisions[(long long)chain * p.tstride + i0 + j] = (unsigned char)acc;
                }
                if (first < 32) {
                    found = base + first;
                    if (lane == first) {
                        c0 = Dl[0] * upd2_rcp(Rv[0]);                       // Delta / R (vldiv22!, fields.jl:176-216)
                        c1 = Dl[nb - 1] * upd2_rcp(Rv[nb - 1]);
                        sconf[j] = (int8_t)(-sconf[j]); conf[i0 + j] = sconf[j];
                        const int a = sh->k;
                        sh->acc_site = found; sh->coef[0] = c0; sh->coef[1] = c1; sh->next = found + 1;
                        sh->xs[a] = found; sh->coefs[0][a] = c0; sh->coefs[1][a] = c1;
                    }
                }
            }
            if (found < 0 && lane == 0) { sh->acc_site = -1; sh->next = kbc; }
        }
        __syncthreads();
        const int j = sh->acc_site;
        if (j < 0) break;
        const int a = sh->k;
        // restricted column / row of the new factors (fields.jl:271-286 on the block), from the owners' registers
        if (owner) {
            if (j >= oyb && j < oyb + PY) {                 // this patch holds part of column j
#pragma unroll
                for (int ix = 0; ix < PX; ++ix) {
                    double v = 0.0;
#pragma unroll
                    for (int iy = 0; iy < PY; ++iy) v = (oyb + iy == j) ? g[ix][iy] : v;
                    const int x = oxb + ix;
                    const double cv = v - ((x == j) ? 1.0 : 0.0);
                    colv[ob * KB + x] = cv;
                    Ub[((size_t)ob * kb + x) * KP + a] = cv;
                }
            }
            if (j >= oxb && j < oxb + PX) {                 // ... part of row j
#pragma unroll
                for (int iy = 0; iy < PY; ++iy) {
                    double v = 0.0;
#pragma unroll
                    for (int ix = 0; ix < PX; ++ix) v = (oxb + ix == j) ? g[ix][iy] : v;
                    const double rv = sh->coef[ob] * v;
                    rowv[ob * KB + oyb + iy] = rv;
                    Wb[((size_t)ob * kb + a) * KP + oyb + iy] = rv;
                }
            }
        }
        __syncthreads();
        if (owner) {
            double cv[PX], rv[PY];
#pragma unroll
            for (int ix = 0; ix < PX; ++ix) cv[ix] = colv[ob * KB + oxb + ix];
#pragma unroll
            for (int iy = 0; iy < PY; ++iy) rv[iy] = rowv[ob * KB + oyb + iy];
#pragma unroll
            for (int iy = 0; iy < PY; ++iy)
#pragma unroll
                for (int ix = 0; ix < PX; ++ix) g[ix][iy] = fma(cv[ix], rv[iy], g[ix][iy]);
        }
        if (warp == 0) {                                    // the diagonal the next decisions read
            for (int e = lane; e < nb * KB; e += 32) gdiag[e] = fma(colv[e], rowv[e], gdiag[e]);
            if (lane == 0) sh->k = a + 1;
            __syncwarp();
        }
    }
    const int k = sh->k;
    __syncthreads();

    // ---- factor build ------------------------------------------------------------------------------------
    // SU[a'][a] = Wb[a'][x_a] over Gblk, then SWt[a'][a] = Ub[x_a][a'] over Wb
    double* SWt = Wb;
    for (int e = tid; e < nb * KB * KB; e += NT) {
        const int b = e / (KB * KB), r = e % (KB * KB);
        const int a = r % KB, ap = r / KB;
        if (a < k && ap < k) SU[((size_t)b * kb + ap) * KP + a] = Wb[((size_t)b * kb + ap) * KP + sh->xs[a]];
    }
    __syncthreads();
    for (int e = tid; e < nb * KB * KB; e += NT) {
        const int b = e / (KB * KB), r = e % (KB * KB);
        const int a = r % KB, ap = r / KB;
        if (a < k && ap < k) SWt[((size_t)b * kb + ap) * KP + a] = Ub[((size_t)b * kb + sh->xs[a]) * KP + ap];
    }
    __syncthreads();
    for (int b = 0; b < nb; ++b) {
        const double* Gb = G + (long long)b * p.strideG;
        double* Ug = p.Ufac + ((long long)chain * nb + b) * p.strideF;
        double* Wg = p.Wfac + ((long long)chain * nb + b) * p.strideF;
        for (int r = tid; r < n; r += NT) {
            build_factor_row<KB, false>(Gb, ld, r, i0, k, sh->xs, SU + (size_t)b * kb * KP, sh->coefs[b], Ug, p.ldf);
            build_factor_row<KB, true>(Gb, ld, r, i0, k, sh->xs, SWt + (size_t)b * kb * KP, sh->coefs[b], Wg, p.ldf);
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
        if (p.accepted) p.accepted[chain] += k;
        if (p.stats && neg_cnt > 0.0) {
            double* s = p.stats + (long long)chain * 4;
            s[0] += neg_cnt; s[1] += neg_sum; s[2] = fmin(s[2], neg_min); s[3] = fmax(s[3], neg_max);
        }
    }
}

int update2_pick_kb(int n, int nb)
{
    int kb = 64;                                            // power of two: the register layout of G[I, I]
    while (kb > 8 && kb / 2 >= n) kb /= 2;
    return kb;
}

static size_t update2_smem(int kb, int nb)
{
    return ((size_t)3 * nb * kb * (kb + 2) + 3 * nb * kb + kb) * sizeof(double) + sizeof(Upd2Shared) + kb + 16;
}

template <int KB, int NB>
static cudaError_t launch_block(const UpdateParams& p, int i0, int kbc, double em2a, double ep2a, cudaStream_t st)
{
    const size_t smem = update2_smem(KB, NB);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(update_block_kernel<KB, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    update_block_kernel<KB, NB><<<(unsigned)p.n_chains, 256, smem, st>>>(p, i0, kbc, em2a, ep2a);
    ++g_kernel_launches;
    return cudaGetLastError();
}

template <int KB>
static cudaError_t launch_block_nb(const UpdateParams& p, int i0, int kbc, double em2a, double ep2a, cudaStream_t st)
{
    return (p.nb == 1) ? launch_block<KB, 1>(p, i0, kbc, em2a, ep2a, st) : launch_block<KB, 2>(p, i0, kbc, em2a, ep2a, st);
}

cudaError_t launch_update2(const UpdateParams& p, cudaStream_t st)
{
    if (p.n_chains <= 0) return cudaSuccess;
    if (!p.Ufac || !p.Wfac || (p.kb != 8 && p.kb != 16 && p.kb != 32 && p.kb != 64) || p.nb < 1 || p.nb > 2) return cudaErrorInvalidValue;
    if (update2_smem(p.kb, p.nb) > 227 * 1024) return cudaErrorInvalidConfiguration;
    const double em2a = exp(-2.0 * p.alpha), ep2a = exp(2.0 * p.alpha);
    for (int i0 = 0; i0 < p.n; i0 += p.kb) {
        const int kbc = (p.n - i0 < p.kb) ? (p.n - i0) : p.kb;
        cudaError_t e;
        switch (p.kb) {
        case 8: e = launch_block_nb<8>(p, i0, kbc, em2a, ep2a, st); break;
        case 16: e = launch_block_nb<16>(p, i0, kbc, em2a, ep2a, st); break;
        case 32: e = launch_block_nb<32>(p, i0, kbc, em2a, ep2a, st); break;
        default: e = launch_block_nb<64>(p, i0, kbc, em2a, ep2a, st); break;
        }
        if (e != cudaSuccess) return e;
        // flush: G += U W^T for every chain and flavor block
        GemmParams g{};
        g.M = g.N = p.n; g.K = p.kb;
        g.A = p.Ufac; g.lda = p.ldf; g.strideA = p.strideF; g.transA = 0;
        g.B = p.Wfac; g.ldb = p.ldf; g.strideB = p.strideF; g.transB = 1;
        g.C = p.G; g.ldc = p.ld; g.strideC = p.strideG;
        g.alpha = 1.0; g.beta = 1.0;
        g.rs = no_scale(); g.ks = no_scale(); g.cs = no_scale();
        g.add_diag = nullptr; g.add_stride = 0; g.batch = p.n_chains * p.nb;
        e = launch_gemm(g, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace dqmc
