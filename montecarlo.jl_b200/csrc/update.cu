// update.cu -- the local Hubbard-Stratonovich flip sweep over one time slice.
//
// Replaces `sweep_spatial` (reference src/flavors/DQMC/updates/local_updates.jl:23-60)
// with `propose_local` / `calculate_detratio!` (fields.jl:388-393, 440-449, 63-84) and
// `accept_local!` -> `update_greens!` (fields.jl:340-344, 271-286; linalg/updates.jl).
//
// The reference applies every accepted flip as an immediate rank-1 update
//     G <- G - (I - G)[:, i] * (Delta / R) * G[i, :]
// which is 16 n^2 bytes of memory traffic per accept.  Here one CTA owns one Markov
// chain (all its flavor blocks) and works through the sites in blocks of `kb`:
// the kb columns and rows of G that the block can touch are staged in shared
// memory, accepted flips are kept as delayed factors (u_a, w_a) that overwrite the
// staged columns/rows in place (compacted: the a-th accept goes to slot a <= j), the
// current diagonal element / column / row are reconstructed on the fly
//     G_cur[x, y] = G0[x, y] - sum_a u_a[x] w_a[y]
// and at the end of the block G is updated once with a rank-k DMMA GEMM
// (G -= U W).  Acceptance ratios are computed by warp 0 with a shuffle reduction.
// The arithmetic per accepted flip is identical to the reference's up to the
// association of the delayed sum.  Roofline: latency-bound proposals + an
// HBM/FP64-balanced flush (2 n^2 k flops over 16 n^2 bytes), see DESIGN.md.
#include "common.cuh"
#include "../../include/dqmc_rng.h"
#include <math.h>

namespace dqmc {

__device__ __forceinline__ void dmma884u(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(a), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
}

__device__ __forceinline__ double upd_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

static inline int update_ldu(int n) { return n + (((4 - n) % 16) + 16) % 16; }   // == 4 mod 16

int update_pick_kb(int n, int nb)
{
    const long long per_slot = (long long)nb * 2 * update_ldu(n) * 8;
    long long kb = (200LL * 1024 - 10LL * n) / per_slot;
    if (kb > 32) kb = 32;
    kb &= ~3LL;
    if (kb < 4) kb = 4;
    if (kb > n) kb = (n + 3) & ~3;
    return (int)kb;
}

struct UpdShared {
    int dec[2];
    double coef[2][2];
    double gdiag[2][32];     // current G_ii of the sites of the block, per flavor (kb <= 32)
};


__global__ void __launch_bounds__(256) update_kernel(const UpdateParams p, const int ldu, const double em2a, const double ep2a)
{
    extern __shared__ __align__(16) double sm[];
    const int n = p.n, nb = p.nb, kb = p.kb, ld = p.ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x, nwarps = NT >> 5;
    const int chain = blockIdx.x;

    double* Uc = sm;                                   // [nb][kb][ldu]   columns of G0 -> u_a
    double* Wr = Uc + (size_t)nb * kb * ldu;           // [nb][kb][ldu]   rows of G0    -> w_a
    UpdShared* sh = (UpdShared*)(Wr + (size_t)nb * kb * ldu);
    double* sunif = (double*)(sh + 1);                 // [n] Metropolis uniforms of this slice visit
    int8_t* sconf = (int8_t*)(sunif + n);              // [n]
    int8_t* sxnew = sconf + n;                         // [n] proposed value of every site (GHQ: drawn up front)

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
        if (p.kind >= 2) {                             // x_new = choices[x_old, rand(1:3)] (fields.jl:528, 590)
            const double u2 = utab ? utab[n + i]
                                   : dqmc_uniform_choice(p.seed, (uint64_t)(p.chain0 + chain), sweep_now, (uint32_t)p.step, (uint32_t)i);
            sxnew[i] = (int8_t)dqmc_ghq_choice((int)x, u2);
        } else sxnew[i] = (int8_t)(-x);
    }

    int accepted = 0;                  // tracked by every thread identically
    double neg_cnt = 0.0, neg_sum = 0.0, neg_min = INFINITY, neg_max = -INFINITY;   // lane 0 of warp 0

    for (int i0 = 0; i0 < n; i0 += kb) {
        const int kbc = (n - i0 < kb) ? (n - i0) : kb;
        __syncthreads();               // previous flush (global G) and sconf visible
        // ---- stage the kbc columns and rows of G ---------------------------------
        for (int b = 0; b < nb; ++b) {
            const double* Gb = G + (long long)b * p.strideG;
            double* ub = Uc + (size_t)b * kb * ldu;
            double* wb = Wr + (size_t)b * kb * ldu;
            for (int j = 0; j < kbc; ++j)                       // columns: contiguous in r
                for (int r = tid; r < n; r += NT) cp_async8(ub + (size_t)j * ldu + r, Gb + r + (long long)(i0 + j) * ld);
            if (lane < kbc)                                     // rows: a warp copies the kbc contiguous entries of one column
                for (int c = warp; c < n; c += nwarps) cp_async8(wb + (size_t)lane * ldu + c, Gb + (i0 + lane) + (long long)c * ld);
        }
        cp_async_wait_all();
        __syncthreads();
        if (tid < nb * kbc) {                                   // G_ii of the block's sites
            const int b = tid / kbc, x = tid - b * kbc;
            sh->gdiag[b][x] = Uc[((size_t)b * kb + x) * ldu + i0 + x];
        }
        __syncthreads();

        int k = 0;                     // accepted flips in this block (delayed factors in slots 0..k-1)
        for (int j = 0; j < kbc; ++j) {
            const int i = i0 + j;
            // ---- decision: warp 0 ---------------------------------------------------
            if (warp == 0) {
                const Proposal pr = make_proposal(p.kind, (int)sconf[i], (int)sxnew[i], p.ghq, em2a, ep2a);
                double Rv[2], Dl[2];
                for (int b = 0; b < nb; ++b) {
                    // current G_ii = G0_ii - sum_a u_a[i] w_a[i], kept as a running value per site of the block
                    const double gii = sh->gdiag[b][j];
                    Dl[b] = pr.Dl[b];
                    Rv[b] = 1.0 + Dl[b] * (1.0 - gii);
                }
                const double prob = proposal_prob(p.kind, pr, (nb == 1) ? Rv[0] * Rv[0] : Rv[0] * Rv[1]);
                __syncwarp();                                   // every lane has read sconf[i] before lane 0 may overwrite it
                if (lane == 0) {
                    if (p.check_sign && prob < 0.0) {
                        neg_cnt += 1.0; neg_sum += log10(fabs(prob));
                        neg_min = fmin(neg_min, prob); neg_max = fmax(neg_max, prob);
                    }
                    int acc;
                    if (forced) acc = forced[i] != 0;
                    else if (prob > 1.0) acc = 1;
                    else {
                        acc = sunif[i] < prob;
                    }
                    if (p.probs) p.probs[(long long)chain * p.tstride + i] = prob;
                    if (p.decisions) p.decisions[(long long)chain * p.tstride + i] = (unsigned char)acc;
                    sh->dec[j & 1] = acc;
                    if (acc) {
                        // Delta / R (vldiv22!, fields.jl:176-216) via a Newton reciprocal: the library division is
                        // a ~15-deep dependent FP64 chain on the serial path of every accepted flip
                        for (int b = 0; b < nb; ++b) sh->coef[j & 1][b] = Dl[b] * upd_rcp(Rv[b]);
                        sconf[i] = sxnew[i]; conf[i] = sconf[i];
                    }
                }
            }
            __syncthreads();
            const int acc = sh->dec[j & 1];
            if (acc) {
                // ---- new delayed factors (fields.jl:271-286) ------------------------------
                double* ub0 = Uc; double* wb0 = Wr;
                double* ub1 = Uc + (size_t)(nb - 1) * kb * ldu; double* wb1 = Wr + (size_t)(nb - 1) * kb * ldu;
                const double coef0 = sh->coef[j & 1][0], coef1 = sh->coef[j & 1][nb - 1];
                for (int r = tid; r < n; r += NT) {
                    // both flavor blocks in one loop body: twice the independent FMA chains per iteration
                    double col0 = ub0[(size_t)j * ldu + r], row0 = wb0[(size_t)j * ldu + r];
                    double col1 = ub1[(size_t)j * ldu + r], row1 = wb1[(size_t)j * ldu + r];
#pragma unroll 2
                    for (int a = 0; a < k; ++a) {
                        col0 = fma(ub0[(size_t)a * ldu + r], wb0[(size_t)a * ldu + i], col0);
                        row0 = fma(ub0[(size_t)a * ldu + i], wb0[(size_t)a * ldu + r], row0);
                        if (nb == 2) {
                            col1 = fma(ub1[(size_t)a * ldu + r], wb1[(size_t)a * ldu + i], col1);
                            row1 = fma(ub1[(size_t)a * ldu + i], wb1[(size_t)a * ldu + r], row1);
                        }
                    }
                    // element (slot k, r) is only ever touched by this thread until the barrier
                    const double un0 = col0 - ((r == i) ? 1.0 : 0.0), wn0 = coef0 * row0;      // -u = G[:, i] - e_i
                    ub0[(size_t)k * ldu + r] = un0; wb0[(size_t)k * ldu + r] = wn0;
                    const bool inblk = (r >= i0) && (r < i0 + kbc);
                    if (inblk) sh->gdiag[0][r - i0] += un0 * wn0;
                    if (nb == 2) {
                        const double un1 = col1 - ((r == i) ? 1.0 : 0.0), wn1 = coef1 * row1;
                        ub1[(size_t)k * ldu + r] = un1; wb1[(size_t)k * ldu + r] = wn1;
                        if (inblk) sh->gdiag[1][r - i0] += un1 * wn1;
                    }
                }
                // the loop above reads slot-a entries at index i written by other threads in
                // earlier steps (already separated by barriers) and slot j/k entries of its own r.
                // BUT when k < j another thread's read of ub[k][i] (a < k only) never hits slot k. ok
                ++k; ++accepted;
                __syncthreads();
            }
        }
        // ---- flush: G_b -= sum_{a<k} u_a w_a^T  (rank-k DMMA update) ---------------------
        if (k > 0) {
            const int g = lane >> 2, t = lane & 3;
            const int tiles = (n + 31) / 32;
            const int k4 = (k + 3) / 4;
            {
                // each warp walks its tiles with a one-tile look-ahead: the loads of G for tile t+1 are in
                // flight while tile t is updated and stored (the flush is bound by memory-level parallelism)
                const int ntl = nb * tiles * tiles;
                auto tile_ptr = [&](int bt, int& tm, int& tn, int& b) -> double* {
                    b = bt / (tiles * tiles);
                    const int tile = bt - b * (tiles * tiles);
                    tm = (tile % tiles) * 32; tn = (tile / tiles) * 32;
                    return G + (long long)b * p.strideG;
                };
                auto load_tile = [&](int bt, double (&dst)[4][4][2]) {
                    int tm, tn, b;
                    const double* Gb = tile_ptr(bt, tm, tn, b);
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        const int r = tm + mi * 8 + g;
#pragma unroll
                        for (int nj = 0; nj < 4; ++nj)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int c = tn + nj * 8 + 2 * t + e;
                                dst[mi][nj][e] = (r < n && c < n) ? Gb[r + (long long)c * ld] : 0.0;
                            }
                    }
                };
                double cur[4][4][2], nxt[4][4][2];
                if (warp < ntl) load_tile(warp, cur);
                for (int bt = warp; bt < ntl; bt += nwarps) {
                    const bool more = bt + nwarps < ntl;
                    if (more) load_tile(bt + nwarps, nxt);
                    int tm, tn, b;
                    double* Gb = tile_ptr(bt, tm, tn, b);
                    const double* ub = Uc + (size_t)b * kb * ldu;
                    const double* wb = Wr + (size_t)b * kb * ldu;
                    for (int kk = 0; kk < k4; ++kk) {
                        const int a = kk * 4 + t;
                        const bool live = a < k;
                        double af[4], bf[4];
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi) {
                            const int r = tm + mi * 8 + g;
                            af[mi] = (live && r < n) ? ub[(size_t)a * ldu + r] : 0.0;
                        }
#pragma unroll
                        for (int nj = 0; nj < 4; ++nj) {
                            const int c = tn + nj * 8 + g;
                            bf[nj] = (live && c < n) ? wb[(size_t)a * ldu + c] : 0.0;
                        }
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                            for (int nj = 0; nj < 4; ++nj)
                                dmma884u(cur[mi][nj][0], cur[mi][nj][1], af[mi], bf[nj]);
                    }
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        const int r = tm + mi * 8 + g;
#pragma unroll
                        for (int nj = 0; nj < 4; ++nj)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int c = tn + nj * 8 + 2 * t + e;
                                if (r < n && c < n) Gb[r + (long long)c * ld] = cur[mi][nj][e];
                                cur[mi][nj][e] = nxt[mi][nj][e];
                            }
                    }
                }
            }
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

cudaError_t launch_update(const UpdateParams& p, cudaStream_t st)
{
    if (p.n_chains <= 0) return cudaSuccess;
    const int ldu = update_ldu(p.n);
    int nt = ((p.n + 31) / 32) * 32;             // one thread per row (flavor blocks in turn)
    if (nt < 64) nt = 64;
    if (nt > 256) nt = 256;
    const size_t smem = (size_t)p.nb * 2 * p.kb * ldu * sizeof(double) + sizeof(UpdShared) + (size_t)p.n * 10 + 16;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    static SmemAttr attr;
    cudaError_t e = attr.ensure(update_kernel, smem);
    if (e != cudaSuccess) return e;
    const double em2a = exp(-2.0 * p.alpha), ep2a = exp(2.0 * p.alpha);
    update_kernel<<<(unsigned)p.n_chains, nt, smem, st>>>(p, ldu, em2a, ep2a);
    count_launch();
    return cudaGetLastError();
}

}  // namespace dqmc
