// udt_reg.cu -- multi-level batched column-pivoted Householder QR -> UDT: level schedule, explicit Q, host side.
//
// Same mathematics and outputs as the reference's udt_AVX_pivot! (src/flavors/DQMC/linalg/UDT.jl:216-334:
// indmaxcolumn :175-192, reflector! :157-172, reflectorApply! :53-70, Q accumulation :272-288,
// D = |diag R| with 0 -> 1 :293-301, T = D^-1 R P^T or the unpivoted upper triangle :311-334).
//
// The Householder steps run register-resident on thread-block clusters (udt_steps.cu).
//
// Levels.  A step costs a few microseconds of latency (barrier, selection, shuffle trees) whatever the size of the
// trailing matrix, and a 512 KB matrix needs 4 SMs, so only 33 of the 296 matrices of a cfg-4 launch
// are in flight.  The factorisation is therefore cut into levels: level 0 does the first steps on the
// cluster, writes the rows of R it has finished and exports the (compacted) trailing block; the next level
// factors that smaller block -- which fits fewer SMs, so more matrices are in flight -- and so on down to 64 columns.
// Every level recomputes the column norms from scratch (as the reference does at every step), so the arithmetic
// is unchanged.
//
// Q.  Formed by a separate full-grid kernel, backwards (UDT.jl:272-288), four reflectors at a time in compact WY form.
#include <algorithm>

#include "udt_level.cuh"

namespace dqmc {

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

// CPW = columns of Q per warp, NW = warps per CTA.  The build uses CPW = 4, NW = 8: half the accumulators of the round-1 layout
// (8 columns per warp, 224 registers, one CTA per SM), so two CTAs share an SM and four warps per scheduler hide the shuffle
// trees and the shared-memory reads of the reflectors.
template <int RPL, int CPW, int NW>
__global__ void __launch_bounds__(32 * NW, (CPW == 4 && NW == 8) ? 2 : 1)
udt_formq4_kernel(const UdtParams p, const double* __restrict__ T4, int ngroups)
{
    constexpr int NV = RPL * 32;
    constexpr int CPC = NW * CPW;                        // columns per CTA
    const int n = p.n, ld = p.ld, ldv = p.ldv;
    const int ctas_per_mat = (n + CPC - 1) / CPC;
    const int mat = blockIdx.x / ctas_per_mat, part_i = blockIdx.x - mat * ctas_per_mat;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col0 = part_i * CPC + warp * CPW;          // this warp owns columns col0 .. col0 + CPW - 1
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

    double a[CPW][RPL];
#pragma unroll
    for (int c = 0; c < CPW; ++c)
#pragma unroll
        for (int r = 0; r < RPL; ++r) a[c][r] = (col0 + c < n && lane + 32 * r == col0 + c) ? 1.0 : 0.0;

    const int ctop = min(n, part_i * CPC + CPC) - 1;     // highest column of this CTA
    const int gtop = ctop >> 2;
    const int wtop = (col0 < n) ? (min(n - 1, col0 + CPW - 1) >> 2) : -1;   // highest block that touches this warp
    const bool dense = (ldv == NV);
    auto stage = [&](int cidx, int s) {                  // blocks CH cidx .. CH cidx + CH - 1 -> stage s
        for (int e = tid; e < CH * 4 * (NV / 2); e += 32 * NW) {
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
        // ---- W = V^T A: 4 CPW independent chains ------------------------------------------------
        double d[4][CPW];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < CPW; ++c) d[j][c] = 0.0;
#pragma unroll
        for (int r = 0; r < RPL; ++r)
            if (r >= r0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double v = VS(s, gb * 4 + j, lane + 32 * r);
#pragma unroll
                    for (int c = 0; c < CPW; ++c) d[j][c] = fma(v, a[c][r], d[j][c]);
                }
            }
        // ---- one reduction tree for the whole block: halve over the columns, butterfly over the rest --
        double w[4];
        if constexpr (CPW == 8) {
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
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double send = h4 ? y2[j][0] : y2[j][1];
                const double keep = h4 ? y2[j][1] : y2[j][0];
                w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
        } else {
            double y1[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const double send = h16 ? d[j][c] : d[j][c + 2];
                    const double keep = h16 ? d[j][c + 2] : d[j][c];
                    y1[j][c] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double send = h8 ? y1[j][0] : y1[j][1];
                const double keep = h8 ? y1[j][1] : y1[j][0];
                w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] += __shfl_xor_sync(0xffffffffu, w[j], 4);
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
        for (int half = 0; half < CPW / 4; ++half) {
            double yy[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int cc = half * 4 + c;
                const int src = (CPW == 8) ? (((cc & 4) ? 16 : 0) | ((cc & 2) ? 8 : 0) | ((cc & 1) ? 4 : 0))
                                           : (((cc & 2) ? 16 : 0) | ((cc & 1) ? 8 : 0));
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
    for (int c = 0; c < CPW; ++c) {
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
template <int RPL, int CPW, int NW>
static cudaError_t launch_formq4(const UdtParams& p, double* T4, cudaStream_t st)
{
    const int ngroups = (p.n + 3) / 4;
    const int warps = p.batch * ngroups;
    count_launch();
    udt_wy_t_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, st>>>(p, T4, ngroups);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int ctas = p.batch * ((p.n + NW * CPW - 1) / (NW * CPW));
    count_launch();
    constexpr int smem = (2 * 4 * 4 * RPL * 32 + 2 * 4 * 16) * (int)sizeof(double);
    // (Q held in the DMMA accumulator layout with the update as one DMMA per 8 rows was measured SLOWER: 1.38 vs
    //  0.96 ms per 296 x 256^2 -- see DESIGN.md; it is not part of the build)
    static SmemAttr attr;
    e = attr.ensure(udt_formq4_kernel<RPL, CPW, NW>, smem);
    if (e != cudaSuccess) return e;
    udt_formq4_kernel<RPL, CPW, NW><<<(unsigned)ctas, 32 * NW, smem, st>>>(p, T4, ngroups);
    return cudaGetLastError();
}

#define DQMC_RPL_SWITCH(rpl, CALL) \
    switch (rpl) { \
    case 1: { constexpr int R = 1; err = CALL; } break; case 2: { constexpr int R = 2; err = CALL; } break; \
    case 3: { constexpr int R = 3; err = CALL; } break; case 4: { constexpr int R = 4; err = CALL; } break; \
    case 5: { constexpr int R = 5; err = CALL; } break; case 6: { constexpr int R = 6; err = CALL; } break; \
    case 7: { constexpr int R = 7; err = CALL; } break; case 8: { constexpr int R = 8; err = CALL; } break; \
    default: { constexpr int R = 9; err = CALL; } break; }

bool udt_reg_supported(int n) { UdtLevel g{}; return n <= 288 && udt_steps_geometry(n, g); }

// Level sizes are the sizes at which the geometry gets cheaper: <= 256 needs 4 SMs per matrix (37 matrices
// in flight), <= 192 two (74: 12 warps of 8 columns x 6 register rows), <= 128 one (148), <= 96 half (296).  Returns the size
// of the trailing block the level that starts with nk columns hands on (0: it finishes the factorisation).
static int udt_next_level_size(int nk)
{
    if (nk <= 64) return 0;
    // (measured, 296 x 256^2: a 224-column level on clusters of 3 -- 48 in flight, 7 waves, 3.4 us per step with 10 fat
    //  warps -- 0.76 ms for its 32 steps against 0.75 ms on the clusters of 4: no gain; 96 columns fit two matrices per SM:
    //  128 -> 64 in 0.27 instead of 0.32 ms)
    // (a 224-column HYBRID level on clusters of 2 between 256 and 192: 0.70 + 0.57 ms against 1.20 ms for the 64 steps on the
    //  256-column level alone -- a level's load / export of the 512 KB matrix costs ~0.1 ms, more than the cheaper steps save)
    static const int sizes[5] = {256, 192, 128, 96, 64};
    for (int k = 0; k < 5; ++k)
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
        if (!udt_steps_geometry(nk, g)) return cudaErrorInvalidConfiguration;
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
        err = launch_udt_steps(p, g, st);
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
    // (Measured alternative, not kept: compact WY with 32-reflector blocks on the batched DMMA GEMM -- T_b and V_b T_b
    //  from one CTA per block, then W = V_b^T Q, Q -= (V_b T_b) W per block: 1.29 ms per 296 x 256^2 against 0.96 ms here;
    //  the K = 32 updates are bound by re-reading Q, the block preparation by shared-memory traffic.)
    // (Also measured and not kept: form-Q in the steps kernel's register tiling -- 3-round butterflies, but four times the
    //  shared-memory reads of the Householder vectors: 1.02 ms against 0.96 ms.)
    // (Measured, 296 x 256^2: 8 columns per warp, one CTA of 8 warps per SM 0.83 ms; 4 columns per warp on two CTAs per SM
    //  0.68 ms; 4 columns per warp on one CTA of 16 warps 0.74 ms; 2 columns per warp on three CTAs per SM 0.86 ms.)
    DQMC_RPL_SWITCH(rpl, (launch_formq4<R, 4, 8>(p, T4, st)))
    if (err != cudaSuccess) return err;
    // the Val(false) form wants the columns of D^-1 R in pivot (logical) order
    if (!direct_T)
        err = launch_permute_cols(Tphys_base, p.T, p.pivot, n, p.ld, strideTp, p.stridePivot, p.batch, st);
    return err;
}

}  // namespace dqmc
