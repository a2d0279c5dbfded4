// udt.cu -- batched column-pivoted Householder QR -> UDT decomposition.
//
// Replaces `udt_AVX_pivot!` (reference src/flavors/DQMC/linalg/UDT.jl:216-334 with
// indmaxcolumn :175-192, reflector! :157-172, reflectorApply! :53-70) for a batch
// of independent n x n matrices:
//     input * diag(colscale) = U * diag(D) * T,   D = |diag R| (exact 0 -> 1),
//     T = D^-1 R P^T (pivot_applied) or the clean upper-triangular D^-1 R plus the
//     pivot vector (the Val(false) form consumed by rdivp!).
//
// B200 design: the matrix never leaves the chip during the factorisation.  The
// columns of one matrix are dealt cyclically over the CTAs of a thread-block
// cluster (1, 2, 4 or 8 CTAs, chosen so that n x n/CS doubles fit in 227 KB of
// shared memory) and stay resident there.  Columns are never physically swapped:
// the column chosen at step j keeps its place, which makes the un-pivoting of T
// free.  Per Householder step there is ONE cluster barrier: every CTA publishes
// its best remaining column (squared norm, index and the column tail) into the
// distributed shared memory of all peers, after the barrier every CTA picks the
// global winner and builds the same reflector redundantly.  Applying the
// reflector to the local columns is fused with the recomputation of their
// remaining squared norms (the reference recomputes them from scratch each step,
// UDT.jl:175-192, and so do we -- same numbers up to summation order).
// Q is then accumulated backwards over the same resident layout.
// Roofline: shared-memory bandwidth / FP64 FMA (level-2 work), see DESIGN.md.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace dqmc {

struct UdtGeom {
    int cs;       // cluster size
    int tpc;      // threads per column
    int nt;       // threads per CTA
    int nloc;     // max local columns
    int lds;      // leading dimension of the local panel (== tpc mod 16)
    int nv;       // n rounded up to even
    size_t smem;
};

static UdtGeom udt_geometry(int n)
{
    UdtGeom best{};
    for (int cs = 1; cs <= 8; cs *= 2) {
        UdtGeom g{};
        g.cs = cs;
        g.nloc = (n + cs - 1) / cs;
        g.tpc = g.nloc <= 16 ? 16 : (g.nloc <= 32 ? 8 : 4);
        g.nt = ((g.nloc * g.tpc + 31) / 32) * 32;
        if (g.nt > 1024) { g.tpc = 2; g.nt = ((g.nloc * g.tpc + 31) / 32) * 32; }
        if (g.nt > 1024) { g.tpc = 1; g.nt = ((g.nloc + 31) / 32) * 32; }
        if (g.nt < 64) g.nt = 64;
        g.lds = n + (((g.tpc - n) % 16) + 16) % 16;
        g.nv = (n + 1) & ~1;
        const size_t dbl = (size_t)g.nloc * g.lds + (size_t)2 * cs * g.nv + (size_t)2 * g.nv +
                           (size_t)2 * n + g.nloc + 2 * 8 + 64;
        g.smem = dbl * sizeof(double) + ((size_t)g.nloc + n + 2 * 8 + 64) * sizeof(int);
        best = g;
        if (g.smem <= 220 * 1024 && g.nt <= 1024) return g;
    }
    best.smem = (size_t)1 << 30;   // does not fit
    return best;
}

int udt_max_n()
{
    int n = 16;
    while (udt_geometry(n + 8).smem <= 220 * 1024) n += 8;
    return n;
}

__device__ __forceinline__ double group_sum(double v, int tpc)
{
    for (int o = tpc >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void udt_kernel(const UdtParams p, const UdtGeom gm)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = gm.cs, TPC = gm.tpc, NT = gm.nt, lds = gm.lds, nv = gm.nv, n = p.n;
    const int rank = (int)cluster.block_rank();
    const int mat = blockIdx.x / CS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    const int nloc = (n - rank + CS - 1) / CS;     // columns c = s * CS + rank owned here

    extern __shared__ __align__(16) double sm[];
    double* Aloc = sm;                                   // [gm.nloc][lds]
    double* vbuf = Aloc + (size_t)gm.nloc * lds;         // [2][CS][nv] published candidate columns
    double* vcur = vbuf + (size_t)2 * CS * nv;           // [2][nv]
    double* dvec = vcur + 2 * nv;                        // [n]
    double* taus = dvec + n;                             // [n]
    double* norms = taus + n;                            // [gm.nloc]
    double* candval = norms + gm.nloc;                   // [2][8]
    double* redval = candval + 16;                       // [<=32]
    int* colstep = (int*)(redval + 64);                  // [gm.nloc]
    int* perm = colstep + gm.nloc;                       // [n]
    int* candcol = perm + n;                             // [2][8]
    int* redcol = candcol + 16;                          // [<=32]

    const double* Ag = p.A + (long long)mat * p.strideA;
    double* Vg = p.Vwork + (long long)mat * p.strideV;

    // ---- load local columns (scaled), initial squared norms ------------------
    for (int s = warp; s < nloc; s += nwarps) {
        const int c = s * CS + rank;
        const double sc = p.colscale.mode ? scale_at(p.colscale, mat, c) : 1.0;
        double part = 0.0;
        for (int i = lane; i < n; i += 32) {
            const double v = Ag[i + (long long)c * p.ld] * sc;
            Aloc[(size_t)s * lds + i] = v;
            part += v * v;
        }
        part = group_sum(part, 32);
        if (lane == 0) { norms[s] = part; colstep[s] = -1; }
    }
    __syncthreads();

    // publishes this CTA's best remaining column for step j into every peer
    auto publish = [&](int j) {
        const int par = j & 1;
        double bv = -1.0; int bc = 0x7fffffff;
        for (int s = tid; s < nloc; s += NT) {
            if (colstep[s] < 0) {
                const double v = norms[s]; const int c = s * CS + rank;
                if (v > bv || (v == bv && c < bc)) { bv = v; bc = c; }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
            if (ov > bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
        }
        if (lane == 0) { redval[warp] = bv; redcol[warp] = bc; }
        __syncthreads();
        bv = redval[0]; bc = redcol[0];
        for (int w = 1; w < nwarps; ++w) {
            const double ov = redval[w]; const int oc = redcol[w];
            if (ov > bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
        }
        const bool have = bv >= 0.0;
        const int s = have ? bc / CS : 0;
        for (int r = 0; r < CS; ++r) {
            double* rv = cluster.map_shared_rank(vbuf, r) + ((size_t)par * CS + rank) * nv;
            if (have)
                for (int i = j + tid; i < n; i += NT) rv[i] = Aloc[(size_t)s * lds + i];
            if (tid == 0) {
                cluster.map_shared_rank(candval, r)[par * 8 + rank] = bv;
                cluster.map_shared_rank(candcol, r)[par * 8 + rank] = bc;
            }
        }
        __syncthreads();   // redval/redcol reuse
    };

    cluster.sync();     // every CTA of the cluster is resident before the first DSMEM store
    publish(0);
    cluster.sync();

    const int grp = tid / TPC, q = tid - grp * TPC, ngrp = NT / TPC;
    const int npass = (gm.nloc + ngrp - 1) / ngrp;   // identical trip count for every thread (shuffles inside)

    for (int j = 0; j < n; ++j) {
        const int par = j & 1;
        // ---- global winner (identical decision in every CTA) ------------------
        double bv = candval[par * 8]; int bc = candcol[par * 8], br = 0;
        for (int r = 1; r < CS; ++r) {
            const double ov = candval[par * 8 + r]; const int oc = candcol[par * 8 + r];
            if (ov > bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; br = r; }
        }
        const double* raw = vbuf + ((size_t)par * CS + br) * nv;
        // ---- reflector (UDT.jl:157-172) ---------------------------------------
        double xi1 = raw[j];
        double tau, rjj;
        if (bv == 0.0) { tau = 0.0; rjj = xi1; }
        else {
            const double nu = copysign(sqrt(bv), xi1);
            xi1 += nu;
            rjj = -nu; tau = xi1 / nu;
        }
        double* v = vcur;            // vcur[0..nv) holds the normalised vector of this step
        for (int i = j + 1 + tid; i < n; i += NT) {
            // the reference divides (x[i] /= xi1); keep the division for bit-closeness
            const double x = (bv == 0.0) ? 0.0 : raw[i] / xi1;
            v[i] = x;
            if (rank == br) Vg[i + (long long)j * p.ldv] = x;
        }
        if (tid == 0) {
            const double ad = fabs(rjj);
            dvec[j] = (ad == 0.0) ? 1.0 : ad;
            taus[j] = tau;
            perm[j] = bc;
            if (rank == br) { const int s = bc / CS; colstep[s] = j; Aloc[(size_t)s * lds + j] = rjj; }
        }
        __syncthreads();

        // ---- apply H_j to the remaining local columns, fused norm recompute ----
        for (int pass = 0; pass < npass; ++pass) {
            const int s = grp + pass * ngrp;
            const bool act = (s < nloc) && (colstep[s] < 0);
            double* a = Aloc + (size_t)(act ? s : 0) * lds;
            double d0 = 0.0, d1 = 0.0;
            if (act) {
                int i = j + 1 + q;
                for (; i + TPC < n; i += 2 * TPC) { d0 += v[i] * a[i]; d1 += v[i + TPC] * a[i + TPC]; }
                if (i < n) d0 += v[i] * a[i];
            }
            double dot = group_sum(d0 + d1, TPC);
            double nrm = 0.0;
            if (act) {
                dot = (a[j] + dot) * tau;
                double n0 = 0.0, n1 = 0.0;
                int i = j + 1 + q;
                for (; i + TPC < n; i += 2 * TPC) {
                    const double x0 = a[i] - v[i] * dot, x1 = a[i + TPC] - v[i + TPC] * dot;
                    a[i] = x0; a[i + TPC] = x1;
                    n0 += x0 * x0; n1 += x1 * x1;
                }
                if (i < n) { const double x0 = a[i] - v[i] * dot; a[i] = x0; n0 += x0 * x0; }
                nrm = n0 + n1;
            }
            nrm = group_sum(nrm, TPC);
            if (act && q == 0) { a[j] -= dot; norms[s] = nrm; }
        }
        __syncthreads();

        if (j + 1 < n) {
            publish(j + 1);
            cluster.sync();
        }
    }
    __threadfence();
    cluster.sync();     // V (global) written by the step owners is visible to every CTA

    // ---- D, pivot, T ------------------------------------------------------------
    if (rank == 0) {
        double* Dg = p.D + (long long)mat * p.strideD;
        int* pg = p.pivot ? p.pivot + (long long)mat * p.stridePivot : nullptr;
        for (int i = tid; i < n; i += NT) { Dg[i] = dvec[i]; if (pg) pg[i] = perm[i]; }
    }
    {
        double* Tg = p.T + (long long)mat * p.strideT;
        for (int s = warp; s < nloc; s += nwarps) {
            const int c = s * CS + rank, js = colstep[s];
            const int oc = p.pivot_applied ? c : js;
            for (int i = lane; i < n; i += 32)
                Tg[i + (long long)oc * p.ld] = (i <= js) ? Aloc[(size_t)s * lds + i] / dvec[i] : 0.0;
        }
    }
    __syncthreads();

    // ---- explicit Q (UDT.jl:272-288): U = H_0 ... H_{n-1} I, backwards ------------
    for (int s = warp; s < nloc; s += nwarps) {
        const int c = s * CS + rank;
        for (int i = lane; i < n; i += 32) Aloc[(size_t)s * lds + i] = (i == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int k = n - 1; k >= 0; --k) {
        double* vk = vcur + (k & 1) * nv;
        // load v_k (rows k+1..n-1) -- for k == n-1 the range is empty
        for (int i = k + 1 + tid; i < n; i += NT) vk[i] = Vg[i + (long long)k * p.ldv];
        __syncthreads();
        const double tau = taus[k];
        for (int pass = 0; pass < npass; ++pass) {
            const int s = grp + pass * ngrp;
            const int c = s * CS + rank;
            const bool act = (s < nloc) && (c >= k);
            double* a = Aloc + (size_t)(act ? s : 0) * lds;
            double d0 = 0.0, d1 = 0.0;
            if (act) {
                int i = k + 1 + q;
                for (; i + TPC < n; i += 2 * TPC) { d0 += vk[i] * a[i]; d1 += vk[i + TPC] * a[i + TPC]; }
                if (i < n) d0 += vk[i] * a[i];
            }
            double dot = group_sum(d0 + d1, TPC);
            if (act) {
                dot = (a[k] + dot) * tau;
                int i = k + 1 + q;
                for (; i < n; i += TPC) a[i] -= vk[i] * dot;
            }
            __syncwarp();            // every lane of the group has read a[k]
            if (act && q == 0) a[k] -= dot;
        }
        // next iteration writes the other vcur buffer; a[] hazards are per-group only,
        // but the a[k] update by q == 0 must be visible to the group's next dot:
        __syncthreads();
    }
    {
        double* Ug = p.U + (long long)mat * p.strideU;
        for (int s = warp; s < nloc; s += nwarps) {
            const int c = s * CS + rank;
            for (int i = lane; i < n; i += 32) Ug[i + (long long)c * p.ld] = Aloc[(size_t)s * lds + i];
        }
    }
}

bool udt_reg_supported(int n);
cudaError_t launch_udt_reg(const UdtParams& p, cudaStream_t st);

cudaError_t launch_udt(const UdtParams& p, cudaStream_t st)
{
    if (p.batch <= 0) return cudaSuccess;
    // v2: register-resident panel (udt_reg.cu); v1 (shared-memory panel, below) covers larger n
    if (udt_reg_supported(p.n)) return launch_udt_reg(p, st);
    const UdtGeom g = udt_geometry(p.n);
    if (g.smem > 220 * 1024) return cudaErrorInvalidConfiguration;
    static SmemAttr attr;
    cudaError_t e = attr.ensure(udt_kernel, g.smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(p.batch * g.cs));
    cfg.blockDim = dim3((unsigned)g.nt);
    cfg.dynamicSmemBytes = g.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)g.cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    count_launch();
    return cudaLaunchKernelEx(&cfg, udt_kernel, p, g);
}

}  // namespace dqmc
