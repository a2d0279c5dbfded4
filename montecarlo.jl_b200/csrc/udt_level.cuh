// udt_level.cuh -- one level of the multi-level pivoted QR (shared by udt_steps.cu and udt_reg.cu).
#pragma once
#include "common.cuh"

namespace dqmc {

struct UdtLevel {
    // geometry (filled by udt_steps_geometry)
    int cs;         // CTAs (cluster size) per matrix
    int cpt;        // register columns per thread
    int cps;        // columns per thread kept in a shared-memory strip (hybrid geometries; 0 otherwise)
    int rpt;        // rows per thread (rows 8 i + g, i < rpt)
    int nwarps;     // warps per CTA; a warp owns 4 * (cpt + cps) local columns
    size_t smem;
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

// picks (cs, cpt, rpt, nwarps, smem) for a level of size nk; false if no kernel geometry fits
bool udt_steps_geometry(int nk, UdtLevel& g);
cudaError_t launch_udt_steps(const UdtParams& p, const UdtLevel& L, cudaStream_t st);

}  // namespace dqmc
