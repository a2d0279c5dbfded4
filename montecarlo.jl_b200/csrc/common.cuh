// common.cuh -- shared declarations of the dqmc_b200 CUDA library (sm_100a only).
//
// Data layout in HBM: every per-chain matrix is stored as
// `nmat = n_chains * n_flavors` independent n x n column-major FP64 matrices with
// leading dimension ld = n rounded up to even (16-byte aligned columns for
// cp.async), matrix m = chain * n_flavors + block at offset m * ld * n.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

namespace dqmc {

// Kernel-launch counter of the context the calling thread is driving (set by the C ABI entry points; a context is
// not thread-safe, but different threads may drive different contexts, so there is no process-wide counter).
extern thread_local long long* t_launch_counter;
inline void count_launch(int n = 1) { if (t_launch_counter) *t_launch_counter += n; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute (primary context): remember what has been
// configured per device ordinal, so a second context on another GPU of the same process configures its own kernels.
struct SmemAttr {
    std::atomic<size_t> configured[64];
    template <class K> cudaError_t ensure(K kern, size_t bytes)
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        std::atomic<size_t>& slot = configured[dev & 63];
        if (bytes <= slot.load(std::memory_order_acquire)) return cudaSuccess;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        size_t cur = slot.load(std::memory_order_relaxed);
        while (cur < bytes && !slot.compare_exchange_weak(cur, bytes, std::memory_order_release)) {}
        return cudaSuccess;
    }
};

// A diagonal factor fused into a kernel.  `field` mode evaluates
// interaction_matrix_exp! (reference src/flavors/DQMC/fields.jl:380-386, 429-438)
// on the fly from the Int8 HS field: exp(+-alpha) chosen by the sign of conf.
struct Scale {
    int mode;            // 0 none, 1 vec[i], 2 1/vec[i], 3 field, 4 min(1, vec[i]), 5 1/max(1, vec[i])
                         // (4, 5: vmin! / vmaxinv!, reference linalg/real.jl:137-164)
    const double* vec;   // mode 1/2: base pointer, matrix m uses vec + m * stride
    long long stride;
    const int8_t* conf;  // mode 3: conf + chain * cstride + i  (already offset to the slice)
    long long cstride;
    double lut[2][4];    // mode 3: exp(+-power alpha eta(x)) per flavor block; index 0/1 = conf > 0 / < 0 (Hirsch),
                         //         conf - 1 (GHQ, 4 values)
    int nb;              // flavor blocks per chain (matrix m -> chain m / nb, block m % nb)
    int ghq;             // mode 3: 4-state field (index = conf - 1)
};

__host__ __device__ inline Scale no_scale() { Scale s{}; s.mode = 0; s.nb = 1; return s; }

__device__ __forceinline__ double scale_at(const Scale& s, int m, int i)
{
    if (s.mode == 1) return s.vec[(long long)m * s.stride + i];
    if (s.mode == 2) return 1.0 / s.vec[(long long)m * s.stride + i];
    if (s.mode == 4) return fmin(1.0, s.vec[(long long)m * s.stride + i]);
    if (s.mode == 5) return 1.0 / fmax(1.0, s.vec[(long long)m * s.stride + i]);
    const int chain = m / s.nb, blk = m - chain * s.nb;
    const int x = s.conf[(long long)chain * s.cstride + i];
    const int code = s.ghq ? ((x - 1) & 3) : ((x > 0) ? 0 : 1);
    return s.lut[blk][code];
}

// C[m] = beta * C[m] + alpha * diag(rs) * op(A[m]) * diag(ks) * op(B[m]) * diag(cs) + diag(add_diag)
struct GemmParams {
    int M, N, K;
    const double* A; int lda; long long strideA; int transA;   // strideA == 0 -> shared by all matrices
    const double* B; int ldb; long long strideB; int transB;
    double* C; int ldc; long long strideC;
    double alpha, beta;
    Scale rs, ks, cs;
    const double* add_diag; long long add_stride;
    int batch;
};

cudaError_t launch_gemm(const GemmParams& p, cudaStream_t st);

// ---- UDT (column-pivoted Householder QR) ----------------------------------
struct UdtParams {
    int n, ld, batch;
    const double* A; long long strideA;      // input (not modified)
    Scale colscale;                          // fused vmul!(tmp, A, Diagonal(d)) on load
    double* U; long long strideU;            // explicit Q
    double* D; long long strideD;            // |diag R| (0 -> 1)
    double* T; long long strideT;            // pivot_applied: D^-1 R P^T ; else clean upper-triangular D^-1 R (logical order)
    int* pivot; long long stridePivot;       // logical column j came from input column pivot[j] (0-based)
    int pivot_applied;
    double* Vwork; long long strideV; int ldv; // n columns x ldv (>= 32 * ceil(n/32)) scratch per matrix (Householder vectors)
    double* tau; long long strideTau;        // n scratch per matrix (Householder taus)
    double* scratch;                         // batch * udt_reg_scratch_doubles(n, ld) doubles (multi-level QR)
    int* iscratch;                           // batch * udt_reg_scratch_ints(n) ints
};
size_t udt_reg_scratch_doubles(int n, int ld);
size_t udt_reg_scratch_ints(int n);
cudaError_t launch_udt(const UdtParams& p, cudaStream_t st);
int udt_max_n();

// ---- rdivp: A <- A[:, pivot] * inv(T), T clean upper triangular ------------
struct RdivpParams {
    int n, ld, batch;
    double* A; long long strideA;
    const double* T; long long strideT;
    const int* pivot; long long stridePivot;
    double* work; long long strideW;         // n x ld scratch per matrix
};
cudaError_t launch_rdivp(const RdivpParams& p, cudaStream_t st);

// ---- local updates (sweep_spatial) -----------------------------------------
// Proposal tables of the 4-state Gauss-Hermite fields (fields.jl:525-556, 587-610), indexed [x_old - 1][x_new - 1]:
// er = exp(alpha (eta_new - eta_old)), ier = 1 / er, ebm = exp(-alpha (eta_new - eta_old)); gam = gamma(x).
// Filled on the host with libm exp in the reference's own expression order.
struct GhqTables { double er[4][4], ier[4][4], ebm[4][4], gam[4]; };

// One proposal of sweep_spatial, independent of G: Delta per flavor block, the boson factor exp(-dE_boson) and, for the
// GHQ fields, gamma(x_new) / gamma(x_old).  Hirsch: fields.jl:388-393, 440-449 (x_new = -x, dE = -2 alpha x);
// GHQ: fields.jl:525-556, 587-610.  em2a / ep2a = exp(-+2 alpha).
struct Proposal { double Dl[2]; double bos, gn, go; int xnew; };
__device__ __forceinline__ Proposal make_proposal(int kind, int x, int xnew_ghq, const GhqTables& T, double em2a, double ep2a)
{
    Proposal q;
    const bool magnetic = (kind & 1) != 0;
    if (kind < 2) {
        const double e_dE = (x > 0) ? em2a : ep2a;       // exp(dE), dE = -2 alpha x
        const double e_mdE = (x > 0) ? ep2a : em2a;
        q.Dl[0] = e_dE - 1.0;
        q.Dl[1] = (magnetic ? e_mdE : e_dE) - 1.0;
        q.bos = magnetic ? 1.0 : e_mdE;
        q.gn = q.go = 1.0;
        q.xnew = -x;
    } else {
        const int xo = (x - 1) & 3, xn = (xnew_ghq - 1) & 3;
        q.Dl[0] = T.er[xo][xn] - 1.0;
        q.Dl[1] = magnetic ? T.ier[xo][xn] - 1.0 : q.Dl[0];
        q.bos = magnetic ? 1.0 : T.ebm[xo][xn];
        q.gn = T.gam[xn]; q.go = T.gam[xo];
        q.xnew = xnew_ghq;
    }
    return q;
}
// p = exp(-dE_boson) * detratio (local_updates.jl:31); GHQ: detratio * gamma_new / gamma_old first (fields.jl:554, 608)
__device__ __forceinline__ double proposal_prob(int kind, const Proposal& q, double det)
{
    if (kind < 2) return (kind == 0) ? q.bos * det : det;
    const double d = (det * q.gn) / q.go;
    return (kind == 2) ? q.bos * d : d;
}

struct UpdateParams {
    int n, ld, nb, kind, n_chains;
    double* G; long long strideG;            // per matrix
    int8_t* conf_slice; long long cstride;   // conf + slice offset; chain stride
    double alpha;
    const double* uniforms; long long ustride; // table for this slice visit (per chain stride) or null; GHQ fields: the
                                             // N choice uniforms follow the N Metropolis uniforms
    GhqTables ghq;                           // kind >= 2 only
    unsigned long long seed; long long sweep; int step; long long chain0;
    const long long* sweep_ptr;              // device copy of the sweep index (read instead of `sweep` when non-null: a captured
                                             // CUDA graph of one sweep is replayed with a new index every time)
    int check_sign;
    int* accepted;                           // per chain counters (accumulated)
    double* stats;                           // per chain: [neg_count, neg_sumlog, neg_min, neg_max]
    const unsigned char* forced;             // optional teacher-forced decisions for this slice visit
    double* probs;                           // optional trace of p
    unsigned char* decisions;                // optional trace of accept decisions
    long long tstride;                       // per-chain stride of forced / probs / decisions
    int kb;                                  // delay block size
};
cudaError_t launch_update(const UpdateParams& p, cudaStream_t st);
int update_pick_kb(int n, int nb);
cudaError_t launch_update3(const UpdateParams& p, cudaStream_t st);   // submatrix form: G0 + Bc X Br, in-kernel flush
int update3_pick_kb(int n, int nb);

// ---- small lattices: a run of slice steps (sweep_spatial + wrap) in one kernel, G on chip (slicestep.cu) ----------
struct SliceStepParams {
    int n, ld, ldg, nb, kind, n_chains;
    double* G; long long strideG;            // per matrix
    int8_t* conf; long long cstride;         // conf of chain 0, slice 1; chain stride
    const double *eT2, *eT2i;                // exp(-+dtau T), leading dimension ld
    double alpha, em2a, ep2a;
    GhqTables ghq;
    double lut[2][2][4];                     // [e^{+V}, e^{-V}][flavor block][field code]: the Scale look-up tables
    const double* uniforms; long long ustride; int uf;   // table at the first step (per chain stride), uf * n per step
    unsigned long long seed; const long long* sweep_ptr; long long sweep; long long chain0;
    int step0, nsteps, slice0, dir;          // steps step0 .. of the sweep visit slices slice0, slice0 + dir, ...
    int check_sign; int* accepted; double* stats;
    const unsigned char* forced; double* probs; unsigned char* decisions; long long tstride;   // at the first step; + n per step
};
bool slice_steps_supported(int n, int nb);
// a run of slice-matrix products on one operand per matrix (slicestep.cu): op 0 eT2 e^V M (slices ascending), 1 e^V eT2^T M,
// 2 e^-V eT2^-1 M (slices descending); src == nullptr starts from the identity
struct SliceChainParams {
    int n, ld, ldg, nb, ghq, n_mats, op, first, count;
    const double* src; double* dst; long long stride;    // per matrix
    const int8_t* conf; long long cstride;               // conf of chain 0, slice 1; chain stride
    const double* E;                                     // eT2 (op 0, 1) or eT2^-1 (op 2), leading dimension ld
    double lut[2][4];                                    // diagonal factor per flavor block and field code
};
cudaError_t launch_slice_chain(SliceChainParams p, cudaStream_t st);
cudaError_t launch_slice_steps(SliceStepParams p, cudaStream_t st);

// ---- small elementwise helpers ---------------------------------------------
cudaError_t launch_set_identity(double* A, int n, int ld, long long stride, int batch, cudaStream_t st);
cudaError_t launch_fill(double* v, double val, long long count, cudaStream_t st);
cudaError_t launch_permute_cols(const double* A, double* O, const int* pivot, int n, int ld,
                                long long stride, long long pstride, int batch, cudaStream_t st);
// per chain: d = max |A - B| over the chain's nb matrices; if d > thresh accumulate MagnitudeStats
cudaError_t launch_prop_error(const double* A, const double* B, int n, int ld, long long stride_chain,
                              int nb, int n_chains, double thresh, double* stats, cudaStream_t st);
cudaError_t launch_accumulate(const double* G, double* sum, double* sumsq, long long count, cudaStream_t st);
// O = diag(rs) * A * diag(cs) + add (add may be null; O may alias A or add); A == nullptr means the identity
cudaError_t launch_scale_add(double* O, const double* A, Scale rs, Scale cs, const double* add, double add_diag,
                             int n, int ld, long long stride, int batch, cudaStream_t st);

// BitArray(conf .== 1) <-> conf for n_chains configurations of nbits sites x slices each
cudaError_t launch_conf_bits(int8_t* conf, unsigned long long* chunks, long long nvalues, int n_chains, int pack, int ghq,
                             cudaStream_t st);

}  // namespace dqmc
