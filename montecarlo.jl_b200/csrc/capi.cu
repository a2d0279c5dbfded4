// capi.cu -- context, the `propagate` state machine and the C ABI (include/dqmc_b200.h).
//
// Host-side control flow mirrors the reference's stack (src/flavors/DQMC/stack.jl) one to
// one -- build_stack / reverse_build_stack (:257-308), add_slice_sequence_left/right
// (:377-416), calculate_greens (:442-516), wrap_greens! (:594-603), propagate (:605-730) and
// local_sweep (updates/local_updates.jl:7-14) -- but every step is ONE batched kernel launch
// over all chains x flavor blocks of the context, and diagonal factors never exist as
// separate passes: they are fused into the GEMM / QR kernels.  All chains of a context walk
// the imaginary-time axis in lockstep (the schedule does not depend on the data).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "ctx.cuh"
#include "../../include/dqmc_rng.h"

namespace dqmc { thread_local long long* t_launch_counter = nullptr; }

static thread_local std::string g_create_error;

// ---- host <-> device matrix copies (host ld = N, device ld = c->ld) ------------------------
cudaError_t h2d_mats(dqmc_ctx* c, double* dst, const double* src, long long nmats)
{
    // equal pitches: ONE contiguous copy (a 2-D copy is issued row by row: 2 KB pieces at n = 256)
    if (c->ld == c->N) return cudaMemcpyAsync(dst, src, (size_t)c->N * c->N * nmats * 8, cudaMemcpyHostToDevice, c->st);
    return cudaMemcpy2DAsync(dst, (size_t)c->ld * 8, src, (size_t)c->N * 8, (size_t)c->N * 8,
                             (size_t)c->N * nmats, cudaMemcpyHostToDevice, c->st);
}
cudaError_t d2h_mats(dqmc_ctx* c, double* dst, const double* src, long long nmats)
{
    if (c->ld == c->N) return cudaMemcpyAsync(dst, src, (size_t)c->N * c->N * nmats * 8, cudaMemcpyDeviceToHost, c->st);
    return cudaMemcpy2DAsync(dst, (size_t)c->N * 8, src, (size_t)c->ld * 8, (size_t)c->N * 8,
                             (size_t)c->N * nmats, cudaMemcpyDeviceToHost, c->st);
}

// ---- fused diagonal factors ------------------------------------------------------------------
// interaction_matrix_exp!(..., slice, power) (fields.jl:380-386, 429-438; GHQ :533-546, 596-602) as a Scale:
// a look-up table over the 2 (Hirsch) or 4 (GHQ) field values, per flavor block (the magnetic fields flip the sign
// of the exponent in block 2)
Scale field_scale(dqmc_ctx* c, int slice, double power)
{
    Scale s{};
    s.mode = 3;
    s.conf = c->conf + (long long)(slice - 1) * c->N;
    s.cstride = (long long)c->M * c->N;
    s.nb = c->nb; s.ghq = c->ghq ? 1 : 0;
    const bool magnetic = (c->kind & 1) != 0;
    for (int k = 0; k < 4; ++k) {
        // Hirsch: x = +1 (index 0), -1 (index 1); GHQ: eta(x), x = 1..4
        const double x = c->ghq ? c->eta[k] : ((k == 0) ? 1.0 : -1.0);
        s.lut[0][k] = exp(power * c->alpha * x);
        s.lut[1][k] = magnetic ? exp(-power * c->alpha * x) : s.lut[0][k];
    }
    return s;
}
Scale vec_scale(dqmc_ctx* c, const double* v, bool inverse)
{
    Scale s{}; s.mode = inverse ? 2 : 1; s.vec = v; s.stride = c->N; s.nb = c->nb; return s;
}

GemmParams gemm_base(dqmc_ctx* c)
{
    GemmParams g{};
    g.M = g.N = g.K = c->N;
    g.lda = g.ldb = g.ldc = c->ld;
    g.strideA = g.strideB = g.strideC = c->ms;
    g.alpha = 1.0; g.beta = 0.0;
    g.rs = no_scale(); g.ks = no_scale(); g.cs = no_scale();
    g.batch = c->nmat;
    return g;
}

// dst = op(A) * op(B) with optional fused factors
cudaError_t mm(dqmc_ctx* c, double* dst, const double* A, bool tA, bool sharedA, const double* Bm, bool tB,
                      bool sharedB, Scale rs, Scale ks, Scale cs, const double* add_diag, double alpha, double beta)
{
    ProfScope ps(c, DQMC_PROF_GEMM);
    GemmParams g = gemm_base(c);
    g.A = A; g.transA = tA; if (sharedA) g.strideA = 0;
    g.B = Bm; g.transB = tB; if (sharedB) g.strideB = 0;
    g.C = dst; g.rs = rs; g.ks = ks; g.cs = cs;
    g.add_diag = add_diag; g.add_stride = c->N;
    g.alpha = alpha; g.beta = beta;
    return launch_gemm(g, c->st);
}

// stack.jl:319-367, out of place
cudaError_t slice_left(dqmc_ctx* c, double* dst, const double* src, int slice)        // eT2 * eV * M
{ return mm(c, dst, c->eT2, false, true, src, false, false, no_scale(), field_scale(c, slice, 1.0)); }
cudaError_t slice_right(dqmc_ctx* c, double* dst, const double* src, int slice)       // M * eT2 * eV
{ return mm(c, dst, src, false, false, c->eT2, false, true, no_scale(), no_scale(), field_scale(c, slice, 1.0)); }
cudaError_t slice_inv_right(dqmc_ctx* c, double* dst, const double* src, int slice)   // M * eV^-1 * eT2^-1
{ return mm(c, dst, src, false, false, c->eT2i, false, true, no_scale(), field_scale(c, slice, -1.0)); }
cudaError_t slice_inv_left(dqmc_ctx* c, double* dst, const double* src, int slice)    // eV^-1 * eT2^-1 * M
{ return mm(c, dst, c->eT2i, false, true, src, false, false, field_scale(c, slice, -1.0)); }
cudaError_t slice_daggered_left(dqmc_ctx* c, double* dst, const double* src, int slice) // eV' * eT2' * M
{ return mm(c, dst, c->eT2, true, true, src, false, false, field_scale(c, slice, 1.0)); }

// wrap_greens! (stack.jl:594-603): gf -> tmp -> gf
cudaError_t wrap_greens(dqmc_ctx* c, double* gf, double* tmp, int curr_slice, int direction)
{
    cudaError_t e;
    if (direction == -1) {
        if ((e = slice_inv_left(c, tmp, gf, curr_slice - 1)) != cudaSuccess) return e;
        return slice_right(c, gf, tmp, curr_slice - 1);
    }
    if ((e = slice_left(c, tmp, gf, curr_slice)) != cudaSuccess) return e;
    return slice_inv_right(c, gf, tmp, curr_slice);
}

cudaError_t udt(dqmc_ctx* c, const double* A, Scale colscale, double* U, double* D, double* T, bool apply_pivot)
{
    ProfScope ps(c, DQMC_PROF_UDT);
    UdtParams p{};
    p.n = c->N; p.ld = c->ld; p.batch = c->nmat;
    p.A = A; p.strideA = c->ms; p.colscale = colscale;
    p.U = U; p.strideU = c->ms; p.D = D; p.strideD = c->N; p.T = T; p.strideT = c->ms;
    p.pivot = c->pivot; p.stridePivot = c->N; p.pivot_applied = apply_pivot ? 1 : 0;
    p.Vwork = c->Vwork; p.ldv = c->ldv; p.strideV = (long long)c->ldv * c->N; p.tau = c->tau; p.strideTau = c->N;
    p.scratch = c->udt_scratch; p.iscratch = c->udt_iscratch;
    return launch_udt(p, c->st);
}

cudaError_t rdivp(dqmc_ctx* c, double* A, const double* T, double* work)
{
    ProfScope ps(c, DQMC_PROF_RDIVP);
    RdivpParams p{};
    p.n = c->N; p.ld = c->ld; p.batch = c->nmat;
    p.A = A; p.strideA = c->ms; p.T = T; p.strideT = c->ms;
    p.pivot = c->pivot; p.stridePivot = c->N; p.work = work; p.strideW = c->ms;
    return launch_rdivp(p, c->st);
}

cudaError_t copy_mats(dqmc_ctx* c, double* dst, const double* src)
{ ProfScope ps(c, DQMC_PROF_OTHER); return cudaMemcpyAsync(dst, src, (size_t)c->nmat * c->ms * 8, cudaMemcpyDeviceToDevice, c->st); }
cudaError_t copy_vecs(dqmc_ctx* c, double* dst, const double* src)
{ ProfScope ps(c, DQMC_PROF_OTHER); return cudaMemcpyAsync(dst, src, (size_t)c->nmat * c->N * 8, cudaMemcpyDeviceToDevice, c->st); }
cudaError_t ident(dqmc_ctx* c, double* A) { ProfScope ps(c, DQMC_PROF_OTHER); return launch_set_identity(A, c->N, c->ld, c->ms, c->nmat, c->st); }
cudaError_t ones(dqmc_ctx* c, double* v) { ProfScope ps(c, DQMC_PROF_OTHER); return launch_fill(v, 1.0, (long long)c->nmat * c->N, c->st); }


// First half of calculate_greens_AVX! (stack.jl:442-480) == calculate_inv_greens_udt
// (updates/global_updates.jl:25-52): afterwards G^-1 = Tl Ul Dr Tr Ur^-1 with det(G) = 1 / prod(Dr).
cudaError_t calculate_inv_greens_udt(dqmc_ctx* c, double* G)
{
    // G = Dl (Tl Tr') Dr                                            :450-452
    CE(mm(c, G, c->Tl, false, false, c->Tr, true, false, vec_scale(c, c->Dl), no_scale(), vec_scale(c, c->Dr)));
    // Tr, Dr, G = udt(G), unpivoted form                            :453
    CE(udt(c, G, no_scale(), c->Tr, c->Dr, G, false));
    CE(mm(c, c->Tl, c->Ul, false, false, c->Tr, false, false));      // Tl = Ul Tr          :464
    CE(rdivp(c, c->Ur, G, c->Ul));                                   // Ur = Ur / G         :465
    // Tr = Tl' Ur + Diagonal(Dr)                                    :466, 472
    CE(mm(c, c->Tr, c->Tl, true, false, c->Ur, false, false, no_scale(), no_scale(), no_scale(), c->Dr));
    return udt(c, c->Tr, no_scale(), c->Ul, c->Dr, c->Tr, false);    // :480
}

// calculate_greens_AVX! (stack.jl:442-496); destroys Ul, Dl, Tl, Ur, Dr, Tr like the reference.
cudaError_t calculate_greens(dqmc_ctx* c, double* G)
{
    CE(calculate_inv_greens_udt(c, G));
    CE(rdivp(c, c->Ur, c->Tr, G));                                   // :481
    CE(mm(c, c->Tr, c->Tl, false, false, c->Ul, false, false));      // Tr = Tl Ul          :482
    // G = (Ur Diagonal(1 / Dr)) Tr'                                 :486-493
    CE(mm(c, G, c->Ur, false, false, c->Tr, true, false, no_scale(), vec_scale(c, c->Dr, true)));
    // the reference leaves 1 / Dr in Dl (:486), which propose_global_from_conf reads as det(G)
    // (global_updates.jl:151-154); Ul..Tr are scratch for the measurement code, so keep a copy
    if (G == c->greens) CE(copy_vecs(c, c->Dgreens, c->Dr));
    return cudaSuccess;
}

// A run of slice-matrix products on one operand: for small lattices one kernel (slicestep.cu), else one GEMM per slice.
// op 0: dst = B_{first + count - 1} ... B_first src; op 1: daggered, slices first, first - 1, ...; op 2: inverse, descending.
// Returns the buffer that holds the result (one of bufs[0], bufs[1]; src itself when count == 0).
cudaError_t slice_chain(dqmc_ctx* c, int op, const double* src, int first, int count, double* bufs[2], const double** out)
{
    if (count <= 0) { *out = src; return cudaSuccess; }
    if (c->fused_steps) {
        ProfScope ps(c, DQMC_PROF_GEMM);
        SliceChainParams p{};
        p.n = c->N; p.ld = c->ld; p.nb = c->nb; p.ghq = c->ghq ? 1 : 0; p.n_mats = c->nmat; p.op = op;
        p.first = first; p.count = count;
        p.src = src; p.dst = bufs[0]; p.stride = c->ms;
        p.conf = c->conf; p.cstride = (long long)c->M * c->N;
        p.E = (op == 2) ? c->eT2i : c->eT2;
        const Scale f = field_scale(c, 1, (op == 2) ? -1.0 : 1.0);
        for (int b = 0; b < 2; ++b)
            for (int k = 0; k < 4; ++k) p.lut[b][k] = f.lut[b][k];
        *out = bufs[0];
        return launch_slice_chain(p, c->st);
    }
    int w = 0;
    for (int i = 0; i < count; ++i) {
        const int s = (op == 0) ? first + i : first - i;
        if (op == 0) CE(slice_left(c, bufs[w], src, s));
        else if (op == 1) CE(slice_daggered_left(c, bufs[w], src, s));
        else CE(slice_inv_left(c, bufs[w], src, s));
        src = bufs[w]; w ^= 1;
    }
    *out = src;
    return cudaSuccess;
}

// add_slice_sequence_left (stack.jl:377-393), idx 1-based
static cudaError_t add_slice_sequence_left(dqmc_ctx* c, int idx)
{
    const double* src = slot_mat(c, c->u_stack, idx - 1);
    double* bufs[2] = {c->curr_U, c->tmp2};
    CE(slice_chain(c, 0, src, c->rfirst[idx - 1], c->rlast[idx - 1] - c->rfirst[idx - 1] + 1, bufs, &src));
    // tmp1 = curr_U * Diagonal(d_stack[idx]) is fused into the QR load
    CE(udt(c, src, vec_scale(c, slot_vec(c, c->d_stack, idx - 1)), slot_mat(c, c->u_stack, idx),
           slot_vec(c, c->d_stack, idx), c->tmp1, true));
    return mm(c, slot_mat(c, c->t_stack, idx), c->tmp1, false, false, slot_mat(c, c->t_stack, idx - 1), false, false);
}

// add_slice_sequence_right (stack.jl:402-416)
static cudaError_t add_slice_sequence_right(dqmc_ctx* c, int idx)
{
    const double* src = slot_mat(c, c->u_stack, idx);
    double* bufs[2] = {c->curr_U, c->tmp2};
    CE(slice_chain(c, 1, src, c->rlast[idx - 1], c->rlast[idx - 1] - c->rfirst[idx - 1] + 1, bufs, &src));
    CE(udt(c, src, vec_scale(c, slot_vec(c, c->d_stack, idx)), slot_mat(c, c->u_stack, idx - 1),
           slot_vec(c, c->d_stack, idx - 1), c->tmp1, true));
    return mm(c, slot_mat(c, c->t_stack, idx - 1), c->tmp1, false, false, slot_mat(c, c->t_stack, idx), false, false);
}

cudaError_t load_udt(dqmc_ctx* c, double* U, double* D, double* T, int slot)   // slot < 0 -> identity
{
    if (slot < 0) { CE(ident(c, U)); CE(ones(c, D)); return ident(c, T); }
    CE(copy_mats(c, U, slot_mat(c, c->u_stack, slot)));
    CE(copy_vecs(c, D, slot_vec(c, c->d_stack, slot)));
    return copy_mats(c, T, slot_mat(c, c->t_stack, slot));
}
static cudaError_t clear_slot(dqmc_ctx* c, int slot)
{
    CE(ident(c, slot_mat(c, c->u_stack, slot)));
    { ProfScope ps(c, DQMC_PROF_OTHER); CE(launch_fill(slot_vec(c, c->d_stack, slot), 1.0, (long long)c->nmat * c->N, c->st)); }
    return ident(c, slot_mat(c, c->t_stack, slot));
}

static cudaError_t prop_check(dqmc_ctx* c)
{
    ProfScope ps(c, DQMC_PROF_OTHER);
    return launch_prop_error(c->greens_temp, c->greens, c->N, c->ld, (long long)c->nb * c->ms, c->nb, c->B, 1e-7,
                             c->stats_prop, c->st);
}

// One side of calculate_greens(mc, slice) (stack.jl:525-583) / inv_det (global_updates.jl:70-137): the UDT
// form of B_slice ... B_1 (dagger == false) or of (B_M ... B_{slice+1})^T (dagger == true), stabilised
// whenever k % safe_mult == 0.
cudaError_t build_chain_udt(dqmc_ctx* c, int slice, int safe_mult, bool dagger, double* U, double* D, double* T)
{
    CE(load_udt(c, U, D, T, -1));
    double* cur = c->curr_U; double* oth = c->tmp2;
    CE(ident(c, cur));
    auto stabilise = [&](double* Uout) -> cudaError_t {
        CE(udt(c, cur, vec_scale(c, D), Uout, D, c->tmp1, true));
        CE(copy_mats(c, c->greens_temp, T));
        return mm(c, T, c->tmp1, false, false, c->greens_temp, false, false);
    };
    if (dagger) {
        for (int k = c->M; k >= slice + 1; --k) {
            CE(slice_daggered_left(c, oth, cur, k)); std::swap(cur, oth);
            if (k % safe_mult == 0) { CE(stabilise(oth)); std::swap(cur, oth); }
        }
    } else {
        for (int k = 1; k <= slice; ++k) {
            CE(slice_left(c, oth, cur, k)); std::swap(cur, oth);
            if (k % safe_mult == 0) { CE(stabilise(oth)); std::swap(cur, oth); }
        }
    }
    return stabilise(U);
}

// build_stack (stack.jl:257-281)
static cudaError_t forward_build(dqmc_ctx* c)
{
    CE(clear_slot(c, 0));
    for (int i = 1; i <= c->C; ++i) CE(add_slice_sequence_left(c, i));
    c->current_slice = c->M + 1; c->current_range = c->C; c->direction = -1;
    CE(load_udt(c, c->Ul, c->Dl, c->Tl, c->C));
    CE(load_udt(c, c->Ur, c->Dr, c->Tr, -1));
    return calculate_greens(c, c->greens);
}

// reverse_build_stack (stack.jl:284-308)
cudaError_t reverse_build(dqmc_ctx* c)
{
    CE(clear_slot(c, c->C));
    for (int i = c->C; i >= 1; --i) CE(add_slice_sequence_right(c, i));
    c->current_slice = 0; c->current_range = 1; c->direction = 1;
    CE(load_udt(c, c->Ul, c->Dl, c->Tl, -1));
    CE(load_udt(c, c->Ur, c->Dr, c->Tr, 0));
    return calculate_greens(c, c->greens);
}

// propagate (stack.jl:605-730)
cudaError_t propagate(dqmc_ctx* c)
{
    c->current_slice += c->direction;
    if (c->direction == 1) {
        if (c->current_slice == 1) {
            CE(clear_slot(c, 0));
        } else if (c->current_slice - 1 == c->rlast[c->current_range - 1]) {
            const int idx = c->current_range;
            CE(load_udt(c, c->Ur, c->Dr, c->Tr, idx));
            CE(add_slice_sequence_left(c, idx));
            CE(load_udt(c, c->Ul, c->Dl, c->Tl, idx));
            if (c->check_prop) {
                CE(copy_mats(c, c->greens_temp, c->greens));
                CE(wrap_greens(c, c->greens_temp, c->tmp1, c->current_slice - 1, 1));   // :638-640
            }
            CE(calculate_greens(c, c->greens));
            if (c->check_prop) CE(prop_check(c));
            if (c->current_range == c->C) { c->direction = -1; return propagate(c); }
            c->current_range += 1;
        } else {
            CE(wrap_greens(c, c->greens, c->tmp1, c->current_slice - 1, 1));
        }
    } else {
        if (c->current_slice == c->M) {
            CE(clear_slot(c, c->C));
            CE(wrap_greens(c, c->greens, c->tmp1, c->current_slice + 1, -1));
        } else if (c->current_slice + 1 == c->rfirst[c->current_range - 1]) {
            const int idx = c->current_range;
            CE(load_udt(c, c->Ul, c->Dl, c->Tl, idx - 1));
            CE(add_slice_sequence_right(c, idx));
            CE(load_udt(c, c->Ur, c->Dr, c->Tr, idx - 1));
            if (c->check_prop) CE(copy_mats(c, c->greens_temp, c->greens));
            CE(calculate_greens(c, c->greens));
            if (c->check_prop) CE(prop_check(c));
            if (c->current_range == 1) { c->direction = 1; return propagate(c); }
            CE(wrap_greens(c, c->greens, c->tmp1, c->current_slice + 1, -1));
            c->current_range -= 1;
        } else {
            CE(wrap_greens(c, c->greens, c->tmp1, c->current_slice + 1, -1));
        }
    }
    return cudaSuccess;
}

// sweep_spatial at the current slice (local_updates.jl:23-60); table pointers are device pointers
// already offset to this slice visit, with per-chain strides ustride / tstride.
static cudaError_t sweep_spatial(dqmc_ctx* c, int step, const double* d_unif, long long ustride,
                                 const unsigned char* d_forced, double* d_probs, unsigned char* d_dec,
                                 long long tstride)
{
    ProfScope ps(c, DQMC_PROF_UPDATE);
    UpdateParams p{};
    p.n = c->N; p.ld = c->ld; p.nb = c->nb; p.kind = c->kind; p.n_chains = c->B;
    p.G = c->greens; p.strideG = c->ms;
    p.conf_slice = c->conf + (long long)(c->current_slice - 1) * c->N; p.cstride = (long long)c->M * c->N;
    p.alpha = c->alpha;
    p.uniforms = d_unif; p.ustride = ustride;
    if (c->ghq) p.ghq = c->ghq_tab;
    p.seed = c->seed; p.sweep = c->sweep_index; p.sweep_ptr = c->d_sweep_index; p.step = step; p.chain0 = c->chain_offset;
    p.check_sign = c->check_sign;
    p.accepted = c->accepted; p.stats = c->stats_neg;
    p.forced = d_forced; p.probs = d_probs; p.decisions = d_dec; p.tstride = tstride;
    p.kb = c->kb;
    c->generation += 1;
    if (c->update_version == 1) return launch_update(p, c->st);
    return launch_update3(p, c->st);
}

__global__ void bump_kernel(long long* counter) { *counter += 1; }

// local_sweep (local_updates.jl:7-14); tables are device pointers [B][2M][N] or null
static cudaError_t local_sweep(dqmc_ctx* c, const double* d_unif, const unsigned char* d_forced, double* d_probs,
                               unsigned char* d_dec)
{
    const long long ts = (long long)2 * c->M * c->N;
    const int uf = c->ghq ? 2 : 1;                       // GHQ: Metropolis + choice uniforms per proposal
    for (int step = 0; step < 2 * c->M;) {
        const long long off = (long long)step * c->N;
        // the propagate after this sweep_spatial is a plain wrap unless the slice closes its range (stack.jl:605-730)
        auto plain = [&](int slice) {
            return (c->direction == 1) ? slice != c->rlast[c->current_range - 1] : slice != c->rfirst[c->current_range - 1];
        };
        if (c->fused_steps && plain(c->current_slice)) {
            int run = 0;
            while (step + run < 2 * c->M && plain(c->current_slice + run * c->direction)) ++run;
            ProfScope ps(c, DQMC_PROF_UPDATE);
            SliceStepParams p{};
            p.n = c->N; p.ld = c->ld; p.nb = c->nb; p.kind = c->kind; p.n_chains = c->B;
            p.G = c->greens; p.strideG = c->ms;
            p.conf = c->conf; p.cstride = (long long)c->M * c->N;
            p.eT2 = c->eT2; p.eT2i = c->eT2i; p.alpha = c->alpha;
            if (c->ghq) p.ghq = c->ghq_tab;
            const Scale sp = field_scale(c, 1, 1.0), sn = field_scale(c, 1, -1.0);
            for (int b = 0; b < 2; ++b)
                for (int k = 0; k < 4; ++k) { p.lut[0][b][k] = sp.lut[b][k]; p.lut[1][b][k] = sn.lut[b][k]; }
            p.uniforms = d_unif ? d_unif + uf * off : nullptr; p.ustride = uf * ts; p.uf = uf;
            p.seed = c->seed; p.sweep_ptr = c->d_sweep_index; p.sweep = c->sweep_index; p.chain0 = c->chain_offset;
            p.step0 = step; p.nsteps = run; p.slice0 = c->current_slice; p.dir = c->direction;
            p.check_sign = c->check_sign; p.accepted = c->accepted; p.stats = c->stats_neg;
            p.forced = d_forced ? d_forced + off : nullptr; p.probs = d_probs ? d_probs + off : nullptr;
            p.decisions = d_dec ? d_dec + off : nullptr; p.tstride = ts;
            CE(launch_slice_steps(p, c->st));
            c->current_slice += run * c->direction;      // what `run` plain propagates do to the state
            c->generation += run;
            step += run;
            continue;
        }
        CE(sweep_spatial(c, step, d_unif ? d_unif + uf * off : nullptr, uf * ts, d_forced ? d_forced + off : nullptr,
                         d_probs ? d_probs + off : nullptr, d_dec ? d_dec + off : nullptr, ts));
        CE(propagate(c));
        ++step;
    }
    c->sweep_index += 1;
    bump_kernel<<<1, 1, 0, c->st>>>(c->d_sweep_index);
    count_launch();
    return cudaGetLastError();
}

// One sweep through a captured CUDA graph.  The first sweeps of a context run eagerly (per-device kernel attributes
// get configured outside of stream capture), then the launch sequence of local_sweep is captured once per RNG mode
// and replayed: the update kernels read the sweep index from device memory, everything else in the sequence is the
// same for every sweep (the stack returns to (slice 1, direction +1) and every buffer to its role).
static cudaError_t local_sweep_graphed(dqmc_ctx* c, const double* d_unif)
{
    const int gi = d_unif ? 1 : 0;
    if (!c->graph_ok || c->prof_on || c->eager_sweeps < 1) {
        c->eager_sweeps += 1;
        return local_sweep(c, d_unif, nullptr, nullptr, nullptr);
    }
    if (!c->sweep_graph[gi]) {
        cudaGraph_t graph = nullptr;
        const long long l0 = c->launches, s0 = c->sweep_index, g0 = c->generation;
        cudaError_t e = cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            e = local_sweep(c, d_unif, nullptr, nullptr, nullptr);
            const cudaError_t e2 = cudaStreamEndCapture(c->st, &graph);
            if (e == cudaSuccess) e = e2;
        }
        if (e == cudaSuccess) e = cudaGraphInstantiate(&c->sweep_graph[gi], graph, 0);
        if (graph) cudaGraphDestroy(graph);
        c->sweep_graph_launches[gi] = c->launches - l0;
        c->launches = l0; c->sweep_index = s0; c->generation = g0;      // nothing has executed yet
        if (e != cudaSuccess) {                              // capture not possible: stay eager for good
            cudaGetLastError();
            c->graph_ok = false; c->sweep_graph[gi] = nullptr;
            c->current_slice = 1; c->current_range = 1; c->direction = 1;
            return local_sweep(c, d_unif, nullptr, nullptr, nullptr);
        }
    }
    CE(cudaGraphLaunch(c->sweep_graph[gi], c->st));
    c->launches += c->sweep_graph_launches[gi];
    c->sweep_index += 1;
    c->generation += 2 * c->M;
    return cudaSuccess;
}

// greens!(mc): target = eThalf^-1 * (G * eThalf) (greens.jl:114-125)
static cudaError_t measured_greens(dqmc_ctx* c, double* out, double* tmp)
{
    CE(mm(c, tmp, c->greens, false, false, c->eTh, false, true));
    return mm(c, out, c->eThi, false, true, tmp, false, false);
}

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int32_t dqmc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int32_t dqmc_max_sites(void) { return udt_max_n(); }

const char* dqmc_last_error(const dqmc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int64_t dqmc_kernel_launches(const dqmc_ctx* c) { return c ? c->launches : 0; }

int32_t dqmc_destroy(dqmc_ctx* c)
{
    if (!c) return DQMC_OK;
    cudaSetDevice(c->device);
    if (t_launch_counter == &c->launches) t_launch_counter = nullptr;
    if (c->st) cudaStreamSynchronize(c->st);
    dqmc_comm_destroy(c);
    ut_destroy(c);
    meas_destroy(c);
    for (cudaGraphExec_t g : c->sweep_graph) if (g) cudaGraphExecDestroy(g);
    for (void* p : c->allocs) cudaFree(p);
    for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
    return DQMC_OK;
}

int32_t dqmc_create(const dqmc_desc* d, dqmc_ctx** out)
{
    if (!d || !out) { g_create_error = "null argument"; return DQMC_ERR_INVALID; }
    *out = nullptr;
    if (d->n_sites < 1 || d->n_slices < 1 || d->n_chains < 1 || d->n_ranges < 1 ||
        d->field_kind < DQMC_FIELD_DENSITY_HIRSCH || d->field_kind > DQMC_FIELD_MAGNETIC_GHQ ||
        !d->range_first || !d->range_last || !d->hopping_exp_squared || !d->hopping_exp_inv_squared ||
        !d->hopping_exp || !d->hopping_exp_inv) {
        g_create_error = "invalid descriptor"; return DQMC_ERR_INVALID;
    }
    if (d->range_first[0] != 1 || d->range_last[d->n_ranges - 1] != d->n_slices) {
        g_create_error = "ranges must cover 1..n_slices (stack.jl:170-171)"; return DQMC_ERR_INVALID;
    }
    for (int i = 0; i < d->n_ranges; ++i) {
        if (d->range_last[i] < d->range_first[i] || (i > 0 && d->range_first[i] != d->range_last[i - 1] + 1)) {
            g_create_error = "ranges must be contiguous and non-empty"; return DQMC_ERR_INVALID;
        }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        g_create_error = "no CUDA device: the dqmc_b200 library has no CPU fallback"; return DQMC_ERR_NO_DEVICE;
    }
    if (d->device < 0 || d->device >= ndev) { g_create_error = "bad device ordinal"; return DQMC_ERR_INVALID; }
    if (d->n_sites > udt_max_n()) {
        g_create_error = "n_sites exceeds the UDT kernel's on-chip capacity (dqmc_max_sites)"; return DQMC_ERR_UNSUPPORTED;
    }
    cudaSetDevice(d->device);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, d->device);
    if (prop.major != 10) {
        g_create_error = "device is not sm_100 (B200): this library is built for sm_100a only"; return DQMC_ERR_NO_DEVICE;
    }

    dqmc_ctx* c = new dqmc_ctx();
    c->N = d->n_sites; c->M = d->n_slices; c->kind = d->field_kind; c->nb = (c->kind & 1) ? 2 : 1;
    c->ghq = c->kind >= DQMC_FIELD_DENSITY_GHQ;
    if (c->ghq) {                                        // fields.jl:513-556, 575-610
        dqmc_ghq_tables(c->eta, c->ghq_tab.gam);
        for (int xo = 0; xo < 4; ++xo)
            for (int xn = 0; xn < 4; ++xn) {
                const double dE = d->alpha * (c->eta[xn] - c->eta[xo]);
                c->ghq_tab.er[xo][xn] = exp(dE);
                c->ghq_tab.ier[xo][xn] = 1.0 / c->ghq_tab.er[xo][xn];
                c->ghq_tab.ebm[xo][xn] = exp(-dE);
            }
    }
    c->B = d->n_chains; c->C = d->n_ranges;
    c->rfirst.assign(d->range_first, d->range_first + c->C);
    c->rlast.assign(d->range_last, d->range_last + c->C);
    c->alpha = d->alpha; c->check_sign = d->check_sign_problem; c->check_prop = d->check_propagation_error;
    c->seed = d->seed; c->chain_offset = d->chain_offset; c->device = d->device;
    c->ld = (c->N + 1) & ~1; c->ms = (long long)c->ld * c->N; c->nmat = c->B * c->nb;
    c->ldv = ((c->N + 31) / 32) * 32;
    // update3.cu (submatrix form) for n >= 96, update.cu (delayed rank-kb factors) below; desc.update_variant forces one
    // (measured: update3 wins from n = 144 up -- cfg 3 / 4 / 5 -- and loses at n = 64, where its per-block overheads
    // outweigh the saved flush passes: cfg 2 5195 vs 5970 sweeps/s)
    if (d->update_variant != 0 && d->update_variant != 1 && d->update_variant != 3) {
        delete c; g_create_error = "update_variant must be 0 (auto), 1 or 3"; return DQMC_ERR_INVALID;
    }
    c->update_version = d->update_variant ? d->update_variant : (c->N >= 96 ? 3 : 1);
    c->fused_steps = d->update_variant == 0 && slice_steps_supported(c->N, c->nb);
    t_launch_counter = &c->launches;
    if (c->update_version == 1) {
        c->kb = d->delay_block > 0 ? ((d->delay_block + 3) & ~3) : update_pick_kb(c->N, c->nb);
        const int kmax = update_pick_kb(c->N, c->nb);
        if (c->kb > kmax) c->kb = kmax;
    } else {
        c->kb = update3_pick_kb(c->N, c->nb);
        if (d->delay_block > 0) c->kb = std::min(c->kb, (d->delay_block + 3) & ~3);
    }
    auto bail = [&](cudaError_t e, const char* what) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(e);
        dqmc_destroy(c);
        return DQMC_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
    const size_t mat = (size_t)c->nmat * c->ms, vec = (size_t)c->nmat * c->N;
#define A_(ptr, cnt) if ((e = dalloc(c, &c->ptr, (cnt))) != cudaSuccess) return bail(e, "cudaMalloc " #ptr)
    A_(eT2, c->ms); A_(eT2i, c->ms); A_(eTh, c->ms); A_(eThi, c->ms);
    A_(conf, (size_t)c->B * c->M * c->N);
    A_(u_stack, mat * (c->C + 1)); A_(t_stack, mat * (c->C + 1)); A_(d_stack, vec * (c->C + 1));
    A_(greens, mat); A_(greens_temp, mat); A_(Ul, mat); A_(Ur, mat); A_(Tl, mat); A_(Tr, mat);
    A_(tmp1, mat); A_(tmp2, mat); A_(curr_U, mat); A_(Vwork, (size_t)c->nmat * c->ldv * c->N);
    A_(Dl, vec); A_(Dr, vec); A_(tau, vec); A_(Dgreens, vec);
    A_(udt_scratch, (size_t)c->nmat * udt_reg_scratch_doubles(c->N, c->ld));
    A_(udt_iscratch, (size_t)c->nmat * udt_reg_scratch_ints(c->N));
    A_(pivot, vec); A_(accepted, (size_t)c->B); A_(d_sweep_index, 1);
    A_(stats_neg, (size_t)c->B * 4); A_(stats_prop, (size_t)c->B * 4);
    c->obs_len = 1 + 2 * (long long)c->nb * c->ms;
    A_(obs, (size_t)c->obs_len);
#undef A_
    if ((e = h2d_mats(c, c->eT2, d->hopping_exp_squared, 1)) != cudaSuccess) return bail(e, "upload");
    if ((e = h2d_mats(c, c->eT2i, d->hopping_exp_inv_squared, 1)) != cudaSuccess) return bail(e, "upload");
    if ((e = h2d_mats(c, c->eTh, d->hopping_exp, 1)) != cudaSuccess) return bail(e, "upload");
    if ((e = h2d_mats(c, c->eThi, d->hopping_exp_inv, 1)) != cudaSuccess) return bail(e, "upload");
    {   // MagnitudeStats start at min = +Inf, max = -Inf (statistics.jl:16)
        std::vector<double> init((size_t)c->B * 4);
        for (int b = 0; b < c->B; ++b) { init[4 * b] = 0; init[4 * b + 1] = 0; init[4 * b + 2] = INFINITY; init[4 * b + 3] = -INFINITY; }
        if ((e = cudaMemcpyAsync(c->stats_neg, init.data(), init.size() * 8, cudaMemcpyHostToDevice, c->st)) != cudaSuccess) return bail(e, "init");
        if ((e = cudaMemcpyAsync(c->stats_prop, init.data(), init.size() * 8, cudaMemcpyHostToDevice, c->st)) != cudaSuccess) return bail(e, "init");
        if ((e = cudaStreamSynchronize(c->st)) != cudaSuccess) return bail(e, "init");
    }
    // a fresh field is all +1 until dqmc_set_conf is called
    if ((e = cudaMemsetAsync(c->conf, 1, (size_t)c->B * c->M * c->N, c->st)) != cudaSuccess) return bail(e, "conf");
    // initialize_stack (stack.jl:186-191)
    if ((e = load_udt(c, c->Ul, c->Dl, c->Tl, -1)) != cudaSuccess) return bail(e, "init stack");
    if ((e = load_udt(c, c->Ur, c->Dr, c->Tr, -1)) != cudaSuccess) return bail(e, "init stack");
    if ((e = cudaStreamSynchronize(c->st)) != cudaSuccess) return bail(e, "init sync");
    *out = c;
    return DQMC_OK;
}


int32_t dqmc_set_conf(dqmc_ctx* c, int32_t chain0, int32_t nchains, const int8_t* conf)
{
    ENTER(c);
    if (!conf || !CHAINS_OK(c, chain0, nchains)) FAIL(c, DQMC_ERR_INVALID, "dqmc_set_conf: bad arguments");
    const size_t per = (size_t)c->M * c->N;
    for (size_t i = 0; i < per * nchains; ++i)
        if (c->ghq ? (conf[i] < 1 || conf[i] > 4) : (conf[i] != 1 && conf[i] != -1))
            FAIL(c, DQMC_ERR_INVALID, c->ghq ? "dqmc_set_conf: conf values must be in 1..4" : "dqmc_set_conf: conf values must be +-1");
    c->generation += 1;
    CK(c, cudaMemcpyAsync(c->conf + per * chain0, conf, per * nchains, cudaMemcpyHostToDevice, c->st));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

int32_t dqmc_get_conf(dqmc_ctx* c, int32_t chain0, int32_t nchains, int8_t* conf)
{
    ENTER(c);
    if (!conf || !CHAINS_OK(c, chain0, nchains)) FAIL(c, DQMC_ERR_INVALID, "dqmc_get_conf: bad arguments");
    const size_t per = (size_t)c->M * c->N;
    CK(c, cudaMemcpyAsync(conf, c->conf + per * chain0, per * nchains, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

// compress / decompress! of the configuration recorder (fields.jl:331-334, configurations.jl): the chunks of
// BitArray(conf .== 1), ceil(N M / 64) UInt64 per chain
static int32_t conf_packed(dqmc_ctx* c, int32_t chain0, int32_t nchains, uint64_t* chunks, int pack)
{
    const long long nval = (long long)c->M * c->N, nbits = nval * (c->ghq ? 2 : 1), wpc = (nbits + 63) / 64;
    unsigned long long* d = nullptr;
    CK(c, cudaMallocAsync((void**)&d, (size_t)wpc * nchains * 8, c->st));
    if (!pack) CK(c, cudaMemcpyAsync(d, chunks, (size_t)wpc * nchains * 8, cudaMemcpyHostToDevice, c->st));
    CK(c, launch_conf_bits(c->conf + nval * chain0, d, nval, nchains, pack, c->ghq ? 1 : 0, c->st));
    if (pack) CK(c, cudaMemcpyAsync(chunks, d, (size_t)wpc * nchains * 8, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaFreeAsync(d, c->st));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

int32_t dqmc_get_conf_packed(dqmc_ctx* c, int32_t chain0, int32_t nchains, uint64_t* chunks)
{
    ENTER(c);
    if (!chunks || !CHAINS_OK(c, chain0, nchains)) FAIL(c, DQMC_ERR_INVALID, "dqmc_get_conf_packed: bad arguments");
    return conf_packed(c, chain0, nchains, chunks, 1);
}

int32_t dqmc_set_conf_packed(dqmc_ctx* c, int32_t chain0, int32_t nchains, const uint64_t* chunks)
{
    ENTER(c);
    if (!chunks || !CHAINS_OK(c, chain0, nchains)) FAIL(c, DQMC_ERR_INVALID, "dqmc_set_conf_packed: bad arguments");
    c->generation += 1;
    return conf_packed(c, chain0, nchains, const_cast<uint64_t*>(chunks), 0);
}

int32_t dqmc_build_stack(dqmc_ctx* c)
{
    ENTER(c);
    CK(c, reverse_build(c));
    CK(c, propagate(c));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

int32_t dqmc_forward_build_stack(dqmc_ctx* c)
{
    ENTER(c);
    CK(c, forward_build(c));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

int32_t dqmc_propagate(dqmc_ctx* c, int32_t n)
{
    ENTER(c);
    for (int i = 0; i < n; ++i) CK(c, propagate(c));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

int32_t dqmc_get_state(const dqmc_ctx* c, int32_t* out3)
{
    if (!c || !out3) return DQMC_ERR_INVALID;
    out3[0] = c->current_slice; out3[1] = c->current_range; out3[2] = c->direction;
    return DQMC_OK;
}

int32_t dqmc_set_sweep_index(dqmc_ctx* c, int64_t s)
{
    ENTER(c);
    c->sweep_index = s;
    const long long v = s;
    CK(c, cudaMemcpyAsync(c->d_sweep_index, &v, sizeof(v), cudaMemcpyHostToDevice, c->st));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

static int32_t fetch_accepted(dqmc_ctx* c, int64_t* accepted)
{
    if (accepted) {
        std::vector<int> h((size_t)c->B);
        CK(c, cudaMemcpyAsync(h.data(), c->accepted, (size_t)c->B * 4, cudaMemcpyDeviceToHost, c->st));
        CK(c, cudaStreamSynchronize(c->st));
        for (int b = 0; b < c->B; ++b) accepted[b] = h[b];
    } else {
        CK(c, cudaStreamSynchronize(c->st));
    }
    return DQMC_OK;
}

static int32_t require_sweep_start(dqmc_ctx* c, const char* who)
{
    if (c->current_slice != 1 || c->direction != 1)
        FAIL(c, DQMC_ERR_INVALID, std::string(who) + ": stack is not at (slice 1, direction +1); call dqmc_build_stack first");
    return DQMC_OK;
}

int32_t dqmc_sweep(dqmc_ctx* c, int32_t nsweeps, const double* uniforms, int64_t* accepted)
{
    ENTER(c);
    if (nsweeps < 0) FAIL(c, DQMC_ERR_INVALID, "dqmc_sweep: nsweeps < 0");
    int32_t rc = require_sweep_start(c, "dqmc_sweep"); if (rc) return rc;
    const size_t per_sweep = (size_t)c->B * 2 * c->M * c->N * (c->ghq ? 2 : 1);
    if (uniforms && !c->d_uniforms) CK(c, dalloc(c, &c->d_uniforms, per_sweep));
    CK(c, cudaMemsetAsync(c->accepted, 0, (size_t)c->B * 4, c->st));
    for (int s = 0; s < nsweeps; ++s) {
        if (uniforms)
            CK(c, cudaMemcpyAsync(c->d_uniforms, uniforms + per_sweep * s, per_sweep * 8, cudaMemcpyHostToDevice, c->st));
        CK(c, local_sweep_graphed(c, uniforms ? c->d_uniforms : nullptr));
    }
    return fetch_accepted(c, accepted);
}

int32_t dqmc_sweep_traced(dqmc_ctx* c, const double* uniforms, const uint8_t* forced, double* probs,
                          uint8_t* decisions, int64_t* accepted)
{
    ENTER(c);
    int32_t rc = require_sweep_start(c, "dqmc_sweep_traced"); if (rc) return rc;
    const size_t per_sweep = (size_t)c->B * 2 * c->M * c->N, per_unif = per_sweep * (c->ghq ? 2 : 1);
    if (uniforms && !c->d_uniforms) CK(c, dalloc(c, &c->d_uniforms, per_unif));
    if (forced && !c->d_forced) CK(c, dalloc(c, &c->d_forced, per_sweep));
    if (probs && !c->d_probs) CK(c, dalloc(c, &c->d_probs, per_sweep));
    if (decisions && !c->d_dec) CK(c, dalloc(c, &c->d_dec, per_sweep));
    if (uniforms) CK(c, cudaMemcpyAsync(c->d_uniforms, uniforms, per_unif * 8, cudaMemcpyHostToDevice, c->st));
    if (forced) CK(c, cudaMemcpyAsync(c->d_forced, forced, per_sweep, cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemsetAsync(c->accepted, 0, (size_t)c->B * 4, c->st));
    CK(c, local_sweep(c, uniforms ? c->d_uniforms : nullptr, forced ? c->d_forced : nullptr,
                      probs ? c->d_probs : nullptr, decisions ? c->d_dec : nullptr));
    if (probs) CK(c, cudaMemcpyAsync(probs, c->d_probs, per_sweep * 8, cudaMemcpyDeviceToHost, c->st));
    if (decisions) CK(c, cudaMemcpyAsync(decisions, c->d_dec, per_sweep, cudaMemcpyDeviceToHost, c->st));
    return fetch_accepted(c, accepted);
}

int32_t dqmc_sweep_spatial(dqmc_ctx* c, const double* uniforms, const uint8_t* forced, double* probs,
                           uint8_t* decisions, int64_t* accepted)
{
    ENTER(c);
    if (c->current_slice < 1 || c->current_slice > c->M) FAIL(c, DQMC_ERR_INVALID, "dqmc_sweep_spatial: no current slice");
    const size_t per = (size_t)c->B * 2 * c->M * c->N;   // reuse the per-sweep buffers
    const size_t cnt = (size_t)c->B * c->N;
    const int uf = c->ghq ? 2 : 1;
    if (uniforms && !c->d_uniforms) CK(c, dalloc(c, &c->d_uniforms, per * uf));
    if (forced && !c->d_forced) CK(c, dalloc(c, &c->d_forced, per));
    if (probs && !c->d_probs) CK(c, dalloc(c, &c->d_probs, per));
    if (decisions && !c->d_dec) CK(c, dalloc(c, &c->d_dec, per));
    if (uniforms) CK(c, cudaMemcpyAsync(c->d_uniforms, uniforms, cnt * uf * 8, cudaMemcpyHostToDevice, c->st));
    if (forced) CK(c, cudaMemcpyAsync(c->d_forced, forced, cnt, cudaMemcpyHostToDevice, c->st));
    CK(c, cudaMemsetAsync(c->accepted, 0, (size_t)c->B * 4, c->st));
    CK(c, sweep_spatial(c, 0, uniforms ? c->d_uniforms : nullptr, (long long)uf * c->N, forced ? c->d_forced : nullptr,
                        probs ? c->d_probs : nullptr, decisions ? c->d_dec : nullptr, c->N));
    if (probs) CK(c, cudaMemcpyAsync(probs, c->d_probs, cnt * 8, cudaMemcpyDeviceToHost, c->st));
    if (decisions) CK(c, cudaMemcpyAsync(decisions, c->d_dec, cnt, cudaMemcpyDeviceToHost, c->st));
    return fetch_accepted(c, accepted);
}

int32_t dqmc_get_stream(dqmc_ctx* c, void** stream)
{
    if (!c || !stream) return DQMC_ERR_INVALID;
    *stream = (void*)c->st;
    return DQMC_OK;
}

int32_t dqmc_profile(dqmc_ctx* c, int32_t enable)
{
    ENTER(c);
    CK(c, cudaStreamSynchronize(c->st));
    c->prof_on = enable != 0;
    if (enable) { c->prof_spans.clear(); c->prof_used = 0; }
    return DQMC_OK;
}

int32_t dqmc_profile_report(dqmc_ctx* c, double* ms, int64_t* count)
{
    ENTER(c);
    if (!ms || !count) FAIL(c, DQMC_ERR_INVALID, "dqmc_profile_report: bad arguments");
    CK(c, cudaStreamSynchronize(c->st));
    for (int i = 0; i < DQMC_PROF_NCAT; ++i) { ms[i] = 0.0; count[i] = 0; }
    for (auto& sp : c->prof_spans) {
        float t = 0.f;
        CK(c, cudaEventElapsedTime(&t, sp.second.first, sp.second.second));
        ms[sp.first] += (double)t; count[sp.first] += 1;
    }
    return DQMC_OK;
}

int32_t dqmc_get_greens(dqmc_ctx* c, int32_t chain0, int32_t nchains, double* G)
{
    ENTER(c);
    if (!G || !CHAINS_OK(c, chain0, nchains)) FAIL(c, DQMC_ERR_INVALID, "dqmc_get_greens: bad arguments");
    CK(c, d2h_mats(c, G, c->greens + (long long)chain0 * c->nb * c->ms, (long long)nchains * c->nb));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

int32_t dqmc_set_greens(dqmc_ctx* c, int32_t chain0, int32_t nchains, const double* G)
{
    ENTER(c);
    if (!G || !CHAINS_OK(c, chain0, nchains)) FAIL(c, DQMC_ERR_INVALID, "dqmc_set_greens: bad arguments");
    CK(c, h2d_mats(c, c->greens + (long long)chain0 * c->nb * c->ms, G, (long long)nchains * c->nb));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

int32_t dqmc_get_measured_greens(dqmc_ctx* c, int32_t chain0, int32_t nchains, double* G)
{
    ENTER(c);
    if (!G || !CHAINS_OK(c, chain0, nchains)) FAIL(c, DQMC_ERR_INVALID, "dqmc_get_measured_greens: bad arguments");
    CK(c, measured_greens(c, c->greens_temp, c->tmp2));
    CK(c, d2h_mats(c, G, c->greens_temp + (long long)chain0 * c->nb * c->ms, (long long)nchains * c->nb));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

// calculate_greens(mc, slice) (stack.jl:525-583)
int32_t dqmc_calculate_greens_at(dqmc_ctx* c, int32_t slice, int32_t safe_mult, double* G)
{
    ENTER(c);
    if (!G || slice < 0 || slice > c->M || safe_mult < 1) FAIL(c, DQMC_ERR_INVALID, "dqmc_calculate_greens_at: bad arguments");
    auto chain = [&](bool dagger, double* U, double* D, double* T) { return build_chain_udt(c, slice, safe_mult, dagger, U, D, T); };
    if (slice + 1 <= c->M) CK(c, chain(true, c->Ur, c->Dr, c->Tr)); else CK(c, load_udt(c, c->Ur, c->Dr, c->Tr, -1));
    if (slice >= 1) CK(c, chain(false, c->Ul, c->Dl, c->Tl)); else CK(c, load_udt(c, c->Ul, c->Dl, c->Tl, -1));
    CK(c, calculate_greens(c, c->greens_temp));
    CK(c, d2h_mats(c, G, c->greens_temp, c->nmat));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

int32_t dqmc_get_stats(dqmc_ctx* c, int32_t chain0, int32_t nchains, dqmc_stats* stats)
{
    ENTER(c);
    if (!stats || !CHAINS_OK(c, chain0, nchains)) FAIL(c, DQMC_ERR_INVALID, "dqmc_get_stats: bad arguments");
    std::vector<double> a((size_t)nchains * 4), b((size_t)nchains * 4);
    CK(c, cudaMemcpyAsync(a.data(), c->stats_neg + 4 * chain0, a.size() * 8, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaMemcpyAsync(b.data(), c->stats_prop + 4 * chain0, b.size() * 8, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaStreamSynchronize(c->st));
    for (int i = 0; i < nchains; ++i) {
        stats[i].neg_count = (int64_t)a[4 * i]; stats[i].neg_sumlog10 = a[4 * i + 1];
        stats[i].neg_min = a[4 * i + 2]; stats[i].neg_max = a[4 * i + 3];
        stats[i].prop_count = (int64_t)b[4 * i]; stats[i].prop_sumlog10 = b[4 * i + 1];
        stats[i].prop_min = b[4 * i + 2]; stats[i].prop_max = b[4 * i + 3];
    }
    return DQMC_OK;
}

int32_t dqmc_get_stack_array(dqmc_ctx* c, int32_t chain, int32_t which, int32_t slot, double* out)
{
    ENTER(c);
    if (!out || chain < 0 || chain >= c->B || which < 0 || which > 8) FAIL(c, DQMC_ERR_INVALID, "dqmc_get_stack_array: bad arguments");
    if (which <= 2 && (slot < 1 || slot > c->C + 1)) FAIL(c, DQMC_ERR_INVALID, "dqmc_get_stack_array: slot out of range");
    const long long mo = (long long)chain * c->nb * c->ms, vo = (long long)chain * c->nb * c->N;
    const double* src = nullptr; bool is_vec = false;
    switch (which) {
    case 0: src = slot_mat(c, c->u_stack, slot - 1) + mo; break;
    case 1: src = slot_vec(c, c->d_stack, slot - 1) + vo; is_vec = true; break;
    case 2: src = slot_mat(c, c->t_stack, slot - 1) + mo; break;
    case 3: src = c->Ul + mo; break;
    case 4: src = c->Dl + vo; is_vec = true; break;
    case 5: src = c->Tl + mo; break;
    case 6: src = c->Ur + mo; break;
    case 7: src = c->Dr + vo; is_vec = true; break;
    default: src = c->Tr + mo; break;
    }
    if (is_vec) CK(c, cudaMemcpyAsync(out, src, (size_t)c->nb * c->N * 8, cudaMemcpyDeviceToHost, c->st));
    else CK(c, d2h_mats(c, out, src, c->nb));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

// ---- observables -------------------------------------------------------------------------------
__global__ void obs_accumulate_kernel(const double* G, long long chain_stride, int n_chains, double* obs,
                                      long long per_chain)
{
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < per_chain;
         e += (long long)gridDim.x * blockDim.x) {
        double s = 0.0, s2 = 0.0;
        for (int b = 0; b < n_chains; ++b) { const double x = G[b * chain_stride + e]; s += x; s2 += x * x; }
        obs[1 + e] += s; obs[1 + per_chain + e] += s2;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) obs[0] += (double)n_chains;
}

int32_t dqmc_accumulate_greens(dqmc_ctx* c)
{
    ENTER(c);
    CK(c, measured_greens(c, c->greens_temp, c->tmp2));
    const long long per = (long long)c->nb * c->ms;
    long long blocks = (per + 255) / 256; if (blocks > 592) blocks = 592;
    obs_accumulate_kernel<<<(unsigned)blocks, 256, 0, c->st>>>(c->greens_temp, per, c->B, c->obs, per);
    count_launch();
    CK(c, cudaGetLastError());
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

int32_t dqmc_observable_buffer(dqmc_ctx* c, void** device_ptr, int64_t* n_doubles)
{
    if (!c || !device_ptr || !n_doubles) return DQMC_ERR_INVALID;
    *device_ptr = c->obs; *n_doubles = c->obs_len;
    return DQMC_OK;
}

// ---- NCCL, resolved at run time (the library does not link libnccl) ---------------------------------------
struct nccl_uid { char internal[128]; };                 // ncclUniqueId (nccl.h)
struct NcclApi {
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(void**, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
static const NcclApi& nccl_api()
{
    static const NcclApi api = [] {
        NcclApi a;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return a;
        a.GetUniqueId = (int (*)(nccl_uid*))dlsym(h, "ncclGetUniqueId");
        a.CommInitRank = (int (*)(void**, int, nccl_uid, int))dlsym(h, "ncclCommInitRank");
        a.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
        a.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
        a.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce;
        return a;
    }();
    return api;
}

int32_t dqmc_comm_unique_id(uint8_t* id128)
{
    if (!id128) { g_create_error = "dqmc_comm_unique_id: null argument"; return DQMC_ERR_INVALID; }
    const NcclApi& n = nccl_api();
    if (!n.ok) { g_create_error = "libnccl.so.2 not loadable"; return DQMC_ERR_UNSUPPORTED; }
    nccl_uid u;
    if (n.GetUniqueId(&u) != 0) { g_create_error = "ncclGetUniqueId failed"; return DQMC_ERR_CUDA; }
    memcpy(id128, u.internal, 128);
    return DQMC_OK;
}

int32_t dqmc_comm_init(dqmc_ctx* c, int32_t n_ranks, int32_t rank, const uint8_t* id128)
{
    ENTER(c);
    if (!id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) FAIL(c, DQMC_ERR_INVALID, "dqmc_comm_init: bad arguments");
    const NcclApi& n = nccl_api();
    if (!n.ok) FAIL(c, DQMC_ERR_UNSUPPORTED, "dqmc_comm_init: libnccl.so.2 not loadable");
    if (c->nccl_comm) { n.CommDestroy(c->nccl_comm); c->nccl_comm = nullptr; }
    nccl_uid u;
    memcpy(u.internal, id128, 128);
    const int rc = n.CommInitRank(&c->nccl_comm, n_ranks, u, rank);
    if (rc != 0) {
        c->nccl_comm = nullptr;
        FAIL(c, DQMC_ERR_CUDA, std::string("ncclCommInitRank: ") + (n.GetErrorString ? n.GetErrorString(rc) : "failed"));
    }
    return DQMC_OK;
}

int32_t dqmc_comm_destroy(dqmc_ctx* c)
{
    if (!c) return DQMC_ERR_INVALID;
    if (c->nccl_comm) { nccl_api().CommDestroy(c->nccl_comm); c->nccl_comm = nullptr; }
    return DQMC_OK;
}

int32_t dqmc_reduce_observables(dqmc_ctx* c, void* comm)
{
    ENTER(c);
    if (!comm) comm = c->nccl_comm;
    if (!comm) FAIL(c, DQMC_ERR_INVALID, "dqmc_reduce_observables: no communicator (pass one or call dqmc_comm_init)");
    const NcclApi& n = nccl_api();
    if (!n.ok) FAIL(c, DQMC_ERR_UNSUPPORTED, "dqmc_reduce_observables: libnccl.so.2 not loadable");
    // ncclDouble = 8, ncclSum = 0 (nccl.h); issued on the context's stream, i.e. ordered after the accumulation kernels
    if (n.AllReduce(c->obs, c->obs, (size_t)c->obs_len, 8, 0, comm, c->st) != 0) FAIL(c, DQMC_ERR_CUDA, "ncclAllReduce failed");
    {   // the Wick-kernel accumulators of measure.cu ride along
        void* mptr = nullptr; int64_t mlen = 0;
        if (c->meas && dqmc_measurement_buffer(c, &mptr, &mlen) == DQMC_OK && mlen > 0)
            if (n.AllReduce(mptr, mptr, (size_t)mlen, 8, 0, comm, c->st) != 0) FAIL(c, DQMC_ERR_CUDA, "ncclAllReduce failed");
        // ... and their log-binning levels ({count, sum, sum of squares} per level)
        if (c->meas && dqmc_measurement_binning_buffer(c, &mptr, &mlen) == DQMC_OK && mlen > 0)
            if (n.AllReduce(mptr, mptr, (size_t)mlen, 8, 0, comm, c->st) != 0) FAIL(c, DQMC_ERR_CUDA, "ncclAllReduce failed");
    }
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

int32_t dqmc_get_observables(dqmc_ctx* c, double* count, double* sum, double* sumsq)
{
    ENTER(c);
    if (!count || !sum || !sumsq) FAIL(c, DQMC_ERR_INVALID, "dqmc_get_observables: bad arguments");
    const long long per = (long long)c->nb * c->ms;
    CK(c, cudaMemcpyAsync(count, c->obs, 8, cudaMemcpyDeviceToHost, c->st));
    CK(c, d2h_mats(c, sum, c->obs + 1, c->nb));
    CK(c, d2h_mats(c, sumsq, c->obs + 1 + per, c->nb));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

// ---- operator level ---------------------------------------------------------------------------
// A throw-away context that only carries geometry + scratch for `batch` matrices of size n.
static int32_t make_op_ctx(int device, int n, int batch, dqmc_ctx** out)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { g_create_error = "no CUDA device"; return DQMC_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev || n < 1 || batch < 1) { g_create_error = "bad arguments"; return DQMC_ERR_INVALID; }
    if (n > udt_max_n()) { g_create_error = "n too large"; return DQMC_ERR_UNSUPPORTED; }
    cudaSetDevice(device);
    dqmc_ctx* c = new dqmc_ctx();
    t_launch_counter = &c->launches;
    c->N = n; c->M = 1; c->nb = 1; c->B = batch; c->C = 1; c->device = device;
    c->ld = (n + 1) & ~1; c->ms = (long long)c->ld * n; c->nmat = batch; c->ldv = ((n + 31) / 32) * 32;
    cudaError_t e = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking);
    const size_t mat = (size_t)c->nmat * c->ms, vec = (size_t)c->nmat * n;
    if (e == cudaSuccess) e = dalloc(c, &c->greens, mat);
    if (e == cudaSuccess) e = dalloc(c, &c->Ul, mat);
    if (e == cudaSuccess) e = dalloc(c, &c->Ur, mat);
    if (e == cudaSuccess) e = dalloc(c, &c->Tl, mat);
    if (e == cudaSuccess) e = dalloc(c, &c->Tr, mat);
    if (e == cudaSuccess) e = dalloc(c, &c->tmp1, mat);
    if (e == cudaSuccess) e = dalloc(c, &c->Vwork, (size_t)c->nmat * c->ldv * n);
    if (e == cudaSuccess) e = dalloc(c, &c->Dl, vec);
    if (e == cudaSuccess) e = dalloc(c, &c->Dr, vec);
    if (e == cudaSuccess) e = dalloc(c, &c->tau, vec);
    if (e == cudaSuccess) e = dalloc(c, &c->Dgreens, vec);
    if (e == cudaSuccess) e = dalloc(c, &c->pivot, vec);
    if (e == cudaSuccess) e = dalloc(c, &c->udt_scratch, (size_t)c->nmat * udt_reg_scratch_doubles(n, c->ld));
    if (e == cudaSuccess) e = dalloc(c, &c->udt_iscratch, (size_t)c->nmat * udt_reg_scratch_ints(n));
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); dqmc_destroy(c); return DQMC_ERR_CUDA; }
    *out = c;
    return DQMC_OK;
}

#define OPCK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    g_create_error = std::string(#call) + ": " + cudaGetErrorString(e__); dqmc_destroy(c); return DQMC_ERR_CUDA; } } while (0)

int32_t dqmc_op_vmul(int32_t device, int32_t n, int32_t batch, int32_t tA, int32_t tB, const double* A,
                     const double* Bm, double* Cm)
{
    dqmc_ctx* c = nullptr;
    int32_t rc = make_op_ctx(device, n, batch, &c); if (rc) return rc;
    OPCK(h2d_mats(c, c->Ul, A, batch));
    OPCK(h2d_mats(c, c->Ur, Bm, batch));
    OPCK(mm(c, c->greens, c->Ul, tA != 0, false, c->Ur, tB != 0, false));
    OPCK(d2h_mats(c, Cm, c->greens, batch));
    OPCK(cudaStreamSynchronize(c->st));
    dqmc_destroy(c);
    return DQMC_OK;
}

int32_t dqmc_op_udt(int32_t device, int32_t n, int32_t batch, int32_t apply_pivot, const double* X, double* U,
                    double* D, double* T, int64_t* pivot)
{
    dqmc_ctx* c = nullptr;
    int32_t rc = make_op_ctx(device, n, batch, &c); if (rc) return rc;
    OPCK(h2d_mats(c, c->greens, X, batch));
    OPCK(udt(c, c->greens, no_scale(), c->Ul, c->Dl, c->Tl, apply_pivot != 0));
    OPCK(d2h_mats(c, U, c->Ul, batch));
    OPCK(d2h_mats(c, T, c->Tl, batch));
    OPCK(cudaMemcpyAsync(D, c->Dl, (size_t)batch * n * 8, cudaMemcpyDeviceToHost, c->st));
    std::vector<int> hp((size_t)batch * n);
    OPCK(cudaMemcpyAsync(hp.data(), c->pivot, hp.size() * 4, cudaMemcpyDeviceToHost, c->st));
    OPCK(cudaStreamSynchronize(c->st));
    if (pivot) for (size_t i = 0; i < hp.size(); ++i) pivot[i] = (int64_t)hp[i] + 1;
    dqmc_destroy(c);
    return DQMC_OK;
}

int32_t dqmc_op_rdivp(int32_t device, int32_t n, int32_t batch, double* A, const double* T, const int64_t* pivot)
{
    dqmc_ctx* c = nullptr;
    int32_t rc = make_op_ctx(device, n, batch, &c); if (rc) return rc;
    std::vector<int> hp((size_t)batch * n);
    for (size_t i = 0; i < hp.size(); ++i) {
        if (pivot[i] < 1 || pivot[i] > n) { g_create_error = "pivot out of range"; dqmc_destroy(c); return DQMC_ERR_INVALID; }
        hp[i] = (int)(pivot[i] - 1);
    }
    OPCK(h2d_mats(c, c->Ur, A, batch));
    OPCK(h2d_mats(c, c->Tr, T, batch));
    OPCK(cudaMemcpyAsync(c->pivot, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice, c->st));
    OPCK(rdivp(c, c->Ur, c->Tr, c->tmp1));
    OPCK(d2h_mats(c, A, c->Ur, batch));
    OPCK(cudaStreamSynchronize(c->st));
    dqmc_destroy(c);
    return DQMC_OK;
}

int32_t dqmc_op_calculate_greens(int32_t device, int32_t n, int32_t batch, const double* Ul, const double* Dl,
                                 const double* Tl, const double* Ur, const double* Dr, const double* Tr, double* G)
{
    dqmc_ctx* c = nullptr;
    int32_t rc = make_op_ctx(device, n, batch, &c); if (rc) return rc;
    OPCK(h2d_mats(c, c->Ul, Ul, batch)); OPCK(h2d_mats(c, c->Tl, Tl, batch));
    OPCK(h2d_mats(c, c->Ur, Ur, batch)); OPCK(h2d_mats(c, c->Tr, Tr, batch));
    OPCK(cudaMemcpyAsync(c->Dl, Dl, (size_t)batch * n * 8, cudaMemcpyHostToDevice, c->st));
    OPCK(cudaMemcpyAsync(c->Dr, Dr, (size_t)batch * n * 8, cudaMemcpyHostToDevice, c->st));
    OPCK(calculate_greens(c, c->greens));
    OPCK(d2h_mats(c, G, c->greens, batch));
    OPCK(cudaStreamSynchronize(c->st));
    dqmc_destroy(c);
    return DQMC_OK;
}

int32_t dqmc_op_multiply_slice_matrix(dqmc_ctx* c, int32_t which, int32_t slice, double* X)
{
    ENTER(c);
    if (!X || slice < 1 || slice > c->M || which < 0 || which > 4) FAIL(c, DQMC_ERR_INVALID, "dqmc_op_multiply_slice_matrix: bad arguments");
    CK(c, h2d_mats(c, c->tmp2, X, c->nmat));
    switch (which) {
    case 0: CK(c, slice_left(c, c->greens_temp, c->tmp2, slice)); break;
    case 1: CK(c, slice_right(c, c->greens_temp, c->tmp2, slice)); break;
    case 2: CK(c, slice_inv_right(c, c->greens_temp, c->tmp2, slice)); break;
    case 3: CK(c, slice_inv_left(c, c->greens_temp, c->tmp2, slice)); break;
    default: CK(c, slice_daggered_left(c, c->greens_temp, c->tmp2, slice)); break;
    }
    CK(c, d2h_mats(c, X, c->greens_temp, c->nmat));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

int32_t dqmc_op_wrap_greens(dqmc_ctx* c, int32_t curr_slice, int32_t direction, double* X)
{
    ENTER(c);
    const int sl = (direction == -1) ? curr_slice - 1 : curr_slice;
    if (!X || (direction != 1 && direction != -1) || sl < 1 || sl > c->M) FAIL(c, DQMC_ERR_INVALID, "dqmc_op_wrap_greens: bad arguments");
    CK(c, h2d_mats(c, c->greens_temp, X, c->nmat));
    CK(c, wrap_greens(c, c->greens_temp, c->tmp2, curr_slice, direction));
    CK(c, d2h_mats(c, X, c->greens_temp, c->nmat));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

}  // extern "C"
