/*
 * dqmc_b200.h -- C ABI of the B200-native DQMC sweep library (libdqmc_b200.so).
 *
 * The reference (carstenbauer/MonteCarlo.jl) is pure Julia and has no FFI for this path;
 * the seam is Julia dispatch.  The entry points below are what a `ccall` from the
 * Julia glue (julia/GPULocalSweep.jl) binds, one per reference
 * interface it replaces.  All paths are relative to /root/reference.
 *
 * Conventions
 *  - every call returns int32 status: 0 = ok, < 0 = error (text via dqmc_last_error);
 *    no exceptions or callbacks cross the ABI (maps to Julia error()/ExitCode,
 *    src/helpers.jl:17-22).
 *  - the caller owns every host buffer; the library copies during the call and never
 *    retains host pointers.  The library owns all device memory and its stream.
 *  - matrices are dense column-major double with leading dimension N (Julia Matrix{Float64});
 *    BlockDiagonal Green's functions are passed as N x N x n_flavors; batches of chains as
 *    one more trailing dimension.  conf is Int8 N x M per chain (field.conf,
 *    src/flavors/DQMC/fields.jl:363-368), values +-1, [site, slice].
 *  - slice / range indices crossing the ABI are 1-based like the reference's.
 *  - a context is not thread-safe; contexts are independent (one per GPU / process).
 *    Calls are synchronous: results are complete when the call returns.
 */
#ifndef DQMC_B200_H
#define DQMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DQMC_OK 0
#define DQMC_ERR_INVALID (-1)
#define DQMC_ERR_CUDA (-2)
#define DQMC_ERR_UNSUPPORTED (-3)
#define DQMC_ERR_NO_DEVICE (-4)

#define DQMC_FIELD_DENSITY_HIRSCH 0   /* DensityHirschField,  fields.jl:363-395, 1 flavor block  */
#define DQMC_FIELD_MAGNETIC_HIRSCH 1  /* MagneticHirschField, fields.jl:412-451, 2 flavor blocks */
#define DQMC_FIELD_DENSITY_GHQ 2      /* DensityGHQField,     fields.jl:565-610, 1 flavor block, conf in 1..4 */
#define DQMC_FIELD_MAGNETIC_GHQ 3     /* MagneticGHQField,    fields.jl:503-556, 2 flavor blocks, conf in 1..4 */
/* Gauss-Hermite fields: alpha = sqrt(+-dtau U / 2) must be real (DensityGHQ: U > 0, MagneticGHQ: U < 0; the complex
 * case is out of scope like every complex matrix type).  A proposal draws x_new = choices[x_old, rand(1:3)] first and
 * the Metropolis uniform afterwards: both are defined by include/dqmc_rng.h (dqmc_uniform_choice, dqmc_ghq_choice,
 * dqmc_uniform); explicit uniform tables carry the choice uniforms as a second block per slice visit (see dqmc_sweep).
 * gamma(x) / eta(x): dqmc_ghq_tables. */

typedef struct dqmc_ctx dqmc_ctx;

/* What `init!(mc, ::GPULocalSweep)` hands over: DQMCParameters (parameters.jl:33-49),
 * the stack's ranges and hopping exponentials (stack.jl:154-158, 235-239) and the field's
 * coupling (fields.jl:370-376, 419-425). */
typedef struct dqmc_desc {
    int32_t n_sites;                  /* N = length(lattice)                                    */
    int32_t n_slices;                 /* M = parameters.slices                                  */
    int32_t field_kind;               /* DQMC_FIELD_*  (n_flavors = 1 density / 2 magnetic)     */
    int32_t n_chains;                 /* independent Markov chains batched in this context      */
    int32_t n_ranges;                 /* C = length(stack.ranges)                               */
    const int32_t* range_first;       /* [C] first(stack.ranges[i]), 1-based                    */
    const int32_t* range_last;        /* [C] last(stack.ranges[i]), 1-based inclusive           */
    double alpha;                     /* field.alpha                                            */
    const double* hopping_exp_squared;     /* exp(-dtau T)    N x N  stack.hopping_matrix_exp_squared     */
    const double* hopping_exp_inv_squared; /* exp(+dtau T)           stack.hopping_matrix_exp_inv_squared */
    const double* hopping_exp;             /* exp(-dtau T/2)         stack.hopping_matrix_exp             */
    const double* hopping_exp_inv;         /* exp(+dtau T/2)         stack.hopping_matrix_exp_inv         */
    int32_t check_sign_problem;       /* parameters.check_sign_problem      (parameters.jl:38)  */
    int32_t check_propagation_error;  /* parameters.check_propagation_error (parameters.jl:39)  */
    uint64_t seed;                    /* key of the counter RNG (include/dqmc_rng.h)            */
    int64_t chain_offset;             /* global index of chain 0 (multi-GPU sharding)           */
    int32_t device;                   /* CUDA device ordinal                                    */
    int32_t delay_block;              /* sites per delayed-update block, 0 = auto               */
    int32_t update_variant;           /* sweep_spatial kernel: 0 = auto (by lattice size), 1 = delayed rank-k factors
                                         (update.cu), 3 = submatrix form (update3.cu); same decisions either way   */
} dqmc_desc;

/* MagnitudeStats (src/flavors/DQMC/statistics.jl:9-38) per chain. */
typedef struct dqmc_stats {
    int64_t neg_count;  double neg_sumlog10, neg_min, neg_max;     /* negative_probability  */
    int64_t prop_count; double prop_sumlog10, prop_min, prop_max;  /* propagation_error     */
} dqmc_stats;

/* ---- lifetime -------------------------------------------------------------------------- */
/* DQMCStack + initialize_stack + init_hopping_matrices upload (stack.jl:1-74, 160-249). */
int32_t dqmc_create(const dqmc_desc* desc, dqmc_ctx** out);
int32_t dqmc_destroy(dqmc_ctx* ctx);
/* message of the last failing call on this context (or of dqmc_create when ctx == NULL). */
const char* dqmc_last_error(const dqmc_ctx* ctx);
int32_t dqmc_device_count(void);

/* ---- field configuration: mc.field.conf (fields.jl:363-368) -------------------------------- */
int32_t dqmc_set_conf(dqmc_ctx* ctx, int32_t chain0, int32_t nchains, const int8_t* conf);
int32_t dqmc_get_conf(dqmc_ctx* ctx, int32_t chain0, int32_t nchains, int8_t* conf);
/* compress(field) = BitArray(conf .== 1) / decompress!(field, bits) (fields.jl:331-334), the wire format of the
 * ConfigRecorder (configurations.jl:92-200): chunks = the UInt64 words of the BitArray (bit i of the column-major
 * N x M array at chunks[i >> 6], position i & 63), ceil(N M / 64) words per chain, packed on the device.
 * GHQ fields (fields.jl:476-489): two bits per value, (v - 1) >> 1 at bit 2 i and (v - 1) & 1 at bit 2 i + 1,
 * ceil(2 N M / 64) words per chain. */
int32_t dqmc_get_conf_packed(dqmc_ctx* ctx, int32_t chain0, int32_t nchains, uint64_t* chunks);
int32_t dqmc_set_conf_packed(dqmc_ctx* ctx, int32_t chain0, int32_t nchains, const uint64_t* chunks);

/* ---- stack ------------------------------------------------------------------------------- */
/* reverse_build_stack + propagate (stack.jl:284-308, 605; DQMC.jl:178-179): afterwards
 * current_slice = 1, direction = +1 and greens = G_eff(1). */
int32_t dqmc_build_stack(dqmc_ctx* ctx);
/* build_stack (stack.jl:257-281): current_slice = M + 1, direction = -1. */
int32_t dqmc_forward_build_stack(dqmc_ctx* ctx);
/* propagate (stack.jl:605-730), n times. */
int32_t dqmc_propagate(dqmc_ctx* ctx, int32_t n);
/* [current_slice, current_range, direction] */
int32_t dqmc_get_state(const dqmc_ctx* ctx, int32_t* out3);

/* ---- the sweep ----------------------------------------------------------------------------- */
/* update(::LocalSweep, mc, model, field) = local_sweep (local_updates.jl:7-14, 82), nsweeps times.
 * uniforms: NULL -> counter RNG (dqmc_rng.h); else host table [nsweeps][n_chains][2M][N] of the
 * Metropolis uniforms (Hirsch fields) or [nsweeps][n_chains][2M][2][N] (GHQ fields: per slice visit the N Metropolis
 * uniforms, then the N choice uniforms u that dqmc_rng.h's dqmc_ghq_choice maps to x_new).
 * accepted: [n_chains] accepted flips summed over the nsweeps. */
int32_t dqmc_sweep(dqmc_ctx* ctx, int32_t nsweeps, const double* uniforms, int64_t* accepted);
/* One sweep with optional teacher forcing and traces, each [n_chains][2M][N]:
 * forced != NULL replays the given accept decisions; probs / decisions (may be NULL) receive
 * p = exp(-dE_boson) * detratio (local_updates.jl:31) and the decision of every proposal. */
int32_t dqmc_sweep_traced(dqmc_ctx* ctx, const double* uniforms, const uint8_t* forced,
                          double* probs, uint8_t* decisions, int64_t* accepted);
/* sweep_spatial at the current slice only (local_updates.jl:23-60); arrays [n_chains][N]. */
int32_t dqmc_sweep_spatial(dqmc_ctx* ctx, const double* uniforms, const uint8_t* forced,
                           double* probs, uint8_t* decisions, int64_t* accepted);
int32_t dqmc_set_sweep_index(dqmc_ctx* ctx, int64_t sweep);

/* ---- global updates (src/flavors/DQMC/updates/global_updates.jl) ----------------------------------------- */
/* global_update (:203-219) for every chain: proposed = [n_chains][M][N] configurations (propose_conf! done by the
 * caller: GlobalShuffle, SpatialShuffle, ...) or NULL for GlobalFlip (:237-248, conf -> -conf; Hirsch fields only).
 * GHQ fields: like the reference, the ratio carries no gamma(x) factors (propose_global_from_conf has no GHQ method,
 * :137-179), which is exact for the shuffle updates (they permute the values).  The weight ratio
 * is det(G_old) / det(G_new) from the diagonal factors (inv_det :70-137, propose_global_from_conf :147-179);
 * accepted chains keep the proposal, rejected ones their old configuration; afterwards the stack is rebuilt
 * (accept_global! :181-198) and sits at (slice 1, direction +1).  uniforms: [n_chains] or NULL: counter RNG
 * (dqmc_rng.h) with sweep = the sweep index, step = 2M and site = the running index of the global update on this
 * context (0, 1, 2, ...: two global updates between local sweeps never share a uniform).
 * accepted / probs: [n_chains], may be NULL; probs = |exp(-dE_boson) detratio|. */
int32_t dqmc_global_update(dqmc_ctx* ctx, const int8_t* proposed, const double* uniforms, int32_t safe_mult,
                           int64_t* accepted, double* probs);
/* running index of the next global update (checkpoint / resume of the counter RNG, like dqmc_set_sweep_index). */
int32_t dqmc_set_global_update_index(dqmc_ctx* ctx, int64_t index);

/* ---- results ------------------------------------------------------------------------------- */
/* mc.stack.greens (effective Green's function), N x N x n_flavors per chain. */
int32_t dqmc_get_greens(dqmc_ctx* ctx, int32_t chain0, int32_t nchains, double* G);
int32_t dqmc_set_greens(dqmc_ctx* ctx, int32_t chain0, int32_t nchains, const double* G);
/* greens!(mc) = exp(+dtau T/2) G_eff exp(-dtau T/2) (greens.jl:94-125). */
int32_t dqmc_get_measured_greens(dqmc_ctx* ctx, int32_t chain0, int32_t nchains, double* G);
/* calculate_greens(mc, slice) from scratch (stack.jl:525-583), all chains; invalidates Ul..Tr. */
int32_t dqmc_calculate_greens_at(dqmc_ctx* ctx, int32_t slice, int32_t safe_mult, double* G);
int32_t dqmc_get_stats(dqmc_ctx* ctx, int32_t chain0, int32_t nchains, dqmc_stats* stats);
/* which: 0 u_stack[slot], 1 d_stack[slot], 2 t_stack[slot], 3 Ul, 4 Dl, 5 Tl, 6 Ur, 7 Dr, 8 Tr
 * (stack.jl:13-22); matrices N x N x n_flavors, vectors N x n_flavors; slot is 1-based. */
int32_t dqmc_get_stack_array(dqmc_ctx* ctx, int32_t chain, int32_t which, int32_t slot, double* out);

/* ---- unequal-time Green's functions (UnequalTimeStack, src/flavors/DQMC/unequal_time_stack.jl) ------- */
/* build_stack(mc, mc.ut_stack) (:128-185): forward, backward and inverse UDT stacks from the current conf. */
int32_t dqmc_ut_build_stack(dqmc_ctx* ctx);
/* lazy_build_forward!(mc, s, upto) / lazy_build_backward!(mc, s, downto) (:207-270); 0 skips that side. */
int32_t dqmc_ut_lazy_build(dqmc_ctx* ctx, int32_t forward_upto, int32_t backward_downto);
/* measured != 0: greens(mc, k, l) (:302-320), G(k <- l) with the exp(+-dtau T/2) transform;
 * measured == 0: calculate_greens(mc, k, l) (:322-335), the effective G(k, l).  0 <= k, l <= M; all chains,
 * N x N x n_flavors x n_chains.  Overwrites Ul..Tr like the reference and ends a running iteration. */
int32_t dqmc_ut_greens(dqmc_ctx* ctx, int32_t k, int32_t l, int32_t measured, double* G);
/* which: 0/1/2 forward_u/d/t_stack, 3/4/5 backward_u/d/t_stack, 6/7/8 inv_u/d/t_stack; slot 1-based. */
int32_t dqmc_ut_get_stack_array(dqmc_ctx* ctx, int32_t chain, int32_t which, int32_t slot, double* out);
/* CombinedGreensIterator(mc; recalculate, start, stop) (measurements/greens_iterators.jl:154-435): every
 * dqmc_cgi_next produces the triple (G(0,l), G(l,0), G(l,l)) of the next l (measured Green's functions of all
 * chains; any of the three pointers may be NULL to leave that matrix on the device) and stores l, or -1 once
 * l > stop.  The stack must be at (slice 1, direction +1) or G(0,0) is recomputed from the ut stack. */
int32_t dqmc_cgi_begin(dqmc_ctx* ctx, int32_t recalculate, int32_t start, int32_t stop, int32_t safe_mult);
int32_t dqmc_cgi_next(dqmc_ctx* ctx, int32_t* l, double* G0l, double* Gl0, double* Gll);

/* ---- observables ---------------------------------------------------------------------------- */
/* Accumulate the measured Green's function of every chain into device accumulators
 * {count, sum, sum of squares} (stands in for push!(LogBinner, G), measurements/generic.jl:586). */
int32_t dqmc_accumulate_greens(dqmc_ctx* ctx);
/* Device pointer / element count of the accumulator block [count | sum | sumsq] so that the host
 * (torch.distributed / NCCL.jl) can all-reduce it in place over NVLink. */
int32_t dqmc_observable_buffer(dqmc_ctx* ctx, void** device_ptr, int64_t* n_doubles);
/* all-reduce (sum) of the accumulator blocks over an NCCL communicator, on the context's stream (ordered after the
 * accumulation kernels).  nccl_comm: an existing ncclComm_t owned by the caller, or NULL for the communicator created
 * by dqmc_comm_init.  NCCL is resolved at run time (dlopen libnccl.so.2). */
int32_t dqmc_reduce_observables(dqmc_ctx* ctx, void* nccl_comm);
/* A communicator owned by the context, for hosts without NCCL bindings of their own (the reference's parallel runs
 * are one process per simulation, test/parallel.jl:62): rank 0 calls dqmc_comm_unique_id (ncclGetUniqueId, 128 bytes)
 * and ships the bytes to every rank by any means (MPI.jl bcast, a file, torch.distributed); every rank then calls
 * dqmc_comm_init (ncclCommInitRank on the context's device -- collective over the n_ranks contexts). */
int32_t dqmc_comm_unique_id(uint8_t* id128);
int32_t dqmc_comm_init(dqmc_ctx* ctx, int32_t n_ranks, int32_t rank, const uint8_t* id128);
int32_t dqmc_comm_destroy(dqmc_ctx* ctx);
/* copy out: count, then mean and variance-of-the-mean inputs (sum, sumsq), N x N x n_flavors each. */
int32_t dqmc_get_observables(dqmc_ctx* ctx, double* count, double* sum, double* sumsq);

/* ---- device-side Wick kernels (src/flavors/DQMC/measurements) ---------------------------------------- */
/* Lattice tables the kernels need: Bravais srctrg2dir (lattices/lattice_cache.jl:69-78, 224-240; n_bravais x
 * n_bravais, column-major [src, trg], 1-based directions), the hopping matrix mc.stack.hopping_matrix (N x N)
 * for the kinetic energy and the model's U (HubbardModel.jl:160-183).  Site = cell + n_bravais * basis. */
int32_t dqmc_set_lattice(dqmc_ctx* ctx, int32_t n_bravais, int32_t n_basis, const int32_t* srctrg2dir,
                         const double* hopping_matrix, double U);
/* Observables, in this order inside a result vector (offsets[k] .. offsets[k+1], offsets[DQMC_OBS_COUNT] = length):
 * occupation (N x n_flavors, occupation.jl:44-70), kinetic / interaction / total energy (energy.jl:119-165),
 * equal-time charge and spin x/y/z density correlations (n_bravais x n_basis x n_basis each, EachSitePairByDistance,
 * full_cdc_kernel / full_sdc_*_kernel), and their time-integrated susceptibilities (TimeIntegral). */
#define DQMC_OBS_OCC 0
#define DQMC_OBS_KINETIC 1
#define DQMC_OBS_INTERACTION 2
#define DQMC_OBS_TOTAL_ENERGY 3
#define DQMC_OBS_CDC 4
#define DQMC_OBS_SDC_X 5
#define DQMC_OBS_SDC_Y 6
#define DQMC_OBS_SDC_Z 7
#define DQMC_OBS_CDS 8
#define DQMC_OBS_SDS_X 9
#define DQMC_OBS_SDS_Y 10
#define DQMC_OBS_SDS_Z 11
#define DQMC_OBS_COUNT 12
int32_t dqmc_measurement_layout(dqmc_ctx* ctx, int32_t* offsets /* [DQMC_OBS_COUNT + 1] */);
/* apply!(::Greens, ...) (measurements/generic.jl:287-310) for every chain: observables 0..7 from greens!(mc);
 * the stack must be at (slice 1, direction +1) like the reference's measurement point (DQMC.jl:217). */
int32_t dqmc_measure_equal_time(dqmc_ctx* ctx);
/* apply!(::TimeIntegral, ...) (generic.jl:337-372): runs the CombinedGreensIterator on the device and sums
 * weight_l * kernel(G00, G0l, Gl0, Gll) into observables 8..11 (weight = dtau, halved at l = 0 and l = M). */
int32_t dqmc_measure_time_integral(dqmc_ctx* ctx, int32_t recalculate, int32_t safe_mult, double delta_tau);
/* values of the last measurement, [nchains][length] (what each chain's LogBinner would be pushed). */
int32_t dqmc_get_measurements(dqmc_ctx* ctx, int32_t chain0, int32_t nchains, double* out);
/* device accumulator block [count_equal_time, count_time_integral | sum[length] | sumsq[length]] over chains and
 * calls; dqmc_reduce_observables all-reduces it together with the Green's function block. */
int32_t dqmc_measurement_buffer(dqmc_ctx* ctx, void** device_ptr, int64_t* n_doubles);
int32_t dqmc_get_measurement_stats(dqmc_ctx* ctx, double* counts /* [2] */, double* sum, double* sumsq);
/* Log-binning of every observable element (the LogBinner of BinningAnalysis.jl 0.6 that backs a DQMCMeasurement,
 * measurements/generic.jl:62-65, 92-95, 586-587): level 0 accumulates every measured value, level l + 1 the mean of two
 * successive values of level l, per chain; {count, sum, sum of squares} of each level are summed over the chains of the
 * context (and over ranks by dqmc_reduce_observables).  L = dqmc_binning_levels().  counts: [2][L] (equal time, time
 * integral), sum / sumsq: [L][length].  std_error(level l) = sqrt((sumsq/count - (sum/count)^2) / (count - 1)); its growth
 * with l is the autocorrelation-aware error estimate.  (Third-party arithmetic: parity unpinned, see DESIGN.md.) */
int32_t dqmc_binning_levels(void);
int32_t dqmc_get_measurement_binning(dqmc_ctx* ctx, double* counts, double* sum, double* sumsq);
int32_t dqmc_measurement_binning_buffer(dqmc_ctx* ctx, void** device_ptr, int64_t* n_doubles);

/* ---- operator level: the reference's linalg "operator API", batched over host arrays ------- */
/* vmul!(C, op(A), op(B)) (linalg/real.jl:7-15, 72-102) */
int32_t dqmc_op_vmul(int32_t device, int32_t n, int32_t batch, int32_t transA, int32_t transB,
                     const double* A, const double* B, double* C);
/* udt_AVX_pivot!(U, D, T, pivot, temp, Val(apply_pivot)) (linalg/UDT.jl:216-334); X is not modified,
 * pivot is 1-based on return. */
int32_t dqmc_op_udt(int32_t device, int32_t n, int32_t batch, int32_t apply_pivot, const double* X,
                    double* U, double* D, double* T, int64_t* pivot);
/* rdivp!(A, T, O, pivot) (linalg/real.jl:198-226); pivot 1-based, A overwritten. */
int32_t dqmc_op_rdivp(int32_t device, int32_t n, int32_t batch, double* A, const double* T,
                      const int64_t* pivot);
/* calculate_greens_AVX!(Ul, Dl, Tl, Ur, Dr, Tr, G) (stack.jl:442-496); inputs are not modified. */
int32_t dqmc_op_calculate_greens(int32_t device, int32_t n, int32_t batch, const double* Ul,
                                 const double* Dl, const double* Tl, const double* Ur,
                                 const double* Dr, const double* Tr, double* G);
/* multiply_*slice_matrix*! (stack.jl:319-367) on the context's conf; which: 0 left, 1 right,
 * 2 inv_right, 3 inv_left, 4 daggered_left; X is N x N x n_flavors x n_chains, in place. */
int32_t dqmc_op_multiply_slice_matrix(dqmc_ctx* ctx, int32_t which, int32_t slice, double* X);
/* wrap_greens!(mc, X, curr_slice, direction) (stack.jl:594-603) on host matrices. */
int32_t dqmc_op_wrap_greens(dqmc_ctx* ctx, int32_t curr_slice, int32_t direction, double* X);

/* ---- instrumentation ------------------------------------------------------------------------ */
/* the CUDA stream (cudaStream_t) every kernel and copy of this context is issued on, so that the
 * host can bracket calls with its own CUDA events. */
int32_t dqmc_get_stream(dqmc_ctx* ctx, void** cuda_stream);
/* per-launch CUDA-event timing by kernel category (the @bm TimerOutputs hooks of the reference,
 * src/helpers.jl:83-104).  enable != 0 resets the counters and starts recording. */
#define DQMC_PROF_GEMM 0     /* n x n x n batched DMMA GEMMs (vmul!, slice matrices, wraps)     */
#define DQMC_PROF_UDT 1      /* udt_AVX_pivot!                                                   */
#define DQMC_PROF_RDIVP 2    /* rdivp! (permute + panel GEMMs + diagonal-block solves)           */
#define DQMC_PROF_UPDATE 3   /* sweep_spatial                                                    */
#define DQMC_PROF_OTHER 4    /* copies, identity fills, propagation-error reduction             */
#define DQMC_PROF_NCAT 5
int32_t dqmc_profile(dqmc_ctx* ctx, int32_t enable);
int32_t dqmc_profile_report(dqmc_ctx* ctx, double* ms /* [DQMC_PROF_NCAT] */, int64_t* count /* [DQMC_PROF_NCAT] */);
/* number of CUDA kernels this context has launched so far. */
int64_t dqmc_kernel_launches(const dqmc_ctx* ctx);
/* largest n_sites the UDT kernel supports in this build. */
int32_t dqmc_max_sites(void);

#ifdef __cplusplus
}
#endif
#endif /* DQMC_B200_H */
