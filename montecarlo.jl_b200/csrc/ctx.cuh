// ctx.cuh -- the context object behind the C ABI and the host-side building blocks shared by
// capi.cu (equal-time stack, sweep) and ut.cu (unequal-time stack, CombinedGreensIterator).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "../../include/dqmc_b200.h"
#include "common.cuh"

using namespace dqmc;

struct dqmc_ut;     // ut.cu
struct dqmc_meas;   // measure.cu

struct dqmc_ctx {
    int N = 0, M = 0, nb = 1, kind = 0, B = 0, C = 0;
    std::vector<int> rfirst, rlast;
    double alpha = 0.0;
    bool ghq = false; double eta[4] = {0, 0, 0, 0}; GhqTables ghq_tab{};   // 4-state fields (kind >= 2)
    int check_sign = 1, check_prop = 1;
    unsigned long long seed = 0; long long chain_offset = 0; int device = 0; int kb = 0;
    int ld = 0; long long ms = 0; int nmat = 0; int ldv = 0;
    cudaStream_t st = nullptr;
    // device state
    double *eT2 = nullptr, *eT2i = nullptr, *eTh = nullptr, *eThi = nullptr;
    int8_t* conf = nullptr;
    double *u_stack = nullptr, *d_stack = nullptr, *t_stack = nullptr;
    double *greens = nullptr, *greens_temp = nullptr, *Ul = nullptr, *Ur = nullptr, *Tl = nullptr, *Tr = nullptr;
    double *tmp1 = nullptr, *tmp2 = nullptr, *curr_U = nullptr, *Dl = nullptr, *Dr = nullptr;
    double* Dgreens = nullptr;         // D of the last Green's function calculation: det(greens) = 1 / prod(Dgreens)
    bool fused_steps = false;          // n <= 64: runs of slice steps in one kernel (slicestep.cu)
    int update_version = 3;            // 3: update3.cu (n >= 96), 1: update.cu; dqmc_desc.update_variant forces one
    int8_t* conf_backup = nullptr;     // temp_conf of the global updates (fields.jl: temp_conf)
    double *Vwork = nullptr, *tau = nullptr, *udt_scratch = nullptr;
    int* pivot = nullptr; int* udt_iscratch = nullptr;
    int* accepted = nullptr;
    double *stats_neg = nullptr, *stats_prop = nullptr;
    double* obs = nullptr; long long obs_len = 0;
    double* d_uniforms = nullptr; unsigned char* d_forced = nullptr; double* d_probs = nullptr;
    unsigned char* d_dec = nullptr;
    double* h_stage = nullptr; size_t h_stage_bytes = 0;
    // stack state (stack.jl:50-52)
    int current_slice = 0, current_range = 1, direction = 1;
    long long sweep_index = 0;
    long long* d_sweep_index = nullptr; // device mirror of sweep_index (read by the update kernels; bumped on the stream)
    // one full local sweep captured as a CUDA graph ([0]: counter RNG, [1]: uniform table): the schedule of a sweep is
    // data independent, so replaying it removes ~2M x 6 host launches per sweep (the small lattices are launch bound)
    cudaGraphExec_t sweep_graph[2] = {nullptr, nullptr}; long long sweep_graph_launches[2] = {0, 0};
    bool graph_ok = true; long long eager_sweeps = 0;
    long long global_index = 0;        // running index of the global updates (site word of their counter-RNG uniform)
    long long launches = 0;            // kernels launched on behalf of this context
    void* nccl_comm = nullptr;         // communicator created by dqmc_comm_init (owned by the context)
    long long generation = 0;          // bumped whenever conf may have changed (mc.last_sweep's role for the ut stack)
    std::string err;
    std::vector<void*> allocs;
    // per-category CUDA-event profiler
    bool prof_on = false; int prof_depth = 0;
    std::vector<cudaEvent_t> prof_pool; size_t prof_used = 0;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> prof_spans;
    dqmc_meas* meas = nullptr;         // lattice tables + observable buffers (measure.cu)
    dqmc_ut* ut = nullptr;             // UnequalTimeStack + iterator state, allocated on first use (ut.cu)
};

inline cudaEvent_t prof_event(dqmc_ctx* c)
{
    if (c->prof_used == c->prof_pool.size()) {
        cudaEvent_t e; cudaEventCreate(&e); c->prof_pool.push_back(e);
    }
    return c->prof_pool[c->prof_used++];
}
// times the outermost category only (rdivp! contains GEMM launches of its own)
struct ProfScope {
    dqmc_ctx* c; cudaEvent_t e1 = nullptr; bool active;
    ProfScope(dqmc_ctx* ctx, int cat) : c(ctx), active(ctx->prof_on && ctx->prof_depth == 0)
    {
        ++c->prof_depth;
        if (active) {
            cudaEvent_t e0 = prof_event(c); e1 = prof_event(c);
            cudaEventRecord(e0, c->st);
            c->prof_spans.push_back({cat, {e0, e1}});
        }
    }
    ~ProfScope() { --c->prof_depth; if (active) cudaEventRecord(e1, c->st); }
};

#define FAIL(ctx, code, msg) do { (ctx)->err = (msg); return (code); } while (0)
#define CK(ctx, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__); return DQMC_ERR_CUDA; } } while (0)

template <class T> inline cudaError_t dalloc(dqmc_ctx* c, T** p, size_t count)
{
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, (count ? count : 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(q, 0, (count ? count : 1) * sizeof(T), c->st);
    c->allocs.push_back(q);
    *p = (T*)q;
    return e;
}

static inline double* slot_mat(dqmc_ctx* c, double* base, int slot) { return base + (long long)slot * c->nmat * c->ms; }
static inline double* slot_vec(dqmc_ctx* c, double* base, int slot) { return base + (long long)slot * c->nmat * c->N; }


#define CE(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return e__; } while (0)
#define ENTER(c) do { if (!(c)) return DQMC_ERR_INVALID; cudaSetDevice((c)->device); \
    dqmc::t_launch_counter = &(c)->launches; } while (0)
#define CHAINS_OK(c, c0, nc) ((c0) >= 0 && (nc) >= 0 && (c0) + (nc) <= (c)->B)

// ---- building blocks defined in capi.cu ----------------------------------------------------------
cudaError_t h2d_mats(dqmc_ctx* c, double* dst, const double* src, long long nmats);
cudaError_t d2h_mats(dqmc_ctx* c, double* dst, const double* src, long long nmats);
Scale field_scale(dqmc_ctx* c, int slice, double power);
Scale vec_scale(dqmc_ctx* c, const double* v, bool inverse = false);
GemmParams gemm_base(dqmc_ctx* c);
// dst = op(A) * op(B) with optional fused diagonal factors (dst must not alias A or B)
cudaError_t mm(dqmc_ctx* c, double* dst, const double* A, bool tA, bool sharedA, const double* Bm, bool tB,
               bool sharedB, Scale rs = no_scale(), Scale ks = no_scale(), Scale cs = no_scale(),
               const double* add_diag = nullptr, double alpha = 1.0, double beta = 0.0);
cudaError_t slice_left(dqmc_ctx* c, double* dst, const double* src, int slice);
cudaError_t slice_right(dqmc_ctx* c, double* dst, const double* src, int slice);
cudaError_t slice_inv_right(dqmc_ctx* c, double* dst, const double* src, int slice);
cudaError_t slice_inv_left(dqmc_ctx* c, double* dst, const double* src, int slice);
cudaError_t slice_daggered_left(dqmc_ctx* c, double* dst, const double* src, int slice);
cudaError_t slice_chain(dqmc_ctx* c, int op, const double* src, int first, int count, double* bufs[2], const double** out);
cudaError_t wrap_greens(dqmc_ctx* c, double* gf, double* tmp, int curr_slice, int direction);
cudaError_t udt(dqmc_ctx* c, const double* A, Scale colscale, double* U, double* D, double* T, bool apply_pivot);
cudaError_t rdivp(dqmc_ctx* c, double* A, const double* T, double* work);
cudaError_t copy_mats(dqmc_ctx* c, double* dst, const double* src);
cudaError_t copy_vecs(dqmc_ctx* c, double* dst, const double* src);
cudaError_t ident(dqmc_ctx* c, double* A);
cudaError_t ones(dqmc_ctx* c, double* v);
cudaError_t load_udt(dqmc_ctx* c, double* U, double* D, double* T, int slot);   // slot < 0 -> identity
cudaError_t build_chain_udt(dqmc_ctx* c, int slice, int safe_mult, bool dagger, double* U, double* D, double* T);
cudaError_t calculate_inv_greens_udt(dqmc_ctx* c, double* G);
cudaError_t calculate_greens(dqmc_ctx* c, double* G);
cudaError_t reverse_build(dqmc_ctx* c);
cudaError_t propagate(dqmc_ctx* c);
void ut_destroy(dqmc_ctx* c);     // ut.cu
void meas_destroy(dqmc_ctx* c);   // measure.cu
