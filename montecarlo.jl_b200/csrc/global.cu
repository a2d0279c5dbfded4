// global.cu -- global Metropolis updates on the device, batched over chains.
//
// Reference: src/flavors/DQMC/updates/global_updates.jl -- inv_det :70-137, propose_global_from_conf
// :147-179, accept_global! :181-198, global_update :203-219, GlobalFlip :237-248; energy_boson from
// fields.jl:395, 451.
//
// Every chain decides on its own proposal.  The weight ratio is det(G_old) / det(G_new) read off the
// diagonal factors of the two Green's function calculations (the reference's "ignore the phases"
// argument, :5-22): one batched chain product + partial Green's function calculation for the proposed
// configurations (inv_det), one tiny decision kernel, then -- like accept_global! -- a full
// reverse_build_stack + propagate.  Chains that rejected get their old configuration back before the
// rebuild, so their stack and G are simply recomputed.
#include "ctx.cuh"
#include "../../include/dqmc_rng.h"

// GlobalFlip's propose_conf! (:243-248)
__global__ void flip_conf_kernel(int8_t* conf, long long count)
{
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < count; e += (long long)gridDim.x * blockDim.x)
        conf[e] = (int8_t)(-conf[e]);
}

// one CTA per chain: p = |exp(-dE_boson) prod_i D_new[i] / D_old[i]| (squared for one flavor block),
// accept iff p > 1 || u < p (:203-219)
__global__ void __launch_bounds__(256)
global_decide_kernel(const double* __restrict__ Dold, const double* __restrict__ Dnew, const int8_t* __restrict__ conf_new,
                     const int8_t* __restrict__ conf_old, int N, int M, int nb, int kind, double alpha,
                     const double* __restrict__ uniforms, unsigned long long seed, long long chain0, long long sweep,
                     unsigned int global_index, int* accept, double* probs)
{
    const int chain = blockIdx.x;
    __shared__ long long red[8];
    long long s = 0;                                             // sum(conf_new) - sum(conf_old), exact
    const long long per = (long long)N * M;
    for (long long e = threadIdx.x; e < per; e += blockDim.x)
        s += (long long)conf_new[chain * per + e] - (long long)conf_old[chain * per + e];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long tot = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
        // same running product as the reference (:167-170): large and small factors interleave
        double detratio = 1.0;
        const long long v0 = (long long)chain * nb * N;
        for (int i = 0; i < nb * N; ++i) detratio *= Dnew[v0 + i] / Dold[v0 + i];
        if (nb == 1) detratio = detratio * detratio;
        const double dE = ((kind & 1) == 0) ? alpha * (double)tot : 0.0;   // energy_boson: density fields alpha sum(conf), magnetic 0
        const double p = fabs(exp(-dE) * detratio);
        const double u = uniforms ? uniforms[chain]
                                  : dqmc_uniform(seed, (uint64_t)(chain0 + chain), (uint64_t)sweep, (uint32_t)(2 * M), global_index);
        accept[chain] = (p > 1.0 || u < p) ? 1 : 0;
        if (probs) probs[chain] = p;
    }
}

__global__ void restore_conf_kernel(int8_t* conf, const int8_t* __restrict__ conf_old, const int* __restrict__ accept, long long per)
{
    const int chain = blockIdx.y;
    if (accept[chain]) return;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < per; e += (long long)gridDim.x * blockDim.x)
        conf[chain * per + e] = conf_old[chain * per + e];
}

extern "C" int32_t dqmc_global_update(dqmc_ctx* c, const int8_t* proposed, const double* uniforms, int32_t safe_mult,
                                      int64_t* accepted, double* probs)
{
    ENTER(c);
    if (safe_mult < 1) FAIL(c, DQMC_ERR_INVALID, "dqmc_global_update: safe_mult < 1");
    if (c->current_slice != 1 || c->direction != 1)
        FAIL(c, DQMC_ERR_INVALID, "dqmc_global_update: the stack must be at (slice 1, direction +1) (global_updates.jl:149-150)");
    const size_t per = (size_t)c->M * c->N, tot = per * c->B;
    if (!proposed && c->ghq)
        FAIL(c, DQMC_ERR_UNSUPPORTED, "dqmc_global_update: GlobalFlip (conf -> -conf) is not defined for the 4-state GHQ fields");
    if (proposed)
        for (size_t i = 0; i < tot; ++i)
            if (c->ghq ? (proposed[i] < 1 || proposed[i] > 4) : (proposed[i] != 1 && proposed[i] != -1))
                FAIL(c, DQMC_ERR_INVALID, "dqmc_global_update: conf values out of range for this field");
    if (!c->conf_backup) CK(c, dalloc(c, &c->conf_backup, tot));
    int* d_acc = nullptr; double* d_p = nullptr; double* d_u = nullptr;
    bool conf_replaced = false;
    // every exit path frees the temporaries; a failure after propose_conf! puts the old configuration back
    auto cleanup = [&](bool failed) {
        if (failed && conf_replaced) cudaMemcpyAsync(c->conf, c->conf_backup, tot, cudaMemcpyDeviceToDevice, c->st);
        if (d_acc) cudaFreeAsync(d_acc, c->st);
        if (d_p) cudaFreeAsync(d_p, c->st);
        if (d_u) cudaFreeAsync(d_u, c->st);
        if (failed) cudaStreamSynchronize(c->st);
    };
#define GCK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(true); \
    c->err = std::string(#call) + ": " + cudaGetErrorString(e__); return DQMC_ERR_CUDA; } } while (0)
    GCK(cudaMallocAsync((void**)&d_acc, (size_t)c->B * sizeof(int), c->st));
    GCK(cudaMallocAsync((void**)&d_p, (size_t)c->B * sizeof(double), c->st));
    if (uniforms) {
        GCK(cudaMallocAsync((void**)&d_u, (size_t)c->B * sizeof(double), c->st));
        GCK(cudaMemcpyAsync(d_u, uniforms, (size_t)c->B * 8, cudaMemcpyHostToDevice, c->st));
    }
    // propose_conf!: temp_conf <- conf, conf <- proposal
    GCK(cudaMemcpyAsync(c->conf_backup, c->conf, tot, cudaMemcpyDeviceToDevice, c->st));
    conf_replaced = true;
    if (proposed) GCK(cudaMemcpyAsync(c->conf, proposed, tot, cudaMemcpyHostToDevice, c->st));
    else {
        flip_conf_kernel<<<592, 256, 0, c->st>>>(c->conf, (long long)tot);
        ++c->launches;
        GCK(cudaGetLastError());
    }
    c->generation += 1;
    // inv_det(mc, current_slice - 1 = 0, field): (Ur Dr Tr)' = B_M ... B_1, Ul Dl Tl = identity
    GCK(build_chain_udt(c, 0, safe_mult, true, c->Ur, c->Dr, c->Tr));
    GCK(load_udt(c, c->Ul, c->Dl, c->Tl, -1));
    GCK(calculate_inv_greens_udt(c, c->greens_temp));
    {
        ProfScope ps(c, DQMC_PROF_OTHER);
        // the counter-RNG uniform of a global update is keyed by (chain, sweep index, step = 2M, site = running index of
        // the global update): consecutive global updates without a local sweep in between draw different uniforms
        global_decide_kernel<<<(unsigned)c->B, 256, 0, c->st>>>(c->Dgreens, c->Dr, c->conf, c->conf_backup, c->N, c->M, c->nb,
                                                                c->kind, c->alpha, d_u, c->seed, c->chain_offset,
                                                                c->sweep_index, (unsigned int)c->global_index, d_acc, d_p);
        ++c->launches;
        GCK(cudaGetLastError());
        dim3 grid(16, (unsigned)c->B);
        restore_conf_kernel<<<grid, 256, 0, c->st>>>(c->conf, c->conf_backup, d_acc, (long long)per);
        ++c->launches;
        GCK(cudaGetLastError());
    }
    c->global_index += 1;
    conf_replaced = false;                               // from here on conf is consistent (accepted or restored per chain)
    // accept_global!: full rebuild (rejected chains rebuild their old configuration)
    GCK(reverse_build(c));
    GCK(propagate(c));
    std::vector<int> h((size_t)c->B);
    GCK(cudaMemcpyAsync(h.data(), d_acc, (size_t)c->B * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    if (probs) GCK(cudaMemcpyAsync(probs, d_p, (size_t)c->B * 8, cudaMemcpyDeviceToHost, c->st));
    GCK(cudaStreamSynchronize(c->st));
    if (accepted) for (int b = 0; b < c->B; ++b) accepted[b] = h[b];
    cleanup(false);
#undef GCK
    return DQMC_OK;
}

extern "C" int32_t dqmc_set_global_update_index(dqmc_ctx* c, int64_t i)
{
    if (!c || i < 0) return DQMC_ERR_INVALID;
    c->global_index = i;
    return DQMC_OK;
}
