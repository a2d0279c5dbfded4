// measure.cu -- device-side Wick kernels: the equal-time and time-integrated observables are evaluated
// where the Green's functions live, so G never crosses PCIe between measurements.
//
// Reference (paths relative to /root/reference/src):
//   flavors/DQMC/measurements/generic.jl:287-310   apply!(::Greens)         (equal time)
//   flavors/DQMC/measurements/generic.jl:337-372   apply!(::TimeIntegral)   (sum over l with weights
//                                                   0.5 dtau at l = 0, M and dtau otherwise)
//   flavors/DQMC/measurements/generic.jl:434-461   apply!(temp, ::EachSitePairByDistance, ...) and
//                                     :578-583      finalize_temp! (divide by length(lattice))
//   .../constructors/charge_density.jl:62-110      full_cdc_kernel, summed over FlavorIterator(mc, 2)
//   .../constructors/spin_density.jl:66-218        full_sdc_{x,y,z}_kernel
//   .../constructors/occupation.jl:44-70           occupation
//   .../constructors/energy.jl:119-165, models/HubbardModel.jl:160-183   kinetic / interaction / total energy
//
// One CTA per (chain, basis pair): the N_bravais^2 site pairs are walked with the row index fastest
// (coalesced reads of Gl0[i, j]); the four pair observables share every load and are binned by
// direction into shared-memory accumulators.  HBM-bound and tiny next to the iterator's GEMMs.
#include "ctx.cuh"

cudaError_t ut_iter_begin(dqmc_ctx* c, int recalculate, int start, int stop, int safe_mult);
cudaError_t ut_iter_next(dqmc_ctx* c, int* l, const double** G0l, const double** Gl0, const double** Gll);

enum { OBS_OCC = 0, OBS_K, OBS_V, OBS_E, OBS_CDC, OBS_SDCX, OBS_SDCY, OBS_SDCZ, OBS_CDS, OBS_SDSX, OBS_SDSY,
       OBS_SDSZ, OBS_COUNT };

struct dqmc_meas {
    int nbr = 0, nbasis = 0;
    int* s2d = nullptr;            // [src + nbr * trg] -> direction, 0-based
    double* thop = nullptr;        // hopping matrix, N x N, leading dimension c->ld
    double U = 0.0;
    int off[OBS_COUNT + 1] = {0};
    double* res = nullptr;         // [chain][len]: values of the last measurement of every chain
    double* acc = nullptr;         // [count_equal_time, count_time_integral | sum[len] | sumsq[len]]
    double* h_res = nullptr;
    // log-binning (LogBinner of BinningAnalysis.jl, the accumulator behind every DQMCMeasurement,
    // measurements/generic.jl:62-65, 586-587): level 0 takes every value, level l + 1 the mean of two successive values
    // of level l.  Every chain is its own time series; the level statistics are summed over the chains.
    double* lb_pend = nullptr;     // [chain][level][len] value waiting for its partner (the compressors)
    double* lb_acc = nullptr;      // [counts: 2 x LB_LEVELS | sum[LB_LEVELS][len] | sumsq[LB_LEVELS][len]]
    long long pushes[2] = {0, 0};  // measurements committed so far: equal time, time integral
};
constexpr int LB_LEVELS = 20;      // 2^19 measurements per chain before the top level stops pairing

void meas_destroy(dqmc_ctx* c) { delete c->meas; c->meas = nullptr; }

// G matrices of one chain: block f at G + (chain * nb + f) * ms
__global__ void __launch_bounds__(256)
pair_obs_kernel(const double* __restrict__ G00, const double* __restrict__ G0l, const double* __restrict__ Gl0,
                const double* __restrict__ Gll, int l_is_zero, double weight, const int* __restrict__ s2d, int nbr,
                int nbasis, int N, int ld, long long ms, int nb, double* res, long long res_stride, int o_cd,
                int o_sx, int o_sy, int o_sz)
{
    extern __shared__ double acc[];                          // [4][nbr]
    const int chain = blockIdx.y;
    const int b1 = blockIdx.x % nbasis, b2 = blockIdx.x / nbasis;
    for (int e = threadIdx.x; e < 4 * nbr; e += blockDim.x) acc[e] = 0.0;
    __syncthreads();
    const long long base = (long long)chain * nb * ms;
    const double* g00[2] = {G00 + base, G00 + base + (nb - 1) * ms};
    const double* g0l[2] = {G0l + base, G0l + base + (nb - 1) * ms};
    const double* gl0[2] = {Gl0 + base, Gl0 + base + (nb - 1) * ms};
    const double* gll[2] = {Gll + base, Gll + base + (nb - 1) * ms};
    const int uc1 = b1 * nbr, uc2 = b2 * nbr;
    for (int e = threadIdx.x; e < nbr * nbr; e += blockDim.x) {
        const int src = e % nbr, trg = e / nbr;
        const int i = src + uc1, j = trg + uc2;
        const int dir = s2d[src + nbr * trg];
        const double id = (i == j && l_is_zero) ? 1.0 : 0.0;
        double cd, sx, sz;
        if (nb == 1) {
            // DiagonallyRepeatingMatrix: flv = 2
            const double a = 1.0 - gll[0][i + (long long)i * ld], b = 1.0 - g00[0][j + (long long)j * ld];
            const double x = (id - g0l[0][j + (long long)i * ld]) * gl0[0][i + (long long)j * ld];
            cd = 4.0 * a * b + 2.0 * x; sx = 2.0 * x; sz = 2.0 * x;
        } else {
            const double a1 = 1.0 - gll[0][i + (long long)i * ld], a2 = 1.0 - gll[1][i + (long long)i * ld];
            const double c1 = 1.0 - g00[0][j + (long long)j * ld], c2 = 1.0 - g00[1][j + (long long)j * ld];
            const double h1 = id - g0l[0][j + (long long)i * ld], h2 = id - g0l[1][j + (long long)i * ld];
            const double p1 = gl0[0][i + (long long)j * ld], p2 = gl0[1][i + (long long)j * ld];
            cd = (a1 + a2) * (c1 + c2) + h1 * p1 + h2 * p2;
            sx = h1 * p2 + h2 * p1;
            sz = (a1 - a2) * (c1 - c2) + h1 * p1 + h2 * p2;
        }
        atomicAdd(&acc[dir], cd);
        atomicAdd(&acc[nbr + dir], sx);
        atomicAdd(&acc[3 * nbr + dir], sz);
    }
    __syncthreads();
    const double f = weight / (double)N;                      // finalize_temp!: temp ./= length(lattice)
    double* r = res + (long long)chain * res_stride;
    const int slot = (b1 + nbasis * b2) * nbr;                // temp[dir, b1, b2], column-major
    for (int d = threadIdx.x; d < nbr; d += blockDim.x) {
        r[o_cd + slot + d] += f * acc[d];
        r[o_sx + slot + d] += f * acc[nbr + d];
        r[o_sy + slot + d] += f * acc[nbr + d];               // sdc_y == sdc_x for real block-diagonal G
        r[o_sz + slot + d] += f * acc[3 * nbr + d];
    }
}

// occupation, kinetic, interaction and total energy of one chain from the measured G
__global__ void __launch_bounds__(256)
scalar_obs_kernel(const double* __restrict__ G, const double* __restrict__ thop, double U, int N, int ld,
                  long long ms, int nb, double* res, long long res_stride, int o_occ, int o_k, int o_v, int o_e)
{
    const int chain = blockIdx.x;
    const double* g = G + (long long)chain * nb * ms;
    double* r = res + (long long)chain * res_stride;
    double kin = 0.0, inter = 0.0;
    for (int f = 0; f < nb; ++f)
        for (long long e = threadIdx.x; e < (long long)ld * N; e += blockDim.x) {
            const int i = (int)(e % ld), j = (int)(e / ld);
            if (i < N) kin += thop[e] * (((i == j) ? 1.0 : 0.0) - g[f * ms + e]);
        }
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double d1 = g[i + (long long)i * ld], d2 = g[(nb - 1) * ms + i + (long long)i * ld];
        inter -= (d1 - 0.5) * (d2 - 0.5);
        r[o_occ + i] = 1.0 - d1;
        if (nb == 2) r[o_occ + N + i] = 1.0 - d2;
    }
    __shared__ double red[2][32];
    for (int o = 16; o > 0; o >>= 1) { kin += __shfl_xor_sync(0xffffffffu, kin, o); inter += __shfl_xor_sync(0xffffffffu, inter, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = kin; red[1][threadIdx.x >> 5] = inter; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double k = 0.0, v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { k += red[0][w]; v += red[1][w]; }
        if (nb == 1) k *= 2.0;
        v *= U;
        r[o_k] = k; r[o_v] = v; r[o_e] = k + v;
    }
}

__global__ void meas_zero_kernel(double* res, long long res_stride, int o0, int o1, int n_chains)
{
    const int len = o1 - o0;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)len * n_chains;
         e += (long long)gridDim.x * blockDim.x)
        res[(e / len) * res_stride + o0 + (e % len)] = 0.0;
}

// push!(observable, temp) for every chain: {count, sum, sum of squares}
// nl = number of levels this push reaches = 1 + trailing ones of the number of earlier pushes (capped at LB_LEVELS)
__global__ void meas_commit_kernel(const double* res, long long res_stride, int o0, int o1, int n_chains,
                                   double* acc, int which_count, int len, double* lb_pend, double* lb_acc, int nl)
{
    for (int e = o0 + blockIdx.x * blockDim.x + threadIdx.x; e < o1; e += gridDim.x * blockDim.x) {
        double s = 0.0, s2 = 0.0;
        double ls[LB_LEVELS], ls2[LB_LEVELS];
        for (int l = 0; l < nl; ++l) ls[l] = ls2[l] = 0.0;
        for (int b = 0; b < n_chains; ++b) {
            const double x = res[(long long)b * res_stride + e];
            s += x; s2 += x * x;
            // LogBinner push!: accumulate at the level, then pair with the waiting value and move up
            double v = x;
            double* pend = lb_pend + ((long long)b * LB_LEVELS) * len + e;
            for (int l = 0; l < nl; ++l) {
                ls[l] += v; ls2[l] += v * v;
                if (l + 1 < nl) v = 0.5 * (pend[(long long)l * len] + v);
                else pend[(long long)l * len] = v;
            }
        }
        acc[2 + e] += s; acc[2 + len + e] += s2;
        for (int l = 0; l < nl; ++l) {
            lb_acc[2 * LB_LEVELS + (long long)l * len + e] += ls[l];
            lb_acc[2 * LB_LEVELS + ((long long)LB_LEVELS + l) * len + e] += ls2[l];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        acc[which_count] += (double)n_chains;
        for (int l = 0; l < nl; ++l) lb_acc[which_count * LB_LEVELS + l] += (double)n_chains;
    }
}

static cudaError_t launch_pair(dqmc_ctx* c, const double* G00, const double* G0l, const double* Gl0,
                               const double* Gll, int l_is_zero, double weight, int o_cd)
{
    dqmc_meas* m = c->meas;
    ProfScope ps(c, DQMC_PROF_OTHER);
    const int per = m->nbr * m->nbasis * m->nbasis;
    dim3 grid((unsigned)(m->nbasis * m->nbasis), (unsigned)c->B);
    pair_obs_kernel<<<grid, 256, (size_t)4 * m->nbr * sizeof(double), c->st>>>(
        G00, G0l, Gl0, Gll, l_is_zero, weight, m->s2d, m->nbr, m->nbasis, c->N, c->ld, c->ms, c->nb, m->res,
        m->off[OBS_COUNT], o_cd, o_cd + per, o_cd + 2 * per, o_cd + 3 * per);
    count_launch();
    return cudaGetLastError();
}
static cudaError_t launch_zero(dqmc_ctx* c, int o0, int o1)
{
    dqmc_meas* m = c->meas;
    meas_zero_kernel<<<148, 256, 0, c->st>>>(m->res, m->off[OBS_COUNT], o0, o1, c->B);
    count_launch();
    return cudaGetLastError();
}
static cudaError_t launch_commit(dqmc_ctx* c, int o0, int o1, int which_count)
{
    dqmc_meas* m = c->meas;
    const int blocks = std::min(148, (o1 - o0 + 255) / 256);
    int nl = 1;
    for (long long t = m->pushes[which_count]; (t & 1) && nl < LB_LEVELS; t >>= 1) ++nl;
    m->pushes[which_count] += 1;
    meas_commit_kernel<<<blocks, 256, 0, c->st>>>(m->res, m->off[OBS_COUNT], o0, o1, c->B, m->acc, which_count,
                                                  m->off[OBS_COUNT], m->lb_pend, m->lb_acc, nl);
    count_launch();
    return cudaGetLastError();
}

static cudaError_t measured_g00(dqmc_ctx* c)   // greens!(mc) -> greens_temp
{
    CE(mm(c, c->tmp2, c->greens, false, false, c->eTh, false, true));
    return mm(c, c->greens_temp, c->eThi, false, true, c->tmp2, false, false);
}

extern "C" {

int32_t dqmc_set_lattice(dqmc_ctx* c, int32_t n_bravais, int32_t n_basis, const int32_t* srctrg2dir,
                         const double* hopping_matrix, double U)
{
    ENTER(c);
    if (n_bravais < 1 || n_basis < 1 || n_bravais * n_basis != c->N || !srctrg2dir || !hopping_matrix)
        FAIL(c, DQMC_ERR_INVALID, "dqmc_set_lattice: need n_bravais * n_basis == n_sites and non-null tables");
    if ((size_t)4 * n_bravais * sizeof(double) > 200 * 1024) FAIL(c, DQMC_ERR_UNSUPPORTED, "dqmc_set_lattice: too many directions");
    std::vector<int> h((size_t)n_bravais * n_bravais);
    for (size_t i = 0; i < h.size(); ++i) {
        if (srctrg2dir[i] < 1 || srctrg2dir[i] > n_bravais) FAIL(c, DQMC_ERR_INVALID, "dqmc_set_lattice: direction out of range");
        h[i] = srctrg2dir[i] - 1;
    }
    if (c->meas) FAIL(c, DQMC_ERR_INVALID, "dqmc_set_lattice: lattice already set");
    dqmc_meas* m = new dqmc_meas();
    c->meas = m;
    m->nbr = n_bravais; m->nbasis = n_basis; m->U = U;
    const int per = n_bravais * n_basis * n_basis;
    int o = 0;
    for (int k = 0; k < OBS_COUNT; ++k) {
        m->off[k] = o;
        o += (k == OBS_OCC) ? c->nb * c->N : ((k <= OBS_E) ? 1 : per);
    }
    m->off[OBS_COUNT] = o;
    CK(c, dalloc(c, &m->s2d, h.size()));
    CK(c, dalloc(c, &m->thop, (size_t)c->ms));
    CK(c, dalloc(c, &m->res, (size_t)c->B * o));
    CK(c, dalloc(c, &m->acc, (size_t)2 + 2 * o));
    CK(c, dalloc(c, &m->lb_pend, (size_t)c->B * LB_LEVELS * o));
    CK(c, dalloc(c, &m->lb_acc, (size_t)2 * LB_LEVELS + (size_t)2 * LB_LEVELS * o));
    CK(c, cudaMemcpyAsync(m->s2d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, c->st));
    CK(c, h2d_mats(c, m->thop, hopping_matrix, 1));
    if ((size_t)4 * n_bravais * sizeof(double) > 48 * 1024)
        CK(c, cudaFuncSetAttribute(pair_obs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * n_bravais * (int)sizeof(double)));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

int32_t dqmc_measurement_layout(dqmc_ctx* c, int32_t* offsets)
{
    if (!c || !offsets) return DQMC_ERR_INVALID;
    if (!c->meas) FAIL(c, DQMC_ERR_INVALID, "dqmc_measurement_layout: call dqmc_set_lattice first");
    for (int k = 0; k <= OBS_COUNT; ++k) offsets[k] = c->meas->off[k];
    return DQMC_OK;
}

int32_t dqmc_measure_equal_time(dqmc_ctx* c)
{
    ENTER(c);
    dqmc_meas* m = c->meas;
    if (!m) FAIL(c, DQMC_ERR_INVALID, "dqmc_measure_equal_time: call dqmc_set_lattice first");
    if (c->current_slice != 1 || c->direction != 1)
        FAIL(c, DQMC_ERR_INVALID, "dqmc_measure_equal_time: measurements are taken at (slice 1, direction +1) (DQMC.jl:217)");
    CK(c, measured_g00(c));
    CK(c, launch_zero(c, m->off[OBS_CDC], m->off[OBS_CDS]));
    {
        ProfScope ps(c, DQMC_PROF_OTHER);
        scalar_obs_kernel<<<(unsigned)c->B, 256, 0, c->st>>>(c->greens_temp, m->thop, m->U, c->N, c->ld, c->ms, c->nb,
                                                              m->res, m->off[OBS_COUNT], m->off[OBS_OCC], m->off[OBS_K],
                                                              m->off[OBS_V], m->off[OBS_E]);
        count_launch();
        CK(c, cudaGetLastError());
    }
    CK(c, launch_pair(c, c->greens_temp, c->greens_temp, c->greens_temp, c->greens_temp, 1, 1.0, m->off[OBS_CDC]));
    CK(c, launch_commit(c, 0, m->off[OBS_CDS], 0));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

int32_t dqmc_measure_time_integral(dqmc_ctx* c, int32_t recalculate, int32_t safe_mult, double delta_tau)
{
    ENTER(c);
    dqmc_meas* m = c->meas;
    if (!m) FAIL(c, DQMC_ERR_INVALID, "dqmc_measure_time_integral: call dqmc_set_lattice first");
    if (c->current_slice != 1 || c->direction != 1)
        FAIL(c, DQMC_ERR_INVALID, "dqmc_measure_time_integral: measurements are taken at (slice 1, direction +1) (DQMC.jl:217)");
    if (recalculate < 1 || safe_mult < 1 || !(delta_tau > 0.0)) FAIL(c, DQMC_ERR_INVALID, "dqmc_measure_time_integral: bad arguments");
    CK(c, measured_g00(c));                                  // G00, constant over the iteration
    CK(c, launch_zero(c, m->off[OBS_CDS], m->off[OBS_COUNT]));
    CK(c, ut_iter_begin(c, recalculate, 0, c->M, safe_mult));
    for (;;) {
        int l = -1; const double *g0l = nullptr, *gl0 = nullptr, *gll = nullptr;
        CK(c, ut_iter_next(c, &l, &g0l, &gl0, &gll));
        if (l < 0) break;
        const double w = ((l == 0 || l == c->M) ? 0.5 : 1.0) * delta_tau;     // generic.jl:348
        CK(c, launch_pair(c, c->greens_temp, g0l, gl0, gll, l == 0, w, m->off[OBS_CDS]));
    }
    CK(c, launch_commit(c, m->off[OBS_CDS], m->off[OBS_COUNT], 1));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

int32_t dqmc_get_measurements(dqmc_ctx* c, int32_t chain0, int32_t nchains, double* out)
{
    ENTER(c);
    if (!c->meas) FAIL(c, DQMC_ERR_INVALID, "dqmc_get_measurements: call dqmc_set_lattice first");
    if (!out || !CHAINS_OK(c, chain0, nchains)) FAIL(c, DQMC_ERR_INVALID, "dqmc_get_measurements: bad arguments");
    const size_t len = (size_t)c->meas->off[OBS_COUNT];
    CK(c, cudaMemcpyAsync(out, c->meas->res + len * chain0, len * nchains * 8, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

int32_t dqmc_measurement_buffer(dqmc_ctx* c, void** device_ptr, int64_t* n_doubles)
{
    if (!c || !device_ptr || !n_doubles) return DQMC_ERR_INVALID;
    if (!c->meas) FAIL(c, DQMC_ERR_INVALID, "dqmc_measurement_buffer: call dqmc_set_lattice first");
    *device_ptr = c->meas->acc; *n_doubles = 2 + 2 * (int64_t)c->meas->off[OBS_COUNT];
    return DQMC_OK;
}

int32_t dqmc_binning_levels(void) { return LB_LEVELS; }

int32_t dqmc_measurement_binning_buffer(dqmc_ctx* c, void** device_ptr, int64_t* n_doubles)
{
    if (!c || !device_ptr || !n_doubles) return DQMC_ERR_INVALID;
    if (!c->meas) FAIL(c, DQMC_ERR_INVALID, "dqmc_measurement_binning_buffer: call dqmc_set_lattice first");
    *device_ptr = c->meas->lb_acc; *n_doubles = 2 * LB_LEVELS + 2 * (int64_t)LB_LEVELS * c->meas->off[OBS_COUNT];
    return DQMC_OK;
}

int32_t dqmc_get_measurement_binning(dqmc_ctx* c, double* counts, double* sum, double* sumsq)
{
    ENTER(c);
    if (!c->meas) FAIL(c, DQMC_ERR_INVALID, "dqmc_get_measurement_binning: call dqmc_set_lattice first");
    if (!counts || !sum || !sumsq) FAIL(c, DQMC_ERR_INVALID, "dqmc_get_measurement_binning: bad arguments");
    const size_t len = (size_t)c->meas->off[OBS_COUNT];
    CK(c, cudaMemcpyAsync(counts, c->meas->lb_acc, 2 * LB_LEVELS * 8, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaMemcpyAsync(sum, c->meas->lb_acc + 2 * LB_LEVELS, LB_LEVELS * len * 8, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaMemcpyAsync(sumsq, c->meas->lb_acc + 2 * LB_LEVELS + LB_LEVELS * len, LB_LEVELS * len * 8, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

int32_t dqmc_get_measurement_stats(dqmc_ctx* c, double* counts, double* sum, double* sumsq)
{
    ENTER(c);
    if (!c->meas) FAIL(c, DQMC_ERR_INVALID, "dqmc_get_measurement_stats: call dqmc_set_lattice first");
    if (!counts || !sum || !sumsq) FAIL(c, DQMC_ERR_INVALID, "dqmc_get_measurement_stats: bad arguments");
    const size_t len = (size_t)c->meas->off[OBS_COUNT];
    CK(c, cudaMemcpyAsync(counts, c->meas->acc, 16, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaMemcpyAsync(sum, c->meas->acc + 2, len * 8, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaMemcpyAsync(sumsq, c->meas->acc + 2 + len, len * 8, cudaMemcpyDeviceToHost, c->st));
    CK(c, cudaStreamSynchronize(c->st));
    return DQMC_OK;
}

}  // extern "C"
