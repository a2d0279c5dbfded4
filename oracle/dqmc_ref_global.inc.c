/*
 * oracle/dqmc_ref_global.inc.c -- CPU restatement of MonteCarlo.jl's global Metropolis update
 * (included at the end of dqmc_ref.c).  TEST INFRASTRUCTURE ONLY (see the header of dqmc_ref.c).
 *
 * Restates src/flavors/DQMC/updates/global_updates.jl: calculate_inv_greens_udt :25-52,
 * inv_det :70-137, propose_global_from_conf :147-179, accept_global! :181-198, global_update :203-219,
 * GlobalFlip's propose_conf! :243-248; energy_boson from fields.jl:395, 451.
 *
 * Parity status: PINNED by tests/test_oracle_global.py against the reference's own backbone check
 * (test/updates.jl:186-245: the global probability equals the product of the local probabilities of the
 * same flips and the final Green's functions agree) and against the brute-force determinant ratio.
 */

/* :25-52 for one block; leaves G = Ur Tr^-1 Dr^-1 Ul' Tl' unassembled, Dr is the result */
static void calculate_inv_greens_udt_block(int n, double *Ul, double *Dl, double *Tl, double *Ur, double *Dr,
                                           double *Tr, double *G, int64_t *pivot, double *temp)
{
    vmul_nt(n, G, Tl, Tr);
    vmul_mat_diag(n, Tr, G, Dr);
    vmul_diag_mat(n, G, Dl, Tr);
    ref_udt_pivot(n, Tr, Dr, G, pivot, temp, 0);
    vmul_nn(n, Tl, Ul, Tr);
    ref_rdivp(n, Ur, G, Ul, pivot);
    vmul_tn(n, Tr, Tl, Ur);
    for (int i = 0; i < n; ++i) Tr[IDX(i, i, n)] += Dr[i];
    ref_udt_pivot(n, Ul, Dr, Tr, pivot, temp, 0);
}

/* inv_det(mc, slice, field) :70-137 -> c->Dr */
static void inv_det(ref_chain *c, int slice, int safe_mult)
{
    const int n = c->N; const size_t nn = (size_t)n * n, tot = nn * c->nb, nv = (size_t)n * c->nb;
    set_identity(n, c->nb, c->curr_U);
    set_identity(n, c->nb, c->Ur); set_ones((int)nv, c->Dr); set_identity(n, c->nb, c->Tr);
    if (slice + 1 <= c->M) {
        for (int k = c->M; k >= slice + 1; --k) {
            multiply_daggered_slice_matrix_left(c, k, c->curr_U);
            if (k % safe_mult == 0) {
                for (int b = 0; b < c->nb; ++b) vmul_mat_diag(n, c->tmp1 + b * nn, c->curr_U + b * nn, c->Dr + b * n);
                udt_blocks(c, c->curr_U, c->Dr, c->tmp1, 1);
                memcpy(c->tmp2, c->Tr, sizeof(double) * tot);
                for (int b = 0; b < c->nb; ++b) vmul_nn(n, c->Tr + b * nn, c->tmp1 + b * nn, c->tmp2 + b * nn);
            }
        }
        for (int b = 0; b < c->nb; ++b) vmul_mat_diag(n, c->tmp1 + b * nn, c->curr_U + b * nn, c->Dr + b * n);
        udt_blocks(c, c->Ur, c->Dr, c->tmp1, 1);
        memcpy(c->tmp2, c->Tr, sizeof(double) * tot);
        for (int b = 0; b < c->nb; ++b) vmul_nn(n, c->Tr + b * nn, c->tmp1 + b * nn, c->tmp2 + b * nn);
    }
    set_identity(n, c->nb, c->curr_U);
    set_identity(n, c->nb, c->Ul); set_ones((int)nv, c->Dl); set_identity(n, c->nb, c->Tl);
    if (slice >= 1) {
        for (int k = 1; k <= slice; ++k) {
            multiply_slice_matrix_left(c, k, c->curr_U);
            if (k % safe_mult == 0) {
                for (int b = 0; b < c->nb; ++b) vmul_mat_diag(n, c->tmp1 + b * nn, c->curr_U + b * nn, c->Dl + b * n);
                udt_blocks(c, c->curr_U, c->Dl, c->tmp1, 1);
                memcpy(c->tmp2, c->Tl, sizeof(double) * tot);
                for (int b = 0; b < c->nb; ++b) vmul_nn(n, c->Tl + b * nn, c->tmp1 + b * nn, c->tmp2 + b * nn);
            }
        }
        for (int b = 0; b < c->nb; ++b) vmul_mat_diag(n, c->tmp1 + b * nn, c->curr_U + b * nn, c->Dl + b * n);
        udt_blocks(c, c->Ul, c->Dl, c->tmp1, 1);
        memcpy(c->tmp2, c->Tl, sizeof(double) * tot);
        for (int b = 0; b < c->nb; ++b) vmul_nn(n, c->Tl + b * nn, c->tmp1 + b * nn, c->tmp2 + b * nn);
    }
    for (int b = 0; b < c->nb; ++b)
        calculate_inv_greens_udt_block(n, c->Ul + b * nn, c->Dl + b * n, c->Tl + b * nn, c->Ur + b * nn,
                                       c->Dr + b * n, c->Tr + b * nn, c->greens_temp + b * nn, c->pivot + b * n,
                                       c->tempv + b * n);
}

/* fields.jl:395, 585 (density Hirsch / GHQ: alpha * sum(conf)) and :451, 530 (magnetic: 0) */
static double energy_boson(const ref_chain *c, const int8_t *conf)
{
    if (c->kind & 1) return 0.0;
    long s = 0;
    for (size_t i = 0; i < (size_t)c->N * c->M; ++i) s += conf[i];
    return c->alpha * (double)s;
}

/* global_update :203-219 with the proposal given as a full configuration (propose_conf! has run: the field
 * holds new_conf, temp_conf the old one).  Returns 1 if accepted; *p_out = |exp(-dE_boson) detratio|.
 * Must be called at (current_slice 1, direction +1), right after the Green's function was calculated
 * (mc.stack.Dl = 1 / D of that calculation, :151-154). */
int ref_global_update(ref_chain *c, const int8_t *new_conf, double uniform, int safe_mult, double *p_out)
{
    const size_t nv = (size_t)c->N * c->nb, nconf = (size_t)c->N * c->M;
    double *tempvf = dalloc(nv);
    int8_t *temp_conf = (int8_t *)malloc(nconf);
    memcpy(tempvf, c->Dl, sizeof(double) * nv);
    memcpy(temp_conf, c->conf, nconf);
    memcpy(c->conf, new_conf, nconf);
    inv_det(c, c->current_slice - 1, safe_mult);
    double detratio = 1.0;
    for (size_t i = 0; i < nv; ++i) detratio *= tempvf[i] * c->Dr[i];
    const double dE = energy_boson(c, c->conf) - energy_boson(c, temp_conf);
    if (c->nb == 1) detratio = detratio * detratio;
    const double p = fabs(exp(-dE) * detratio);
    if (p_out) *p_out = p;
    int acc = 0;
    if (p > 1.0 || uniform < p) {
        ref_reverse_build_stack(c);      /* accept_global! :181-198 */
        ref_propagate(c);
        acc = 1;
    } else {
        memcpy(c->conf, temp_conf, nconf);
    }
    free(tempvf); free(temp_conf);
    return acc;
}
