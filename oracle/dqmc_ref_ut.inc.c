/*
 * oracle/dqmc_ref_ut.inc.c -- CPU restatement of MonteCarlo.jl's unequal-time Green's function path
 * (included at the end of dqmc_ref.c; it uses that file's static helpers).
 *
 * TEST INFRASTRUCTURE ONLY (see the header of dqmc_ref.c).
 *
 * Restates, buffer for buffer (the reference reuses mc.stack's Ul..Tr, curr_U, tmp1, tmp2 as storage and
 * the aliasing matters):
 *   src/flavors/DQMC/unequal_time_stack.jl   UnequalTimeStack :1-118, build_stack :128-185,
 *       lazy_build_* :187-283, calculate_greens(mc, k, l) :322-335, _find_range_with_value :353-384,
 *       compute_inverse_udt_block! :400-457, compute_forward_udt_block! :472-494,
 *       compute_backward_udt_block! :509-533, calculate_greens_full1! :537-618, _full2! :621-697
 *   src/flavors/DQMC/measurements/greens_iterators.jl   CombinedGreensIterator :154-435
 *   src/flavors/DQMC/linalg/real.jl   rvadd! :117-121, vsub! :122-132, vmin!/vmax!/vmaxinv!/vinv! :137-178
 *   src/flavors/DQMC/measurements/generic.jl  apply!(::TimeIntegral) :337-368 and :430-461
 *   src/flavors/DQMC/measurements/constructors/charge_density.jl :62-110, spin_density.jl
 *
 * Parity status: PINNED against the reference's own tests for this path (tests/test_oracle_ut.py):
 * test/DQMC/unequal_time_stack.jl (stack equalities, G(k,k) == calculate_greens(k) < 1e-14,
 * G(t,0) == -G(t,M), iterator vs greens(k,l) < 2e-14 / 1e-10, high-precision G(37,14)) and the U = 0
 * analytic G(k,l) of test/ED/ED_tests.jl.
 */

typedef struct ref_ut {
    int forward_idx, backward_idx;          /* 1-based, like the reference                    */
    double *fu, *fd, *ft;                   /* forward  stack, C+1 slots  (:64-67)            */
    double *bu, *bd, *bt;                   /* backward stack, C+1 slots  (:70-73)            */
    int *inv_done; double *iu, *id, *it;    /* inverse  stack, C slots    (:76-79)            */
    double *greens, *tmp, *U, *D, *T;       /* :82-86                                         */
    int64_t last_update; int last_k, last_l;
    int it_recalc, it_start, it_stop, safe_mult;   /* CombinedGreensIterator spec             */
} ref_ut;

/* ---- per-block forwards (blockdiagonal.jl:198-249) ---- */
#define NB_LOOP(c) for (int b = 0; b < (c)->nb; ++b)
static void bd_nn(ref_chain *c, double *C_, const double *A, const double *B)
{ const size_t nn = (size_t)c->N * c->N; NB_LOOP(c) vmul_nn(c->N, C_ + b * nn, A + b * nn, B + b * nn); }
static void bd_nt(ref_chain *c, double *C_, const double *A, const double *B)
{ const size_t nn = (size_t)c->N * c->N; NB_LOOP(c) vmul_nt(c->N, C_ + b * nn, A + b * nn, B + b * nn); }
static void bd_tn(ref_chain *c, double *C_, const double *A, const double *B)
{ const size_t nn = (size_t)c->N * c->N; NB_LOOP(c) vmul_tn(c->N, C_ + b * nn, A + b * nn, B + b * nn); }
static void bd_md(ref_chain *c, double *C_, const double *A, const double *d)
{ const size_t nn = (size_t)c->N * c->N; NB_LOOP(c) vmul_mat_diag(c->N, C_ + b * nn, A + b * nn, d + b * c->N); }
static void bd_dm(ref_chain *c, double *C_, const double *d, const double *B)
{ const size_t nn = (size_t)c->N * c->N; NB_LOOP(c) vmul_diag_mat(c->N, C_ + b * nn, d + b * c->N, B + b * nn); }
static void bd_rdivp(ref_chain *c, double *A, const double *T, double *O)
{ const size_t nn = (size_t)c->N * c->N; NB_LOOP(c) ref_rdivp(c->N, A + b * nn, T + b * nn, O + b * nn, c->pivot + b * c->N); }
static void bd_copy(ref_chain *c, double *dst, const double *src)
{ memcpy(dst, src, sizeof(double) * (size_t)c->N * c->N * c->nb); }
static void bd_vcopy(ref_chain *c, double *dst, const double *src)
{ memcpy(dst, src, sizeof(double) * (size_t)c->N * c->nb); }
/* copyto!(A, Diagonal(d)) */
static void bd_from_diag(ref_chain *c, double *A, const double *d)
{
    const int n = c->N; const size_t nn = (size_t)n * n;
    memset(A, 0, sizeof(double) * nn * c->nb);
    NB_LOOP(c) for (int i = 0; i < n; ++i) A[b * nn + IDX(i, i, n)] = d[b * n + i];
}
/* real.jl:117-121 */
static void bd_rvadd(ref_chain *c, double *A, const double *B)
{ const size_t tot = (size_t)c->N * c->N * c->nb; for (size_t i = 0; i < tot; ++i) A[i] = A[i] + B[i]; }
/* real.jl:122-132 */
static void bd_vsub_I(ref_chain *c, double *O, const double *A)
{
    const int n = c->N; const size_t nn = (size_t)n * n;
    bd_copy(c, O, A);
    NB_LOOP(c) for (int i = 0; i < n; ++i) O[b * nn + IDX(i, i, n)] -= 1.0;
}
/* real.jl:137-178 */
static void vmin_(int len, double *v, const double *w) { for (int i = 0; i < len; ++i) v[i] = fmin(1.0, w[i]); }
static void vmaxinv_(int len, double *v, const double *w) { for (int i = 0; i < len; ++i) v[i] = 1.0 / fmax(1.0, w[i]); }
static void vinv_(int len, double *v) { for (int i = 0; i < len; ++i) v[i] = 1.0 / v[i]; }

/* greens.jl:114-125 with explicit buffers: temp = source * eThalf; target = eThalf^-1 * temp */
static void measured_greens_into(ref_chain *c, double *target, const double *source, double *temp)
{
    const int n = c->N; const size_t nn = (size_t)n * n;
    NB_LOOP(c) {
        vmul_nn(n, temp + b * nn, source + b * nn, c->eTh);
        vmul_nn(n, target + b * nn, c->eThi, temp + b * nn);
    }
}

/* unequal_time_stack.jl:60-118 */
static ref_ut *ut_get(ref_chain *c)
{
    if (c->ut) return (ref_ut *)c->ut;
    ref_ut *s = (ref_ut *)calloc(1, sizeof(ref_ut));
    const size_t nn = (size_t)c->N * c->N * c->nb, nv = (size_t)c->N * c->nb;
    const int E = c->C + 1;
    s->fu = dalloc(nn * E); s->fd = dalloc(nv * E); s->ft = dalloc(nn * E);
    s->bu = dalloc(nn * E); s->bd = dalloc(nv * E); s->bt = dalloc(nn * E);
    s->inv_done = (int *)calloc((size_t)c->C, sizeof(int));
    s->iu = dalloc(nn * c->C); s->id = dalloc(nv * c->C); s->it = dalloc(nn * c->C);
    s->greens = dalloc(nn); s->tmp = dalloc(nn); s->U = dalloc(nn); s->D = dalloc(nv); s->T = dalloc(nn);
    s->forward_idx = 1; s->backward_idx = E - 1;
    s->last_update = -1; s->last_k = s->last_l = -1;
    set_identity(c->N, c->nb, MAT(c, s->fu, 0)); set_ones((int)nv, VEC(c, s->fd, 0));
    set_identity(c->N, c->nb, MAT(c, s->ft, 0));
    set_identity(c->N, c->nb, MAT(c, s->bu, E - 1)); set_ones((int)nv, VEC(c, s->bd, E - 1));
    set_identity(c->N, c->nb, MAT(c, s->bt, E - 1));
    c->ut = s;
    return s;
}

static void ut_free(ref_chain *c)
{
    ref_ut *s = (ref_ut *)c->ut;
    if (!s) return;
    free(s->fu); free(s->fd); free(s->ft); free(s->bu); free(s->bd); free(s->bt);
    free(s->inv_done); free(s->iu); free(s->id); free(s->it);
    free(s->greens); free(s->tmp); free(s->U); free(s->D); free(s->T);
    free(s); c->ut = NULL;
}

/* one range of the forward / backward / inverse builds; idx is the 1-based range index */
static void ut_forward_step(ref_chain *c, ref_ut *s, int idx)        /* :132-143 */
{
    bd_copy(c, c->curr_U, MAT(c, s->fu, idx - 1));
    for (int sl = c->rfirst[idx - 1]; sl <= c->rlast[idx - 1]; ++sl) multiply_slice_matrix_left(c, sl, c->curr_U);
    bd_md(c, c->tmp1, c->curr_U, VEC(c, s->fd, idx - 1));
    udt_blocks(c, MAT(c, s->fu, idx), VEC(c, s->fd, idx), c->tmp1, 1);
    bd_nn(c, MAT(c, s->ft, idx), c->tmp1, MAT(c, s->ft, idx - 1));
}
static void ut_backward_step(ref_chain *c, ref_ut *s, int idx)       /* :148-159 */
{
    bd_copy(c, c->curr_U, MAT(c, s->bu, idx));
    for (int sl = c->rlast[idx - 1]; sl >= c->rfirst[idx - 1]; --sl) multiply_daggered_slice_matrix_left(c, sl, c->curr_U);
    bd_md(c, c->tmp1, c->curr_U, VEC(c, s->bd, idx));
    udt_blocks(c, MAT(c, s->bu, idx - 1), VEC(c, s->bd, idx - 1), c->tmp1, 1);
    bd_nn(c, MAT(c, s->bt, idx - 1), c->tmp1, MAT(c, s->bt, idx));
}
static void ut_inv_step(ref_chain *c, ref_ut *s, int idx)            /* :165-174 */
{
    double *t = MAT(c, s->it, idx - 1);
    set_identity(c->N, c->nb, t);
    for (int sl = c->rlast[idx - 1]; sl >= c->rfirst[idx - 1]; --sl) multiply_slice_matrix_inv_left(c, sl, t);
    udt_blocks(c, MAT(c, s->iu, idx - 1), VEC(c, s->id, idx - 1), t, 1);
}

/* build_stack(mc, ::UnequalTimeStack)  :128-185 */
void ref_ut_build_stack(ref_chain *c)
{
    ref_ut *s = ut_get(c);
    for (int idx = 1; idx <= c->C; ++idx) ut_forward_step(c, s, idx);
    for (int idx = c->C; idx >= 1; --idx) ut_backward_step(c, s, idx);
    for (int idx = 1; idx <= c->C; ++idx) { ut_inv_step(c, s, idx); s->inv_done[idx - 1] = 1; }
    s->forward_idx = c->C + 1; s->backward_idx = 1;
    s->last_update = c->sweep_index; s->last_k = s->last_l = -1;
}

static void ut_lazy_reset(ref_chain *c, ref_ut *s)                   /* :209-214 */
{
    if (s->last_update != c->sweep_index) {
        s->last_update = c->sweep_index;
        for (int i = 0; i < c->C; ++i) s->inv_done[i] = 0;
        s->forward_idx = 1; s->backward_idx = c->C + 1;
    }
}
static void ut_lazy_build_forward(ref_chain *c, ref_ut *s, int upto)     /* :207-238 */
{
    ut_lazy_reset(c, s);
    for (int idx = s->forward_idx; idx <= upto - 1; ++idx) ut_forward_step(c, s, idx);
    if (upto > s->forward_idx) s->forward_idx = upto;
}
static void ut_lazy_build_backward(ref_chain *c, ref_ut *s, int downto)  /* :240-270 */
{
    ut_lazy_reset(c, s);
    for (int idx = s->backward_idx - 1; idx >= downto; --idx) ut_backward_step(c, s, idx);
    if (downto < s->backward_idx) s->backward_idx = downto;
}
static void ut_lazy_build_inv(ref_chain *c, ref_ut *s, int from, int to) /* :272-300 */
{
    ut_lazy_reset(c, s);
    for (int idx = from; idx <= to; ++idx) {
        if (s->inv_done[idx - 1]) continue;
        s->inv_done[idx - 1] = 1;
        ut_inv_step(c, s, idx);
    }
}

/* :353-384: index of the range containing val; 0 below, C + 1 above */
static int find_range_with_value(const ref_chain *c, int val)
{
    if (val < 1) return 0;
    if (val > c->rlast[c->C - 1]) return c->C + 1;
    for (int i = 0; i < c->C; ++i) if (c->rfirst[i] <= val && val <= c->rlast[i]) return i + 1;
    return -1;
}
int ref_find_range_with_value(const ref_chain *c, int val) { return find_range_with_value(c, val); }

/* :400-457  U D T = B_{low+1}^-1 ... B_high^-1 into (s->U, s->D, s->T) */
static void compute_inverse_udt_block(ref_chain *c, ref_ut *s, int low, int high)
{
    const int n = c->N; const size_t nv = (size_t)n * c->nb;
    double *U = s->U, *D = s->D, *T = s->T, *tmp1 = c->tmp1, *tmp2 = c->tmp2;
    const int lower = find_range_with_value(c, low) + 1;
    const int upper = find_range_with_value(c, high + 1) - 1;
    ut_lazy_build_inv(c, s, lower, upper);
    set_identity(n, c->nb, U); set_ones((int)nv, D); set_identity(n, c->nb, T);
    for (int idx = lower; idx <= upper; ++idx) {
        bd_nn(c, tmp1, T, MAT(c, s->iu, idx - 1));
        bd_dm(c, tmp2, D, tmp1);
        bd_md(c, tmp1, tmp2, VEC(c, s->id, idx - 1));
        udt_blocks(c, tmp2, D, tmp1, 1);
        bd_nn(c, T, tmp1, MAT(c, s->it, idx - 1));
        bd_nn(c, tmp1, U, tmp2);
        bd_copy(c, U, tmp1);
    }
    const int lower_slice = (lower <= c->C) ? c->rfirst[lower - 1] : c->rlast[c->C - 1] + 1;
    const int upper_slice = (upper > 0) ? c->rlast[upper - 1] : 0;
    const int top = (lower_slice - 1 < high) ? lower_slice - 1 : high;
    for (int sl = top; sl >= low + 1; --sl) multiply_slice_matrix_inv_left(c, sl, U);
    if (top >= low + 1) {
        bd_md(c, tmp1, U, D);
        udt_blocks(c, U, D, tmp1, 1);
        bd_nn(c, tmp2, tmp1, T);
        bd_copy(c, T, tmp2);
    }
    const int from = (upper_slice + 1 > top + 1) ? upper_slice + 1 : top + 1;
    for (int sl = from; sl <= high; ++sl) multiply_slice_matrix_inv_right(c, sl, T);
}

/* :472-494  Ul Dl Tl = B_slice ... B_1 */
static void compute_forward_udt_block(ref_chain *c, ref_ut *s, int slice)
{
    int idx = find_range_with_value(c, slice) - 1;
    if (idx < 0) idx = 0;
    ut_lazy_build_forward(c, s, idx + 1);
    bd_copy(c, c->Tl, MAT(c, s->fu, idx));
    const int target = (idx > 0) ? c->rlast[idx - 1] + 1 : 1;
    for (int l = target; l <= slice; ++l) multiply_slice_matrix_left(c, l, c->Tl);
    bd_md(c, c->tmp1, c->Tl, VEC(c, s->fd, idx));
    udt_blocks(c, c->Ul, c->Dl, c->tmp1, 1);
    bd_nn(c, c->Tl, c->tmp1, MAT(c, s->ft, idx));
}

/* :509-533  (Ur Dr Tr)' = B_M ... B_{slice+1} */
static void compute_backward_udt_block(ref_chain *c, ref_ut *s, int slice)
{
    const int idx = find_range_with_value(c, slice) + 1;
    ut_lazy_build_backward(c, s, idx);
    bd_copy(c, c->Ur, MAT(c, s->bu, idx - 1));
    const int target = (idx <= c->C) ? c->rfirst[idx - 1] - 1 : c->rlast[c->C - 1];
    for (int l = target; l >= slice + 1; --l) multiply_daggered_slice_matrix_left(c, l, c->Ur);
    bd_md(c, c->tmp1, c->Ur, VEC(c, s->bd, idx - 1));
    udt_blocks(c, c->Ur, c->Dr, c->tmp1, 1);
    bd_nn(c, c->Tr, c->tmp1, MAT(c, s->bt, idx - 1));
}

/* :537-618, slice1 >= slice2 */
static void calculate_greens_full1(ref_chain *c, ref_ut *s, int slice1, int slice2)
{
    const int nv = c->N * c->nb;
    compute_inverse_udt_block(c, s, slice2, slice1);
    compute_forward_udt_block(c, s, slice2);
    compute_backward_udt_block(c, s, slice1);
    /* B1 */
    bd_nt(c, s->greens, c->Tl, c->Tr);
    bd_md(c, c->tmp1, s->greens, c->Dr);
    bd_dm(c, s->greens, c->Dl, c->tmp1);
    udt_blocks(c, c->Tr, c->Dr, s->greens, 0);
    /* B2 */
    bd_nn(c, c->Tl, c->Ul, c->Tr);
    bd_rdivp(c, c->Ur, s->greens, c->Ul);
    /* B3 */
    bd_tn(c, c->Tr, s->U, c->Tl);
    vmaxinv_(nv, c->Dl, s->D);
    bd_dm(c, c->tmp1, c->Dl, c->Tr);
    vmin_(nv, c->Dl, c->Dr);
    bd_md(c, c->Tr, c->tmp1, c->Dl);
    /* B4 */
    bd_nn(c, c->Tl, s->T, c->Ur);
    vmin_(nv, c->Dl, s->D);
    bd_dm(c, c->tmp1, c->Dl, c->Tl);
    vmaxinv_(nv, c->Dl, c->Dr);
    bd_md(c, c->Tl, c->tmp1, c->Dl);
    /* sum, UDT */
    bd_rvadd(c, c->Tl, c->Tr);
    udt_blocks(c, c->Tr, c->Dl, c->Tl, 0);
    /* B5 */
    vmaxinv_(nv, c->Dr, c->Dr);
    bd_from_diag(c, c->Ul, c->Dr);
    bd_rdivp(c, c->Ul, c->Tl, c->tmp1);
    vinv_(nv, c->Dl);
    bd_md(c, c->tmp1, c->Ul, c->Dl);
    bd_nt(c, c->Ul, c->tmp1, c->Tr);
    vmaxinv_(nv, c->Dl, s->D);
    bd_md(c, s->greens, c->Ul, c->Dl);
    /* B6 */
    bd_nt(c, c->Tr, s->greens, s->U);
    bd_nn(c, s->greens, c->Ur, c->Tr);
}

/* :621-697, slice1 <= slice2 */
static void calculate_greens_full2(ref_chain *c, ref_ut *s, int slice1, int slice2)
{
    const int nv = c->N * c->nb;
    const size_t tot = (size_t)c->N * c->N * c->nb;
    compute_inverse_udt_block(c, s, slice1, slice2);
    compute_forward_udt_block(c, s, slice1);
    compute_backward_udt_block(c, s, slice2);
    /* B1 */
    bd_nt(c, s->greens, c->Tl, c->Tr);
    bd_dm(c, c->tmp1, c->Dl, s->greens);
    bd_md(c, s->greens, c->tmp1, c->Dr);
    udt_blocks(c, c->Tr, c->Dr, s->greens, 0);
    /* B2 */
    bd_nn(c, c->Tl, c->Ul, c->Tr);
    bd_tn(c, c->Ul, s->U, c->Tl);
    vmaxinv_(nv, c->Dl, s->D);
    bd_dm(c, s->U, c->Dl, c->Ul);
    vmin_(nv, c->Dl, c->Dr);
    bd_md(c, c->Ul, s->U, c->Dl);
    /* B3 */
    bd_nn(c, s->U, s->T, c->Ur);
    bd_rdivp(c, s->U, s->greens, c->Ur);
    vmin_(nv, c->Dl, s->D);
    bd_dm(c, c->Ur, c->Dl, s->U);
    vmaxinv_(nv, c->Dl, c->Dr);
    bd_md(c, c->Tr, c->Ur, c->Dl);
    /* sum, udt */
    bd_rvadd(c, c->Tr, c->Ul);
    udt_blocks(c, c->Ul, c->Dl, c->Tr, 0);
    /* B4 */
    vmin_(nv, c->Dr, c->Dr);
    bd_from_diag(c, s->U, c->Dr);
    bd_rdivp(c, s->U, c->Tr, c->Ur);
    vinv_(nv, c->Dl);
    bd_md(c, c->Ur, s->U, c->Dl);
    bd_nt(c, s->U, c->Ur, c->Ul);
    vmin_(nv, s->D, s->D);
    bd_md(c, c->Ur, s->U, s->D);
    /* B6 */
    bd_nn(c, c->Tr, c->Ur, s->T);
    bd_nn(c, s->greens, c->Tl, c->Tr);
    for (size_t i = 0; i < tot; ++i) s->greens[i] *= -1.0;
}

/* calculate_greens(mc, slice1, slice2) :322-335 -> effective G(k, l) (no eThalf transform) */
void ref_ut_calculate_greens(ref_chain *c, int slice1, int slice2, double *out)
{
    ref_ut *s = ut_get(c);
    if (s->last_k != slice1 || s->last_l != slice2 || s->last_update != c->sweep_index) {
        s->last_k = slice1; s->last_l = slice2;
        if (slice1 >= slice2) calculate_greens_full1(c, s, slice1, slice2);
        else calculate_greens_full2(c, s, slice1, slice2);
    }
    if (out) bd_copy(c, out, s->greens);
}

/* greens(mc, k, l) :302-320 = _greens!(calculate_greens(mc, k, l)) -> measured G(k, l) in greens_temp */
void ref_ut_greens(ref_chain *c, int slice1, int slice2, double *out)
{
    ref_ut *s = ut_get(c);
    ref_ut_calculate_greens(c, slice1, slice2, NULL);
    measured_greens_into(c, c->greens_temp, s->greens, c->tmp1);
    bd_copy(c, out, c->greens_temp);
}

/* which: 0/1/2 forward u/d/t, 3/4/5 backward u/d/t, 6/7/8 inverse u/d/t; slot 0-based */
void ref_ut_get_array(ref_chain *c, int which, int slot, double *out)
{
    ref_ut *s = ut_get(c);
    const size_t tot = (size_t)c->N * c->N * c->nb, nv = (size_t)c->N * c->nb;
    double *bases[9] = {s->fu, s->fd, s->ft, s->bu, s->bd, s->bt, s->iu, s->id, s->it};
    if (which % 3 == 1) memcpy(out, VEC(c, bases[which], slot), sizeof(double) * nv);
    else memcpy(out, MAT(c, bases[which], slot), sizeof(double) * tot);
}
void ref_ut_lazy_build(ref_chain *c, int forward_upto, int backward_downto)
{
    ref_ut *s = ut_get(c);
    if (forward_upto > 0) ut_lazy_build_forward(c, s, forward_upto);
    if (backward_downto > 0) ut_lazy_build_backward(c, s, backward_downto);
}

/* ---- CombinedGreensIterator (greens_iterators.jl:198-435) ------------------------------------ */
/* outputs of one iteration live in (G0l, Gl0, Gll) = (stack.tmp2, stack.tmp1, uts.greens) */
static void cgi_emit(ref_chain *c, ref_ut *s, double *G0l, double *Gl0, double *Gll)
{
    if (G0l) bd_copy(c, G0l, c->tmp2);
    if (Gl0) bd_copy(c, Gl0, c->tmp1);
    if (Gll) bd_copy(c, Gll, s->greens);
}

static void cgi_recalculate(ref_chain *c, ref_ut *s, int l, int first)
{
    /* :253-272 (first iteration with start > 1) and :312-331 (recalculation) differ only in the
     * temporaries handed to _greens!; the values are the same */
    calculate_greens_full1(c, s, l, 0);
    bd_copy(c, c->curr_U, s->greens);
    calculate_greens_full2(c, s, 0, l);
    bd_copy(c, s->tmp, s->greens);
    calculate_greens_full1(c, s, l, l);
    bd_copy(c, s->T, s->greens);
    if (first) {
        measured_greens_into(c, s->greens, s->T, c->tmp1);       /* Gll */
        measured_greens_into(c, c->tmp1, c->curr_U, c->tmp2);    /* Gl0 */
        bd_copy(c, c->Tl, c->curr_U);
        measured_greens_into(c, c->tmp2, s->tmp, c->curr_U);     /* G0l */
        bd_copy(c, c->Tr, s->tmp);
    } else {
        measured_greens_into(c, s->greens, s->T, c->tmp2);       /* Gll */
        bd_copy(c, c->Tl, c->curr_U);
        measured_greens_into(c, c->tmp1, c->curr_U, c->tmp2);    /* Gl0 */
        bd_copy(c, c->Tr, s->tmp);
        measured_greens_into(c, c->tmp2, s->tmp, c->curr_U);     /* G0l */
    }
    udt_blocks(c, s->U, s->D, s->T, 1);
    udt_blocks(c, c->Ul, c->Dl, c->Tl, 1);
    udt_blocks(c, c->Ur, c->Dr, c->Tr, 1);
}

int ref_cgi_next(ref_chain *c, int l, double *G0l, double *Gl0, double *Gll);

/* iterate(it) :198-293.  Returns the next state l, or -1 when the iteration is over. */
int ref_cgi_first(ref_chain *c, int recalculate, int start, int stop, int safe_mult,
                  double *G0l, double *Gl0, double *Gll)
{
    ref_ut *s = ut_get(c);
    s->it_recalc = recalculate; s->it_start = start; s->it_stop = stop; s->safe_mult = safe_mult;
    ref_ut_build_stack(c);
    s->last_k = s->last_l = -1;
    if (start == 0 || start == 1) {
        if (c->current_slice == 1) bd_copy(c, c->Tl, c->greens);
        else { calculate_greens_full1(c, s, 0, 0); bd_copy(c, c->Tl, s->greens); }
        bd_copy(c, c->tmp1, c->Tl);
        bd_vsub_I(c, c->Tr, c->Tl);
        udt_blocks(c, c->Ul, c->Dl, c->Tl, 1);
        bd_copy(c, s->U, c->Ul); bd_vcopy(c, s->D, c->Dl); bd_copy(c, s->T, c->Tl);
        udt_blocks(c, c->Ur, c->Dr, c->Tr, 1);
        if (start == 0) {
            measured_greens_into(c, s->greens, c->tmp1, c->tmp2);
            bd_copy(c, c->tmp1, s->greens);
            bd_copy(c, c->tmp2, s->greens);
            cgi_emit(c, s, G0l, Gl0, Gll);
            return 1;
        }
        return ref_cgi_next(c, 1, G0l, Gl0, Gll);
    }
    cgi_recalculate(c, s, start, 1);
    cgi_emit(c, s, G0l, Gl0, Gll);
    return start + 1;
}

/* iterate(it, l) :295-435 */
int ref_cgi_next(ref_chain *c, int l, double *G0l, double *Gl0, double *Gll)
{
    ref_ut *s = ut_get(c);
    s->last_k = s->last_l = -1;
    const int shift = (s->it_start != 1) ? s->it_start : 0;
    if (l > s->it_stop) return -1;
    if ((l - shift) % s->it_recalc == 0) {
        cgi_recalculate(c, s, l, 0);
    } else if (((l - shift) % s->it_recalc) % s->safe_mult == 0) {
        /* stabilisation :343-390 */
        multiply_slice_matrix_left(c, l, c->Ul);
        multiply_slice_matrix_inv_right(c, l, c->Tr);
        multiply_slice_matrix_left(c, l, s->U);
        multiply_slice_matrix_inv_right(c, l, s->T);
        /* Gl0 */
        bd_md(c, c->tmp1, c->Ul, c->Dl);
        bd_nn(c, c->tmp2, c->tmp1, c->Tl);
        udt_blocks(c, c->Ul, c->Dl, c->tmp1, 1);
        bd_nn(c, c->curr_U, c->tmp1, c->Tl);
        bd_copy(c, c->Tl, c->curr_U);
        measured_greens_into(c, c->tmp1, c->tmp2, c->curr_U);
        /* G0l */
        bd_dm(c, c->curr_U, c->Dr, c->Tr);
        bd_nn(c, s->greens, c->Ur, c->curr_U);
        bd_copy(c, c->Tr, c->curr_U);
        udt_blocks(c, c->tmp2, c->Dr, c->Tr, 1);
        bd_nn(c, c->curr_U, c->Ur, c->tmp2);
        bd_copy(c, c->Ur, c->curr_U);
        measured_greens_into(c, c->tmp2, s->greens, c->curr_U);
        /* Gll */
        bd_md(c, s->tmp, s->U, s->D);
        bd_nn(c, s->greens, s->tmp, s->T);
        udt_blocks(c, c->curr_U, s->D, s->tmp, 1);
        bd_nn(c, s->U, s->tmp, s->T);
        bd_dm(c, s->T, s->D, s->U);
        udt_blocks(c, s->tmp, s->D, s->T, 1);
        bd_nn(c, s->U, c->curr_U, s->tmp);
        measured_greens_into(c, s->greens, s->greens, c->curr_U);
    } else {
        /* quick advance :392-420 */
        multiply_slice_matrix_left(c, l, c->Ul);
        multiply_slice_matrix_inv_right(c, l, c->Tr);
        multiply_slice_matrix_left(c, l, s->U);
        multiply_slice_matrix_inv_right(c, l, s->T);
        bd_md(c, c->curr_U, c->Ul, c->Dl);
        bd_nn(c, c->tmp2, c->curr_U, c->Tl);
        measured_greens_into(c, c->tmp1, c->tmp2, c->curr_U);
        bd_md(c, c->curr_U, c->Ur, c->Dr);
        bd_nn(c, s->greens, c->curr_U, c->Tr);
        measured_greens_into(c, c->tmp2, s->greens, c->curr_U);
        bd_md(c, c->curr_U, s->U, s->D);
        bd_nn(c, s->greens, c->curr_U, s->T);
        measured_greens_into(c, s->greens, s->greens, c->curr_U);
    }
    cgi_emit(c, s, G0l, Gl0, Gll);
    return l + 1;
}
