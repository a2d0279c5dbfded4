/*
 * oracle/dqmc_ref.c -- CPU restatement of MonteCarlo.jl's DQMC sweep hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA library
 * (montecarlo.jl_b200/csrc) and the CPU baseline of bench.py.  Nothing in the
 * product path may include, link or call it.
 *
 * Parity status: PINNED against the reference's own known-answer tests
 * (tests/test_oracle_*.py: U=0 analytic G, independent LAPACK-QR Green's
 * function, forward/reverse stack equivalence, slice-matrix products, UDT
 * identities, rank-1 update formula, local-vs-global determinant ratios,
 * lattice bond golden lists, chunk invariants).  The reference itself (Julia)
 * cannot run in this image, so no outputs of the real package are used.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose loop structure it restates.  All matrices are column-major double.
 * A "block" is one flavor block of the BlockDiagonal Green's matrix
 * (src/flavors/DQMC/linalg/blockdiagonal.jl:22-45): nb = 1 for
 * DensityHirschField, nb = 2 for MagneticHirschField; every linalg routine
 * is applied per block (blockdiagonal.jl:198-249,286-324,352-370).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "../include/dqmc_rng.h"

#define IDX(i, j, n) ((size_t)(i) + (size_t)(j) * (size_t)(n))

/* ------------------------------------------------------------------------ */
/* dense products: src/flavors/DQMC/linalg/real.jl:7-15, 72-102              */
/* ------------------------------------------------------------------------ */

/* The reference's products are LoopVectorization @turbo loops (real.jl:7-15, 72-102): the macro register-tiles the
 * two output indices, keeps the C tile in vector registers across the whole k loop and reassociates reductions; it
 * does no cache blocking or packing.  The three kernels below restate exactly that (and nothing more: no blocking
 * over k, no packing), so the CPU baseline is neither flattered nor penalised relative to the package. */
typedef double v8d __attribute__((vector_size(64), aligned(8), may_alias));
static inline v8d ld8(const double *p) { return *(const v8d *)p; }
static inline void st8(double *p, v8d v) { *(v8d *)p = v; }

/* C(i,j) = sum_k A[i + k n] * B[k bk + j bj]: A is contiguous along the output row index, so rows are the vector
 * dimension.  Tile = 16 rows x 4 columns. */
static void gemm_rows_contiguous(int n, double *restrict C, const double *restrict A, const double *restrict B,
                                 size_t bk, size_t bj)
{
    const size_t ld = (size_t)n;
    for (int j = 0; j < n; j += 4) {
        const int jw = (n - j < 4) ? n - j : 4;
        int i = 0;
        if (jw == 4) {
            for (; i + 16 <= n; i += 16) {
                v8d c00 = {0}, c01 = {0}, c10 = {0}, c11 = {0}, c20 = {0}, c21 = {0}, c30 = {0}, c31 = {0};
                const double *b0 = B + (size_t)j * bj, *b1 = b0 + bj, *b2 = b1 + bj, *b3 = b2 + bj;
                for (int k = 0; k < n; ++k) {
                    const v8d a0 = ld8(A + i + k * ld), a1 = ld8(A + i + 8 + k * ld);
                    const double x0 = b0[k * bk], x1 = b1[k * bk], x2 = b2[k * bk], x3 = b3[k * bk];
                    c00 += a0 * x0; c01 += a1 * x0; c10 += a0 * x1; c11 += a1 * x1;
                    c20 += a0 * x2; c21 += a1 * x2; c30 += a0 * x3; c31 += a1 * x3;
                }
                double *c = C + i + (size_t)j * ld;
                st8(c, c00); st8(c + 8, c01); st8(c + ld, c10); st8(c + ld + 8, c11);
                st8(c + 2 * ld, c20); st8(c + 2 * ld + 8, c21); st8(c + 3 * ld, c30); st8(c + 3 * ld + 8, c31);
            }
        }
        for (; i < n; ++i)
            for (int jj = 0; jj < jw; ++jj) {
                double sacc = 0.0;
                const double *bb = B + (size_t)(j + jj) * bj;
                for (int k = 0; k < n; ++k) sacc += A[i + k * ld] * bb[k * bk];
                C[i + (size_t)(j + jj) * ld] = sacc;
            }
    }
}

/* C = A * B   (real.jl:7-15) */
static void vmul_nn(int n, double *restrict C, const double *restrict A, const double *restrict B)
{ gemm_rows_contiguous(n, C, A, B, 1, (size_t)n); }

/* C = A * B'   (real.jl:72-81) */
static void vmul_nt(int n, double *restrict C, const double *restrict A, const double *restrict B)
{ gemm_rows_contiguous(n, C, A, B, (size_t)n, 1); }

/* C = A' * B   (real.jl:82-91): both operands are contiguous along k, so k is the vector dimension: a 4 x 4 tile of
 * vector partial sums, reduced at the end. */
static void vmul_tn(int n, double *restrict C, const double *restrict A, const double *restrict B)
{
    const size_t ld = (size_t)n;
    const int n4 = n & ~3, k8 = n & ~7;
    for (int j = 0; j < n4; j += 4)
        for (int i = 0; i < n4; i += 4) {
            v8d acc[4][4];
            for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) acc[a][b] = (v8d){0};
            const double *a0 = A + i * ld, *b0 = B + j * ld;
            for (int k = 0; k < k8; k += 8) {
                const v8d x0 = ld8(a0 + k), x1 = ld8(a0 + ld + k), x2 = ld8(a0 + 2 * ld + k), x3 = ld8(a0 + 3 * ld + k);
                for (int b = 0; b < 4; ++b) {
                    const v8d y = ld8(b0 + b * ld + k);
                    acc[0][b] += x0 * y; acc[1][b] += x1 * y; acc[2][b] += x2 * y; acc[3][b] += x3 * y;
                }
            }
            for (int a = 0; a < 4; ++a)
                for (int b = 0; b < 4; ++b) {
                    double sacc = 0.0;
                    for (int l = 0; l < 8; ++l) sacc += acc[a][b][l];
                    for (int k = k8; k < n; ++k) sacc += a0[a * ld + k] * b0[b * ld + k];
                    C[(i + a) + (size_t)(j + b) * ld] = sacc;
                }
        }
    for (int j = 0; j < n; ++j)
        for (int i = (j < n4 ? n4 : 0); i < n; ++i) {
            double sacc = 0.0;
            for (int k = 0; k < n; ++k) sacc += A[k + i * ld] * B[k + j * ld];
            C[i + j * ld] = sacc;
        }
}

/* C = A * Diagonal(d)  (real.jl:16-20) */
static void vmul_mat_diag(int n, double *C, const double *A, const double *d)
{
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) C[IDX(i, j, n)] = A[IDX(i, j, n)] * d[j];
}

/* C = Diagonal(d) * B  (real.jl:27-31) */
static void vmul_diag_mat(int n, double *C, const double *d, const double *B)
{
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) C[IDX(i, j, n)] = d[i] * B[IDX(i, j, n)];
}

/* ------------------------------------------------------------------------ */
/* UDT: src/flavors/DQMC/linalg/UDT.jl:157-192 (reflector!, indmaxcolumn),   */
/*      :53-70 (reflectorApply!), :216-334 (udt_AVX_pivot!, _apply_pivot!)   */
/* ------------------------------------------------------------------------ */

/* UDT.jl:175-192.  First index wins ties (strict >). 0-based j. */
static int indmaxcolumn(int n, const double *A, int j, double *maxval)
{
    double best = 0.0;
    for (int k = j; k < n; ++k) best += A[IDX(k, j, n)] * A[IDX(k, j, n)];
    int ii = j;
    for (int i = j + 1; i < n; ++i) {
        double mi = 0.0;
        for (int k = j; k < n; ++k) mi += A[IDX(k, i, n)] * A[IDX(k, i, n)];
        if (fabs(mi) > best) { best = mi; ii = i; }
    }
    *maxval = best;
    return ii;
}

/* UDT.jl:157-172 */
static double reflector(int n, double *x, double normu, int j)
{
    double xi1 = x[IDX(j, j, n)];
    if (normu == 0.0) return 0.0;
    normu = sqrt(normu);
    const double nu = copysign(normu, xi1);
    xi1 += nu;
    x[IDX(j, j, n)] = -nu;
    for (int i = j + 1; i < n; ++i) x[IDX(i, j, n)] /= xi1;
    return xi1 / nu;
}

/* UDT.jl:53-70 */
static void reflector_apply(int n, double *M, double tau, int k)
{
    const double *v = M + IDX(0, k, n);
    for (int j = k + 1; j < n; ++j) {
        double *c = M + IDX(0, j, n);
        double vAj = c[k];
        for (int i = k + 1; i < n; ++i) vAj += v[i] * c[i];
        vAj = tau * vAj;
        c[k] -= vAj;
        for (int i = k + 1; i < n; ++i) c[i] -= v[i] * vAj;
    }
}

/* UDT.jl:216-334.  apply_pivot != 0 <=> Val(true).  pivot is 0-based here. */
void ref_udt_pivot(int n, double *U, double *D, double *input, int64_t *pivot, double *temp,
                   int apply_pivot)
{
    for (int i = 0; i < n; ++i) pivot[i] = i;

    for (int j = 0; j < n; ++j) {
        double maxval;
        const int jm = indmaxcolumn(n, input, j, &maxval);
        if (jm != j) {
            const int64_t tp = pivot[jm]; pivot[jm] = pivot[j]; pivot[j] = tp;
            for (int i = 0; i < n; ++i) {
                const double t = input[IDX(i, jm, n)];
                input[IDX(i, jm, n)] = input[IDX(i, j, n)];
                input[IDX(i, j, n)] = t;
            }
        }
        const double tau = reflector(n, input, maxval, j);
        temp[j] = tau;
        reflector_apply(n, input, tau, j);
    }

    /* "Calculate Q", UDT.jl:272-288 */
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) U[IDX(i, j, n)] = (i == j) ? 1.0 : 0.0;
    U[IDX(n - 1, n - 1, n)] -= temp[n - 1];
    for (int k = n - 2; k >= 0; --k) {
        const double *v = input + IDX(0, k, n);
        for (int j = k; j < n; ++j) {
            double *c = U + IDX(0, j, n);
            double vBj = c[k];
            for (int i = k + 1; i < n; ++i) vBj += v[i] * c[i];
            vBj = temp[k] * vBj;
            c[k] -= vBj;
            for (int i = k + 1; i < n; ++i) c[i] -= v[i] * vBj;
        }
    }

    /* "Calculate D", UDT.jl:293-301 */
    for (int i = 0; i < n; ++i) {
        const double x = fabs(input[IDX(i, i, n)]);
        D[i] = (x == 0.0) ? 1.0 : x;
    }

    if (apply_pivot) {
        /* UDT.jl:311-324 */
        for (int i = 0; i < n; ++i) {
            const double d = 1.0 / D[i];
            for (int j = 0; j < i; ++j) temp[pivot[j]] = 0.0;
            for (int j = i; j < n; ++j) temp[pivot[j]] = d * input[IDX(i, j, n)];
            for (int j = 0; j < n; ++j) input[IDX(i, j, n)] = temp[j];
        }
    } else {
        /* UDT.jl:325-333: "dirty" upper triangle */
        for (int i = 0; i < n; ++i) {
            const double d = 1.0 / D[i];
            for (int j = i; j < n; ++j) input[IDX(i, j, n)] = d * input[IDX(i, j, n)];
        }
    }
}

/* real.jl:198-226: A <- A[:, pivot] * inv(triu(T)); O is scratch */
void ref_rdivp(int n, double *A, const double *T, double *O, const int64_t *pivot)
{
    for (int j = 0; j < n; ++j) {
        const int64_t p = pivot[j];
        for (int i = 0; i < n; ++i) O[IDX(i, j, n)] = A[IDX(i, p, n)];
    }
    for (int i = 0; i < n; ++i) A[IDX(i, 0, n)] = O[IDX(i, 0, n)] / T[IDX(0, 0, n)];
    for (int j = 1; j < n; ++j) {
        double *aj = A + IDX(0, j, n);
        const double *oj = O + IDX(0, j, n);
        for (int i = 0; i < n; ++i) aj[i] = oj[i];
        for (int k = 0; k < j; ++k) {
            const double t = T[IDX(k, j, n)];
            const double *ak = A + IDX(0, k, n);
            for (int i = 0; i < n; ++i) aj[i] -= ak[i] * t;
        }
        const double tjj = T[IDX(j, j, n)];
        for (int i = 0; i < n; ++i) aj[i] /= tjj;
    }
}

/* stack.jl:442-496 for ONE block; destroys all six inputs. */
void ref_calculate_greens_block(int n, double *Ul, double *Dl, double *Tl, double *Ur, double *Dr,
                                double *Tr, double *G, int64_t *pivot, double *temp)
{
    vmul_nt(n, G, Tl, Tr);                 /* G  = Tl * Tr'            :450 */
    vmul_mat_diag(n, Tr, G, Dr);           /* Tr = G * Diagonal(Dr)    :451 */
    vmul_diag_mat(n, G, Dl, Tr);           /* G  = Diagonal(Dl) * Tr   :452 */
    ref_udt_pivot(n, Tr, Dr, G, pivot, temp, 0);                    /* :453 */

    vmul_nn(n, Tl, Ul, Tr);                /* Tl = Ul * Tr             :464 */
    ref_rdivp(n, Ur, G, Ul, pivot);        /* Ur = Ur / G              :465 */
    vmul_tn(n, Tr, Tl, Ur);                /* Tr = Tl' * Ur            :466 */

    for (int i = 0; i < n; ++i) Tr[IDX(i, i, n)] += Dr[i];          /* :472 */

    ref_udt_pivot(n, Ul, Dr, Tr, pivot, temp, 0);                   /* :480 */
    ref_rdivp(n, Ur, Tr, G, pivot);                                 /* :481 */
    vmul_nn(n, Tr, Tl, Ul);                /* Tr = Tl * Ul             :482 */

    for (int i = 0; i < n; ++i) Dl[i] = 1.0 / Dr[i];                /* :486 */

    vmul_mat_diag(n, Ul, Ur, Dl);          /* Ul = Ur * Diagonal(Dl)   :492 */
    vmul_nt(n, G, Ul, Tr);                 /* G  = Ul * Tr'            :493 */
}

/* ------------------------------------------------------------------------ */
/* The chain object: DQMCStack (stack.jl:1-74, 160-206) + field + analysis   */
/* ------------------------------------------------------------------------ */

typedef struct ref_stats {
    /* MagnitudeStats (statistics.jl:9-38): count, sum of log10|x|, min, max */
    int64_t neg_count;  double neg_sumlog, neg_min, neg_max;
    int64_t prop_count; double prop_sumlog, prop_min, prop_max;
} ref_stats;

typedef struct ref_chain {
    int N, M, nb, kind, C;               /* kind: 0 DensityHirsch, 1 MagneticHirsch, 2 DensityGHQ, 3 MagneticGHQ */
    int *rfirst, *rlast;                 /* 1-based inclusive ranges (stack.jl:154-158) */
    double alpha;
    double eta[4], gam[4];               /* GHQ nodes / weights (fields.jl:517-519, 579-581) */
    const double *eT2, *eT2i, *eTh, *eThi; /* exp(-dt T), exp(+dt T), exp(-dt T/2), exp(+dt T/2) */
    int8_t *conf;                        /* N x M, +-1 (fields.jl:363-368) or 1..4 (GHQ, fields.jl:471) */
    double *u_stack, *d_stack, *t_stack; /* (C+1) slots, each nb blocks */
    double *greens, *greens_temp, *Ul, *Ur, *Tl, *Tr, *tmp1, *tmp2, *curr_U;
    double *Dl, *Dr, *eV, *tempv;
    double *cIG, *cG;                    /* StandardFieldCache IG, G (fields.jl:1-48) */
    int64_t *pivot;
    int current_slice, current_range, direction;
    int check_sign_problem, check_propagation_error;
    uint64_t seed; int64_t chain_id; int64_t sweep_index;
    const double *uniforms;              /* optional table [2M][N] for the current sweep */
    int64_t step_in_sweep;
    ref_stats stats;
    void *ut;                            /* UnequalTimeStack (dqmc_ref_ut.inc.c), lazily allocated */
} ref_chain;

static void ut_free(ref_chain *c);

static double *dalloc(size_t k) { return (double *)calloc(k ? k : 1, sizeof(double)); }

static void set_identity(int n, int nb, double *A)
{
    memset(A, 0, sizeof(double) * (size_t)n * n * nb);
    for (int b = 0; b < nb; ++b)
        for (int i = 0; i < n; ++i) A[(size_t)b * n * n + IDX(i, i, n)] = 1.0;
}
static void set_ones(int len, double *d) { for (int i = 0; i < len; ++i) d[i] = 1.0; }

#define MAT(c, base, slot) ((base) + (size_t)(slot) * (c)->N * (c)->N * (c)->nb)
#define VEC(c, base, slot) ((base) + (size_t)(slot) * (c)->N * (c)->nb)

ref_chain *ref_chain_create(int N, int M, int nb, int kind, int C, const int *rfirst,
                            const int *rlast, double alpha, const double *eT2, const double *eT2i,
                            const double *eTh, const double *eThi, int check_sign,
                            int check_prop, uint64_t seed, int64_t chain_id)
{
    ref_chain *c = (ref_chain *)calloc(1, sizeof(ref_chain));
    c->N = N; c->M = M; c->nb = nb; c->kind = kind; c->C = C;
    c->rfirst = (int *)malloc(sizeof(int) * C); c->rlast = (int *)malloc(sizeof(int) * C);
    memcpy(c->rfirst, rfirst, sizeof(int) * C); memcpy(c->rlast, rlast, sizeof(int) * C);
    c->alpha = alpha; c->eT2 = eT2; c->eT2i = eT2i; c->eTh = eTh; c->eThi = eThi;
    dqmc_ghq_tables(c->eta, c->gam);
    const size_t nn = (size_t)N * N * nb, nv = (size_t)N * nb;
    c->conf = (int8_t *)malloc((size_t)N * M);
    for (size_t i = 0; i < (size_t)N * M; ++i) c->conf[i] = 1;
    c->u_stack = dalloc(nn * (C + 1)); c->t_stack = dalloc(nn * (C + 1));
    c->d_stack = dalloc(nv * (C + 1));
    c->greens = dalloc(nn); c->greens_temp = dalloc(nn);
    c->Ul = dalloc(nn); c->Ur = dalloc(nn); c->Tl = dalloc(nn); c->Tr = dalloc(nn);
    c->tmp1 = dalloc(nn); c->tmp2 = dalloc(nn); c->curr_U = dalloc(nn);
    c->Dl = dalloc(nv); c->Dr = dalloc(nv); c->eV = dalloc(nv); c->tempv = dalloc(nv);
    c->cIG = dalloc(nv); c->cG = dalloc(nv);
    c->pivot = (int64_t *)calloc(nv, sizeof(int64_t));
    /* initialize_stack, stack.jl:186-191 */
    set_identity(N, nb, c->Ul); set_identity(N, nb, c->Ur);
    set_identity(N, nb, c->Tl); set_identity(N, nb, c->Tr);
    set_ones((int)nv, c->Dl); set_ones((int)nv, c->Dr);
    c->check_sign_problem = check_sign; c->check_propagation_error = check_prop;
    c->seed = seed; c->chain_id = chain_id; c->sweep_index = 0;
    c->stats.neg_min = c->stats.prop_min = INFINITY;
    c->stats.neg_max = c->stats.prop_max = -INFINITY;
    c->current_slice = 0; c->current_range = 1; c->direction = 1;
    return c;
}

void ref_chain_destroy(ref_chain *c)
{
    if (!c) return;
    ut_free(c);
    free(c->rfirst); free(c->rlast); free(c->conf); free(c->u_stack); free(c->t_stack);
    free(c->d_stack); free(c->greens); free(c->greens_temp); free(c->Ul); free(c->Ur);
    free(c->Tl); free(c->Tr); free(c->tmp1); free(c->tmp2); free(c->curr_U); free(c->Dl);
    free(c->Dr); free(c->eV); free(c->tempv); free(c->cIG); free(c->cG); free(c->pivot);
    free(c);
}

/* fields.jl:380-386 (density: every block +p*alpha*x) and :429-438 (magnetic:
 * block 1 +p*alpha*x, block 2 -p*alpha*x).  slice is 1-based. */
static void interaction_matrix_exp(ref_chain *c, int slice, double power)
{
    const int N = c->N;
    const int8_t *x = c->conf + (size_t)(slice - 1) * N;
    if (c->kind >= 2) {
        /* fields.jl:533-546 (magnetic GHQ: +eta in block 1, -eta in block 2), :596-602 (density GHQ) */
        for (int i = 0; i < N; ++i) c->eV[i] = exp(power * c->alpha * c->eta[x[i] - 1]);
        if (c->nb == 2)
            for (int i = 0; i < N; ++i) c->eV[N + i] = exp(-power * c->alpha * c->eta[x[i] - 1]);
        return;
    }
    for (int i = 0; i < N; ++i) c->eV[i] = exp(power * c->alpha * (double)x[i]);
    if (c->nb == 2) {
        const double s = (c->kind == 1) ? -1.0 : 1.0;
        for (int i = 0; i < N; ++i) c->eV[N + i] = exp(s * power * c->alpha * (double)x[i]);
    }
}

/* stack.jl:319-327   M <- eT2 * (eV * M) */
static void multiply_slice_matrix_left(ref_chain *c, int slice, double *Mx)
{
    const int n = c->N; const size_t nn = (size_t)n * n;
    interaction_matrix_exp(c, slice, 1.0);
    for (int b = 0; b < c->nb; ++b) {
        vmul_diag_mat(n, c->tmp1 + b * nn, c->eV + b * n, Mx + b * nn);
        vmul_nn(n, Mx + b * nn, c->eT2, c->tmp1 + b * nn);
    }
}
/* stack.jl:329-337   M <- (M * eT2) * eV */
static void multiply_slice_matrix_right(ref_chain *c, int slice, double *Mx)
{
    const int n = c->N; const size_t nn = (size_t)n * n;
    interaction_matrix_exp(c, slice, 1.0);
    for (int b = 0; b < c->nb; ++b) {
        vmul_nn(n, c->tmp1 + b * nn, Mx + b * nn, c->eT2);
        vmul_mat_diag(n, Mx + b * nn, c->tmp1 + b * nn, c->eV + b * n);
    }
}
/* stack.jl:339-347   M <- (M * eV^-1) * eT2^-1 */
static void multiply_slice_matrix_inv_right(ref_chain *c, int slice, double *Mx)
{
    const int n = c->N; const size_t nn = (size_t)n * n;
    interaction_matrix_exp(c, slice, -1.0);
    for (int b = 0; b < c->nb; ++b) {
        vmul_mat_diag(n, c->tmp1 + b * nn, Mx + b * nn, c->eV + b * n);
        vmul_nn(n, Mx + b * nn, c->tmp1 + b * nn, c->eT2i);
    }
}
/* stack.jl:349-357   M <- eV^-1 * (eT2^-1 * M) */
static void multiply_slice_matrix_inv_left(ref_chain *c, int slice, double *Mx)
{
    const int n = c->N; const size_t nn = (size_t)n * n;
    interaction_matrix_exp(c, slice, -1.0);
    for (int b = 0; b < c->nb; ++b) {
        vmul_nn(n, c->tmp1 + b * nn, c->eT2i, Mx + b * nn);
        vmul_diag_mat(n, Mx + b * nn, c->eV + b * n, c->tmp1 + b * nn);
    }
}
/* stack.jl:359-367   M <- eV' * (eT2' * M) */
static void multiply_daggered_slice_matrix_left(ref_chain *c, int slice, double *Mx)
{
    const int n = c->N; const size_t nn = (size_t)n * n;
    interaction_matrix_exp(c, slice, 1.0);
    for (int b = 0; b < c->nb; ++b) {
        vmul_tn(n, c->tmp1 + b * nn, c->eT2, Mx + b * nn);
        vmul_diag_mat(n, Mx + b * nn, c->eV + b * n, c->tmp1 + b * nn);
    }
}

static void udt_blocks(ref_chain *c, double *U, double *D, double *input, int apply_pivot)
{
    const int n = c->N; const size_t nn = (size_t)n * n;
    for (int b = 0; b < c->nb; ++b)   /* blockdiagonal.jl:352-370: per-block pivoting */
        ref_udt_pivot(n, U + b * nn, D + b * n, input + b * nn, c->pivot + b * n, c->tempv + b * n,
                      apply_pivot);
}

/* stack.jl:377-393; idx is the 1-based range index */
static void add_slice_sequence_left(ref_chain *c, int idx)
{
    const int n = c->N; const size_t nn = (size_t)n * n, tot = nn * c->nb;
    memcpy(c->curr_U, MAT(c, c->u_stack, idx - 1), sizeof(double) * tot);
    for (int s = c->rfirst[idx - 1]; s <= c->rlast[idx - 1]; ++s)
        multiply_slice_matrix_left(c, s, c->curr_U);
    for (int b = 0; b < c->nb; ++b)
        vmul_mat_diag(n, c->tmp1 + b * nn, c->curr_U + b * nn, VEC(c, c->d_stack, idx - 1) + b * n);
    udt_blocks(c, MAT(c, c->u_stack, idx), VEC(c, c->d_stack, idx), c->tmp1, 1);
    for (int b = 0; b < c->nb; ++b)
        vmul_nn(n, MAT(c, c->t_stack, idx) + b * nn, c->tmp1 + b * nn,
                MAT(c, c->t_stack, idx - 1) + b * nn);
}

/* stack.jl:402-416 */
static void add_slice_sequence_right(ref_chain *c, int idx)
{
    const int n = c->N; const size_t nn = (size_t)n * n, tot = nn * c->nb;
    memcpy(c->curr_U, MAT(c, c->u_stack, idx), sizeof(double) * tot);
    for (int s = c->rlast[idx - 1]; s >= c->rfirst[idx - 1]; --s)
        multiply_daggered_slice_matrix_left(c, s, c->curr_U);
    for (int b = 0; b < c->nb; ++b)
        vmul_mat_diag(n, c->tmp1 + b * nn, c->curr_U + b * nn, VEC(c, c->d_stack, idx) + b * n);
    udt_blocks(c, MAT(c, c->u_stack, idx - 1), VEC(c, c->d_stack, idx - 1), c->tmp1, 1);
    for (int b = 0; b < c->nb; ++b)
        vmul_nn(n, MAT(c, c->t_stack, idx - 1) + b * nn, c->tmp1 + b * nn,
                MAT(c, c->t_stack, idx) + b * nn);
}

/* stack.jl:509-516 */
static void calculate_greens(ref_chain *c, double *out)
{
    const int n = c->N; const size_t nn = (size_t)n * n;
    for (int b = 0; b < c->nb; ++b)
        ref_calculate_greens_block(n, c->Ul + b * nn, c->Dl + b * n, c->Tl + b * nn,
                                   c->Ur + b * nn, c->Dr + b * n, c->Tr + b * nn, out + b * nn,
                                   c->pivot + b * n, c->tempv + b * n);
}

/* stack.jl:594-603 */
static void wrap_greens(ref_chain *c, double *gf, int curr_slice, int direction)
{
    if (direction == -1) {
        multiply_slice_matrix_inv_left(c, curr_slice - 1, gf);
        multiply_slice_matrix_right(c, curr_slice - 1, gf);
    } else {
        multiply_slice_matrix_left(c, curr_slice, gf);
        multiply_slice_matrix_inv_right(c, curr_slice, gf);
    }
}

static void push_prop_error(ref_chain *c, double v)
{
    ref_stats *s = &c->stats;
    s->prop_count++; s->prop_sumlog += log10(fabs(v));
    if (v < s->prop_min) s->prop_min = v;
    if (v > s->prop_max) s->prop_max = v;
}

static double max_abs_diff(size_t len, const double *a, const double *b)
{
    double m = 0.0;
    for (size_t i = 0; i < len; ++i) { const double d = fabs(a[i] - b[i]); if (d > m) m = d; }
    return m;
}

/* stack.jl:257-281 */
void ref_build_stack(ref_chain *c)
{
    const size_t tot = (size_t)c->N * c->N * c->nb, nv = (size_t)c->N * c->nb;
    set_identity(c->N, c->nb, MAT(c, c->u_stack, 0));
    set_ones((int)nv, VEC(c, c->d_stack, 0));
    set_identity(c->N, c->nb, MAT(c, c->t_stack, 0));
    for (int i = 1; i <= c->C; ++i) add_slice_sequence_left(c, i);
    c->current_slice = c->M + 1; c->current_range = c->C; c->direction = -1;
    memcpy(c->Ul, MAT(c, c->u_stack, c->C), sizeof(double) * tot);
    memcpy(c->Dl, VEC(c, c->d_stack, c->C), sizeof(double) * nv);
    memcpy(c->Tl, MAT(c, c->t_stack, c->C), sizeof(double) * tot);
    set_identity(c->N, c->nb, c->Ur); set_ones((int)nv, c->Dr); set_identity(c->N, c->nb, c->Tr);
    calculate_greens(c, c->greens);
}

/* stack.jl:284-308 */
void ref_reverse_build_stack(ref_chain *c)
{
    const size_t tot = (size_t)c->N * c->N * c->nb, nv = (size_t)c->N * c->nb;
    set_identity(c->N, c->nb, MAT(c, c->u_stack, c->C));
    set_ones((int)nv, VEC(c, c->d_stack, c->C));
    set_identity(c->N, c->nb, MAT(c, c->t_stack, c->C));
    for (int i = c->C; i >= 1; --i) add_slice_sequence_right(c, i);
    c->current_slice = 0; c->current_range = 1; c->direction = 1;
    set_identity(c->N, c->nb, c->Ul); set_ones((int)nv, c->Dl); set_identity(c->N, c->nb, c->Tl);
    memcpy(c->Ur, MAT(c, c->u_stack, 0), sizeof(double) * tot);
    memcpy(c->Dr, VEC(c, c->d_stack, 0), sizeof(double) * nv);
    memcpy(c->Tr, MAT(c, c->t_stack, 0), sizeof(double) * tot);
    calculate_greens(c, c->greens);
}

/* stack.jl:605-730 */
void ref_propagate(ref_chain *c)
{
    const size_t tot = (size_t)c->N * c->N * c->nb, nv = (size_t)c->N * c->nb;
    c->current_slice += c->direction;

    if (c->direction == 1) {
        if (c->current_slice == 1) {
            set_identity(c->N, c->nb, MAT(c, c->u_stack, 0));
            set_ones((int)nv, VEC(c, c->d_stack, 0));
            set_identity(c->N, c->nb, MAT(c, c->t_stack, 0));
        } else if (c->current_slice - 1 == c->rlast[c->current_range - 1]) {
            const int idx = c->current_range;
            memcpy(c->Ur, MAT(c, c->u_stack, idx), sizeof(double) * tot);
            memcpy(c->Dr, VEC(c, c->d_stack, idx), sizeof(double) * nv);
            memcpy(c->Tr, MAT(c, c->t_stack, idx), sizeof(double) * tot);
            add_slice_sequence_left(c, idx);
            memcpy(c->Ul, MAT(c, c->u_stack, idx), sizeof(double) * tot);
            memcpy(c->Dl, VEC(c, c->d_stack, idx), sizeof(double) * nv);
            memcpy(c->Tl, MAT(c, c->t_stack, idx), sizeof(double) * tot);

            if (c->check_propagation_error) {
                memcpy(c->greens_temp, c->greens, sizeof(double) * tot);
                /* stack.jl:638-640 wraps greens_temp unconditionally; it is only
                 * consumed by the check, so it is skipped when the check is off. */
                wrap_greens(c, c->greens_temp, c->current_slice - 1, 1);
            }
            calculate_greens(c, c->greens);
            if (c->check_propagation_error) {
                const double d = max_abs_diff(tot, c->greens_temp, c->greens);
                if (d > 1e-7) push_prop_error(c, d);
            }
            if (c->current_range == c->C) {
                c->direction = -1;
                ref_propagate(c);
            } else {
                c->current_range += 1;
            }
        } else {
            wrap_greens(c, c->greens, c->current_slice - 1, 1);
        }
    } else {
        if (c->current_slice == c->M) {
            set_identity(c->N, c->nb, MAT(c, c->u_stack, c->C));
            set_ones((int)nv, VEC(c, c->d_stack, c->C));
            set_identity(c->N, c->nb, MAT(c, c->t_stack, c->C));
            wrap_greens(c, c->greens, c->current_slice + 1, -1);
        } else if (c->current_slice + 1 == c->rfirst[c->current_range - 1]) {
            const int idx = c->current_range;
            memcpy(c->Ul, MAT(c, c->u_stack, idx - 1), sizeof(double) * tot);
            memcpy(c->Dl, VEC(c, c->d_stack, idx - 1), sizeof(double) * nv);
            memcpy(c->Tl, MAT(c, c->t_stack, idx - 1), sizeof(double) * tot);
            add_slice_sequence_right(c, idx);
            memcpy(c->Ur, MAT(c, c->u_stack, idx - 1), sizeof(double) * tot);
            memcpy(c->Dr, VEC(c, c->d_stack, idx - 1), sizeof(double) * nv);
            memcpy(c->Tr, MAT(c, c->t_stack, idx - 1), sizeof(double) * tot);

            if (c->check_propagation_error)
                memcpy(c->greens_temp, c->greens, sizeof(double) * tot);
            calculate_greens(c, c->greens);
            if (c->check_propagation_error) {
                const double d = max_abs_diff(tot, c->greens_temp, c->greens);
                if (d > 1e-7) push_prop_error(c, d);
            }
            if (c->current_range == 1) {
                c->direction = 1;
                ref_propagate(c);
            } else {
                wrap_greens(c, c->greens, c->current_slice + 1, -1);
                c->current_range -= 1;
            }
        } else {
            wrap_greens(c, c->greens, c->current_slice + 1, -1);
        }
    }
}

/* stack.jl:525-583: G at `slice` from scratch, stabilising when k % safe_mult == 0 */
void ref_calculate_greens_at(ref_chain *c, int slice, int safe_mult, double *out)
{
    const int n = c->N; const size_t nn = (size_t)n * n, tot = nn * c->nb, nv = (size_t)n * c->nb;
    set_identity(n, c->nb, c->curr_U);
    set_identity(n, c->nb, c->Ur); set_ones((int)nv, c->Dr); set_identity(n, c->nb, c->Tr);
    if (slice + 1 <= c->M) {
        for (int k = c->M; k >= slice + 1; --k) {
            multiply_daggered_slice_matrix_left(c, k, c->curr_U);
            if (k % safe_mult == 0) {
                for (int b = 0; b < c->nb; ++b)
                    vmul_mat_diag(n, c->tmp1 + b * nn, c->curr_U + b * nn, c->Dr + b * n);
                udt_blocks(c, c->curr_U, c->Dr, c->tmp1, 1);
                memcpy(c->tmp2, c->Tr, sizeof(double) * tot);
                for (int b = 0; b < c->nb; ++b)
                    vmul_nn(n, c->Tr + b * nn, c->tmp1 + b * nn, c->tmp2 + b * nn);
            }
        }
        for (int b = 0; b < c->nb; ++b)
            vmul_mat_diag(n, c->tmp1 + b * nn, c->curr_U + b * nn, c->Dr + b * n);
        udt_blocks(c, c->Ur, c->Dr, c->tmp1, 1);
        memcpy(c->tmp2, c->Tr, sizeof(double) * tot);
        for (int b = 0; b < c->nb; ++b)
            vmul_nn(n, c->Tr + b * nn, c->tmp1 + b * nn, c->tmp2 + b * nn);
    }
    set_identity(n, c->nb, c->curr_U);
    set_identity(n, c->nb, c->Ul); set_ones((int)nv, c->Dl); set_identity(n, c->nb, c->Tl);
    if (slice >= 1) {
        for (int k = 1; k <= slice; ++k) {
            multiply_slice_matrix_left(c, k, c->curr_U);
            if (k % safe_mult == 0) {
                for (int b = 0; b < c->nb; ++b)
                    vmul_mat_diag(n, c->tmp1 + b * nn, c->curr_U + b * nn, c->Dl + b * n);
                udt_blocks(c, c->curr_U, c->Dl, c->tmp1, 1);
                memcpy(c->tmp2, c->Tl, sizeof(double) * tot);
                for (int b = 0; b < c->nb; ++b)
                    vmul_nn(n, c->Tl + b * nn, c->tmp1 + b * nn, c->tmp2 + b * nn);
            }
        }
        for (int b = 0; b < c->nb; ++b)
            vmul_mat_diag(n, c->tmp1 + b * nn, c->curr_U + b * nn, c->Dl + b * n);
        udt_blocks(c, c->Ul, c->Dl, c->tmp1, 1);
        memcpy(c->tmp2, c->Tl, sizeof(double) * tot);
        for (int b = 0; b < c->nb; ++b)
            vmul_nn(n, c->Tl + b * nn, c->tmp1 + b * nn, c->tmp2 + b * nn);
    }
    calculate_greens(c, out);
}

/* ------------------------------------------------------------------------ */
/* local updates: local_updates.jl:7-60, fields.jl:63-84, 271-286, 340-344,  */
/* 388-393, 440-449; linalg/updates.jl:7-11, 48-54, 92-97                     */
/* ------------------------------------------------------------------------ */

/* one proposal at (site i 0-based, current slice); returns p and fills R/Delta; *x_new = the proposed field value
 * (the reference's `passthrough`); u_choice: the uniform behind `rand(1:3)` of the GHQ fields */
static double propose_local(const ref_chain *c, int i, double *Delta, double *R, int8_t *x_new, double u_choice)
{
    const int n = c->N; const size_t nn = (size_t)n * n;
    const int8_t xi = c->conf[(size_t)(c->current_slice - 1) * n + i];
    if (c->kind >= 2) {
        /* fields.jl:525-556 (magnetic), :587-610 (density) */
        const int xo = xi, xn = dqmc_ghq_choice(xo, u_choice);
        *x_new = (int8_t)xn;
        const double dEb = c->alpha * (c->eta[xn - 1] - c->eta[xo - 1]);
        const double exp_ratio = exp(dEb);
        double detratio;
        if (c->kind == 3) {
            Delta[0] = exp_ratio - 1.0; Delta[1] = 1.0 / exp_ratio - 1.0;
            R[0] = 1.0 + Delta[0] * (1.0 - c->greens[IDX(i, i, n)]);
            R[1] = 1.0 + Delta[1] * (1.0 - c->greens[nn + IDX(i, i, n)]);
            detratio = R[0] * R[1];
            return detratio * c->gam[xn - 1] / c->gam[xo - 1];          /* exp(-0.0) * ... */
        }
        Delta[0] = exp_ratio - 1.0;
        R[0] = 1.0 + Delta[0] * (1.0 - c->greens[IDX(i, i, n)]);
        detratio = R[0] * R[0];
        return exp(-dEb) * (detratio * c->gam[xn - 1] / c->gam[xo - 1]);  /* local_updates.jl:31 */
    }
    const double x = (double)xi;
    *x_new = (int8_t)(-xi);
    const double dE = -2.0 * c->alpha * x;
    if (c->kind == 0) {
        /* fields.jl:388-393 + :63-66 (nb==1) / :68-75 (nb==2, scalar Delta) */
        double det = 1.0;
        for (int b = 0; b < c->nb; ++b) {
            Delta[b] = exp(dE) - 1.0;
            R[b] = 1.0 + Delta[b] * (1.0 - c->greens[b * nn + IDX(i, i, n)]);
            det *= R[b];
        }
        if (c->nb == 1) det = R[0] * R[0];
        return exp(-dE) * det;     /* local_updates.jl:31 */
    }
    /* fields.jl:440-449 + :77-84 */
    Delta[0] = exp(+dE) - 1.0; Delta[1] = exp(-dE) - 1.0;
    R[0] = 1.0 + Delta[0] * (1.0 - c->greens[IDX(i, i, n)]);
    R[1] = 1.0 + Delta[1] * (1.0 - c->greens[nn + IDX(i, i, n)]);
    return R[0] * R[1];            /* exp(-0.0) * detratio */
}

/* fields.jl:271-286 + 340-344 */
static void accept_local(ref_chain *c, int i, const double *Delta, const double *R, int8_t x_new)
{
    const int n = c->N; const size_t nn = (size_t)n * n;
    for (int b = 0; b < c->nb; ++b) {
        double *G = c->greens + b * nn, *IG = c->cIG + b * n, *g = c->cG + b * n;
        const double invRD = Delta[b] / R[b];                      /* vldiv22!  fields.jl:176-216 */
        for (int j = 0; j < n; ++j) IG[j] = -G[IDX(j, i, n)];      /* vsub!     updates.jl:48-54  */
        IG[i] += 1.0;
        for (int j = 0; j < n; ++j) g[j] = invRD * G[IDX(i, j, n)];/* vmul!     updates.jl:92-97  */
        for (int l = 0; l < n; ++l) {                              /* vsubkron! updates.jl:7-11   */
            const double gl = g[l];
            double *col = G + IDX(0, l, n);
            for (int k = 0; k < n; ++k) col[k] -= IG[k] * gl;
        }
    }
    c->conf[(size_t)(c->current_slice - 1) * n + i] = x_new;     /* fields.jl:342 / :497 */
}

/* tables: [2M][N] (Hirsch) or [2M][2][N] (GHQ: Metropolis uniforms, then choice uniforms) */
static double next_uniform(ref_chain *c, int site)
{
    const size_t uf = (c->kind >= 2) ? 2 : 1;
    if (c->uniforms) return c->uniforms[(size_t)c->step_in_sweep * c->N * uf + site];
    return dqmc_uniform(c->seed, (uint64_t)c->chain_id, (uint64_t)c->sweep_index,
                        (uint32_t)c->step_in_sweep, (uint32_t)site);
}
static double next_choice_uniform(ref_chain *c, int site)
{
    if (c->kind < 2) return 0.0;
    if (c->uniforms) return c->uniforms[(size_t)c->step_in_sweep * c->N * 2 + c->N + site];
    return dqmc_uniform_choice(c->seed, (uint64_t)c->chain_id, (uint64_t)c->sweep_index,
                               (uint32_t)c->step_in_sweep, (uint32_t)site);
}

/* local_updates.jl:23-60.  forced != NULL replays given accept decisions
 * (teacher forcing; one byte per site) instead of the Metropolis test;
 * probs != NULL receives p for every site. */
int ref_sweep_spatial(ref_chain *c, const uint8_t *forced, double *probs, uint8_t *decisions)
{
    int accepted = 0;
    double Delta[2], R[2];
    for (int i = 0; i < c->N; ++i) {
        int8_t x_new;
        const double p = propose_local(c, i, Delta, R, &x_new, next_choice_uniform(c, i));
        if (probs) probs[i] = p;
        if (c->check_sign_problem && p < 0.0) {       /* local_updates.jl:40-46 */
            ref_stats *s = &c->stats;
            s->neg_count++; s->neg_sumlog += log10(fabs(p));
            if (p < s->neg_min) s->neg_min = p;
            if (p > s->neg_max) s->neg_max = p;
        }
        int acc;
        if (forced) acc = forced[i] != 0;
        else acc = (p > 1.0) || (next_uniform(c, i) < p);   /* :53 */
        if (decisions) decisions[i] = (uint8_t)acc;
        if (acc) { accept_local(c, i, Delta, R, x_new); accepted++; }
    }
    return accepted;
}

/* local_updates.jl:7-14.  Optional trace buffers are [2M][N]. */
int64_t ref_local_sweep(ref_chain *c, const double *uniforms, const uint8_t *forced,
                        double *probs, uint8_t *decisions)
{
    int64_t accepted = 0;
    c->uniforms = uniforms;
    for (int step = 0; step < 2 * c->M; ++step) {
        c->step_in_sweep = step;
        accepted += ref_sweep_spatial(c, forced ? forced + (size_t)step * c->N : NULL,
                                      probs ? probs + (size_t)step * c->N : NULL,
                                      decisions ? decisions + (size_t)step * c->N : NULL);
        ref_propagate(c);
    }
    c->uniforms = NULL;
    c->sweep_index++;
    return accepted;
}

/* greens.jl:114-125: target = eThalf^-1 * (G_eff * eThalf) per block */
void ref_measured_greens(ref_chain *c, double *out)
{
    const int n = c->N; const size_t nn = (size_t)n * n;
    for (int b = 0; b < c->nb; ++b) {
        vmul_nn(n, c->curr_U + b * nn, c->greens + b * nn, c->eTh);
        vmul_nn(n, out + b * nn, c->eThi, c->curr_U + b * nn);
    }
}

/* ------------------------------------------------------------------------ */
/* thin accessors for ctypes                                                  */
/* ------------------------------------------------------------------------ */
void ref_set_conf(ref_chain *c, const int8_t *conf) { memcpy(c->conf, conf, (size_t)c->N * c->M); }
void ref_get_conf(const ref_chain *c, int8_t *conf) { memcpy(conf, c->conf, (size_t)c->N * c->M); }
void ref_get_greens(const ref_chain *c, double *G)
{ memcpy(G, c->greens, sizeof(double) * (size_t)c->N * c->N * c->nb); }
void ref_set_greens(ref_chain *c, const double *G)
{ memcpy(c->greens, G, sizeof(double) * (size_t)c->N * c->N * c->nb); }
void ref_get_state(const ref_chain *c, int *out3)
{ out3[0] = c->current_slice; out3[1] = c->current_range; out3[2] = c->direction; }
void ref_set_state(ref_chain *c, int slice, int range, int direction)
{ c->current_slice = slice; c->current_range = range; c->direction = direction; }
void ref_get_stats(const ref_chain *c, ref_stats *s) { *s = c->stats; }
void ref_set_sweep_index(ref_chain *c, int64_t s) { c->sweep_index = s; }
/* which: 0 u_stack, 1 d_stack, 2 t_stack, 3 Ul, 4 Dl, 5 Tl, 6 Ur, 7 Dr, 8 Tr, 9 greens_temp */
void ref_get_array(const ref_chain *c, int which, int slot, double *out)
{
    const size_t tot = (size_t)c->N * c->N * c->nb, nv = (size_t)c->N * c->nb;
    switch (which) {
    case 0: memcpy(out, MAT(c, c->u_stack, slot), sizeof(double) * tot); break;
    case 1: memcpy(out, VEC(c, c->d_stack, slot), sizeof(double) * nv); break;
    case 2: memcpy(out, MAT(c, c->t_stack, slot), sizeof(double) * tot); break;
    case 3: memcpy(out, c->Ul, sizeof(double) * tot); break;
    case 4: memcpy(out, c->Dl, sizeof(double) * nv); break;
    case 5: memcpy(out, c->Tl, sizeof(double) * tot); break;
    case 6: memcpy(out, c->Ur, sizeof(double) * tot); break;
    case 7: memcpy(out, c->Dr, sizeof(double) * nv); break;
    case 8: memcpy(out, c->Tr, sizeof(double) * tot); break;
    case 9: memcpy(out, c->greens_temp, sizeof(double) * tot); break;
    default: break;
    }
}

/* operator-level entry points (reference's linalg "operator API") */
void ref_vmul(int n, int ta, int tb, double *C, const double *A, const double *B)
{
    if (!ta && !tb) vmul_nn(n, C, A, B);
    else if (!ta && tb) vmul_nt(n, C, A, B);
    else if (ta && !tb) vmul_tn(n, C, A, B);
    else { /* real.jl:92-102  C = A' * B' */
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                double s = 0.0;
                for (int k = 0; k < n; ++k) s += A[IDX(k, i, n)] * B[IDX(j, k, n)];
                C[IDX(i, j, n)] = s;
            }
    }
}
/* which: 0 left, 1 right, 2 inv_right, 3 inv_left, 4 daggered_left  (stack.jl:319-367) */
void ref_multiply_slice_matrix(ref_chain *c, int which, int slice, double *Mx)
{
    switch (which) {
    case 0: multiply_slice_matrix_left(c, slice, Mx); break;
    case 1: multiply_slice_matrix_right(c, slice, Mx); break;
    case 2: multiply_slice_matrix_inv_right(c, slice, Mx); break;
    case 3: multiply_slice_matrix_inv_left(c, slice, Mx); break;
    case 4: multiply_daggered_slice_matrix_left(c, slice, Mx); break;
    default: break;
    }
}
void ref_wrap_greens(ref_chain *c, double *gf, int curr_slice, int direction)
{ wrap_greens(c, gf, curr_slice, direction); }
/* one proposal + optional accept at (site 0-based) on the current slice */
double ref_propose_local_choice(ref_chain *c, int site, int accept, double u_choice)
{
    double Delta[2], R[2];
    int8_t x_new;
    const double p = propose_local(c, site, Delta, R, &x_new, u_choice);
    if (accept) accept_local(c, site, Delta, R, x_new);
    return p;
}
double ref_propose_local(ref_chain *c, int site, int accept) { return ref_propose_local_choice(c, site, accept, 0.0); }

/* ------------------------------------------------------------------------ */
/* CPU baseline driver: one chain per thread (the reference is single-        */
/* threaded per simulation and parallel only across processes,               */
/* test/parallel.jl:62).  Returns elapsed seconds of the timed sweeps.        */
/* ------------------------------------------------------------------------ */
#include <pthread.h>
#include <time.h>
#include <unistd.h>

static double now_s(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int ref_max_threads(void)
{
    const long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct run_job {
    ref_chain **chains; int nchains; int phase; int warm; int nsweeps; int64_t *accepted;
    int next; pthread_mutex_t mu;
} run_job;

static void *run_worker(void *arg)
{
    run_job *j = (run_job *)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        const int b = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (b >= j->nchains) break;
        ref_chain *c = j->chains[b];
        if (j->phase == 0) {
            ref_reverse_build_stack(c);
            ref_propagate(c);
            for (int s = 0; s < j->warm; ++s) ref_local_sweep(c, NULL, NULL, NULL, NULL);
        } else {
            int64_t acc = 0;
            for (int s = 0; s < j->nsweeps; ++s) acc += ref_local_sweep(c, NULL, NULL, NULL, NULL);
            if (j->accepted) j->accepted[b] = acc;
        }
    }
    return NULL;
}

static void run_phase(run_job *j, int phase, int nthreads)
{
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    j->phase = phase; j->next = 0;
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, run_worker, j);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
}

double ref_run_chains(ref_chain **chains, int nchains, int nthreads, int warm, int nsweeps,
                      int64_t *accepted)
{
    run_job j; memset(&j, 0, sizeof(j));
    j.chains = chains; j.nchains = nchains; j.warm = warm; j.nsweeps = nsweeps; j.accepted = accepted;
    pthread_mutex_init(&j.mu, NULL);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nchains) nthreads = nchains;
    run_phase(&j, 0, nthreads);
    const double t0 = now_s();
    run_phase(&j, 1, nthreads);
    const double dt = now_s() - t0;
    pthread_mutex_destroy(&j.mu);
    return dt;
}

/* unequal-time Green's functions: UnequalTimeStack + CombinedGreensIterator */
#include "dqmc_ref_ut.inc.c"

/* global Metropolis updates */
#include "dqmc_ref_global.inc.c"
