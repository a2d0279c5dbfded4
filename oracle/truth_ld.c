/*
 * oracle/truth_ld.c -- extended-precision arbiter for the equal-time Green's function.
 *
 * TEST INFRASTRUCTURE ONLY (tests/ and bench.py's cpu_baseline leg).  Not a restatement of the reference's
 * algorithm: an INDEPENDENT evaluation of the quantity the reference defines,
 *
 *      G = [1 + B_M ... B_1]^-1,   B_l = exp(-dtau T) * diag(exp(+-alpha x_l))
 *
 * (src/flavors/DQMC/stack.jl:525-583 `calculate_greens(mc, 0)`; slice matrices stack.jl:319-327), carried out in
 * x87 `long double` (64-bit mantissa, eps = 1.1e-19) with its own stabilisation (pivoted Householder QR every
 * `chunk` slices; final inverse in the scale-separated form G = (Db^-1 U^T + Ds T)^-1 Db^-1 U^T with
 * Db = max(D, 1), Ds = min(D, 1)).  It plays the role the BigFloat evaluations play in the reference's own tests
 * (test/DQMC/unequal_time_stack.jl:176-304): both the double-precision oracle (oracle/dqmc_ref.c) and the CUDA
 * library are measured against it, so a parity tolerance can be stated as "device error vs truth" next to "oracle
 * error vs truth" instead of being guessed.  The double-precision inputs (hopping exponential, exp(+-alpha)) are
 * taken as exact.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef long double R;
#define IX(i, j, n) ((size_t)(i) + (size_t)(j) * (size_t)(n))

/* C = A * B, column-major n x n; columns of C are dealt to threads (each element is one fixed-order dot product, so the
 * result does not depend on the thread count) */
struct mm_job { int n, j0, j1; R *C; const R *A, *B; };
static void *mm_worker(void *arg)
{
    const struct mm_job *w = (const struct mm_job *)arg;
    const int n = w->n;
    int j = w->j0;
    for (; j + 4 <= w->j1; j += 4) {                     /* four columns of C per pass over A */
        R *c0 = w->C + IX(0, j, n), *c1 = c0 + n, *c2 = c1 + n, *c3 = c2 + n;
        for (int i = 0; i < n; ++i) c0[i] = c1[i] = c2[i] = c3[i] = 0.0L;
        for (int k = 0; k < n; ++k) {
            const R b0 = w->B[IX(k, j, n)], b1 = w->B[IX(k, j + 1, n)], b2 = w->B[IX(k, j + 2, n)], b3 = w->B[IX(k, j + 3, n)];
            const R *a = w->A + IX(0, k, n);
            for (int i = 0; i < n; ++i) { const R x = a[i]; c0[i] += x * b0; c1[i] += x * b1; c2[i] += x * b2; c3[i] += x * b3; }
        }
    }
    for (; j < w->j1; ++j) {
        R *c = w->C + IX(0, j, n);
        for (int i = 0; i < n; ++i) c[i] = 0.0L;
        for (int k = 0; k < n; ++k) {
            const R b = w->B[IX(k, j, n)];
            const R *a = w->A + IX(0, k, n);
            for (int i = 0; i < n; ++i) c[i] += a[i] * b;
        }
    }
    return NULL;
}
static void mm(int n, R *restrict C, const R *restrict A, const R *restrict B)
{
    long nt = sysconf(_SC_NPROCESSORS_ONLN);
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    if (nt > n / 8) nt = n / 8 > 0 ? n / 8 : 1;
    pthread_t th[64]; struct mm_job job[64];
    for (long t = 0; t < nt; ++t) {
        job[t] = (struct mm_job){n, (int)((long)n * t / nt), (int)((long)n * (t + 1) / nt), C, A, B};
        if (t + 1 < nt) pthread_create(&th[t], NULL, mm_worker, &job[t]);
    }
    mm_worker(&job[nt - 1]);
    for (long t = 0; t + 1 < nt; ++t) pthread_join(th[t], NULL);
}

/* columns [lo, hi) dealt to threads in contiguous ranges; every column's arithmetic is independent of the split */
typedef void (*col_fn)(int c, void *ctx);
struct pf_job { int lo, hi; col_fn f; void *ctx; };
static void *pf_worker(void *arg)
{
    const struct pf_job *w = (const struct pf_job *)arg;
    for (int c = w->lo; c < w->hi; ++c) w->f(c, w->ctx);
    return NULL;
}
static void parfor(int lo, int hi, col_fn f, void *ctx, long min_per_thread)
{
    long nt = sysconf(_SC_NPROCESSORS_ONLN);
    const long cnt = hi - lo;
    if (cnt <= 0) return;
    if (nt > 64) nt = 64;
    if (nt > cnt / min_per_thread) nt = cnt / min_per_thread;
    if (nt < 1) nt = 1;
    pthread_t th[64]; struct pf_job job[64];
    for (long t = 0; t < nt; ++t) {
        job[t] = (struct pf_job){lo + (int)(cnt * t / nt), lo + (int)(cnt * (t + 1) / nt), f, ctx};
        if (t + 1 < nt) pthread_create(&th[t], NULL, pf_worker, &job[t]);
    }
    pf_worker(&job[nt - 1]);
    for (long t = 0; t + 1 < nt; ++t) pthread_join(th[t], NULL);
}

struct refl_ctx { int n, j; R tau; R *W; const R *v; R *norms; };
/* W[:, c] <- H_j W[:, c] on rows j.., and the squared norm of its rows j+1.. (recomputed, not down-dated) */
static void refl_apply(int c, void *p)
{
    const struct refl_ctx *x = (const struct refl_ctx *)p;
    const int n = x->n, j = x->j;
    R *w = x->W + IX(0, c, n);
    const R *v = x->v;
    R s = w[j];
    for (int i = j + 1; i < n; ++i) s += v[i] * w[i];
    s *= x->tau;
    w[j] -= s;
    R nr = 0.0L;
    for (int i = j + 1; i < n; ++i) { w[i] -= v[i] * s; nr += w[i] * w[i]; }
    if (x->norms) x->norms[c] = nr;
}

/* column-pivoted Householder QR of W (overwritten): W P = Q R; returns Q, d = |diag R|, Tn = d^-1 R P^T */
static void qrp(int n, R *W, R *Q, R *d, R *Tn)
{
    int *perm = (int *)malloc(sizeof(int) * (size_t)n);
    R *tau = (R *)malloc(sizeof(R) * (size_t)n);
    R *norms = (R *)malloc(sizeof(R) * (size_t)n);
    for (int j = 0; j < n; ++j) {
        perm[j] = j;
        R sacc = 0.0L;
        for (int i = 0; i < n; ++i) sacc += W[IX(i, j, n)] * W[IX(i, j, n)];
        norms[j] = sacc;
    }
    for (int j = 0; j < n; ++j) {
        /* pivot: largest remaining column norm (recomputed from scratch during the previous step) */
        int best = j; R bn = -1.0L;
        for (int c = j; c < n; ++c) if (norms[c] > bn) { bn = norms[c]; best = c; }
        if (best != j) {
            for (int i = 0; i < n; ++i) { R t = W[IX(i, j, n)]; W[IX(i, j, n)] = W[IX(i, best, n)]; W[IX(i, best, n)] = t; }
            int t = perm[j]; perm[j] = perm[best]; perm[best] = t;
            norms[best] = norms[j];
        }
        /* reflector H = 1 - tau v v^T, v = (1, W[j+1:, j] / (x + nu)) */
        R x = W[IX(j, j, n)];
        R nrm = sqrtl(bn);
        if (nrm == 0.0L) { tau[j] = 0.0L; continue; }
        R nu = (x >= 0.0L) ? nrm : -nrm;
        x += nu;
        for (int i = j + 1; i < n; ++i) W[IX(i, j, n)] /= x;
        tau[j] = x / nu;
        W[IX(j, j, n)] = -nu;
        struct refl_ctx rc = {n, j, tau[j], W, W + IX(0, j, n), norms};
        parfor(j + 1, n, refl_apply, &rc, 1 << 20);   /* serial: a thread team per step costs more than it saves */
    }
    free(norms);
    /* Q = H_0 ... H_{n-1} */
    for (size_t e = 0; e < (size_t)n * n; ++e) Q[e] = 0.0L;
    for (int i = 0; i < n; ++i) Q[IX(i, i, n)] = 1.0L;
    for (int j = n - 1; j >= 0; --j) {
        if (tau[j] == 0.0L) continue;
        struct refl_ctx rc = {n, j, tau[j], Q, W + IX(0, j, n), NULL};
        parfor(j, n, refl_apply, &rc, 1 << 20);
    }
    for (int j = 0; j < n; ++j) { R a = fabsl(W[IX(j, j, n)]); d[j] = (a == 0.0L) ? 1.0L : a; }
    for (size_t e = 0; e < (size_t)n * n; ++e) Tn[e] = 0.0L;
    for (int c = 0; c < n; ++c)
        for (int i = 0; i <= c; ++i) Tn[IX(i, perm[c], n)] = W[IX(i, c, n)] / d[i];
    free(perm); free(tau);
}

/* solve A X = B in place (LU with partial pivoting); A, B n x n; B <- X */
static int lusolve(int n, R *A, R *B)
{
    for (int k = 0; k < n; ++k) {
        int p = k; R mx = fabsl(A[IX(k, k, n)]);
        for (int i = k + 1; i < n; ++i) if (fabsl(A[IX(i, k, n)]) > mx) { mx = fabsl(A[IX(i, k, n)]); p = i; }
        if (mx == 0.0L) return -1;
        if (p != k) {
            for (int j = 0; j < n; ++j) {
                R t = A[IX(k, j, n)]; A[IX(k, j, n)] = A[IX(p, j, n)]; A[IX(p, j, n)] = t;
                t = B[IX(k, j, n)]; B[IX(k, j, n)] = B[IX(p, j, n)]; B[IX(p, j, n)] = t;
            }
        }
        const R inv = 1.0L / A[IX(k, k, n)];
        for (int i = k + 1; i < n; ++i) A[IX(i, k, n)] *= inv;
        for (int j = k + 1; j < n; ++j) {
            const R a = A[IX(k, j, n)];
            for (int i = k + 1; i < n; ++i) A[IX(i, j, n)] -= A[IX(i, k, n)] * a;
        }
        for (int j = 0; j < n; ++j) {
            const R b = B[IX(k, j, n)];
            for (int i = k + 1; i < n; ++i) B[IX(i, j, n)] -= A[IX(i, k, n)] * b;
        }
    }
    for (int j = 0; j < n; ++j)
        for (int k = n - 1; k >= 0; --k) {
            const R x = B[IX(k, j, n)] / A[IX(k, k, n)];
            B[IX(k, j, n)] = x;
            for (int i = 0; i < k; ++i) B[IX(i, j, n)] -= A[IX(i, k, n)] * x;
        }
    return 0;
}

/* eT2: n x n col-major; ev: [M][n] diagonal of exp(V_l) for slice l = 1..M (row l - 1); G out n x n.
 * slice0: G(slice0) = [1 + B_slice0 ... B_1 B_M ... B_{slice0+1}]^-1 (slice0 = 0: the sweep-end Green's function). */
int truth_greens_ld(int n, int M, int chunk, int slice0, const double *eT2, const double *ev, double *G)
{
    const size_t nn = (size_t)n * n;
    R *E = (R *)malloc(sizeof(R) * nn), *U = (R *)malloc(sizeof(R) * nn), *T = (R *)malloc(sizeof(R) * nn);
    R *W = (R *)malloc(sizeof(R) * nn), *W2 = (R *)malloc(sizeof(R) * nn), *Tn = (R *)malloc(sizeof(R) * nn);
    R *D = (R *)malloc(sizeof(R) * (size_t)n), *dn = (R *)malloc(sizeof(R) * (size_t)n);
    if (!E || !U || !T || !W || !W2 || !Tn || !D || !dn) return -2;
    for (size_t e = 0; e < nn; ++e) { E[e] = (R)eT2[e]; U[e] = 0.0L; T[e] = 0.0L; }
    for (int i = 0; i < n; ++i) { U[IX(i, i, n)] = 1.0L; T[IX(i, i, n)] = 1.0L; D[i] = 1.0L; }
    int done = 0;
    while (done < M) {
        const int cnt = (M - done < chunk) ? (M - done) : chunk;
        memcpy(W, U, sizeof(R) * nn);
        for (int s = 0; s < cnt; ++s) {
            const int l = (slice0 + done + s) % M;          /* 0-based slice index, order slice0+1, ..., M, 1, ..., slice0 */
            const double *v = ev + (size_t)l * n;
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) W[IX(i, j, n)] *= (R)v[i];
            mm(n, W2, E, W);
            R *t = W; W = W2; W2 = t;
        }
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) W[IX(i, j, n)] *= D[j];
        qrp(n, W, U, dn, Tn);
        mm(n, W2, Tn, T);
        memcpy(T, W2, sizeof(R) * nn);
        memcpy(D, dn, sizeof(R) * (size_t)n);
        done += cnt;
    }
    /* A = Db^-1 U^T + Ds T ; B = Db^-1 U^T ; G = A^-1 B */
    for (int i = 0; i < n; ++i) {
        const R db = (D[i] > 1.0L) ? D[i] : 1.0L, ds = (D[i] > 1.0L) ? 1.0L : D[i];
        for (int j = 0; j < n; ++j) {
            const R b = U[IX(j, i, n)] / db;
            W2[IX(i, j, n)] = b;
            W[IX(i, j, n)] = b + ds * T[IX(i, j, n)];
        }
    }
    const int rc = lusolve(n, W, W2);
    for (size_t e = 0; e < nn; ++e) G[e] = (double)W2[e];
    free(E); free(U); free(T); free(W); free(W2); free(Tn); free(D); free(dn);
    return rc;
}
