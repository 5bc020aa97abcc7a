/*
 * kkt_oracle.c -- CPU restatement of OnePhase.jl's per-iteration KKT solve.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (onephase.jl_b200/,
 * include/) may link, import or execute this file.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker or the reported CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" at the CHOLMOD boundary.  The reference
 * reaches its arithmetic through Julia stdlib SuiteSparse/CHOLMOD (not vendored,
 * version fixed by the Julia binary: 1.6/1.7 => SuiteSparse 5.4-5.10) and holds
 * no numeric golden vectors for this path.  What the reference's
 * tests do pin -- relational properties (inertia==1 on I10 and I10+two 0.1 lower
 * entries, ls_solve! == ls_solve bitwise, LDLt vs Cholesky < 1e-9, lower-only ==
 * symmetrised < 1e-9; test/linear_system_solvers.jl:58-116) -- are checked in
 * tests/test_oracle.py, and this restatement is cross-checked there against dense
 * numpy Cholesky and scipy SuperLU.
 *
 * All indices are 0-based int64 in this file; the Python wrapper converts.
 * Compile with -ffp-contract=off so no FMA is formed.
 *
 * Reference lines restated (paths under /root/reference/src):
 *   assembly            kkt_system_solver/schur.jl:47-62, utils/eval.jl:85-87,132-134
 *   shift               kkt_system_solver/kkt_system_solver.jl:109-113, schur.jl:64-83
 *   Cholesky / PD flag  linear_system_solvers/julia.jl:28-46   (CHOLMOD: pivot <= 0 or NaN => not PD)
 *   LDLt / inertia      linear_system_solvers/julia.jl:47-90, linear_system_solvers.jl:48-91
 *   solve               linear_system_solvers/julia.jl:99-113
 *   delta loop          IPM/delta_strategy.jl:37-114, parameters.jl:147-158
 *   direction + refine  kkt_system_solver/schur.jl:89-182
 *   KKT error ("N err") kkt_system_solver/kkt_system_solver.jl:27-96
 *   H symmetric product utils/eval.jl:221-234
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

/* ------------------------------------------------------------------ */
/* small helpers                                                      */
/* ------------------------------------------------------------------ */
static void *xmalloc(size_t n) { void *p = malloc(n ? n : 1); return p; }

static int cmp_i64(const void *a, const void *b) {
    i64 x = *(const i64 *)a, y = *(const i64 *)b;
    return (x > y) - (x < y);
}

/* ------------------------------------------------------------------ */
/* Assembly:  Q = (J_T * diag(sig)) * J + H     (schur.jl:55)          */
/* Q holds both triangles of J'DJ plus the lower-triangular H.         */
/* ------------------------------------------------------------------ */
typedef struct {
    i64 n, nnz;
    i64 *colptr, *rowval;
    double *nzval;
} orc_csc;

void orc_csc_free(orc_csc *A) {
    if (!A) return;
    free(A->colptr); free(A->rowval); free(A->nzval); free(A);
}
i64 orc_csc_nnz(const orc_csc *A) { return A->nnz; }
void orc_csc_copy(const orc_csc *A, i64 *colptr, i64 *rowval, double *nzval) {
    memcpy(colptr, A->colptr, sizeof(i64) * (size_t)(A->n + 1));
    memcpy(rowval, A->rowval, sizeof(i64) * (size_t)A->nnz);
    memcpy(nzval, A->nzval, sizeof(double) * (size_t)A->nnz);
}

/* J is m x n CSC.  Returns Q (n x n CSC, rows sorted). */
orc_csc *orc_form_system(i64 n, i64 m,
                         const i64 *Jp, const i64 *Ji, const double *Jx,
                         const i64 *Hp, const i64 *Hi, const double *Hx,
                         const double *y, const double *s)
{
    i64 nnzJ = Jp[n];
    /* J_T as CSC (n x m): column k of J_T = row k of J; rows sorted because we
       sweep the columns of J in increasing order (Class_iterate.jl:339). */
    i64 *Tp = (i64 *)calloc((size_t)m + 2, sizeof(i64));
    i64 *Ti = (i64 *)xmalloc(sizeof(i64) * (size_t)nnzJ);
    double *Tx = (double *)xmalloc(sizeof(double) * (size_t)nnzJ);
    for (i64 p = 0; p < nnzJ; p++) Tp[Ji[p] + 2]++;
    for (i64 k = 0; k < m; k++) Tp[k + 2] += Tp[k + 1];
    for (i64 j = 0; j < n; j++)
        for (i64 p = Jp[j]; p < Jp[j + 1]; p++) {
            i64 k = Ji[p];
            i64 q = Tp[k + 1]++;
            Ti[q] = j;
            /* A = J_T * Diagonal(sig): rounded once (eval.jl:86, left-assoc) */
            double sig = y[k] / s[k];
            Tx[q] = Jx[p] * sig;
        }
    /* Gustavson C = A * J, column by column, k ascending */
    i64 cap = nnzJ * 4 + n + 16;
    orc_csc *Q = (orc_csc *)xmalloc(sizeof(orc_csc));
    Q->n = n;
    Q->colptr = (i64 *)xmalloc(sizeof(i64) * (size_t)(n + 1));
    Q->rowval = (i64 *)xmalloc(sizeof(i64) * (size_t)cap);
    Q->nzval = (double *)xmalloc(sizeof(double) * (size_t)cap);
    i64 *mark = (i64 *)xmalloc(sizeof(i64) * (size_t)n);
    double *acc = (double *)xmalloc(sizeof(double) * (size_t)n);
    for (i64 i = 0; i < n; i++) mark[i] = -1;
    i64 nz = 0;
    for (i64 j = 0; j < n; j++) {
        Q->colptr[j] = nz;
        i64 worst = 0;
        for (i64 p = Jp[j]; p < Jp[j + 1]; p++) worst += Tp[Ji[p] + 1] - Tp[Ji[p]];
        worst += Hp[j + 1] - Hp[j];
        if (nz + worst > cap) {
            cap = (nz + worst) * 2;
            Q->rowval = (i64 *)realloc(Q->rowval, sizeof(i64) * (size_t)cap);
            Q->nzval = (double *)realloc(Q->nzval, sizeof(double) * (size_t)cap);
        }
        i64 start = nz;
        for (i64 p = Jp[j]; p < Jp[j + 1]; p++) {
            i64 k = Ji[p];
            double b = Jx[p];
            for (i64 q = Tp[k]; q < Tp[k + 1]; q++) {
                i64 i = Ti[q];
                double prod = Tx[q] * b;
                if (mark[i] != j) { mark[i] = j; acc[i] = prod; Q->rowval[nz++] = i; }
                else acc[i] = acc[i] + prod;
            }
        }
        /* + H (lower triangular, eval.jl:132-134): union pattern */
        for (i64 p = Hp[j]; p < Hp[j + 1]; p++) {
            i64 i = Hi[p];
            if (mark[i] != j) { mark[i] = j; acc[i] = Hx[p]; Q->rowval[nz++] = i; }
            else acc[i] = acc[i] + Hx[p];
        }
        qsort(Q->rowval + start, (size_t)(nz - start), sizeof(i64), cmp_i64);
        for (i64 p = start; p < nz; p++) Q->nzval[p] = acc[Q->rowval[p]];
    }
    Q->colptr[n] = nz;
    Q->nnz = nz;
    free(Tp); free(Ti); free(Tx); free(mark); free(acc);
    return Q;
}

/* ------------------------------------------------------------------ */
/* Sparse factorisation of Symmetric(Q,:L) with a given permutation.   */
/* Up-looking simplicial algorithm on C = P Q P' (upper part of C).    */
/* mode 0: Cholesky  L L'   (pivot <= 0 or NaN => fail)                */
/* mode 1: LDL'       unit L, D; no pivoting (zero pivot => fail)      */
/* ------------------------------------------------------------------ */
typedef struct {
    i64 n, lnz;
    int mode;
    i64 *perm;   /* perm[newidx] = oldidx */
    i64 *parent; /* etree */
    i64 *Lp, *Li;
    double *Lx;
    double *D;   /* mode 1 */
    /* permuted upper pattern (symbolic, reusable) */
    i64 *Cp, *Ci, *Cmap; /* Cmap: index into caller's nzval */
    double flops;
} orc_factor;

void orc_factor_free(orc_factor *F) {
    if (!F) return;
    free(F->perm); free(F->parent); free(F->Lp); free(F->Li); free(F->Lx); free(F->D);
    free(F->Cp); free(F->Ci); free(F->Cmap); free(F);
}
i64 orc_factor_lnz(const orc_factor *F) { return F->lnz; }
double orc_factor_flops(const orc_factor *F) { return F->flops; }
void orc_factor_diag(const orc_factor *F, double *d) {
    for (i64 k = 0; k < F->n; k++) d[k] = (F->mode == 1) ? F->D[k] : F->Lx[F->Lp[k]];
}

/* Symbolic analysis: permuted upper pattern, etree, column counts, Lp.
   Only entries with row >= col of the input are read (the reference relies on
   this: test/linear_system_solvers.jl:74-84). perm may be NULL (identity). */
orc_factor *orc_analyze(i64 n, const i64 *Ap, const i64 *Ai, const i64 *perm)
{
    orc_factor *F = (orc_factor *)calloc(1, sizeof(orc_factor));
    F->n = n;
    F->perm = (i64 *)xmalloc(sizeof(i64) * (size_t)n);
    i64 *iperm = (i64 *)xmalloc(sizeof(i64) * (size_t)n);
    for (i64 k = 0; k < n; k++) F->perm[k] = perm ? perm[k] : k;
    for (i64 k = 0; k < n; k++) iperm[F->perm[k]] = k;
    /* count entries of upper(C) per column */
    i64 *Cp = (i64 *)calloc((size_t)n + 2, sizeof(i64));
    i64 cnz = 0;
    for (i64 j = 0; j < n; j++)
        for (i64 p = Ap[j]; p < Ap[j + 1]; p++) {
            i64 i = Ai[p];
            if (i < j) continue; /* upper triangle of the input is never read */
            i64 a = iperm[i], b = iperm[j];
            i64 col = a > b ? a : b;
            Cp[col + 2]++; cnz++;
        }
    for (i64 k = 0; k < n; k++) Cp[k + 2] += Cp[k + 1];
    i64 *Ci = (i64 *)xmalloc(sizeof(i64) * (size_t)cnz);
    i64 *Cmap = (i64 *)xmalloc(sizeof(i64) * (size_t)cnz);
    for (i64 j = 0; j < n; j++)
        for (i64 p = Ap[j]; p < Ap[j + 1]; p++) {
            i64 i = Ai[p];
            if (i < j) continue;
            i64 a = iperm[i], b = iperm[j];
            i64 col = a > b ? a : b, row = a > b ? b : a;
            i64 q = Cp[col + 1]++;
            Ci[q] = row; Cmap[q] = p;
        }
    F->Cp = Cp; F->Ci = Ci; F->Cmap = Cmap;
    /* elimination tree (Liu) with path compression */
    i64 *parent = (i64 *)xmalloc(sizeof(i64) * (size_t)n);
    i64 *anc = (i64 *)xmalloc(sizeof(i64) * (size_t)n);
    for (i64 k = 0; k < n; k++) {
        parent[k] = -1; anc[k] = -1;
        for (i64 p = Cp[k]; p < Cp[k + 1]; p++) {
            i64 i = Ci[p];
            while (i != -1 && i < k) {
                i64 nxt = anc[i];
                anc[i] = k;
                if (nxt == -1) parent[i] = k;
                i = nxt;
            }
        }
    }
    F->parent = parent;
    /* column counts by walking row subtrees (O(nnz(L)), fine for an oracle) */
    i64 *cnt = (i64 *)calloc((size_t)n, sizeof(i64));
    i64 *flag = anc; /* reuse */
    for (i64 k = 0; k < n; k++) flag[k] = -1;
    for (i64 k = 0; k < n; k++) {
        flag[k] = k;
        for (i64 p = Cp[k]; p < Cp[k + 1]; p++) {
            i64 i = Ci[p];
            while (i != -1 && i < k && flag[i] != k) { cnt[i]++; flag[i] = k; i = parent[i]; }
        }
    }
    F->Lp = (i64 *)xmalloc(sizeof(i64) * (size_t)(n + 1));
    F->Lp[0] = 0;
    double fl = 0.0;
    for (i64 k = 0; k < n; k++) {
        F->Lp[k + 1] = F->Lp[k] + cnt[k] + 1;
        fl += (double)(cnt[k] + 1) * (double)(cnt[k] + 1);
    }
    F->flops = fl;
    F->lnz = F->Lp[n];
    F->Li = (i64 *)xmalloc(sizeof(i64) * (size_t)F->lnz);
    F->Lx = (double *)xmalloc(sizeof(double) * (size_t)F->lnz);
    F->D = (double *)xmalloc(sizeof(double) * (size_t)n);
    free(cnt); free(anc); free(iperm);
    return F;
}

/* Numeric factorisation.  Ax = caller's nzval array (same pattern as analyze).
   dshift (may be NULL): value used INSTEAD of the stored diagonal is
   diag_override[i] (schur.jl:75-77 writes schur_diag[i]+delta into Q[i,i]);
   when diag_override is NULL the stored values are used.
   Returns 1 if the factorisation completed, 0 if a pivot failed. */
int orc_factorize(orc_factor *F, const double *Ax, const double *diag_override, int mode)
{
    i64 n = F->n;
    const i64 *Cp = F->Cp, *Ci = F->Ci, *Cmap = F->Cmap, *parent = F->parent;
    i64 *Lp = F->Lp, *Li = F->Li;
    double *Lx = F->Lx;
    F->mode = mode;
    double *x = (double *)calloc((size_t)n, sizeof(double));
    i64 *stack = (i64 *)xmalloc(sizeof(i64) * (size_t)n);
    i64 *flag = (i64 *)xmalloc(sizeof(i64) * (size_t)n);
    i64 *fill = (i64 *)xmalloc(sizeof(i64) * (size_t)n); /* next free slot per column */
    for (i64 k = 0; k < n; k++) { flag[k] = -1; fill[k] = Lp[k] + 1; }
    int ok = 1;
    for (i64 k = 0; k < n && ok; k++) {
        /* pattern of row k of L = reach of C[:,k] in the etree, topological order */
        i64 top = n;
        flag[k] = k;
        double d = 0.0;
        int have_d = 0;
        for (i64 p = Cp[k]; p < Cp[k + 1]; p++) {
            i64 i = Ci[p];
            double v = Ax[Cmap[p]];
            if (i == k) { d = have_d ? d + v : v; have_d = 1; continue; }
            x[i] += v; /* duplicates cannot occur for a valid CSC; += keeps it total */
            i64 len = 0;
            while (flag[i] != k) { stack[len++] = i; flag[i] = k; i = parent[i]; }
            while (len > 0) stack[--top] = stack[--len];
        }
        if (diag_override) d = diag_override[F->perm[k]];
        for (i64 t = top; t < n; t++) {
            i64 i = stack[t];
            double xi = x[i];
            x[i] = 0.0;
            double lki;
            if (mode == 0) {
                lki = xi / Lx[Lp[i]];
                for (i64 p = Lp[i] + 1; p < fill[i]; p++) x[Li[p]] -= Lx[p] * lki;
                d -= lki * lki;
            } else {
                double di = F->D[i];
                lki = xi / di;
                for (i64 p = Lp[i] + 1; p < fill[i]; p++) x[Li[p]] -= Lx[p] * xi;
                d -= lki * xi;
            }
            i64 q = fill[i]++;
            Li[q] = k; Lx[q] = lki;
        }
        Li[Lp[k]] = k;
        if (mode == 0) {
            if (!(d > 0.0)) { ok = 0; Lx[Lp[k]] = d; break; } /* d <= 0 or NaN */
            Lx[Lp[k]] = sqrt(d);
        } else {
            F->D[k] = d;
            Lx[Lp[k]] = 1.0;
            if (d == 0.0 || d != d) { ok = 0; break; } /* ZeroPivotException / NaN */
        }
    }
    free(x); free(stack); free(flag); free(fill);
    return ok;
}

/* inertia_status (linear_system_solvers.jl:48-91) applied to LDL' pivots with
   the +-1e-20 classification of julia.jl:72-80.  Returns 1 iff pos==n_ && neg==m_. */
int orc_ldlt_inertia_ok(const orc_factor *F, i64 n_, i64 m_)
{
    const double tol = 1e-20;
    i64 pos = 0, neg = 0, zer = 0;
    for (i64 k = 0; k < F->n; k++) {
        double d = F->D[k];
        if (d != d || isinf(d)) return 0;
        if (d > tol) pos++; else if (d < -tol) neg++; else zer++;
    }
    (void)zer;
    return (pos == n_ && neg == m_) ? 1 : 0;
}

/* x = F \ b   (julia.jl:99-113): permute, L, (D), L', un-permute */
void orc_solve(const orc_factor *F, const double *b, double *xout)
{
    i64 n = F->n;
    const i64 *Lp = F->Lp, *Li = F->Li;
    const double *Lx = F->Lx;
    double *w = (double *)xmalloc(sizeof(double) * (size_t)n);
    for (i64 k = 0; k < n; k++) w[k] = b[F->perm[k]];
    if (F->mode == 0) {
        for (i64 j = 0; j < n; j++) {
            w[j] /= Lx[Lp[j]];
            double wj = w[j];
            for (i64 p = Lp[j] + 1; p < Lp[j + 1]; p++) w[Li[p]] -= Lx[p] * wj;
        }
        for (i64 j = n - 1; j >= 0; j--) {
            double acc = w[j];
            for (i64 p = Lp[j] + 1; p < Lp[j + 1]; p++) acc -= Lx[p] * w[Li[p]];
            w[j] = acc / Lx[Lp[j]];
        }
    } else {
        for (i64 j = 0; j < n; j++) {
            double wj = w[j];
            for (i64 p = Lp[j] + 1; p < Lp[j + 1]; p++) w[Li[p]] -= Lx[p] * wj;
        }
        for (i64 j = 0; j < n; j++) w[j] /= F->D[j];
        for (i64 j = n - 1; j >= 0; j--) {
            double acc = w[j];
            for (i64 p = Lp[j] + 1; p < Lp[j + 1]; p++) acc -= Lx[p] * w[Li[p]];
            w[j] = acc;
        }
    }
    for (i64 k = 0; k < n; k++) xout[F->perm[k]] = w[k];
    free(w);
}

/* ------------------------------------------------------------------ */
/* delta rule (delta_strategy.jl:37-114)                               */
/* Q lower pattern (Ap, Ai, Ax) with schur_diag the unshifted diagonal.*/
/* deltas_out (len >= max_record) receives every delta tried.          */
/* returns status: 1 = :success, 0 = :failure (delta > delta_max),     */
/*                -1 = "max it" error                                  */
/* ------------------------------------------------------------------ */
int orc_delta_loop(orc_factor *F, const double *Ax, const double *schur_diag,
                   double delta_prev, double delta_zero, double delta_min,
                   double delta_max, double delta_start, double inc, double dec,
                   double *delta_out, i64 *num_fac_out,
                   double *deltas_out, i64 max_record)
{
    i64 n = F->n;
    double dmin = INFINITY;
    for (i64 i = 0; i < n; i++) if (schur_diag[i] < dmin || schur_diag[i] != schur_diag[i]) dmin = schur_diag[i];
    double tau = 1.5 * dmin; /* diag_min, kkt_system_solver.jl:291-294 */
    double *dg = (double *)xmalloc(sizeof(double) * (size_t)n);
    i64 num_fac = 0;
    double delta = delta_zero;
    int status = -1;
    if (tau > 0.0) {
        tau = 0.0;
        for (i64 i = 0; i < n; i++) dg[i] = schur_diag[i] + delta;
        int ok = orc_factorize(F, Ax, dg, 0);
        if (num_fac < max_record) deltas_out[num_fac] = delta;
        num_fac++;
        if (ok == 1) { status = 1; goto done; }
    }
    for (int it = 1; it <= 500; it++) {
        if (it == 1) {
            if (delta_prev != 0.0) {
                double a = delta_min - tau, b = delta_prev * dec;
                delta = (a != a || b != b) ? a + b : (a > b ? a : b);   /* Julia's max propagates NaN */
            } else {
                delta = delta_start - tau;
            }
        } else {
            delta = delta * inc;
        }
        for (i64 i = 0; i < n; i++) dg[i] = schur_diag[i] + delta;
        int ok = orc_factorize(F, Ax, dg, 0);
        if (num_fac < max_record) deltas_out[num_fac] = delta;
        num_fac++;
        if (ok == 1) { status = 1; goto done; }
        if (delta > delta_max) { status = 0; goto done; }
    }
done:
    free(dg);
    *delta_out = delta;
    *num_fac_out = num_fac;
    return status;
}

/* ------------------------------------------------------------------ */
/* Products on the cached matrices (eval.jl:102-108, 221-234)          */
/* ------------------------------------------------------------------ */
static void jac_prod(i64 n, i64 m, const i64 *Jp, const i64 *Ji, const double *Jx,
                     const double *x, double *out)
{
    for (i64 k = 0; k < m; k++) out[k] = 0.0;
    for (i64 j = 0; j < n; j++) {
        double xj = x[j];
        for (i64 p = Jp[j]; p < Jp[j + 1]; p++) out[Ji[p]] += Jx[p] * xj;
    }
}
static void jac_T_prod(i64 n, const i64 *Jp, const i64 *Ji, const double *Jx,
                       const double *yv, double *out)
{
    for (i64 j = 0; j < n; j++) {
        double acc = 0.0;
        for (i64 p = Jp[j]; p < Jp[j + 1]; p++) acc += Jx[p] * yv[Ji[p]];
        out[j] = acc;
    }
}
/* L v + L' v - diag(L) .* v */
static void hess_prod(i64 n, const i64 *Hp, const i64 *Hi, const double *Hx,
                      const double *v, double *out, double *tmp)
{
    for (i64 i = 0; i < n; i++) { out[i] = 0.0; tmp[i] = 0.0; }
    for (i64 j = 0; j < n; j++) {
        double vj = v[j], acc = 0.0, dj = 0.0;
        for (i64 p = Hp[j]; p < Hp[j + 1]; p++) {
            i64 i = Hi[p];
            out[i] += Hx[p] * vj;       /* L v   */
            acc += Hx[p] * v[i];        /* L' v  */
            if (i == j) dj += Hx[p];
        }
        tmp[j] = acc - dj * vj;
    }
    for (i64 i = 0; i < n; i++) out[i] = out[i] + tmp[i];
}

static double nrm_inf3(const double *a, i64 na, const double *b, i64 nb, const double *c, i64 nc)
{
    double r = 0.0; int nan = 0;
    for (i64 i = 0; i < na; i++) { double v = fabs(a[i]); if (v != v) nan = 1; if (v > r) r = v; }
    for (i64 i = 0; i < nb; i++) { double v = fabs(b[i]); if (v != v) nan = 1; if (v > r) r = v; }
    for (i64 i = 0; i < nc; i++) { double v = fabs(c[i]); if (v != v) nan = 1; if (v > r) r = v; }
    return nan ? NAN : r;
}

/* compute_direction_implementation! + solver_schur_rhs + update_kkt_error!
   (schur.jl:89-182, kkt_system_solver.jl:27-96).  kkt_err = [error_D, error_P,
   error_mu, overall, rhs_norm, ratio] in the inf norm. */
void orc_direction(const orc_factor *F, i64 n, i64 m,
                   const i64 *Jp, const i64 *Ji, const double *Jx,
                   const i64 *Hp, const i64 *Hi, const double *Hx,
                   const double *y, const double *s, double delta,
                   const double *dual_r, const double *primal_r, const double *comp_r,
                   int n_refine,
                   double *dx, double *dy, double *ds, double *kkt_err)
{
    double *S = (double *)xmalloc(sizeof(double) * (size_t)m);
    double *sym_p = (double *)xmalloc(sizeof(double) * (size_t)m);
    double *ym = (double *)xmalloc(sizeof(double) * (size_t)m);
    double *tm = (double *)xmalloc(sizeof(double) * (size_t)m);
    double *b = (double *)xmalloc(sizeof(double) * (size_t)n);
    double *res = (double *)xmalloc(sizeof(double) * (size_t)n);
    double *sol = (double *)xmalloc(sizeof(double) * (size_t)n);
    double *jr = (double *)xmalloc(sizeof(double) * (size_t)n);
    double *hr = (double *)xmalloc(sizeof(double) * (size_t)n);
    double *tn = (double *)xmalloc(sizeof(double) * (size_t)n);
    for (i64 k = 0; k < m; k++) {
        sym_p[k] = primal_r[k] + comp_r[k] / y[k];
        S[k] = y[k] / s[k];
        ym[k] = primal_r[k] * S[k] + comp_r[k] / s[k];
    }
    jac_T_prod(n, Jp, Ji, Jx, ym, b);
    for (i64 i = 0; i < n; i++) { b[i] = dual_r[i] + b[i]; res[i] = b[i]; dx[i] = 0.0; }
    for (int it = 0; it < n_refine; it++) {
        orc_solve(F, res, sol);
        for (i64 i = 0; i < n; i++) dx[i] += sol[i];
        jac_prod(n, m, Jp, Ji, Jx, dx, tm);
        for (i64 k = 0; k < m; k++) tm[k] = S[k] * tm[k];
        jac_T_prod(n, Jp, Ji, Jx, tm, jr);
        hess_prod(n, Hp, Hi, Hx, dx, hr, tn);
        for (i64 i = 0; i < n; i++) {
            double hess_res = hr[i] + delta * dx[i];
            res[i] = b[i] - (jr[i] + hess_res);
        }
    }
    jac_prod(n, m, Jp, Ji, Jx, dx, tm);
    for (i64 k = 0; k < m; k++) {
        dy[k] = -(tm[k] - sym_p[k]) * S[k];
        ds[k] = tm[k] - primal_r[k];
    }
    /* update_kkt_error! */
    jac_T_prod(n, Jp, Ji, Jx, dy, jr);     /* J_err */
    hess_prod(n, Hp, Hi, Hx, dx, hr, tn);  /* H_err */
    double *eD = res, *eP = tm, *eM = ym;
    for (i64 i = 0; i < n; i++) {
        double delta_err = delta * dx[i] + 0.0; /* delta_s_vec is all zero */
        eD[i] = (delta_err + hr[i] - jr[i]) - dual_r[i];
    }
    jac_prod(n, m, Jp, Ji, Jx, dx, sym_p);
    for (i64 k = 0; k < m; k++) {
        eP[k] = sym_p[k] - ds[k] - primal_r[k];
        eM[k] = s[k] * dy[k] + y[k] * ds[k] - comp_r[k];
    }
    kkt_err[0] = nrm_inf3(eD, n, NULL, 0, NULL, 0);
    kkt_err[1] = nrm_inf3(eP, m, NULL, 0, NULL, 0);
    kkt_err[2] = nrm_inf3(eM, m, NULL, 0, NULL, 0);
    kkt_err[3] = nrm_inf3(eD, n, eP, m, eM, m);
    kkt_err[4] = nrm_inf3(dual_r, n, primal_r, m, comp_r, m);
    kkt_err[5] = kkt_err[3] / kkt_err[4];
    free(S); free(sym_p); free(ym); free(tm); free(b); free(res); free(sol);
    free(jr); free(hr); free(tn);
}

/* ---------------------------------------------------------------------------
 * Helper of oracle/supernodal.py (the multifrontal CPU baseline): extend-add of a child's
 * update block (rc x rc, column-major, lower part) into its parent's front, split into the
 * pivot-column panel P (N x c, leading dimension ldp) and the parent's own update block U
 * (r x r, leading dimension r).  rel[t] = position of child row t inside the parent front.
 * Plain index arithmetic: the dense algebra of that baseline runs in BLAS. */
void orc_extend_add(double *P, i64 ldp, i64 c, double *U, i64 r, const i64 *rel, i64 rc, const double *cb) {
    for (i64 u = 0; u < rc; u++) {
        const i64 pj = rel[u];
        const double *col = cb + u * rc;
        if (pj < c) {
            double *dst = P + pj * ldp;
            for (i64 t = u; t < rc; t++) dst[rel[t]] += col[t];
        } else {
            double *dst = U + (pj - c) * r - c;
            for (i64 t = u; t < rc; t++) dst[rel[t]] += col[t];
        }
    }
}

/* Supernodal triangular solves x <- (L L')^-1 x (x in the permuted order) for oracle/supernodal.py.
 * Panels: supernode s has c = sfirst[s+1]-sfirst[s] pivot columns and r = rowptr[s+1]-rowptr[s]
 * rows below them (global indices rowidx), stored column-major at L + Loff[s] with leading
 * dimension ld = (c+r+1) & ~1.  u (length rowptr[ns]) is scratch for the update vectors;
 * children of s are child_list[child_ptr[s] .. child_ptr[s+1]) in ascending order, rel maps a
 * child's rows to positions of the parent front.  Column-oriented forward, row-oriented backward:
 * every entry of L is read once per sweep (CHOLMOD's supernodal solve does the same with BLAS-2). */
void orc_snode_solve(i64 ns, const i64 *sfirst, const i64 *rowptr, const i64 *rowidx, const i64 *rel,
                     const i64 *Loff, const i64 *child_ptr, const i64 *child_list, const double *L,
                     double *x, double *u) {
    for (i64 s = 0; s < ns; s++) {
        const i64 f = sfirst[s], c = sfirst[s + 1] - f, rp = rowptr[s], r = rowptr[s + 1] - rp;
        const i64 N = c + r, ld = (N + 1) & ~(i64)1;
        const double *P = L + Loff[s];
        double *xs = x + f, *us = u + rp;
        for (i64 t = 0; t < r; t++) us[t] = 0.0;
        for (i64 k = child_ptr[s]; k < child_ptr[s + 1]; k++) {
            const i64 ch = child_list[k], rpc = rowptr[ch], rc = rowptr[ch + 1] - rpc;
            for (i64 t = 0; t < rc; t++) {
                const i64 dst = rel[rpc + t];
                if (dst < c) xs[dst] += u[rpc + t]; else us[dst - c] += u[rpc + t];
            }
        }
        for (i64 j = 0; j < c; j++) {
            const double *col = P + j * ld;
            const double xj = xs[j] / col[j];
            xs[j] = xj;
            for (i64 i = j + 1; i < c; i++) xs[i] -= col[i] * xj;
            for (i64 i = c; i < N; i++) us[i - c] -= col[i] * xj;
        }
    }
    for (i64 s = ns - 1; s >= 0; s--) {
        const i64 f = sfirst[s], c = sfirst[s + 1] - f, rp = rowptr[s], r = rowptr[s + 1] - rp;
        const i64 N = c + r, ld = (N + 1) & ~(i64)1;
        const double *P = L + Loff[s];
        double *xs = x + f, *us = u + rp;
        for (i64 t = 0; t < r; t++) us[t] = x[rowidx[rp + t]];
        for (i64 j = c - 1; j >= 0; j--) {
            const double *col = P + j * ld;
            double acc = xs[j];
            for (i64 i = j + 1; i < c; i++) acc -= col[i] * xs[i];
            for (i64 i = c; i < N; i++) acc -= col[i] * us[i - c];
            xs[j] = acc / col[j];
        }
    }
}
