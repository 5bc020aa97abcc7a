/*
 * snode.c -- supernodal multifrontal CPU restatement of the reference's linear-algebra layer,
 * with its OWN fill-reducing orderings and its OWN symbolic analysis.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as kkt_oracle.c): loaded by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs; never by the product package.  Nothing in
 * this file comes from, or links against, the library under test.
 *
 * What it stands for.  The reference factorises with `cholesky(Symmetric(Q,:L))`
 * (src/linear_system_solvers/julia.jl:34), i.e. CHOLMOD: fill-reducing ordering, elimination
 * tree, column counts, relaxed supernodes, supernodal numeric factorisation on a threaded BLAS,
 * "pivot <= 0 => not positive definite" (julia.jl:39-41), and solves with `F \ b`
 * (julia.jl:99-113).  CHOLMOD is not vendored in /root/reference and is absent from this image
 * (Julia stdlib SuiteSparse 5.4-5.10, fixed by the Julia binary), so this file restates the
 * published algorithms: Liu's elimination tree, the Gilbert-Ng-Peyton skeleton column counts,
 * fundamental supernodes with relaxed amalgamation (CHOLMOD's documented nrelax = 4/16/48,
 * zrelax = 0.8/0.1/0.05 rule), a multifrontal numeric phase with dpotrf / dtrsm / dsyrk on the
 * fronts and dtrsv / dgemv supernodal solves.  The analysis is redone on every call by the
 * callers that mirror `linear_solver_recycle = false` (parameters.jl:38).
 * Parity status: "parity unpinned" (SURVEY.md 8c); pinned against kkt_oracle.c and dense LAPACK
 * in tests/test_oracle.py.
 *
 * Orderings: METIS_NodeND (the static library that ships with the CUDA toolkit; CHOLMOD's own
 * default tries AMD then METIS), an exact minimum-degree ordering for small graphs, natural.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

/* ---- BLAS / LAPACK entry points, handed over by the Python side (scipy's OpenBLAS) ---- */
typedef void (*dpotrf_t)(char*, int*, double*, int*, int*);
typedef void (*dtrsm_t)(char*, char*, char*, char*, int*, int*, double*, double*, int*, double*, int*);
typedef void (*dsyrk_t)(char*, char*, int*, int*, double*, double*, int*, double*, double*, int*);
typedef void (*dgemv_t)(char*, int*, int*, double*, double*, int*, double*, int*, double*, double*, int*);
typedef void (*dtrsv_t)(char*, char*, char*, int*, double*, int*, double*, int*);
static dpotrf_t f_dpotrf; static dtrsm_t f_dtrsm; static dsyrk_t f_dsyrk; static dgemv_t f_dgemv; static dtrsv_t f_dtrsv;

void orc_sn_set_blas(void* potrf, void* trsm, void* syrk, void* gemv, void* trsv) {
    f_dpotrf = (dpotrf_t)potrf; f_dtrsm = (dtrsm_t)trsm; f_dsyrk = (dsyrk_t)syrk;
    f_dgemv = (dgemv_t)gemv; f_dtrsv = (dtrsv_t)trsv;
}
int orc_sn_has_blas(void) { return f_dpotrf && f_dtrsm && f_dsyrk && f_dgemv && f_dtrsv; }

/* ======================================================================================
 * Orderings
 * ====================================================================================== */
/* symmetric adjacency (no self loops) of the pattern of a lower-triangular CSC matrix */
static void build_adjacency(i64 n, const i64* Ap, const i64* Ai, i64** xadj_out, i64** adj_out) {
    i64* deg = (i64*)calloc((size_t)n + 1, sizeof(i64));
    for (i64 j = 0; j < n; j++)
        for (i64 p = Ap[j]; p < Ap[j + 1]; p++) {
            const i64 i = Ai[p];
            if (i != j) { deg[i]++; deg[j]++; }
        }
    i64* xadj = (i64*)malloc(((size_t)n + 1) * sizeof(i64));
    xadj[0] = 0;
    for (i64 j = 0; j < n; j++) xadj[j + 1] = xadj[j] + deg[j];
    i64* adj = (i64*)malloc((size_t)(xadj[n] ? xadj[n] : 1) * sizeof(i64));
    for (i64 j = 0; j < n; j++) deg[j] = xadj[j];
    for (i64 j = 0; j < n; j++)
        for (i64 p = Ap[j]; p < Ap[j + 1]; p++) {
            const i64 i = Ai[p];
            if (i != j) { adj[deg[i]++] = j; adj[deg[j]++] = i; }
        }
    free(deg);
    *xadj_out = xadj; *adj_out = adj;
}

extern int METIS_NodeND(i64* nvtxs, i64* xadj, i64* adjncy, i64* vwgt, i64* options, i64* perm, i64* iperm);
extern int METIS_SetDefaultOptions(i64* options);

/* perm[new] = old.  Returns 1 on success. */
int orc_order_metis(i64 n, const i64* Ap, const i64* Ai, i64* perm) {
    if (n <= 0) return 1;
    if (n == 1) { perm[0] = 0; return 1; }
    i64 *xadj, *adj;
    build_adjacency(n, Ap, Ai, &xadj, &adj);
    i64 options[64];
    METIS_SetDefaultOptions(options);
    i64* iperm = (i64*)malloc((size_t)n * sizeof(i64));
    i64 nv = n;
    const int rc = METIS_NodeND(&nv, xadj, adj, NULL, options, perm, iperm);
    free(iperm); free(xadj); free(adj);
    return rc == 1;
}

/* Exact minimum-degree ordering on an explicit elimination graph held as bit sets
 * (ties: lowest index).  O(n^2/64) memory: for the small cases of the test-suite only. */
int orc_order_mindeg(i64 n, const i64* Ap, const i64* Ai, i64* perm) {
    if (n > 16384) return 0;
    const i64 W = (n + 63) / 64;
    uint64_t* G = (uint64_t*)calloc((size_t)(n * W ? n * W : 1), sizeof(uint64_t));
    i64* deg = (i64*)calloc((size_t)n + 1, sizeof(i64));
    char* gone = (char*)calloc((size_t)n + 1, 1);
    for (i64 j = 0; j < n; j++)
        for (i64 p = Ap[j]; p < Ap[j + 1]; p++) {
            const i64 i = Ai[p];
            if (i == j) continue;
            G[i * W + (j >> 6)] |= 1ull << (j & 63);
            G[j * W + (i >> 6)] |= 1ull << (i & 63);
        }
    for (i64 v = 0; v < n; v++) { i64 d = 0; for (i64 w = 0; w < W; w++) d += __builtin_popcountll(G[v * W + w]); deg[v] = d; }
    for (i64 k = 0; k < n; k++) {
        i64 best = -1;
        for (i64 v = 0; v < n; v++) if (!gone[v] && (best < 0 || deg[v] < deg[best])) best = v;
        perm[k] = best; gone[best] = 1;
        const uint64_t* gb = G + best * W;
        for (i64 w = 0; w < W; w++) {
            uint64_t bits = gb[w];
            while (bits) {
                const i64 v = w * 64 + __builtin_ctzll(bits);
                bits &= bits - 1;
                uint64_t* gv = G + v * W;
                i64 d = 0;
                for (i64 q = 0; q < W; q++) { gv[q] |= gb[q]; }
                gv[best >> 6] &= ~(1ull << (best & 63));
                gv[v >> 6] &= ~(1ull << (v & 63));
                for (i64 q = 0; q < W; q++) d += __builtin_popcountll(gv[q]);
                deg[v] = d;
            }
        }
    }
    free(G); free(deg); free(gone);
    return 1;
}

/* ======================================================================================
 * Symbolic analysis
 * ====================================================================================== */
typedef struct {
    i64 n, nnzA;
    i64 *perm, *iperm;          /* perm[new] = old (fill-reducing ordering composed with the postorder) */
    i64 nsuper;
    i64 *sfirst;                /* nsuper+1: column ranges (new numbering) */
    i64 *sparent;               /* supernodal elimination tree */
    i64 *child_ptr, *child_list;
    i64 *rowptr, *rowidx;       /* rows below the pivot block, ascending */
    i64 *Loff;                  /* panel of supernode s: (c + r) x c, column-major, ld = c + r */
    i64 *amap;                  /* input entry -> position in L (or -1 for an ignored upper entry) */
    i64 *dpos;                  /* new column -> position of its diagonal in L */
    i64 nnzL;                   /* doubles of panel storage */
    i64 nnzL_true;              /* sum of the column counts (no relaxation zeros) */
    double flops;               /* sum of squared column counts */
    i64 max_front;
    double* L;                  /* numeric factor (NULL until factorised) */
    i64* pos;                   /* workspace n: global row -> local front index */
} SN;

static void* xmalloc(size_t b) { void* p = malloc(b ? b : 1); if (!p) { fprintf(stderr, "snode.c: out of memory (%zu bytes)\n", b); abort(); } return p; }
static void* xcalloc(size_t n, size_t b) { void* p = calloc(n ? n : 1, b); if (!p) { fprintf(stderr, "snode.c: out of memory\n"); abort(); } return p; }

/* permuted lower pattern B (CSC, columns in new numbering) and its transpose (rows) */
static void permuted_lower(i64 n, const i64* Ap, const i64* Ai, const i64* iperm, i64** Bp_o, i64** Bi_o) {
    i64* Bp = (i64*)xcalloc((size_t)n + 2, sizeof(i64));
    for (i64 j = 0; j < n; j++)
        for (i64 p = Ap[j]; p < Ap[j + 1]; p++) {
            const i64 i = Ai[p];
            if (i < j) continue;
            const i64 a = iperm[i], b = iperm[j];
            Bp[(a < b ? a : b) + 1]++;
        }
    for (i64 j = 0; j < n; j++) Bp[j + 1] += Bp[j];
    i64* Bi = (i64*)xmalloc((size_t)Bp[n] * sizeof(i64));
    i64* nx = (i64*)xmalloc(((size_t)n + 1) * sizeof(i64));
    memcpy(nx, Bp, ((size_t)n + 1) * sizeof(i64));
    for (i64 j = 0; j < n; j++)
        for (i64 p = Ap[j]; p < Ap[j + 1]; p++) {
            const i64 i = Ai[p];
            if (i < j) continue;
            const i64 a = iperm[i], b = iperm[j];
            const i64 col = a < b ? a : b, row = a < b ? b : a;
            Bi[nx[col]++] = row;
        }
    free(nx);
    *Bp_o = Bp; *Bi_o = Bi;
}
static void transpose_pattern(i64 n, const i64* Bp, const i64* Bi, i64** Tp_o, i64** Ti_o) {
    i64* Tp = (i64*)xcalloc((size_t)n + 2, sizeof(i64));
    for (i64 p = 0; p < Bp[n]; p++) Tp[Bi[p] + 1]++;
    for (i64 j = 0; j < n; j++) Tp[j + 1] += Tp[j];
    i64* Ti = (i64*)xmalloc((size_t)Bp[n] * sizeof(i64));
    i64* nx = (i64*)xmalloc(((size_t)n + 1) * sizeof(i64));
    memcpy(nx, Tp, ((size_t)n + 1) * sizeof(i64));
    for (i64 j = 0; j < n; j++)
        for (i64 p = Bp[j]; p < Bp[j + 1]; p++) Ti[nx[Bi[p]]++] = j;
    free(nx);
    *Tp_o = Tp; *Ti_o = Ti;
}

/* Liu's algorithm: row k of the lower triangle lists the columns i < k it touches; every such i
 * is walked up to its current root (virtual ancestors with path compression) and hung under k. */
static void elimination_tree(i64 n, const i64* Tp, const i64* Ti, i64* parent) {
    i64* anc = (i64*)xmalloc((size_t)n * sizeof(i64));
    for (i64 k = 0; k < n; k++) {
        parent[k] = -1; anc[k] = -1;
        for (i64 p = Tp[k]; p < Tp[k + 1]; p++) {
            i64 r = Ti[p];
            while (r != -1 && r < k) {
                const i64 next = anc[r];
                anc[r] = k;
                if (next == -1) parent[r] = k;
                r = next;
            }
        }
    }
    free(anc);
}

/* depth-first postorder, children visited in ascending order */
static void tree_postorder(i64 n, const i64* parent, i64* post) {
    i64* head = (i64*)xmalloc((size_t)n * sizeof(i64));
    i64* next = (i64*)xmalloc((size_t)n * sizeof(i64));
    i64* stack = (i64*)xmalloc((size_t)n * sizeof(i64));
    for (i64 j = 0; j < n; j++) head[j] = -1;
    for (i64 j = n - 1; j >= 0; j--)
        if (parent[j] != -1) { next[j] = head[parent[j]]; head[parent[j]] = j; }
    i64 k = 0;
    for (i64 root = 0; root < n; root++) {
        if (parent[root] != -1) continue;
        i64 top = 0;
        stack[0] = root;
        while (top >= 0) {
            const i64 v = stack[top];
            const i64 c = head[v];
            if (c == -1) { post[k++] = v; top--; }
            else { head[v] = next[c]; stack[++top] = c; }
        }
    }
    free(head); free(next); free(stack);
}

/* Column counts of L for a matrix whose elimination tree is already postordered (parent[j] > j,
 * subtrees are index ranges).  Skeleton-matrix algorithm of Gilbert, Ng and Peyton: column j
 * is a leaf of the row subtree of i exactly when its first descendant lies beyond every first
 * descendant seen for row i so far; each new leaf adds one to its column and takes one away at
 * the least common ancestor with the previous leaf of the same row. */
static i64 lca_find(i64* anc, i64 v) {
    i64 r = v;
    while (anc[r] != r) r = anc[r];
    while (anc[v] != r) { const i64 t = anc[v]; anc[v] = r; v = t; }
    return r;
}
static void column_counts(i64 n, const i64* Bp, const i64* Bi, const i64* parent, i64* cc) {
    i64* first = (i64*)xmalloc((size_t)n * sizeof(i64));
    i64* maxfirst = (i64*)xmalloc((size_t)n * sizeof(i64));
    i64* prevleaf = (i64*)xmalloc((size_t)n * sizeof(i64));
    i64* anc = (i64*)xmalloc((size_t)n * sizeof(i64));
    i64* w = (i64*)xcalloc((size_t)n, sizeof(i64));
    for (i64 j = 0; j < n; j++) { first[j] = -1; maxfirst[j] = -1; prevleaf[j] = -1; anc[j] = j; }
    for (i64 j = 0; j < n; j++) {
        if (first[j] == -1) w[j] = 1;                 /* a leaf of the elimination tree */
        for (i64 a = j; a != -1 && first[a] == -1; a = parent[a]) first[a] = j;
    }
    for (i64 j = 0; j < n; j++) {
        if (parent[j] != -1) w[parent[j]]--;          /* the diagonal entry is counted once per column */
        for (i64 p = Bp[j]; p < Bp[j + 1]; p++) {
            const i64 i = Bi[p];
            if (i <= j || first[j] <= maxfirst[i]) continue;
            maxfirst[i] = first[j];
            const i64 jp = prevleaf[i];
            prevleaf[i] = j;
            w[j]++;
            if (jp != -1) w[lca_find(anc, jp)]--;
        }
        if (parent[j] != -1) anc[j] = parent[j];
    }
    for (i64 j = 0; j < n; j++) cc[j] = w[j];
    for (i64 j = 0; j < n; j++) if (parent[j] != -1) cc[parent[j]] += cc[j];
    free(first); free(maxfirst); free(prevleaf); free(anc); free(w);
}

static int cmp_i64(const void* a, const void* b) {
    const i64 x = *(const i64*)a, y = *(const i64*)b;
    return (x > y) - (x < y);
}

void orc_sn_free(void* h);

/* perm_in: perm[new] = old, or NULL for the natural order.  relax != 0 enables amalgamation. */
void* orc_sn_analyze(i64 n, const i64* Ap, const i64* Ai, const i64* perm_in, int relax) {
    SN* S = (SN*)xcalloc(1, sizeof(SN));
    S->n = n; S->nnzA = Ap[n];
    i64* perm0 = (i64*)xmalloc((size_t)n * sizeof(i64));
    i64* iperm = (i64*)xmalloc((size_t)n * sizeof(i64));
    for (i64 k = 0; k < n; k++) perm0[k] = perm_in ? perm_in[k] : k;
    for (i64 k = 0; k < n; k++) iperm[perm0[k]] = k;
    i64 *Bp, *Bi, *Tp, *Ti;
    i64* parent = (i64*)xmalloc((size_t)n * sizeof(i64));
    i64* post = (i64*)xmalloc((size_t)n * sizeof(i64));
    permuted_lower(n, Ap, Ai, iperm, &Bp, &Bi);
    transpose_pattern(n, Bp, Bi, &Tp, &Ti);
    elimination_tree(n, Tp, Ti, parent);
    tree_postorder(n, parent, post);
    free(Bp); free(Bi); free(Tp); free(Ti);
    /* compose with the postorder and redo the tree in the final numbering */
    S->perm = (i64*)xmalloc((size_t)n * sizeof(i64));
    for (i64 k = 0; k < n; k++) S->perm[k] = perm0[post[k]];
    for (i64 k = 0; k < n; k++) iperm[S->perm[k]] = k;
    S->iperm = iperm;
    free(perm0); free(post);
    permuted_lower(n, Ap, Ai, iperm, &Bp, &Bi);
    transpose_pattern(n, Bp, Bi, &Tp, &Ti);
    elimination_tree(n, Tp, Ti, parent);
    free(Tp); free(Ti);
    i64* cc = (i64*)xmalloc((size_t)n * sizeof(i64));
    column_counts(n, Bp, Bi, parent, cc);
    S->flops = 0.0; S->nnzL_true = 0;
    for (i64 j = 0; j < n; j++) { S->flops += (double)cc[j] * (double)cc[j]; S->nnzL_true += cc[j]; }

    /* ---- fundamental supernodes: j+1 joins j when it is j's parent, j is its only child and
     *      the structures nest exactly */
    i64* nchild = (i64*)xcalloc((size_t)n, sizeof(i64));
    for (i64 j = 0; j < n; j++) if (parent[j] != -1) nchild[parent[j]]++;
    i64* fs_first = (i64*)xmalloc(((size_t)n + 1) * sizeof(i64));
    i64 nf = 0;
    for (i64 j = 0; j < n; j++) {
        const int joins = j > 0 && parent[j - 1] == j && nchild[j] == 1 && cc[j - 1] == cc[j] + 1;
        if (!joins) fs_first[nf++] = j;
    }
    fs_first[nf] = n;
    free(nchild);
    /* ---- relaxed amalgamation: a supernode absorbs its LAST child (the one whose columns end
     *      right before its own) when the explicit zeros this creates stay below the threshold
     *      of the merged width (<= 4 columns: always; <= 16: 80 %; <= 48: 10 %; else 5 %) */
    i64* first = (i64*)xmalloc((size_t)nf * sizeof(i64));
    i64* ncol = (i64*)xmalloc((size_t)nf * sizeof(i64));
    i64* nrow = (i64*)xmalloc((size_t)nf * sizeof(i64));     /* rows below the pivot block */
    double* zeros = (double*)xcalloc((size_t)nf, sizeof(double));
    char* alive = (char*)xmalloc((size_t)nf);
    i64* ends_at = (i64*)xmalloc((size_t)n * sizeof(i64));   /* column -> alive supernode ending there, or -1 */
    for (i64 j = 0; j < n; j++) ends_at[j] = -1;
    for (i64 s = 0; s < nf; s++) {
        first[s] = fs_first[s]; ncol[s] = fs_first[s + 1] - fs_first[s];
        nrow[s] = cc[first[s]] - ncol[s];
        alive[s] = 1;
        ends_at[first[s] + ncol[s] - 1] = s;
    }
    if (relax) {
        for (i64 p = 0; p < nf; p++) {
            for (;;) {
                if (first[p] == 0) break;
                const i64 s = ends_at[first[p] - 1];
                if (s < 0) break;
                const i64 slast = first[s] + ncol[s] - 1;
                if (parent[slast] < first[p] || parent[slast] >= first[p] + ncol[p]) break;   /* not a child of p */
                if (parent[slast] != first[p]) break;     /* keeps struct(s) inside cols(p) + struct(p) trivially */
                const double ns = (double)ncol[s], np_ = (double)ncol[p];
                const double extra = ns * ((np_ + (double)nrow[p]) - (double)nrow[s]);
                const double z = zeros[s] + zeros[p] + extra;
                const double tot = ns + np_;
                const double lnz = tot * (tot + 1.0) / 2.0 + tot * (double)nrow[p];
                int merge;
                if (tot <= 4.0) merge = 1;
                else if (tot <= 16.0) merge = z / lnz < 0.8;
                else if (tot <= 48.0) merge = z / lnz < 0.1;
                else merge = z / lnz < 0.05;
                if (!merge) break;
                ends_at[slast] = -1;
                alive[s] = 0;
                first[p] = first[s]; ncol[p] += ncol[s]; zeros[p] = z;
            }
        }
    }
    i64 ns = 0;
    for (i64 s = 0; s < nf; s++) ns += alive[s];
    S->nsuper = ns;
    S->sfirst = (i64*)xmalloc(((size_t)ns + 1) * sizeof(i64));
    {
        i64 k = 0;
        for (i64 s = 0; s < nf; s++) if (alive[s]) S->sfirst[k++] = first[s];
        S->sfirst[ns] = n;
    }
    free(first); free(ncol); free(nrow); free(zeros); free(alive); free(ends_at); free(fs_first);
    i64* col2sn = (i64*)xmalloc((size_t)n * sizeof(i64));
    for (i64 s = 0; s < ns; s++) for (i64 j = S->sfirst[s]; j < S->sfirst[s + 1]; j++) col2sn[j] = s;
    S->sparent = (i64*)xmalloc((size_t)ns * sizeof(i64));
    S->child_ptr = (i64*)xcalloc((size_t)ns + 2, sizeof(i64));
    for (i64 s = 0; s < ns; s++) {
        const i64 pj = parent[S->sfirst[s + 1] - 1];
        S->sparent[s] = pj == -1 ? -1 : col2sn[pj];
        if (pj != -1) S->child_ptr[col2sn[pj] + 1]++;
    }
    for (i64 s = 0; s < ns; s++) S->child_ptr[s + 1] += S->child_ptr[s];
    S->child_list = (i64*)xmalloc((size_t)(S->child_ptr[ns] ? S->child_ptr[ns] : 1) * sizeof(i64));
    {
        i64* nx = (i64*)xmalloc(((size_t)ns + 1) * sizeof(i64));
        memcpy(nx, S->child_ptr, ((size_t)ns + 1) * sizeof(i64));
        for (i64 s = 0; s < ns; s++) if (S->sparent[s] != -1) S->child_list[nx[S->sparent[s]]++] = s;
        free(nx);
    }
    /* ---- row structures: own entries below the pivot block + the children's structures */
    S->rowptr = (i64*)xmalloc(((size_t)ns + 1) * sizeof(i64));
    i64 cap = S->nnzA + n + 16, used = 0;
    S->rowidx = (i64*)xmalloc((size_t)cap * sizeof(i64));
    i64* mark = (i64*)xmalloc((size_t)n * sizeof(i64));
    for (i64 j = 0; j < n; j++) mark[j] = -1;
    S->max_front = 0;
    for (i64 s = 0; s < ns; s++) {
        const i64 last = S->sfirst[s + 1] - 1;
        S->rowptr[s] = used;
        i64 need = 0;
        for (i64 j = S->sfirst[s]; j <= last; j++) need += Bp[j + 1] - Bp[j];
        for (i64 q = S->child_ptr[s]; q < S->child_ptr[s + 1]; q++) { const i64 c = S->child_list[q]; need += S->rowptr[c + 1] - S->rowptr[c]; }
        if (used + need > cap) { cap = (used + need) * 2; S->rowidx = (i64*)realloc(S->rowidx, (size_t)cap * sizeof(i64)); if (!S->rowidx) abort(); }
        for (i64 j = S->sfirst[s]; j <= last; j++)
            for (i64 p = Bp[j]; p < Bp[j + 1]; p++) {
                const i64 i = Bi[p];
                if (i > last && mark[i] != s) { mark[i] = s; S->rowidx[used++] = i; }
            }
        for (i64 q = S->child_ptr[s]; q < S->child_ptr[s + 1]; q++) {
            const i64 c = S->child_list[q];
            for (i64 p = S->rowptr[c]; p < S->rowptr[c + 1]; p++) {
                const i64 i = S->rowidx[p];
                if (i > last && mark[i] != s) { mark[i] = s; S->rowidx[used++] = i; }
            }
        }
        qsort(S->rowidx + S->rowptr[s], (size_t)(used - S->rowptr[s]), sizeof(i64), cmp_i64);
        const i64 N = (last + 1 - S->sfirst[s]) + (used - S->rowptr[s]);
        if (N > S->max_front) S->max_front = N;
    }
    S->rowptr[ns] = used;
    free(mark);
    /* ---- storage offsets, scatter map of the input entries, diagonal positions */
    S->Loff = (i64*)xmalloc(((size_t)ns + 1) * sizeof(i64));
    S->Loff[0] = 0;
    for (i64 s = 0; s < ns; s++) {
        const i64 c = S->sfirst[s + 1] - S->sfirst[s], r = S->rowptr[s + 1] - S->rowptr[s];
        S->Loff[s + 1] = S->Loff[s] + (c + r) * c;
    }
    S->nnzL = S->Loff[ns];
    S->dpos = (i64*)xmalloc((size_t)n * sizeof(i64));
    S->pos = (i64*)xmalloc((size_t)n * sizeof(i64));
    S->amap = (i64*)xmalloc((size_t)(S->nnzA ? S->nnzA : 1) * sizeof(i64));
    for (i64 s = 0; s < ns; s++) {
        const i64 f = S->sfirst[s], c = S->sfirst[s + 1] - f, r = S->rowptr[s + 1] - S->rowptr[s];
        for (i64 j = 0; j < c; j++) S->dpos[f + j] = S->Loff[s] + j + j * (c + r);
    }
    {
        /* position of global row i inside the front of the supernode that owns column j */
        i64 cur = -1;
        /* entries are visited per owning supernode so that `pos` is valid: bucket them first */
        i64* cnt = (i64*)xcalloc((size_t)ns + 2, sizeof(i64));
        for (i64 j = 0; j < n; j++)
            for (i64 p = Ap[j]; p < Ap[j + 1]; p++) {
                const i64 i = Ai[p];
                if (i < j) continue;
                const i64 a = iperm[i], b = iperm[j];
                cnt[col2sn[a < b ? a : b] + 1]++;
            }
        for (i64 s = 0; s < ns; s++) cnt[s + 1] += cnt[s];
        i64* ent = (i64*)xmalloc((size_t)(cnt[ns] ? cnt[ns] : 1) * sizeof(i64));
        i64* nx = (i64*)xmalloc(((size_t)ns + 1) * sizeof(i64));
        memcpy(nx, cnt, ((size_t)ns + 1) * sizeof(i64));
        for (i64 j = 0; j < n; j++)
            for (i64 p = Ap[j]; p < Ap[j + 1]; p++) {
                const i64 i = Ai[p];
                if (i < j) { S->amap[p] = -1; continue; }
                const i64 a = iperm[i], b = iperm[j];
                ent[nx[col2sn[a < b ? a : b]]++] = p;
            }
        /* input column of entry p: recover by a second pass (colof) */
        i64* colof = (i64*)xmalloc((size_t)(S->nnzA ? S->nnzA : 1) * sizeof(i64));
        for (i64 j = 0; j < n; j++) for (i64 p = Ap[j]; p < Ap[j + 1]; p++) colof[p] = j;
        for (i64 s = 0; s < ns; s++) {
            const i64 f = S->sfirst[s], c = S->sfirst[s + 1] - f, r = S->rowptr[s + 1] - S->rowptr[s];
            for (i64 j = 0; j < c; j++) S->pos[f + j] = j;
            for (i64 t = 0; t < r; t++) S->pos[S->rowidx[S->rowptr[s] + t]] = c + t;
            for (i64 q = cnt[s]; q < cnt[s + 1]; q++) {
                const i64 p = ent[q];
                const i64 a = iperm[Ai[p]], b = iperm[colof[p]];
                const i64 col = a < b ? a : b, row = a < b ? b : a;
                S->amap[p] = S->Loff[s] + S->pos[row] + (col - f) * (c + r);
            }
        }
        (void)cur;
        free(cnt); free(ent); free(nx); free(colof);
    }
    free(Bp); free(Bi); free(parent); free(cc); free(col2sn);
    return S;
}

void orc_sn_free(void* h) {
    SN* S = (SN*)h;
    if (!S) return;
    free(S->perm); free(S->iperm); free(S->sfirst); free(S->sparent); free(S->child_ptr); free(S->child_list);
    free(S->rowptr); free(S->rowidx); free(S->Loff); free(S->amap); free(S->dpos); free(S->L); free(S->pos);
    free(S);
}

double orc_sn_info(void* h, int what) {
    SN* S = (SN*)h;
    switch (what) {
        case 0: return (double)S->n;
        case 1: return (double)S->nsuper;
        case 2: return (double)S->nnzL;
        case 3: return (double)S->nnzL_true;
        case 4: return S->flops;
        case 5: return (double)S->max_front;
        case 6: return (double)S->rowptr[S->nsuper];
        default: return -1.0;
    }
}
/* which: 0 perm, 1 sfirst, 2 sparent, 3 rowptr, 4 rowidx, 5 Loff, 6 amap, 7 dpos */
void orc_sn_copy(void* h, int which, i64* out) {
    SN* S = (SN*)h;
    const i64* src = NULL; i64 cnt = 0;
    switch (which) {
        case 0: src = S->perm; cnt = S->n; break;
        case 1: src = S->sfirst; cnt = S->nsuper + 1; break;
        case 2: src = S->sparent; cnt = S->nsuper; break;
        case 3: src = S->rowptr; cnt = S->nsuper + 1; break;
        case 4: src = S->rowidx; cnt = S->rowptr[S->nsuper]; break;
        case 5: src = S->Loff; cnt = S->nsuper + 1; break;
        case 6: src = S->amap; cnt = S->nnzA; break;
        case 7: src = S->dpos; cnt = S->n; break;
    }
    if (src && cnt) memcpy(out, src, (size_t)cnt * sizeof(i64));
}

/* ======================================================================================
 * Numeric multifrontal Cholesky:  Q + diag shift  (update_delta_vecs! + ls_factor!(:definite),
 * schur.jl:64-87, julia.jl:28-46).  Ax: values of the input lower-triangular CSC matrix;
 * diag (optional, n, ORIGINAL numbering): replaces the diagonal, Q[i,i] = schur_diag[i] + delta.
 * Returns 1 when every pivot is > 0 (positive definite), else 0.
 * ====================================================================================== */
static int dense_chol_lower(double* A, i64 n, i64 lda) {
    for (i64 j = 0; j < n; j++) {
        double d = A[j + j * lda];
        for (i64 k = 0; k < j; k++) d -= A[j + k * lda] * A[j + k * lda];
        if (!(d > 0.0)) return 0;
        d = sqrt(d);
        A[j + j * lda] = d;
        for (i64 i = j + 1; i < n; i++) {
            double v = A[i + j * lda];
            for (i64 k = 0; k < j; k++) v -= A[i + k * lda] * A[j + k * lda];
            A[i + j * lda] = v / d;
        }
    }
    return 1;
}

int orc_sn_factorize(void* h, const double* Ax, const double* diag) {
    SN* S = (SN*)h;
    const i64 ns = S->nsuper;
    if (!S->L) S->L = (double*)xmalloc((size_t)(S->nnzL ? S->nnzL : 1) * sizeof(double));
    double* L = S->L;
    memset(L, 0, (size_t)S->nnzL * sizeof(double));
    for (i64 p = 0; p < S->nnzA; p++) if (S->amap[p] >= 0) L[S->amap[p]] += Ax[p];
    if (diag) for (i64 k = 0; k < S->n; k++) L[S->dpos[k]] = diag[S->perm[k]];
    double** CB = (double**)xcalloc((size_t)ns + 1, sizeof(double*));
    int ok = 1;
    for (i64 s = 0; s < ns && ok; s++) {
        const i64 f = S->sfirst[s], c = S->sfirst[s + 1] - f, r = S->rowptr[s + 1] - S->rowptr[s], N = c + r;
        double* P = L + S->Loff[s];
        double* U = r ? (double*)xcalloc((size_t)(r * r), sizeof(double)) : NULL;
        if (S->child_ptr[s + 1] > S->child_ptr[s]) {
            for (i64 j = 0; j < c; j++) S->pos[f + j] = j;
            for (i64 t = 0; t < r; t++) S->pos[S->rowidx[S->rowptr[s] + t]] = c + t;
        }
        for (i64 q = S->child_ptr[s]; q < S->child_ptr[s + 1]; q++) {       /* extend-add, ascending child order */
            const i64 ch = S->child_list[q];
            const i64 rc = S->rowptr[ch + 1] - S->rowptr[ch];
            const i64* rows = S->rowidx + S->rowptr[ch];
            const double* cb = CB[ch];
            for (i64 u = 0; u < rc; u++) {
                const i64 pj = S->pos[rows[u]];
                const double* col = cb + u * rc;
                if (pj < c) { double* dst = P + pj * N; for (i64 t = u; t < rc; t++) dst[S->pos[rows[t]]] += col[t]; }
                else { double* dst = U + (pj - c) * r - c; for (i64 t = u; t < rc; t++) dst[S->pos[rows[t]]] += col[t]; }
            }
            free(CB[ch]); CB[ch] = NULL;
        }
        /* pivot block, rows below it, update block */
        if (orc_sn_has_blas() && N >= 8) {
            int ci = (int)c, ri = (int)r, Ni = (int)N, info = 0;
            char lo = 'L', rt = 'R', tr = 'T', nn = 'N';
            double one = 1.0, mone = -1.0;
            f_dpotrf(&lo, &ci, P, &Ni, &info);
            if (info != 0) ok = 0;
            for (i64 j = 0; j < c && ok; j++) { const double d = P[j + j * N]; if (!(d > 0.0) || isinf(d)) ok = 0; }
            if (ok && r) {
                f_dtrsm(&rt, &lo, &tr, &nn, &ri, &ci, &one, P, &Ni, P + c, &Ni);
                f_dsyrk(&lo, &nn, &ri, &ci, &mone, P + c, &Ni, &one, U, &ri);
            }
        } else {
            if (!dense_chol_lower(P, c, N)) ok = 0;
            for (i64 j = 0; j < c && ok; j++) {
                const double d = P[j + j * N];
                if (isinf(d)) { ok = 0; break; }
                for (i64 i = c; i < N; i++) {
                    double v = P[i + j * N];
                    for (i64 k = 0; k < j; k++) v -= P[i + k * N] * P[j + k * N];
                    P[i + j * N] = v / d;
                }
            }
            if (ok)
                for (i64 k = 0; k < c; k++)
                    for (i64 u = 0; u < r; u++) {
                        const double l = P[c + u + k * N];
                        double* dst = U + u * r;
                        for (i64 t = u; t < r; t++) dst[t] -= P[c + t + k * N] * l;
                    }
        }
        CB[s] = U;
    }
    for (i64 s = 0; s < ns; s++) free(CB[s]);
    free(CB);
    return ok;
}

/* pivots L[j,j] in the final (new) numbering */
void orc_sn_diag(void* h, double* d) {
    SN* S = (SN*)h;
    for (i64 k = 0; k < S->n; k++) d[k] = S->L[S->dpos[k]];
}

/* ls_solve (julia.jl:99-113): x = P' (L L')^-1 P b */
void orc_sn_solve(void* h, const double* b, double* xout) {
    SN* S = (SN*)h;
    const i64 n = S->n, ns = S->nsuper;
    double* x = (double*)xmalloc((size_t)n * sizeof(double));
    double* tmp = (double*)xmalloc((size_t)(S->max_front + 1) * sizeof(double));
    for (i64 k = 0; k < n; k++) x[k] = b[S->perm[k]];
    const int blas = orc_sn_has_blas();
    char lo = 'L', nn = 'N', tr = 'T';
    int one_i = 1;
    double one = 1.0, zero = 0.0, mone = -1.0;
    for (i64 s = 0; s < ns; s++) {
        const i64 f = S->sfirst[s], c = S->sfirst[s + 1] - f, r = S->rowptr[s + 1] - S->rowptr[s], N = c + r;
        double* P = S->L + S->Loff[s];
        double* x1 = x + f;
        const i64* rows = S->rowidx + S->rowptr[s];
        if (blas && N >= 16) {
            int ci = (int)c, ri = (int)r, Ni = (int)N;
            f_dtrsv(&lo, &nn, &nn, &ci, P, &Ni, x1, &one_i);
            if (r) {
                f_dgemv(&nn, &ri, &ci, &one, P + c, &Ni, x1, &one_i, &zero, tmp, &one_i);
                for (i64 t = 0; t < r; t++) x[rows[t]] -= tmp[t];
            }
        } else {
            for (i64 j = 0; j < c; j++) {
                const double v = x1[j] / P[j + j * N];
                x1[j] = v;
                for (i64 i = j + 1; i < c; i++) x1[i] -= P[i + j * N] * v;
                for (i64 t = 0; t < r; t++) x[rows[t]] -= P[c + t + j * N] * v;
            }
        }
    }
    for (i64 s = ns - 1; s >= 0; s--) {
        const i64 f = S->sfirst[s], c = S->sfirst[s + 1] - f, r = S->rowptr[s + 1] - S->rowptr[s], N = c + r;
        double* P = S->L + S->Loff[s];
        double* x1 = x + f;
        const i64* rows = S->rowidx + S->rowptr[s];
        if (blas && N >= 16) {
            int ci = (int)c, ri = (int)r, Ni = (int)N;
            if (r) {
                for (i64 t = 0; t < r; t++) tmp[t] = x[rows[t]];
                f_dgemv(&tr, &ri, &ci, &mone, P + c, &Ni, tmp, &one_i, &one, x1, &one_i);
            }
            f_dtrsv(&lo, &tr, &nn, &ci, P, &Ni, x1, &one_i);
        } else {
            for (i64 j = c - 1; j >= 0; j--) {
                double v = x1[j];
                for (i64 i = j + 1; i < c; i++) v -= P[i + j * N] * x1[i];
                for (i64 t = 0; t < r; t++) v -= P[c + t + j * N] * x[rows[t]];
                x1[j] = v / P[j + j * N];
            }
        }
    }
    for (i64 k = 0; k < n; k++) xout[S->perm[k]] = x[k];
    free(x); free(tmp);
}
