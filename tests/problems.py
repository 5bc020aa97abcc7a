"""Seeded synthetic KKT inputs of the shapes BASELINE.json names.

Each generator returns a KKTProblem: the data `form_system!` reads from a
Class_iterate (J m x n CSC, H n x n lower-triangular CSC, y, s;
Class_iterate.jl:4-84, schur.jl:47-62), rhs triples in System_rhs layout
(system_rhs.jl:39-74) and the previous delta (get_delta(iter)).

Row order of J follows Class_cutest.jl:392,458-459:
    [ c(x)[l_i] - l ; u - c(x)[u_i] ; x[lb_i] - lb ; ub - x[ub_i] ]
so an equality yields two parallel rows (+grad, -grad) and bounds yield +-e_i rows.
"""
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np
import scipy.sparse as sp


@dataclass
class KKTProblem:
    name: str
    J: sp.csc_matrix
    H: sp.csc_matrix          # lower triangular, incl. diagonal where present
    y: np.ndarray
    s: np.ndarray
    rhs: List[Tuple[np.ndarray, np.ndarray, np.ndarray]] = field(default_factory=list)
    delta_prev: float = 0.0

    @property
    def n(self):
        return self.J.shape[1]

    @property
    def m(self):
        return self.J.shape[0]


def _csc(A):
    A = sp.csc_matrix(A, dtype=np.float64)
    A.sum_duplicates()
    A.sort_indices()
    A.indptr = A.indptr.astype(np.int64)
    A.indices = A.indices.astype(np.int64)
    return A


def _ys(rng, m, lo=1e-4, hi=1e2):
    y = np.exp(rng.uniform(np.log(lo), np.log(hi), m))
    s = np.exp(rng.uniform(np.log(lo), np.log(hi), m))
    return y, s


def _rhs(rng, n, m, k=2):
    return [(rng.standard_normal(n), rng.standard_normal(m), rng.standard_normal(m)) for _ in range(k)]


def _rows_from_parts(n, Jc, l_idx, u_idx, lb_idx, ub_idx):
    """[J_c[l_i]; -J_c[u_i]; I[lb_i]; -I[ub_i]]  (Class_cutest.jl:451-503)."""
    Jc = sp.csr_matrix(Jc) if Jc is not None else sp.csr_matrix((0, n))
    I = sp.identity(n, format="csr")
    parts = []
    if len(l_idx):
        parts.append(Jc[l_idx])
    if len(u_idx):
        parts.append(-Jc[u_idx])
    if len(lb_idx):
        parts.append(I[lb_idx])
    if len(ub_idx):
        parts.append(-I[ub_idx])
    return _csc(sp.vstack(parts)) if parts else _csc(sp.csr_matrix((0, n)))


# ---------------------------------------------------------------------------
# C1: the reference's toy problems (README.md:34-45, test/problems.jl:108-296)
# ---------------------------------------------------------------------------
def toy(name="toy_lp1", seed=0, h_scale=0.0):
    rng = np.random.default_rng(seed)
    a = np.arange
    if name == "readme":
        # min x  s.t. x^2 - 1 >= 0 (NL), x + 1 >= 0 ; evaluated at x0
        x0 = 1.5
        J = _csc(np.array([[2 * x0], [1.0]]))
        y, s = _ys(rng, 2, 0.1, 10)
        H = _csc(np.array([[-2.0 * y[0]]]))  # -y_1 * d2(x^2-1)
        return KKTProblem(name, J, H, y, s, _rhs(rng, 1, 2))
    n = 1 if name == "toy_lp0" else 2
    spec = {
        # name: (Jc rows, l_idx, u_idx, lb_idx, ub_idx)
        "toy_lp0": ([[1.0]], [0], [], [], []),
        "toy_lp1": ([[1.0, 1.0]], [], [0], [0, 1], []),
        "toy_lp2": ([[1.0, 1.0]], [], [0], [0, 1], [0, 1]),
        "toy_lp3": ([[1.0, 1.0]], [0], [0], [0, 1], [0, 1]),
        "toy_lp4": ([[1.0, 1.0]], [0], [0], [0, 1], [0, 1]),
        "toy_lp5": ([[1.0, 1.0], [32.5, 32.5], [3.0, 3.0]], [0, 1], [0, 1, 2], [0, 1], [0, 1]),
        "toy_lp6": ([[1.0, 1.0], [5.5, 5.5]], [0, 1], [0, 1], [0, 1], [0, 1]),
        "toy_lp7": ([[2.0, 1.0]], [0], [0], [0, 1], [0, 1]),
        "toy_lp8": ([[1.0, 1.0], [5.5, 5.5]], [0], [1], [0, 1], [0, 1]),
    }[name]
    Jc, l_idx, u_idx, lb_idx, ub_idx = spec
    J = _rows_from_parts(n, np.array(Jc), l_idx, u_idx, lb_idx, ub_idx)
    m = J.shape[0]
    y, s = _ys(rng, m, 0.1, 10)
    if h_scale == 0.0:
        H = sp.csc_matrix((n, n))  # LP: structurally empty Hessian (gotcha 9.7-2)
        H.indptr = H.indptr.astype(np.int64); H.indices = H.indices.astype(np.int64)
    else:
        H = _csc(np.tril(h_scale * np.ones((n, n))))
    return KKTProblem(name, J, H, y, s, _rhs(rng, n, m))


TOY_NAMES = ["readme"] + ["toy_lp%d" % i for i in range(9)]


# ---------------------------------------------------------------------------
# C2: banded collocation stand-in for COPS chain / camshape
# ---------------------------------------------------------------------------
def chain(nh=2500, seed=0, neq=3, neg_curv=0.0, offdiag_curv=0.0):
    """nh intervals, 4 variables per node, `neq` equalities per interval each
    touching the 8 variables of two adjacent nodes (duplicated with sign), plus
    lower and upper bound rows on every variable.  n = 4(nh+1).  H is block
    diagonal (4x4 lower blocks), shifted by -neg_curv*I."""
    rng = np.random.default_rng(seed)
    n = 4 * (nh + 1)
    k = np.arange(nh)
    rows = np.repeat(np.arange(nh * neq), 8)
    base = np.repeat(4 * k, neq)
    cols = (base[:, None] + np.arange(8)[None, :]).ravel()
    vals = rng.standard_normal(rows.shape[0])
    Jc = sp.csr_matrix((vals, (rows, cols)), shape=(nh * neq, n))
    eq = np.arange(nh * neq)
    allv = np.arange(n)
    J = _rows_from_parts(n, Jc, eq, eq, allv, allv)
    m = J.shape[0]
    # block-diagonal 4x4 lower blocks
    bi, bj = np.tril_indices(4)
    nb = nh + 1
    B = rng.standard_normal((nb, 4, 4))
    # neg_curv > 0 makes the Hessian blocks indefinite so the delta loop has work to do
    B = B @ B.transpose(0, 2, 1) * 0.25 + (0.5 - neg_curv) * np.eye(4)
    # offdiag_curv = g adds g*(ones - I): eigenvalues 3g and -g (x3) with a zero diagonal, i.e.
    # negative curvature that the diagonal test of the delta rule cannot see (probe fails, x8 retries)
    B = B + offdiag_curv * (np.ones((4, 4)) - np.eye(4))
    hr = (4 * np.arange(nb)[:, None] + bi[None, :]).ravel()
    hc = (4 * np.arange(nb)[:, None] + bj[None, :]).ravel()
    H = _csc(sp.csc_matrix((B[:, bi, bj].ravel(), (hr, hc)), shape=(n, n)))
    y, s = _ys(rng, m)
    return KKTProblem("chain_nh%d" % nh, J, H, y, s, _rhs(rng, n, m))


# ---------------------------------------------------------------------------
# C3: sparse convex QP with 2-D locality
# ---------------------------------------------------------------------------
def sparse_qp(n=200_000, m_gen=100_000, nnz_row=10, seed=0, bounds=True, local=True, win=7):
    rng = np.random.default_rng(seed)
    W = int(np.ceil(np.sqrt(n))) + 1          # 448 for n = 200000
    full_rows = n // W                        # rows of the grid that are complete
    if local and full_rows > win and W > win:
        h = win // 2
        cr = rng.integers(h, full_rows - h, m_gen)
        cc = rng.integers(h, W - h, m_gen)
        # nnz_row distinct offsets out of win*win
        key = rng.random((m_gen, win * win))
        off = np.argpartition(key, nnz_row, axis=1)[:, :nnz_row]
        dr = off // win - h
        dc = off % win - h
        cols = ((cr[:, None] + dr) * W + (cc[:, None] + dc)).ravel()
    else:
        key = rng.random((m_gen, 0))
        cols = np.empty((m_gen, nnz_row), np.int64)
        for r in range(m_gen):  # uniform-random columns (fill study only)
            cols[r] = rng.choice(n, nnz_row, replace=False)
        cols = cols.ravel()
    rows = np.repeat(np.arange(m_gen), nnz_row)
    vals = rng.standard_normal(rows.shape[0])
    Jc = sp.csr_matrix((vals, (rows, cols)), shape=(m_gen, n))
    g = np.arange(m_gen)
    lb = np.arange(n) if bounds else np.arange(0)
    J = _rows_from_parts(n, Jc, g, [], lb, [])
    m = J.shape[0]
    H = _csc(sp.diags(rng.uniform(0.1, 1.0, n)))
    y, s = _ys(rng, m)
    return KKTProblem("sparse_qp_n%d" % n, J, H, y, s, _rhs(rng, n, m))


# ---------------------------------------------------------------------------
# C4: COPS electron (benchmark/COPS/2-electron.jl:9-25), dense Hessian
# ---------------------------------------------------------------------------
def elec(n_p=400, seed=0):
    rng = np.random.default_rng(seed)
    P = rng.standard_normal((n_p, 3))
    P /= np.linalg.norm(P, axis=1, keepdims=True)
    n = 3 * n_p
    # variable order: x_1..x_np, y_1..y_np, z_1..z_np
    rows = np.repeat(np.arange(n_p), 3)
    cols = (np.arange(n_p)[:, None] + n_p * np.arange(3)[None, :]).ravel()
    Jc = sp.csr_matrix((2.0 * P.ravel(), (rows, cols)), shape=(n_p, n))
    e = np.arange(n_p)
    J = _rows_from_parts(n, Jc, e, e, [], [])
    m = J.shape[0]
    y, s = _ys(rng, m, 1e-2, 1e1)
    # Hessian of sum_{i<j} 1/|p_i - p_j|
    D = P[:, None, :] - P[None, :, :]
    r2 = (D ** 2).sum(-1)
    np.fill_diagonal(r2, 1.0)
    r5 = r2 ** 2.5
    blk = (3.0 * D[:, :, :, None] * D[:, :, None, :] - r2[:, :, None, None] * np.eye(3)) / r5[:, :, None, None]
    idx = np.arange(n_p)
    blk[idx, idx] = 0.0
    Hd = np.zeros((n, n))
    diag_blk = blk.sum(1)
    for a in range(3):
        for b in range(3):
            Hd[a * n_p:(a + 1) * n_p, b * n_p:(b + 1) * n_p] = -blk[:, :, a, b]
            Hd[a * n_p + idx, b * n_p + idx] = diag_blk[:, a, b]
    # - sum y_k d2 a_k : l-row a = c - 1 -> -2 y_l I ; u-row a = 1 - c -> +2 y_u I
    mult = 2.0 * (y[n_p:] - y[:n_p])
    for a in range(3):
        Hd[a * n_p + idx, a * n_p + idx] += mult
    H = _csc(sp.csc_matrix(np.tril(Hd)))
    return KKTProblem("elec_np%d" % n_p, J, H, y, s, _rhs(rng, n, m))


# ---------------------------------------------------------------------------
# C5: 3-D PDE-constrained optimal control, control eliminated
# ---------------------------------------------------------------------------
def laplace3d(N):
    n = N ** 3
    idx = np.arange(n).reshape(N, N, N)
    rows = [np.arange(n)]
    cols = [np.arange(n)]
    vals = [np.full(n, 6.0)]
    for ax in range(3):
        a = np.take(idx, np.arange(N - 1), axis=ax).ravel()
        b = np.take(idx, np.arange(1, N), axis=ax).ravel()
        rows += [a, b]; cols += [b, a]
        vals += [np.full(a.shape[0], -1.0)] * 2
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))


def pde_control(N=100, seed=0):
    rng = np.random.default_rng(seed)
    n = N ** 3
    A = laplace3d(N)
    allv = np.arange(n)
    J = _rows_from_parts(n, A, allv, allv, allv, allv)
    m = J.shape[0]
    H = _csc(sp.diags(rng.uniform(0.5, 1.5, n)))
    y, s = _ys(rng, m)
    return KKTProblem("pde_control_N%d" % N, J, H, y, s, _rhs(rng, n, m))


def grid_nd_perm(N, leaf=4):
    """Geometric nested-dissection order for the N^3 grid whose M pattern reaches
    L1-distance 2 (separators two planes thick).  Returns perm (new -> old)."""
    order = []

    def rec(lo, hi):
        ext = [hi[a] - lo[a] for a in range(3)]
        if max(ext) <= leaf:
            g = np.mgrid[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]].reshape(3, -1)
            order.append((g[0] * N + g[1]) * N + g[2])
            return
        ax = int(np.argmax(ext))
        mid = (lo[ax] + hi[ax]) // 2
        l1 = list(hi); l1[ax] = mid - 1
        rec(lo, l1)
        l2 = list(lo); l2[ax] = mid + 1
        rec(l2, hi)
        slo = list(lo); slo[ax] = mid - 1
        shi = list(hi); shi[ax] = mid + 1
        g = np.mgrid[slo[0]:shi[0], slo[1]:shi[1], slo[2]:shi[2]].reshape(3, -1)
        order.append((g[0] * N + g[1]) * N + g[2])

    rec([0, 0, 0], [N, N, N])
    return np.concatenate(order).astype(np.int64)


# ---------------------------------------------------------------------------
# IPM-like sequence: same pattern, drifting (y, s), occasionally indefinite H
# ---------------------------------------------------------------------------
def ipm_sequence(prob: KKTProblem, steps=6, seed=0, indefinite_every=3, shift=5.0, offdiag=0.0):
    """Yield KKTProblems sharing prob's sparsity pattern, mimicking outer
    iterations: s*y is driven towards a shrinking mu, J's values drift, and every
    `indefinite_every`-th iterate gets `shift` subtracted from H's stored diagonal
    so the delta loop (delta_strategy.jl:37-114) has work to do; `offdiag` is added to H's stored
    off-diagonal entries on those iterates instead (negative curvature the diagonal test of the
    delta rule cannot see: the probe at delta = 0 fails and the x8 retries run).  delta_prev is
    left at 0; the caller threads the accepted delta through like one_phase.jl:205-206."""
    rng = np.random.default_rng(seed)
    y = prob.y.copy(); s = prob.s.copy()
    mu = float(np.exp(np.mean(np.log(y * s))))
    H0 = prob.H
    # positions of the stored diagonal entries of H (lower CSC: first entry of a column if row == col)
    cols = np.repeat(np.arange(prob.n), np.diff(H0.indptr))
    dpos = np.nonzero(H0.indices == cols)[0]
    opos = np.nonzero(H0.indices != cols)[0]
    for t in range(steps):
        mu *= 0.3
        s = np.sqrt(s * (mu / y)) * np.exp(0.3 * rng.standard_normal(s.shape[0]))
        y = mu / s * np.exp(0.1 * rng.standard_normal(s.shape[0]))
        H = H0.copy()
        if indefinite_every and (t % indefinite_every) == indefinite_every - 1:
            H.data = H.data.copy()
            H.data[dpos] -= shift
            H.data[opos] += offdiag
        J = prob.J.copy()
        J.data = prob.J.data * (1.0 + 0.01 * rng.standard_normal(J.nnz))
        yield KKTProblem("%s_it%d" % (prob.name, t), J, H, y.copy(), s.copy(),
                         _rhs(rng, prob.n, prob.m, 2), 0.0)
