"""ctypes wrapper around oracle/kkt_oracle.c -- the CPU restatement of the
reference's KKT path (see the header of kkt_oracle.c for the file:line map).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(onephase.jl_b200/) never imports this module.

Parity status: "parity unpinned" (no CHOLMOD, no Julia in this image; the
reference holds no numeric goldens for this path -- SURVEY.md 8c).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    so = os.path.join(_HERE, "libkkt_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("kkt_oracle.c", "snode.c", "Makefile")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.orc_form_system.restype = ctypes.c_void_p
        L.orc_form_system.argtypes = [ctypes.c_int64, ctypes.c_int64, _i64p, _i64p, _f64p,
                                      _i64p, _i64p, _f64p, _f64p, _f64p]
        L.orc_csc_nnz.restype = ctypes.c_int64
        L.orc_csc_nnz.argtypes = [ctypes.c_void_p]
        L.orc_csc_copy.argtypes = [ctypes.c_void_p, _i64p, _i64p, _f64p]
        L.orc_csc_free.argtypes = [ctypes.c_void_p]
        L.orc_analyze.restype = ctypes.c_void_p
        L.orc_analyze.argtypes = [ctypes.c_int64, _i64p, _i64p, _i64p]
        L.orc_factorize.restype = ctypes.c_int
        L.orc_factorize.argtypes = [ctypes.c_void_p, _f64p, _f64p, ctypes.c_int]
        L.orc_factor_free.argtypes = [ctypes.c_void_p]
        L.orc_factor_lnz.restype = ctypes.c_int64
        L.orc_factor_lnz.argtypes = [ctypes.c_void_p]
        L.orc_factor_flops.restype = ctypes.c_double
        L.orc_factor_flops.argtypes = [ctypes.c_void_p]
        L.orc_factor_diag.argtypes = [ctypes.c_void_p, _f64p]
        L.orc_ldlt_inertia_ok.restype = ctypes.c_int
        L.orc_ldlt_inertia_ok.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64]
        L.orc_solve.argtypes = [ctypes.c_void_p, _f64p, _f64p]
        L.orc_delta_loop.restype = ctypes.c_int
        L.orc_delta_loop.argtypes = [ctypes.c_void_p, _f64p, _f64p] + [ctypes.c_double] * 7 + \
            [_f64p, _i64p, _f64p, ctypes.c_int64]
        L.orc_direction.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                    _i64p, _i64p, _f64p, _i64p, _i64p, _f64p,
                                    _f64p, _f64p, ctypes.c_double, _f64p, _f64p, _f64p,
                                    ctypes.c_int, _f64p, _f64p, _f64p, _f64p]
        L.orc_extend_add.argtypes = [_f64p, ctypes.c_int64, ctypes.c_int64, _f64p, ctypes.c_int64,
                                     _i64p, ctypes.c_int64, _f64p]
        L.orc_snode_solve.argtypes = [ctypes.c_int64] + [_i64p] * 7 + [_f64p, _f64p, _f64p]
        _LIB = L
    return _LIB


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_i64p)


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_f64p)


# Default delta parameters: parameters.jl:147-158
DELTA_PARS = dict(delta_zero=0.0, delta_min=1e-12, delta_max=1e50, delta_start=1e-6,
                  inc=8.0, dec=1.0 / np.pi)


def form_system(J, H, y, s):
    """schur.jl:55-56.  J: scipy CSC (m x n); H: scipy CSC lower-triangular (n x n).
    Returns (Q as scipy csc holding full J'DJ + lower H, schur_diag)."""
    import scipy.sparse as sp
    J = sp.csc_matrix(J); H = sp.csc_matrix(H)
    J.sort_indices(); H.sort_indices()
    m, n = J.shape
    Jp, Jpp = _i(J.indptr); Ji, Jip = _i(J.indices); Jx, Jxp = _f(J.data)
    Hp, Hpp = _i(H.indptr); Hi, Hip = _i(H.indices); Hx, Hxp = _f(H.data)
    yy, yp = _f(y); ss, spp = _f(s)
    L = lib()
    h = L.orc_form_system(n, m, Jpp, Jip, Jxp, Hpp, Hip, Hxp, yp, spp)
    nnz = L.orc_csc_nnz(h)
    cp = np.empty(n + 1, np.int64); ri = np.empty(nnz, np.int64); nz = np.empty(nnz, np.float64)
    L.orc_csc_copy(h, cp.ctypes.data_as(_i64p), ri.ctypes.data_as(_i64p), nz.ctypes.data_as(_f64p))
    L.orc_csc_free(h)
    Q = sp.csc_matrix((nz, ri, cp), shape=(n, n))
    Q.has_sorted_indices = True
    return Q, Q.diagonal()


class Factor:
    """Symbolic + numeric sparse factor of Symmetric(Q,:L) (julia.jl:21-97)."""

    def __init__(self, Q, perm=None):
        import scipy.sparse as sp
        Q = sp.csc_matrix(Q)
        Q.sort_indices()
        self.n = Q.shape[0]
        self._Qp, p1 = _i(Q.indptr)
        self._Qi, p2 = _i(Q.indices)
        self.nnz = Q.nnz
        if perm is not None:
            self._perm, pp = _i(perm)
        else:
            pp = None
        self._h = lib().orc_analyze(self.n, p1, p2, pp)
        self.mode = 0

    def __del__(self):
        try:
            if self._h:
                lib().orc_factor_free(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def lnz(self):
        return lib().orc_factor_lnz(self._h)

    @property
    def flops(self):
        return lib().orc_factor_flops(self._h)

    def factorize(self, nzval, diag_override=None, mode="chol"):
        """Returns 1 if completed (PD for chol), else 0 (julia.jl:39-45,61-63)."""
        x, xp = _f(nzval)
        assert x.shape[0] == self.nnz
        dp = None
        if diag_override is not None:
            d, dp = _f(diag_override)
        self.mode = 0 if mode == "chol" else 1
        return lib().orc_factorize(self._h, xp, dp, self.mode)

    def ldlt_inertia_ok(self, n, m):
        return lib().orc_ldlt_inertia_ok(self._h, n, m)

    def diag(self):
        d = np.empty(self.n)
        lib().orc_factor_diag(self._h, d.ctypes.data_as(_f64p))
        return d

    def solve(self, b):
        bb, bp = _f(b)
        x = np.empty(self.n)
        lib().orc_solve(self._h, bp, x.ctypes.data_as(_f64p))
        return x

    def delta_loop(self, nzval, schur_diag, delta_prev, **kw):
        """ipopt_strategy! (delta_strategy.jl:37-114).  Returns
        (status, num_fac, delta, deltas_tried)."""
        p = dict(DELTA_PARS); p.update(kw)
        x, xp = _f(nzval); d, dp = _f(schur_diag)
        delta = ctypes.c_double(0.0); nf = ctypes.c_int64(0)
        rec = np.zeros(600)
        st = lib().orc_delta_loop(self._h, xp, dp, float(delta_prev), p["delta_zero"], p["delta_min"],
                                  p["delta_max"], p["delta_start"], p["inc"], p["dec"],
                                  ctypes.cast(ctypes.byref(delta), _f64p),
                                  ctypes.cast(ctypes.byref(nf), _i64p),
                                  rec.ctypes.data_as(_f64p), 600)
        self.mode = 0
        status = {1: "success", 0: "failure", -1: "max_it"}[st]
        return status, int(nf.value), float(delta.value), rec[: nf.value].copy()

    def direction(self, J, H, y, s, delta, dual_r, primal_r, comp_r, n_refine=3):
        """compute_direction_implementation! (schur.jl:89-128) + N err."""
        import scipy.sparse as sp
        J = sp.csc_matrix(J); H = sp.csc_matrix(H)
        J.sort_indices(); H.sort_indices()
        m, n = J.shape
        Jp, Jpp = _i(J.indptr); Ji, Jip = _i(J.indices); Jx, Jxp = _f(J.data)
        Hp, Hpp = _i(H.indptr); Hi, Hip = _i(H.indices); Hx, Hxp = _f(H.data)
        yy, yp = _f(y); ss, spp = _f(s)
        a, ap = _f(dual_r); b, bp = _f(primal_r); c, cp = _f(comp_r)
        dx = np.empty(n); dy = np.empty(m); ds = np.empty(m); err = np.empty(6)
        lib().orc_direction(self._h, n, m, Jpp, Jip, Jxp, Hpp, Hip, Hxp, yp, spp, float(delta),
                            ap, bp, cp, int(n_refine),
                            dx.ctypes.data_as(_f64p), dy.ctypes.data_as(_f64p),
                            ds.ctypes.data_as(_f64p), err.ctypes.data_as(_f64p))
        return dx, dy, ds, err


# ---------------------------------------------------------------------------
# SURVEY 8 f4 restatements (numpy; small cases)
# ---------------------------------------------------------------------------
def eval_diag_J_T_J(J, diag_vals):
    """utils/eval.jl:89-100: di[i] += a[j]^2 * diag_vals[j] over the stored entries of column i,
    rows ascending (square first, then the product, then the sum)."""
    import scipy.sparse as sp
    J = sp.csc_matrix(J); J.sort_indices()
    d = np.asarray(diag_vals, dtype=np.float64)
    n = J.shape[1]
    di = np.zeros(n)
    for i in range(n):
        acc = 0.0
        for p in range(J.indptr[i], J.indptr[i + 1]):
            a = J.data[p]
            acc = acc + (a * a) * d[J.indices[p]]
        di[i] = acc
    return di


def compute_schur_diag(J, H, y, s):
    """kkt_system_solver.jl:296-300: diag(get_lag_hess(iter)) + eval_diag_J_T_J(iter, y ./ s)."""
    import scipy.sparse as sp
    return sp.csc_matrix(H).diagonal() + eval_diag_J_T_J(J, np.asarray(y, float) / np.asarray(s, float))


def estimate_y_tilde(J, g, lam=1e-4):
    """init/guess-vars.jl:128-169 (Cholesky branch): H = lam I + J'J; dx = H \\ -g; y = -J dx."""
    import scipy.sparse as sp
    J = sp.csc_matrix(J)
    m, n = J.shape
    Q, sd = form_system(J, sp.identity(n, format="csc") * lam, np.ones(m), np.ones(m))
    QL = sp.tril(Q, format="csc"); QL.sort_indices()
    F = Factor(QL)
    if F.factorize(QL.data, mode="chol") != 1:
        return np.ones(m)
    dx = F.solve(-np.asarray(g, dtype=np.float64))
    return -(J @ dx)


# ---------------------------------------------------------------------------
# SURVEY 8 f3 restatements (numpy, explicit loops where the operation order matters)
# ---------------------------------------------------------------------------
def system_rhs(J, y, s, grad, cons, mu, a_norm_penalty, eta_P, eta_D, eta_mu):
    """System_rhs(it, reduct) (kkt_system_solver/system_rhs.jl:57-73) with eval_grad_lag / eval_grad_r
    (utils/eval.jl:59-63,136-142); J'y and J'1 accumulate over the rows ascending like Julia's CSC product."""
    import scipy.sparse as sp
    J = sp.csc_matrix(J); J.sort_indices()
    y = np.asarray(y, float); s = np.asarray(s, float)
    n = J.shape[1]
    mu_t = mu * eta_mu
    dual = np.empty(n)
    for j in range(n):
        jty = 0.0; jt1 = 0.0
        for p in range(J.indptr[j], J.indptr[j + 1]):
            jty = jty + J.data[p] * y[J.indices[p]]
            jt1 = jt1 + J.data[p]
        gl = (grad[j] - jty) + mu_t * (a_norm_penalty * jt1)
        dual[j] = (-gl) * (1.0 - eta_D)
    primal = -(np.asarray(cons, float) - s) * (1.0 - eta_P)
    comp = mu_t - s * y
    return dual, primal, comp


def step_bounds(s, dx, dy, ds, frac_bd, predict_exp):
    """frac_boundary.jl:3-35: inf norms, lb_s = frac_bd * min.(s, |dx| * |dx|^ex), simple_max_step(s, ds, lb_s)."""
    ndx = float(np.abs(dx).max()) if len(dx) else 0.0
    lb = frac_bd * np.minimum(s, ndx * ndx ** predict_exp)
    ratio = max(1.0, float(np.max(-np.asarray(ds) / (np.asarray(s) - lb)))) if len(s) else 1.0
    return dict(norm_dx=ndx, norm_dy=float(np.abs(dy).max()) if len(dy) else 0.0,
                norm_ds=float(np.abs(ds).max()) if len(ds) else 0.0, max_step_s=1.0 / ratio)
