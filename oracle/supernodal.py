"""Supernodal (multifrontal) CPU restatement of the reference's KKT path with BLAS-3 dense
kernels -- the performance class of what the reference actually runs: `cholesky(Symmetric(Q,:L))`
(linear_system_solvers/julia.jl:34) is CHOLMOD's supernodal factorisation on top of a threaded
BLAS.  oracle/kkt_oracle.c (scalar, up-looking, one core) is the small-case checker; this module
(on oracle/snode.c) is the large-case checker and the CPU BASELINE that bench.py times with all
host cores (scipy's OpenBLAS threads).

TEST INFRASTRUCTURE ONLY (same rule as oracle.py): imported by tests/ and by bench.py's
cpu_baseline / --impl reference legs, never by the product package -- and it does not use the
product either: ordering (METIS_NodeND / minimum degree / natural), elimination tree, column
counts, supernodes and row structures all come from oracle/snode.c.  Parity status: "parity
unpinned" (no CHOLMOD, no Julia here); tests/test_oracle.py pins it against kkt_oracle.c.

What follows which reference lines:
  SupernodalFactor(...)   CHOLMOD analyze (redone per ls_factor!, recycle = false)   julia.jl:34, parameters.jl:38
  delta_loop   ipopt_strategy!                      IPM/delta_strategy.jl:37-114, parameters.jl:147-158
  factorize    update_delta_vecs! + ls_factor!      schur.jl:64-87, julia.jl:28-46 (1 = PD, 0 = PosDefException)
  solve        ls_solve                             julia.jl:99-113
  direction    compute_direction_implementation!    schur.jl:89-182, kkt_system_solver.jl:27-96
"""
import ctypes
import os

import numpy as np
import scipy.sparse as sp

from . import oracle as _orc

_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)

DELTA_PARS = dict(delta_zero=0.0, delta_min=1e-12, delta_max=1e50, delta_start=1e-6, inc=8.0, dec=1.0 / np.pi)

_BLAS_SET = False


def _capsule_pointer(mod, name):
    cap = mod.__pyx_capi__[name]
    api = ctypes.pythonapi
    api.PyCapsule_GetName.restype = ctypes.c_char_p
    api.PyCapsule_GetName.argtypes = [ctypes.py_object]
    api.PyCapsule_GetPointer.restype = ctypes.c_void_p
    api.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
    return api.PyCapsule_GetPointer(cap, api.PyCapsule_GetName(cap))


def lib():
    """libkkt_oracle.so with the snode.c entry points typed and the BLAS / LAPACK routines of
    scipy's OpenBLAS (dpotrf, dtrsm, dsyrk, dgemv, dtrsv) handed over as function pointers."""
    global _BLAS_SET
    L = _orc.lib()
    if not _BLAS_SET:
        vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        L.orc_sn_set_blas.argtypes = [vp] * 5
        L.orc_sn_has_blas.restype = ci
        L.orc_order_metis.restype = ci
        L.orc_order_metis.argtypes = [i64, _i64p, _i64p, _i64p]
        L.orc_order_mindeg.restype = ci
        L.orc_order_mindeg.argtypes = [i64, _i64p, _i64p, _i64p]
        L.orc_sn_analyze.restype = vp
        L.orc_sn_analyze.argtypes = [i64, _i64p, _i64p, _i64p, ci]
        L.orc_sn_free.argtypes = [vp]
        L.orc_sn_info.restype = ctypes.c_double
        L.orc_sn_info.argtypes = [vp, ci]
        L.orc_sn_copy.argtypes = [vp, ci, _i64p]
        L.orc_sn_factorize.restype = ci
        L.orc_sn_factorize.argtypes = [vp, _f64p, _f64p]
        L.orc_sn_diag.argtypes = [vp, _f64p]
        L.orc_sn_solve.argtypes = [vp, _f64p, _f64p]
        from scipy.linalg import cython_blas, cython_lapack
        L.orc_sn_set_blas(_capsule_pointer(cython_lapack, "dpotrf"), _capsule_pointer(cython_blas, "dtrsm"),
                          _capsule_pointer(cython_blas, "dsyrk"), _capsule_pointer(cython_blas, "dgemv"),
                          _capsule_pointer(cython_blas, "dtrsv"))
        assert L.orc_sn_has_blas() == 1
        _BLAS_SET = True
    return L


def blas_threads(n=None):
    """Context manager that pins the BLAS thread count (torchrun exports OMP_NUM_THREADS=1, which
    would silently serialise the baseline); n = None -> all host cores."""
    from threadpoolctl import threadpool_limits
    return threadpool_limits(limits=int(n or os.cpu_count() or 1), user_api="blas")


def _ip(a):
    return a.ctypes.data_as(_i64p)


def order(QL, method="metis"):
    """Fill-reducing ordering of the pattern of the lower-triangular CSC matrix QL, computed by
    the oracle itself: 'metis' (METIS_NodeND), 'mindeg' (exact minimum degree, small n),
    'natural'.  Returns perm with perm[new] = old."""
    QL = sp.csc_matrix(QL)
    n = QL.shape[0]
    if method == "natural":
        return np.arange(n, dtype=np.int64)
    Ap = np.ascontiguousarray(QL.indptr, dtype=np.int64)
    Ai = np.ascontiguousarray(QL.indices, dtype=np.int64)
    perm = np.empty(n, np.int64)
    fn = {"metis": lib().orc_order_metis, "mindeg": lib().orc_order_mindeg}[method]
    if fn(n, _ip(Ap), _ip(Ai), _ip(perm)) != 1:
        raise RuntimeError("ordering %s failed" % method)
    return perm


class SupernodalFactor:
    def __init__(self, QL, perm=None, ordering="metis", relax=True):
        """QL: scipy CSC, lower triangle (entries above the diagonal are ignored).
        perm: explicit permutation (perm[new] = old) or None -> `ordering` is computed here.
        The whole analysis (ordering, elimination tree, column counts, supernodes, row
        structures) runs in this constructor, like CHOLMOD's analyze inside cholesky()."""
        QL = sp.csc_matrix(QL)
        self.n = QL.shape[0]
        self._Ap = np.ascontiguousarray(QL.indptr, dtype=np.int64)
        self._Ai = np.ascontiguousarray(QL.indices, dtype=np.int64)
        self.nnz = int(self._Ap[-1])
        if perm is None:
            perm = order(QL, ordering)
        p = np.ascontiguousarray(perm, dtype=np.int64)
        self._h = lib().orc_sn_analyze(self.n, _ip(self._Ap), _ip(self._Ai), _ip(p), 1 if relax else 0)
        self.delta = 0.0
        self._factored = False

    def __del__(self):
        try:
            if self._h:
                lib().orc_sn_free(self._h)
                self._h = None
        except Exception:
            pass

    def info(self, key):
        keys = {"n": 0, "nsuper": 1, "nnzL": 2, "nnzL_true": 3, "flops": 4, "max_front": 5, "sum_rows": 6}
        return lib().orc_sn_info(self._h, keys[key])

    def array(self, name):
        which = {"perm": 0, "sfirst": 1, "sparent": 2, "rowptr": 3, "rowidx": 4, "Loff": 5, "amap": 6, "dpos": 7}[name]
        ns = int(self.info("nsuper"))
        cnt = {"perm": self.n, "sfirst": ns + 1, "sparent": ns, "rowptr": ns + 1, "rowidx": int(self.info("sum_rows")),
               "Loff": ns + 1, "amap": self.nnz, "dpos": self.n}[name]
        out = np.empty(max(cnt, 1), np.int64)
        lib().orc_sn_copy(self._h, which, _ip(out))
        return out[:cnt]

    @property
    def perm(self):
        return self.array("perm")

    # -- update_delta_vecs! + ls_factor!(:definite)
    def factorize(self, nzval, delta=0.0, schur_diag=None):
        """Factorise Q with its diagonal replaced by schur_diag + delta (schur.jl:76) when
        schur_diag is given, else Q + nothing (delta must then be 0).  Returns 1 when positive
        definite (every pivot > 0), else 0."""
        x = np.ascontiguousarray(nzval, dtype=np.float64)
        assert x.shape[0] == self.nnz
        dp = None
        if schur_diag is not None:
            d = np.ascontiguousarray(np.asarray(schur_diag, dtype=np.float64) + delta)
            dp = d.ctypes.data_as(_f64p)
        else:
            assert delta == 0.0
        ok = lib().orc_sn_factorize(self._h, x.ctypes.data_as(_f64p), dp)
        self._factored = ok == 1
        self.delta = delta
        return ok

    def diag(self):
        d = np.empty(self.n)
        lib().orc_sn_diag(self._h, d.ctypes.data_as(_f64p))
        return d

    # -- ls_solve
    def solve(self, b):
        assert self._factored
        bb = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty(self.n)
        lib().orc_sn_solve(self._h, bb.ctypes.data_as(_f64p), x.ctypes.data_as(_f64p))
        return x

    # -- ipopt_strategy!
    def delta_loop(self, nzval, schur_diag, delta_prev, **kw):
        p = dict(DELTA_PARS); p.update(kw)
        tried = []
        num_fac = 0
        tau = 1.5 * float(np.min(schur_diag))
        delta = p["delta_zero"]
        if tau > 0.0:
            tau = 0.0
            ok = self.factorize(nzval, delta, schur_diag); num_fac += 1; tried.append(delta)
            if ok == 1:
                return "success", num_fac, delta, np.array(tried)
        for i in range(1, 501):
            if i == 1:
                delta = float(np.maximum(p["delta_min"] - tau, delta_prev * p["dec"])) if delta_prev != 0.0 else p["delta_start"] - tau
            else:
                delta = delta * p["inc"]
            ok = self.factorize(nzval, delta, schur_diag); num_fac += 1; tried.append(delta)
            if ok == 1:
                return "success", num_fac, delta, np.array(tried)
            if delta > p["delta_max"]:
                return "failure", num_fac, delta, np.array(tried)
        return "max_it", num_fac, delta, np.array(tried)

    # -- compute_direction_implementation! + update_kkt_error!
    def direction(self, J, H, y, s, delta, dual_r, primal_r, comp_r, n_refine=3):
        J = sp.csr_matrix(J); JT = sp.csr_matrix(J.T)
        Hl = sp.csr_matrix(H)
        Hs = Hl + sp.tril(Hl, -1).T                     # H_sym v = L v + L' v - diag(L) v (eval.jl:221-230)
        y = np.asarray(y, float); s = np.asarray(s, float)
        rD, rP, rC = (np.asarray(v, float) for v in (dual_r, primal_r, comp_r))
        sig = y / s
        p = rP + rC / y
        b = rD + JT @ (rP * sig + rC / s)
        x = np.zeros(self.n)
        r = b.copy()
        for k in range(n_refine):
            x = x + self.solve(r)
            if k + 1 < n_refine:      # the reference also forms the last residual but only prints it
                r = b - (JT @ (sig * (J @ x)) + Hs @ x + delta * x)
        dx = x
        Jdx = J @ dx
        dy = -(Jdx - p) * sig
        ds = Jdx - rP
        eD = (delta * dx + Hs @ dx - JT @ dy) - rD
        eP = Jdx - ds - rP
        eM = s * dy + y * ds - rC
        nrm = lambda v: float(np.abs(v).max()) if v.size else 0.0      # noqa: E731
        overall = max(nrm(eD), nrm(eP), nrm(eM))
        rhs_norm = max(nrm(rD), nrm(rP), nrm(rC))
        err = np.array([nrm(eD), nrm(eP), nrm(eM), overall, rhs_norm, overall / rhs_norm if rhs_norm else np.nan])
        return dx, dy, ds, err
