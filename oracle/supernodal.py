"""Supernodal (multifrontal) CPU restatement of the reference's KKT path with BLAS-3 dense
kernels -- the performance class of what the reference actually runs: `cholesky(Symmetric(Q,:L))`
(linear_system_solvers/julia.jl:34) is CHOLMOD's supernodal factorisation on top of a threaded
BLAS.  oracle/kkt_oracle.c (scalar, up-looking, one core) is the checker; this module is the
CPU BASELINE that bench.py times with all host cores (numpy/scipy -> OpenBLAS threads).

TEST INFRASTRUCTURE ONLY (same rule as oracle.py): imported by tests/ and by bench.py's
cpu_baseline / --impl reference legs, never by the product package.  Parity status: "parity
unpinned" (no CHOLMOD, no Julia here); tests/test_oracle.py pins it against kkt_oracle.c.

What follows which reference lines:
  delta_loop   ipopt_strategy!                      IPM/delta_strategy.jl:37-114, parameters.jl:147-158
  factorize    update_delta_vecs! + ls_factor!      schur.jl:64-87, julia.jl:28-46 (1 = PD, 0 = PosDefException)
  solve        ls_solve                             julia.jl:99-113
  direction    compute_direction_implementation!    schur.jl:89-182, kkt_system_solver.jl:27-96

The elimination-tree structures (ordering, supernodes, row structures, child-to-parent maps)
come from the host-side symbolic analysis of the library under test (a host-only handle, no
device): they are index structures, the arithmetic below is independent of the CUDA code.
"""
import ctypes

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
from scipy.linalg import blas

from . import oracle as _orc

_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)

DELTA_PARS = dict(delta_zero=0.0, delta_min=1e-12, delta_max=1e50, delta_start=1e-6, inc=8.0, dec=1.0 / np.pi)


class SupernodalFactor:
    def __init__(self, QL, handle):
        """QL: scipy CSC, lower triangle with the full diagonal, sorted indices.
        handle: host-only opb handle whose structure was set from the same (J, H) patterns."""
        g = handle.symbolic
        QL = sp.csc_matrix(QL)
        self.n = QL.shape[0]
        Mp, Mi = g("Mp"), g("Mi")
        if not (np.array_equal(Mp, QL.indptr) and np.array_equal(Mi, QL.indices)):
            raise ValueError("QL does not have the pattern the symbolic analysis was made for")
        self.perm = g("perm"); self.sfirst = g("sfirst"); self.sparent = g("sparent")
        self.rowptr = g("rowptr"); self.rowidx = g("rowidx"); self.rel = g("rel")
        self.Loff = g("Loff"); self.amap = g("amap"); self.dpos = g("dpos")
        self.nsuper = len(self.sfirst) - 1
        self.nnzL = int(handle.info("nnzL"))
        self.children = [[] for _ in range(self.nsuper)]
        for s in range(self.nsuper):
            if self.sparent[s] >= 0:
                self.children[self.sparent[s]].append(s)
        self.L = None
        self.delta = 0.0
        # index arrays of the C solve (kept alive here)
        cp = np.zeros(self.nsuper + 1, np.int64)
        for s in range(self.nsuper):
            cp[s + 1] = cp[s] + len(self.children[s])
        cl = np.array([ch for s in range(self.nsuper) for ch in self.children[s]] or [0], np.int64)
        self._keep = [np.ascontiguousarray(v, dtype=np.int64) for v in
                      (self.sfirst, self.rowptr, self.rowidx if len(self.rowidx) else np.zeros(1), self.rel if len(self.rel) else np.zeros(1),
                       self.Loff, cp, cl)]
        self._solve_args = [v.ctypes.data_as(_i64p) for v in self._keep]

    # -- update_delta_vecs! + ls_factor!(:definite)
    def factorize(self, nzval, delta):
        """Returns 1 when Q + delta*I is positive definite (every pivot > 0), else 0."""
        L = np.zeros(self.nnzL)
        L[self.amap] = nzval
        L[self.dpos] += delta                 # Q[i,i] = schur_diag[i] + delta (schur.jl:76)
        CB = {}
        sfirst, rowptr, Loff = self.sfirst, self.rowptr, self.Loff
        rel = np.ascontiguousarray(self.rel, dtype=np.int64)
        relp = rel.ctypes.data
        extend_add = _orc.lib().orc_extend_add
        for s in range(self.nsuper):          # supernodes are numbered in postorder
            c = int(sfirst[s + 1] - sfirst[s])
            r = int(rowptr[s + 1] - rowptr[s])
            N = c + r
            ld = (N + 1) & ~1
            P = L[Loff[s]:Loff[s] + ld * c]                           # the panel in place: N x c, leading dimension ld
            U = np.zeros((r, r), order="F")
            for ch in self.children[s]:
                cb = CB.pop(ch)
                rc = cb.shape[0]
                extend_add(P.ctypes.data_as(_f64p), ld, c, U.ctypes.data_as(_f64p), r,
                           ctypes.cast(relp + 8 * int(rowptr[ch]), _i64p), rc, cb.ctypes.data_as(_f64p))
            view = P.reshape(c, ld).T                                   # (ld x c) column-major view
            L11, info = sla.lapack.dpotrf(view[:c], lower=1, clean=1, overwrite_a=0)
            if info != 0 or not np.isfinite(L11[np.diag_indices(c)]).all():
                self.L = None                                           # pivot <= 0 or NaN: not positive definite
                return 0
            view[:c] = L11
            if r:
                # L21 = A21 * L11^-T (dtrsm), update block U -= L21 L21' (dsyrk, lower part)
                L21 = blas.dtrsm(1.0, L11, np.asfortranarray(view[c:N]), side=1, lower=1, trans_a=1)
                view[c:N] = L21
                U = blas.dsyrk(-1.0, L21, beta=1.0, c=U, lower=1, overwrite_c=1)
                CB[s] = U
        self.L = L
        self.delta = delta
        return 1

    # -- ls_solve
    def solve(self, b):
        x = np.ascontiguousarray(np.asarray(b, dtype=np.float64)[self.perm])
        u = np.empty(max(int(self.rowptr[-1]), 1))
        a = self._solve_args
        _orc.lib().orc_snode_solve(self.nsuper, *a, self.L.ctypes.data_as(_f64p), x.ctypes.data_as(_f64p),
                                   u.ctypes.data_as(_f64p))
        out = np.empty(self.n)
        out[self.perm] = x
        return out

    # -- ipopt_strategy!
    def delta_loop(self, nzval, schur_diag, delta_prev, **kw):
        p = dict(DELTA_PARS); p.update(kw)
        tried = []
        num_fac = 0
        tau = 1.5 * float(np.min(schur_diag))
        delta = p["delta_zero"]
        if tau > 0.0:
            tau = 0.0
            ok = self.factorize(nzval, delta); num_fac += 1; tried.append(delta)
            if ok == 1:
                return "success", num_fac, delta, np.array(tried)
        for i in range(1, 501):
            if i == 1:
                delta = max(p["delta_min"] - tau, delta_prev * p["dec"]) if delta_prev != 0.0 else p["delta_start"] - tau
            else:
                delta = delta * p["inc"]
            ok = self.factorize(nzval, delta); num_fac += 1; tried.append(delta)
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
