"""Host-side mirror of the reference's two plugin interfaces for the KKT path,
bound to libonephase_b200.so.  Julia is not available in this image, so this
Python mirror is what the parity tests drive; the Julia shim with the same
structure is in julia/ (INTEGRATION.md).  Names, argument meaning and error
behaviour follow the reference (Julia's `!` suffix is dropped):

  linear_solver_B200      <: abstract_linear_system_solver
      initialize / finalize / ls_factor / ls_solve_inplace (ls_solve!) / ls_solve
      (src/linear_system_solvers/linear_system_solvers.jl:11,40-46; julia.jl:1-113)
  Schur_B200_KKT_solver   <: abstract_schur_solver
      form_system / update_delta_vecs / factor_implementation /
      compute_direction_implementation / kkt_associate_rhs, plus the generic
      factor / compute_direction / update_delta / diag_min
      (src/kkt_system_solver/kkt_system_solver.jl:10-25,98-113,178-204,291-294;
       schur.jl:3-182)
  ipopt_strategy          (src/IPM/delta_strategy.jl:37-114), specialised on the
      B200 solver so the whole delta loop runs on the device
  pick_KKT_solver         (kkt_system_solver.jl:232-287) with the new symbols
      kkt_solver_type = :schur_b200, linear_solver_type = :b200
"""
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from . import _lib


# ---------------------------------------------------------------------------
# parameter mirror (src/parameters.jl:17-45,147-158)
# ---------------------------------------------------------------------------
@dataclass
class Class_delta_pars:
    max: float = 1e50
    start: float = 1e-6
    zero: float = 0.0
    min: float = 1e-12
    inc: float = 8.0
    dec: float = 1.0 / np.pi


@dataclass
class Class_kkt_pars:
    ItRefine_Num: int = 3
    ItRefine_BigFloat: bool = False
    kkt_solver_type: str = "schur_b200"
    linear_solver_type: str = "b200"
    linear_solver_safe_mode: bool = False
    linear_solver_recycle: bool = False


@dataclass
class Class_parameters:
    output_level: int = 0
    kkt: Class_kkt_pars = field(default_factory=Class_kkt_pars)
    delta: Class_delta_pars = field(default_factory=Class_delta_pars)
    device: int = 0


@dataclass
class Class_point:
    """utils/Class_point.jl:2-13 (only the fields the path touches)."""
    x: np.ndarray
    y: np.ndarray
    s: np.ndarray
    mu: float = 0.0
    primal_scale: float = 1.0


@dataclass
class Class_iterate:
    """The slice of Class_iterate (utils/Class_iterate.jl:4-84) the path reads:
    cached J (m x n CSC), H (lower CSC), point.y, point.s and local_info.delta."""
    J: sp.csc_matrix
    H: sp.csc_matrix
    y: np.ndarray
    s: np.ndarray
    delta: float = 0.0     # get_delta(iter)
    mu: float = 0.0
    primal_scale: float = 1.0


@dataclass
class System_rhs:
    """system_rhs.jl:39-74."""
    dual_r: np.ndarray
    primal_r: np.ndarray
    comp_r: np.ndarray


@dataclass
class Class_kkt_error:
    """kkt_system_solver.jl:49-65."""
    error_D: float = 0.0
    error_P: float = 0.0
    error_mu: float = 0.0
    overall: float = 0.0
    rhs_norm: float = 0.0
    ratio: float = 0.0


def dim(it):
    return it.J.shape[1]


def ncon(it):
    return it.J.shape[0]


def get_delta(it):
    return it.delta


def set_delta(it, d):
    it.delta = d


def _csc(A):
    if sp.isspmatrix_csc(A) and A.has_sorted_indices:
        return A              # the usual case: no new object, the sortedness flag stays cached on A
    A = sp.csc_matrix(A)
    if not A.has_sorted_indices:
        A = A.copy(); A.sort_indices()
    return A


# ---------------------------------------------------------------------------
# L1: linear system solver
# ---------------------------------------------------------------------------
class linear_solver_B200:
    """Drop-in for linear_solver_JULIA (julia.jl:1-19)."""

    def __init__(self, sym, safe_mode=False, recycle=False, device=0):
        if sym not in ("definite", "symmetric"):
            # julia.jl:92: error("this.options.sym = ... not supported"); :unsymmetric is dead upstream
            raise ValueError("this.options.sym = %s not supported" % sym)
        self.sym = sym
        self.safe_mode = safe_mode
        self.recycle = recycle   # symbolic analysis is always reused (pattern-hash cache)
        self.device = device
        self._h = None
        self._factor_defined = False

    def initialize(self):
        if self._h is None:
            self._h = _lib.Handle(self.device)

    def finalize(self):
        if self._h is not None:
            self._h.close(); self._h = None
        self._factor_defined = False

    def ls_factor(self, SparseMatrix, n, m, timer=None):
        """julia.jl:21-97.  Returns 1 when the inertia is (n, m), else 0."""
        self.initialize()
        Q = _csc(SparseMatrix)
        if self.sym == "definite":
            assert m == 0
            mode = _lib.MODE_CHOLESKY
        else:
            mode = _lib.MODE_LDLT
        ok = self._h.ls_factor_csc(Q.shape[0], Q.indptr, Q.indices, Q.data, 0, mode, n, m)
        self._factor_defined = True
        return ok

    def ls_solve_inplace(self, my_rhs, my_sol, timer=None):
        """ls_solve!  (julia.jl:99-103)."""
        self._h.ls_solve(np.asarray(my_rhs, dtype=np.float64), out=my_sol)

    def ls_solve(self, my_rhs, timer=None):
        """julia.jl:105-113; a sparse rhs is densified like Vector(my_rhs)."""
        if sp.issparse(my_rhs):
            my_rhs = np.asarray(my_rhs.todense()).ravel()
        return self._h.ls_solve(np.asarray(my_rhs, dtype=np.float64))


# ---------------------------------------------------------------------------
# L2: KKT system solver
# ---------------------------------------------------------------------------
class Schur_B200_KKT_solver:
    """Drop-in for Schur_KKT_solver (schur.jl:3-31).  The matrix Q, its factor and
    the cached (J, H, y, s) of factor_it live in HBM behind one opb handle."""

    def __init__(self, device=0, shard=None):
        """shard: None, or an object with `rank`, `world` and `exchange(bytes) -> list[bytes]`
        (ThreadShard / DistShard below): ONE instance is then factorised and solved by `world`
        GPUs, every rank calling the same methods with the same data."""
        self.shard = shard
        self.ls_solver = None
        self.factor_it = None
        self.delta_x_vec = None
        self.delta_s_vec = None
        self.rhs = None
        self.dir = None
        self.kkt_err_norm = Class_kkt_error()
        self.rhs_norm = 0.0
        self.pars = None
        self.schur_diag = None
        self.ready = "not_ready"
        self.Q = None            # host copy of tril(Q), materialised on demand (is_diag_dom)
        self.current_it = None
        self.reduct_factors = None
        self.device = device
        self._h = None
        self._delta = 0.0
        self._diag_min = np.nan
        self._pattern_key = None
        self._pattern_ident = None
        self._pattern_refs = None

    # -- initialize!(kkt_solver, it)  kkt_system_solver.jl:21-25
    def initialize(self, initial_it):
        if self._h is None:
            self._h = _lib.Handle(self.device)
            if self.shard is not None and self.shard.world > 1:
                self._h.shard_init(self.shard.rank, self.shard.world)
        self.dir = Class_point(np.zeros(dim(initial_it)), np.zeros(ncon(initial_it)), np.zeros(ncon(initial_it)))

    def finalize(self):
        if self._h is not None:
            self._h.close(); self._h = None

    def set_permutation(self, perm):
        self._h.set_permutation(perm)
        self._pattern_key = None
        self._pattern_ident = None

    # -- form_system!  schur.jl:47-62
    @staticmethod
    def _pattern_sample(J, H):
        """Cheap content fingerprint of the index arrays (a strided sample of <= 4096 entries of
        each plus the ends): catches a caller that refills the SAME preallocated index arrays in
        place with another pattern of equal nnz, which the id()-based shortcut alone would miss."""
        out = []
        for a in (J.indptr, J.indices, H.indptr, H.indices):
            step = max(1, a.shape[0] // 4096)
            out.append(hash(a[::step].tobytes()))
            out.append(int(a[-1]) if a.shape[0] else 0)
        return tuple(out)

    def form_system(self, it, timer=None):
        J = _csc(it.J); H = _csc(it.H)
        # pattern identity: same index arrays as last time (the usual case: the iterate's cached
        # matrices keep their structure) and the same sampled content -> no full hashing; the full
        # pattern hash is re-verified on every 16th call and whenever the shortcut does not apply
        ident = (id(J.indptr), id(J.indices), id(H.indptr), id(H.indices), J.shape, J.nnz, H.nnz,
                 self._pattern_sample(J, H))
        self._form_calls = getattr(self, "_form_calls", 0) + 1
        if ident != self._pattern_ident or self._form_calls % 16 == 0:
            key = (J.shape, hash(J.indptr.tobytes()), hash(J.indices.tobytes()),
                   hash(H.indptr.tobytes()), hash(H.indices.tobytes()))
            if key != self._pattern_key:
                # pattern changed (Class_cutest.jl:490-502 can drop numerical zeros): new symbolic analysis
                self._h.set_structure(J.shape[1], J.shape[0], J.indptr, J.indices, H.indptr, H.indices, 0)
                self._attach_peers()
                self._pattern_key = key
            self._pattern_ident = ident
            self._pattern_refs = (J.indptr, J.indices, H.indptr, H.indices)   # keep the ids alive
        # schur_diag is refilled in place (same array object across iterations, like the reference's
        # kkt_solver.schur_diag field) unless the dimension changed
        reuse = self.schur_diag if (self.schur_diag is not None and self.schur_diag.shape == (J.shape[1],)) else None
        self.schur_diag, self._diag_min = self._h.form(J.data, H.data, it.y, it.s, out=reuse)
        self.factor_it = it
        self.Q = None
        self.ready = "system_formed"

    def _attach_peers(self):
        """Sharded instance: after every symbolic analysis the ranks exchange the CUDA IPC
        descriptors of their peer-visible buffers and map each other's memory."""
        if self.shard is None or self.shard.world <= 1:
            return
        import hashlib
        # every rank must have derived the same ordering and the same supernode-to-rank map
        sig = hashlib.md5(self._h.symbolic("perm").tobytes() + self._h.symbolic("owner").tobytes()).digest()
        blobs = self.shard.exchange(self._h.shard_export() + sig)
        for p, b in enumerate(blobs):
            if b[_lib.SHARD_BLOB_BYTES:] != sig:
                raise RuntimeError("sharded instance: rank %d derived a different symbolic analysis" % p)
            if p != self.shard.rank:
                self._h.shard_attach(p, b[:_lib.SHARD_BLOB_BYTES])
        self.shard.exchange(b"ok")            # nobody launches before every rank has attached

    def get_Q(self):
        """Lower triangle of Q with the current shift, as scipy CSC (for is_diag_dom)."""
        cp, ri = self._h.M_pattern()
        v = self._h.M_values()
        Q = sp.csc_matrix((v, ri, cp), shape=(self._h.n, self._h.n))
        Q.setdiag(self.schur_diag + (self._delta if self.delta_x_vec is not None else 0.0))
        return Q

    # -- update_delta!/update_delta_vecs!  kkt_system_solver.jl:109-113, schur.jl:64-83
    def update_delta(self, delta_x, delta_s, timer=None):
        n = dim(self.factor_it)
        self.update_delta_vecs(delta_x * np.ones(n), delta_s * self.factor_it.s ** (-2.0), timer)

    def update_delta_vecs(self, delta_x_vec, delta_s_vec, timer=None):
        self.delta_x_vec = delta_x_vec
        self.delta_s_vec = delta_s_vec
        if np.sum(np.abs(delta_s_vec)) > 0.0:
            raise RuntimeError("Not implemented")          # schur.jl:71
        if delta_x_vec.size and np.any(delta_x_vec != delta_x_vec[0]):
            raise RuntimeError("Not implemented: non-uniform delta_x_vec")
        self._delta = float(delta_x_vec[0]) if delta_x_vec.size else 0.0
        self.ready = "delta_updated"

    # -- factor!  kkt_system_solver.jl:98-107,190-204
    def factor(self, delta_x, timer=None, delta_s=0.0):
        self.update_delta(delta_x, delta_s, timer)
        return self._factor(timer)

    def _factor(self, timer=None):
        if self.ready != "delta_updated":
            raise RuntimeError("kkt solver not ready to factor kkt_solver.ready = %s != :delta_updated" % self.ready)
        self.ready = "factored"
        return self.factor_implementation(timer)

    def factor_implementation(self, timer=None):
        return self._h.factor(self._delta)

    # -- kkt_associate_rhs!  schur.jl:34-45
    def kkt_associate_rhs(self, it, rhs, reduct_factors=None, timer=None):
        """The reference builds System_rhs(iter, reduct_factors) from the NLP
        (system_rhs.jl:57-73); the NLP lives outside the path, so the caller
        passes the three vectors."""
        self.rhs = rhs
        if reduct_factors is not None:
            self.dir.mu = -(1.0 - reduct_factors[2]) * it.mu
            self.dir.primal_scale = -(1.0 - reduct_factors[0]) * it.primal_scale
        self.reduct_factors = reduct_factors
        self.current_it = it

    # -- compute_direction!  kkt_system_solver.jl:178-188
    def compute_direction(self, timer=None):
        if self.ready != "factored":
            raise RuntimeError("kkt solver not ready to compute direction!")
        self.compute_direction_implementation(timer)
        for v in (self.dir.x, self.dir.y, self.dir.s):      # check_for_nan, IPM_tools.jl:32-49
            # a NaN anywhere makes the sum NaN: one pass without a temporary in the usual case
            if np.isnan(np.sum(v)) and np.isnan(v).any():
                raise FloatingPointError("NaN in direction")

    def compute_direction_implementation(self, timer=None):
        n_ref = self.pars.kkt.ItRefine_Num if self.pars is not None else 3
        # kkt_solver.dir.x/y/s are filled in place, as in the reference (schur.jl:115-123)
        n, m = self._h.n, self._h.m
        d = self.dir
        ok = all(isinstance(v, np.ndarray) and v.dtype == np.float64 and v.flags.c_contiguous and v.shape == (k,)
                 for v, k in ((d.x, n), (d.y, m), (d.s, m)))
        out = (d.x, d.y, d.s) if ok else None
        dx, dy, ds, err = self._h.direction(self.rhs.dual_r, self.rhs.primal_r, self.rhs.comp_r, n_ref, out=out)
        self.dir.x, self.dir.y, self.dir.s = dx, dy, ds
        self.kkt_err_norm = Class_kkt_error(*[float(v) for v in err])
        self.rhs_norm = float(err[4])

    def diag_min(self):
        return self._diag_min


def diag_min(kkt_solver):
    """kkt_system_solver.jl:291-294."""
    return kkt_solver.diag_min()


# ---------------------------------------------------------------------------
# helpers of kkt_system_solver.jl:296-300 / utils/eval.jl:89-100 / init/guess-vars.jl:128-169
# ---------------------------------------------------------------------------
def eval_diag_J_T_J(it, diag_vals, handle=None, device=0):
    """eval.jl:89-100: di[i] = sum_j J[j,i]^2 * diag_vals[j], on the device (column-parallel kernel,
    the reference's operation order: square first, rows ascending)."""
    J = _csc(it.J)
    h = handle or _scratch_handle(device)
    return h.diag_JtDJ(J.shape[1], J.shape[0], J.indptr, J.indices, J.data, diag_vals)


def compute_schur_diag(it, handle=None, device=0):
    """kkt_system_solver.jl:296-300: diag(H) + eval_diag_J_T_J(iter, y ./ s)."""
    H = _csc(it.H)
    return H.diagonal() + eval_diag_J_T_J(it, np.asarray(it.y, dtype=np.float64) / np.asarray(it.s, dtype=np.float64),
                                          handle, device)


_SCRATCH = {}


def _scratch_handle(device):
    if device not in _SCRATCH:
        _SCRATCH[device] = _lib.Handle(device)
    return _SCRATCH[device]


def estimate_y_tilde(J, g, pars=None, device=0):
    """init/guess-vars.jl:128-169 (the Cholesky branch): y = -J * (cholesky(lambda I + J'J) solve of -g), lambda = 1e-4.
    The matrix is the primal Schur complement with y./s = 1 and H = lambda I, so it is assembled,
    factorised and solved by the same device path (opb_form / opb_factor / opb_ls_solve).  Like
    the reference, any failure returns ones(m)."""
    J = _csc(J)
    m, n = J.shape
    lam = 1e-4
    try:
        h = _lib.Handle(pars.device if pars is not None else device)
        try:
            eye = sp.identity(n, format="csc")
            h.set_structure(n, m, J.indptr, J.indices, eye.indptr, eye.indices, 0)
            h.form(J.data, np.full(n, lam), np.ones(m), np.ones(m), want_diag=False)
            if h.factor(0.0) != 1:
                raise RuntimeError("lambda I + J'J not positive definite")      # PosDefException in the reference
            dx = h.ls_solve(-np.asarray(g, dtype=np.float64))
        finally:
            h.close()
        return -(J @ dx)
    except Exception as e:                  # guess-vars.jl:163-167
        print("error in estimate_y_tilde")
        print(repr(e))
        return np.ones(m)


# ---------------------------------------------------------------------------
# L2: Symmetric_KKT_solver on the LDL' back-end (kkt_system_solver/symmetric.jl)
# ---------------------------------------------------------------------------
class Symmetric_B200_KKT_solver:
    """Drop-in for Symmetric_KKT_solver (symmetric.jl:2-33): the quasi-definite system
        [[H + delta I, J'], [J, -S/Y]]
    factorised by the device LDL' (linear_solver_B200(:symmetric)) with the inertia test
    pos == n, neg == m (julia.jl:72-80, linear_system_solvers.jl:48-91).  The lower triangle of
    the matrix keeps one pattern across iterations, so its symbolic analysis is cached by the
    library; only the values are refilled here."""

    def __init__(self, device=0):
        self.ls_solver = None
        self.factor_it = None
        self.delta_x_vec = None
        self.delta_s_vec = None
        self.rhs = None
        self.dir = None
        self.kkt_err_norm = Class_kkt_error()
        self.rhs_norm = 0.0
        self.pars = None
        self.schur_diag = None
        self.true_x_diag = None
        self.ready = "not_ready"
        self.Q = None
        self.device = device
        self._pat = None

    def initialize(self, initial_it):
        if self.ls_solver is None:
            self.ls_solver = linear_solver_B200("symmetric", False, False, self.device)
        self.ls_solver.initialize()
        self.dir = Class_point(np.zeros(dim(initial_it)), np.zeros(ncon(initial_it)), np.zeros(ncon(initial_it)))

    def finalize(self):
        if self.ls_solver is not None:
            self.ls_solver.finalize()

    # -- form_system!  symmetric.jl:35-54
    def form_system(self, it, timer=None):
        J = _csc(it.J); H = _csc(it.H)
        m, n = J.shape
        key = (J.shape, J.nnz, H.nnz, id(J.indptr), id(J.indices), id(H.indptr), id(H.indices),
               Schur_B200_KKT_solver._pattern_sample(J, H))
        if self._pat is None or self._pat[0] != key:
            # pattern of the lower triangle [[tril(H) + full diagonal, 0], [J, diag]] and where the
            # values of H, J and -s./y go in its nzval (symmetric.jl:41: M = [[H J_T]; [J B]])
            one = lambda A: sp.csc_matrix((np.ones(A.nnz), A.indices, A.indptr), shape=A.shape)      # noqa: E731
            P = sp.bmat([[one(H) + sp.identity(n, format="csc"), None], [one(J), sp.identity(m, format="csc")]], format="csc")
            P.sort_indices()
            cols = np.repeat(np.arange(n + m), np.diff(P.indptr))
            lookup = {}
            pos = np.arange(P.nnz)
            order = np.lexsort((P.indices, cols))
            assert np.array_equal(order, pos)
            keyarr = cols.astype(np.int64) * (n + m) + P.indices
            def locate(rows, cls):
                want = cls.astype(np.int64) * (n + m) + rows
                idx = np.searchsorted(keyarr, want)
                assert np.array_equal(keyarr[idx], want)
                return idx
            hcols = np.repeat(np.arange(n), np.diff(H.indptr))
            jcols = np.repeat(np.arange(n), np.diff(J.indptr))
            self._pat = (key, P.indptr.astype(np.int64), P.indices.astype(np.int64),
                         locate(H.indices, hcols), locate(J.indices + n, jcols),
                         locate(np.arange(n, n + m), np.arange(n, n + m)), locate(np.arange(n), np.arange(n)))
        _, cp, ri, hpos, jpos, bpos, xdiag = self._pat
        v = np.zeros(ri.shape[0])
        np.add.at(v, hpos, H.data)
        v[jpos] = J.data
        v[bpos] = -np.asarray(it.s, dtype=np.float64) / np.asarray(it.y, dtype=np.float64)
        self.Q = sp.csc_matrix((v, ri, cp), shape=(n + m, n + m))
        self._xdiag = xdiag
        self.factor_it = it
        self.schur_diag = compute_schur_diag(it, self.ls_solver._h)
        self.true_x_diag = v[xdiag].copy()
        self.ready = "system_formed"

    # -- update_delta_vecs!  symmetric.jl:85-102
    def update_delta(self, delta_x, delta_s, timer=None):
        n = dim(self.factor_it)
        self.update_delta_vecs(delta_x * np.ones(n), delta_s * self.factor_it.s ** (-2.0), timer)

    def update_delta_vecs(self, delta_x_vec, delta_s_vec, timer=None):
        self.delta_x_vec = delta_x_vec
        self.delta_s_vec = delta_s_vec
        if np.sum(np.abs(delta_s_vec)) > 0.0:
            raise RuntimeError("not implemented")          # symmetric.jl:94
        self.Q.data[self._xdiag] = self.true_x_diag + delta_x_vec
        self.ready = "delta_updated"

    def factor(self, delta_x, timer=None, delta_s=0.0):
        self.update_delta(delta_x, delta_s, timer)
        if self.ready != "delta_updated":
            raise RuntimeError("kkt solver not ready to factor kkt_solver.ready = %s != :delta_updated" % self.ready)
        self.ready = "factored"
        return self.factor_implementation(timer)

    # -- factor_implementation!  symmetric.jl:56-58
    def factor_implementation(self, timer=None):
        return self.ls_solver.ls_factor(self.Q, dim(self.factor_it), ncon(self.factor_it), timer)

    def kkt_associate_rhs(self, it, rhs, reduct_factors=None, timer=None):
        self.rhs = rhs
        if reduct_factors is not None:
            self.dir.mu = -(1.0 - reduct_factors[2]) * it.mu
            self.dir.primal_scale = -(1.0 - reduct_factors[0]) * it.primal_scale

    def compute_direction(self, timer=None):
        if self.ready != "factored":
            raise RuntimeError("kkt solver not ready to compute direction!")
        self.compute_direction_implementation(timer)
        for v in (self.dir.x, self.dir.y, self.dir.s):
            if np.isnan(np.sum(v)) and np.isnan(v).any():
                raise FloatingPointError("NaN in direction")

    # -- compute_direction_implementation!  symmetric.jl:60-85
    def compute_direction_implementation(self, timer=None):
        it = self.factor_it
        y_org = np.asarray(it.y, dtype=np.float64)
        rhs = self.rhs
        n = dim(it)
        symmetric_rhs = np.concatenate([rhs.dual_r, rhs.primal_r + rhs.comp_r / y_org])
        sol = self.ls_solver.ls_solve(symmetric_rhs, timer)
        d = self.dir
        d.x = sol[:n].copy()
        d.y = -sol[n:]
        J = _csc(it.J)
        d.s = J @ d.x - rhs.primal_r                         # eval_jac_prod(factor_it, dir.x) - rhs.primal_r
        self.update_kkt_error()

    def update_kkt_error(self):
        """update_kkt_error! (kkt_system_solver.jl:67-96): generic host code of the reference,
        evaluated from the factorisation iterate's matrices."""
        it, d, rhs = self.factor_it, self.dir, self.rhs
        J = _csc(it.J); H = _csc(it.H)
        Hs = H + sp.tril(H, -1).T
        eD = (self.delta_x_vec * d.x + Hs @ d.x - J.T @ d.y) - rhs.dual_r
        eP = J @ d.x - d.s - rhs.primal_r
        eM = np.asarray(it.s) * d.y + np.asarray(it.y) * d.s - rhs.comp_r
        nrm = lambda v: float(np.abs(v).max()) if v.size else 0.0      # noqa: E731
        overall = max(nrm(eD), nrm(eP), nrm(eM))
        rn = max(nrm(rhs.dual_r), nrm(rhs.primal_r), nrm(rhs.comp_r))
        with np.errstate(divide="ignore", invalid="ignore"):
            self.kkt_err_norm = Class_kkt_error(nrm(eD), nrm(eP), nrm(eM), overall, rn, float(np.float64(overall) / np.float64(rn)))
        self.rhs_norm = rn

    def diag_min(self):
        return float(np.min(self.schur_diag))


def ipopt_strategy(it, kkt_solver, pars, timer=None):
    """delta_strategy.jl:37-114 for the B200 solver: one library call runs the
    probe at delta.zero, the first shift and the x8 retries on the device.
    Returns (status, num_fac, delta) with status 'success' or 'failure'."""
    if kkt_solver.ready not in ("system_formed", "delta_updated", "factored"):
        raise RuntimeError("kkt solver not ready to factor")
    d = pars.delta
    st, num_fac, delta = kkt_solver._h.factor_delta_loop(get_delta(it), d.zero, d.min, d.max, d.start,
                                                         d.inc, d.dec, 500)
    n = dim(it)
    # delta_x_vec = delta * ones(n), delta_s_vec = zeros(m) (delta_strategy.jl via update_delta!):
    # refilled in place when the sizes are unchanged
    dxv, dsv = kkt_solver.delta_x_vec, kkt_solver.delta_s_vec
    if isinstance(dxv, np.ndarray) and dxv.shape == (n,) and isinstance(dsv, np.ndarray) and dsv.shape == (ncon(it),):
        dxv.fill(delta); dsv.fill(0.0)
    else:
        kkt_solver.delta_x_vec = np.full(n, delta)
        kkt_solver.delta_s_vec = np.zeros(ncon(it))
    kkt_solver._delta = delta
    kkt_solver.ready = "factored"
    if st == 1:
        return "success", num_fac, delta
    if st == 0:
        return "failure", num_fac, delta
    raise RuntimeError("max it")                         # delta_strategy.jl:113


class DistShard:
    """Rank bookkeeping of a sharded instance over torch.distributed (one process per GPU).
    Only the IPC descriptors travel through the process group; the numeric data moves between
    the GPUs inside the kernels (peer-mapped HBM over NVLink)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def exchange(self, blob):
        out = [None] * self.world
        self._dist.all_gather_object(out, blob, group=self.group)
        return out


class ThreadShard:
    """The same bookkeeping for `world` host threads of ONE process, each driving its own handle
    (used by the single-GPU tests: several virtual ranks on one device)."""

    class _Hub:
        def __init__(self, world):
            import threading
            self.world, self.slots, self.barrier = world, [None] * world, threading.Barrier(world)

    def __init__(self, hub, rank):
        self.hub, self.rank, self.world = hub, rank, hub.world

    @classmethod
    def make(cls, world):
        hub = cls._Hub(world)
        return [cls(hub, r) for r in range(world)]

    def exchange(self, blob):
        self.hub.slots[self.rank] = blob
        self.hub.barrier.wait()
        out = list(self.hub.slots)
        self.hub.barrier.wait()
        return out


def respond_to_failed_step(it, kkt_solver, pars, old_delta, grad_lag_inf=None, response="lag_delta_inc"):
    """The failure-driven delta increase of the outer loop (IPM/one_phase.jl:231-242): after a
    failed line search the caller raises delta and refactorises ONCE (the inertia flag is ignored
    there, `tot_num_fac += 1`).  `grad_lag_inf` = norm(eval_grad_lag(iter, mu), Inf), which lives
    outside the path (the NLP); `old_delta` = get_delta(iter) before this outer iteration.
    Returns (new_delta, inertia)."""
    d = pars.delta
    floor = max(d.start, old_delta * d.dec)
    if response == "lag_delta_inc":
        if grad_lag_inf is None:
            raise ValueError("response_to_failure = :lag_delta_inc needs norm(grad L, Inf)")
        # Julia semantics (one_phase.jl:233): a zero direction gives Inf (or NaN for 0/0), no exception,
        # and max() propagates NaN
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = np.float64(grad_lag_inf) / np.float64(np.abs(kkt_solver.dir.x).max())
        new_delta = float(np.maximum(np.maximum(ratio, get_delta(it) * d.inc), floor))
    elif response == "default":
        new_delta = max(get_delta(it) * d.inc, floor)
    else:
        raise ValueError("pars.test.response_to_failure parameter incorrectly set")      # one_phase.jl:238
    set_delta(it, new_delta)
    inertia = kkt_solver.factor(new_delta)       # factor!(kkt_solver, get_delta(iter), timer), one_phase.jl:241
    return new_delta, inertia


def pick_KKT_solver(pars, shard=None):
    """kkt_system_solver.jl:232-287 extended with the B200 symbols."""
    t, ls = pars.kkt.kkt_solver_type, pars.kkt.linear_solver_type
    if t == "schur_b200":
        if ls != "b200":
            raise ValueError("pick a valid solver!")
        k = Schur_B200_KKT_solver(pars.device, shard)
        k.ls_solver = linear_solver_B200("definite", pars.kkt.linear_solver_safe_mode,
                                         pars.kkt.linear_solver_recycle, pars.device)
    elif t == "symmetric_b200":
        if ls != "b200":
            raise ValueError("pick a valid solver!")
        k = Symmetric_B200_KKT_solver(pars.device)
        k.ls_solver = linear_solver_B200("symmetric", pars.kkt.linear_solver_safe_mode,
                                         pars.kkt.linear_solver_recycle, pars.device)
    else:
        raise ValueError("pick a solver!")
    k.kkt_err_norm = Class_kkt_error()
    k.pars = pars
    return k
