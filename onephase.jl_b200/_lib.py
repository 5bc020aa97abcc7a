"""ctypes binding of libonephase_b200.so (include/onephase_b200.h).

The shared library is the product; this module only loads it.  There is no
Python or CPU implementation behind these calls: if the library is missing or
no CUDA device is usable the calls raise.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libonephase_b200.so")
_lib = None

c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_intp = ctypes.POINTER(ctypes.c_int)

OPB_OK = 0
OPB_ERR_INVALID, OPB_ERR_STATE, OPB_ERR_CUDA, OPB_ERR_NO_DEVICE, OPB_ERR_INTERNAL = -1, -2, -3, -4, -5
MODE_CHOLESKY, MODE_LDLT = 0, 1

# every symbol include/onephase_b200.h declares
EXPORTS = [
    "opb_create", "opb_destroy", "opb_last_error", "opb_set_stream", "opb_set_option",
    "opb_set_permutation", "opb_set_structure", "opb_form", "opb_get_M_pattern", "opb_get_M_values",
    "opb_factor_delta_loop", "opb_factor", "opb_direction", "opb_ls_factor_csc", "opb_ls_solve",
    "opb_upload_values", "opb_upload_rhs", "opb_form_resident", "opb_delta_loop_resident",
    "opb_direction_resident", "opb_solve_resident", "opb_sync_state", "opb_get_info",
    "opb_get_symbolic", "opb_get_L_values", "opb_launch_count", "opb_version",
    "opb_shard_init", "opb_shard_export", "opb_shard_attach", "opb_profile_factor", "opb_eval_diag_JtDJ",
    "opb_system_rhs", "opb_step_bounds", "opb_get_direction", "opb_profile_levels", "opb_cache_clear",
]
SHARD_BLOB_BYTES = 384


class OPBError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libonephase_b200 error %d: %s" % (code, msg))
        self.code = code


def build(force=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8", "-s"]
    if force:
        args.append("-B")
    subprocess.check_call(args)
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libonephase_b200.so is not built (%s). Run `make -C onephase.jl_b200/csrc` or "
            "__graft_entry__.build(); there is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, f64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_int
    L.opb_create.argtypes = [ctypes.POINTER(vp), ci, ctypes.c_uint]
    L.opb_destroy.argtypes = [vp]
    L.opb_last_error.argtypes = [vp]; L.opb_last_error.restype = ctypes.c_char_p
    L.opb_set_stream.argtypes = [vp, vp]
    L.opb_set_option.argtypes = [vp, ctypes.c_char_p, f64]
    L.opb_set_permutation.argtypes = [vp, i64, c_i64p]
    L.opb_set_structure.argtypes = [vp, i64, i64, c_i64p, c_i64p, c_i64p, c_i64p, ci]
    L.opb_form.argtypes = [vp, c_f64p, c_f64p, c_f64p, c_f64p, c_f64p, c_f64p]
    L.opb_get_M_pattern.argtypes = [vp, c_i64p, c_i64p]
    L.opb_get_M_values.argtypes = [vp, c_f64p]
    L.opb_factor_delta_loop.argtypes = [vp] + [f64] * 7 + [ci, c_f64p, c_intp, c_intp]
    L.opb_factor.argtypes = [vp, f64, c_intp]
    L.opb_direction.argtypes = [vp, c_f64p, c_f64p, c_f64p, ci, c_f64p, c_f64p, c_f64p, c_f64p]
    L.opb_ls_factor_csc.argtypes = [vp, i64, c_i64p, c_i64p, c_f64p, ci, ci, i64, i64, c_intp]
    L.opb_ls_solve.argtypes = [vp, c_f64p, c_f64p]
    L.opb_upload_values.argtypes = [vp, c_f64p, c_f64p, c_f64p, c_f64p]
    L.opb_upload_rhs.argtypes = [vp, c_f64p, c_f64p, c_f64p]
    L.opb_form_resident.argtypes = [vp]
    L.opb_delta_loop_resident.argtypes = [vp] + [f64] * 7 + [ci]
    L.opb_direction_resident.argtypes = [vp, ci]
    L.opb_solve_resident.argtypes = [vp, ci]
    L.opb_sync_state.argtypes = [vp, c_f64p, c_intp, c_intp, c_f64p]
    L.opb_get_info.argtypes = [vp, ctypes.c_char_p, c_f64p]
    L.opb_get_symbolic.argtypes = [vp, ctypes.c_char_p, c_i64p, i64]
    L.opb_get_symbolic.restype = i64
    L.opb_get_L_values.argtypes = [vp, c_f64p, i64]
    L.opb_shard_init.argtypes = [vp, ci, ci]
    L.opb_shard_export.argtypes = [vp, ctypes.c_char_p]
    L.opb_shard_attach.argtypes = [vp, ci, ctypes.c_char_p]
    L.opb_profile_factor.argtypes = [vp, f64, c_f64p, c_f64p, c_f64p, c_f64p, c_f64p, c_intp]
    L.opb_profile_levels.argtypes = [vp, f64, c_f64p, ci, c_intp, c_f64p]
    L.opb_eval_diag_JtDJ.argtypes = [vp, i64, i64, c_i64p, c_i64p, c_f64p, ci, c_f64p, c_f64p]
    L.opb_system_rhs.argtypes = [vp, c_f64p, c_f64p] + [f64] * 5 + [c_f64p, c_f64p, c_f64p]
    L.opb_step_bounds.argtypes = [vp, f64, f64, c_f64p]
    L.opb_get_direction.argtypes = [vp, c_f64p, c_f64p, c_f64p, c_f64p]
    L.opb_launch_count.restype = ctypes.c_longlong
    L.opb_version.restype = ctypes.c_char_p
    _lib = L
    return L


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def pf(a):
    return a.ctypes.data_as(c_f64p) if a is not None else None


def pi(a):
    return a.ctypes.data_as(c_i64p) if a is not None else None


class Handle:
    """Thin RAII wrapper of an opb_handle."""

    def __init__(self, device=0):
        self.L = load()
        self.h = ctypes.c_void_p()
        rc = self.L.opb_create(ctypes.byref(self.h), int(device), 0)
        if rc != OPB_OK:
            msg = self.L.opb_last_error(self.h).decode() if self.h else "create failed"
            self.L.opb_destroy(self.h)
            self.h = None
            raise OPBError(rc, msg)
        self.device = device

    def check(self, rc):
        if rc != OPB_OK:
            raise OPBError(rc, self.L.opb_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.opb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- configuration
    def set_option(self, key, value):
        self.check(self.L.opb_set_option(self.h, key.encode(), float(value)))

    def set_stream(self, cuda_stream_ptr):
        self.check(self.L.opb_set_stream(self.h, ctypes.c_void_p(int(cuda_stream_ptr))))

    def set_permutation(self, perm):
        if perm is None:
            self.check(self.L.opb_set_permutation(self.h, 0, None))
        else:
            p = i64(perm)
            self.check(self.L.opb_set_permutation(self.h, p.shape[0], pi(p)))

    def set_structure(self, n, m, Jp, Ji, Hp, Hi, index_base=0):
        Jp, Ji, Hp, Hi = i64(Jp), i64(Ji), i64(Hp), i64(Hi)
        self.check(self.L.opb_set_structure(self.h, n, m, pi(Jp), pi(Ji), pi(Hp), pi(Hi), index_base))
        self.n, self.m = int(n), int(m)

    # --- one instance sharded over several GPUs (include/onephase_b200.h, opb_shard_*)
    def shard_init(self, rank, world):
        self.check(self.L.opb_shard_init(self.h, int(rank), int(world)))

    def shard_export(self):
        buf = ctypes.create_string_buffer(SHARD_BLOB_BYTES)
        self.check(self.L.opb_shard_export(self.h, buf))
        return buf.raw

    def shard_attach(self, peer, blob):
        assert len(blob) == SHARD_BLOB_BYTES
        self.check(self.L.opb_shard_attach(self.h, int(peer), blob))

    # --- numeric
    def form(self, Jx, Hx, y, s, want_diag=True, out=None):
        Jx, Hx, y, s = f64(Jx), f64(Hx), f64(y), f64(s)
        if out is not None:
            assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == (self.n,)
        sd = out if out is not None else (np.empty(self.n) if want_diag else None)
        dmin = ctypes.c_double()
        self.check(self.L.opb_form(self.h, pf(Jx), pf(Hx), pf(y), pf(s), pf(sd),
                                   ctypes.cast(ctypes.byref(dmin), c_f64p)))
        return sd, dmin.value

    def M_pattern(self):
        nnz = int(self.info("nnzM"))
        cp = np.empty(self.n + 1, np.int64); ri = np.empty(nnz, np.int64)
        self.check(self.L.opb_get_M_pattern(self.h, pi(cp), pi(ri)))
        return cp, ri

    def M_values(self):
        v = np.empty(int(self.info("nnzM")))
        self.check(self.L.opb_get_M_values(self.h, pf(v)))
        return v

    def factor_delta_loop(self, delta_prev, delta_zero=0.0, delta_min=1e-12, delta_max=1e50,
                          delta_start=1e-6, inc=8.0, dec=1.0 / np.pi, max_it=500):
        d = ctypes.c_double(); nf = ctypes.c_int(); st = ctypes.c_int()
        self.check(self.L.opb_factor_delta_loop(self.h, delta_prev, delta_zero, delta_min, delta_max,
                                                delta_start, inc, dec, max_it,
                                                ctypes.cast(ctypes.byref(d), c_f64p),
                                                ctypes.byref(nf), ctypes.byref(st)))
        return st.value, nf.value, d.value

    def factor(self, delta):
        ok = ctypes.c_int()
        self.check(self.L.opb_factor(self.h, float(delta), ctypes.byref(ok)))
        return ok.value

    def direction(self, dual_r, primal_r, comp_r, n_refine=3, out=None):
        """out = (dx, dy, ds): caller-owned float64 arrays written in place (the reference fills
        kkt_solver.dir.x/y/s in place, schur.jl:89-128); fresh arrays when omitted."""
        a, b, c = f64(dual_r), f64(primal_r), f64(comp_r)
        if out is not None:
            dx, dy, ds = out
            for v, k in ((dx, self.n), (dy, self.m), (ds, self.m)):
                assert v.dtype == np.float64 and v.flags.c_contiguous and v.shape == (k,)
        else:
            dx = np.empty(self.n); dy = np.empty(self.m); ds = np.empty(self.m)
        err = np.empty(6)
        self.check(self.L.opb_direction(self.h, pf(a), pf(b), pf(c), n_refine, pf(dx), pf(dy), pf(ds), pf(err)))
        return dx, dy, ds, err

    def ls_factor_csc(self, dim, colptr, rowval, nzval, index_base, mode, n_pos, m_neg):
        cp, ri, nz = i64(colptr), i64(rowval), f64(nzval)
        ok = ctypes.c_int()
        self.check(self.L.opb_ls_factor_csc(self.h, dim, pi(cp), pi(ri), pf(nz), index_base, mode,
                                            n_pos, m_neg, ctypes.byref(ok)))
        self.n = int(dim)
        return ok.value

    def ls_solve(self, rhs, out=None):
        r = f64(rhs)
        sol = out if out is not None else np.empty(self.n)
        assert sol.dtype == np.float64 and sol.flags.c_contiguous
        self.check(self.L.opb_ls_solve(self.h, pf(r), pf(sol)))
        return sol

    def diag_JtDJ(self, n, m, Jp, Ji, Jx, diag_vals, index_base=0):
        Jp, Ji, Jx, d = i64(Jp), i64(Ji), f64(Jx), f64(diag_vals)
        assert d.shape[0] == m
        out = np.empty(n)
        self.check(self.L.opb_eval_diag_JtDJ(self.h, n, m, pi(Jp), pi(Ji), pf(Jx), index_base, pf(d), pf(out)))
        return out

    # --- resident variants (bench)
    def upload_values(self, Jx, Hx, y, s):
        Jx, Hx, y, s = f64(Jx), f64(Hx), f64(y), f64(s)
        self.check(self.L.opb_upload_values(self.h, pf(Jx), pf(Hx), pf(y), pf(s)))
        self.sync_state()

    def upload_rhs(self, a, b, c):
        a, b, c = f64(a), f64(b), f64(c)
        self.check(self.L.opb_upload_rhs(self.h, pf(a), pf(b), pf(c)))
        self.sync_state()

    def form_resident(self):
        self.check(self.L.opb_form_resident(self.h))

    def delta_loop_resident(self, delta_prev, delta_zero=0.0, delta_min=1e-12, delta_max=1e50,
                            delta_start=1e-6, inc=8.0, dec=1.0 / np.pi, max_it=500):
        self.check(self.L.opb_delta_loop_resident(self.h, delta_prev, delta_zero, delta_min, delta_max,
                                                  delta_start, inc, dec, max_it))

    def direction_resident(self, n_refine=3):
        self.check(self.L.opb_direction_resident(self.h, n_refine))

    def system_rhs(self, grad, cons, mu, a_norm_penalty, eta_P, eta_D, eta_mu, fetch=False):
        """System_rhs(iter, reduct_factors) on the device; the result is the resident rhs of the next
        direction_resident.  fetch=True also returns (dual_r, primal_r, comp_r)."""
        g, c = f64(grad), f64(cons)
        out = (np.empty(self.n), np.empty(self.m), np.empty(self.m)) if fetch else (None, None, None)
        self.check(self.L.opb_system_rhs(self.h, pf(g), pf(c), float(mu), float(a_norm_penalty), float(eta_P),
                                         float(eta_D), float(eta_mu), pf(out[0]), pf(out[1]), pf(out[2])))
        return out if fetch else None

    def step_bounds(self, frac_bd, predict_exp):
        out = np.empty(4)
        self.check(self.L.opb_step_bounds(self.h, float(frac_bd), float(predict_exp), pf(out)))
        return dict(norm_dx=out[0], norm_dy=out[1], norm_ds=out[2], max_step_s=out[3])

    def get_direction(self):
        dx = np.empty(self.n); dy = np.empty(self.m); ds = np.empty(self.m); err = np.empty(6)
        self.check(self.L.opb_get_direction(self.h, pf(dx), pf(dy), pf(ds), pf(err)))
        return dx, dy, ds, err

    def solve_resident(self, nsolves=1):
        self.check(self.L.opb_solve_resident(self.h, nsolves))

    def profile_factor(self, delta):
        """One attempt with per-kernel CUDA-event timing (opb_profile_factor)."""
        v = [ctypes.c_double() for _ in range(5)]
        ok = ctypes.c_int()
        self.check(self.L.opb_profile_factor(self.h, float(delta),
                                             *[ctypes.cast(ctypes.byref(x), c_f64p) for x in v], ctypes.byref(ok)))
        keys = ("total_ms", "cb_ms", "update_ms", "cb_flops", "update_flops")
        out = {k: x.value for k, x in zip(keys, v)}
        out["inertia_ok"] = ok.value
        return out

    def profile_levels(self, delta):
        """One attempt with the look-ahead streams as configured and CUDA-event marks at the phase
        boundaries of every level (opb_profile_levels): rows [pre_ms, panel_ms, cb_ms] per level."""
        nl = int(self.info("nlevels"))
        out = np.zeros(3 * nl + 2)
        n = ctypes.c_int(); tot = ctypes.c_double()
        self.check(self.L.opb_profile_levels(self.h, float(delta), pf(out), out.size, ctypes.byref(n),
                                             ctypes.cast(ctypes.byref(tot), c_f64p)))
        return {"levels": out[:3 * nl].reshape(nl, 3), "fill_ms": out[3 * nl], "trtri_ms": out[3 * nl + 1],
                "total_ms": tot.value}

    def sync_state(self):
        d = ctypes.c_double(); nf = ctypes.c_int(); st = ctypes.c_int(); err = np.empty(6)
        self.check(self.L.opb_sync_state(self.h, ctypes.cast(ctypes.byref(d), c_f64p),
                                         ctypes.byref(nf), ctypes.byref(st), pf(err)))
        return d.value, nf.value, st.value, err

    # --- introspection
    def info(self, key):
        v = ctypes.c_double()
        self.check(self.L.opb_get_info(self.h, key.encode(), ctypes.cast(ctypes.byref(v), c_f64p)))
        return v.value

    def symbolic(self, name):
        cnt = self.L.opb_get_symbolic(self.h, name.encode(), None, 0)
        if cnt < 0:
            raise OPBError(int(cnt), self.L.opb_last_error(self.h).decode())
        out = np.empty(cnt, np.int64)
        self.L.opb_get_symbolic(self.h, name.encode(), pi(out), cnt)
        return out

    def L_values(self):
        v = np.empty(int(self.info("nnzL")))
        self.check(self.L.opb_get_L_values(self.h, pf(v), v.shape[0]))
        return v


def launch_count():
    return int(load().opb_launch_count())


def cache_clear():
    """Drop the per-process cache of symbolic analyses (structures bound to live handles stay valid)."""
    return int(load().opb_cache_clear())
