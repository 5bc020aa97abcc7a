"""The C-ABI library loads without a GPU, exports every symbol the header declares,
and refuses numeric work on a host-only handle (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header(pkg):
    hdr = open(os.path.join(ROOT, "include", "onephase_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(opb_[a-zA-Z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    L = ctypes.CDLL(pkg._lib.LIB_PATH)
    for sym in declared:
        assert hasattr(L, sym), "library does not export %s" % sym
    assert sorted(pkg._lib.EXPORTS) == declared


def test_library_exports_nothing_but_the_abi(pkg):
    """csrc/exports.map: the statically linked METIS / GKlib and the internal C++ symbols stay local, so
    another METIS in the host process can neither clash with them nor be called in their place."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", pkg._lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    names = [ln.split()[-1] for ln in out.splitlines() if ln.strip()]
    assert names and all(nm.startswith("opb_") for nm in names), [nm for nm in names if not nm.startswith("opb_")][:10]
    assert sorted(set(names)) == sorted(pkg._lib.EXPORTS)


def test_cache_clear_drops_cached_analyses_but_not_bound_ones(pkg):
    prob = problems.chain(7, seed=3)
    args = (prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)
    h1 = pkg.Handle(-1)
    h1.set_structure(*args)
    h2 = pkg.Handle(-1)
    h2.set_structure(*args)
    assert h2.info("symbolic_cached") == 1          # the second solver object of a solve shares the analysis
    assert pkg.cache_clear() >= 1
    assert pkg.cache_clear() == 0
    assert h1.info("nsuper") == h2.info("nsuper") > 0 and len(h1.symbolic("perm")) == prob.n   # still bound, still valid
    h3 = pkg.Handle(-1)
    h3.set_structure(*args)
    assert h3.info("symbolic_cached") == 0          # analysed again
    assert np.array_equal(h3.symbolic("perm"), h1.symbolic("perm"))


def test_version_and_launch_counter(pkg):
    L = pkg._lib.load()
    assert b"sm_100a" in L.opb_version()
    assert pkg.launch_count() >= 0


def test_library_is_sm100a_only(pkg):
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", pkg._lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_numeric_calls_fail_loudly_without_device(pkg):
    prob = problems.chain(5, seed=0)
    h = pkg.Handle(-1)
    h.set_structure(prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)
    with pytest.raises(pkg.OPBError) as e:
        h.form(prob.J.data, prob.H.data, prob.y, prob.s)
    assert e.value.code == pkg._lib.OPB_ERR_NO_DEVICE
    for call in (lambda: h.factor(0.0), lambda: h.factor_delta_loop(0.0),
                 lambda: h.direction(*prob.rhs[0]), lambda: h.ls_solve(np.zeros(prob.n)),
                 lambda: h.ls_factor_csc(2, [0, 1, 2], [0, 1], [1.0, 1.0], 0, 0, 2, 0)):
        with pytest.raises(pkg.OPBError) as e:
            call()
        assert e.value.code == pkg._lib.OPB_ERR_NO_DEVICE


def test_no_gpu_means_create_fails(pkg, have_gpu):
    if have_gpu:
        pytest.skip("GPU present")
    with pytest.raises(pkg.OPBError) as e:
        pkg.Handle(0)
    assert e.value.code == pkg._lib.OPB_ERR_CUDA


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package or include/ refers to it."""
    bad = []
    for base in ("onephase.jl_b200", "include", "julia"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cpp", ".h", ".jl")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(import\s+oracle|from\s+oracle|kkt_oracle|libkkt_oracle)", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_host_mirror_state_machine(pkg):
    """ready-state errors of kkt_system_solver.jl:181-183,195-199 without touching the device."""
    k = pkg.Schur_B200_KKT_solver()
    with pytest.raises(RuntimeError, match="not ready to compute direction"):
        k.compute_direction()
    with pytest.raises(RuntimeError, match="not ready to factor"):
        k._factor()
    with pytest.raises(ValueError):
        pkg.linear_solver_B200("unsymmetric")
    pars = pkg.Class_parameters()
    pars.kkt.linear_solver_type = "julia"
    with pytest.raises(ValueError, match="pick a valid solver"):
        pkg.pick_KKT_solver(pars)


def test_options_are_validated(pkg):
    """opb_set_option (include/onephase_b200.h): known keys are accepted on a host-only handle,
    unknown keys are an error, outer_block is rounded to a multiple of the 128-column block."""
    h = pkg.Handle(-1)
    for key, v in (("ordering", 0), ("nd_leaf", 64), ("nd_balance", 0.35), ("relax", 1), ("attempts_per_sync", 3), ("graphs", 0),
                   ("outer_block", 1000), ("lookahead", 0), ("barrier_timeout_s", 2.5), ("metis_max_n", 1000)):
        h.set_option(key, v)
    with pytest.raises(pkg.OPBError):
        h.set_option("no_such_option", 1)
    prob = problems.chain(nh=30, seed=1)
    h.set_structure(prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)
    assert h.info("n") == prob.n
    h.close()


def test_failure_driven_delta_increase_rule(pkg):
    """respond_to_failed_step = one_phase.jl:231-242: delta <- max(|grad L|/|dx|, 8 delta,
    max(delta.start, delta_old/pi)) (or without the first term for :default), then ONE factor!."""
    class FakeSolver:
        def __init__(self):
            self.dir = type("P", (), {"x": np.array([0.5, -2.0, 1.0])})()
            self.calls = []

        def factor(self, delta):
            self.calls.append(delta)
            return 1
    pars = pkg.Class_parameters()
    prob = problems.toy("toy_lp1")
    it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s, delta=1e-3)
    k = FakeSolver()
    new, inertia = pkg.respond_to_failed_step(it, k, pars, old_delta=1e-4, grad_lag_inf=10.0)
    assert new == max(10.0 / 2.0, 8e-3, max(1e-6, 1e-4 / np.pi)) == 5.0 and it.delta == 5.0
    assert k.calls == [5.0] and inertia == 1
    it.delta = 1e-3
    new, _ = pkg.respond_to_failed_step(it, k, pars, old_delta=1.0, response="default")
    assert new == max(8e-3, 1.0 / np.pi)
    with pytest.raises(ValueError):
        pkg.respond_to_failed_step(it, k, pars, old_delta=0.0, response="nonsense")
