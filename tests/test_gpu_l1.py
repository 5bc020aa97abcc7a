"""GPU tests of the L1 boundary (`linear_solver_B200` = linear_solver_JULIA's contract,
linear_system_solvers/julia.jl:1-113) and of the single-shot `factor!` entry, all through the
C ABI on the device:

  * the reference's own L1 test (test/linear_system_solvers.jl:58-116) run on the CUDA path:
    both matrices, both `sym` modes, inertia == 1, `ls_solve!` bitwise == `ls_solve`,
    LDL' vs Cholesky < 1e-9, lower-only vs symmetrised < 1e-9;
  * LDL' (`OPB_MODE_LDLT`) on quasi-definite KKT matrices [[H + dI, J'], [J, -S/Y]] with inertia
    (n, m), on indefinite matrices, on zero / NaN pivots -- pivots, flags and solutions against
    the oracle's LDL' with the same permutation; sizes that reach every front class
    (shared-memory fronts, the blocked big-front kernels);
  * `factor!(kkt_solver, 1e-8)` followed by `compute_direction!` (test/kkt_system_solvers.jl:75-81)
    and the failure-driven refactorisation of one_phase.jl:231-242 with the real solver.
"""
import numpy as np
import pytest
import scipy.sparse as sp

import problems

pytestmark = pytest.mark.gpu

TOL = 1e-9          # the reference's own tolerance in run_linear_solvers (:62)
REL_TOL = 1e-10     # BASELINE.json north_star


def _ref_matrices():
    # test/linear_system_solvers.jl:94-116
    A1 = sp.identity(10, format="csc")
    A2 = sp.identity(10, format="lil")
    A2[9, 0] = 0.1
    A2[8, 1] = 0.1
    return [A1, sp.csc_matrix(A2)]


def _solver(pkg, sym):
    s = pkg.linear_solver_B200(sym, False, False)
    s.initialize()
    return s


def _run_one(pkg, sym, A, b, n, m, inertia):
    """test_julia_sym / test_julia_chol (test/linear_system_solvers.jl:18-44)."""
    s = _solver(pkg, sym)
    assert inertia == s.ls_factor(A, n, m)
    res1 = np.zeros(len(b))
    s.ls_solve_inplace(b, res1)
    res2 = s.ls_solve(b)
    assert np.array_equal(res1, res2)                      # @test res1 == res2 (bitwise)
    assert res2 is not res1
    s.finalize()
    return res1


@pytest.mark.parametrize("which", [0, 1])
def test_reference_l1_contract_on_device(pkg, orc, which):
    """run_linear_solvers (test/linear_system_solvers.jl:58-92) on the CUDA path."""
    A = _ref_matrices()[which]
    b = np.random.default_rng(which).random(10)
    n, m, inertia = 10, 0, 1
    d_sym = _run_one(pkg, "symmetric", A, b, n, m, inertia)
    d_chol = _run_one(pkg, "definite", A, b, n, m, inertia)
    assert np.linalg.norm(d_sym - d_chol) < TOL            # :67
    A_2 = sp.lil_matrix(A + A.T)
    A_2.setdiag(A.diagonal())
    A_2 = sp.csc_matrix(A_2)
    d_sym2 = _run_one(pkg, "symmetric", A_2, b, n, m, inertia)
    assert np.linalg.norm(d_sym - d_sym2) < TOL            # :82 lower-only == symmetrised
    d_chol2 = _run_one(pkg, "definite", A_2, b, n, m, inertia)
    assert np.linalg.norm(d_chol - d_chol2) < TOL          # :84
    # and against the oracle / the mathematics
    F = orc.Factor(sp.tril(A, format="csc"))
    assert F.factorize(sp.tril(A, format="csc").data, mode="chol") == 1
    assert np.linalg.norm(d_chol - F.solve(b)) <= REL_TOL * np.linalg.norm(b)
    Afull = (sp.tril(A) + sp.tril(A, -1).T).toarray()
    assert np.allclose(Afull @ d_chol, b, atol=1e-14)
    assert np.allclose(Afull @ d_sym, b, atol=1e-14)


def test_ls_solve_accepts_sparse_rhs(pkg):
    """julia.jl:107-110: a SparseVector rhs is densified."""
    A = _ref_matrices()[1]
    s = _solver(pkg, "definite")
    assert s.ls_factor(A, 10, 0) == 1
    b = np.zeros(10); b[3] = 2.0; b[9] = -1.0
    x_dense = s.ls_solve(b)
    x_sparse = s.ls_solve(sp.csc_matrix(b.reshape(-1, 1)))
    assert np.array_equal(x_dense, x_sparse)
    s.finalize()


def _kkt_matrix(prob, delta):
    """Lower triangle of the symmetric KKT matrix [[H + delta I, J'], [J, -S/Y]]
    (kkt_system_solver/symmetric.jl:35-52,87-99)."""
    n, m = prob.n, prob.m
    Hd = sp.csc_matrix(prob.H) + delta * sp.identity(n, format="csc")
    B = sp.diags(-prob.s / prob.y)
    K = sp.bmat([[Hd, None], [prob.J, B]], format="csc")
    K = sp.tril(K, format="csc")
    K.sort_indices()
    return K


def _ldlt_compare(pkg, orc, K, n_pos, m_neg, expect=None, rel_tol=REL_TOL, seed=0):
    """GPU LDL' against the oracle's LDL' under the same permutation: completion flag, inertia
    flag, the pivots D and a solve."""
    s = _solver(pkg, "symmetric")
    ok = s.ls_factor(K, n_pos, m_neg)
    h = s._h
    perm = h.symbolic("perm")
    F = orc.Factor(K, perm)
    done = F.factorize(K.data, mode="ldlt")
    ok_o = 1 if (done == 1 and F.ldlt_inertia_ok(n_pos, m_neg) == 1) else 0
    assert ok == ok_o, (ok, ok_o)
    if expect is not None:
        assert ok == expect
    if done == 1:
        Dg = h.L_values()[h.symbolic("dpos")][perm]
        Do = F.diag()
        assert np.array_equal(np.sign(Dg), np.sign(Do))
        assert np.allclose(Dg, Do, rtol=1e-9, atol=0.0), np.abs(Dg / Do - 1).max()
        b = np.random.default_rng(seed).standard_normal(K.shape[0])
        x = s.ls_solve(b)
        xo = F.solve(b)
        rel = np.linalg.norm(x - xo) / np.linalg.norm(xo)
        assert rel <= rel_tol, rel
        Kf = K + sp.tril(K, -1).T
        res = np.abs(Kf @ x - b).max() / max(np.abs(b).max(), np.abs(Kf).max() * np.abs(x).max())
        assert res <= 1e-10, res
    info = dict(n_big=h.info("n_big"), n_small=h.info("n_small"), n_tiny=h.info("n_tiny"), max_front=h.info("max_front"))
    s.finalize()
    return ok, info


@pytest.mark.parametrize("name", ["toy_lp1", "toy_lp5", "readme"])
def test_ldlt_quasidefinite_toys(pkg, orc, name):
    p = problems.toy(name)
    K = _kkt_matrix(p, 1e-8 if name != "readme" else 1.0)
    _ldlt_compare(pkg, orc, K, p.n, p.m, expect=1)


def test_ldlt_quasidefinite_chain(pkg, orc):
    p = problems.chain(nh=150, seed=2)
    K = _kkt_matrix(p, 1e-6)
    ok, info = _ldlt_compare(pkg, orc, K, p.n, p.m, expect=1)
    # wrong expected inertia -> 0 (inertia_status, linear_system_solvers.jl:48-91)
    s = _solver(pkg, "symmetric")
    assert s.ls_factor(K, p.n + 1, p.m - 1) == 0
    assert s.ls_factor(K, p.n, p.m) == 1
    s.finalize()


def test_ldlt_quasidefinite_reaches_big_front_kernels(pkg, orc):
    """The KKT matrix of a 3-D grid problem: top separators far beyond the 152-row shared-memory
    fronts, so the blocked LDL' path of the big fronts runs -- the tensor-core tile engine with the
    pivots applied inside it, the diagonal-block / panel kernels in their LDL' variant, the pivot-block
    inverses and the multi-CTA solves."""
    p = problems.pde_control(9, seed=1)
    K = _kkt_matrix(p, 1e-6)
    ok, info = _ldlt_compare(pkg, orc, K, p.n, p.m, expect=1)
    assert info["n_big"] >= 1 and info["max_front"] > 152, info


def test_ldlt_tensor_path_matches_the_scalar_path(pkg):
    """The two LDL' implementations of the big fronts (tile engine, default; scalar 32-column blocks,
    option ldlt_scalar) on a quasi-definite matrix with multi-block pivot blocks: same completion /
    inertia flags, pivots to 1e-9, solutions to 1e-10, both with a small residual."""
    p = problems.pde_control(12, seed=3)
    K = _kkt_matrix(p, 1e-6)
    b = np.random.default_rng(5).standard_normal(K.shape[0])
    Kf = K + sp.tril(K, -1).T
    out = []
    for scalar in (0, 1):
        s = _solver(pkg, "symmetric")
        s._h.set_option("ldlt_scalar", scalar)
        ok = s.ls_factor(K, p.n, p.m)
        h = s._h
        D = h.L_values()[h.symbolic("dpos")]
        x = s.ls_solve(b)
        res = np.abs(Kf @ x - b).max() / max(np.abs(b).max(), np.abs(Kf).max() * np.abs(x).max())
        assert res <= 1e-10, (scalar, res)
        out.append((ok, D, x, h.info("n_big"), h.info("n_trtri")))
        s._h.set_option("ldlt_scalar", 0)
        s.finalize()
    (ok0, D0, x0, nbig, ntr), (ok1, D1, x1, _, _) = out
    assert ok0 == ok1 == 1
    assert nbig >= 1 and ntr >= 1, (nbig, ntr)       # pivot blocks of more than 128 columns took part
    assert np.array_equal(np.sign(D0), np.sign(D1))
    assert np.allclose(D0, D1, rtol=1e-9, atol=0.0)
    assert np.linalg.norm(x0 - x1) <= REL_TOL * np.linalg.norm(x1)


def test_ldlt_dense_front(pkg, orc):
    """Dense Schur complement (COPS elec): one front of 3 n_p rows through the blocked LDL' path,
    positive definite after the shift the Cholesky delta loop accepts."""
    p = problems.elec(80, seed=3)
    Q, sd = orc.form_system(p.J, p.H, p.y, p.s)
    QL = sp.tril(Q, format="csc"); QL.sort_indices()
    F = orc.Factor(QL)
    st, nf, delta, _ = F.delta_loop(QL.data, sd, 0.0)
    assert st == "success"
    Qs = QL.copy(); Qs.setdiag(sd + delta)
    ok, info = _ldlt_compare(pkg, orc, sp.csc_matrix(Qs), p.n, 0, expect=1)
    assert info["max_front"] == p.n
    # unshifted: indefinite -> inertia flag 0, but the factorisation itself completes
    if nf > 1:
        s = _solver(pkg, "symmetric")
        assert s.ls_factor(QL, p.n, 0) == 0      # (no-pivot LDL' of an indefinite matrix: flag only, the pivots may grow)
        s.finalize()


def test_ldlt_indefinite_zero_and_nan_pivots(pkg, orc):
    A = sp.csc_matrix(np.array([[2.0, 0, 0], [1.0, -3.0, 0], [0.5, 0.2, 4.0]]))
    sc = _solver(pkg, "definite")
    assert sc.ls_factor(A, 3, 0) == 0                      # PosDefException -> 0 (julia.jl:39-41)
    sc.finalize()
    s = _solver(pkg, "symmetric")
    s._h.set_option("ordering", 1)                         # natural order: the pivots below are those of the text
    assert s.ls_factor(A, 3, 0) == 0
    assert s.ls_factor(A, 2, 1) == 1
    D = s._h.L_values()[s._h.symbolic("dpos")]
    assert np.allclose(D, [2.0, -3.5, 4.0 - 0.125 - (0.2 - 0.25) ** 2 / -3.5])
    Z = sp.csc_matrix(np.array([[0.0, 0], [1.0, 1.0]]))
    assert s.ls_factor(Z, 2, 0) == 0                       # ZeroPivotException -> 0 (julia.jl:61-63)
    assert s.ls_factor(Z, 1, 1) == 0
    N = sp.csc_matrix(np.array([[np.nan, 0], [1.0, 1.0]]))
    assert s.ls_factor(N, 2, 0) == 0
    s.finalize()
    sc = _solver(pkg, "definite")
    assert sc.ls_factor(N, 2, 0) == 0
    sc.finalize()
    _ldlt_compare(pkg, orc, A, 2, 1, expect=1)


def test_ldlt_vs_cholesky_on_a_sparse_spd_matrix(pkg, orc):
    """LDL' and Cholesky of the same SPD matrix agree (< 1e-9, the reference's :67 relation) on a
    matrix with all front classes; D = diag(L_chol)^2."""
    p = problems.sparse_qp(3000, 1500, seed=11)
    Q, sd = orc.form_system(p.J, p.H, p.y, p.s)
    QL = sp.tril(Q, format="csc"); QL.sort_indices()
    b = p.rhs[0][0]
    sc = _solver(pkg, "definite")
    assert sc.ls_factor(QL, p.n, 0) == 1
    xc = sc.ls_solve(b)
    Lc = sc._h.L_values()[sc._h.symbolic("dpos")]
    ss = _solver(pkg, "symmetric")
    assert ss.ls_factor(QL, p.n, 0) == 1
    xs = ss.ls_solve(b)
    D = ss._h.L_values()[ss._h.symbolic("dpos")]
    assert np.linalg.norm(xc - xs) < TOL * max(1.0, np.linalg.norm(xc))
    assert np.allclose(D, Lc ** 2, rtol=1e-8)
    F = orc.Factor(QL, sc._h.symbolic("perm"))
    assert F.factorize(QL.data, mode="chol") == 1
    xo = F.solve(b)
    assert np.linalg.norm(xc - xo) <= REL_TOL * np.linalg.norm(xo)
    sc.finalize(); ss.finalize()


def test_l1_refactor_same_pattern_new_values(pkg, orc):
    """ls_factor! is called once per delta attempt with the same pattern (delta_strategy.jl:66,93):
    the cached analysis is reused and the new values are factorised."""
    p = problems.chain(nh=60, seed=5)
    Q, sd = orc.form_system(p.J, p.H, p.y, p.s)
    QL = sp.tril(Q, format="csc"); QL.sort_indices()
    s = _solver(pkg, "definite")
    b = p.rhs[0][0]
    for delta in (0.0, 1e-3, 7.0):
        Qd = QL.copy(); Qd.setdiag(sd + delta); Qd = sp.csc_matrix(Qd)
        assert s.ls_factor(Qd, p.n, 0) == 1
        F = orc.Factor(Qd, s._h.symbolic("perm"))
        assert F.factorize(Qd.data, mode="chol") == 1
        x, xo = s.ls_solve(b), F.solve(b)
        assert np.linalg.norm(x - xo) <= REL_TOL * np.linalg.norm(xo)
    assert s._h.info("symbolic_cached") == 1
    s.finalize()


# ---------------------------------------------------------------------------
# single-shot factor!(kkt_solver, delta) and the failure-driven refactorisation
# ---------------------------------------------------------------------------
def _oracle_single(orc, prob, delta, perm):
    Q, sd = orc.form_system(prob.J, prob.H, prob.y, prob.s)
    QL = sp.tril(Q, format="csc"); QL.sort_indices()
    F = orc.Factor(QL, perm)
    ok = F.factorize(QL.data, sd + delta)
    return F, ok


@pytest.mark.parametrize("name", ["toy_lp%d" % i for i in range(9)] + ["chain", "sparse_qp", "pde"])
def test_single_shot_factor_then_direction(pkg, orc, name):
    """test/kkt_system_solvers.jl:75-81: form_system!, factor!(kkt_solver, 1e-8), kkt_associate_rhs!,
    compute_direction!."""
    if name == "chain":
        prob = problems.chain(nh=400, seed=9)
    elif name == "sparse_qp":
        prob = problems.sparse_qp(4000, 2000, seed=2)
    elif name == "pde":
        prob = problems.pde_control(10, seed=2)
    else:
        prob = problems.toy(name)
    pars = pkg.Class_parameters()
    it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s)
    k = pkg.pick_KKT_solver(pars)
    k.initialize(it)
    k.form_system(it)
    with pytest.raises(RuntimeError, match="not ready to compute direction"):
        k.compute_direction()
    inertia = k.factor(1e-8)
    F, ok = _oracle_single(orc, prob, 1e-8, k._h.symbolic("perm"))
    assert inertia == ok
    assert k.ready == "factored" and np.all(k.delta_x_vec == 1e-8) and not np.any(k.delta_s_vec)
    if ok == 1:
        for r in prob.rhs:
            k.kkt_associate_rhs(it, pkg.System_rhs(*r))
            k.compute_direction()
            dxo, dyo, dso, erro = F.direction(prob.J, prob.H, prob.y, prob.s, 1e-8, *r)
            for a, b_ in ((k.dir.x, dxo), (k.dir.y, dyo), (k.dir.s, dso)):
                assert np.linalg.norm(a - b_) <= REL_TOL * max(np.linalg.norm(b_), 1e-300)
            assert k.kkt_err_norm.ratio <= max(10 * erro[5], 1e-13)          # a ratio of rounding residuals
    k.finalize()


def test_single_shot_factor_reports_not_pd(pkg, orc):
    prob = problems.chain(nh=80, seed=1, offdiag_curv=25.0)
    pars = pkg.Class_parameters()
    it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s)
    k = pkg.pick_KKT_solver(pars)
    k.initialize(it)
    k.form_system(it)
    perm = k._h.symbolic("perm")
    seen = set()
    for delta in (0.0, 1e-8, 1e-2, 1.0, 1e2, 1e4):
        got = k.factor(delta)
        _, want = _oracle_single(orc, prob, delta, perm)
        assert got == want, (delta, got, want)
        seen.add(got)
    assert seen == {0, 1}, "the sweep must cross the PD threshold"
    k.finalize()


def test_respond_to_failed_step_with_the_real_solver(pkg, orc):
    """one_phase.jl:231-242: after a failed line search delta <- max(|grad L|/|dx|, 8 delta,
    max(1e-6, delta_old/pi)), ONE factor! at that delta, then directions come from the new factor."""
    prob = problems.chain(nh=200, seed=4, offdiag_curv=5.0)
    pars = pkg.Class_parameters()
    it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s, delta=0.0)
    k = pkg.pick_KKT_solver(pars)
    k.initialize(it)
    k.form_system(it)
    st, nf, delta = pkg.ipopt_strategy(it, k, pars)
    assert st == "success"
    pkg.kkt.set_delta(it, delta)
    r = prob.rhs[0]
    k.kkt_associate_rhs(it, pkg.System_rhs(*r))
    k.compute_direction()
    dx_inf = np.abs(k.dir.x).max()
    grad_lag_inf = 3.0 * dx_inf * max(delta, 1e-3)            # makes the first term the active one
    new_delta, inertia = pkg.respond_to_failed_step(it, k, pars, old_delta=0.0, grad_lag_inf=grad_lag_inf)
    assert new_delta == max(grad_lag_inf / dx_inf, delta * 8.0, max(1e-6, 0.0)) and it.delta == new_delta
    F, ok = _oracle_single(orc, prob, new_delta, k._h.symbolic("perm"))
    assert inertia == ok == 1
    k.compute_direction()
    dxo, dyo, dso, erro = F.direction(prob.J, prob.H, prob.y, prob.s, new_delta, *r)
    for a, b_ in ((k.dir.x, dxo), (k.dir.y, dyo), (k.dir.s, dso)):
        assert np.linalg.norm(a - b_) <= REL_TOL * max(np.linalg.norm(b_), 1e-300)
    # a zero direction gives Inf like Julia's 1/0 (no ZeroDivisionError): delta = Inf, factor! still runs
    k.dir.x[:] = 0.0
    new2, inertia2 = pkg.respond_to_failed_step(it, k, pars, old_delta=new_delta, grad_lag_inf=1.0)
    assert np.isinf(new2)
    k.finalize()


# ---------------------------------------------------------------------------
# Symmetric_KKT_solver on the device LDL' -- the reference's cross-formulation check
# ---------------------------------------------------------------------------
def _direction_through(pkg, kind, prob, delta, r):
    """test_kkt_solver (test/kkt_system_solvers.jl:61-90) for one kkt_solver_type."""
    pars = pkg.Class_parameters()
    pars.kkt.kkt_solver_type = kind
    it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s)
    k = pkg.pick_KKT_solver(pars)
    k.initialize(it)
    k.form_system(it)
    inertia = k.factor(delta)
    k.kkt_associate_rhs(it, pkg.System_rhs(*r))
    k.compute_direction()
    out = (inertia, k.dir.x.copy(), k.dir.y.copy(), k.dir.s.copy(), k.kkt_err_norm, k.schur_diag.copy())
    k.finalize()
    return out


@pytest.mark.parametrize("name", ["toy_lp%d" % i for i in range(9)] + ["chain", "sparse_qp", "pde"])
def test_schur_and_symmetric_formulations_agree(pkg, orc, name):
    """test_kkt_solvers (test/kkt_system_solvers.jl:92-120): the Schur-complement solver (Cholesky)
    and the symmetric solver (LDL' of the quasi-definite system, inertia (n, m)) give the same
    direction, norm(diff, 2) < 1e-6 at delta = 1e-8 -- both on the device."""
    if name == "chain":
        prob = problems.chain(nh=300, seed=6)
    elif name == "sparse_qp":
        prob = problems.sparse_qp(3000, 1500, seed=6)
    elif name == "pde":
        prob = problems.pde_control(8, seed=6)
    else:
        prob = problems.toy(name)
    r = prob.rhs[0]
    i1, x1, y1, s1, e1, sd1 = _direction_through(pkg, "schur_b200", prob, 1e-8, r)
    i2, x2, y2, s2, e2, sd2 = _direction_through(pkg, "symmetric_b200", prob, 1e-8, r)
    assert i1 == 1 and i2 == 1
    scale = max(1.0, np.linalg.norm(x1), np.linalg.norm(y1), np.linalg.norm(s1))
    assert np.linalg.norm(x1 - x2) < 1e-6 * scale
    assert np.linalg.norm(y1 - y2) < 1e-6 * scale
    assert np.linalg.norm(s1 - s2) < 1e-6 * scale
    assert e2.ratio < 1e-6 and e2.rhs_norm == e1.rhs_norm
    # compute_schur_diag (kkt_system_solver.jl:296-300) equals the assembled diagonal up to rounding
    assert np.allclose(sd2, sd1, rtol=1e-12, atol=0.0)


def test_symmetric_solver_inertia_drives_delta(pkg):
    """An indefinite Hessian: the symmetric solver's inertia flag is 0 until delta makes
    H + delta I + J' S^-1 Y J positive definite, exactly when the Cholesky flag of the Schur
    solver turns 1 (the two formulations have the same inertia)."""
    prob = problems.chain(nh=60, seed=1, offdiag_curv=25.0)
    r = prob.rhs[0]
    seen = set()
    for delta in (1e-8, 1.0, 1e2, 1e4):
        pars = pkg.Class_parameters()
        flags = []
        for kind in ("schur_b200", "symmetric_b200"):
            pars.kkt.kkt_solver_type = kind
            it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s)
            k = pkg.pick_KKT_solver(pars); k.initialize(it); k.form_system(it)
            flags.append(k.factor(delta))
            k.finalize()
        assert flags[0] == flags[1], (delta, flags)
        seen.add(flags[0])
    assert seen == {0, 1}


# ---------------------------------------------------------------------------
# estimate_y_tilde, compute_schur_diag / eval_diag_J_T_J
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["toy_lp5", "chain", "sparse_qp"])
def test_estimate_y_tilde_and_schur_diag(pkg, orc, name):
    prob = {"toy_lp5": lambda: problems.toy("toy_lp5"), "chain": lambda: problems.chain(nh=100, seed=3),
            "sparse_qp": lambda: problems.sparse_qp(2000, 1000, seed=3)}[name]()
    g = np.random.default_rng(1).standard_normal(prob.n)
    y = pkg.estimate_y_tilde(prob.J, g)
    yo = orc.estimate_y_tilde(prob.J, g)
    assert np.linalg.norm(y - yo) <= 1e-9 * max(np.linalg.norm(yo), 1e-300)
    it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s)
    sd = pkg.compute_schur_diag(it)
    assert np.array_equal(sd, orc.compute_schur_diag(prob.J, prob.H, prob.y, prob.s))      # same operation order: bit-exact
    d = np.random.default_rng(2).random(prob.m)
    assert np.array_equal(pkg.eval_diag_J_T_J(it, d), orc.eval_diag_J_T_J(prob.J, d))


# ---------------------------------------------------------------------------
# System_rhs and the step bounds on the device, resident iterate
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["toy_lp5", "chain", "sparse_qp"])
@pytest.mark.parametrize("etas", [(0.0, 0.0, 0.0), (1.0, 0.0, 1.0), (0.3, 0.0, 0.3)])      # affine, stabilisation, aggressive
def test_device_system_rhs_and_step_bounds(pkg, orc, name, etas):
    prob = {"toy_lp5": lambda: problems.toy("toy_lp5"), "chain": lambda: problems.chain(nh=80, seed=8),
            "sparse_qp": lambda: problems.sparse_qp(1500, 700, seed=8)}[name]()
    rng = np.random.default_rng(5)
    grad, cons = rng.standard_normal(prob.n), prob.s + 0.1 * rng.standard_normal(prob.m)
    mu, a_pen = 0.37, 1e-4
    pars = pkg.Class_parameters()
    it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s)
    k = pkg.pick_KKT_solver(pars); k.initialize(it); k.form_system(it)
    st, nf, delta = pkg.ipopt_strategy(it, k, pars)
    assert st == "success"
    h = k._h
    got = h.system_rhs(grad, cons, mu, a_pen, *etas, fetch=True)
    want = orc.system_rhs(prob.J, prob.y, prob.s, grad, cons, mu, a_pen, *etas)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)                       # same operation order: bit-exact
    # the resident rhs feeds the next direction: identical to passing the same vectors through the host
    h.direction_resident(3)
    dx, dy, ds, err = h.get_direction()
    k.kkt_associate_rhs(it, pkg.System_rhs(*want))
    k.compute_direction()
    assert np.array_equal(dx, k.dir.x) and np.array_equal(dy, k.dir.y) and np.array_equal(ds, k.dir.s)
    assert err[5] == k.kkt_err_norm.ratio
    # fraction-to-the-boundary scalars: only four doubles cross PCIe
    sb = h.step_bounds(0.1, 1.5)
    so = orc.step_bounds(prob.s, dx, dy, ds, 0.1, 1.5)
    assert sb["norm_dx"] == so["norm_dx"] and sb["norm_dy"] == so["norm_dy"] and sb["norm_ds"] == so["norm_ds"]
    assert sb["max_step_s"] == pytest.approx(so["max_step_s"], rel=1e-14)
    k.finalize()
