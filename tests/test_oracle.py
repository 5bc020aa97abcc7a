"""Pins the CPU oracle (oracle/kkt_oracle.c).

The reference has no numeric golden vectors for this path, so the
oracle is pinned by (i) the relational properties the reference's own tests assert
(test/linear_system_solvers.jl:58-116), (ii) dense numpy / SuperLU cross-checks of
the same mathematics, (iii) a line-by-line Python restatement of the delta rule
(delta_strategy.jl:37-114) and (iv) committed golden fixtures minted from the
oracle itself (tests/golden/make_golden.py) that freeze its behaviour.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import problems

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ref_matrices():
    # test/linear_system_solvers.jl:94-116
    A1 = sp.identity(10, format="csc")
    A2 = sp.identity(10, format="lil")
    A2[9, 0] = 0.1
    A2[8, 1] = 0.1
    return [A1, sp.csc_matrix(A2)]


@pytest.mark.parametrize("which", [0, 1])
def test_reference_linear_solver_relations(orc, which):
    """run_linear_solvers (test/linear_system_solvers.jl:58-92)."""
    A = _ref_matrices()[which]
    rng = np.random.default_rng(which)
    b = rng.random(10)
    n, m, tol = 10, 0, 1e-9
    Fs = orc.Factor(A)
    assert Fs.factorize(A.data, mode="ldlt") == 1 and Fs.ldlt_inertia_ok(n, m) == 1   # :23
    x_sym = Fs.solve(b)
    assert np.array_equal(x_sym, Fs.solve(b))                                         # :27 ls_solve! == ls_solve
    Fc = orc.Factor(A)
    assert Fc.factorize(A.data, mode="chol") == 1                                     # :36
    x_chol = Fc.solve(b)
    assert np.linalg.norm(x_sym - x_chol) < tol                                       # :67
    # A_2 = A + A' with the original diagonal: only the lower triangle may be read (:74-84)
    A_2 = sp.lil_matrix(A + A.T)
    A_2.setdiag(A.diagonal())
    A_2 = sp.csc_matrix(A_2)
    F2 = orc.Factor(A_2)
    assert F2.factorize(A_2.data, mode="ldlt") == 1
    assert np.linalg.norm(x_sym - F2.solve(b)) < tol
    assert F2.factorize(A_2.data, mode="chol") == 1
    assert np.linalg.norm(x_chol - F2.solve(b)) < tol
    # and it is the right answer
    Afull = (sp.tril(A) + sp.tril(A, -1).T).toarray()
    assert np.allclose(Afull @ x_chol, b, atol=1e-14)


def test_form_system_matches_dense(pkg, orc):
    for prob in [problems.toy("toy_lp5"), problems.chain(9, seed=1), problems.elec(12),
                 problems.sparse_qp(300, 150, win=5), problems.pde_control(4)]:
        Q, sd = orc.form_system(prob.J, prob.H, prob.y, prob.s)
        Jd = prob.J.toarray(); Hd = prob.H.toarray()
        Qd = Jd.T @ np.diag(prob.y / prob.s) @ Jd + Hd           # both triangles of J'DJ + lower H (schur.jl:55)
        assert np.allclose(Q.toarray(), Qd, rtol=1e-13, atol=1e-13 * np.abs(Qd).max())
        assert np.array_equal(sd, Q.diagonal())
        # structural product pattern: every pair of variables sharing a row is present
        pat = ((abs(prob.J).T @ abs(prob.J)) + abs(prob.H)).tocsc()
        assert Q.nnz == pat.nnz


def test_assembly_operation_order_is_k_ascending_no_fma(orc):
    """t = fl(J[k,i]*sig_k), p = fl(t*J[k,j]), summed over k ascending."""
    rng = np.random.default_rng(0)
    m, n = 7, 3
    J = sp.csc_matrix(rng.standard_normal((m, n)))
    H = sp.csc_matrix(np.tril(rng.standard_normal((n, n))))
    y = rng.random(m) + 0.1; s = rng.random(m) + 0.1
    Q, _ = orc.form_system(J, H, y, s)
    Jd = J.toarray(); Hd = H.toarray(); sig = y / s
    for i in range(n):
        for j in range(n):
            acc = None
            for k in range(m):
                p = np.float64(np.float64(Jd[k, i] * sig[k]) * Jd[k, j])
                acc = p if acc is None else np.float64(acc + p)
            if i >= j:
                acc = np.float64(acc + Hd[i, j])
            assert Q[i, j] == acc, (i, j)


@pytest.mark.parametrize("perm_kind", ["natural", "random", "rcm"])
def test_cholesky_and_ldlt_match_dense(pkg, orc, perm_kind):
    prob = problems.sparse_qp(400, 200, win=5, seed=3)
    Q, sd = orc.form_system(prob.J, prob.H, prob.y, prob.s)
    QL = sp.tril(Q, format="csc")
    n = prob.n
    if perm_kind == "natural":
        perm = None
    elif perm_kind == "random":
        perm = np.random.default_rng(1).permutation(n)
    else:
        from scipy.sparse.csgraph import reverse_cuthill_mckee
        perm = reverse_cuthill_mckee(sp.csr_matrix(QL + QL.T), symmetric_mode=True).astype(np.int64)
    F = orc.Factor(QL, perm)
    assert F.factorize(QL.data, mode="chol") == 1
    Md = (QL + sp.tril(QL, -1).T).toarray()
    b = prob.rhs[0][0]
    x = F.solve(b)
    xd = np.linalg.solve(Md, b)
    assert np.linalg.norm(x - xd) / np.linalg.norm(xd) < 1e-9
    # the diagonal of L equals that of the dense Cholesky factor of P M P'
    p = perm if perm is not None else np.arange(n)
    Ld = np.linalg.cholesky(Md[np.ix_(p, p)])
    assert np.allclose(F.diag(), np.diag(Ld), rtol=1e-9)
    # LDL'
    assert F.factorize(QL.data, mode="ldlt") == 1 and F.ldlt_inertia_ok(n, 0) == 1
    assert np.allclose(F.diag(), np.diag(Ld) ** 2, rtol=1e-8)
    assert np.linalg.norm(F.solve(b) - xd) / np.linalg.norm(xd) < 1e-9
    # SuperLU on the symmetrised matrix
    xs = spla.splu(sp.csc_matrix(Md)).solve(b)
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-9


def test_not_positive_definite_and_inertia(orc):
    A = sp.csc_matrix(np.array([[2.0, 0, 0], [1.0, -3.0, 0], [0.5, 0.2, 4.0]]))
    F = orc.Factor(A)
    assert F.factorize(A.data, mode="chol") == 0          # PosDefException -> 0 (julia.jl:39-41)
    assert F.factorize(A.data, mode="ldlt") == 1
    assert F.ldlt_inertia_ok(3, 0) == 0 and F.ldlt_inertia_ok(2, 1) == 1
    Z = sp.csc_matrix(np.array([[0.0, 0], [1.0, 1.0]]))
    Fz = orc.Factor(Z)
    assert Fz.factorize(Z.data, mode="ldlt") == 0         # ZeroPivotException -> 0 (julia.jl:61-63)
    N = sp.csc_matrix(np.array([[np.nan, 0], [1.0, 1.0]]))
    assert orc.Factor(N).factorize(N.data, mode="chol") == 0


def _delta_rule_python(try_factor, diag_min, delta_prev, zero=0.0, dmin=1e-12, dmax=1e50, start=1e-6,
                       inc=8.0, dec=1 / np.pi):
    """Literal transcription of ipopt_strategy! (delta_strategy.jl:37-114)."""
    num_fac = 0
    tau = 1.5 * diag_min
    delta = zero
    tried = []
    if tau > 0.0:
        tau = 0.0
        ok = try_factor(delta); num_fac += 1; tried.append(delta)
        if ok == 1:
            return "success", num_fac, delta, tried
    for i in range(1, 501):
        if i == 1:
            if delta_prev != 0.0:
                delta = max(dmin - tau, delta_prev * dec)
            else:
                delta = start - tau
        else:
            delta = delta * inc
        ok = try_factor(delta); num_fac += 1; tried.append(delta)
        if ok == 1:
            return "success", num_fac, delta, tried
        if delta > dmax:
            return "failure", num_fac, delta, tried
    raise RuntimeError("max it")


@pytest.mark.parametrize("case", ["pd", "indefinite", "neg_diag", "warm", "hopeless"])
def test_delta_loop_matches_transcription(pkg, orc, case):
    prob = problems.chain(40, seed=7, neg_curv=(40.0 if case == "neg_diag" else 0.0),
                              offdiag_curv=(25.0 if case in ("indefinite", "warm") else 0.0))
    if case == "neg_diag":
        H = prob.H.tolil(); H[0, 0] = -5e3; prob.H = sp.csc_matrix(H)
    Q, sd = orc.form_system(prob.J, prob.H, prob.y, prob.s)
    QL = sp.tril(Q, format="csc")
    if case == "hopeless":
        QL = QL.copy(); QL.data[:] = np.nan; sd = QL.diagonal()
    Md = (QL + sp.tril(QL, -1).T).toarray()
    n = prob.n

    def try_dense(delta):
        A = Md.copy()
        A[np.arange(n), np.arange(n)] = sd + delta
        try:
            if np.isnan(A).any():
                return 0
            np.linalg.cholesky(A)
            return 1
        except np.linalg.LinAlgError:
            return 0
    delta_prev = 0.0037 if case == "warm" else 0.0
    F = orc.Factor(QL)
    got = F.delta_loop(QL.data, sd, delta_prev)
    if case == "hopeless":
        # NaN diagonal: tau = NaN, every delta is NaN, `delta > DELTA_MAX` is never true -> error("max it")
        with pytest.raises(RuntimeError):
            _delta_rule_python(try_dense, np.nan, delta_prev)
        assert got[0] == "max_it" and got[1] == 500
        return
    want = _delta_rule_python(try_dense, sd.min(), delta_prev)
    assert got[0] == want[0] and got[1] == want[1]
    assert got[2] == want[2] or (np.isnan(got[2]) and np.isnan(want[2]))
    assert np.array_equal(got[3], np.array(want[3]), equal_nan=True)
    if case in ("indefinite", "warm"):
        assert got[1] >= 2
    if case == "neg_diag":
        assert got[3][0] == 1e-6 - 1.5 * sd.min()       # probe skipped, first shift lifted by -tau


def test_direction_matches_dense_kkt(pkg, orc):
    """schur.jl:89-128 solves the full KKT system  [H+dI  -J'; ... ] by elimination:
    check (dx,dy,ds) against a dense solve of the unreduced 3x3 block system."""
    prob = problems.chain(12, seed=2)
    rng = np.random.default_rng(0)
    prob.y = rng.uniform(0.5, 2, prob.m); prob.s = rng.uniform(0.5, 2, prob.m)
    Q, sd = orc.form_system(prob.J, prob.H, prob.y, prob.s)
    QL = sp.tril(Q, format="csc")
    F = orc.Factor(QL)
    delta = 1e-8
    assert F.factorize(QL.data, sd + delta) == 1
    rD, rP, rC = prob.rhs[0]
    dx, dy, ds, err = F.direction(prob.J, prob.H, prob.y, prob.s, delta, rD, rP, rC)
    n, m = prob.n, prob.m
    J = prob.J.toarray(); Hs = (prob.H + sp.tril(prob.H, -1).T).toarray()
    K = np.block([[Hs + delta * np.eye(n), -J.T, np.zeros((n, m))],
                  [J, np.zeros((m, m)), -np.eye(m)],
                  [np.zeros((m, n)), np.diag(prob.s), np.diag(prob.y)]])
    sol = np.linalg.solve(K, np.concatenate([rD, rP, rC]))
    assert np.allclose(dx, sol[:n], rtol=1e-8, atol=1e-10)
    assert np.allclose(dy, sol[n:n + m], rtol=1e-8, atol=1e-10)
    assert np.allclose(ds, sol[n + m:], rtol=1e-8, atol=1e-10)
    assert err[4] == max(np.abs(rD).max(), np.abs(rP).max(), np.abs(rC).max())
    assert err[5] < 1e-10 and err[3] == max(err[0], err[1], err[2])


def test_golden_fixtures(pkg, orc):
    """Frozen oracle outputs (minted by tests/golden/make_golden.py)."""
    import glob
    files = sorted(glob.glob(os.path.join(GOLD, "*.npz")))
    assert files, "golden fixtures missing"
    from golden.make_golden import CASES, run_case
    for f in files:
        g = np.load(f)
        name = os.path.basename(f)[:-4]
        out = run_case(pkg, orc, CASES[name])
        for key in g.files:
            assert np.array_equal(np.asarray(out[key]), g[key], equal_nan=True), (name, key)


@pytest.mark.parametrize("gen,kw,delta_prev", [("chain", dict(nh=120, offdiag_curv=5.0), 0.0),
                                               ("chain", dict(nh=60, neg_curv=30.0), 1e-3),
                                               ("sparse_qp", dict(n=2500, m_gen=1200), 0.0),
                                               ("pde_control", dict(N=9), 0.0), ("elec", dict(n_p=15), 0.0)])
def test_supernodal_cpu_baseline_matches_the_scalar_oracle(pkg, orc, gen, kw, delta_prev):
    """oracle/supernodal.py + snode.c (multifrontal + BLAS-3: the CPU baseline bench.py times with
    all host cores, and the large-case checker) against oracle/kkt_oracle.c: identical
    (status, #fac, delta) sequence, solves and directions within 1e-10 -- under each of the
    oracle's own orderings.  No product code is involved."""
    from oracle import supernodal
    prob = getattr(problems, gen)(seed=4, **kw)
    Q, sd = orc.form_system(prob.J, prob.H, prob.y, prob.s)
    QL = sp.tril(Q, format="csc"); QL.sort_indices()
    for ordering in ("metis", "mindeg", "natural"):
        F1 = supernodal.SupernodalFactor(QL, ordering=ordering)
        perm = F1.perm
        assert np.array_equal(np.sort(perm), np.arange(prob.n))
        F0 = orc.Factor(QL, perm)
        # symbolic: same fill and flops as the scalar up-looking oracle under the same permutation
        assert F1.info("nnzL_true") == F0.lnz and F1.info("flops") == pytest.approx(F0.flops, rel=1e-12)
        st0, nf0, d0, tried0 = F0.delta_loop(QL.data, sd, delta_prev)
        st1, nf1, d1, tried1 = F1.delta_loop(QL.data, sd, delta_prev)
        assert (st0, nf0, d0) == (st1, nf1, d1), ((st0, nf0, d0), (st1, nf1, d1))
        assert np.array_equal(tried0, tried1)
        assert np.allclose(F1.diag(), F0.diag(), rtol=1e-9)
        b = np.random.default_rng(0).standard_normal(prob.n)
        x0, x1 = F0.solve(b), F1.solve(b)
        assert np.linalg.norm(x0 - x1) <= 1e-9 * np.linalg.norm(x0)
        for r in prob.rhs:
            a = F0.direction(prob.J, prob.H, prob.y, prob.s, d0, *r)
            c = F1.direction(prob.J, prob.H, prob.y, prob.s, d1, *r)
            for u, v in zip(a[:3], c[:3]):
                assert np.linalg.norm(u - v) <= 1e-10 * max(np.linalg.norm(u), 1e-300)
            assert c[3][4] == pytest.approx(a[3][4], rel=1e-14)          # rhs norm
            assert c[3][5] <= max(30 * a[3][5], 1e-8)                    # N err (a ratio of rounding residuals: noisy)


def test_supernodal_oracle_without_relaxation_and_against_dense(orc):
    """snode.c against dense LAPACK: Cholesky pivots of P M P' and the solution, with and without
    relaxed amalgamation; not-PD and NaN matrices return 0 (julia.jl:39-41)."""
    from oracle import supernodal
    prob = problems.sparse_qp(500, 250, win=5, seed=9)
    Q, sd = orc.form_system(prob.J, prob.H, prob.y, prob.s)
    QL = sp.tril(Q, format="csc"); QL.sort_indices()
    Md = (QL + sp.tril(QL, -1).T).toarray()
    b = prob.rhs[0][0]
    xd = np.linalg.solve(Md, b)
    for relax in (True, False):
        F = supernodal.SupernodalFactor(QL, ordering="metis", relax=relax)
        assert F.factorize(QL.data) == 1
        p = F.perm
        Ld = np.linalg.cholesky(Md[np.ix_(p, p)])
        assert np.allclose(F.diag(), np.diag(Ld), rtol=1e-9)
        assert np.linalg.norm(F.solve(b) - xd) <= 1e-9 * np.linalg.norm(xd)
        assert F.info("nnzL") >= F.info("nnzL_true")
        if not relax:
            assert F.info("nnzL") == F.info("nnzL_true") + sum(c * (c - 1) // 2 for c in np.diff(F.array("sfirst")))
    bad = QL.copy(); bad.data = bad.data.copy()
    F = supernodal.SupernodalFactor(bad, ordering="natural")
    assert F.factorize(bad.data, 0.0, sd - 2.0 * abs(sd).max()) == 0
    nanv = bad.data.copy(); nanv[3] = np.nan
    assert F.factorize(nanv) == 0
