"""GPU parity tests: the CUDA path, called through the C ABI (ctypes) and the
host-side mirror of the reference's plugin interface, against the CPU oracle on
identical seeded inputs.  Tolerances follow BASELINE.json's north_star:
assembly bit-exact; relative direction difference <= 1e-10; identical
delta/#fac sequence; N err within 10x of the oracle's."""
import numpy as np
import pytest
import scipy.sparse as sp

import problems

pytestmark = pytest.mark.gpu

REL_TOL = 1e-10   # BASELINE.json: "relative residual and direction difference <= 1e-10 relative"


def _own_perm(QL):
    """Fill-reducing ordering computed by the ORACLE (oracle/snode.c): exact minimum degree on
    small graphs, METIS_NodeND above -- independent of the product's symbolic analysis."""
    from oracle import supernodal
    return supernodal.order(QL, "mindeg" if QL.shape[0] <= 3000 else "metis")


def _oracle_factor(orc, QL, perm, big):
    if big:      # supernodal multifrontal oracle (BLAS-3): the sizes the scalar oracle cannot do in seconds
        from oracle import supernodal
        return supernodal.SupernodalFactor(QL, perm=perm)
    return orc.Factor(QL, perm)


def _oracle_iteration(orc, prob, delta_prev, perm=None, big=False, QL_sd=None):
    if QL_sd is None:
        Q, sd = orc.form_system(prob.J, prob.H, prob.y, prob.s)
        QL = sp.tril(Q, format="csc"); QL.sort_indices()
    else:
        QL, sd = QL_sd
    F = _oracle_factor(orc, QL, perm, big)
    st, nf, delta, tried = F.delta_loop(QL.data, sd, delta_prev)
    dirs = []
    if st == "success":
        for r in prob.rhs:
            dirs.append(F.direction(prob.J, prob.H, prob.y, prob.s, delta, *r))
    return QL, sd, st, nf, delta, tried, dirs, F


def _gpu_iteration(pkg, prob, delta_prev, opts=None):
    pars = pkg.Class_parameters()
    it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s, delta=delta_prev)
    k = pkg.pick_KKT_solver(pars)
    k.initialize(it)
    for key, v in (opts or {}).items():
        k._h.set_option(key, v)
    k.form_system(it)
    st, nf, delta = pkg.ipopt_strategy(it, k, pars)
    dirs = []
    if st == "success":
        for r in prob.rhs:
            k.kkt_associate_rhs(it, pkg.System_rhs(*r))
            k.compute_direction()
            e = k.kkt_err_norm
            dirs.append((k.dir.x.copy(), k.dir.y.copy(), k.dir.s.copy(),
                         np.array([e.error_D, e.error_P, e.error_mu, e.overall, e.rhs_norm, e.ratio])))
    return k, st, nf, delta, dirs


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _compare(pkg, orc, prob, delta_prev=0.0, opts=None, rel_tol=REL_TOL, big=False, own=True):
    """One outer iteration on the GPU against the oracle, twice: with the oracle factorising
    under the product's permutation (same elimination order: the PD decisions are comparable
    pivot by pivot) and under the oracle's OWN ordering (a bad permutation from the product
    cannot hide).  (status, #fac, delta) must be identical both times; dx, dy, ds within
    `rel_tol` (BASELINE: 1e-10 relative).  Where the two ORACLE runs themselves differ by more
    than that -- the two orderings bound what the conditioning of the system lets any correct
    FP64 solver reproduce -- the bar is twice that spread (reported in the failure message)."""
    k, st, nf, delta, dirs = _gpu_iteration(pkg, prob, delta_prev, opts)
    perm = k._h.symbolic("perm")
    QL, sd, st_o, nf_o, delta_o, tried, dirs_o, F = _oracle_iteration(orc, prob, delta_prev, perm, big)
    # assembly: bit-exact
    cp, ri = k._h.M_pattern()
    assert np.array_equal(cp, QL.indptr) and np.array_equal(ri, QL.indices)
    Mv = k._h.M_values()
    assert np.array_equal(Mv, QL.data), "assembly differs from the oracle: max %g" % np.abs(Mv - QL.data).max()
    assert np.array_equal(k.schur_diag, sd)
    assert k.diag_min() == sd.min()
    # delta loop
    assert (st, nf, delta) == (st_o, nf_o, delta_o), ((st, nf, delta), (st_o, nf_o, delta_o), tried)
    dirs_own = None
    if own:
        _, _, st_w, nf_w, delta_w, tried_w, dirs_own, _ = _oracle_iteration(orc, prob, delta_prev, _own_perm(QL), big, (QL, sd))
        assert (st, nf, delta) == (st_w, nf_w, delta_w), ("oracle with its own ordering", (st, nf, delta), (st_w, nf_w, delta_w), tried_w)
    for q, ((dx, dy, ds, err), (dxo, dyo, dso, erro)) in enumerate(zip(dirs, dirs_o)):
        for idx, (a, b, nm) in enumerate(((dx, dxo, "dx"), (dy, dyo, "dy"), (ds, dso, "ds"))):
            rel = _rel(a, b)
            bar = rel_tol
            if dirs_own is not None:
                spread = _rel(dirs_own[q][idx], b)
                bar = max(rel_tol, 2.0 * spread)
                assert _rel(a, dirs_own[q][idx]) <= max(rel_tol, 2.0 * spread), (prob.name, nm, "vs own-ordering oracle", spread)
            assert rel <= bar, (prob.name, nm, rel, "oracle spread between orderings", bar / 2.0)
        assert err[4] == pytest.approx(erro[4], rel=1e-14)
        assert err[5] <= max(10 * erro[5], 1e-13), (err[5], erro[5])      # a ratio of rounding residuals
    k.finalize()
    return nf, delta


@pytest.mark.parametrize("name", ["readme"] + ["toy_lp%d" % i for i in range(9)])
def test_toys(pkg, orc, name):
    _compare(pkg, orc, problems.toy(name))


def test_toy_with_hessian_needs_delta(pkg, orc):
    # indefinite H: the delta loop must raise delta exactly like the oracle
    p = problems.toy("toy_lp1", h_scale=-50.0)
    nf, delta = _compare(pkg, orc, p)
    assert nf > 1 and delta > 0


@pytest.mark.parametrize("nh", [7, 200, 2500])
def test_chain(pkg, orc, nh):
    _compare(pkg, orc, problems.chain(nh=nh, seed=nh))


def test_chain_indefinite_delta_sequence(pkg, orc):
    p = problems.chain(nh=300, seed=1, offdiag_curv=25.0)
    nf, delta = _compare(pkg, orc, p, delta_prev=0.0)
    assert nf >= 2
    nf2, delta2 = _compare(pkg, orc, p, delta_prev=delta)   # warm start from the previous delta
    assert nf2 >= 1


@pytest.mark.parametrize("n,m", [(600, 300), (5000, 2500), (20000, 10000)])
def test_sparse_qp(pkg, orc, n, m):
    _compare(pkg, orc, problems.sparse_qp(n, m, seed=n), big=n > 10000)


@pytest.mark.parametrize("n_p", [20, 100])
def test_elec_dense_front(pkg, orc, n_p):
    _compare(pkg, orc, problems.elec(n_p, seed=n_p))


@pytest.mark.parametrize("N", [4, 8, 14])
def test_pde_control(pkg, orc, N):
    _compare(pkg, orc, problems.pde_control(N, seed=N))


def test_options_do_not_change_results(pkg, orc):
    p = problems.sparse_qp(3000, 1500, seed=5)
    _compare(pkg, orc, p, opts={"relax": 0})
    _compare(pkg, orc, p, opts={"nd_leaf": 16})
    _compare(pkg, orc, p, opts={"ordering": 1})


@pytest.mark.parametrize("opts", [{"lookahead": 0}, {"lookahead": 1}, {"lookahead": 2, "outer_block": 256},
                                  {"cb_small_k": 0}, {"cb_small_k": 1}, {"graphs": 0}, {"loop_graph": 0}])
def test_schedule_options_do_not_change_results(pkg, orc, opts):
    """The numeric tuning options only change HOW the big fronts are scheduled (one stream / two streams / deep
    look-ahead with an outer block small enough to exercise every piece class, 64- vs 128-row tiles, graphs or
    plain launches): same delta sequence and directions <= 1e-10 against the oracle on a grid problem whose
    top fronts have several 128-column blocks."""
    _compare(pkg, orc, problems.pde_control(14, seed=2), opts=opts, own=False)


def test_profiling_entry_points(pkg):
    """opb_profile_factor / opb_profile_levels run one attempt like opb_factor: same factor (checked through a
    direction), phase times that add up, flops of the two tensor-pipe kernels reported."""
    prob = problems.pde_control(14, seed=2)
    pars = pkg.Class_parameters()
    it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s, delta=0.0)
    k = pkg.pick_KKT_solver(pars)
    k.initialize(it)
    k.form_system(it)
    st, nf, delta = pkg.ipopt_strategy(it, k, pars)
    assert st == "success"
    k.kkt_associate_rhs(it, pkg.System_rhs(*prob.rhs[0]))
    k.compute_direction()
    ref = k.dir.x.copy()
    h = k._h
    pf = h.profile_factor(delta)
    assert pf["inertia_ok"] == 1 and pf["total_ms"] > 0 and pf["cb_ms"] > 0 and pf["cb_flops"] > 0 and pf["update_flops"] >= 0
    k.compute_direction()
    assert np.linalg.norm(k.dir.x - ref) <= 1e-12 * np.linalg.norm(ref)
    pl = h.profile_levels(delta)
    T = pl["levels"]
    assert T.shape == (int(h.info("nlevels")), 3) and (T >= 0).all()
    assert T.sum() + pl["fill_ms"] + pl["trtri_ms"] <= pl["total_ms"] * 1.05 + 0.05
    assert T.sum() >= 0.5 * pl["total_ms"]
    k.compute_direction()
    assert np.linalg.norm(k.dir.x - ref) <= 1e-12 * np.linalg.norm(ref)
    # the factor and the update-block arena alone are 8 (nnzL + cb_total) bytes of device memory
    assert h.info("device_bytes") >= 8 * (h.info("nnzL") + h.info("cb_total"))
    assert h.info("t_upload") > 0 and h.info("t_analyze") > 0
    k.finalize()


def _full_size_properties(pkg, prob, expect_fac=None):
    """BASELINE.json sizes: the oracle is too slow here, so check size-independent
    properties of the result instead: the delta loop accepts the convex system at once,
    the direction satisfies the Newton system it claims to solve (relative residual in
    the inf norm, evaluated on the host with scipy from the inputs, not from device data)."""
    pars = pkg.Class_parameters()
    it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s, delta=0.0)
    k = pkg.pick_KKT_solver(pars)
    k.initialize(it)
    k.form_system(it)
    st, nf, delta = pkg.ipopt_strategy(it, k, pars)
    assert st == "success"
    if expect_fac is not None:
        assert nf == expect_fac, (nf, delta)
    sig = prob.y / prob.s
    Hs = prob.H + sp.tril(prob.H, -1).T
    for r in prob.rhs:
        k.kkt_associate_rhs(it, pkg.System_rhs(*r))
        k.compute_direction()
        dx, dy, ds = k.dir.x, k.dir.y, k.dir.s
        b = r[0] + prob.J.T @ (r[1] * sig + r[2] / prob.s)
        Mdx = prob.J.T @ (sig * (prob.J @ dx)) + Hs @ dx + delta * dx
        rel = np.abs(Mdx - b).max() / np.abs(b).max()
        assert rel <= 1e-9, (prob.name, "schur residual", rel)
        assert k.kkt_err_norm.ratio <= 1e-6, k.kkt_err_norm
        Jdx = prob.J @ dx
        assert np.allclose(ds, Jdx - r[1], rtol=1e-12, atol=1e-12 * np.abs(Jdx).max())
    k.finalize()


# ---- BASELINE.json configurations at FULL size, against the oracle (not only self-consistency)
def test_full_size_chain_c2(pkg, orc):
    """C2: chain n = 1e5 against the scalar oracle."""
    nf, _ = _compare(pkg, orc, problems.chain(nh=25000))
    assert nf == 1


def test_full_size_elec_dense_c4(pkg, orc):
    """C4: COPS elec n_p = 400 (n = 1200, dense front, indefinite Hessian: the delta loop runs)."""
    nf, delta = _compare(pkg, orc, problems.elec(400))
    assert nf >= 2 and delta > 0


def test_full_size_sparse_qp_c3(pkg, orc):
    """C3: sparse QP n = 2e5 against the supernodal oracle (own METIS ordering and the product's)."""
    nf, _ = _compare(pkg, orc, problems.sparse_qp(), big=True)
    assert nf == 1


def test_pde_56_c5_sample(pkg, orc):
    """C5 at 56^3 (the bounded CPU sample of bench.py) against the supernodal oracle."""
    nf, _ = _compare(pkg, orc, problems.pde_control(56), big=True)
    assert nf == 1


def test_mid_size_pde(pkg):
    _full_size_properties(pkg, problems.pde_control(40), expect_fac=1)


def test_full_size_pde_100_c5(pkg):
    """C5 at 100^3 (n = 1e6, the headline instance): no CPU oracle finishes this in test time, so the
    size-independent properties: convex system accepted at delta = 0 with one factorisation, the
    direction satisfies the Schur system and the recovery identities (host-evaluated)."""
    _full_size_properties(pkg, problems.pde_control(100), expect_fac=1)


def test_ipm_like_iterate_sequence(pkg, orc):
    """Six outer iterations on ONE solver object: same sparsity pattern (one symbolic analysis),
    drifting (J, y, s), every third iterate with an indefinite Hessian, the accepted delta threaded
    through as delta_prev like one_phase.jl:205-206.  The (status, #fac, delta) sequence must equal
    the oracle's and every direction must agree to 1e-10."""
    base = problems.chain(nh=500, seed=1)
    pars = pkg.Class_parameters()
    k = pkg.pick_KKT_solver(pars)
    k.initialize(pkg.Class_iterate(base.J, base.H, base.y, base.s))
    delta_prev, perm, seq = 0.0, None, []
    for t, prob in enumerate(problems.ipm_sequence(base, steps=6, seed=3, shift=0.0, offdiag=5.0)):
        it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s, delta=delta_prev)
        k.form_system(it)
        if perm is None:
            perm = k._h.symbolic("perm")
        else:
            assert np.array_equal(perm, k._h.symbolic("perm"))      # the analysis was reused
        st, nf, delta = pkg.ipopt_strategy(it, k, pars)
        Q, sd = orc.form_system(prob.J, prob.H, prob.y, prob.s)
        QL = sp.tril(Q, format="csc"); QL.sort_indices()
        F = orc.Factor(QL, perm)
        st_o, nf_o, delta_o, tried = F.delta_loop(QL.data, sd, delta_prev)
        assert (st, nf, delta) == (st_o, nf_o, delta_o), (t, (st, nf, delta), (st_o, nf_o, delta_o), tried)
        seq.append(nf)
        if st == "success":
            r = prob.rhs[0]
            k.kkt_associate_rhs(it, pkg.System_rhs(*r))
            k.compute_direction()
            dxo, dyo, dso, erro = F.direction(prob.J, prob.H, prob.y, prob.s, delta, *r)
            for a, b in ((k.dir.x, dxo), (k.dir.y, dyo), (k.dir.s, dso)):
                assert np.linalg.norm(a - b) <= REL_TOL * max(np.linalg.norm(b), 1e-300), t
        delta_prev = delta
    assert max(seq) > 1, "the sequence never exercised the delta loop: %r" % seq
    k.finalize()
