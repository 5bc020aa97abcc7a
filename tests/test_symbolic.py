"""CPU tests of the host-side symbolic analysis (csrc/symbolic.cpp) through a
host-only handle (device_id = -1): the gather map, ordering, supernodes,
extend-add maps and panel offsets that the CUDA kernels consume are validated by
emulating the device algorithms in numpy (tests/emulate.py) and comparing with the
oracle.  No numeric library call is made (there is no CPU fallback to call)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import emulate

import problems


def _handle(pkg, prob, **opts):
    h = pkg.Handle(-1)
    for k, v in opts.items():
        h.set_option(k, v)
    h.set_structure(prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)
    return h


def _problems(pkg):
    P = problems
    return [P.toy(nm) for nm in P.TOY_NAMES] + [
        P.chain(50, seed=1), P.sparse_qp(600, 300, win=5, seed=2), P.elec(20, seed=3), P.pde_control(6, seed=4)]


@pytest.mark.parametrize("opts", [{}, {"relax": 0}, {"nd_leaf": 8}, {"ordering": 1}])
def test_maps_reproduce_oracle(pkg, orc, opts):
    for prob in _problems(pkg):
        h = _handle(pkg, prob, **opts)
        S = emulate.Sym(h)
        assert sorted(S.perm.tolist()) == list(range(prob.n))
        # gather map -> tril(Q) bit for bit
        Mv = emulate.assemble_M_values(h, prob.J.indptr, prob.J.indices, prob.J.data, prob.H.data, prob.y, prob.s)
        Q, sd = orc.form_system(prob.J, prob.H, prob.y, prob.s)
        QL = sp.tril(Q, format="csc"); QL.sort_indices()
        assert np.array_equal(S.Mp, QL.indptr) or prob.H.nnz == 0   # LP: diagonal may be inserted
        M = sp.csc_matrix((Mv, S.Mi, S.Mp), shape=(prob.n, prob.n))
        assert abs(M - QL).max() == 0.0 if (M - QL).nnz else True
        assert np.all(S.Mi[S.Mp[:-1]] == np.arange(prob.n)), "diagonal must lead every column"
        # multifrontal emulation with the exported maps == oracle factor/solve
        delta = 0.5 + abs(min(sd.min(), 0.0)) * 2 + (300.0 if prob.name.startswith("elec") else 0.0)
        F = orc.Factor(QL, S.perm)
        ok_o = F.factorize(QL.data, sd + delta)
        ok, L = emulate.factor(S, Mv, delta)
        assert ok == bool(ok_o)
        if ok:
            b = prob.rhs[0][0]
            x = emulate.solve(S, L, b)
            xo = F.solve(b)
            assert np.linalg.norm(x - xo) <= 1e-10 * np.linalg.norm(xo)
            # same pivots as the simplicial oracle under the same permutation
            assert np.allclose(np.sort(L[S.dpos]), np.sort(F.diag()), rtol=1e-9)
            assert int(h.info("nnzL_true")) == F.lnz     # skeleton column counts are exact
        h.close()


def test_ldlt_emulation_inertia(pkg, orc):
    prob = problems.chain(20, seed=5, neg_curv=3.0)
    h = _handle(pkg, prob)
    S = emulate.Sym(h)
    Mv = emulate.assemble_M_values(h, prob.J.indptr, prob.J.indices, prob.J.data, prob.H.data, prob.y, prob.s)
    ok, L = emulate.factor(S, Mv, 0.0, mode="ldlt")
    Q, sd = orc.form_system(prob.J, prob.H, prob.y, prob.s)
    QL = sp.tril(Q, format="csc")
    F = orc.Factor(QL, S.perm)
    assert F.factorize(QL.data, mode="ldlt") == int(ok)
    if ok:
        d = L[S.dpos]
        do = F.diag()
        assert (d > 0).sum() == (do > 0).sum() and (d < 0).sum() == (do < 0).sum()
        b = prob.rhs[0][0]
        assert np.linalg.norm(emulate.solve(S, L, b, mode="ldlt") - F.solve(b)) <= 1e-9 * np.linalg.norm(F.solve(b))


def test_structure_invariants(pkg):
    prob = problems.sparse_qp(3000, 1500, seed=7)
    h = _handle(pkg, prob)
    S = emulate.Sym(h)
    ns = S.nsuper
    assert S.sfirst[0] == 0 and S.sfirst[-1] == prob.n and np.all(np.diff(S.sfirst) > 0)
    for s in range(ns):
        rows = S.rowidx[S.rowptr[s]:S.rowptr[s + 1]]
        assert np.all(np.diff(rows) > 0)
        if len(rows):
            assert rows[0] >= S.sfirst[s + 1]
            p = S.sparent[s]
            assert p > s and S.sfirst[p] <= rows[0] < S.sfirst[p + 1]
            assert S.level[p] > S.level[s]
            rel = S.rel[S.rowptr[s]:S.rowptr[s + 1]]
            assert np.all(np.diff(rel) > 0)
            # rel really points at the same global row in the parent's front
            pc = S.sfirst[p + 1] - S.sfirst[p]
            prow = np.concatenate([np.arange(S.sfirst[p], S.sfirst[p + 1]), S.rowidx[S.rowptr[p]:S.rowptr[p + 1]]])
            assert np.array_equal(prow[rel], rows)
            assert rel.max() < pc + (S.rowptr[p + 1] - S.rowptr[p])
        else:
            assert S.sparent[s] == -1
    # panels tile the L storage exactly, maps are injective
    c = np.diff(S.sfirst); r = np.diff(S.rowptr)
    assert np.array_equal(np.diff(S.Loff), ((c + r + 1) & ~1) * c)
    assert len(np.unique(S.amap)) == len(S.amap) and S.amap.min() >= 0 and S.amap.max() < S.Loff[-1]
    h.close()


def test_missing_structural_diagonal_is_inserted(pkg):
    # LP with an empty Hessian and a variable that appears in no constraint (gotcha 9.7-2)
    J = sp.csc_matrix(np.array([[1.0, 0.0, 2.0], [0.0, 0.0, 1.0]]))
    H = sp.csc_matrix((3, 3))
    h = pkg.Handle(-1)
    h.set_structure(3, 2, J.indptr, J.indices, H.indptr, H.indices, 0)
    Mp, Mi = h.symbolic("Mp"), h.symbolic("Mi")
    assert np.all(Mi[Mp[:-1]] == np.arange(3))
    pp = h.symbolic("pair_ptr")
    assert pp[Mp[1] + 1] - pp[Mp[1]] == 0      # column 1: diagonal present with no products


def test_one_based_indices_and_cache(pkg):
    prob = problems.chain(30, seed=9)
    h0 = _handle(pkg, prob)
    h1 = pkg.Handle(-1)
    h1.set_structure(prob.n, prob.m, prob.J.indptr + 1, prob.J.indices + 1, prob.H.indptr + 1, prob.H.indices + 1, 1)
    for nm in ("perm", "sfirst", "rowidx", "amap", "pairA", "pairB", "hmap"):
        assert np.array_equal(h0.symbolic(nm), h1.symbolic(nm)), nm
    # a second handle on the same pattern reuses the analysis (two solver objects per solve)
    h2 = _handle(pkg, prob)
    assert h2.info("symbolic_cached") == 1.0


def test_bad_patterns_are_rejected(pkg):
    h = pkg.Handle(-1)
    J = sp.csc_matrix(np.array([[1.0, 2.0], [3.0, 4.0]]))
    Hup = sp.csc_matrix(np.array([[1.0, 1.0], [0.0, 1.0]]))   # upper entry: not lower triangular
    with pytest.raises(pkg.OPBError) as e:
        h.set_structure(2, 2, J.indptr, J.indices, Hup.indptr, Hup.indices, 0)
    assert e.value.code == -1
    bad = J.indices.copy(); bad[0] = 5
    Hl = sp.csc_matrix(np.tril(np.ones((2, 2))))
    with pytest.raises(pkg.OPBError):
        h.set_structure(2, 2, J.indptr, bad, Hl.indptr, Hl.indices, 0)


def test_user_permutation(pkg, orc):
    N = 5
    prob = problems.pde_control(N, seed=1)
    perm = problems.grid_nd_perm(N, leaf=2)
    h = pkg.Handle(-1)
    h.set_permutation(perm)
    h.set_structure(prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)
    S = emulate.Sym(h)
    assert sorted(S.perm.tolist()) == list(range(prob.n))
    Mv = emulate.assemble_M_values(h, prob.J.indptr, prob.J.indices, prob.J.data, prob.H.data, prob.y, prob.s)
    ok, L = emulate.factor(S, Mv, 1e-3)
    assert ok
    with pytest.raises(pkg.OPBError):
        hb = pkg.Handle(-1)
        hb.set_permutation(np.zeros(prob.n, np.int64))
        hb.set_structure(prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)


def test_forward_gather_lists_are_the_transpose_of_rel(pkg):
    """gptr/gsrc/gch (forward-solve gather lists) against the child-by-child scatter through `rel`."""
    prob = problems.sparse_qp(3000, 1500, seed=4)
    h = pkg.Handle(-1)
    h.set_structure(prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)
    sfirst, sparent, rowptr, rel = (h.symbolic(k) for k in ("sfirst", "sparent", "rowptr", "rel"))
    gptr, gsrc, gch = h.symbolic("gptr"), h.symbolic("gsrc"), h.symbolic("gch")
    ns = len(sparent)
    assert len(gptr) == rowptr[-1] + prob.n + 1 and gptr[0] == 0
    rng = np.random.default_rng(0)
    u = rng.standard_normal(rowptr[-1])
    # reference: scatter every child's update vector into its parent, children ascending
    want = np.zeros(len(gptr) - 1)
    nsrc = 0
    for s in range(ns):
        p = sparent[s]
        if p < 0:
            continue
        gb = rowptr[p] + sfirst[p]
        for t in range(rowptr[s], rowptr[s + 1]):
            want[gb + rel[t]] += u[t]
            nsrc += 1
    assert gptr[-1] == nsrc == len(gsrc) == len(gch)
    got = np.array([u[gsrc[gptr[g]:gptr[g + 1]]].sum() for g in range(len(gptr) - 1)])
    assert np.allclose(got, want, rtol=0, atol=1e-12)
    for g in range(len(gptr) - 1):                       # ascending child order inside a destination
        ch = gch[gptr[g]:gptr[g + 1]]
        assert np.all(np.diff(ch) > 0)
    # every entry belongs to the child it claims
    assert np.all((gsrc >= rowptr[gch]) & (gsrc < rowptr[gch + 1]))


@pytest.mark.parametrize("gen,kw", [("sparse_qp", dict(n=4000, m_gen=2000)), ("pde_control", dict(N=12)),
                                    ("chain", dict(nh=300)), ("elec", dict(n_p=10))])
def test_update_block_storage_reuses_memory_without_overlap(pkg, gen, kw):
    """CBoff comes from a level-lifetime allocator: block s is live on the levels
    [level(s), level(parent(s))]; two blocks that are live on a common level must not overlap, and
    the arena must not be larger than the prefix-sum layout."""
    prob = getattr(problems, gen)(seed=3, **kw)
    h = pkg.Handle(-1)
    h.set_structure(prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)
    sparent, rowptr, level, cboff = (h.symbolic(k) for k in ("sparent", "rowptr", "level", "CBoff"))
    ns = len(sparent)
    r = np.diff(rowptr)
    total = int(h.info("cb_total"))
    assert cboff[ns] == total and total <= int((r * r + 1).sum())
    nlev = int(level.max()) + 1
    live = [[] for _ in range(nlev)]
    for s in range(ns):
        p = sparent[s]
        if p < 0 or r[s] == 0:
            continue
        assert cboff[s] >= 0 and cboff[s] + r[s] * r[s] <= total
        for l in range(level[s], level[p] + 1):
            live[l].append((int(cboff[s]), int(cboff[s] + r[s] * r[s])))
    for l in range(nlev):
        iv = sorted(live[l])
        for (a0, a1), (b0, b1) in zip(iv, iv[1:]):
            assert a1 <= b0, "update blocks live on level %d overlap" % l
    h.close()


def test_tile_cut_table(pkg):
    """tcut (symbolic.cpp): row position at which every update block crosses the 64-row tile
    boundaries of its parent's update block == a binary search in `rel` (what front_cb_kernel did
    per tile before the table existed)."""
    for prob in [problems.sparse_qp(600, 300, win=5, seed=2), problems.pde_control(9, seed=4), problems.elec(60, seed=3)]:
        h = _handle(pkg, prob)
        sfirst = np.array(h.symbolic("sfirst")); rowptr = np.array(h.symbolic("rowptr"))
        rel = np.array(h.symbolic("rel")); par = np.array(h.symbolic("sparent"))
        tp = np.array(h.symbolic("tcut_ptr")); tc = np.array(h.symbolic("tcut"))
        assert len(tp) == len(par) + 1 and tp[-1] == len(tc)
        for s in range(len(par)):
            p = par[s]
            if p < 0:
                assert tp[s + 1] == tp[s]
                continue
            c = sfirst[p + 1] - sfirst[p]; N = c + rowptr[p + 1] - rowptr[p]
            ce = c & ~1
            nt = (N - ce + 63) // 64
            assert tp[s + 1] - tp[s] == nt + 2
            want = np.searchsorted(rel[rowptr[s]:rowptr[s + 1]], ce + 64 * np.arange(nt + 2), side="left")
            assert np.array_equal(tc[tp[s]:tp[s + 1]], want)
        h.close()


def _cb_tiles(wm, n_minus_ce):
    """Mirror of cb_tiles (opb_internal.h): tiles of BM = 64 * wm rows x 128 columns over the lower triangle."""
    if wm == 2:
        nt = (n_minus_ce + 127) // 128
        return nt * (nt + 1) // 2
    n64 = (n_minus_ce + 63) // 64
    a = n64 >> 1
    return a * (a + 1) + ((a + 1) if (n64 & 1) else 0)


def _decode_wm1(tp):
    """Mirror of front_cb_kernel<1>'s tile decode: the row tiles 2a and 2a+1 hold a+1 tiles each."""
    a = int((np.sqrt(4.0 * tp + 1.0) - 1.0) * 0.5)
    while a * (a + 1) > tp:
        a -= 1
    while (a + 1) * (a + 2) <= tp:
        a += 1
    rem = tp - a * (a + 1)
    return (2 * a, rem) if rem <= a else (2 * a + 1, rem - (a + 1))


def _decode_update_wm1(tp, nrow):
    """Mirror of chol_panel_update_kernel<1>'s decode: column tile J (128 wide) starts at row tile 2J and
    J (nrow + 1) - J^2 tiles precede it."""
    b = nrow + 1.0
    disc = b * b - 4.0 * tp
    J = int((b - np.sqrt(disc)) * 0.5) if disc > 0 else nrow // 2
    while J > 0 and J * (nrow + 1) - J * J > tp:
        J -= 1
    while (J + 1) * (nrow + 1) - (J + 1) * (J + 1) <= tp and 2 * (J + 1) < nrow:
        J += 1
    off = tp - (J * (nrow + 1) - J * J)
    return 2 * J + off, J


@pytest.mark.parametrize("rows", [1, 63, 64, 65, 127, 128, 129, 200, 1000, 4097])
def test_tile_enumerations_cover_the_lower_triangle_once(rows):
    """The linear tile index of the 64-row-tile kernels (update blocks and panel updates) is a bijection onto the
    tiles that meet the lower triangle: every (row tile I, column tile J) with 64 I + 63 >= 128 J exactly once."""
    n64 = (rows + 63) // 64
    want = {(I, J) for I in range(n64) for J in range((rows + 127) // 128) if 64 * I + 63 >= 128 * J and 128 * J < rows}
    nt = _cb_tiles(1, rows)
    got = [_decode_wm1(t) for t in range(nt)]
    assert len(set(got)) == nt and set(got) == want
    # 128-row tiles: the plain triangle
    n128 = (rows + 127) // 128
    assert _cb_tiles(2, rows) == n128 * (n128 + 1) // 2
    # panel update over `ncol` column tiles of a panel with n64 row tiles below its first column
    for ncol in (1, 2, (rows + 127) // 128):
        tiles = sum(n64 - 2 * J for J in range(ncol) if 2 * J < n64)
        got = [_decode_update_wm1(t, n64) for t in range(tiles)]
        want_u = {(I, J) for J in range(ncol) if 2 * J < n64 for I in range(2 * J, n64)}
        assert len(set(got)) == tiles and set(got) == want_u


def test_analysis_does_not_depend_on_the_host_thread_count():
    """The one-off analysis runs its ordering candidates, the Schur pattern and the entry maps on host
    threads (symbolic.cpp: run_chunks, nd_rec); every array it produces must be the same for any number
    of threads -- the ranks of a sharded instance derive their maps independently and compare hashes."""
    import subprocess
    import sys
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "symbolic_digest.py")
    outs = []
    for nt in ("1", "3", "16"):
        env = dict(os.environ, OPB_HOST_THREADS=nt)
        r = subprocess.run([sys.executable, tool, "c3_small", "c5_pde_40"], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout)
    assert len(outs[0].splitlines()) > 40
    assert outs[0] == outs[1] == outs[2]


def test_analysis_phase_times_are_reported(pkg):
    prob = problems.pde_control(8, seed=0)
    h = pkg.Handle(-1)
    h.set_structure(prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)
    total = h.info("t_pattern") + h.info("t_analyze") + h.info("t_plan")
    assert total > 0
    parts = sum(h.info("t_" + k) for k in ("order_own", "order_candidates", "order_compare", "order", "etree_counts",
                                           "supernodes", "row_structures", "storage", "rel_gather", "tile_cuts", "amap"))
    assert 0 < parts <= h.info("t_analyze") * 1.001 + 1e-6
    assert h.info("t_upload") == 0          # host-only handle
    assert h.info("device_bytes") == 0      # ... which owns no device memory
