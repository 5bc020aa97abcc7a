"""One instance sharded over several GPUs.

CPU part: the elimination-tree mapping (opb_shard_init + opb_set_structure on host-only handles)
-- every rank derives the same owner map, subtrees are private to one rank, cross-rank edges only
leave `top` supernodes, barriers are flagged on exactly the levels that need them -- and the same
bookkeeping through two gloo ranks (the all_gather that carries the IPC descriptors).

GPU part (one device is enough): `world` host threads drive `world` handles as virtual ranks; the
update blocks, forward update vectors and the solution cross the handles through the same
peer-pointer tables and flag barriers that CUDA IPC + NVLink serve between processes.  The result
must equal the single-handle result: same (status, #fac, delta), directions within 1e-12."""
import os
import socket
import sys
import numpy as np
import pytest

import problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _maps(pkg, prob, world):
    out = []
    for r in range(world):
        h = pkg.Handle(-1)
        h.shard_init(r, world)
        h.set_structure(prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)
        out.append(h)
    return out


@pytest.mark.parametrize("gen,kw", [("pde_control", dict(N=12)), ("sparse_qp", dict(n=6000, m_gen=3000)),
                                    ("chain", dict(nh=400)), ("elec", dict(n_p=12))])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_mapping_properties(pkg, gen, kw, world):
    prob = getattr(problems, gen)(seed=1, **kw)
    hs = _maps(pkg, prob, world)
    owner = hs[0].symbolic("owner"); top = hs[0].symbolic("top")
    sparent = hs[0].symbolic("sparent"); level = hs[0].symbolic("level")
    for h in hs[1:]:                                    # every rank derives the same map
        assert np.array_equal(h.symbolic("owner"), owner) and np.array_equal(h.symbolic("top"), top)
    assert owner.min() >= 0 and owner.max() < world
    ns = len(owner)
    barrier_levels = set()
    for s in range(ns):
        p = sparent[s]
        if p < 0:
            continue
        if top[s]:
            assert top[p], "ancestors of a top supernode are top supernodes"
        if owner[p] != owner[s]:
            assert top[p], "a cross-rank edge must enter a top supernode"
            barrier_levels.add(int(level[p]))
        if not top[p]:
            assert owner[p] == owner[s] and not top[s], "below the top the tree is private to one rank"
    # a rank takes part in the barrier of a level only when it is connected to a cross-rank edge there
    nb = [h.info("shard_barriers") for h in hs]
    assert all(0 <= b <= len(barrier_levels) for b in nb) and (not barrier_levels or max(nb) >= 1)
    for l in barrier_levels:
        assert sum(1 for h in hs if h.symbolic("level_mask")[l] != 0) >= 2
    # loads = flops of the (amalgamated) supernode panels each rank owns
    sf = hs[0].symbolic("sfirst"); rp = hs[0].symbolic("rowptr")
    c = np.diff(sf).astype(float); N = c + np.diff(rp)
    w = c * N * N - N * c * (c - 1) + (c - 1) * c * (2 * c - 1) / 6
    for r, h in enumerate(hs):
        assert h.info("shard_load") == pytest.approx(w[owner == r].sum(), rel=1e-9)
    assert hs[0].info("shard_top_flops") == pytest.approx(w[top == 1].sum(), rel=1e-9)
    # every rank's schedule holds exactly the supernodes it owns
    for r, h in enumerate(hs):
        assert h.info("n_tiny") + h.info("n_small") + h.info("n_big") == int((owner == r).sum())


@pytest.mark.parametrize("world", [2, 4, 8])
def test_split_fronts_are_top_fronts_with_a_rank_range(pkg, world):
    """ShardMap::split: only top supernodes with rows below the pivot block, the owner is the first rank of
    the range, non-top supernodes have the one-rank range of their owner; every rank derives the same flags."""
    prob = problems.pde_control(12, seed=1)
    seen = None
    for rank in range(world):
        h = pkg.Handle(-1)
        h.set_option("shard_split_flops", 0.0)
        h.shard_init(rank, world)
        h.set_structure(prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)
        owner = np.array(h.symbolic("owner")); top = np.array(h.symbolic("top")); split = np.array(h.symbolic("split"))
        ra = np.array(h.symbolic("range_a")); rb = np.array(h.symbolic("range_b"))
        r = np.diff(np.array(h.symbolic("rowptr")))
        assert np.all(top[split == 1] == 1) and np.all(r[split == 1] > 0)
        assert np.all(ra <= owner) and np.all(owner < rb) and np.all(rb <= world)
        assert np.all(ra[split == 1] == owner[split == 1]) and np.all((rb - ra)[split == 1] >= 2)
        assert np.all((rb - ra)[top == 0] == 1)
        # threshold 0: every top front with rows below its pivot block and a range of >= 2 ranks is split
        assert np.array_equal(split == 1, (top == 1) & (r > 0) & (rb - ra >= 2))
        cur = (owner.tolist(), split.tolist(), ra.tolist(), rb.tolist())
        assert seen is None or seen == cur
        seen = cur
        h.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_barrier_groups_are_consistent(pkg, world):
    """ShardMap::level_mask: at every level the ranks fall into groups that synchronise among themselves
    only -- every rank of a group holds the same mask, the owners of a supernode and of its children on
    other ranks share a group at the supernode's level, and so do the ranks of a split front's range."""
    prob = problems.pde_control(12, seed=1)
    masks, meta = [], None
    for rank in range(world):
        h = pkg.Handle(-1)
        h.set_option("shard_split_flops", 0.0)
        h.shard_init(rank, world)
        h.set_structure(prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)
        masks.append(np.array(h.symbolic("level_mask")))
        if meta is None:
            meta = [np.array(h.symbolic(k)) for k in ("owner", "sparent", "level", "split", "range_a", "range_b")]
        h.close()
    owner, par, level, split, ra, rb = meta
    nl = len(masks[0])
    for l in range(nl):
        for r in range(world):
            m = int(masks[r][l])
            if m == 0:
                continue
            assert (m >> r) & 1 and m != (1 << r)
            for q in range(world):
                if (m >> q) & 1:
                    assert int(masks[q][l]) == m, (l, r, q)
    for s in range(len(par)):
        p = par[s]
        if p >= 0 and owner[p] != owner[s]:
            m = int(masks[owner[p]][level[p]])
            assert (m >> owner[s]) & 1 and (m >> owner[p]) & 1
        if split[s]:
            m = int(masks[owner[s]][level[s]])
            assert all((m >> q) & 1 for q in range(ra[s], rb[s]))
    # private subtrees do not wait for anybody at the low levels
    assert all(int(masks[r][0]) == 0 for r in range(world))


def test_mapping_balances_subtrees(pkg):
    prob = problems.sparse_qp(20000, 10000, seed=0)
    hs = _maps(pkg, prob, 4)
    owner = hs[0].symbolic("owner"); top = hs[0].symbolic("top")
    sf = hs[0].symbolic("sfirst"); rp = hs[0].symbolic("rowptr")
    c = np.diff(sf).astype(float); N = c + np.diff(rp)
    w = c * N * N - N * c * (c - 1) + (c - 1) * c * (2 * c - 1) / 6
    sub = np.array([w[(owner == r) & (top == 0)].sum() for r in range(4)])
    assert sub.min() > 0 and sub.max() <= 1.5 * sub.mean(), sub
    assert w[top == 1].sum() < 0.8 * w.sum()


def test_unsharded_world_one(pkg):
    prob = problems.chain(nh=50, seed=0)
    h = pkg.Handle(-1)
    h.shard_init(0, 1)
    h.set_structure(prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)
    assert h.symbolic("owner").max() == 0 and h.info("shard_barriers") == 0
    with pytest.raises(pkg.OPBError):
        pkg.Handle(-1).shard_init(3, 2)
    with pytest.raises(pkg.OPBError):
        h.shard_init(0, 2)          # after the structure


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = g.package()
    prob = problems.pde_control(N=10, seed=0)            # the SAME instance on every rank
    shard = pkg.DistShard()
    h = pkg.Handle(-1)
    h.shard_init(shard.rank, shard.world)
    h.set_structure(prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)
    owner = h.symbolic("owner")
    blobs = shard.exchange(owner.tobytes())                   # the path the IPC descriptors take
    mine = int((owner == rank).sum())
    import hashlib
    out.put((rank, [hashlib.md5(b).hexdigest() for b in blobs], mine, len(owner), h.info("shard_load")))
    dist.destroy_process_group()


def test_two_gloo_ranks_agree_on_the_mapping():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, h0, m0, n0, l0), (r1, h1, m1, n1, l1) = res
    assert h0 == h1 and h0[0] == h0[1]          # both ranks computed and received the same owner map
    assert m0 + m1 == n0 == n1 and m0 > 0 and m1 > 0
    assert l0 > 0 and l1 > 0


# --------------------------------------------------------------------------- GPU
def _case(gen, world, *extra, timeout=240):
    """One sharded case in a fresh process (tests/shard_case.py explains why)."""
    import json
    import subprocess
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32")
    cmd = [sys.executable, os.path.join(ROOT, "tests", "shard_case.py"), gen, str(world)] + [str(e) for e in extra]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert lines, "no report (rc %d)\n%s\n%s" % (p.returncode, p.stdout[-2000:], p.stderr[-2000:])
    rep = json.loads(lines[-1])
    assert rep["ok"], rep
    return rep


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("gen,kw", [("chain", ["nh=300"]), ("sparse_qp", ["n=20000", "m_gen=10000"]),
                                    ("pde_control", ["N=16"])])
def test_virtual_ranks_match_single_gpu(gen, kw, world):
    rep = _case(gen, world, *kw)
    assert all(r["worst_rel_diff"] <= 1e-12 for r in rep["ranks"])


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])      # virtual ranks share one context: more handles than hardware queues stall in the barriers
@pytest.mark.parametrize("gen,kw", [("sparse_qp", ["n=20000", "m_gen=10000"]), ("pde_control", ["N=16"])])
def test_virtual_ranks_split_update_blocks(gen, kw, world):
    """The update blocks of the top separators formed by ALL ranks of the separator's range (the owner
    factorises the panel, the helpers pull it and store their tiles into the owner's arena): with the
    split threshold at zero every top front with rows below its pivot block is split; same delta
    sequence, directions <= 1e-12 against one GPU."""
    rep = _case(gen, world, *kw, "--opt", "shard_split_flops=0")
    if world >= 4:       # with two ranks the only top front may be the root (no update block)
        assert rep["split_fronts"] >= 1 and sum(rep["helped"]) >= 1, rep
    assert all(r["worst_rel_diff"] <= 1e-12 for r in rep["ranks"])


@pytest.mark.gpu
def test_virtual_ranks_delta_loop_failure_is_seen_by_every_rank():
    # indefinite Hessian: attempts fail on SOME rank (near the accepted delta only in the top
    # separator); every rank must take the same delta decisions (fail bit exchanged at the
    # barriers) and end with the same factor
    # (offdiag_curv: negative curvature the diagonal test of the delta rule cannot see -> x8 retries)
    rep = _case("chain", 2, "nh=300", "offdiag_curv=5.0")
    assert rep["ref"][0][1] > 3 and rep["ref"][0][2] > 0
    _case("chain", 3, "nh=300", "offdiag_curv=5.0", "--delta-prev", "1e-3")


@pytest.mark.gpu
def test_sharded_handle_requires_attached_peers(pkg):
    prob = problems.chain(nh=40, seed=0)
    h = pkg.Handle(0)
    h.shard_init(0, 2)
    h.set_structure(prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)
    with pytest.raises(pkg.OPBError):
        h.form(prob.J.data, prob.H.data, prob.y, prob.s)
    assert len(h.shard_export()) == 384


# --------------------------------------------------------------------------- several GPUs, one process each
def _mp_worker(rank, world, port, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)     # carries the IPC descriptors only
    try:
        pkg = g.package()
        report = []
        for gen, kw, neg in (("chain", dict(nh=300), 0.0), ("chain", dict(nh=300), 5.0),
                             ("sparse_qp", dict(n=20000, m_gen=10000), None), ("pde_control", dict(N=20), None)):
            prob = getattr(problems, gen)(seed=2, **kw) if neg is None else \
                getattr(problems, gen)(seed=2, offdiag_curv=neg, **kw)
            pars = pkg.Class_parameters(device=rank)

            def solve(shard):
                it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s, delta=0.0)
                k = pkg.pick_KKT_solver(pars, shard=shard)
                k.initialize(it)
                res = []
                for _ in range(2):                      # second pass replays the CUDA graphs
                    k.form_system(it)
                    st, nf, delta = pkg.ipopt_strategy(it, k, pars)
                    k.kkt_associate_rhs(it, pkg.System_rhs(*prob.rhs[0]))
                    k.compute_direction()
                    res.append((st, nf, delta, k.dir.x.copy(), k.dir.y.copy(), k.kkt_err_norm.ratio))
                k.finalize()
                return res
            ref = solve(None)
            got = solve(pkg.DistShard())
            for (st, nf, d, dx, dy, ne), (st0, nf0, d0, dx0, dy0, ne0) in zip(got, ref):
                rel = float(np.linalg.norm(dx - dx0) / np.linalg.norm(dx0))
                rely = float(np.linalg.norm(dy - dy0) / np.linalg.norm(dy0))
                report.append((gen, neg, (st, nf, d) == (st0, nf0, d0), nf0, rel, rely, ne, ne0))
            dist.barrier()
        out.put((rank, "ok", report))
    except Exception as e:      # noqa: BLE001
        out.put((rank, "error", repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_two_processes_two_gpus_match_single_gpu():
    """The real thing: one process per GPU, buffers mapped through CUDA IPC, data over NVLink."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=400) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank, status, report in res:
        assert status == "ok", (rank, report)
        assert len(report) == 8
        for gen, neg, same_delta_seq, nf0, rel, rely, ne, ne0 in report:
            assert same_delta_seq, (rank, gen, neg)
            assert rel <= 1e-12 and rely <= 1e-12, (rank, gen, rel, rely)
            if neg:
                assert nf0 > 1          # the delta loop had to retry, on every rank alike
