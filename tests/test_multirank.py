"""Multi-GPU path = batches of instances, one independent solve per rank, no data-path
collective.  This covers the rank bookkeeping bench.py relies on with two
gloo ranks on CPU: every rank builds ITS OWN instance (seed = rank), runs the symbolic analysis
through the C ABI on a host-only handle, and the only collective is the MAX reduction of the
per-rank time.  No compute calls: there is no GPU here."""
import os
import socket
import sys

import numpy as np
import pytest

import problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = g.package()
    prob = problems.sparse_qp(2000, 1000, seed=rank)          # one instance per rank
    h = pkg.Handle(-1)
    h.set_structure(prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)
    nnzL = h.info("nnzL_true")
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)       # stand-in for the per-rank ms
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.tensor([nnzL], dtype=torch.float64))
    out.put((rank, float(t[0]), [float(v[0]) for v in gathered], float(np.abs(prob.J.data).sum())))
    dist.destroy_process_group()


def test_two_ranks_independent_instances():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, t0, g0, j0), (r1, t1, g1, j1) = res
    assert t0 == t1 == 2.0                      # MAX over ranks, as bench.py reports it
    assert g0 == g1 and len(g0) == 2            # every rank sees every rank's instance size
    assert j0 != j1                             # the instances differ (seed = rank): no shared data path
