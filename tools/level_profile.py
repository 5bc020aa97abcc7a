"""Per-level phase times of one factorisation attempt (opb_profile_levels) next to the level's
fronts and flops:  python tools/level_profile.py [workload] [KEY=VALUE ...]   (needs a GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import __graft_entry__ as g  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 and "=" not in sys.argv[1] else bench.DEFAULT_WORKLOAD
    opts = [a for a in sys.argv[1:] if "=" in a]
    pkg = g.package()
    gen, kw = bench.WORKLOADS[wl][0], bench.WORKLOADS[wl][1]
    prob = getattr(g.problems(), gen)(**kw)
    # under torchrun: ONE instance sharded over the ranks, every rank prints its own table
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, shard, dev = 0, None, 0
    if world > 1:
        import torch
        import torch.distributed as dist
        dev = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
        rank = dist.get_rank()
        shard = pkg.DistShard()
    pars = pkg.Class_parameters(device=dev)
    it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s, delta=0.0)
    k = pkg.pick_KKT_solver(pars, shard=shard)
    k.initialize(it)
    for o in opts:
        key, val = o.split("=")
        k._h.set_option(key, float(val))
    k.form_system(it)
    h = k._h
    sf = np.array(h.symbolic("sfirst")); rp = np.array(h.symbolic("rowptr")); lev = np.array(h.symbolic("level"))
    c = np.diff(sf).astype(float); r = np.diff(rp).astype(float); N = c + r
    for rep in range(2):
        P = h.profile_levels(0.0)
    T = P["levels"]
    if world > 1:
        import io
        import contextlib
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            report(wl, h, P, T, lev, c, r, N, "rank %d of %d: " % (rank, world))
        for q in range(world):
            dist.barrier()
            if q == rank:
                sys.stdout.write(buf.getvalue()); sys.stdout.flush()
        dist.barrier()
        dist.destroy_process_group()
        return
    report(wl, h, P, T, lev, c, r, N, "")


def report(wl, h, P, T, lev, c, r, N, tag):
    print(tag + "resident CTAs per SM of the 64-row-tile kernels:", h.info("occ_small_tiles"))
    print("workload %s  total %.2f ms  fill %.2f  trtri %.2f   (columns: ms before the big panels | big panels | update blocks)"
          % (wl, P["total_ms"], P["fill_ms"], P["trtri_ms"]))
    print("%3s %6s %7s %7s %10s %10s %8s %8s %8s %7s %7s" % ("lvl", "fronts", "max c", "max N", "panel Gf", "cb Gf", "pre", "panel", "cb", "pan TF", "cb TF"))
    for l in range(T.shape[0]):
        m = lev == l
        big = m & (N > 152)
        fpan = float(np.sum((N[big] ** 3 - r[big] ** 3) / 3 - r[big] ** 2 * c[big]))
        fcb = float(np.sum(r[big] ** 2 * c[big]))
        print("%3d %6d %7d %7d %10.1f %10.1f %8.3f %8.3f %8.3f %7.2f %7.2f" % (
            l, m.sum(), c[m].max(), N[m].max(), fpan / 1e9, fcb / 1e9, T[l, 0], T[l, 1], T[l, 2],
            fpan / max(T[l, 1], 1e-9) / 1e9, fcb / max(T[l, 2], 1e-9) / 1e9))
    print("sums: pre %.2f  panel %.2f  cb %.2f" % tuple(T.sum(axis=0)))


if __name__ == "__main__":
    main()
