"""Digest of every array of the symbolic analysis (host only) -- to check that a change of the
analysis code (threading, data structures) leaves its RESULT bit-identical.

    python tools/symbolic_digest.py [workload ...] > before.txt   # then rebuild, rerun, diff
"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np          # noqa: E402
import scipy.sparse as sp   # noqa: E402

import bench                # noqa: E402

NAMES = ("perm", "sfirst", "sparent", "rowptr", "rowidx", "rel", "Loff", "CBoff", "level", "amap", "dpos", "Mp", "Mi",
         "pair_ptr", "pairA", "pairB", "hmap", "gptr", "gsrc", "gch", "tcut_ptr", "tcut")
SHARD = ("owner", "level_mask", "split", "range_a", "range_b", "top")


def digest(pkg, name, opts=(), world=1):
    prob = bench.make_problem(name, seed=0)
    J = sp.csc_matrix(prob.J)
    H = sp.csc_matrix(sp.tril(prob.H))
    out = []
    for rank in range(world):
        h = pkg.Handle(-1)
        if world > 1:
            h.shard_init(rank, world)
        for k, v in opts:
            h.set_option(k, float(v))
        h.set_structure(prob.n, prob.m, J.indptr.astype(np.int64), J.indices.astype(np.int64),
                        H.indptr.astype(np.int64), H.indices.astype(np.int64))
        for nm in NAMES + (SHARD if world > 1 else ()):
            out.append("%s %s w%d r%d %-9s %s" % (name, ",".join("%s=%s" % kv for kv in opts), world, rank, nm,
                                                   hashlib.md5(h.symbolic(nm).tobytes()).hexdigest()))
        out.append("%s flops %.17g cb_total %d nsuper %d" % (name, h.info("flops"), h.info("cb_total"), h.info("nsuper")))
    return out


def main():
    pkg = bench.graft.package()
    names = sys.argv[1:] or ["c3_small", "c5_pde_40", "c2_chain_n100k", "c4_elec_n1200"]
    for nm in names:
        for line in digest(pkg, nm):
            print(line)
        for line in digest(pkg, nm, opts=(("ordering", 0),)):
            print(line)
    if not sys.argv[1:]:
        for line in digest(pkg, "c5_pde_40", world=4):
            print(line)


if __name__ == "__main__":
    main()
