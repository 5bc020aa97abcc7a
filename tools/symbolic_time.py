"""Host time of the one-off symbolic analysis (opb_set_structure on a host-only handle), by phase.

    python tools/symbolic_time.py c3_sparse_qp_n200k [option=value ...]

Runs without a GPU: the analysis is host code.  Prints the info keys t_<phase> of the C ABI.
"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np          # noqa: E402
import scipy.sparse as sp   # noqa: E402

import bench                # noqa: E402

PHASES = ("pattern", "analyze", "order_own", "order_candidates", "order_compare", "order", "etree_counts", "supernodes",
          "row_structures", "storage", "rel_gather", "tile_cuts", "amap", "shard_map", "plan", "upload")


def main():
    name = sys.argv[1]
    opts = [a.split("=") for a in sys.argv[2:]]
    pkg = bench.graft.package()
    prob = bench.make_problem(name, seed=0)
    J = sp.csc_matrix(prob.J)
    H = sp.csc_matrix(sp.tril(prob.H))
    h = pkg.Handle(-1)
    for k, v in opts:
        h.set_option(k, float(v))
    t0 = time.perf_counter()
    h.set_structure(prob.n, prob.m, J.indptr.astype(np.int64), J.indices.astype(np.int64),
                    H.indptr.astype(np.int64), H.indices.astype(np.int64))
    t = time.perf_counter() - t0
    print("%s: n = %d, m = %d, opb_set_structure %.3f s, flops %.3e, nnz(L) %.3e, supernodes %d" %
          (name, prob.n, prob.m, t, h.info("flops"), h.info("nnzL_true"), h.info("nsuper")))
    for p in PHASES:
        v = h.info("t_" + p)
        if v > 0:
            print("  %-16s %8.3f s" % (p, v))


if __name__ == "__main__":
    main()
