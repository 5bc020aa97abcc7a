"""LDL' (sym == :symmetric) through the L1 entry point on the Schur complement of a workload: wall time of
ls_factor! (values up, numeric factorisation, inertia back) and of one ls_solve, tensor path vs the scalar path:
    python tools/ldlt_time.py [workload]      (needs a GPU)"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import __graft_entry__ as g  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c5_pde_60"
    pkg = g.package()
    gen, kw = bench.WORKLOADS[wl][0], bench.WORKLOADS[wl][1]
    prob = getattr(g.problems(), gen)(**kw)
    pars = pkg.Class_parameters()
    it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s, delta=0.0)
    k = pkg.pick_KKT_solver(pars)
    k.initialize(it)
    k.form_system(it)
    Mp, Mi = k._h.M_pattern()
    M = sp.csc_matrix((k._h.M_values(), Mi, Mp), shape=(prob.n, prob.n))
    k.finalize()
    b = prob.rhs[0][0]
    ref = None
    for scalar in (0, 1):
        for sym in ("definite", "symmetric"):
            if scalar and sym == "definite":
                continue
            ls = pkg.linear_solver_B200(sym)
            ls.initialize()
            ls._h.set_option("ldlt_scalar", scalar)
            ts = []
            for rep in range(3):
                t0 = time.perf_counter(); ok = ls.ls_factor(M, prob.n, 0); ts.append(time.perf_counter() - t0)
            t0 = time.perf_counter(); x = ls.ls_solve(b); tsol = time.perf_counter() - t0
            res = np.linalg.norm((M + sp.tril(M, -1).T) @ x - b) / np.linalg.norm(b)
            if ref is None:
                ref = x
            print("%s sym=%s scalar_path=%d inertia_ok=%d ls_factor %.1f ms (first %.1f) ls_solve %.1f ms  residual %.2e  |x - x_chol|/|x| %.2e"
                  % (wl, sym, scalar, ok, min(ts[1:]) * 1e3, ts[0] * 1e3, tsol * 1e3, res,
                     np.linalg.norm(x - ref) / np.linalg.norm(ref)), flush=True)
            ls.finalize()


if __name__ == "__main__":
    main()
