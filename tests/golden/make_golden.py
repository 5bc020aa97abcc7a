"""Mints the golden fixtures in tests/golden/*.npz from the CPU oracle
(oracle/kkt_oracle.c) on seeded inputs.  The reference holds no numeric goldens
for this path and cannot be executed here (no Julia / CHOLMOD), so these freeze the
restated oracle: "parity unpinned" with respect to CHOLMOD itself.

    python tests/golden/make_golden.py        # rewrites the fixtures
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if os.path.dirname(HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(HERE))
import problems  # noqa: E402  (tests/problems.py)

CASES = {
    "toy_lp1": ("toy", dict(name="toy_lp1"), 0.0),
    "toy_lp5": ("toy", dict(name="toy_lp5"), 0.0),
    "readme": ("toy", dict(name="readme"), 0.0),
    "toy_lp1_indef": ("toy", dict(name="toy_lp1", h_scale=-50.0), 0.0),
    "chain_nh30": ("chain", dict(nh=30, seed=11), 0.0),
    "chain_nh30_indef_warm": ("chain", dict(nh=30, seed=12, offdiag_curv=25.0), 0.0025),
    "sparse_qp_500": ("sparse_qp", dict(n=500, m_gen=250, win=5, seed=13), 0.0),
    "elec_np15": ("elec", dict(n_p=15, seed=14), 0.0),
    "chain_nh30_indef_cold": ("chain", dict(nh=30, seed=16, offdiag_curv=25.0), 0.0),
    "pde_N5": ("pde_control", dict(N=5, seed=15), 0.0),
}


def run_case(pkg, orc, case):
    gen, kw, delta_prev = case
    prob = getattr(problems, gen)(**kw)
    Q, sd = orc.form_system(prob.J, prob.H, prob.y, prob.s)
    QL = sp.tril(Q, format="csc"); QL.sort_indices()
    F = orc.Factor(QL)          # natural ordering: independent of the library's ordering code
    st, nf, delta, tried = F.delta_loop(QL.data, sd, delta_prev)
    out = dict(M_colptr=QL.indptr.astype(np.int64), M_rowval=QL.indices.astype(np.int64), M_nzval=QL.data,
               schur_diag=sd, status=np.array([1 if st == "success" else 0]), num_fac=np.array([nf]),
               delta=np.array([delta]), deltas_tried=tried)
    if st == "success":
        dx, dy, ds, err = F.direction(prob.J, prob.H, prob.y, prob.s, delta, *prob.rhs[0])
        out.update(dx=dx, dy=dy, ds=ds, kkt_err=err)
    return out


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg = g.package(); orc = g.oracle()
    for name, case in CASES.items():
        out = run_case(pkg, orc, case)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: np.asarray(v).shape for k, v in out.items()})
