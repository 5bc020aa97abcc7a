"""Prints whether the device-side WHILE graph of the delta loop could be built on this box."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.package(); P = g.problems()
for prob in (P.chain(nh=300, seed=1, offdiag_curv=25.0), P.sparse_qp(20000, 10000, seed=1), P.pde_control(16, seed=1), P.elec(100, seed=1)):
    pars = pkg.Class_parameters()
    it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s)
    k = pkg.pick_KKT_solver(pars); k.initialize(it); k.form_system(it)
    l0 = pkg.launch_count()
    res = pkg.ipopt_strategy(it, k, pars)
    print(prob.name, res, "launches", pkg.launch_count() - l0, "n_big", k._h.info("n_big"))
    print("   active", k._h.info("loop_graph_active"), k._h.L.opb_last_error(k._h.h).decode())
    k.finalize()
