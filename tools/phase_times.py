"""Quick CUDA-event phase times of one workload (form / factor / direction / solve pair), without the
rest of bench.py:

    python tools/phase_times.py [workload] [KEY=VALUE ...] [-- KEY=VALUE ...]   (needs a GPU)

Every `--` starts another option group measured in the same process on the same problem."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import __graft_entry__ as g  # noqa: E402


def measure(pkg, torch, wl, prob, opts, reps):
    inst = bench.Instance(pkg, torch, prob, 0, opts)
    nf, delta = inst.e2e_step()
    inst.make_resident()
    h = inst.h
    ph = {"form_ms": bench.timed_events(torch, h.form_resident, reps),
          "factor_ms": bench.timed_events(torch, lambda: h.delta_loop_resident(*inst.dl_args), reps) / max(nf, 1),
          "direction_ms": bench.timed_events(torch, lambda: h.direction_resident(bench.N_REFINE), reps),
          "solve_pair_ms": bench.timed_events(torch, lambda: h.solve_resident(1), 4 * reps)}
    cnt = bench.algorithmic_counts(h)
    print(wl, opts, "num_fac", nf, {k: round(v, 4) for k, v in ph.items()},
          "solve GB/s %.0f" % (cnt["B_solve"] / ph["solve_pair_ms"] / 1e6),
          "factor TFLOP/s %.2f" % (cnt["F_chol"] / ph["factor_ms"] / 1e9),
          "flops %.4e nnzL %.4e symbolic_s %.2f" % (cnt["F_chol"], h.info("nnzL_true"), inst.t_symbolic), flush=True)
    inst.close()
    del inst, h
    pkg.cache_clear()
    torch.cuda.empty_cache()


def main():
    import torch
    argv = sys.argv[1:]
    wl = argv.pop(0) if argv and "=" not in argv[0] and argv[0] != "--" else bench.DEFAULT_WORKLOAD
    groups, cur = [], []
    for a in argv:
        if a == "--":
            groups.append(cur); cur = []
        else:
            cur.append(a)
    groups.append(cur)
    reps = int(os.environ.get("OPB_PHASE_REPS", "5"))
    pkg = g.package()
    gen, kw = bench.WORKLOADS[wl][0], bench.WORKLOADS[wl][1]
    prob = getattr(g.problems(), gen)(**kw)
    torch.cuda.set_device(0)
    for opts in groups:
        measure(pkg, torch, wl, prob, opts, reps)


if __name__ == "__main__":
    main()
