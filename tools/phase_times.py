"""Quick CUDA-event phase times of one workload (form / factor / direction / solve pair), without the
rest of bench.py:  python tools/phase_times.py [workload] [KEY=VALUE ...]   (needs a GPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import __graft_entry__ as g  # noqa: E402


def main():
    import torch
    wl = sys.argv[1] if len(sys.argv) > 1 and "=" not in sys.argv[1] else bench.DEFAULT_WORKLOAD
    opts = [a for a in sys.argv[1:] if "=" in a]
    pkg = g.package()
    gen, kw = bench.WORKLOADS[wl][0], bench.WORKLOADS[wl][1]
    prob = getattr(g.problems(), gen)(**kw)
    torch.cuda.set_device(0)
    inst = bench.Instance(pkg, torch, prob, 0, opts)
    nf, delta = inst.e2e_step()
    inst.make_resident()
    h = inst.h
    reps = 5
    ph = {"form_ms": bench.timed_events(torch, h.form_resident, reps),
          "factor_ms": bench.timed_events(torch, lambda: h.delta_loop_resident(*inst.dl_args), reps) / max(nf, 1),
          "direction_ms": bench.timed_events(torch, lambda: h.direction_resident(bench.N_REFINE), reps),
          "solve_pair_ms": bench.timed_events(torch, lambda: h.solve_resident(1), 4 * reps)}
    cnt = bench.algorithmic_counts(h)
    print(wl, opts, "num_fac", nf, {k: round(v, 4) for k, v in ph.items()},
          "solve GB/s %.0f" % (cnt["B_solve"] / ph["solve_pair_ms"] / 1e6),
          "factor TFLOP/s %.2f" % (cnt["F_chol"] / ph["factor_ms"] / 1e9))


if __name__ == "__main__":
    main()
