"""Helper of tests/test_shard.py: ONE sharded case in a fresh process.

`world` host threads drive `world` handles (virtual ranks) on one GPU and the result is compared
with the single-handle solve of the same instance.  A fresh process per case, with
CUDA_DEVICE_MAX_CONNECTIONS raised, keeps every handle's stream on its own hardware queue: inside
one CUDA context two streams that share a queue serialise, and a rank waiting in a barrier kernel
would then block the very peer it waits for (between processes, the real deployment, every rank
has its own context and queues).  Prints one JSON line.
    python tests/shard_case.py GEN WORLD [key=value ...] [--delta-prev X] [--iters K] [--opt LIBOPTION=VALUE]"""
import faulthandler
import json
import os
import sys
import threading

import numpy as np

faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems  # noqa: E402
import __graft_entry__ as g  # noqa: E402


def main():
    gen, world = sys.argv[1], int(sys.argv[2])
    kw, delta_prev, iters, opts = {}, 0.0, 2, {}
    args = sys.argv[3:]
    while args:
        a = args.pop(0)
        if a == "--delta-prev":
            delta_prev = float(args.pop(0))
        elif a == "--iters":
            iters = int(args.pop(0))
        elif a == "--opt":
            ok_, ov_ = args.pop(0).split("=")
            opts[ok_] = float(ov_)
        else:
            k, v = a.split("=")
            kw[k] = float(v) if "." in v else int(v)
    pkg = g.package()
    prob = getattr(problems, gen)(seed=2, **kw)
    pars = pkg.Class_parameters()

    def solve(shard):
        it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s, delta=delta_prev)
        k = pkg.pick_KKT_solver(pars, shard=shard)
        k.initialize(it)
        if shard is not None:
            k._h.set_option("barrier_timeout_s", 5.0)
            for ok_, ov_ in opts.items():
                k._h.set_option(ok_, ov_)
        out = []
        for _ in range(iters):
            k.form_system(it)
            st, nf, delta = pkg.ipopt_strategy(it, k, pars)
            dirs = []
            if st == "success":
                for rr in prob.rhs:
                    k.kkt_associate_rhs(it, pkg.System_rhs(*rr))
                    k.compute_direction()
                    dirs.append((k.dir.x.copy(), k.dir.y.copy(), k.dir.s.copy(), k.kkt_err_norm.ratio))
            out.append((st, nf, delta, dirs))
        return k, out

    k0, ref = solve(None)
    k0.finalize()
    shards = pkg.ThreadShard.make(world)
    res, keep, errors = [None] * world, [None] * world, []

    def work(r):
        try:
            keep[r], res[r] = solve(shards[r])
        except Exception as e:  # noqa: BLE001
            errors.append("rank %d: %r" % (r, e))
            try:
                shards[r].hub.barrier.abort()
            except Exception:
                pass

    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=90)
    report = {"gen": gen, "world": world, "errors": errors, "ref": [list(r[:3]) for r in ref], "ranks": []}
    if not errors and keep[0] is not None:
        report["split_fronts"] = int(keep[0]._h.info("shard_split"))
        report["helped"] = [int(k._h.info("shard_helped")) for k in keep if k is not None]
    ok = not errors and all(r is not None for r in res)
    if ok:
        for r in range(world):
            worst, same = 0.0, True
            for (st, nf, d, dirs), (st0, nf0, d0, dirs0) in zip(res[r], ref):
                same = same and (st, nf, d) == (st0, nf0, d0) and len(dirs) == len(dirs0)
                for (dx, dy, ds, ne), (dx0, dy0, ds0, ne0) in zip(dirs, dirs0):
                    for a, b in ((dx, dx0), (dy, dy0), (ds, ds0)):
                        worst = max(worst, float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)))
            report["ranks"].append({"same_delta_sequence": same, "worst_rel_diff": worst})
            ok = ok and same and worst <= 1e-12
    report["ok"] = bool(ok)
    print(json.dumps(report), flush=True)
    os._exit(0 if ok else 1)      # skip the destructors of possibly wedged handles


if __name__ == "__main__":
    main()
