#!/usr/bin/env python
"""bench.py -- per-iteration KKT factor+solve on B200 (BASELINE.json's metric).

One "step" = one outer IPM iteration's linear algebra on one synthetic instance
(SURVEY.md 8d "unit of work", canonical F = as the delta loop decides, S = 2):
    form_system!  ->  ipopt_strategy! (delta loop, #fac attempts)  ->  2 x compute_direction!
      (each direction = 3 x [triangular solves + residual] + recovery + N err)

  value : ms per iteration with all inputs already resident in HBM (kernels only)
  e2e   : the same step through the reference-facing plugin API (host numpy
          buffers in, host buffers out: H2D of J/H values, y, s, rhs; D2H of
          schur_diag, dx, dy, ds, N err) -- the headline against --impl reference
  --impl reference : the CPU restatement of the reference path (oracle/) on the
          box's host cores, on a bounded sample of the same workload.

Multi-GPU (torchrun): the path shards by instance -- one independent solve per GPU,
no data-path collective ("replicas", scaling = weak).  value = wall ms / (N * K).  The same
line carries `sharded_single_instance`: ONE instance whose elimination-tree subtrees are mapped
to the N GPUs (top separators pull their children's update blocks over NVLink).
`--shard` makes that the headline instead (scaling = strong, value = ms per iteration).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

WORKLOADS = {
    # name: (generator, kwargs, cpu-sample kwargs)
    "c3_sparse_qp_n200k": ("sparse_qp", dict(n=200_000, m_gen=100_000), dict(n=50_000, m_gen=25_000)),
    "c2_chain_n100k": ("chain", dict(nh=25_000), dict(nh=25_000)),
    "c4_elec_n1200": ("elec", dict(n_p=400), dict(n_p=400)),
    "c5_pde_100": ("pde_control", dict(N=100), dict(N=56)),
    "c5_pde_60": ("pde_control", dict(N=60), dict(N=24)),
    "c5_pde_40": ("pde_control", dict(N=40), dict(N=24)),
    "c3_small": ("sparse_qp", dict(n=20_000, m_gen=10_000), dict(n=20_000, m_gen=10_000)),
}
DEFAULT_WORKLOAD = "c5_pde_100"
N_DIRECTIONS = 2
N_REFINE = 3


def make_problem(pkg, workload, seed, sample=False):
    gen, kw, kws = WORKLOADS[workload]
    return getattr(pkg.problems, gen)(seed=seed, **(kws if sample else kw))


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def algorithmic_counts(h):
    """compulsory bytes / flops of each stage (values only)."""
    g = h.info
    n, m = g("n"), g("m")
    nnzJ, nnzH, nnzM, nnzL = g("nnzJ"), g("nnzH"), g("nnzM"), g("nnzL_true")
    B_asm = 8 * (nnzJ + m + nnzH + nnzM)
    B_solve = 2 * 8 * nnzL + 4 * 8 * n
    B_res = 8 * (2 * nnzJ + 2 * nnzH) + 8 * (3 * m + 4 * n)
    B_dir = N_REFINE * B_solve + (N_REFINE - 1) * B_res + 8 * (2 * nnzJ + 6 * m + 2 * n)
    return dict(B_asm=B_asm, B_solve=B_solve, B_res=B_res, B_dir=B_dir, F_chol=g("flops"),
                B_fac=8 * (nnzM + nnzL))


def measure_fp64_peak(torch, dev):
    """cuBLAS DGEMM 8192^3 (MEASURED_PEAKS.json carries no FP64 figure): best of 5."""
    n = 8192
    a = torch.randn(n, n, device=dev, dtype=torch.float64)
    b = torch.randn(n, n, device=dev, dtype=torch.float64)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def cpu_iteration(orc, pkg, prob, hs):
    """One step of the CPU baseline: oracle/supernodal.py, the multifrontal restatement of the
    reference path with BLAS-3 dense kernels on all host cores (the performance class of the
    CHOLMOD supernodal factorisation behind julia.jl:34); assembly by oracle/kkt_oracle.c.  `hs` is a
    host-only handle holding the symbolic analysis; its cost is added per step by the caller
    (the reference analyses on every call: linear_solver_recycle=false)."""
    import scipy.sparse as sp
    from oracle import supernodal
    t0 = time.perf_counter()
    Q, sd = orc.form_system(prob.J, prob.H, prob.y, prob.s)
    QL = sp.tril(Q, format="csc"); QL.sort_indices()
    t1 = time.perf_counter()
    F = supernodal.SupernodalFactor(QL, hs)
    st, nf, delta, _ = F.delta_loop(QL.data, sd, prob.delta_prev)
    t2 = time.perf_counter()
    for r in prob.rhs[:N_DIRECTIONS]:
        F.direction(prob.J, prob.H, prob.y, prob.s, delta, *r, n_refine=N_REFINE)
    t3 = time.perf_counter()
    return dict(form_ms=(t1 - t0) * 1e3, factor_ms=(t2 - t1) * 1e3, direction_ms=(t3 - t2) * 1e3,
                total_ms=(t3 - t0) * 1e3, num_fac=nf)


CPU_SAMPLE_NOTE = ("oracle/supernodal.py: multifrontal Cholesky with LAPACK/BLAS-3 fronts (dpotrf, dtrsm, dsyrk on all host "
                   "cores), C supernodal solves, scipy products; symbolic analysis redone per call like the reference "
                   "(recycle=false)")


def measure_sharded(pkg, torch, dist, args, local, world, workload):
    """ONE instance of `workload` over all `world` GPUs: subtrees of the elimination
    tree mapped to ranks, top separators pulling their children's update blocks over NVLink.
    Every rank makes the same calls with the same data; time = max over ranks (CUDA events)."""
    prob = make_problem(pkg, workload, seed=0)
    pars = pkg.Class_parameters(device=local)
    it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s, delta=prob.delta_prev)
    k = pkg.pick_KKT_solver(pars, shard=pkg.DistShard())
    k.initialize(it)
    stream = torch.cuda.current_stream()
    k._h.set_stream(stream.cuda_stream)
    for kv in args.opt:
        key, val = kv.split("=")
        k._h.set_option(key, float(val))
    k.form_system(it)
    h = k._h
    d = pars.delta
    dl_args = (prob.delta_prev, d.zero, d.min, d.max, d.start, d.inc, d.dec, 500)
    rhs = [pkg.System_rhs(*r) for r in prob.rhs[:N_DIRECTIONS]]

    def e2e_step():
        k.form_system(it)
        pkg.ipopt_strategy(it, k, pars)
        for r in rhs:
            k.kkt_associate_rhs(it, r)
            k.compute_direction()

    def resident_step():
        h.form_resident()
        h.delta_loop_resident(*dl_args)
        for _ in range(N_DIRECTIONS):
            h.direction_resident(N_REFINE)

    def timed(fn, reps, warm):
        for _ in range(warm):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        dist.barrier(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    steps = max(2, min(args.steps, 5))
    e2e_step()
    t0 = time.perf_counter()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    dist.barrier(); torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    h.upload_values(prob.J.data, prob.H.data, prob.y, prob.s)
    h.upload_rhs(*prob.rhs[0])
    ms = timed(resident_step, steps, 2)
    fac = timed(lambda: h.delta_loop_resident(*dl_args), 2, 0)
    sol = timed(lambda: h.solve_resident(1), 3, 1)
    delta_res, nf_res, st_res, kkt_err = h.sync_state()
    t = torch.tensor([ms, e2e_ms, fac, sol], device=torch.device("cuda", local), dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    loads = torch.zeros(world, device=torch.device("cuda", local), dtype=torch.float64)
    loads[dist.get_rank()] = h.info("shard_load")
    dist.all_reduce(loads)
    out = {"workload": workload, "n_gpus": world, "ms_per_iter": float(t[0]), "e2e_ms_per_iter": float(t[1]),
           "factor_ms": float(t[2]) / max(nf_res, 1), "solve_pair_ms": float(t[3]), "num_fac": nf_res,
           "N_err": float(kkt_err[5]), "scaling": "strong",
           "rank_flops": [float(v) for v in loads], "top_flops": h.info("shard_top_flops"),
           "barrier_levels": int(h.info("shard_barriers")),
           "how": "subtree-to-GPU mapping by factorisation flops; update blocks, forward update vectors and the "
                  "solution cross GPUs through peer-mapped HBM (CUDA IPC over NVLink) inside the consuming kernels; "
                  "flag barriers on the stream; no collective in the data path"}
    k.finalize()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = graft.package(); orc = graft.oracle()
    prob = make_problem(pkg, args.workload, seed=0, sample=True)
    hs = pkg.Handle(-1)
    t0 = time.perf_counter()
    hs.set_structure(prob.n, prob.m, prob.J.indptr, prob.J.indices, prob.H.indptr, prob.H.indices, 0)
    t_order = time.perf_counter() - t0
    times = []
    for i in range(args.warmup + args.steps):
        r = cpu_iteration(orc, pkg, prob, hs)
        if i >= args.warmup:
            times.append(r["total_ms"] + t_order * 1e3)
    ms = float(np.mean(times))
    gen, kw, kws = WORKLOADS[args.workload]
    sample = "%s%s: %s; analysis %.0f ms per call included" % (gen, kws, CPU_SAMPLE_NOTE, t_order * 1e3)
    cores = os.cpu_count()
    out = {"metric": "kkt_factor_solve_ms_per_iter", "value": ms, "unit": "ms/iter", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
           "config": {"workload": args.workload, "sample": kws, "directions_per_iter": N_DIRECTIONS,
                      "refine": N_REFINE},
           "cpu_baseline": {"value": ms, "unit": "ms/iter", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": ms, "unit": "ms/iter", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the kernels-only timings of the other BASELINE shapes")
    ap.add_argument("--phase-repeat", type=int, default=3)
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="library tuning option passed to opb_set_option (e.g. outer_block=1024)")
    ap.add_argument("--shard", action="store_true",
                    help="N > 1: one instance sharded over the N GPUs (strong scaling) instead of N replicas")
    ap.add_argument("--ncu", action="store_true",
                    help="profiling run: resident steps only (no e2e, phase or CPU legs); never a bench value")
    args = ap.parse_args()
    if args.impl == "b200" and not args.ncu:
        args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the KKT path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = graft.package()
    lib0 = pkg.launch_count()

    sharded = bool(args.shard and world > 1)
    # replicas: one independent instance per GPU;  --shard: the SAME instance on every rank
    prob = make_problem(pkg, args.workload, seed=0 if sharded else rank)
    # the e2e leg copies its inputs from PINNED host memory every step
    def pinned(a):
        t = torch.empty(a.shape[0], dtype=torch.float64, pin_memory=True)
        v = t.numpy(); v[:] = a
        return v
    import scipy.sparse as sp
    prob.J = sp.csc_matrix((pinned(prob.J.data), prob.J.indices, prob.J.indptr), shape=prob.J.shape)
    prob.H = sp.csc_matrix((pinned(prob.H.data), prob.H.indices, prob.H.indptr), shape=prob.H.shape)
    prob.y = pinned(prob.y); prob.s = pinned(prob.s)
    prob.rhs = [tuple(pinned(v) for v in r) for r in prob.rhs]
    pars = pkg.Class_parameters(device=local)
    it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s, delta=prob.delta_prev)
    k = pkg.pick_KKT_solver(pars, shard=pkg.DistShard() if sharded else None)
    k.initialize(it)
    stream = torch.cuda.current_stream()
    k._h.set_stream(stream.cuda_stream)
    for kv in args.opt:
        key, val = kv.split("=")
        k._h.set_option(key, float(val))
    t0 = time.perf_counter()
    k.form_system(it)                                           # symbolic analysis happens here, once
    t_symbolic = time.perf_counter() - t0
    h = k._h
    cnt = algorithmic_counts(h)
    rhs = [pkg.System_rhs(*r) for r in prob.rhs[:N_DIRECTIONS]]
    d = pars.delta
    dl_args = (prob.delta_prev, d.zero, d.min, d.max, d.start, d.inc, d.dec, 500)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- e2e: plugin API with host buffers ----------------
    def e2e_step():
        k.form_system(it)
        st, nf, delta = pkg.ipopt_strategy(it, k, pars)
        for r in rhs:
            k.kkt_associate_rhs(it, r)
            k.compute_direction()
        return nf, delta

    e2e_ms = float("nan")
    if not args.ncu:
        for _ in range(args.warmup):
            nf, delta = e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            nf, delta = e2e_step()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    n, m = prob.n, prob.m
    h2d = 8 * (prob.J.nnz + prob.H.nnz + 2 * m) + N_DIRECTIONS * 8 * (n + 2 * m)
    d2h = 8 * n + 8 + N_DIRECTIONS * (8 * (n + 2 * m) + 48) + 3 * 8

    # ---------------- value: inputs resident in HBM ----------------
    h.upload_values(prob.J.data, prob.H.data, prob.y, prob.s)
    h.upload_rhs(*prob.rhs[0])

    def resident_step():
        h.form_resident()
        h.delta_loop_resident(*dl_args)
        for _ in range(N_DIRECTIONS):
            h.direction_resident(N_REFINE)

    for _ in range(args.warmup):
        resident_step()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    l0 = pkg.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        resident_step()
    e1.record()
    barrier()
    launches = pkg.launch_count() - l0
    clocks = sampler.stop()
    ms_step = e0.elapsed_time(e1) / args.steps
    delta_res, nf_res, st_res, kkt_err = h.sync_state()
    if world > 1:
        t = torch.tensor([ms_step, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms = float(t[0]), float(t[1])

    # ---------------- per-phase timing (rank 0) for the rooflines ----------------
    phases = {}
    if args.ncu:
        if rank == 0:
            print(json.dumps({"ncu_run": True, "workload": args.workload, "ms_per_step_under_profiler": ms_step,
                              "launches": int(launches), "num_fac": nf_res}))
        k.finalize()
        return
    if rank == 0 or sharded:        # a sharded instance needs every rank in every call
        def timed(fn, reps):
            fn(); torch.cuda.synchronize()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        reps = args.phase_repeat
        phases["form_ms"] = timed(h.form_resident, reps)
        phases["factor_ms"] = timed(lambda: h.delta_loop_resident(*dl_args), reps) / max(nf_res, 1)
        phases["direction_ms"] = timed(lambda: h.direction_resident(N_REFINE), reps)
        phases["solve_pair_ms"] = timed(lambda: h.solve_resident(1), reps)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        fp64_peak = measure_fp64_peak(torch, dev) * (world if sharded else 1)
        fac_tflops = cnt["F_chol"] / (phases["factor_ms"] * 1e-3) / 1e12
        asm_gbs = cnt["B_asm"] / (phases["form_ms"] * 1e-3) / 1e9
        solve_gbs = cnt["B_solve"] / (phases["solve_pair_ms"] * 1e-3) / 1e9
        dir_gbs = cnt["B_dir"] / (phases["direction_ms"] * 1e-3) / 1e9
        share = {kk: phases[kk] for kk in ("form_ms", "factor_ms", "direction_ms")}
        tot = share["form_ms"] + share["factor_ms"] * nf_res + share["direction_ms"] * N_DIRECTIONS
        if share["factor_ms"] * nf_res >= 0.5 * tot:
            roofline = {"bound": "tensor", "kernel": "numeric factorisation (all fronts of all levels)",
                        "achieved": fac_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                        "frac": fac_tflops / fp64_peak, "traffic": None,
                        "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (no FP64 entry in MEASURED_PEAKS.json)"}
        else:
            roofline = {"bound": "hbm", "kernel": "direction (3 x triangular solves + residuals)",
                        "achieved": dir_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": dir_gbs / hbm_peak,
                        "traffic": None, "peak_source": hbm_src}
        tr = {}
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_dram_traffic.json"))).get(args.workload, {})
        except Exception:
            pass
        roofline["traffic"] = tr.get("factor_bytes_per_attempt" if roofline["bound"] == "tensor" else "direction_bytes")
        # the dominant KERNEL of a factorisation-bound step: CUDA events around its launches in one
        # extra attempt (opb_profile_factor: plain launches, no look-ahead, so the timed kernels do
        # not overlap anything); achieved = algorithmic flops it serves / its summed launch time
        kernel_rooflines = {}
        if roofline["bound"] == "tensor" and not sharded:
            try:
                pf = h.profile_factor(delta_res)
                for kname, fl, ms, tkey in (("front_cb_kernel", pf["cb_flops"], pf["cb_ms"], "front_cb_bytes_per_attempt"),
                                            ("chol_panel_update_kernel", pf["update_flops"], pf["update_ms"],
                                             "panel_update_bytes_per_attempt")):
                    if ms > 0:
                        tf = fl / (ms * 1e-3) / 1e12
                        kernel_rooflines[kname] = {"bound": "tensor", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                                                   "frac": tf / fp64_peak, "traffic": tr.get(tkey),
                                                   "flops_per_attempt": fl, "ms_per_attempt": ms}
                kernel_rooflines["attempt_ms_plain_launches"] = pf["total_ms"]
                dom = max(("front_cb_kernel", "chol_panel_update_kernel"),
                          key=lambda kk: kernel_rooflines.get(kk, {}).get("ms_per_attempt", 0.0))
                if dom in kernel_rooflines:
                    phase_roofline = roofline
                    roofline = dict(kernel_rooflines[dom])
                    roofline["kernel"] = dom + " (all its launches of one factorisation attempt; per-attempt sums)"
                    roofline["peak_source"] = phase_roofline["peak_source"]
                    roofline["share_of_step"] = roofline["ms_per_attempt"] * max(nf_res, 1) / tot
            except Exception as e:      # the phase-level roofline stays
                kernel_rooflines = {"error": str(e)[:200]}
        extra_rooflines = {
            "assembly": {"bound": "hbm", "achieved": asm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": asm_gbs / hbm_peak},
            "factor": {"bound": "tensor", "achieved": fac_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fac_tflops / fp64_peak},
            "solve_pair": {"bound": "hbm", "achieved": solve_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": solve_gbs / hbm_peak},
            "direction": {"bound": "hbm", "achieved": dir_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": dir_gbs / hbm_peak},
        }

    # ---------------- the other BASELINE.json shapes, kernels only (rank 0, N = 1) ----------------
    others = {}
    if rank == 0 and world == 1 and not args.no_extra:
        for wname in ("c2_chain_n100k", "c3_sparse_qp_n200k", "c4_elec_n1200"):
            if wname == args.workload:
                continue
            try:
                p2 = make_problem(pkg, wname, seed=0)
                it2 = pkg.Class_iterate(p2.J, p2.H, p2.y, p2.s, delta=p2.delta_prev)
                k2 = pkg.pick_KKT_solver(pars); k2.initialize(it2)
                k2._h.set_stream(stream.cuda_stream)
                k2.form_system(it2)
                h2 = k2._h
                h2.upload_values(p2.J.data, p2.H.data, p2.y, p2.s); h2.upload_rhs(*p2.rhs[0])

                def step2():
                    h2.form_resident(); h2.delta_loop_resident(*dl_args)
                    for _ in range(N_DIRECTIONS):
                        h2.direction_resident(N_REFINE)
                for _ in range(3):
                    step2()
                torch.cuda.synchronize()
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    step2()
                b.record(); torch.cuda.synchronize()
                d2, nf2, st2, err2 = h2.sync_state()
                others[wname] = {"ms_per_iter": a.elapsed_time(b) / 5, "n": p2.n, "m": p2.m, "num_fac": nf2,
                                 "N_err": float(err2[5]), "factor_flops": h2.info("flops"),
                                 "nnz_L": int(h2.info("nnzL_true"))}
                k2.finalize()
            except Exception as e:      # never lose the headline line to an extra
                others[wname] = {"error": str(e)[:200]}

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            orc = graft.oracle()
            sp_prob = make_problem(pkg, args.workload, seed=0, sample=True)
            hs = pkg.Handle(-1)
            t0 = time.perf_counter()
            hs.set_structure(sp_prob.n, sp_prob.m, sp_prob.J.indptr, sp_prob.J.indices,
                             sp_prob.H.indptr, sp_prob.H.indices, 0)
            t_order = time.perf_counter() - t0
            r = cpu_iteration(orc, pkg, sp_prob, hs)
            # the GPU on the same sample, through the plugin API (host buffers)
            it_s = pkg.Class_iterate(sp_prob.J, sp_prob.H, sp_prob.y, sp_prob.s, delta=sp_prob.delta_prev)
            ks = pkg.pick_KKT_solver(pars); ks.initialize(it_s)
            rhs_s = [pkg.System_rhs(*q) for q in sp_prob.rhs[:N_DIRECTIONS]]

            def step_s():
                ks.form_system(it_s)
                pkg.ipopt_strategy(it_s, ks, pars)
                for q in rhs_s:
                    ks.kkt_associate_rhs(it_s, q); ks.compute_direction()
            step_s(); step_s()
            t0 = time.perf_counter()
            for _ in range(3):
                step_s()
            gpu_same = (time.perf_counter() - t0) * 1e3 / 3
            gen, kw, kws = WORKLOADS[args.workload]
            cpu = {"value": r["total_ms"] + t_order * 1e3, "unit": "ms/iter", "cores": os.cpu_count(), "kind": "port",
                   "sample": "%s%s, 1 iteration: %s; analysis %.0f ms included" % (gen, kws, CPU_SAMPLE_NOTE, t_order * 1e3),
                   "breakdown_ms": r, "gpu_e2e_same_sample_ms": gpu_same, "host_cores_available": os.cpu_count()}
            ks.finalize()
        except Exception as e:      # never lose the headline line to the baseline leg
            cpu = {"error": str(e)[:300]}

    if rank == 0:
        gen, kw, kws = WORKLOADS[args.workload]
        out = {
            "metric": "kkt_factor_solve_ms_per_iter", "value": ms_step if sharded else ms_step / world, "unit": "ms/iter",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": False, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "generator": gen, "generator_args": kw,
                       "n": n, "m": m, "nnz_J": int(prob.J.nnz), "nnz_M_lower": int(h.info("nnzM")),
                       "nnz_L": int(h.info("nnzL_true")), "factor_flops": cnt["F_chol"],
                       "supernodes": int(h.info("nsuper")), "etree_levels": int(h.info("nlevels")),
                       "max_front": int(h.info("max_front")),
                       "directions_per_iter": N_DIRECTIONS, "refine": N_REFINE, "num_fac": nf_res,
                       "delta": delta_res, "N_err": float(kkt_err[5]),
                       "instances": 1 if sharded else world,
                       "parallelism": ("one instance, elimination-tree subtrees mapped to %d GPUs, top separators over "
                                       "NVLink peer memory" % world) if sharded else
                                      "one independent instance per GPU (replicas)",
                       "l2_policy": "working set larger than L2: factor L alone is %.0f MB and is streamed by every "
                                    "factorisation and solve" % (8 * h.info("nnzL") / 1e6),
                       "symbolic_s_once": t_symbolic},
            "e2e": {"value": e2e_ms if sharded else e2e_ms / world, "unit": "ms/iter", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "phases_ms": phases,
            "roofline": roofline,
            "rooflines_by_phase": extra_rooflines,
            "rooflines_by_kernel": kernel_rooflines,
            "cpu_baseline": cpu,
            "other_workloads_kernels_only": others,
            "lib_launches_total": pkg.launch_count() - lib0,
        }
    k.finalize()
    del k, h
    torch.cuda.empty_cache()
    if world > 1 and not sharded and not args.no_extra:
        # the same line also carries ONE instance sharded over the N GPUs
        try:
            sh = measure_sharded(pkg, torch, dist, args, local, world, args.workload)
        except Exception as e:      # never lose the headline line to an extra
            sh = {"error": str(e)[:300]}
        if rank == 0:
            out["sharded_single_instance"] = sh
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
