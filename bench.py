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
  --impl reference : the CPU restatement of the reference path (oracle/snode.c +
          oracle/supernodal.py: own METIS ordering, own symbolic analysis redone every step
          like `linear_solver_recycle = false`, multifrontal Cholesky on the host BLAS with all
          cores) on the SAME workload at FULL size, for as many steps as the time budget
          allows (at least one).  That process never imports the product package.

The same line carries the other BASELINE.json shapes (C2 chain, C3 sparse QP, C4 elec) with
their own value / e2e / per-phase rooflines / full-size CPU baseline (`workloads`).

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
    # name: (generator in tests/problems.py, kwargs, kwargs of the bounded CPU sample of the b200 arm's cpu_baseline leg)
    "c3_sparse_qp_n200k": ("sparse_qp", dict(n=200_000, m_gen=100_000), dict(n=200_000, m_gen=100_000)),
    "c2_chain_n100k": ("chain", dict(nh=25_000), dict(nh=25_000)),
    "c4_elec_n1200": ("elec", dict(n_p=400), dict(n_p=400)),
    "c5_pde_100": ("pde_control", dict(N=100), dict(N=56)),
    "c5_pde_60": ("pde_control", dict(N=60), dict(N=40)),
    "c5_pde_40": ("pde_control", dict(N=40), dict(N=40)),
    "c3_small": ("sparse_qp", dict(n=20_000, m_gen=10_000), dict(n=20_000, m_gen=10_000)),
}
DEFAULT_WORKLOAD = "c5_pde_100"
OTHER_WORKLOADS = ("c2_chain_n100k", "c3_sparse_qp_n200k", "c4_elec_n1200")
N_DIRECTIONS = 2
N_REFINE = 3
REFERENCE_BUDGET_S = 200.0       # --impl reference: timed steps stop once this much wall time is spent


def make_problem(workload, seed, sample=False, override=None):
    gen, kw, kws = WORKLOADS[workload]
    args = dict(kws if sample else kw)
    if override:
        args.update(override)
    return getattr(graft.problems(), gen)(seed=seed, **args)


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


# ---------------------------------------------------------------------------
# CPU side (oracle/ only -- nothing of the product is used here)
# ---------------------------------------------------------------------------
CPU_NOTE = ("oracle/snode.c + supernodal.py: own METIS_NodeND ordering + elimination tree / column counts / relaxed "
            "supernodes redone every step (the reference analyses on every ls_factor!: recycle=false), multifrontal "
            "Cholesky with dpotrf/dtrsm/dsyrk fronts and dtrsv/dgemv solves on scipy's OpenBLAS, scipy sparse "
            "products, assembly by oracle/kkt_oracle.c")


def cpu_iteration(prob, threads=None):
    """One step of the CPU baseline (form_system! -> ipopt_strategy! -> 2 x compute_direction!)."""
    import scipy.sparse as sp
    from oracle import oracle as orc
    from oracle import supernodal
    with supernodal.blas_threads(threads):
        t0 = time.perf_counter()
        Q, sd = orc.form_system(prob.J, prob.H, prob.y, prob.s)
        QL = sp.tril(Q, format="csc"); QL.sort_indices()
        t1 = time.perf_counter()
        F = supernodal.SupernodalFactor(QL, ordering="metis")
        t2 = time.perf_counter()
        st, nf, delta, _ = F.delta_loop(QL.data, sd, prob.delta_prev)
        t3 = time.perf_counter()
        err = None
        for r in prob.rhs[:N_DIRECTIONS]:
            err = F.direction(prob.J, prob.H, prob.y, prob.s, delta, *r, n_refine=N_REFINE)[3]
        t4 = time.perf_counter()
    return dict(form_ms=(t1 - t0) * 1e3, analysis_ms=(t2 - t1) * 1e3, factor_ms=(t3 - t2) * 1e3,
                direction_ms=(t4 - t3) * 1e3, total_ms=(t4 - t0) * 1e3, num_fac=nf, delta=delta,
                N_err=float(err[5]) if err is not None else None, flops=F.info("flops"),
                nnz_L=int(F.info("nnzL_true")), factor_gflops=F.info("flops") * nf / max(t3 - t2, 1e-9) / 1e9)


def cpu_memory_estimate_bytes(workload_args, gen):
    """Host memory the full-size CPU step needs (factor + update-block stack + matrices): 36 GB peak
    measured for pde N = 100 (METIS fill: nnz(L) = 3.6e9 with the relaxation zeros), ~ N^4 scaling."""
    if gen != "pde_control":
        return 4e9
    return 36e9 * (workload_args["N"] / 100.0) ** 4 + 4e9


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    assert "onephase_jl_b200" not in sys.modules
    gen, kw, _ = WORKLOADS[args.workload]
    name, override, same_config = args.workload, None, True
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64e9
    if gen == "pde_control" and cpu_memory_estimate_bytes(kw, gen) > 0.85 * avail:
        for N in (80, 64, 56, 40):
            if N < kw["N"] and cpu_memory_estimate_bytes(dict(N=N), gen) <= 0.85 * avail:
                override = dict(N=N); name = "c5_pde_%d_sample" % N; same_config = False
                break
    prob = make_problem(args.workload, seed=0, override=override)
    cores = os.cpu_count() or 1
    budget = float(os.environ.get("OPB_REFERENCE_BUDGET_S", REFERENCE_BUDGET_S))
    t_start = time.perf_counter()
    times, runs, warm_done = [], [], 0
    # The first step always runs.  It is a warm-up step only when the whole schedule (W + K steps at
    # that cost) fits the budget; otherwise it is too expensive to discard and counts as the first
    # timed step, and further steps run while the budget lasts.
    t0 = time.perf_counter()
    r = cpu_iteration(prob, cores)
    first_s = time.perf_counter() - t0
    if args.warmup >= 1 and first_s * (args.warmup + args.steps) <= budget:
        warm_done = 1
        while warm_done < args.warmup:
            cpu_iteration(prob, cores); warm_done += 1
    else:
        times.append(r["total_ms"]); runs.append(r)
    while len(times) < args.steps and (not times or time.perf_counter() - t_start + 1.1 * first_s <= budget):
        r = cpu_iteration(prob, cores)
        times.append(r["total_ms"]); runs.append(r)
    ms = float(np.mean(times))
    last = runs[-1]
    sample = ("%s(%s), full size, %d timed step(s) of %d requested within a %.0f s budget: %s"
              % (gen, override or kw, len(times), args.steps, budget, CPU_NOTE))
    out = {"metric": "kkt_factor_solve_ms_per_iter", "value": ms, "unit": "ms/iter", "n_gpus": args.gpus,
           "steps": len(times), "warmup": warm_done, "steps_requested": args.steps, "warmup_requested": args.warmup,
           "ms_per_step": ms, "higher_is_better": False,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
           "same_config": same_config,
           "config": {"workload": name, "generator": gen, "generator_args": override or kw,
                      "n": int(prob.n), "m": int(prob.m), "nnz_J": int(prob.J.nnz),
                      "directions_per_iter": N_DIRECTIONS, "refine": N_REFINE, "num_fac": last["num_fac"],
                      "delta": last["delta"], "N_err": last["N_err"], "factor_flops": last["flops"], "nnz_L": last["nnz_L"],
                      "instances": 1, "parallelism": "one instance on the host cores (BLAS threads = %d)" % cores},
           "cpu_baseline": {"value": ms, "unit": "ms/iter", "cores": cores, "kind": "port", "sample": sample,
                            "breakdown_ms": {k: last[k] for k in ("form_ms", "analysis_ms", "factor_ms", "direction_ms")},
                            "factor_gflops": last["factor_gflops"]},
           "e2e": {"value": ms, "unit": "ms/iter", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ---------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------
def pin_problem(torch, prob):
    """The e2e leg copies its inputs from PINNED host memory every step (a Julia Vector{Float64} is
    pageable: `e2e_pageable` times the same step from ordinary numpy buffers)."""
    import scipy.sparse as sp

    def pinned(a):
        t = torch.empty(a.shape[0], dtype=torch.float64, pin_memory=True)
        v = t.numpy(); v[:] = a
        return v
    prob.J = sp.csc_matrix((pinned(prob.J.data), prob.J.indices, prob.J.indptr), shape=prob.J.shape)
    prob.H = sp.csc_matrix((pinned(prob.H.data), prob.H.indices, prob.H.indptr), shape=prob.H.shape)
    prob.y = pinned(prob.y); prob.s = pinned(prob.s)
    prob.rhs = [tuple(pinned(v) for v in r) for r in prob.rhs]
    return prob


def symbolic_breakdown(h):
    """Host seconds of the phases of the one-off analysis (info keys t_<phase> of the C ABI); what is left of
    `symbolic_s_once` is device allocation and the first assembly."""
    out = {}
    for key in ("pattern", "analyze", "plan", "upload"):
        try:
            out[key] = round(h.info("t_" + key), 4)
        except Exception:
            return None
    return out


class Instance:
    """One KKT instance bound to a solver object (the plugin API) on the current torch stream."""

    def __init__(self, pkg, torch, prob, local, opts=(), shard=None, own_stream=False):
        self.pkg, self.torch, self.prob = pkg, torch, prob
        self.pars = pkg.Class_parameters(device=local)
        self.it = pkg.Class_iterate(prob.J, prob.H, prob.y, prob.s, delta=prob.delta_prev)
        self.k = pkg.pick_KKT_solver(self.pars, shard=shard)
        self.k.initialize(self.it)
        if not own_stream:       # own_stream: the handle keeps its private non-blocking stream (instance batches)
            self.k._h.set_stream(torch.cuda.current_stream().cuda_stream)
        for kv in opts:
            key, val = kv.split("=")
            self.k._h.set_option(key, float(val))
        t0 = time.perf_counter()
        self.k.form_system(self.it)              # symbolic analysis happens here, once
        self.t_symbolic = time.perf_counter() - t0
        self.h = self.k._h
        self.symbolic_breakdown = symbolic_breakdown(self.h)
        self.rhs = [pkg.System_rhs(*r) for r in prob.rhs[:N_DIRECTIONS]]
        d = self.pars.delta
        self.dl_args = (prob.delta_prev, d.zero, d.min, d.max, d.start, d.inc, d.dec, 500)

    def e2e_step(self):
        k, it = self.k, self.it
        k.form_system(it)
        st, nf, delta = self.pkg.ipopt_strategy(it, k, self.pars)
        for r in self.rhs:
            k.kkt_associate_rhs(it, r)
            k.compute_direction()
        return nf, delta

    def make_resident(self):
        p = self.prob
        self.h.upload_values(p.J.data, p.H.data, p.y, p.s)
        self.h.upload_rhs(*p.rhs[0])

    def resident_step(self):
        h = self.h
        h.form_resident()
        h.delta_loop_resident(*self.dl_args)
        for _ in range(N_DIRECTIONS):
            h.direction_resident(N_REFINE)

    def bytes_per_step(self):
        p = self.prob
        n, m = p.n, p.m
        h2d = 8 * (p.J.nnz + p.H.nnz + 2 * m) + N_DIRECTIONS * 8 * (n + 2 * m)
        d2h = 8 * n + 8 + N_DIRECTIONS * (8 * (n + 2 * m) + 48) + 3 * 8
        return int(h2d), int(d2h)

    def close(self):
        self.k.finalize()


def timed_events(torch, fn, reps, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def phase_rooflines(torch, inst, reps, hbm_peak, fp64_peak, nf):
    """CUDA-event time of each phase and its algorithmic bytes / flops against the measured peaks."""
    h = inst.h
    cnt = algorithmic_counts(h)
    ph = {"form_ms": timed_events(torch, h.form_resident, reps),
          "factor_ms": timed_events(torch, lambda: h.delta_loop_resident(*inst.dl_args), reps) / max(nf, 1),
          "direction_ms": timed_events(torch, lambda: h.direction_resident(N_REFINE), reps),
          "solve_pair_ms": timed_events(torch, lambda: h.solve_resident(1), reps)}
    fac_tflops = cnt["F_chol"] / (ph["factor_ms"] * 1e-3) / 1e12
    fac_gbs = cnt["B_fac"] / (ph["factor_ms"] * 1e-3) / 1e9
    asm_gbs = cnt["B_asm"] / (ph["form_ms"] * 1e-3) / 1e9
    solve_gbs = cnt["B_solve"] / (ph["solve_pair_ms"] * 1e-3) / 1e9
    dir_gbs = cnt["B_dir"] / (ph["direction_ms"] * 1e-3) / 1e9
    rl = {
        "assembly": {"bound": "hbm", "achieved": asm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": asm_gbs / hbm_peak,
                     "algorithmic_bytes": cnt["B_asm"]},
        "factor": {"bound": "tensor", "achieved": fac_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                   "frac": fac_tflops / fp64_peak, "algorithmic_flops": cnt["F_chol"],
                   "hbm_view": {"achieved": fac_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": fac_gbs / hbm_peak,
                                "algorithmic_bytes": cnt["B_fac"]}},
        "solve_pair": {"bound": "hbm", "achieved": solve_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": solve_gbs / hbm_peak,
                       "algorithmic_bytes": cnt["B_solve"]},
        "direction": {"bound": "hbm", "achieved": dir_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": dir_gbs / hbm_peak,
                      "algorithmic_bytes": cnt["B_dir"]},
    }
    return ph, rl, cnt


def measure_other_workload(pkg, torch, wname, local, hbm_peak, fp64_peak, cpu=True, steps=5):
    """value / e2e / per-phase rooflines / full-size CPU baseline of one of the other BASELINE shapes."""
    prob = pin_problem(torch, make_problem(wname, seed=0))
    inst = Instance(pkg, torch, prob, local)
    for _ in range(3):
        inst.e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        inst.e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    inst.make_resident()
    ms = timed_events(torch, inst.resident_step, steps, warm=3)
    d2, nf2, st2, err2 = inst.h.sync_state()
    ph, rl, cnt = phase_rooflines(torch, inst, 3, hbm_peak, fp64_peak, nf2)
    h2d, d2h = inst.bytes_per_step()
    out = {"value": ms, "unit": "ms/iter", "e2e": {"value": e2e_ms, "unit": "ms/iter", "h2d_bytes_per_step": h2d,
                                                    "d2h_bytes_per_step": d2h},
           "n": prob.n, "m": prob.m, "num_fac": nf2, "delta": d2, "N_err": float(err2[5]),
           "factor_flops": inst.h.info("flops"), "nnz_L": int(inst.h.info("nnzL_true")),
           "phases_ms": ph, "rooflines_by_phase": rl, "symbolic_s_once": inst.t_symbolic,
           "symbolic_breakdown_s": inst.symbolic_breakdown}
    inst.close()
    if cpu:
        try:
            r = cpu_iteration(make_problem(wname, seed=0), os.cpu_count())
            out["cpu_baseline"] = {"value": r["total_ms"], "unit": "ms/iter", "cores": os.cpu_count(), "kind": "port",
                                   "sample": "full size, 1 iteration: " + CPU_NOTE, "breakdown_ms": r}
        except Exception as e:
            out["cpu_baseline"] = {"error": str(e)[:200]}
    return out


def measure_batched(pkg, torch, wname, local, nbatch, steps=5):
    """B independent instances of one shape INSIDE one GPU (north_star: "batches of instances"): B solver
    objects sharing one symbolic analysis (pattern cache), each on its own stream; a step enqueues
    every instance's form / delta loop (one graph launch: the loop runs on the device) / directions
    without waiting, then synchronises once.  Latency-bound shapes (C2: tiny fronts, C4: one dense
    front) overlap on the SMs.  value = wall ms per instance-iteration."""
    insts = []
    for b in range(nbatch):
        insts.append(Instance(pkg, torch, make_problem(wname, seed=b), local, own_stream=True))
        insts[-1].make_resident()

    def step():
        for i in insts:
            i.resident_step()
    for _ in range(3):
        step()
    for i in insts:
        i.h.sync_state()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    nfs = [i.h.sync_state()[1] for i in insts]
    ms = (time.perf_counter() - t0) * 1e3 / steps
    out = {"instances": nbatch, "ms_per_batch_step": ms, "ms_per_instance_iter": ms / nbatch, "num_fac": int(max(nfs)),
           "symbolic_shared": int(sum(i.h.info("symbolic_cached") for i in insts)),
           "how": "one handle per instance, shared symbolic bundle, private streams, host enqueues all instances then syncs once; wall clock"}
    for i in insts:
        i.close()
    return out


def measure_sharded(pkg, torch, dist, args, local, world, workload):
    """ONE instance of `workload` over all `world` GPUs: subtrees of the elimination
    tree mapped to ranks, top separators pulling their children's update blocks over NVLink.
    Every rank makes the same calls with the same data; time = max over ranks (CUDA events)."""
    prob = make_problem(workload, seed=0)
    inst = Instance(pkg, torch, prob, local, args.opt, shard=pkg.DistShard())
    h = inst.h

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
    inst.e2e_step()
    dist.barrier(); torch.cuda.synchronize()
    # correctness of the sharded solve, evaluated on the host from the INPUTS (rank 0): the last direction of the
    # step must satisfy the Schur system  (J' S J + H + delta I) dx = dual_r + J'(S primal_r + comp_r / s)
    schur_res = None
    if dist.get_rank() == 0:
        import scipy.sparse as sp
        r = prob.rhs[N_DIRECTIONS - 1]
        sig = prob.y / prob.s
        Hs = prob.H + sp.tril(prob.H, -1).T
        dx = inst.k.dir.x
        b = r[0] + prob.J.T @ (r[1] * sig + r[2] / prob.s)
        Mdx = prob.J.T @ (sig * (prob.J @ dx)) + Hs @ dx + float(inst.k._delta) * dx
        schur_res = float(np.abs(Mdx - b).max() / np.abs(b).max())
    t0 = time.perf_counter()
    for _ in range(steps):
        inst.e2e_step()
    dist.barrier(); torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    inst.make_resident()
    ms = timed(inst.resident_step, steps, 2)
    fac = timed(lambda: h.delta_loop_resident(*inst.dl_args), 2, 0)
    sol = timed(lambda: h.solve_resident(1), 3, 1)
    delta_res, nf_res, st_res, kkt_err = h.sync_state()
    t = torch.tensor([ms, e2e_ms, fac, sol], device=torch.device("cuda", local), dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    loads = torch.zeros(world, device=torch.device("cuda", local), dtype=torch.float64)
    loads[dist.get_rank()] = h.info("shard_load")
    dist.all_reduce(loads)
    out = {"workload": workload, "n_gpus": world, "ms_per_iter": float(t[0]), "e2e_ms_per_iter": float(t[1]),
           "factor_ms": float(t[2]) / max(nf_res, 1), "solve_pair_ms": float(t[3]), "num_fac": nf_res,
           "N_err": float(kkt_err[5]), "schur_residual_host_check": schur_res, "scaling": "strong",
           "split_fronts": int(h.info("shard_split")),
           "rank_flops": [float(v) for v in loads], "top_flops": h.info("shard_top_flops"),
           "barrier_levels": int(h.info("shard_barriers")),
           "how": "subtree-to-GPU mapping by factorisation flops; the update blocks of the top separators are formed by all "
                  "GPUs of the separator's range (panel pulled over NVLink, tiles stored into the owner's arena); children's "
                  "update blocks, forward update vectors and the solution cross GPUs through peer-mapped HBM (CUDA IPC over "
                  "NVLink), in bulk copies or inside the consuming kernels; flag barriers on the stream among the ranks that "
                  "exchange data at a level; no collective in the data path"}
    inst.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE shapes")
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

    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # instance batches: more concurrent streams
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
    prob = pin_problem(torch, make_problem(args.workload, seed=0 if sharded else rank))
    inst = Instance(pkg, torch, prob, local, args.opt, shard=pkg.DistShard() if sharded else None)
    h = inst.h
    cnt = algorithmic_counts(h)
    n, m = prob.n, prob.m

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- e2e: plugin API with host buffers ----------------
    e2e_ms = float("nan")
    if not args.ncu:
        for _ in range(args.warmup):
            nf, delta = inst.e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            nf, delta = inst.e2e_step()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    h2d, d2h = inst.bytes_per_step()

    # ---------------- value: inputs resident in HBM ----------------
    inst.make_resident()
    for _ in range(args.warmup):
        inst.resident_step()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    l0 = pkg.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        inst.resident_step()
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

    if args.ncu:
        if rank == 0:
            print(json.dumps({"ncu_run": True, "workload": args.workload, "ms_per_step_under_profiler": ms_step,
                              "launches": int(launches), "num_fac": nf_res}))
        inst.close()
        return

    # ---------------- per-phase timing (rank 0) for the rooflines ----------------
    phases, roofline, extra_rooflines, kernel_rooflines = {}, None, {}, {}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    fp64_peak_1 = measure_fp64_peak(torch, dev) if (rank == 0 or sharded) else None
    if rank == 0 or sharded:        # a sharded instance needs every rank in every call
        fp64_peak = fp64_peak_1 * (world if sharded else 1)
        phases, extra_rooflines, _ = phase_rooflines(torch, inst, args.phase_repeat, hbm_peak, fp64_peak, nf_res)
        fac_tflops = extra_rooflines["factor"]["achieved"]
        dir_gbs = extra_rooflines["direction"]["achieved"]
        tot = phases["form_ms"] + phases["factor_ms"] * nf_res + phases["direction_ms"] * N_DIRECTIONS
        if phases["factor_ms"] * nf_res >= 0.5 * tot:
            roofline = {"bound": "tensor", "kernel": "numeric factorisation (all fronts of all levels)",
                        "achieved": fac_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                        "frac": fac_tflops / fp64_peak, "traffic": None,
                        "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (no FP64 entry in MEASURED_PEAKS.json)"}
        else:
            roofline = {"bound": "hbm", "kernel": "direction (3 x triangular solves + residuals)",
                        "achieved": dir_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": dir_gbs / hbm_peak,
                        "traffic": None, "peak_source": hbm_src}
        tr = {}
        for fn in ("r2_dram_traffic.json", "r1_dram_traffic.json"):
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", fn))).get(args.workload, {})
                if tr:
                    break
            except Exception:
                pass
        roofline["traffic"] = tr.get("factor_bytes_per_attempt" if roofline["bound"] == "tensor" else "direction_bytes")
        roofline["traffic_source"] = tr.get("source")
        # the dominant KERNEL of a factorisation-bound step: CUDA events around its launches in one
        # extra attempt (opb_profile_factor: plain launches, no look-ahead, so the timed kernels do
        # not overlap anything); achieved = algorithmic flops it serves / its summed launch time
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
                    roofline["traffic_source"] = tr.get("source")
                    roofline["share_of_step"] = roofline["ms_per_attempt"] * max(nf_res, 1) / tot
            except Exception as e:      # the phase-level roofline stays
                kernel_rooflines = {"error": str(e)[:200]}

    # ---------------- e2e from pageable host memory (what a Julia caller has), rank 0 ----------------
    e2e_pageable = None
    if rank == 0 and world == 1:
        try:
            p2 = make_problem(args.workload, seed=rank)
            inst.prob, inst.it = p2, pkg.Class_iterate(p2.J, p2.H, p2.y, p2.s, delta=p2.delta_prev)
            inst.rhs = [pkg.System_rhs(*r) for r in p2.rhs[:N_DIRECTIONS]]
            inst.e2e_step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = max(2, min(args.steps, 5))
            for _ in range(reps):
                inst.e2e_step()
            torch.cuda.synchronize()
            e2e_pageable = (time.perf_counter() - t0) * 1e3 / reps
        except Exception:
            e2e_pageable = None
    config_extra = dict(nnz_M_lower=int(h.info("nnzM")), nnz_L=int(h.info("nnzL_true")), supernodes=int(h.info("nsuper")),
                        etree_levels=int(h.info("nlevels")), max_front=int(h.info("max_front")),
                        L_mb=8 * h.info("nnzL") / 1e6)
    t_symbolic = inst.t_symbolic
    t_symbolic_parts = inst.symbolic_breakdown
    inst.close()
    del inst, h
    torch.cuda.empty_cache()

    # ---------------- the other BASELINE.json shapes (rank 0, N = 1) ----------------
    others = {}
    if rank == 0 and world == 1 and not args.no_extra:
        for wname in OTHER_WORKLOADS:
            if wname == args.workload:
                continue
            try:
                others[wname] = measure_other_workload(pkg, torch, wname, local, hbm_peak, fp64_peak_1,
                                                       cpu=not args.no_cpu_baseline)
                nb = {"c2_chain_n100k": 64, "c4_elec_n1200": 64, "c3_sparse_qp_n200k": 8}.get(wname)
                if nb:
                    others[wname]["batched"] = measure_batched(pkg, torch, wname, local, nb)
                    others[wname]["batched"]["speedup_vs_one_at_a_time"] = others[wname]["value"] / others[wname]["batched"]["ms_per_instance_iter"]
            except Exception as e:      # never lose the headline line to an extra
                others[wname] = {"error": str(e)[:200]}

    # ---------------- CPU baseline of the headline workload: bounded sample (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            gen, kw, kws = WORKLOADS[args.workload]
            sp_prob = make_problem(args.workload, seed=0, sample=True)
            r = cpu_iteration(sp_prob, os.cpu_count())
            # the GPU on the same sample, through the plugin API (host buffers)
            si = Instance(pkg, torch, sp_prob, local)
            si.e2e_step(); si.e2e_step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                si.e2e_step()
            gpu_same = (time.perf_counter() - t0) * 1e3 / 3
            si.close()
            cpu = {"value": r["total_ms"], "unit": "ms/iter", "cores": os.cpu_count(), "kind": "port",
                   "sample": "%s%s (%s of the headline's size), 1 iteration: %s; the full-size CPU time is what "
                             "`--impl reference` measures" % (gen, kws, "bounded sample" if kws != kw else "all", CPU_NOTE),
                   "breakdown_ms": r, "gpu_e2e_same_sample_ms": gpu_same, "host_cores_available": os.cpu_count()}
        except Exception as e:      # never lose the headline line to the baseline leg
            cpu = {"error": str(e)[:300]}

    out = None
    if rank == 0:
        gen, kw, kws = WORKLOADS[args.workload]
        cfg = {"workload": args.workload, "generator": gen, "generator_args": kw,
               "n": n, "m": m, "nnz_J": int(prob.J.nnz), "nnz_M_lower": config_extra["nnz_M_lower"],
               "nnz_L": config_extra["nnz_L"], "factor_flops": cnt["F_chol"],
               "supernodes": config_extra["supernodes"], "etree_levels": config_extra["etree_levels"],
               "max_front": config_extra["max_front"],
               "directions_per_iter": N_DIRECTIONS, "refine": N_REFINE, "num_fac": nf_res,
               "delta": delta_res, "N_err": float(kkt_err[5]),
               "instances": 1 if sharded else world,
               "parallelism": ("one instance, elimination-tree subtrees mapped to %d GPUs, top separators over "
                               "NVLink peer memory" % world) if sharded else
                              "one independent instance per GPU (replicas)",
               "l2_policy": "working set larger than L2: factor L alone is %.0f MB and is streamed by every "
                            "factorisation and solve" % config_extra["L_mb"],
               "symbolic_s_once": t_symbolic, "symbolic_breakdown_s": t_symbolic_parts}
        out = {
            "metric": "kkt_factor_solve_ms_per_iter", "value": ms_step if sharded else ms_step / world, "unit": "ms/iter",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": False, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": cfg,
            "e2e": {"value": e2e_ms if sharded else e2e_ms / world, "unit": "ms/iter", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "host_memory": "pinned",
                    "pageable_host_memory_ms": e2e_pageable},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "phases_ms": phases,
            "roofline": roofline,
            "rooflines_by_phase": extra_rooflines,
            "rooflines_by_kernel": kernel_rooflines,
            "cpu_baseline": cpu,
            "workloads": others,
            "lib_launches_total": pkg.launch_count() - lib0,
        }
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
