"""bench.py's reference arm (the CPU restatement timed on the host cores) runs without a GPU:
check that it prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c3_small",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["higher_is_better"] is False and d["unit"] == "ms/iter"
    assert d["config"]["workload"] == "c3_small" and d["value"] > 0
    # the arm runs the workload it names: the FULL generator arguments of the table, not the bounded sample
    sys.path.insert(0, ROOT)
    import bench
    gen, kw, _ = bench.WORKLOADS["c3_small"]
    assert d["same_config"] is True and d["config"]["generator"] == gen and d["config"]["generator_args"] == kw
    assert d["config"]["n"] == kw["n"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "supernodal" in cb["sample"]


def test_reference_arm_under_torchrun_runs_on_rank_zero_only():
    """N > 1: the driver launches the arm under torchrun; rank 0 alone runs and prints, the others exit 0."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c3_small",
                        "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300,
                       cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    assert not [ln for ln in p.stdout.splitlines() if ln.startswith("{")]


def test_workload_table_is_consistent():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.DEFAULT_WORKLOAD in bench.WORKLOADS
    import __graft_entry__ as g
    pkg = g.package()
    for name, (gen, kw, kws) in bench.WORKLOADS.items():
        assert hasattr(problems, gen), name
        assert set(kws) <= set(kw) | {"N", "n", "m_gen", "nh", "n_p"}
