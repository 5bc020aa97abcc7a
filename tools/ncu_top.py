"""Profile helper (runs on the GPU box): time-only launch list of one bench step, then one
`ncu --set full` capture of the longest launch of each named kernel.
    python tools/ncu_top.py WORKLOAD OUT_PREFIX kernel_name [kernel_name ...]"""
import csv
import re
import subprocess
import sys


def durations(path):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1].split("<")[0].strip()
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        out.append((name, v))
    return out


def main():
    workload, prefix, names = sys.argv[1], sys.argv[2], sys.argv[3:]
    bench = ["python", "bench.py", "--ncu", "--workload", workload, "--warmup", "0", "--steps", "1"]
    lst = prefix + "_launches.csv"
    subprocess.run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--csv",
                    "--log-file", lst] + bench, check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    d = durations(lst)
    for nm in names:
        mine = [v for n, v in d if n == nm]
        if not mine:
            print("no launches of", nm)
            continue
        idx = max(range(len(mine)), key=lambda i: mine[i])
        print("%s: %d launches, longest #%d = %.1f us" % (nm, len(mine), idx, mine[idx]), flush=True)
        subprocess.run(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on",
                        "-k", "regex:" + nm, "-s", str(idx), "-c", "1", "-f", "-o", "%s_%s" % (prefix, nm)] + bench,
                       check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


if __name__ == "__main__":
    main()
