"""Text summary of `ncu -i X.ncu-rep --page raw --csv` files (one capture each):
    python tools/ncu_summary.py out.txt title kernel1.raw.csv [kernel2.raw.csv ...]"""
import csv
import sys

KEYS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def one(path):
    rows = list(csv.reader(open(path, errors="replace")))
    rows = [r for r in rows if r and not r[0].startswith("==")]
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else path
    out = ["== " + name.replace("opb::<unnamed>::", "")]
    col = {h: i for i, h in enumerate(hdr)}
    for k in KEYS:
        if k in col:
            out.append("  %-82s %s %s" % (k, vals[col[k]], units[col[k]]))
    stalls = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(vals[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    out.append("  top stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:4]))
    return "\n".join(out)


if __name__ == "__main__":
    dst, title, files = sys.argv[1], sys.argv[2], sys.argv[3:]
    with open(dst, "w") as f:
        f.write(title + "\n\n")
        for p in files:
            try:
                f.write(one(p) + "\n\n")
            except Exception as e:  # noqa: BLE001
                f.write("== %s: %r\n\n" % (p, e))
    print(open(dst).read())
