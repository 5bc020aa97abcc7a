"""Per-kernel time and DRAM bytes of one step from an ncu launch list taken with
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv:
    python tools/dram_table.py launches.csv [out.json [workload]]
Prints a per-kernel table for the LAST real factorisation attempt + the rest of the step and, with out.json,
writes the totals bench.py reads as `roofline.traffic`."""
import collections
import csv
import json
import sys


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= col["Metric Value"]:
            continue
        lid = int(r[col["ID"]])
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").strip().rsplit(">::", 1)[-1].replace("opb::", "")
        try:
            v = float(r[col["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        unit = r[col["Metric Unit"]]
        m = r[col["Metric Name"]]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)          # -> us
        else:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)
        d = launches.setdefault(lid, {"name": name, "us": 0.0, "rd": 0.0, "wr": 0.0})
        if m == "gpu__time_duration.sum":
            d["us"] = v
        elif m == "dram__bytes_read.sum":
            d["rd"] = v
        elif m == "dram__bytes_write.sum":
            d["wr"] = v
    return list(launches.values())


def main():
    L = load(sys.argv[1])
    tot = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in L:
        t = tot[d["name"]]
        t[0] += 1; t[1] += d["us"]; t[2] += d["rd"]; t[3] += d["wr"]
    T = sum(t[1] for t in tot.values())
    print("launches %d  total %.1f ms  DRAM read %.1f GB  write %.1f GB" % (
        len(L), T / 1e3, sum(t[2] for t in tot.values()) / 1e9, sum(t[3] for t in tot.values()) / 1e9))
    print("%-44s %7s %11s %7s %10s %10s %9s" % ("kernel", "n", "ms", "share", "read GB", "write GB", "GB/s"))
    for k, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %7d %11.3f %6.2f%% %10.3f %10.3f %9.0f" % (k[:44], t[0], t[1] / 1e3, 100 * t[1] / T, t[2] / 1e9, t[3] / 1e9,
                                                        (t[2] + t[3]) / max(t[1], 1e-9) / 1e3))
    if len(sys.argv) > 2:
        workload = sys.argv[3] if len(sys.argv) > 3 else "c5_pde_100"
        fac = ("front_cb", "chol_panel_update", "chol_trsm", "chol_diag", "big_extend_add_panel", "zero_kernel", "scatter_kernel",
               "front_small", "mid_panel", "trtri_merge", "ctl_")
        asm = ("prep_kernel", "assemble_M", "diag_extract")

        def is_in(n, names):
            return any(n.startswith(f) for f in names)
        by = lambda names: sum(t[2] + t[3] for k, t in tot.items() if is_in(k, names))      # noqa: E731
        rest = sum(t[2] + t[3] for k, t in tot.items() if not is_in(k, fac) and not is_in(k, asm))
        out = {workload: {
            "factor_bytes_per_attempt": by(fac),
            "front_cb_bytes_per_attempt": by(("front_cb",)),
            "panel_update_bytes_per_attempt": by(("chol_panel_update",)),
            "assembly_bytes": by(asm),
            "direction_bytes": rest / 2.0,
            "source": "profiles/%s: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over every launch of "
                      "one bench step (one factorisation attempt, two directions); direction_bytes = everything that is not "
                      "assembly or factorisation, per direction" % sys.argv[1].split("/")[-1]}}
        json.dump(out, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
