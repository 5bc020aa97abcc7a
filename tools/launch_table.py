"""Per-kernel table (launches, total us, share) from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def table(path, first=0, last=None):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = []
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").strip().rsplit(">::", 1)[-1].replace("opb::", "")
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        if r[ui] == "ns":
            v /= 1e3
        elif r[ui] == "ms":
            v *= 1e3
        seq.append((name, v))
    seq = seq[first:last]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for n, v in seq:
        tot[n] += v
        cnt[n] += 1
    T = sum(tot.values())
    out = ["launches %d  total %.1f us" % (len(seq), T)]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        out.append("%-46s %7d %12.1f us %6.2f%%" % (k, cnt[k], v, 100 * v / T))
    return "\n".join(out)


if __name__ == "__main__":
    a = sys.argv
    print(table(a[1], int(a[2]) if len(a) > 2 else 0, int(a[3]) if len(a) > 3 else None))
