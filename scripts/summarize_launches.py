"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / mean / share."""
import collections
import csv
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10]
hdr = rows[0]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg = collections.OrderedDict()
tot = 0.0
for r in rows[1 + skip:]:
    n = r[ki].split("(")[0][-48:]
    v = float(r[vi].replace(",", "")) / 1e3
    agg.setdefault(n, []).append((v, r[gi]))
    tot += v
print(f"{'kernel':50s} {'n':>5s} {'mean us':>10s} {'total us':>10s} {'share':>7s}  grid(example)")
for k, v in sorted(agg.items(), key=lambda kv: -sum(x[0] for x in kv[1])):
    s = sum(x[0] for x in v)
    print(f"{k:50s} {len(v):5d} {s / len(v):10.2f} {s:10.1f} {100 * s / tot:6.1f}%  {v[len(v) // 2][1]}")
print(f"total {tot:.1f} us over {len(rows) - 1 - skip} launches")
