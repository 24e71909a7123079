"""Per-source-line stall-sample shares (where the time goes for latency-bound kernels) of an ncu source page:
   ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > f.csv ; python scripts/ncu_samples.py f.csv [top] [function substring]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
want = sys.argv[3] if len(sys.argv) > 3 else None
cur = None; fn = None; per = {}
for row in rows:
    if not row: continue
    if row[0] == 'File Path': cur = row[1].split('/')[-1]; continue
    if row[0] in ('Function Name', 'Kernel Name'): fn = row[1]; continue
    if row[0] == 'Line No':
        ii = row.index('Instructions Executed'); si = row.index('# Samples'); continue
    if row[0] != '':
        try: n = int(row[ii]); s = int(row[si]); ln = int(row[0])
        except ValueError: continue
        agg = per.setdefault(fn, {})
        a = agg.get((cur, ln), (0, 0, row[1])); agg[(cur, ln)] = (a[0] + n, a[1] + s, row[1])
for fn, agg in per.items():
    if want and want not in fn: continue
    tot = max(sum(v[0] for v in agg.values()), 1); stot = max(sum(v[1] for v in agg.values()), 1)
    print('==', fn[:100], 'warp instructions', tot, 'samples', stot)
    for (f, ln), (n, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{f}:{ln:4d} smp {s/stot*100:5.2f}% inst {n/tot*100:5.2f}%  {src.strip()[:110]}")
