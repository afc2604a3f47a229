"""Sums gpu__time_duration per kernel name from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value"); u = hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) <= v: continue
    name = r[k].split('(')[0][:70]; val = float(r[v].replace(',', ''))
    val *= {"ms": 1e6, "us": 1e3, "ns": 1.0, "s": 1e9}.get(r[u], 1.0)
    agg[name][0] += 1; agg[name][1] += val
tot = sum(t for _, t in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t / 1e6:10.3f} ms {100 * t / tot:5.1f}% {c:5d} launches  {n}")
