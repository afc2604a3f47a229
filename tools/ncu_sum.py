#!/usr/bin/env python
"""tools/ncu_sum.py LOG.csv ... — per kernel, over every launch in an `ncu --metrics ... --csv --log-file` list:
launches, warp instructions, thread instructions per warp instruction, device time, SM-active share of the elapsed
cycles and instructions per active SM cycle. (Launches under ncu run alone and cold: compare SHARES and counts.)"""
import collections
import csv
import sys

for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    if not rows:
        print(f, "empty")
        continue
    hdr = rows[0]
    iname, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    iu = hdr.index("Metric Unit")
    d = collections.defaultdict(dict)
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        if r[im] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        d[(int(r[iid]), r[iname].split("(")[0].split("<")[0].split("::")[-1])][r[im]] = v
    agg = collections.defaultdict(collections.Counter)
    for (_, k), m in d.items():
        a = agg[k]
        a["n"] += 1
        a["inst"] += m.get("smsp__inst_executed.sum", 0)
        a["thr"] += m.get("smsp__thread_inst_executed.sum", 0)
        a["us"] += m.get("gpu__time_duration.sum", 0)
        a["act"] += m.get("sm__cycles_active.avg", 0)
        a["el"] += m.get("sm__cycles_elapsed.max", 0)
    print(f)
    for k, a in sorted(agg.items()):
        n = a["n"]
        print(f"  {k:28s} launches {n:5d}  warp-instr {a['inst'] / 1e6:9.1f} M  thread/warp-instr {a['thr'] / max(1, a['inst']):5.1f}  "
              f"time {a['us'] / 1e3:8.2f} ms  SM-active {a['act'] / max(1, a['el']):.2f} of elapsed  instr/active SM cycle {a['inst'] / 148 / max(1, a['act']):.2f}")
