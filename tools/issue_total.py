#!/usr/bin/env python
"""Sums the warp / thread instructions of every wavefront launch of ONE render from an ncu CSV log
(`ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum -k regex:wf_
 -s <launches of the warm-up renders> -c <launches of one render> --csv --log-file LOG python bench.py ...`)
and relates them to the issue slots of the render's un-profiled duration.

usage: tools/issue_total.py LOG.csv RENDER_MS [SM_MHZ=1965] [SMS=148]
"""
import csv
import sys


def main():
    log, render_ms = sys.argv[1], float(sys.argv[2])
    mhz = float(sys.argv[3]) if len(sys.argv) > 3 else 1965.0
    sms = int(sys.argv[4]) if len(sys.argv) > 4 else 148
    rows = list(csv.reader(open(log, errors="replace")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ix = {h: j for j, h in enumerate(hdr)}
    per = {}
    for r in rows[start + 1:]:
        if len(r) <= ix["Metric Value"]:
            continue
        k = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("rtx::", "").split("<")[0]
        per.setdefault(k, {}).setdefault(r[ix["Metric Name"]], []).append(float(r[ix["Metric Value"]].replace(",", "")))
    total_w = 0.0
    for k, m in per.items():
        w, t, d = m["smsp__inst_executed.sum"], m["smsp__thread_inst_executed.sum"], m["gpu__time_duration.sum"]
        total_w += sum(w)
        print(f"{k}: {len(w)} launches, warp instructions {sum(w):.4g} (mean {sum(w) / len(w):.4g}, min {min(w):.3g}, max {max(w):.3g}), "
              f"lanes per warp instruction {sum(t) / sum(w):.2f}, serialised time {sum(d) / 1e6:.1f} ms")
    slots = render_ms * 1e-3 * mhz * 1e6 * 4 * sms
    print(f"all: {total_w:.4g} warp instructions in a render of {render_ms:.1f} ms = {total_w / slots:.3f} of the issue slots "
          f"({sms} SMs x 4 per clock x {mhz:.0f} MHz)")


if __name__ == "__main__":
    main()
