#!/usr/bin/env python
"""Reads `ncu --set full` reports and writes the per-launch counters bench.py quotes (profiles/dram_traffic.json).

usage: tools/ncu_counters.py OUT.json REPORT.ncu-rep [REPORT2.ncu-rep ...] [--note TEXT]
Each report is read with `ncu -i REPORT --page raw --csv`; launches of the same kernel are averaged.
"""
import csv
import io
import json
import re
import subprocess
import sys

METRICS = {
    "gpu__time_duration.sum": ("duration_us_alone", {"us": 1.0, "ns": 1e-3, "ms": 1e3}),
    "dram__bytes_read.sum": ("dram_read_bytes", {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}),
    "dram__bytes_write.sum": ("dram_write_bytes", {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}),
    "lts__t_sectors.sum": ("l2_sectors", {"sector": 1.0}),
    "smsp__inst_executed.sum": ("warp_instructions", {"inst": 1.0}),
    "smsp__thread_inst_executed_per_inst_executed.ratio": ("lanes_active_per_warp_instruction", None),
    "sm__inst_issued.avg.pct_of_peak_sustained_active": ("issue_slots_busy_pct", None),
    "sm__warps_active.avg.pct_of_peak_sustained_active": ("achieved_occupancy_pct", None),
    "l1tex__t_sector_hit_rate.pct": ("l1_hit_pct", None),
    "lts__t_sector_hit_rate.pct": ("l2_hit_pct", None),
    "sm__cycles_active.avg": ("sm_active_cycles", None),
    "sm__cycles_elapsed.max": ("elapsed_cycles", None),
}


def read_report(path):
    text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        name = re.sub(r"^void\s+", "", r[col["Kernel Name"]])
        name = re.sub(r"^rtx::", "", name)
        name = re.match(r"[A-Za-z_0-9]+", name).group(0)
        rec = {}
        for m, (key, scale) in METRICS.items():
            if m not in col:
                continue
            v = float(r[col[m]].replace(",", ""))
            if scale:
                v *= scale.get(units[col[m]], 1.0)
            rec[key] = v
        out.append((name, rec))
    return out


def main():
    args = sys.argv[1:]
    note = ""
    if "--note" in args:
        i = args.index("--note")
        note = args[i + 1]
        del args[i:i + 2]
    out_path, reports = args[0], args[1:]
    per = {}
    for rp in reports:
        for name, rec in read_report(rp):
            per.setdefault(name, []).append(rec)
    res = {}
    for name, recs in per.items():
        mean = {k: sum(r[k] for r in recs) / len(recs) for k in recs[0]}
        res[f"{name}_launches_captured"] = len(recs)
        res[f"{name}_dram_bytes_per_launch"] = mean["dram_read_bytes"] + mean["dram_write_bytes"]
        res[f"{name}_l2_bytes_per_launch"] = 32.0 * mean["l2_sectors"]
        for k in ("duration_us_alone", "warp_instructions", "lanes_active_per_warp_instruction", "issue_slots_busy_pct",
                  "achieved_occupancy_pct", "l1_hit_pct", "l2_hit_pct"):
            res[f"{name}_{k}"] = round(mean[k], 3)
        res[f"{name}_sm_active_fraction_of_launch"] = round(mean["sm_active_cycles"] / mean["elapsed_cycles"], 3)
    res["source"] = note
    json.dump(res, open(out_path, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
