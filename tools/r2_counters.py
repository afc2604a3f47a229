#!/usr/bin/env python
"""tools/r2_counters.py LOG.csv SPP_TOTAL OUT.json [SCENE] — per-path-sample hardware counters of the wavefront kernels
from an `ncu --metrics ... --csv --log-file LOG.csv` pass over EVERY launch of a small render (tools/r2_profile.sh):
warp instructions, thread instructions (lanes), DRAM and L2 bytes, device time, per kernel and per path sample, tagged
with a hash of the kernel sources. bench.py turns them into live roofline fractions (instructions per sample x measured
samples/s against the issue slots of the measured clock) and refuses them when the hash is not the current sources'."""
import collections
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernels_sha():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "rttnw_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp")) and name not in ("cli.cpp", "png_io.cpp"):  # (what shapes the kernels' work)
            h.update(name.encode())
            h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


def main():
    log, spp_total, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    scene = int(sys.argv[4]) if len(sys.argv) > 4 else 9
    sys.path.insert(0, ROOT)
    from bench import SCENE_TABLE
    w, h = SCENE_TABLE[scene][:2]
    samples = w * h * spp_total
    rows = [r for r in csv.reader(open(log)) if len(r) > 10]
    hdr = rows[0]
    iname, im, iv, iid, iu = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
    per = collections.defaultdict(dict)
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        if r[im] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        if r[iu] in ("Kbyte", "Mbyte", "Gbyte"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[iu]]
        per[(int(r[iid]), r[iname].split("(")[0].split("<")[0].split("::")[-1].replace("void ", ""))][r[im]] = v
    agg = collections.defaultdict(collections.Counter)
    for (_, k), m in per.items():
        a = agg[k]
        a["launches"] += 1
        for key, name in (("warp_instructions", "smsp__inst_executed.sum"), ("thread_instructions", "smsp__thread_inst_executed.sum"),
                          ("dram_bytes", "dram__bytes_read.sum"), ("dram_bytes", "dram__bytes_write.sum"), ("l2_bytes", "lts__t_bytes.sum"),
                          ("time_us", "gpu__time_duration.sum"), ("sm_active_cycles", "sm__cycles_active.avg"), ("elapsed_cycles", "sm__cycles_elapsed.max")):
            a[key] += m.get(name, 0.0)
    doc = {"kernels_sha": kernels_sha(), "scene": scene, "width": w, "height": h, "spp_total": spp_total, "path_samples": samples,
           "source": "ncu --metrics (all launches of one render, each launch alone and cold under ncu): tools/r2_profile.sh", "kernels": {}}
    for k, a in sorted(agg.items()):
        doc["kernels"][k] = {
            "launches": int(a["launches"]),
            "warp_instructions_per_sample": a["warp_instructions"] / samples,
            "lanes_per_warp_instruction": a["thread_instructions"] / max(1.0, a["warp_instructions"]),
            "dram_bytes_per_sample": a["dram_bytes"] / samples,
            "dram_bytes_per_launch": a["dram_bytes"] / a["launches"],
            "l2_bytes_per_sample": a["l2_bytes"] / samples,
            "time_us_per_launch_alone": a["time_us"] / a["launches"],
            "sm_active_share_of_launch": a["sm_active_cycles"] / max(1.0, a["elapsed_cycles"]),
            "instructions_per_active_sm_cycle": a["warp_instructions"] / 148.0 / max(1.0, a["sm_active_cycles"]),
        }
    json.dump(doc, open(out, "w"), indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
