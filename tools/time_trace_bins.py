"""Scratch: how much of the gain of origin-sorted rays (tools/time_trace.py) is left with coarse bins, random order inside a
bin, a launch-sized ray set (256 Ki) and the cell grid laid over the min/max box of the origins instead of percentiles."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rttnw_b200 as R
from rttnw_b200 import abi
from tests import _rays as RY

scene = int(sys.argv[1]) if len(sys.argv) > 1 else 9
n = int(sys.argv[2]) if len(sys.argv) > 2 else 530_000
ctx = R.Context(0)
desc = R.BuiltinDesc(scene)
gsc = R.DeviceScene(ctx, desc)
rng = np.random.default_rng(1)
rays = RY.camera_rays(desc.desc.camera, n, rng)
def run(rays, label, quiet=False):
    m = rays.shape[0]
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
    d_hits = torch.empty(m * 88, dtype=torch.uint8, device="cuda")
    for _ in range(3): gsc.trace_device(d_rays, d_hits, m)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): gsc.trace_device(d_rays, d_hits, m)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if not quiet: print(f"scene {scene} {label}: {m} rays, {ms*1000:.1f} us, {m / ms / 1e6:.3f} Grays/s", flush=True)
    return d_hits.cpu().numpy().view(abi.HIT_DTYPE)
h1 = run(rays, "", True)
sec = RY.secondary_rays(h1, rays, rng)
h2 = run(sec, "", True)
ter = RY.secondary_rays(h2, sec, rng)
src = h2["prim_id"][h2["prim_id"] >= 0]
perm = rng.permutation(ter.shape[0])
ter, src = ter[perm], src[perm]   # random order: what is inside a bin stays random under the stable sorts below
run(ter, "random order")
d, o = ter["direction"], ter["origin"]
octant = (d[:, 0] > 0).astype(np.int64) | ((d[:, 1] > 0).astype(np.int64) << 1) | ((d[:, 2] > 0).astype(np.int64) << 2)
def spread(v):
    v = (v | (v << 16)) & 0x30000FF; v = (v | (v << 8)) & 0x300F00F; v = (v | (v << 4)) & 0x30C30C3; return (v | (v << 2)) & 0x9249249
for box in ("minmax", "p1-p99"):
    lo, hi = (o.min(axis=0), o.max(axis=0)) if box == "minmax" else (np.percentile(o, 1, axis=0), np.percentile(o, 99, axis=0))
    print(f"  box {box}: lo {lo} hi {hi}")
    q = np.clip(((o - lo) / np.maximum(hi - lo, 1e-30) * 1024).astype(np.int64), 0, 1023)
    morton = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    for bits in (3, 6, 9, 12, 15, 30):
        cell = morton >> (30 - bits)
        key = (octant << bits) | cell
        run(ter[np.argsort(key, kind="stable")], f"{box}: (octant, {bits}-bit cell) = {8 << bits} bins, {len(np.unique(key))} used")
    cell = morton >> 21
    run(ter[np.argsort((cell << 3) | octant, kind="stable")], f"{box}: (9-bit cell, octant)")
    run(ter[np.argsort(cell, kind="stable")], f"{box}: 9-bit cell alone")
for sh in (0, 2, 4):
    key = octant * (src.max() + 1) + (src >> sh)
    run(ter[np.argsort(key, kind="stable")], f"(octant, source primitive >> {sh}), {len(np.unique(key))} used")
