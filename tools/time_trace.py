"""Scratch: time the fixed-ray kernel (K1) on primary and secondary rays of a builtin scene."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rttnw_b200 as R
from rttnw_b200 import abi
from tests import _rays as RY

scene = int(sys.argv[1]) if len(sys.argv) > 1 else 9
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
ctx = R.Context(0)
desc = R.BuiltinDesc(scene)
gsc = R.DeviceScene(ctx, desc)
cam = desc.desc.camera
rng = np.random.default_rng(1)
rays = RY.camera_rays(cam, n, rng)
def run(rays, label):
    m = rays.shape[0]
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
    d_hits = torch.empty(m * 88, dtype=torch.uint8, device="cuda")
    for _ in range(2): gsc.trace_device(d_rays, d_hits, m)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): gsc.trace_device(d_rays, d_hits, m)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"scene {scene} {label}: {m} rays, {ms:.3f} ms, {m / ms / 1e6:.3f} Grays/s", flush=True)
    return d_hits.cpu().numpy().view(abi.HIT_DTYPE)
h1 = run(rays, "primary")
sec = RY.secondary_rays(h1, rays, rng)
h2 = run(sec, "secondary")
ter = RY.secondary_rays(h2, sec, rng)
run(ter, "tertiary")
perm = rng.permutation(ter.shape[0])
run(ter[perm], "tertiary shuffled")
# --- coherence experiments: the same tertiary rays in different orders ---
d = ter["direction"]
run(ter[np.argsort(d[:, 1] > 0, kind="stable")], "tertiary sorted by sign(dy)")
run(ter[np.argsort(d[:, 1] / np.linalg.norm(d, axis=1))], "tertiary sorted by dy/|d|")
octant = (d[:, 0] > 0).astype(int) | ((d[:, 1] > 0).astype(int) << 1) | ((d[:, 2] > 0).astype(int) << 2)
run(ter[np.argsort(octant, kind="stable")], "tertiary sorted by octant")
h3 = run(ter, "tertiary again")
run(ter[np.argsort(np.where(h3["prim_id"] < 0, 1e30, h3["t"]))], "tertiary sorted by hit distance (oracle-ish)")
# --- round 2: order by where the ray STARTS (the primitive its parent hit), alone and with the octant ---
src = h2["prim_id"][h2["prim_id"] >= 0]  # tertiary ray k starts on the primitive secondary ray k hit
assert src.shape[0] == ter.shape[0]
run(ter[np.argsort(src, kind="stable")], "tertiary sorted by source primitive")
run(ter[np.lexsort((src, octant))], "tertiary sorted by (octant, source primitive)")
run(ter[np.lexsort((octant, src // 4))], "tertiary sorted by (source primitive / 4, octant)")
o = ter["origin"]
lo, hi = np.percentile(o, 1, axis=0), np.percentile(o, 99, axis=0)
q = np.clip(((o - lo) / np.maximum(hi - lo, 1e-30) * 1024).astype(np.int64), 0, 1023)
def spread(v):
    v = (v | (v << 16)) & 0x30000FF; v = (v | (v << 8)) & 0x300F00F; v = (v | (v << 4)) & 0x30C30C3; return (v | (v << 2)) & 0x9249249
morton = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
run(ter[np.argsort(morton, kind="stable")], "tertiary sorted by origin Morton code (30 bits)")
run(ter[np.lexsort((morton >> 12, octant))], "tertiary sorted by (octant, Morton >> 12)")
run(ter[np.lexsort((octant, morton >> 15))], "tertiary sorted by (Morton >> 15, octant)")
