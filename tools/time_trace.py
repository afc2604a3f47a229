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
