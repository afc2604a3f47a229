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
