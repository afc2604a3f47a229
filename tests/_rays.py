"""Fixed-ray sets for the closest-hit parity tests: half camera rays on a
jittered grid, half secondary-like rays leaving surfaces the first half hit."""
from __future__ import annotations

import numpy as np

from rttnw_b200 import abi

DBL_MAX = np.finfo(np.float64).max


def camera_basis(cam: abi.Camera):
    """Camera::new, camera.rs:32-61, in numpy (test-side ray generation only)."""
    lookfrom, lookat, vup = (np.array(list(v), dtype=np.float64) for v in (cam.lookfrom, cam.lookat, cam.view_up))
    theta = cam.vertical_fov * np.pi / 180.0
    hh = np.tan(theta / 2.0)
    hw = cam.aspect_ratio * hh
    w = lookfrom - lookat
    w /= np.linalg.norm(w)
    u = np.cross(vup, w)
    u /= np.linalg.norm(u)
    v = np.cross(w, u)
    fd = cam.focus_distance
    llc = lookfrom - hw * fd * u - hh * fd * v - fd * w
    return lookfrom, llc, 2 * hw * fd * u, 2 * hh * fd * v, u, v


def camera_rays(cam: abi.Camera, n: int, rng: np.random.Generator) -> np.ndarray:
    origin, llc, hor, ver, u, v = camera_basis(cam)
    s, t = rng.random(n), rng.random(n)
    ang, rad = rng.random(n) * 2 * np.pi, np.sqrt(rng.random(n)) * cam.aperture / 2.0
    off = np.outer(rad * np.cos(ang), u) + np.outer(rad * np.sin(ang), v)
    rays = np.zeros(n, dtype=abi.RAY_DTYPE)
    rays["origin"] = origin + off
    rays["direction"] = llc + np.outer(s, hor) + np.outer(t, ver) - origin - off
    rays["time"] = cam.open_time + (cam.close_time - cam.open_time) * rng.random(n)
    rays["t_min"], rays["t_max"] = 0.001, DBL_MAX
    rays["xi"] = rng.random(n) * 0.999 + 0.0005
    return rays


def secondary_rays(hits: np.ndarray, parent: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    """Rays starting at hit points, direction = normal + point in the unit ball (the Lambertian
    lobe of material.rs:90-99), or a uniform direction for volume hits / a mirror-like one."""
    ok = hits["prim_id"] >= 0
    h, par = hits[ok], parent[ok]
    n = h.shape[0]
    ball = rng.normal(size=(n, 3))
    ball /= np.linalg.norm(ball, axis=1, keepdims=True)
    ball *= np.cbrt(rng.random((n, 1)))
    nrm = h["normal"] / np.maximum(np.linalg.norm(h["normal"], axis=1, keepdims=True), 1e-300)
    d = nrm + ball
    mode = rng.random(n)
    d = np.where((mode < 0.15)[:, None], ball, d)  # isotropic-like
    d = np.where((mode > 0.9)[:, None], -nrm + 0.3 * ball, d)  # going through (dielectric-like)
    rays = np.zeros(n, dtype=abi.RAY_DTYPE)
    rays["origin"], rays["direction"] = h["p"], d
    rays["time"] = par["time"]
    rays["t_min"], rays["t_max"] = 0.001, DBL_MAX
    rays["xi"] = rng.random(n) * 0.999 + 0.0005
    return rays


def compare_hits(gpu: np.ndarray, ref: np.ndarray, fragile: np.ndarray, rtol: float = 1e-5) -> dict:
    """The fixed-ray contract of BASELINE.json: primitive id bit-exact (grazing ties excluded),
    t / p / normal / u / v within `rtol` relative. Returns counts and the largest relative errors seen
    (what profiles/r2_parity.json records); raises on violation."""
    ok = ~fragile
    ids_g, ids_r = gpu["prim_id"][ok], ref["prim_id"][ok]
    bad = np.nonzero(ids_g != ids_r)[0]
    assert bad.size == 0, f"{bad.size} primitive-id mismatches outside grazing ties, first: gpu={gpu[ok][bad[:3]]} ref={ref[ok][bad[:3]]}"
    hit = ok & (ref["prim_id"] >= 0)
    g, r = gpu[hit], ref[hit]
    assert np.array_equal(g["front_face"], r["front_face"])
    known = r["material"] != -1  # the oracle's own scenes carry no material indices
    assert np.array_equal(g["material"][known], r["material"][known])

    def rel(a, b, scale=None):
        scale = np.maximum(np.abs(b), 1e-300) if scale is None else scale
        e = np.abs(a - b) / scale
        return float(e.max()) if e.size else 0.0
    err = {"t": rel(g["t"], r["t"])}
    pscale = np.maximum(np.linalg.norm(r["p"], axis=1, keepdims=True), 1.0)
    err["p"] = rel(g["p"], r["p"], pscale)
    nscale = np.maximum(np.linalg.norm(r["normal"], axis=1, keepdims=True), 1e-300)
    err["normal"] = rel(g["normal"], r["normal"], nscale)
    # u wraps at the sphere seam; the relative scale is floored at 1e-4
    du = np.abs(g["u"] - r["u"])
    du = np.minimum(du, 1.0 - du)
    err["u"] = float((du / np.maximum(np.abs(r["u"]), 1e-4)).max()) if du.size else 0.0
    err["v"] = rel(g["v"], r["v"], np.maximum(np.abs(r["v"]), 1e-4))
    for k, e in err.items():
        assert e <= rtol, (k, e)
    return {"rays": int(gpu.shape[0]), "fragile": int(fragile.sum()), "hits": int(hit.sum()),
            "id_mismatches_outside_ties": 0, "max_rel_err": err}


def merge_stats(a: dict | None, b: dict) -> dict:
    """Sum of the counts, maximum of the errors, of two compare_hits() results."""
    if a is None:
        return b
    out = {k: a[k] + b[k] for k in ("rays", "fragile", "hits", "id_mismatches_outside_ties")}
    out["max_rel_err"] = {k: max(a["max_rel_err"][k], b["max_rel_err"][k]) for k in b["max_rel_err"]}
    return out
