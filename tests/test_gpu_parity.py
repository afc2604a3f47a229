"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle.

Fixed rays (BASELINE.json): primitive id bit-exact outside grazing ties, t / p / normal / u / v
within 1e-5 relative. Renders: PSNR >= 35 dB and mean per-channel error <= 1/255 at equal high
spp, plus same-counter (Philox) low-depth comparisons that are nearly sample-exact.
"""
import math
import os

import numpy as np
import pytest

import rttnw_b200 as R
from rttnw_b200 import abi
from rttnw_b200 import scene as S
from tests import _oracle as O
from tests import _rays as RY

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = R.Context(0)
    yield c
    c.close()


def both(ctx, desc, bvh_seed=7):
    return R.DeviceScene(ctx, desc), O.OracleScene.from_desc(desc, bvh_seed)


def check_rays(gpu_scene, osc, rays, min_ok=0.98):
    ref, fragile = osc.trace(rays)
    got = gpu_scene.trace(rays)
    st = RY.compare_hits(got, ref, fragile)
    assert 1.0 - st["fragile"] / max(1, st["rays"]) >= min_ok, st
    return got, ref, st


# ---------------------------------------------------------------------------
# known answers through the C ABI (same cases that pin the oracle)
# ---------------------------------------------------------------------------
def test_primitive_known_answers(ctx):
    mat = S.Lambertian((0.5, 0.5, 0.5))
    sc = R.DeviceScene(ctx, S.Scene(S.List([S.Sphere((0, 0, 0), 1.0, mat)])).to_desc())
    h = sc.trace(O.make_rays((0, 0, -5), (0, 0, 1)))[0]
    assert h["prim_id"] == 0 and h["t"] == pytest.approx(4.0) and h["front_face"] == 1
    assert tuple(h["normal"]) == pytest.approx((0, 0, -1)) and (h["u"], h["v"]) == pytest.approx((0.75, 0.5))
    assert sc.trace(O.make_rays((0, 0, -5), (0, 0, 1), t_max=4.0))[0]["prim_id"] == 0  # inclusive t_max (Q9)
    assert sc.trace(O.make_rays((0, 0, -5), (0, 0, 1), t_max=3.999))[0]["prim_id"] == abi.RTX_MISS
    h = sc.trace(O.make_rays((0, 0, 0), (0, 0, 1)))[0]
    assert h["t"] == pytest.approx(1.0) and h["front_face"] == 0
    sc = R.DeviceScene(ctx, S.Scene(S.List([S.XY.rectangle(mat, (0, 2), (0, 4), 1.0)])).to_desc())
    assert sc.trace(O.make_rays((0, 1, 0), (0, 0, 1)))[0]["prim_id"] == 0  # half-open ranges (Q11)
    assert sc.trace(O.make_rays((2, 1, 0), (0, 0, 1)))[0]["prim_id"] == abi.RTX_MISS
    assert sc.trace(O.make_rays((1, 1, 0), (1, 0, 0)))[0]["prim_id"] == abi.RTX_MISS  # parallel ray
    h = sc.trace(O.make_rays((1, 1, 0), (0, 0, 1)))[0]
    assert (h["u"], h["v"]) == pytest.approx((0.5, 0.25)) and h["front_face"] == 0
    sc = R.DeviceScene(ctx, S.Scene(S.List([S.Cube((0, 0, 0), (1, 2, 3), mat)])).to_desc())
    for d, face in {(0, 0, 1): 0, (0, 0, -1): 1, (0, 1, 0): 2, (0, -1, 0): 3, (1, 0, 0): 4, (-1, 0, 0): 5}.items():
        o = tuple((0.5, 1.0, 1.5)[i] - 10 * d[i] for i in range(3))
        assert sc.trace(O.make_rays(o, d))[0]["prim_id"] == face  # Cube::new order (Q12)


def test_yrotate_quirk_and_translate(ctx):  # Q13 / Q14 closed forms
    mat = S.Lambertian((0.5, 0.5, 0.5))
    th = math.radians(15.0)
    s, c = math.sin(th), math.cos(th)
    rect = S.XY.rectangle(mat, (-5, 5), (-5, 5), 1.0)
    sc = R.DeviceScene(ctx, S.Scene(S.List([rect.rotate_y(15.0).translate((10, 20, 30))])).to_desc())

    def to_world(v):
        return np.array([c * v[0] + s * v[2], v[1], -s * v[0] + c * v[2]])
    o = to_world(np.array([0.3, 0.2, 5.0])) + np.array([10, 20, 30.0])
    h = sc.trace(O.make_rays(o, to_world(np.array([0.0, 0.0, -1.0]))))[0]
    p0 = c * 0.3 + s
    assert h["t"] == pytest.approx(4.0)
    assert tuple(h["normal"]) == pytest.approx((s, 0.0, -s * s + c))
    assert tuple(h["p"]) == pytest.approx((p0 + 10, 20.2, -s * p0 + c + 30))


def test_constant_medium_known_answer(ctx):  # Q16
    density, xi = 0.5, 0.6
    dist = -math.log(xi) / density
    for boundary in (S.Sphere((0, 0, 0), 2.0, S.Dielectric(1.5)),  # analytic fast path
                     S.Sphere((0, 0, 0), 2.0, S.Dielectric(1.5)).translate((0, 0, 0))):  # general boundary query
        sc = R.DeviceScene(ctx, S.Scene(S.List([S.ConstantMedium(boundary, density, (1, 1, 1))])).to_desc())
        h = sc.trace(O.make_rays((0, 0, -5), (0, 0, 2), xi=xi))[0]
        assert h["prim_id"] == 1 and h["t"] == pytest.approx(1.5 + dist / 2.0)
        assert tuple(h["normal"]) == (1.0, 0.0, 0.0) and h["front_face"] == 1 and h["material"] == -1
        assert sc.trace(O.make_rays((0, 0, -5), (0, 0, 2), xi=math.exp(-density * 4.0 * 1.01)))[0]["prim_id"] == abi.RTX_MISS
        assert sc.trace(O.make_rays((0, 0, 0), (0, 0, 1), xi=xi))[0]["t"] == pytest.approx(0.001 + dist)
        assert sc.trace(O.make_rays((0, 0, -5), (0, 0, 2), xi=xi, t_max=1.4))[0]["prim_id"] == abi.RTX_MISS


def test_empty_and_degenerate_inputs(ctx):
    mat = S.Lambertian((0.5, 0.5, 0.5))
    sc = R.DeviceScene(ctx, S.Scene(S.List([])).to_desc())  # empty world
    assert sc.trace(O.make_rays((0, 0, 0), (0, 0, 1)))[0]["prim_id"] == abi.RTX_MISS
    assert sc.trace(np.zeros(0, dtype=abi.RAY_DTYPE)).shape == (0,)  # zero rays
    sc, osc = both(ctx, S.Scene(S.List([S.Sphere((0, 0, 0), 1.0, mat), S.XZ.rectangle(mat, (-3, 3), (-3, 3), -1.0)])).to_desc())
    rays = O.make_rays([(0, 0, -5), (0, 0, -5), (0, 5, 0), (0, 0, -5), (1e30, 0, 0)],
                       [(0, 0, 0), (0, 0, 1e-300), (0, -1, 0), (1, 0, 0), (-1, 0, 0)])
    got, (ref, fragile) = sc.trace(rays), osc.trace(rays)
    # zero / denormal / axis-aligned directions, far origins. A zero direction makes every reference
    # test compare NaNs (Sphere::hit then "hits" at t = NaN) and the 1e30 origin cancels the whole
    # discriminant: the oracle flags both as grazing ties, where only "no crash, valid id" is required.
    assert list(fragile) == [True, True, False, False, True]
    assert np.array_equal(got["prim_id"][~fragile], ref["prim_id"][~fragile])
    assert ((got["prim_id"] >= -1) & (got["prim_id"] < 2)).all()


# ---------------------------------------------------------------------------
# fixed rays on the nine scenes of scenes.rs
# ---------------------------------------------------------------------------
# Largest share of rays the oracle may flag as grazing ties (a comparison of the reference within 1e-9 relative of
# flipping, for a candidate at or before the final hit) per scene and ray generation: twice what 10^6 rays per scene
# showed in round 2 (scene 7: 0 / 0.18 % / 1.46 %; scene 9: 0 / 1.48 % / 3.51 %; scene 1: < 0.003 %; the others: none —
# profiles/r2_parity.json has the counts of every run), floored at 0.02 %. Scene 9's later generations travel INSIDE the
# ground boxes and meet the exactly coplanar side faces adjacent boxes share (scenes.rs:244-253); scene 7's graze the
# walls the two boxes stand on: genuine two-primitive ties.
MAX_TIES = {7: (2e-4, 4e-3, 3e-2), 9: (2e-4, 3e-2, 7e-2)}


@pytest.mark.parametrize("number", range(1, 10))
def test_fixed_rays_builtin_scene(ctx, number, earth_rgba):
    from tests._record import record
    n = 60000 if number in (1, 9) else 100000
    rng = np.random.default_rng(0xF17ED + number)
    gsc = R.DeviceScene(ctx, R.BuiltinDesc(number))
    osc = O.OracleScene.builtin(number, earth=earth_rgba)  # the oracle's OWN restatement of scenes.rs
    cam, _ = osc.camera()
    limits = MAX_TIES.get(number, (2e-4, 2e-4, 2e-4))
    primary = RY.camera_rays(cam, n, rng)
    got, ref, st1 = check_rays(gsc, osc, primary, min_ok=1.0 - limits[0])
    secondary = RY.secondary_rays(ref, primary, rng)
    _, ref2, st2 = check_rays(gsc, osc, secondary, min_ok=1.0 - limits[1])
    tertiary = RY.secondary_rays(ref2, secondary, rng)
    _, _, st3 = check_rays(gsc, osc, tertiary, min_ok=1.0 - limits[2])
    assert st1["hits"] > 0.3 * n
    record("fixed_rays_three_generations", f"scene_{number}", {"primary": st1, "secondary": st2, "tertiary": st3, "max_tie_share_allowed": limits})
    print(f"scene {number}: {st1} {st2} {st3}")


def random_tree(rng, with_media=True):
    mats = [S.Lambertian((0.5, 0.5, 0.5)), S.Metal((0.8, 0.8, 0.8), 0.3), S.Dielectric(1.5),
            S.DiffuseLight((4, 4, 4)), S.Lambertian(S.CheckerTexture((0.1, 0.1, 0.1), (0.9, 0.9, 0.9)))]

    def leaf():
        kind = int(rng.integers(0, 4))
        m = mats[int(rng.integers(0, len(mats)))]
        c = tuple(rng.uniform(-20, 20, 3))
        if kind == 0:
            return S.Sphere(c, float(rng.uniform(0.3, 4)), m)
        if kind == 1:
            return S.MovingSphere((c, tuple(np.array(c) + rng.uniform(-1, 1, 3))), (0, 1), float(rng.uniform(0.3, 2)), m)
        if kind == 2:
            plane = [S.XY, S.XZ, S.YZ][int(rng.integers(0, 3))]
            a0, b0 = rng.uniform(-20, 10, 2)
            return plane.rectangle(m, (a0, a0 + rng.uniform(1, 15)), (b0, b0 + rng.uniform(1, 15)), float(rng.uniform(-20, 20)))
        lo = np.array(c)
        return S.Cube(tuple(lo), tuple(lo + rng.uniform(0.5, 8, 3)), m)

    def node(depth):
        r = rng.random()
        if depth >= 3 or r < 0.35:
            return leaf()
        if r < 0.55:
            return S.List([node(depth + 1) for _ in range(int(rng.integers(1, 6)))])
        if r < 0.7:
            return S.BvhTree(S.List([leaf() for _ in range(int(rng.integers(1, 40)))]))
        if r < 0.85:
            return node(depth + 1).translate(tuple(rng.uniform(-10, 10, 3)))
        return node(depth + 1).rotate_y(float(rng.uniform(-180, 180)))
    items = [node(0) for _ in range(int(rng.integers(2, 12)))]
    if with_media:
        for _ in range(int(rng.integers(0, 3))):
            c = np.array(rng.uniform(-10, 10, 3))
            choice = rng.random()
            if choice < 0.34:
                b = S.Sphere(tuple(c), float(rng.uniform(2, 8)), mats[2])
            elif choice < 0.67:
                b = S.Cube(tuple(c), tuple(c + rng.uniform(2, 8, 3)), mats[0]).rotate_y(float(rng.uniform(-90, 90))).translate(tuple(rng.uniform(-5, 5, 3)))
            else:
                b = S.MovingSphere((tuple(c), tuple(c + 1.0)), (0, 1), 3.0, mats[0])
            med = S.ConstantMedium(b, float(rng.uniform(0.01, 0.5)), (0.5, 0.5, 0.5))
            items.append(med if rng.random() < 0.7 else med.rotate_y(float(rng.uniform(-45, 45))).translate(tuple(rng.uniform(-3, 3, 3))))
    return S.List(items)


@pytest.mark.parametrize("seed", range(12))
def test_fixed_rays_random_scene_trees(ctx, seed):
    """Arbitrary trees of the reference's API: nested List / BvhTree / translate / rotate_y chains,
    instanced media, dielectric under wrappers — everything the description can express."""
    rng = np.random.default_rng(1000 + seed)
    desc = S.Scene(random_tree(rng)).to_desc()
    gsc, osc = both(ctx, desc, bvh_seed=seed)
    n = 30000
    o = rng.uniform(-40, 40, (n, 3))
    target = rng.uniform(-15, 15, (n, 3))
    rays = O.make_rays(o, target - o, time=rng.random(n), xi=rng.random(n) * 0.999 + 0.0005)
    rays["t_min"] = np.where(rng.random(n) < 0.2, -np.inf, 0.001)  # also the (-inf, x) ranges media use
    got, ref, st = check_rays(gsc, osc, rays)
    sec = RY.secondary_rays(ref, rays, rng)
    if sec.shape[0]:
        check_rays(gsc, osc, sec)
    assert st["hits"] > 0.2 * n


TEN_MILLION = {  # scene: (instanced-geometry id range whose p is Q14-displaced, or None)
    1: None, 2: None, 3: None, 4: None, 5: None, 6: None, 7: (6, 18), 8: (6, 20), 9: (2411, 3411),
}
# Largest share of the 10^7 rays (half primary, half secondary) the oracle may flag as grazing ties, per scene: twice the
# share 10^6 rays per scene showed (half of the secondary shares above), floored at 0.01 %.
MAX_TIE_SHARE = {1: 1e-4, 2: 1e-4, 3: 1e-4, 4: 1e-4, 5: 1e-4, 6: 1e-4, 7: 2e-3, 8: 1e-4, 9: 1.5e-2}


def trace_on_device(gsc, rays):
    import torch
    n = rays.shape[0]
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
    d_hits = torch.empty(n * 88, dtype=torch.uint8, device="cuda")
    gsc.trace_device(d_rays, d_hits, n)
    return d_hits.cpu().numpy().view(abi.HIT_DTYPE)


@pytest.mark.parametrize("number", sorted(TEN_MILLION))
def test_ten_million_fixed_rays(ctx, number, earth_rgba):
    """BASELINE.json's size and contract: 10^7 fixed rays per scene (half camera rays, half secondary rays leaving
    the surfaces the first half hit), EVERY ONE compared with the oracle's hit() — primitive id bit-exact outside
    grazing ties, t / p / normal / u / v within 1e-5 relative — plus two size-independent properties of the answers
    (p = o + t d, and nothing closer than the reported hit). The counts and the largest errors go to r2_parity.json."""
    from tests._record import record
    n = 10_000_000
    rng = np.random.default_rng(0xF17ED + number)
    gsc = R.DeviceScene(ctx, R.BuiltinDesc(number))
    osc = O.OracleScene.builtin(number, earth=earth_rgba)
    n_prims = osc.prim_count
    cam, _ = osc.camera()
    half = n // 2
    chunk = 2_500_000
    displaced = TEN_MILLION[number]
    stats = {"primary": None, "secondary": None}
    seeds_for_secondary = []

    def check_chunk(kind, rays):
        hits = trace_on_device(gsc, rays)
        ref, fragile = osc.trace(rays)
        stats[kind] = RY.merge_stats(stats[kind], RY.compare_hits(hits, ref, fragile))
        hit = hits["prim_id"] >= 0
        assert ((hits["prim_id"] >= -1) & (hits["prim_id"] < n_prims)).all()
        assert (hits["t"][hit] >= 0.001).all() and np.isfinite(hits["t"][hit]).all()
        # p = o + t d for everything except geometry under a YRotate (whose p is Q14-displaced)
        p = rays["origin"] + hits["t"][:, None] * rays["direction"]
        plain = hit if displaced is None else hit & ~((hits["prim_id"] >= displaced[0]) & (hits["prim_id"] < displaced[1]))
        assert np.allclose(p[plain], hits["p"][plain], rtol=1e-9, atol=1e-6)
        # idempotence: nothing is closer than the reported hit (media aside: their hit is a draw)
        cand = np.nonzero(hit & (hits["material"] >= 0))[0]
        sub = rng.choice(cand, min(100000, cand.shape[0]), replace=False)
        again = rays[sub].copy()
        again["t_max"] = hits["t"][sub] * (1 - 1e-9)
        again["xi"] = 1e-300  # an (essentially) infinite free flight: media never scatter
        assert (gsc.trace(again)["prim_id"] == abi.RTX_MISS).mean() > 0.9999
        return hits

    for _ in range(half // chunk):
        primary = RY.camera_rays(cam, chunk, rng)
        h1 = check_chunk("primary", primary)
        seeds_for_secondary.append((h1, primary))
    # the second half: secondary rays off the first half's hits, recycled until there are `half` of them
    done = 0
    k = 0
    while done < half:
        h1, primary = seeds_for_secondary[k % len(seeds_for_secondary)]
        k += 1
        secondary = RY.secondary_rays(h1, primary, rng)[:min(chunk, half - done)]
        assert secondary.shape[0] > 0
        check_chunk("secondary", secondary)
        done += secondary.shape[0]
    total = RY.merge_stats(stats["primary"], stats["secondary"])
    assert total["rays"] == n
    entry = {"rays": n, "primary": stats["primary"], "secondary": stats["secondary"],
             "tie_share": total["fragile"] / n, "max_tie_share_allowed": MAX_TIE_SHARE[number]}
    record("fixed_rays_10M", f"scene_{number}", entry)
    print(f"scene {number}: {entry}")
    assert total["fragile"] <= MAX_TIE_SHARE[number] * n, entry


# ---------------------------------------------------------------------------
# renders
# ---------------------------------------------------------------------------
def psnr(a_u8, b_u8):
    mse = np.mean((a_u8[..., :3].astype(np.float64) - b_u8[..., :3].astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10.0 * math.log10(255.0 ** 2 / mse)


def gpu_sum(gsc, w, h, spp, seed=1, max_depth=50, chunk=None):
    acc = gsc.new_accum(w, h)
    chunk = chunk or spp
    for b in range(0, spp, chunk):
        gsc.render_into(acc, b, min(chunk, spp - b), seed=seed, max_depth=max_depth)
    return acc


def test_furnace_and_tonemap(ctx):
    albedo, bg = (0.5, 0.25, 0.75), (0.8, 0.6, 0.4)
    cam = S.CameraDescriptor(lookfrom=(0, 0, -4), lookat=(0, 0, 0), vertical_fov=60.0)
    desc = S.Scene(S.List([S.Sphere((0, 0, 0), 1.0, S.Lambertian(albedo))]), cam, bg).to_desc()
    gsc, osc = both(ctx, desc)
    acc = gpu_sum(gsc, 16, 16, 8)
    img = acc.cpu().numpy()
    assert np.allclose(img[..., 3], 8.0)
    assert np.allclose(img[7:9, 7:9, :3] / 8, np.array(albedo) * np.array(bg), rtol=1e-5)
    assert np.allclose(img[0, 0, :3] / 8, bg, rtol=1e-5)
    ref, _ = osc.render_sum(16, 16, 8)
    assert np.allclose(img[..., :3], ref, rtol=1e-4, atol=1e-5)  # same Philox counters on both sides
    assert np.array_equal(gsc.tonemap(acc), osc.tonemap(ref, 8))


@pytest.mark.parametrize("number", [2, 3, 4, 7, 9])
def test_same_counter_low_depth_render(ctx, number, earth_rgba):
    """Same (pixel, sample, bounce) Philox counters on both sides: at depth <= 2 the per-pixel sums
    agree almost everywhere (they differ only where fp32 shading / an f64 rounding flips a branch)."""
    gsc = R.DeviceScene(ctx, R.BuiltinDesc(number))
    osc = O.OracleScene.builtin(number, earth=earth_rgba)
    w, h, spp = (64, 36, 4) if number < 6 else (48, 48, 4)
    for depth in (1, 2):
        got = gpu_sum(gsc, w, h, spp, seed=3, max_depth=depth).cpu().numpy()[..., :3]
        ref, _ = osc.render_sum(w, h, spp, seed=3, max_depth=depth)
        close = np.isclose(got, ref, rtol=2e-3, atol=2e-3).all(axis=2)
        from tests._record import record
        record("same_counter_render", f"scene_{number}_depth_{depth}", {"width": w, "height": h, "spp": spp, "pixels_agreeing": float(close.mean())})
        assert close.mean() > 0.97, (number, depth, close.mean())


RENDER_CASES = {  # scene: (width, height, spp)
    2: (64, 36, 2048), 3: (64, 36, 2048), 4: (64, 36, 2048), 1: (48, 27, 2048),
    5: (48, 27, 16384), 6: (40, 40, 16384), 7: (40, 40, 24576), 8: (32, 32, 24576), 9: (40, 40, 12288),
}


@pytest.mark.parametrize("number", sorted(RENDER_CASES))
def test_render_parity_psnr(ctx, number, earth_rgba):
    """BASELINE.json: at equal high spp, PSNR >= 35 dB and mean per-channel error <= 1/255. The numbers go to
    r2_parity.json."""
    from tests._record import record
    w, h, spp = RENDER_CASES[number]
    gsc = R.DeviceScene(ctx, R.BuiltinDesc(number))
    osc = O.OracleScene.builtin(number, earth=earth_rgba)
    acc = gpu_sum(gsc, w, h, spp, seed=11, chunk=1024)
    got = gsc.tonemap(acc)
    ref_sum, _ = osc.render_sum(w, h, spp, seed=12)  # independent samples: the comparison is statistical
    ref = osc.tonemap(ref_sum, spp)
    p = psnr(got, ref)
    err = np.abs(got[..., :3].astype(np.float64).mean(axis=(0, 1)) - ref[..., :3].astype(np.float64).mean(axis=(0, 1)))
    record("render_psnr", f"scene_{number}", {"width": w, "height": h, "spp": spp, "psnr_db": p, "mean_abs_error_per_channel_8bit": err.tolist(),
                                              "mean_abs_pixel_error_8bit": float(np.abs(got[..., :3].astype(float) - ref[..., :3].astype(float)).mean())})
    print(f"scene {number}: PSNR {p:.2f} dB, mean per-channel error {err} /255")
    assert p >= 35.0
    assert (err <= 1.0).all()


def test_render_parity_at_the_reference_resolution(ctx):
    """The Cornell box at the reference's own 600x600 (src/main.rs:137-149) against the oracle: the GPU frame's linear
    sums are box-filtered 4x4 to the 150x150 grid the oracle can afford at the same total sample count per output pixel
    (the same camera, so an oracle pixel integrates exactly the area of 16 GPU pixels), then both are tonemapped."""
    from tests._record import record
    gsc = R.DeviceScene(ctx, R.BuiltinDesc(7))
    osc = O.OracleScene.builtin(7)
    spp = 640
    acc = gpu_sum(gsc, 600, 600, spp, seed=31).cpu().numpy()
    assert (acc[..., 3] == spp).all()
    small = acc[..., :3].astype(np.float64).reshape(150, 4, 150, 4, 3).sum(axis=(1, 3))
    got = osc.tonemap(small, 16 * spp)
    ref_sum, _ = osc.render_sum(150, 150, 16 * spp, seed=32)
    ref = osc.tonemap(ref_sum, 16 * spp)
    p = psnr(got, ref)
    err = np.abs(got[..., :3].astype(np.float64).mean(axis=(0, 1)) - ref[..., :3].astype(np.float64).mean(axis=(0, 1)))
    record("render_psnr", "scene_7_600x600_vs_oracle_150x150", {"gpu": "600x600 x %d spp, linear sums box-filtered 4x4" % spp,
           "oracle": "150x150 x %d spp" % (16 * spp), "psnr_db": p, "mean_abs_error_per_channel_8bit": err.tolist()})
    print(f"scene 7 at 600x600: PSNR {p:.2f} dB, mean per-channel error {err} /255")
    assert p >= 35.0 and (err <= 1.0).all()


def test_spp_chunking_and_sharding_invariance(ctx):
    """Samples are keyed by their global index: 64 spp in one launch == 4 launches of 16 == two
    'ranks' summed (up to fp32 summation order). Different seeds differ."""
    gsc = R.DeviceScene(ctx, R.BuiltinDesc(7))
    one = gpu_sum(gsc, 40, 40, 64, seed=5).cpu().numpy()
    four = gpu_sum(gsc, 40, 40, 64, seed=5, chunk=16).cpu().numpy()
    assert np.allclose(one, four, rtol=1e-5, atol=1e-5)
    parts = []
    for r in range(2):
        b, c = R.shard_spp(64, r, 2)
        acc = gsc.new_accum(40, 40)
        gsc.render_into(acc, b, c, seed=5)
        parts.append(acc.cpu().numpy())
    assert np.allclose(parts[0] + parts[1], one, rtol=1e-5, atol=1e-5)
    other = gpu_sum(gsc, 40, 40, 64, seed=6).cpu().numpy()
    assert not np.allclose(one[..., :3], other[..., :3])


def test_ragged_image_sizes(ctx):
    """Widths / heights that are not multiples of the 8x4 warp tile."""
    gsc = R.DeviceScene(ctx, R.BuiltinDesc(2))
    osc = O.OracleScene.builtin(2)
    for (w, h) in [(1, 1), (7, 3), (9, 5), (33, 17)]:
        acc = gpu_sum(gsc, w, h, 4, seed=2, max_depth=1).cpu().numpy()
        ref, _ = osc.render_sum(w, h, 4, seed=2, max_depth=1)
        assert np.allclose(acc[..., 3], 4.0)
        assert np.isclose(acc[..., :3], ref, rtol=2e-3, atol=2e-3).mean() > 0.97


def test_reduce_tonemap_kernel_single_device(ctx):
    """The fused sum + tonemap kernel with 'peers' on the same device (the NVLink path is
    exercised by bench.py --gpus N)."""
    import ctypes as C
    import torch
    gsc = R.DeviceScene(ctx, R.BuiltinDesc(7))
    accs = []
    for r in range(3):
        b, c = R.shard_spp(48, r, 3)
        a = gsc.new_accum(32, 32)
        gsc.render_into(a, b, c, seed=9)
        accs.append(a)
    total = (accs[0] + accs[1] + accs[2])
    expect = gsc.tonemap(total)
    out = torch.zeros((32, 32, 4), dtype=torch.uint8, device="cuda")
    peers = (C.c_void_p * 2)(accs[1].data_ptr(), accs[2].data_ptr())
    abi.check(ctx.lib.rtx_reduce_tonemap_peers(ctx.h, accs[0].data_ptr(), peers, 2, 32, 32, out.data_ptr()))
    ctx.sync()
    assert np.array_equal(out.cpu().numpy(), expect)
    assert torch.allclose(accs[0], total)


# ---------------------------------------------------------------------------
# the CLI (`cargo run --release -- <scene>`, src/main.rs:236-258)
# ---------------------------------------------------------------------------
CLI = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rttnw_b200", "lib", "rttnw")


def run_cli(args, cwd):
    import subprocess
    return subprocess.run([CLI] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True, timeout=300)


def test_cli_writes_image_png_like_the_reference(tmp_path, ctx):
    from PIL import Image
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.symlink(os.path.join(root, "assets"), tmp_path / "assets")  # ImageTexture::new("assets/earth.png") is CWD-relative
    r = run_cli([4, "--spp", 16, "--width", 80, "--height", 45], tmp_path)
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines()[0] == "Scene number: 4" and "Running scene earth" in r.stdout
    img = np.asarray(Image.open(tmp_path / "image.png"))
    assert img.shape == (45, 80, 4) and (img[..., 3] == 255).all()
    # the same render through the library, same seeds: the same pixels (up to the order of the fp32 atomics)
    gsc = R.DeviceScene(ctx, R.BuiltinDesc(4))
    acc = gsc.new_accum(80, 45)
    gsc.render_into(acc, 0, 16, seed=1)
    assert np.abs(gsc.tonemap(acc).astype(int) - img.astype(int)).max() <= 1
    # usage and unknown scenes behave like main.rs:238-251,179-182
    r = run_cli([], tmp_path)
    assert r.returncode != 0 and "Usage:" in r.stderr and "9: final_scene" in r.stderr
    assert run_cli([12], tmp_path).returncode != 0


def test_cli_checkpoint_and_resume(tmp_path):
    """a killed render is resumed from its fp32 accumulator; samples are keyed by their
    global index, so the result equals the uninterrupted render up to fp32 summation order."""
    from PIL import Image
    common = [7, "--spp", 48, "--width", 64, "--height", 64, "--chunk", 8]
    r = run_cli(common + ["--out", "full.png"], tmp_path)
    assert r.returncode == 0, r.stderr
    r = run_cli(common + ["--out", "part.png", "--checkpoint", "ck", "--stop-after-chunks", 3], tmp_path)
    assert r.returncode == 3 and os.path.exists(tmp_path / "ck.0") and not os.path.exists(tmp_path / "part.png")
    r = run_cli(common + ["--out", "resumed.png", "--resume", "ck", "--checkpoint", "ck"], tmp_path)
    assert r.returncode == 0, r.stderr
    full = np.asarray(Image.open(tmp_path / "full.png")).astype(int)
    resumed = np.asarray(Image.open(tmp_path / "resumed.png")).astype(int)
    assert np.abs(full - resumed).max() <= 1
    # a checkpoint of another render is refused
    r = run_cli([7, "--spp", 64, "--width", 64, "--height", 64, "--chunk", 8, "--resume", "ck"], tmp_path)
    assert r.returncode == 1 and "not a checkpoint of this render" in r.stderr


# ---------------------------------------------------------------------------
# device-side BVH construction (replaces BvhTree::new; topology is not part of the contract)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("builder", ["lbvh", "ploc"])
@pytest.mark.parametrize("number", [1, 2, 7, 8, 9])
def test_device_built_bvh_fixed_rays(number, builder, earth_rgba):
    """The same closest-hit answers whichever builder made the tree: host binned SAH, device LBVH or device PLOC
    (scenes 1 and 9 hold moving spheres: their boxes cover the whole shutter interval under every builder)."""
    c = R.Context(0)
    try:
        c.set_bvh_builder(builder)
        gsc = R.DeviceScene(c, R.BuiltinDesc(number))
        c.set_bvh_builder("sah")
        ref_sc = R.DeviceScene(c, R.BuiltinDesc(number))
        osc = O.OracleScene.builtin(number, earth=earth_rgba)
        rng = np.random.default_rng(0xB0 + number)
        cam, _ = osc.camera()
        primary = RY.camera_rays(cam, 200000, rng)
        got, ref, st = check_rays(gsc, osc, primary)
        secondary = RY.secondary_rays(ref, primary, rng)
        got2, _, _ = check_rays(gsc, osc, secondary, min_ok=0.95)
        # and against the host-built tree, ray for ray (ties aside, the answers are the same records)
        same = ref_sc.trace(secondary)
        agree = (same["prim_id"] == got2["prim_id"]) & np.isclose(same["t"], got2["t"], rtol=1e-12, atol=0)
        assert agree.mean() > 0.95
        info = gsc.info()
        assert info["bvh_nodes"] >= 1
    finally:
        c.close()


@pytest.mark.parametrize("builder", ["lbvh", "ploc"])
def test_device_built_bvh_random_trees_and_render(builder):
    c = R.Context(0)
    try:
        c.set_bvh_builder(builder)
        for seed in range(4):
            rng = np.random.default_rng(3000 + seed)
            desc = S.Scene(random_tree(rng)).to_desc()
            gsc, osc = both(c, desc, bvh_seed=seed)
            n = 30000
            o = rng.uniform(-40, 40, (n, 3))
            target = rng.uniform(-15, 15, (n, 3))
            rays = O.make_rays(o, target - o, time=rng.random(n), xi=rng.random(n) * 0.999 + 0.0005)
            check_rays(gsc, osc, rays)
        # a render through the device-built tree: same counters, same image as the oracle at depth 2
        gsc = R.DeviceScene(c, R.BuiltinDesc(9))
        osc = O.OracleScene.builtin(9)
        got = gpu_sum(gsc, 48, 48, 4, seed=3, max_depth=2).cpu().numpy()[..., :3]
        ref, _ = osc.render_sum(48, 48, 4, seed=3, max_depth=2)
        assert np.isclose(got, ref, rtol=2e-3, atol=2e-3).all(axis=2).mean() > 0.97
    finally:
        c.close()


def test_image_texture_on_rectangles_and_boxes(ctx, earth_rgba):
    """Rectangle u, v (hittable.rs:515-516) reach an ImageTexture — also on the faces of a wrapped Cube, whose
    hit record comes from the face's rectangle record — and an image-emitting light. Same counters as the oracle."""
    img = S.ImageTexture(earth_rgba[::8, ::8].copy())
    cam = S.CameraDescriptor(lookfrom=(0, 3, -9), lookat=(0, 1, 0), vertical_fov=50.0)
    world = S.List([
        S.XZ.rectangle(S.Lambertian(img), (-6, 6), (-6, 6), 0.0),
        S.XY.rectangle(S.DiffuseLight(img), (-3, 3), (0.5, 3.5), 4.0),
        S.Cube((-1, 0, -1), (1, 2, 1), S.Lambertian(img)).rotate_y(30.0).translate((1.5, 0.0, 0.5)),
        S.YZ.rectangle(S.Lambertian(S.CheckerTexture(img, (0.2, 0.3, 0.9))), (0, 3), (-3, 3), -4.0),
    ])
    desc = S.Scene(world, cam, (0.05, 0.05, 0.05)).to_desc()
    gsc, osc = both(ctx, desc)
    for depth in (1, 2):
        got = gpu_sum(gsc, 64, 48, 4, seed=5, max_depth=depth).cpu().numpy()[..., :3]
        ref, _ = osc.render_sum(64, 48, 4, seed=5, max_depth=depth)
        close = np.isclose(got, ref, rtol=2e-3, atol=2e-3).all(axis=2)
        assert close.mean() > 0.97, (depth, close.mean())
    assert got.std() > 0.01  # the texture is really there


def test_full_size_final_scene_properties(ctx, earth_rgba):
    """BASELINE.json's frame (scene 9, 800x800) at a spp the GPU renders in a blink, checked through
    size-independent properties: every pixel got exactly its samples, nothing is NaN / negative, splitting the
    samples over two calls changes nothing but the summation order, and the frame's mean colour agrees with the
    oracle's render of a 10x smaller frame of the same camera (a different, much sparser set of paths)."""
    gsc = R.DeviceScene(ctx, R.BuiltinDesc(9))
    w = h = 800
    spp = 64
    one = gpu_sum(gsc, w, h, spp, seed=21).cpu().numpy()
    two = gpu_sum(gsc, w, h, spp, seed=21, chunk=24).cpu().numpy()
    assert np.array_equal(one[..., 3], np.full((h, w), float(spp), dtype=np.float32))
    assert np.isfinite(one).all() and (one[..., :3] >= 0).all()
    assert np.allclose(one, two, rtol=2e-5, atol=2e-4)
    osc = O.OracleScene.builtin(9, earth=earth_rgba)
    ref, _ = osc.render_sum(80, 80, 256, seed=22)
    mean_gpu = one[..., :3].reshape(-1, 3).mean(axis=0) / spp
    mean_ref = ref.reshape(-1, 3).mean(axis=0) / 256
    assert np.allclose(mean_gpu, mean_ref, rtol=0.03), (mean_gpu, mean_ref)
    rgba = gsc.tonemap(gsc.new_accum(w, h) + torch_from(one, ctx))
    assert rgba.shape == (h, w, 4) and rgba[..., 3].min() == 255


def torch_from(arr, ctx):
    import torch
    return torch.from_numpy(arr).to(f"cuda:{ctx.device}")


def test_bad_arguments_are_refused(ctx):
    import ctypes as C
    gsc = R.DeviceScene(ctx, R.BuiltinDesc(2))
    acc = gsc.new_accum(8, 8)
    lib = ctx.lib
    for bad in (abi.RenderParams(0, 8, 0, 1, 50, 0, 1), abi.RenderParams(8, 8, -1, 1, 50, 0, 1), abi.RenderParams(8, 8, 0, -1, 50, 0, 1),
                abi.RenderParams(8, 8, 0, 1, -1, 0, 1), abi.RenderParams(8, 8, 0, (1 << 24) + 1, 50, 0, 1)):
        assert lib.rtx_render(ctx.h, gsc.h, C.byref(bad), acc.data_ptr(), None) == -1
    assert lib.rtx_render(ctx.h, gsc.h, None, acc.data_ptr(), None) == -1
    assert lib.rtx_trace_rays(ctx.h, gsc.h, -1, None, None) == -1
    assert lib.rtx_ctx_set_bvh_builder(ctx.h, 7) == -1
    # zero samples and depth 0 are valid requests: nothing traced, counts move by spp
    gsc.render_into(acc, 0, 0)
    assert float(acc.sum()) == 0.0
    gsc.render_into(acc, 0, 5, max_depth=0)
    a = acc.cpu().numpy()
    assert (a[..., 3] == 5).all() and (a[..., :3] == 0).all()
    h = C.c_void_p()
    assert lib.rtx_ctx_create(99, None, C.byref(h)) == -1 and b"no such CUDA device" in lib.rtx_last_error()


def test_final_scene_matches_the_reference_shipped_image(ctx):
    """The reference's own 800x800 render of the final scene (image.png, ~10k spp; committed 8x8
    box-filtered as tests/golden/final_ref_100.npy) against this backend's render of scene 9. The ground
    boxes and the sphere cluster are drawn from an unseeded RNG in the reference, so the lower third of
    the frame only agrees in the mean; everything else — camera, fog, light, glass, metal, earth and
    marble spheres, gamma — is deterministic and agrees block by block."""
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "final_ref_100.npy")).astype(np.float64)
    gsc = R.DeviceScene(ctx, R.BuiltinDesc(9))
    acc = gpu_sum(gsc, 800, 800, 4096, seed=4, chunk=512)  # (sqrt gamma per pixel biases dark regions low at small spp)
    rgb = gsc.tonemap(acc)[..., :3].astype(np.float64)
    ours = rgb.reshape(100, 8, 100, 8, 3).mean(axis=(1, 3))
    diff = ours - ref
    assert np.abs(diff.mean(axis=(0, 1))).max() < 1.5, diff.mean(axis=(0, 1))  # whole-frame mean colour, 8-bit units
    blocks = np.abs(diff.reshape(10, 10, 10, 10, 3).mean(axis=(1, 3))).mean(axis=2)  # 80x80-pixel regions
    upper = blocks[:5]  # no random geometry above the horizon of the ground boxes
    assert np.median(upper) < 0.8 and upper.max() < 7.0, np.round(blocks, 1)
    assert np.corrcoef(ours.mean(axis=2).ravel(), ref.mean(axis=2).ravel())[0, 1] > 0.85


def test_cornell_box_matches_the_reference_shipped_image(ctx):
    """cornel_box.png (the reference's render of its deterministic Cornell box, 600x600 at ~200 spp; committed
    8x8 box-filtered) against this backend's render at the same size and spp — same sqrt-gamma bias, so the
    comparison is region by region in 8-bit units, including the nearly black rotated-box faces that only
    YRotate's sequential-update behaviour (hittable.rs:700-705, Q14) produces."""
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cornell_ref_75.npy")).astype(np.float64)
    gsc = R.DeviceScene(ctx, R.BuiltinDesc(7))
    rgb = gsc.tonemap(gpu_sum(gsc, 600, 600, 200, seed=8))[..., :3].astype(np.float64)
    ours = rgb.reshape(75, 8, 75, 8, 3).mean(axis=(1, 3))
    diff = ours - ref
    assert abs(diff.mean()) < 1.0
    blocks = diff.reshape(15, 5, 15, 5, 3).mean(axis=(1, 3))
    assert np.abs(blocks).max() < 14.0 and np.sqrt((blocks ** 2).mean()) < 4.0, np.round(np.abs(blocks).mean(axis=2), 1)
    assert ours[37:55, 27:33].mean() < 25.0 and ours[12:15, 32:42].mean() > 120.0  # dark side faces, bright top


@pytest.mark.gpu
@pytest.mark.parametrize("number", [1, 7, 8, 9])
def test_wide_bvh_traversal_finds_the_same_hits(number, monkeypatch):
    """RTX_BVH_WIDE=1 (experimental): the trace kernel walks the 4-wide copy of the world BVH. The closest hit of a
    ray does not depend on the tree it was found through, so the same rays are traced and the same samples land in
    the same pixels (the number of primitive tests may differ: the pruning order does)."""
    plain = R.Context(0)
    monkeypatch.setenv("RTX_BVH_WIDE", "1")
    wide = R.Context(0)
    try:
        out = []
        for c in (plain, wide):
            gsc = R.DeviceScene(c, R.BuiltinDesc(number))
            acc = gsc.new_accum(160, 120)
            st = gsc.render_counted(acc, 0, 24, seed=11, max_depth=50)
            out.append((acc.cpu().numpy(), st))
            gsc.close()
        (a, sa), (b, sb) = out
        assert sa["rays"] == sb["rays"]
        assert (a[..., 3] == 24).all() and (b[..., 3] == 24).all()
        assert np.allclose(a, b, rtol=1e-4, atol=1e-4)  # fp32 atomics: the summation order differs
        assert sb["node_visits"] < sa["node_visits"] * 1.35  # (counted in 64-byte entries: two per wide step)
    finally:
        plain.close()
        wide.close()


# ---------------------------------------------------------------------------
# round 2: the other kernel forms, the spread-out combine, NCCL behind the ABI, one process per GPU
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("form", ["RTX_TRACE=2", "RTX_TRACE=2 RTX_T_REFILL=33 RTX_T_LEAF=33 RTX_T_BURST=100000", "RTX_TRACE=3 RTX_T_BURST=4",
                                  "RTX_TRACE=4", "RTX_SHADE=1", "RTX_PERLIN_SMEM=0", "RTX_ORDER=1", "RTX_ORDER=1 RTX_ORDER_GROUPS=4096"])
@pytest.mark.parametrize("number", [3, 7, 9])
def test_other_kernel_forms_trace_the_same_rays(form, number, monkeypatch, earth_rgba):
    """Every opt-in form of the trace kernel (shared-memory BVH + persistent voted warps, sorted through shared memory,
    shared stack + 32-byte loads) and the first form of the shade kernel: the same rays, the same samples in the same
    pixels as the default pair (closest hits do not depend on the traversal order; fp32 atomics reorder the sums) —
    and the default pair is what every other test compares with the oracle."""
    base = R.Context(0)
    for kv in form.split():
        k, v = kv.split("=")
        monkeypatch.setenv(k, v)
    other = R.Context(0)
    try:
        out = []
        for c in (base, other):
            gsc = R.DeviceScene(c, R.BuiltinDesc(number))
            acc = gsc.new_accum(160, 120)
            st = gsc.render_counted(acc, 0, 24, seed=11, max_depth=50)
            out.append((acc.cpu().numpy(), st))
            gsc.close()
        (a, sa), (b, sb) = out
        assert sa["rays"] == sb["rays"] and sa["rays"] > 0
        assert (a[..., 3] == 24).all() and (b[..., 3] == 24).all()
        assert np.allclose(a, b, rtol=1e-4, atol=1e-4)
    finally:
        base.close()
        other.close()


def test_reduce_tonemap_slice_single_device(ctx):
    """rtx_reduce_tonemap_slice with every 'rank' on one device: the slices tile the frame, nothing is written back."""
    import ctypes as C
    import torch
    gsc = R.DeviceScene(ctx, R.BuiltinDesc(7))
    accs = []
    for r in range(3):
        b, c = R.shard_spp(48, r, 3)
        a = gsc.new_accum(33, 31)  # 1023 pixels: the slices are ragged
        gsc.render_into(a, b, c, seed=9)
        accs.append(a)
    keep = [a.clone() for a in accs]
    expect = gsc.tonemap(accs[0] + accs[1] + accs[2])
    out = torch.zeros((31, 33, 4), dtype=torch.uint8, device="cuda")
    ptrs = (C.c_void_p * 3)(*[a.data_ptr() for a in accs])
    for r in range(3):
        abi.check(ctx.lib.rtx_reduce_tonemap_slice(ctx.h, ptrs, 3, r, 33, 31, out.data_ptr()))
    ctx.sync()
    assert np.array_equal(out.cpu().numpy(), expect)
    assert all(torch.equal(a, k) for a, k in zip(accs, keep))
    assert ctx.lib.rtx_reduce_tonemap_slice(ctx.h, ptrs, 3, 3, 33, 31, out.data_ptr()) == -1  # rank out of range


def test_nccl_reduce_behind_the_abi_one_rank(ctx):
    """rtx_comm_* / rtx_accum_reduce with a communicator of one rank (the N-rank path runs in bench.py --combine nccl):
    libnccl is found, the reduce runs on the ctx stream and leaves the sum (= the accumulator) in place."""
    import ctypes as C
    lib = ctx.lib
    ident = (C.c_uint8 * 128)()
    rc = lib.rtx_comm_unique_id(C.byref(ident))
    if rc == -5:
        pytest.skip("no libnccl.so.2 on this host")
    abi.check(rc)
    comm = C.c_void_p()
    abi.check(lib.rtx_comm_create(ctx.h, 1, 0, C.byref(ident), C.byref(comm)))
    gsc = R.DeviceScene(ctx, R.BuiltinDesc(2))
    acc = gsc.new_accum(40, 24)
    gsc.render_into(acc, 0, 8, seed=2)
    before = acc.clone()
    abi.check(lib.rtx_accum_reduce(ctx.h, comm, acc.data_ptr(), 40, 24, 0))
    ctx.sync()
    import torch
    assert torch.equal(acc, before)
    abi.check(lib.rtx_comm_destroy(comm))


def test_cli_one_process_per_gpu(tmp_path):
    """`rttnw --gpus 2`: the parent forks rank 1 before touching CUDA, the ranks render their shares of the sample
    indices, the parent combines over CUDA IPC. Same image as one GPU (up to fp32 summation order)."""
    import torch
    from PIL import Image
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    common = [7, "--spp", 32, "--width", 96, "--height", 96]
    assert run_cli(common + ["--out", "one.png"], tmp_path).returncode == 0
    r = run_cli(common + ["--gpus", 2, "--out", "two.png"], tmp_path)
    assert r.returncode == 0, r.stderr
    one = np.asarray(Image.open(tmp_path / "one.png")).astype(int)
    two = np.asarray(Image.open(tmp_path / "two.png")).astype(int)
    assert np.abs(one - two).max() <= 1


def test_asynchronous_render_returns_at_once_and_gives_the_same_frame():
    """rtx_ctx_set_async: rtx_render hands the wavefront driver's loop to the context's own thread and returns
    immediately; rtx_ctx_sync (and every other call on the context) waits for it. Same frame as the synchronous call."""
    import ctypes as C
    import time
    import torch
    c = R.Context(0)
    try:
        gsc = R.DeviceScene(c, R.BuiltinDesc(9))
        w = h = 800
        sync_acc = gsc.new_accum(w, h)
        gsc.render_into(sync_acc, 0, 64, seed=3)  # also brings the pool to its size for this job
        c.sync()
        sync_acc.zero_()
        gsc.render_into(sync_acc, 0, 16, seed=3)
        c.sync()
        t0 = time.perf_counter()
        gsc.render_into(sync_acc, 16, 64, seed=3)
        t_sync_call = time.perf_counter() - t0
        c.sync()
        abi.check(c.lib.rtx_ctx_set_async(c.h, 1))
        async_acc = gsc.new_accum(w, h)
        gsc.render_into(async_acc, 0, 16, seed=3)
        c.sync()
        t0 = time.perf_counter()
        gsc.render_into(async_acc, 16, 64, seed=3)
        t_async_call = time.perf_counter() - t0
        c.sync()  # waits for the worker thread, then for the stream
        t_async_total = time.perf_counter() - t0
        from tests._record import record
        record("async_render", "scene_9_800x800_64spp", {"sync_call_ms": 1e3 * t_sync_call, "async_call_ms": 1e3 * t_async_call,
                                                         "async_call_plus_sync_ms": 1e3 * t_async_total})
        assert t_async_call < 0.005 and t_async_call < 0.1 * t_sync_call, (t_async_call, t_sync_call)
        assert t_async_total > 0.5 * t_sync_call
        assert torch.allclose(sync_acc, async_acc, rtol=1e-4, atol=1e-3)
        # an invalid request fails at once (argument checks run in the caller), a valid one after an error still works
        bad = abi.RenderParams(0, 8, 0, 1, 50, 0, 1)
        assert c.lib.rtx_render(c.h, gsc.h, C.byref(bad), async_acc.data_ptr(), None) == 0  # queued ...
        assert c.lib.rtx_ctx_sync(c.h) == -1  # ... and reported by the call that joins
        abi.check(c.lib.rtx_ctx_sync(c.h))
        abi.check(c.lib.rtx_ctx_set_async(c.h, 0))
        gsc.close()
    finally:
        c.close()
