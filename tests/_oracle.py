"""ctypes bindings of oracle/_build/liboracle.so — the CHECKER. Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rttnw_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
_P = C.c_void_p
_lib = None


def build() -> None:
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        build()
    lib = C.CDLL(LIB)
    sig = {
        "orc_scene_builtin": (_P, [C.c_int, C.c_uint64, _P, C.c_int, C.c_int]),
        "orc_scene_from_desc": (_P, [C.POINTER(abi.SceneDesc), C.c_uint64]),
        "orc_scene_free": (None, [_P]),
        "orc_scene_prim_count": (C.c_int, [_P]),
        "orc_scene_camera": (None, [_P, C.POINTER(abi.Camera), C.POINTER(C.c_double * 3)]),
        "orc_trace_rays": (None, [_P, C.c_int64, _P, _P, _P, C.c_int]),
        "orc_render": (C.c_uint64, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int,
                                    C.c_int, C.c_int, _P, C.c_int]),
        "orc_tonemap": (None, [_P, C.c_int, C.c_double, _P]),
        "orc_philox4x32_10": (None, [C.POINTER(C.c_uint32 * 4), C.POINTER(C.c_uint32 * 2), C.POINTER(C.c_uint32 * 4)]),
        "orc_sphere_uv": (None, [C.POINTER(C.c_double * 3), C.POINTER(C.c_double * 2)]),
        "orc_bound_hit": (C.c_int, [C.POINTER(C.c_double * 3), C.POINTER(C.c_double * 3), C.POINTER(abi.Ray)]),
        "orc_perlin_noise": (C.c_double, [C.POINTER(abi.Perlin), C.POINTER(C.c_double * 3)]),
        "orc_perlin_turbulence": (C.c_double, [C.POINTER(abi.Perlin), C.POINTER(C.c_double * 3), C.c_int]),
        "orc_texture_value": (None, [_P, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double * 3),
                                     C.POINTER(C.c_double * 3)]),
        "orc_perlin_generate": (None, [C.c_uint64, C.POINTER(abi.Perlin)]),
        "orc_hardware_threads": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


SCENE_SEED_BASE = 0x5254544E57  # "RTTNW" (SURVEY.md §8d); scene seed = base + scene number


class OracleScene:
    def __init__(self, handle):
        assert handle, "oracle could not build the scene"
        self.h = handle
        self.lib = load()

    @classmethod
    def builtin(cls, number: int, seed: int | None = None, earth: np.ndarray | None = None) -> "OracleScene":
        lib = load()
        seed = SCENE_SEED_BASE + number if seed is None else seed
        if earth is not None:
            earth = np.ascontiguousarray(earth, dtype=np.uint8)
            h = lib.orc_scene_builtin(number, seed, earth.ctypes.data, earth.shape[1], earth.shape[0])
        else:
            h = lib.orc_scene_builtin(number, seed, None, 0, 0)
        return cls(h)

    @classmethod
    def from_desc(cls, desc, bvh_seed: int = 7) -> "OracleScene":
        d = desc.desc if hasattr(desc, "desc") else desc
        s = cls(load().orc_scene_from_desc(C.byref(d), bvh_seed))
        s._keep = desc
        return s

    def __del__(self):
        try:
            self.lib.orc_scene_free(self.h)
        except Exception:
            pass

    @property
    def prim_count(self) -> int:
        return self.lib.orc_scene_prim_count(self.h)

    def camera(self):
        cam, bg = abi.Camera(), (C.c_double * 3)()
        self.lib.orc_scene_camera(self.h, C.byref(cam), C.byref(bg))
        return cam, tuple(bg)

    def trace(self, rays: np.ndarray, threads: int = 0):
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == np.dtype(abi.RAY_DTYPE)
        hits = np.zeros(rays.shape[0], dtype=abi.HIT_DTYPE)
        fragile = np.zeros(rays.shape[0], dtype=np.uint8)
        self.lib.orc_trace_rays(self.h, rays.shape[0], rays.ctypes.data, hits.ctypes.data, fragile.ctypes.data, threads)
        return hits, fragile.astype(bool)

    def render_sum(self, width: int, height: int, spp: int, seed: int = 1, spp_begin: int = 0, max_depth: int = 50,
                   rows=None, row_stride: int = 1, threads: int = 0):
        """Returns (rgb_sum float64 (H,W,3), n_rays)."""
        out = np.zeros((height, width, 3), dtype=np.float64)
        r0, r1 = rows if rows is not None else (0, height)
        n = self.lib.orc_render(self.h, width, height, spp_begin, spp, max_depth, seed, r0, r1, row_stride,
                                out.ctypes.data, threads)
        return out, int(n)

    def tonemap(self, rgb_sum: np.ndarray, samples: float) -> np.ndarray:
        h, w, _ = rgb_sum.shape
        out = np.zeros((h, w, 4), dtype=np.uint8)
        rgb_sum = np.ascontiguousarray(rgb_sum, dtype=np.float64)
        self.lib.orc_tonemap(rgb_sum.ctypes.data, h * w, float(samples), out.ctypes.data)
        return out


def make_rays(origin, direction, time=0.0, t_min=0.001, t_max=np.finfo(np.float64).max, xi=0.5) -> np.ndarray:
    origin = np.atleast_2d(np.asarray(origin, dtype=np.float64))
    direction = np.atleast_2d(np.asarray(direction, dtype=np.float64))
    n = max(origin.shape[0], direction.shape[0])
    rays = np.zeros(n, dtype=abi.RAY_DTYPE)
    rays["origin"], rays["direction"] = origin, direction
    rays["time"], rays["t_min"], rays["t_max"], rays["xi"] = time, t_min, t_max, xi
    return rays
