#!/usr/bin/env python
"""tools/quick_ab.py — torch-free A/B of library configurations on one GPU, inside ONE process and ONE gpurun call.

usage: quick_ab.py [--scene 9] [--spp 128] [--reps 3] [--lib PATH] "VAR=a VAR2=b" "VAR=c" ...

Every argument is one configuration: environment variables the library reads in rtx_ctx_create (RTX_TRACE,
RTX_TRACE_THREADS, RTX_T_LEAF, RTX_T_REFILL, RTX_T_BURST, RTX_WF_STREAMS, RTX_WF_SLOTS, ...). Per configuration:
a fresh context and scene, one warm-up render, `reps` timed renders of `spp` samples per pixel at the scene's default
resolution (host clock around rtx_render + rtx_ctx_sync; a render is >= 100 ms), the per-kernel times of the event-
bracketed launches (rtx_ctx_set_profiling) and the frame mean as a checksum (same seeds: equal up to summation order).
Prints one line per configuration: best and median M samples/s."""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", type=int, default=9)
    ap.add_argument("--spp", type=int, default=128)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--lib", default=None)
    ap.add_argument("--warm", type=int, default=0, help="spp of the warm-up render (default spp / 4, at least 8)")
    ap.add_argument("--counted", action="store_true", help="also run the counting build once and print the per-ray means")
    ap.add_argument("--prof", action="store_true", help="bracket every 8th iteration with CUDA events (costs ~1-2 %)")
    ap.add_argument("configs", nargs="*")
    args = ap.parse_args()
    if args.lib:
        os.environ["RTTNW_B200_LIB"] = args.lib
    import numpy as np
    from rttnw_b200 import abi
    from rttnw_b200.render import BuiltinDesc, scene_defaults
    lib = abi.load()
    d = scene_defaults(args.scene)
    w, h = args.width or d["width"], args.height or d["height"]
    desc = BuiltinDesc(args.scene)
    n_px = w * h
    host = np.zeros((h, w, 4), dtype=np.float32)
    base_env = dict(os.environ)
    for cfg in (args.configs or [""]):
        for k in list(os.environ):
            if k.startswith("RTX_") and k not in base_env:
                del os.environ[k]
        for kv in cfg.split():
            k, v = kv.split("=", 1)
            os.environ[k] = v
        ctx = C.c_void_p()
        abi.check(lib.rtx_ctx_create(0, None, C.byref(ctx)))
        sc = C.c_void_p()
        abi.check(lib.rtx_scene_create(ctx, C.byref(desc.desc), C.byref(sc)))
        acc = C.c_void_p()
        abi.check(lib.rtx_malloc(ctx, n_px * 16, C.byref(acc)))
        rays = C.c_void_p()
        abi.check(lib.rtx_malloc(ctx, 8, C.byref(rays)))

        def render(k, spp):
            abi.check(lib.rtx_memset_zero(ctx, acc, n_px * 16))
            abi.check(lib.rtx_memset_zero(ctx, rays, 8))
            p = abi.RenderParams(w, h, k * spp, spp, d["max_depth"], 0, 1)
            t0 = time.perf_counter()
            abi.check(lib.rtx_render(ctx, sc, C.byref(p), acc, rays))
            abi.check(lib.rtx_ctx_sync(ctx))
            return time.perf_counter() - t0
        render(0, args.warm or max(8, args.spp // 4))
        if args.prof:
            abi.check(lib.rtx_ctx_set_profiling(ctx, 1))
        ts = sorted(render(k + 1, args.spp) for k in range(args.reps))
        a, b, n = C.c_double(), C.c_double(), C.c_uint64()
        abi.check(lib.rtx_ctx_profile_read(ctx, C.byref(a), C.byref(b), C.byref(n), 1))
        abi.check(lib.rtx_memcpy_d2h(ctx, host.ctypes.data, acc, n_px * 16))
        nr = np.zeros(1, dtype=np.uint64)
        abi.check(lib.rtx_memcpy_d2h(ctx, nr.ctypes.data, rays, 8))
        mean = (host[..., :3].sum(axis=(0, 1)) / host[..., 3].sum()).tolist()
        it = max(1, n.value)
        nl = C.c_uint64()
        abi.check(lib.rtx_ctx_kernel_launches(ctx, C.byref(nl)))
        best, med = n_px * args.spp / ts[0] / 1e6, n_px * args.spp / ts[len(ts) // 2] / 1e6
        prof = (f" shade {1e3 * a.value / it:6.1f} us trace {1e3 * b.value / it:6.1f} us/launch (kernels {a.value + b.value:.1f} ms of "
                f"{1e3 * sum(ts):.1f} ms wall)") if args.prof else ""
        print(f"{cfg or '(default)':60s} | best {best:7.1f} med {med:7.1f} Msamples/s  {float(nr[0]) / (n_px * args.spp):.3f} rays/sample{prof}"
              f"  mean rgb {mean[0]:.5f} {mean[1]:.5f} {mean[2]:.5f}  launches {nl.value}", flush=True)
        if args.counted:
            st = abi.TraceStats()
            abi.check(lib.rtx_memset_zero(ctx, acc, n_px * 16))
            p = abi.RenderParams(w, h, 0, 4, d["max_depth"], 0, 1)
            abi.check(lib.rtx_render_counted(ctx, sc, C.byref(p), acc, C.byref(st)))
            print("    per ray: " + "  ".join(f"{k} {getattr(st, k):.3f}" for k, _ in abi.TraceStats._fields_ if k != "rays"), flush=True)
        lib.rtx_free(ctx, acc)
        lib.rtx_free(ctx, rays)
        lib.rtx_scene_destroy(sc)
        lib.rtx_ctx_destroy(ctx)


if __name__ == "__main__":
    main()
