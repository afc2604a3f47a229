#!/usr/bin/env python
"""tools/combine_ab.py — the end-of-frame combine alone, three ways, on N GPUs (launch under torch.distributed.run):
peer (rank 0's fused reduce + tonemap kernel reads every other accumulator over NVLink), slice (every rank reduces and
tonemaps 1/N of the pixels, rtx_reduce_tonemap_slice), nccl (rtx_accum_reduce = ncclReduce behind the C ABI, then
rtx_tonemap_rgba8 on rank 0). Microseconds per combine, device-timed (events on the ctx stream, max over ranks, mean of
`--reps` after a warm-up), barriers included where the mode needs them, at 800x800 and 600x600."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=50)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import rttnw_b200 as R
    from rttnw_b200 import abi
    import bench
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = abi.load()
    ctx = R.Context(local)
    dev = torch.device("cuda", local)
    env = {"lib": lib, "ctx": ctx, "abi": abi, "dist": dist, "rank": rank, "world": world, "dev": dev, "barrier": dist.barrier}
    scene = R.DeviceScene(ctx, R.BuiltinDesc(7))
    out = {"n_gpus": world, "reps": args.reps, "unit": "us per combine (device time, max over ranks)", "sizes": {}}
    for (w, h) in ((800, 800), (600, 600)):
        row = {}
        ref = None
        for mode in ("peer", "slice", "nccl"):
            fr = bench.Frame(env, w, h, mode)
            fr.zero()
            fr.render(scene.h, rank * 4, 4, 50)
            ctx.sync()
            # peer and nccl add into rank 0's accumulator: re-render it each rep would dominate, so they run on whatever
            # the accumulators hold (the cost of the combine does not depend on the values)
            for _ in range(5):
                fr.combine_and_tonemap()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                fr.combine_and_tonemap()
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            row[mode] = 1e3 * ms.item() / args.reps
            # same frame from all three (slice is non-destructive: check it first in a fresh state)
            if mode == "slice" and rank == 0:
                ref = fr.d_rgba.clone()
        out["sizes"][f"{w}x{h}"] = row
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
