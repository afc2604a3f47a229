#!/bin/bash
# dispenser tile order in 64x64 blocks: parity, then A/B against the row-major build on the same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "tile_shards or same_counter or render_matches or other_kernel_forms" > gpurun_out/j27_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j27_tests.log
tail -6 gpurun_out/j27_tests.log
for s in 9 7 8 1; do
  for spp in 32 128 512; do
    timeout 300 python tools/quick_ab.py --scene $s --spp $spp --reps 3 --lib old_lib/librttnw_b200_old.so "RTX_X=rowmajor_s${s}_$spp" 2>&1 | grep RTX_X | tee -a gpurun_out/j27_order.log
    timeout 300 python tools/quick_ab.py --scene $s --spp $spp --reps 3 "RTX_X=blocks_s${s}_$spp" 2>&1 | grep RTX_X | tee -a gpurun_out/j27_order.log
  done
done
