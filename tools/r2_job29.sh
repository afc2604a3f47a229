#!/bin/bash
# block shape of the dispenser's tile order (tiles of 8x4 px): W x H tiles
mkdir -p gpurun_out
for s in 9 7 1; do
  for spp in 128 512; do
    for v in 4x32 2x64 4x64 1x128 2x32 4x16; do
      timeout 300 python tools/quick_ab.py --scene $s --spp $spp --reps 3 --lib old_lib/lib_b$v.so "RTX_X=b${v}_s${s}_$spp" 2>&1 | grep RTX_X | tee -a gpurun_out/j29_shape.log
    done
  done
done
