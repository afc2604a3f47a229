#!/bin/bash
mkdir -p gpurun_out
for s in 9 7 8; do
  for spp in 128 512; do
    for rep in 1 2; do
    timeout 300 python tools/quick_ab.py --scene $s --spp $spp --reps 3 --lib old_lib/librttnw_b200_old.so "RTX_X=rowmajor_s${s}_$spp" 2>&1 | grep RTX_X | tee -a gpurun_out/j28_order.log
    timeout 300 python tools/quick_ab.py --scene $s --spp $spp --reps 3 "RTX_X=blocks_s${s}_$spp" 2>&1 | grep RTX_X | tee -a gpurun_out/j28_order.log
    done
  done
done
