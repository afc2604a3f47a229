#!/bin/bash
# pool slot without the radiance triple (96 B instead of 108): parity, A/B against the previous build
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "other_kernel_forms or same_counter or render_matches or furnace or emissive or light" > gpurun_out/j34_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j34_tests.log
tail -3 gpurun_out/j34_tests.log
for s in 9 7; do
  for rep in 1 2; do
  timeout 300 python tools/quick_ab.py --scene $s --spp 128 --reps 3 --lib old_lib/librttnw_b200_old.so "RTX_X=slot108_s$s" 2>&1 | grep RTX_X | tee -a gpurun_out/j34_slot.log
  timeout 300 python tools/quick_ab.py --scene $s --spp 128 --reps 3 "RTX_X=slot96_s$s" 2>&1 | grep RTX_X | tee -a gpurun_out/j34_slot.log
  done
done
