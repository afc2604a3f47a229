#!/bin/bash
# shade2: deferred dispenser compare + early dispenser request, against the previous build on the same box
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "other_kernel_forms or same_counter or furnace or render_matches" > gpurun_out/j21_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j21_tests.log
for s in 9 7 3; do
  for rep in 1 2; do
    timeout 300 python tools/quick_ab.py --scene $s --spp 64 --reps 5 --lib old_lib/librttnw_b200_old.so "RTX_X=0" >> gpurun_out/j21_ab.log 2>&1
    timeout 300 python tools/quick_ab.py --scene $s --spp 64 --reps 5 "RTX_EARLY_ASK=0" "RTX_EARLY_ASK=1" >> gpurun_out/j21_ab.log 2>&1
  done
done
tail -5 gpurun_out/j21_tests.log
grep -E "scene|M samples|RTX_" gpurun_out/j21_ab.log | tail -60
