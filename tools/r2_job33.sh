#!/bin/bash
# shade kernel leaves the rays of each CTA sorted by octant (RTX_LOCAL_ORDER, default on): parity, A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "other_kernel_forms or same_counter or render_matches or furnace" > gpurun_out/j33_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j33_tests.log
tail -5 gpurun_out/j33_tests.log
for s in 9 7 8 3 1; do
  timeout 300 python tools/quick_ab.py --scene $s --spp 128 --reps 3 --prof "RTX_LOCAL_ORDER=0" "RTX_LOCAL_ORDER=1" 2>&1 | grep RTX_ | tee -a gpurun_out/j33_local.log
done
