#!/bin/bash
# ray ordering (RTX_ORDER=1): parity, then A/B on scenes 9, 7, 3
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "other_kernel_forms and ORDER" > gpurun_out/j22_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j22_tests.log
tail -15 gpurun_out/j22_tests.log
for s in 9 7 3; do
  timeout 300 python tools/quick_ab.py --scene $s --spp 64 --reps 5 "RTX_ORDER=0" "RTX_ORDER=1" "RTX_ORDER=1 RTX_ORDER_GROUPS=64" "RTX_ORDER=1 RTX_ORDER_GROUPS=4096" "RTX_ORDER=1 RTX_WF_STREAMS=1" "RTX_ORDER=0 RTX_WF_STREAMS=1" --prof 2>&1 | tee -a gpurun_out/j22_ab.log | grep -E "RTX_|shade|trace"
done
