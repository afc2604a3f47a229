#!/bin/bash
# block tile order in: the whole GPU suite, then the pool size revisited (the optimum was found under row-major order)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/j31_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j31_tests.log
tail -5 gpurun_out/j31_tests.log
cp gpurun_out/r2_parity.json gpurun_out/r2_parity_full.json 2>/dev/null
for s in 9 7; do
  timeout 600 python tools/quick_ab.py --scene $s --spp 128 --reps 3 "RTX_WF_SLOTS=262144" "RTX_WF_SLOTS=393216" "RTX_WF_SLOTS=524288" "RTX_WF_SLOTS=655360" "RTX_WF_SLOTS=786432" "RTX_WF_SLOTS=1048576" "RTX_WF_SLOTS=393216 RTX_WF_STREAMS=3" "RTX_WF_SLOTS=786432 RTX_WF_STREAMS=3" 2>&1 | grep RTX_ | tee -a gpurun_out/j31_slots.log
done
