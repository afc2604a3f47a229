#!/bin/bash
mkdir -p gpurun_out
for v in "RTX_TRACE=1 RTX_SHADE=1" "RTX_TRACE=1 RTX_SHADE=2"; do
  echo "== $v"
  env RTX_DEBUG_BATCHES=1 RTX_WF_STREAMS=1 python tools/quick_ab.py --spp 16 --warm 1 --reps 1 "$v RTX_WF_STREAMS=1 RTX_DEBUG_BATCHES=1" 2>&1 | tail -60
done > gpurun_out/j6.log 2>&1
cat gpurun_out/j6.log
