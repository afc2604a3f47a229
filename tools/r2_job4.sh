#!/bin/bash
# round 2, job 4: shade2 (media after the surface search, early dispenser request): parity subset, then A/B
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "not ten_million" 2>&1 | tail -15 ) > gpurun_out/j4_pytest.log
( timeout 600 python tools/quick_ab.py --prof \
  "RTX_TRACE=1 RTX_SHADE=1" \
  "RTX_TRACE=1 RTX_SHADE=2" \
  "RTX_TRACE=1 RTX_SHADE=2 RTX_WF_SLOTS=655360" \
  "RTX_TRACE=1 RTX_SHADE=2 RTX_WF_SLOTS=786432" \
  "RTX_TRACE=1 RTX_SHADE=2 RTX_WF_SLOTS=1048576" \
  "RTX_TRACE=1 RTX_SHADE=1 RTX_WF_STREAMS=1" \
  "RTX_TRACE=1 RTX_SHADE=2 RTX_WF_STREAMS=1" \
  "RTX_TRACE=2 RTX_SHADE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8" \
  "RTX_TRACE=2 RTX_SHADE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8 RTX_TRACE_THREADS=640" \
  "RTX_TRACE=2 RTX_SHADE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8 RTX_WF_SLOTS=1048576" \
  "RTX_TRACE=2 RTX_SHADE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8 RTX_WF_SLOTS=2097152 RTX_WF_STREAMS=1" \
  "RTX_TRACE=2 RTX_SHADE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8 RTX_WF_SLOTS=2097152" \
  2>&1 ) > gpurun_out/j4_ab.log
for sc in 1 2 3 7 8; do
( timeout 300 python tools/quick_ab.py --scene $sc --spp 256 "RTX_TRACE=1 RTX_SHADE=1" "RTX_TRACE=1 RTX_SHADE=2" "RTX_TRACE=2 RTX_SHADE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8" 2>&1 | sed "s/^/scene $sc: /" ) >> gpurun_out/j4_ab_scenes.log
done
cat gpurun_out/j4_pytest.log gpurun_out/j4_ab.log gpurun_out/j4_ab_scenes.log
