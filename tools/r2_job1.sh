#!/bin/bash
# round 2, job 1: parity of the shared-memory trace kernel (the GPU suite runs with it as the default), then the
# first A/B sweep of its knobs against the first form, then one ncu --set full capture of a mid-frame launch.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/j1_gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/j1_pytest.log
( timeout 600 python tools/quick_ab.py --prof \
  "RTX_TRACE=1" \
  "RTX_TRACE=2" \
  "RTX_TRACE=2 RTX_TRACE_THREADS=640" \
  "RTX_TRACE=2 RTX_TRACE_THREADS=512" \
  "RTX_TRACE=2 RTX_TRACE_THREADS=448" \
  "RTX_TRACE=2 RTX_T_REFILL=24" \
  "RTX_TRACE=2 RTX_T_REFILL=16" \
  "RTX_TRACE=2 RTX_T_REFILL=8" \
  "RTX_TRACE=2 RTX_T_REFILL=4" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=16 RTX_T_BURST=8" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=4" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=4 RTX_T_BURST=4" \
  "RTX_TRACE=2 RTX_T_REFILL=8 RTX_T_LEAF=8 RTX_T_BURST=4" \
  "RTX_TRACE=2 RTX_T_REFILL=8 RTX_T_LEAF=4 RTX_T_BURST=2" \
  "RTX_TRACE=2 RTX_WF_STREAMS=1" \
  "RTX_TRACE=2 RTX_WF_STREAMS=1 RTX_WF_SLOTS=1048576" \
  "RTX_TRACE=2 RTX_WF_STREAMS=3" \
  "RTX_TRACE=2 RTX_WF_SLOTS=1048576" \
  "RTX_TRACE=2 RTX_WF_SLOTS=1048576 RTX_T_REFILL=16" \
  "RTX_TRACE=2 RTX_TRACE_THREADS=640 RTX_T_REFILL=16" \
  "RTX_TRACE=1" \
  2>&1 ) > gpurun_out/j1_ab.log
( timeout 300 python tools/quick_ab.py --scene 7 --spp 256 "RTX_TRACE=1" "RTX_TRACE=2" "RTX_TRACE=2 RTX_T_REFILL=16" 2>&1 ) > gpurun_out/j1_ab_s7.log
( timeout 300 python tools/quick_ab.py --scene 1 --spp 256 "RTX_TRACE=1" "RTX_TRACE=2" "RTX_TRACE=2 RTX_T_REFILL=16" 2>&1 ) > gpurun_out/j1_ab_s1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wf_trace2 -s 150 -c 1 -f -o gpurun_out/j1_trace2 \
  python tools/quick_ab.py --spp 64 --reps 1 "RTX_TRACE=2" > gpurun_out/j1_ncu.log 2>&1
cat gpurun_out/j1_pytest.log gpurun_out/j1_ab.log gpurun_out/j1_ab_s7.log gpurun_out/j1_ab_s1.log
