#!/bin/bash
# round 2, job 2: trace2 with interleaved slot pieces; 1- and 2-stream A/B; per-launch instruction counts of both forms
mkdir -p gpurun_out
( timeout 600 python tools/quick_ab.py --prof \
  "RTX_TRACE=1" \
  "RTX_TRACE=1 RTX_WF_STREAMS=1" \
  "RTX_TRACE=2" \
  "RTX_TRACE=2 RTX_WF_STREAMS=1" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8 RTX_WF_STREAMS=1" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=12 RTX_T_BURST=8" \
  "RTX_TRACE=2 RTX_T_REFILL=12 RTX_T_LEAF=8 RTX_T_BURST=8" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=16" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8 RTX_TRACE_THREADS=1024" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8 RTX_TRACE_THREADS=1024 RTX_WF_STREAMS=1" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8 RTX_TRACE_THREADS=640" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8 RTX_TRACE_THREADS=512" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8 RTX_WF_SLOTS=1048576" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8 RTX_WF_SLOTS=1048576 RTX_WF_STREAMS=1" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8 RTX_WF_SLOTS=2097152 RTX_WF_STREAMS=1" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8 RTX_WF_STREAMS=3 RTX_WF_SLOTS=786432" \
  2>&1 ) > gpurun_out/j2_ab.log
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_active.min,gpu__time_duration.sum,sm__cycles_elapsed.max
for v in "RTX_TRACE=1" "RTX_TRACE=2" "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8" "RTX_TRACE=2 RTX_T_REFILL=8 RTX_T_LEAF=4 RTX_T_BURST=4"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v RTX_WF_STREAMS=1 timeout 300 ncu --metrics $M --clock-control none -k regex:wf_ -s 240 -c 24 --csv --log-file gpurun_out/j2_m_$tag.csv \
    python tools/quick_ab.py --spp 32 --reps 1 "$v RTX_WF_STREAMS=1" > gpurun_out/j2_m_$tag.log 2>&1
done
cat gpurun_out/j2_ab.log
