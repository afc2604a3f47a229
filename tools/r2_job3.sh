#!/bin/bash
# round 2, job 3: whole-render instruction totals of the two trace kernel forms (one stream, 16 spp), ncu metrics only
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_active.avg,gpu__time_duration.sum,sm__cycles_elapsed.max
i=0
for v in "RTX_TRACE=1" "RTX_TRACE=2" "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8" "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8 RTX_WF_SLOTS=2097152" "RTX_TRACE=1 RTX_WF_SLOTS=2097152"; do
  i=$((i+1))
  env $v RTX_WF_STREAMS=1 timeout 400 ncu --metrics $M --clock-control none -k regex:wf_ --csv --log-file gpurun_out/j3_m$i.csv \
    python tools/quick_ab.py --spp 16 --warm 1 --reps 1 "$v RTX_WF_STREAMS=1" > gpurun_out/j3_m$i.log 2>&1
  echo "== $v" >> gpurun_out/j3_sum.txt
  python tools/ncu_sum.py gpurun_out/j3_m$i.csv >> gpurun_out/j3_sum.txt
done
cat gpurun_out/j3_sum.txt
