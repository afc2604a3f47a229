#!/bin/bash
# round 2, job 5: shade2 (media after the surface search; same-pass refill) and the leaner trace2 loop
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "not ten_million" 2>&1 | tail -5 ) > gpurun_out/j5_pytest.log
T2="RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=8"
( timeout 600 python tools/quick_ab.py --prof \
  "RTX_TRACE=1 RTX_SHADE=1" \
  "RTX_TRACE=1 RTX_SHADE=2" \
  "RTX_TRACE=1 RTX_SHADE=1 RTX_WF_STREAMS=1" \
  "RTX_TRACE=1 RTX_SHADE=2 RTX_WF_STREAMS=1" \
  "$T2 RTX_SHADE=1" \
  "$T2 RTX_SHADE=2" \
  "$T2 RTX_SHADE=2 RTX_WF_STREAMS=1" \
  "$T2 RTX_SHADE=2 RTX_TRACE_THREADS=640" \
  "$T2 RTX_SHADE=2 RTX_WF_SLOTS=1048576" \
  "$T2 RTX_SHADE=2 RTX_WF_SLOTS=2097152 RTX_WF_STREAMS=1" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=4 RTX_SHADE=2" \
  "RTX_TRACE=2 RTX_T_REFILL=16 RTX_T_LEAF=8 RTX_T_BURST=16 RTX_SHADE=2" \
  "RTX_TRACE=2 RTX_T_REFILL=12 RTX_T_LEAF=6 RTX_T_BURST=8 RTX_SHADE=2" \
  "RTX_TRACE=2 RTX_T_REFILL=20 RTX_T_LEAF=12 RTX_T_BURST=8 RTX_SHADE=2" \
  2>&1 ) > gpurun_out/j5_ab.log
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_active.avg,gpu__time_duration.sum,sm__cycles_elapsed.max
rm -f gpurun_out/j5_sum.txt
i=0
for v in "RTX_TRACE=1 RTX_SHADE=2" "$T2 RTX_SHADE=2"; do
  i=$((i+1))
  env $v RTX_WF_STREAMS=1 timeout 400 ncu --metrics $M --clock-control none -k regex:wf_ --csv --log-file gpurun_out/j5_m$i.csv \
    python tools/quick_ab.py --spp 16 --warm 1 --reps 1 "$v RTX_WF_STREAMS=1" > gpurun_out/j5_m$i.log 2>&1
  echo "== $v" >> gpurun_out/j5_sum.txt
  python tools/ncu_sum.py gpurun_out/j5_m$i.csv >> gpurun_out/j5_sum.txt
done
cat gpurun_out/j5_pytest.log gpurun_out/j5_ab.log gpurun_out/j5_sum.txt
