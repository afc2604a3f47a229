#!/bin/bash
mkdir -p gpurun_out
( RTX_TRACE=3 timeout 900 python -m pytest tests -m gpu -x -q -k "not ten_million" 2>&1 | tail -15 ) > gpurun_out/j10_pytest.log
( timeout 600 python tools/quick_ab.py --prof \
  "RTX_TRACE=1 RTX_SHADE=2" \
  "RTX_TRACE=3 RTX_SHADE=2 RTX_T_BURST=4" \
  "RTX_TRACE=3 RTX_SHADE=2 RTX_T_BURST=2" \
  "RTX_TRACE=3 RTX_SHADE=2 RTX_T_BURST=8" \
  "RTX_TRACE=3 RTX_SHADE=2 RTX_T_BURST=4 RTX_TRACE_THREADS=768" \
  "RTX_TRACE=3 RTX_SHADE=2 RTX_T_BURST=4 RTX_TRACE_THREADS=640" \
  "RTX_TRACE=3 RTX_SHADE=2 RTX_T_BURST=4 RTX_WF_STREAMS=1" \
  "RTX_TRACE=3 RTX_SHADE=2 RTX_T_BURST=4 RTX_WF_SLOTS=1048576" \
  "RTX_TRACE=3 RTX_SHADE=2 RTX_T_BURST=4 RTX_WF_SLOTS=2097152 RTX_WF_STREAMS=1" \
  2>&1 ) > gpurun_out/j10_ab.log
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_active.avg,gpu__time_duration.sum,sm__cycles_elapsed.max
rm -f gpurun_out/j10_sum.txt
i=0
for v in "RTX_TRACE=3 RTX_SHADE=2 RTX_T_BURST=4" "RTX_TRACE=3 RTX_SHADE=2 RTX_T_BURST=2"; do
  i=$((i+1))
  env $v RTX_WF_STREAMS=1 timeout 400 ncu --metrics $M --clock-control none -k regex:wf_ --csv --log-file gpurun_out/j10_m$i.csv \
    python tools/quick_ab.py --spp 16 --warm 1 --reps 1 "$v RTX_WF_STREAMS=1" > gpurun_out/j10_m$i.log 2>&1
  echo "== $v" >> gpurun_out/j10_sum.txt
  python tools/ncu_sum.py gpurun_out/j10_m$i.csv >> gpurun_out/j10_sum.txt
done
cat gpurun_out/j10_pytest.log gpurun_out/j10_ab.log gpurun_out/j10_sum.txt
