#!/bin/bash
# round 2, final job 1 (one GPU): the whole GPU suite (writes gpurun_out/r2_parity.json), the per-sample counters of scenes 9
# and 7, and --set full captures of the default kernel pair and of the two opt-in trace forms
mkdir -p gpurun_out; rm -f gpurun_out/r2_parity.json
( timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -14 ) > gpurun_out/f1_pytest.log
cp gpurun_out/r2_parity.json gpurun_out/r2_parity_full.json
tools/r2_profile.sh 9 16 > gpurun_out/f1_profile9.txt 2>&1
tools/r2_profile.sh 7 32 > gpurun_out/f1_profile7.txt 2>&1
cap() {  # name, env, regex, skip
  env $2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 2 -f -o gpurun_out/r2_$1 \
    python tools/quick_ab.py --spp 64 --warm 1 --reps 1 "$2" > gpurun_out/r2_$1.log 2>&1
  ncu -i gpurun_out/r2_$1.ncu-rep --page details > gpurun_out/r2_$1_details.txt 2>/dev/null
}
cap default "RTX_TRACE=1" wf_ 1850
cap trace2 "RTX_TRACE=2" wf_trace2 900
cap trace3 "RTX_TRACE=3" wf_trace3 900
cat gpurun_out/f1_pytest.log gpurun_out/f1_profile9.txt gpurun_out/f1_profile7.txt; ls -la gpurun_out/r2_*details.txt gpurun_out/r2_counters*
