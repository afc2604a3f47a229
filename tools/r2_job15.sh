#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "asynchronous or other_kernel_forms" 2>&1 | tail -6 ) > gpurun_out/j15_pytest.log
tools/r2_profile.sh 9 16 > gpurun_out/j15_profile9.txt 2>&1
tools/r2_profile.sh 7 32 > gpurun_out/j15_profile7.txt 2>&1
cap() {  # name, env, regex
  env $2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1850 -c 2 -f -o gpurun_out/r2_$1 \
    python tools/quick_ab.py --spp 64 --warm 1 --reps 1 "$2" > gpurun_out/r2_$1.log 2>&1
  ncu -i gpurun_out/r2_$1.ncu-rep --page details > gpurun_out/r2_$1_details.txt 2>/dev/null
}
cap default "RTX_TRACE=1" wf_
cap trace2 "RTX_TRACE=2" wf_trace2
cap trace3 "RTX_TRACE=3 RTX_T_BURST=8" wf_trace3
cat gpurun_out/j15_pytest.log gpurun_out/j15_profile9.txt gpurun_out/j15_profile7.txt; ls -la gpurun_out/r2_*
