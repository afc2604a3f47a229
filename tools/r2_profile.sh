#!/bin/bash
# usage (inside ONE gpurun call): tools/r2_profile.sh [scene] [spp]   -> gpurun_out/r2_counters_s<scene>.json, r2_launches_s<scene>.csv
# Hardware counters of every wavefront launch of one small render (metrics only: two ncu passes per launch).
scene=${1:-9}; spp=${2:-16}
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max
timeout 900 ncu --metrics $M --clock-control none -k regex:wf_ --csv --log-file gpurun_out/r2_launches_s$scene.csv \
  python tools/quick_ab.py --scene $scene --spp $spp --warm 1 --reps 1 "" > gpurun_out/r2_profile_s$scene.log 2>&1
python tools/r2_counters.py gpurun_out/r2_launches_s$scene.csv $((spp + 1)) gpurun_out/r2_counters_s$scene.json $scene > /dev/null
python tools/ncu_sum.py gpurun_out/r2_launches_s$scene.csv
