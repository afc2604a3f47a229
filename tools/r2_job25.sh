#!/bin/bash
# throughput against samples per call (scene 9): where do the 654 M of the 10 000-spp frame come from?
mkdir -p gpurun_out
for spp in 32 128 512 2048; do timeout 300 python tools/quick_ab.py --scene 9 --spp $spp --reps 3 --prof "RTX_X=$spp" 2>&1 | grep RTX_X | tee -a gpurun_out/j25_spp.log; done
timeout 300 python tools/quick_ab.py --scene 9 --spp 10000 --reps 1 --warm 64 --prof "RTX_X=10000" 2>&1 | grep RTX_X | tee -a gpurun_out/j25_spp.log
