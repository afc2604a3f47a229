#!/bin/bash
# PLOC builder: parity, then trees compared (host SAH / LBVH / PLOC) on random spheres and on the shipped scenes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "device_built_bvh" > gpurun_out/j23_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j23_tests.log
tail -15 gpurun_out/j23_tests.log
RTX_DEBUG_BUILD=1 timeout 600 python tools/big_scene.py 200000 2>&1 | tee gpurun_out/j23_big200k.log | tail -12
RTX_DEBUG_BUILD=1 timeout 600 python tools/big_scene.py 2000000 2>&1 | tee gpurun_out/j23_big2m.log | tail -8
for r in 8 32; do RTX_PLOC_RADIUS=$r RTX_DEBUG_BUILD=1 timeout 600 python tools/big_scene.py 2000000 2>&1 | grep -E "ploc|PLOC" | tail -3 | tee -a gpurun_out/j23_radius.log; done
for b in sah lbvh ploc; do RTX_BVH=$b timeout 300 python tools/quick_ab.py --scene 9 --spp 64 --reps 3 "RTX_X=$b" 2>&1 | grep RTX_X | tee -a gpurun_out/j23_s9.log; done
