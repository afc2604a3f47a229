#!/bin/bash
# L2 prefetch of deferred children in the fixed-ray kernel: big scene (LBVH), on/off; small scene on/off; parity of the fixed rays
mkdir -p gpurun_out
for p in 0 1 0 1; do RTX_PREFETCH=$p timeout 600 python tools/big_scene.py 2000000 2>&1 | grep -E "^lbvh|^ploc" | sed "s/^/prefetch=$p /" | tee -a gpurun_out/j24_prefetch.log; done
for p in 0 1; do RTX_PREFETCH=$p timeout 300 python tools/time_trace.py 9 2000000 2>&1 | sed -n 1,4p | sed "s/^/prefetch=$p /" | tee -a gpurun_out/j24_prefetch.log; done
RTX_PREFETCH=1 timeout 900 python -m pytest tests -m gpu -x -q -k "fixed_rays and not 10M" > gpurun_out/j24_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j24_tests.log
tail -4 gpurun_out/j24_tests.log
