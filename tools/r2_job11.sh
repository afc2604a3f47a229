#!/bin/bash
mkdir -p gpurun_out
( RTX_TRACE=4 timeout 900 python -m pytest tests -m gpu -x -q -k "not ten_million" 2>&1 | tail -5 ) > gpurun_out/j11_pytest.log
( timeout 600 python tools/quick_ab.py --prof \
  "RTX_TRACE=1 RTX_SHADE=1" "RTX_TRACE=1 RTX_SHADE=2" "RTX_TRACE=4 RTX_SHADE=2" "RTX_TRACE=4 RTX_SHADE=1" \
  "RTX_TRACE=1 RTX_SHADE=2 RTX_WF_STREAMS=1" "RTX_TRACE=4 RTX_SHADE=2 RTX_WF_STREAMS=1" \
  "RTX_TRACE=4 RTX_SHADE=2 RTX_WF_STREAMS=3 RTX_WF_SLOTS=786432" \
  "RTX_TRACE=4 RTX_SHADE=2 RTX_WF_SLOTS=786432" \
  2>&1 ) > gpurun_out/j11_ab.log
for sc in 1 7 8; do
( timeout 300 python tools/quick_ab.py --scene $sc --spp 256 "RTX_TRACE=1 RTX_SHADE=2" "RTX_TRACE=4 RTX_SHADE=2" 2>&1 | sed "s/^/scene $sc: /" ) >> gpurun_out/j11_ab.log
done
cat gpurun_out/j11_pytest.log gpurun_out/j11_ab.log
