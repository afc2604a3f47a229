#!/bin/bash
mkdir -p gpurun_out
( python tools/quick_ab.py --counted --prof "RTX_TRACE=1 RTX_SHADE=1" "RTX_TRACE=1 RTX_SHADE=2" 2>&1;
  python tools/quick_ab.py --counted --scene 8 --spp 256 "RTX_TRACE=1 RTX_SHADE=1" "RTX_TRACE=1 RTX_SHADE=2" 2>&1 ) > gpurun_out/j9.log
cat gpurun_out/j9.log
