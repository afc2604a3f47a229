#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "other_kernel_forms or slice_single or nccl_reduce or cli_" 2>&1 | tail -8 ) > gpurun_out/j12_pytest.log
( timeout 900 python bench.py --steps 4 --cpu-seconds 4 --final-spp 1000 --cornell-spp 800 2> gpurun_out/j12_bench.err ) > gpurun_out/j12_bench.json
tools/r2_profile.sh 9 16 > gpurun_out/j12_profile.txt 2>&1
cat gpurun_out/j12_pytest.log; tail -5 gpurun_out/j12_bench.err; head -c 3000 gpurun_out/j12_bench.json; cat gpurun_out/j12_profile.txt
