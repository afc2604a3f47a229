#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "asynchronous or cli_ or reference_resolution or (psnr and 9) or bad_arguments" 2>&1 | tail -8 ) > gpurun_out/j14_pytest.log
for sc in 3 5 9; do
( timeout 300 python tools/quick_ab.py --scene $sc --spp 128 "RTX_PERLIN_SMEM=0" "RTX_PERLIN_SMEM=1" "RTX_PERLIN_SMEM=0" "RTX_PERLIN_SMEM=1" 2>&1 | sed "s/^/scene $sc: /" ) >> gpurun_out/j14_perlin.log
done
( time timeout 900 python bench.py 2> gpurun_out/j14_bench.err > gpurun_out/j14_bench.json ) 2> gpurun_out/j14_bench_time.txt
cat gpurun_out/j14_pytest.log gpurun_out/j14_perlin.log gpurun_out/j14_bench_time.txt; tail -3 gpurun_out/j14_bench.err
python -c "
import json; d=json.load(open('gpurun_out/j14_bench.json')); print(d['value'], d['e2e']['value'], d.get('big_scene')); print(d['cpu_baseline']['value'], [ (o['scene'], o['value']) for o in d['cpu_baseline']['other_scenes']])"
