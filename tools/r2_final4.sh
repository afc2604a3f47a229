#!/bin/bash
# round 2, final job 4 (one GPU): counters of the final sources (96-byte pool slot) and the default bench line
mkdir -p gpurun_out
bash tools/r2_profile.sh 9 16 > gpurun_out/f5_profile9.txt 2>&1
cp gpurun_out/r2_counters_s9.json profiles/
( time python bench.py 2> gpurun_out/f5_bench.err > gpurun_out/f5_bench.json ) 2> gpurun_out/f5_time.txt
bash tools/r2_profile.sh 7 32 > gpurun_out/f5_profile7.txt 2>&1
cat gpurun_out/f5_time.txt
python -c "
import json; d=json.load(open('gpurun_out/f5_bench.json')); r=d['roofline']; print(d['value'], d['e2e']['value'], r['bound'], r['frac'], r['achieved'], r['peak'], r['traffic'], r['hbm']['dram']['frac'], r['l2']['frac'], d['big_scene']['roofline']['frac'])
print({k: (v['value'], v['seconds']) for k, v in d['frames'].items()})"
