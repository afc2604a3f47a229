#!/bin/bash
# round 2, final job 2 (one GPU): compute-sanitizer over small renders under every kernel form, then the default bench line
mkdir -p gpurun_out
{
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool: scene 9, 96x96, 1 + 1 spp, every kernel form"
  timeout 1200 compute-sanitizer --tool $tool --print-limit 3 python tools/quick_ab.py --scene 9 --spp 1 --warm 1 --reps 1 --width 96 --height 96 \
    "RTX_TRACE=1" "RTX_TRACE=2" "RTX_TRACE=3" "RTX_TRACE=4" "RTX_SHADE=1 RTX_PERLIN_SMEM=0" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Race reported|Invalid|hazard|RTX_|Error" | head -40
done
echo "== compute-sanitizer --tool memcheck: the combine kernels and the fixed-ray known answers"
timeout 900 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests -m gpu -x -q -k "slice_single or reduce_tonemap_kernel or known_answers or empty_and_degenerate or nccl_reduce" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid" | head
} > gpurun_out/r2_sanitizer.txt 2>&1
( time python bench.py 2> gpurun_out/f2_bench.err > gpurun_out/f2_bench.json ) 2> gpurun_out/f2_time.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f2_bench_ref.json 2>/dev/null
cat gpurun_out/r2_sanitizer.txt gpurun_out/f2_time.txt
python -c "
import json; d=json.load(open('gpurun_out/f2_bench.json')); r=d['roofline']; print(d['value'], d['e2e']['value'], r['bound'], r['frac'], r['achieved'], r['peak'], r['traffic'], r['hbm']['dram']['frac'], r['l2']['frac'], d['big_scene']['roofline']['frac'])
e=json.load(open('gpurun_out/f2_bench_ref.json')); print(e['value'], e['config']==d['config'])"
