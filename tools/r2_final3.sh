#!/bin/bash
# round 2, final job 3 (one GPU): counters of the final sources (scenes 9, 7), ncu --set full of one mid-frame (shade, trace) pair,
# compute-sanitizer over the forms added since final job 2, the default bench line and the reference arm
mkdir -p gpurun_out
bash tools/r2_profile.sh 9 16 > gpurun_out/f3_profile9.txt 2>&1
bash tools/r2_profile.sh 7 32 > gpurun_out/f3_profile7.txt 2>&1
cp gpurun_out/r2_counters_s9.json gpurun_out/r2_counters_s7.json profiles/   # (box-local: the bench line below reads them; the same files come back in gpurun_out/)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wf_ -s 900 -c 2 -f -o gpurun_out/r2_default3 python tools/quick_ab.py --scene 9 --spp 16 --warm 1 --reps 1 "" > gpurun_out/f3_ncu.log 2>&1
for k in wf_shade2_kernel wf_trace_kernel; do ncu -i gpurun_out/r2_default3.ncu-rep --page details -k regex:$k > gpurun_out/r3_${k}_details.txt 2>/dev/null; done
{
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool: scene 9, 96x96, 1 + 1 spp: default pair, ordered trace pass, PLOC-built tree"
  timeout 1200 compute-sanitizer --tool $tool --print-limit 3 python tools/quick_ab.py --scene 9 --spp 1 --warm 1 --reps 1 --width 96 --height 96 \
    "RTX_TRACE=1" "RTX_ORDER=1" "RTX_BVH=ploc" "RTX_BVH=lbvh" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Race reported|Invalid|hazard|RTX_|Error" | head -40
done
} > gpurun_out/r3_sanitizer.txt 2>&1
( time python bench.py 2> gpurun_out/f3_bench.err > gpurun_out/f3_bench.json ) 2> gpurun_out/f3_time.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f3_bench_ref.json 2>/dev/null
cat gpurun_out/r3_sanitizer.txt gpurun_out/f3_time.txt
python -c "
import json; d=json.load(open('gpurun_out/f3_bench.json')); r=d['roofline']; print(d['value'], d['e2e']['value'], r['bound'], r['frac'], r['achieved'], r['peak'], r['traffic'], r['hbm']['dram']['frac'], r['l2']['frac'], d['big_scene']['roofline']['frac'])
print({k: (v['value'], v['seconds']) for k, v in d['frames'].items()})
e=json.load(open('gpurun_out/f3_bench_ref.json')); print(e['value'], e['config']==d['config'])"
