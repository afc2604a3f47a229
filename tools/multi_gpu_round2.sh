#!/bin/bash
# usage: tools/multi_gpu_round2.sh N   (inside one `gpurun --gpus N` call): the N-GPU measurements of round 2
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29571"
( timeout 300 python -m pytest tests -m gpu -x -q -k "cli_one_process" 2>&1 | tail -3 ) > gpurun_out/mg${N}_pytest.log
# 1. the combine alone, three ways
for n in 2 4 8; do
  [ $n -le $N ] && timeout 300 $TR --nproc-per-node $n tools/combine_ab.py > gpurun_out/mg${N}_combine_$n.json 2> gpurun_out/mg${N}_combine_$n.err
done
# 2. the bench line at 1, 2, 4, ... N GPUs on this box (strong-scaling frames inside)
for n in ${BENCH_NS:-$N}; do
  [ $n -gt $N ] && continue
  if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 --steps 4 --no-cpu-baseline > gpurun_out/mg${N}_bench_1.json 2> gpurun_out/mg${N}_bench_1.err
  else timeout 600 $TR --nproc-per-node $n bench.py --gpus $n --steps 4 > gpurun_out/mg${N}_bench_$n.json 2> gpurun_out/mg${N}_bench_$n.err; fi
done
# 3. the reference's default frames through the CLI, wall clock of the whole process
cd gpurun_out && ln -sf ../assets assets
for n in 1 2 4 8; do
  [ $n -gt $N ] && continue
  ( t0=$(date +%s.%N); RTTNW_VERBOSE=1 ../rttnw_b200/lib/rttnw 9 --gpus $n --out mg_final_$n.png; echo "process wall $(echo "$(date +%s.%N) - $t0" | bc) s" ) > mg${N}_cli9_$n.txt 2>&1
  ( t0=$(date +%s.%N); RTTNW_VERBOSE=1 ../rttnw_b200/lib/rttnw 7 --gpus $n --out mg_cornell_$n.png; echo "process wall $(echo "$(date +%s.%N) - $t0" | bc) s" ) > mg${N}_cli7_$n.txt 2>&1
done
( t0=$(date +%s.%N); RTTNW_VERBOSE=1 RTTNW_SINGLE_PROCESS=1 ../rttnw_b200/lib/rttnw 9 --gpus $N --out mg_final_sp.png; echo "process wall $(echo "$(date +%s.%N) - $t0" | bc) s" ) > mg${N}_cli9_single_process.txt 2>&1
cd ..
python tools/shipped_compare.py gpurun_out/mg_final_$N.png > gpurun_out/mg${N}_compare.txt 2>&1
python tools/shipped_compare.py gpurun_out/mg_cornell_$N.png >> gpurun_out/mg${N}_compare.txt 2>&1
rm -f gpurun_out/mg_final_[124].png gpurun_out/mg_cornell_[124].png gpurun_out/mg_final_sp.png
tail -3 gpurun_out/mg${N}_pytest.log; cat gpurun_out/mg${N}_combine_*.json; for f in gpurun_out/mg${N}_cli*.txt; do echo "== $f"; tail -12 $f; done; cat gpurun_out/mg${N}_compare.txt
for n in 1 2 4 8; do [ -f gpurun_out/mg${N}_bench_$n.json ] && python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/mg${N}_bench_$n.json'))
    print($n, 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), {k:(round(v['seconds'],3), round(v['value']/1e6,1)) for k,v in d['frames'].items()})
except Exception as e: print($n, 'bench failed', e)
"; done
