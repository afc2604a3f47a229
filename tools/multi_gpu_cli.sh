#!/bin/bash
# usage: tools/multi_gpu_cli.sh N  (inside one `gpurun --gpus N` call): wall clock of the reference's default frames through
# the CLI at 1, 2, 4, ... N GPUs (process start to exit, measured by the shell), then the bench line at N GPUs
N=${1:-8}
mkdir -p gpurun_out; cd gpurun_out && ln -sf ../assets assets
now() { python3 -c 'import time; print(time.time())'; }
run() {  # label, command...
  label=$1; shift
  t0=$(date +%s%N)
  ( RTTNW_VERBOSE=1 "$@" ) > cli_$label.txt 2>&1
  t1=$(date +%s%N)
  echo "process wall $(( (t1 - t0) / 1000000 )) ms" >> cli_$label.txt
}
for n in 1 2 4 8; do
  [ $n -gt $N ] && continue
  run s9_$n ../rttnw_b200/lib/rttnw 9 --gpus $n --out cli_final_$n.png
  run s7_$n ../rttnw_b200/lib/rttnw 7 --gpus $n --out cli_cornell_$n.png
done
RTTNW_SINGLE_PROCESS=1 run s9_${N}_single_process ../rttnw_b200/lib/rttnw 9 --gpus $N --out cli_final_sp.png
cd ..
for n in 1 2 4 8; do [ $n -le $N ] && { echo "== rttnw 9 --gpus $n"; cat gpurun_out/cli_s9_$n.txt; echo "== rttnw 7 --gpus $n"; cat gpurun_out/cli_s7_$n.txt; }; done
echo "== RTTNW_SINGLE_PROCESS=1 rttnw 9 --gpus $N"; cat gpurun_out/cli_s9_${N}_single_process.txt
python tools/shipped_compare.py gpurun_out/cli_final_$N.png; python tools/shipped_compare.py gpurun_out/cli_cornell_$N.png
python - <<PY
import numpy as np, sys
sys.path.insert(0, '.')
from rttnw_b200.render import png_read_rgba8
a = png_read_rgba8('gpurun_out/cli_final_1.png').astype(int); b = png_read_rgba8('gpurun_out/cli_final_$N.png').astype(int)
print('1-GPU frame vs $N-GPU frame: max |difference|', np.abs(a - b).max(), 'mean', np.abs(a - b).mean())
PY
rm -f gpurun_out/cli_final_*.png gpurun_out/cli_cornell_[124].png
[ -n "$SKIP_BENCH" ] && exit 0
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29573 --nproc-per-node $N bench.py --gpus $N --steps 4 > gpurun_out/cli_bench_$N.json 2> gpurun_out/cli_bench_$N.err
python -c "
import json; d=json.load(open('gpurun_out/cli_bench_$N.json')); print('bench', d['n_gpus'], d['value']/1e6, d['e2e']['value']/1e6, d['details']['combine'], {k:(round(v['seconds'],3)) for k,v in d['frames'].items()})"
