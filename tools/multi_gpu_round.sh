#!/bin/bash
# usage: tools/multi_gpu_round.sh N   — the N-GPU measurements of a round, one gpurun --gpus N call:
# Cornell-box bench line (scene 7, 600x600), the reference's default frames through the CLI (scene 9 at 10 000 spp,
# scene 7 at 200 spp) timed end to end, and both frames compared with the reference's shipped renders.
N=${1:-8}
mkdir -p gpurun_out
[ -n "$SKIP_BENCH" ] || python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --scene 7 --steps 4 --warmup 3 > gpurun_out/bench_s7_${N}gpu.json 2> gpurun_out/bench_s7_${N}gpu.err
[ -n "$SKIP_BENCH" ] || python -c "import json;d=json.load(open('gpurun_out/bench_s7_${N}gpu.json'));print('scene 7, $N GPUs:', round(d['value']/1e6,1), 'M samples/s, e2e', round(d['e2e']['value']/1e6,1))"
for job in "9 final_10k" "7 cornell_200"; do set -- $job
  t0=$(date +%s%N)
  RTTNW_VERBOSE=1 rttnw_b200/lib/rttnw $1 --gpus $N --out gpurun_out/$2_${N}gpu.png 2>&1 | tail -8
  echo "  wall $(( ($(date +%s%N) - t0) / 1000000 )) ms (process start to image.png written)"
  python tools/shipped_compare.py gpurun_out/$2_${N}gpu.png
done
