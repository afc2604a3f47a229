#!/bin/bash
# usage: tools_tune.sh "NODE LEAF SHADE BURST" ...   (scratch helper for GPU tuning runs)
for w in "$@"; do set -- $w
RTX_W_NODE=$1 RTX_W_LEAF=$2 RTX_W_SHADE=$3 RTX_NODE_BURST=$4 python bench.py --steps 4 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('W $w', round(d['value']/1e6,1), 'Msamples/s', round(d['rays_per_sec']/1e9,3), 'Grays/s e2e', round(d['e2e']['value']/1e6,1))"
done
