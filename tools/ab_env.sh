#!/bin/bash
# usage: tools/ab_env.sh [bench args --] "VAR=a VAR2=b" "VAR=c" ...   one bench.py run per environment, one line each
# (A/B inside ONE gpurun call: box-to-box variation is +-2 %, within a box +-0.1 %)
extra=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do case "$1" in *=*) break;; *) extra+=("$1"); shift;; esac; done
[ "$1" = "--" ] && shift
for cfg in "$@"; do
  env $cfg python bench.py --steps 4 --no-cpu-baseline "${extra[@]}" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$cfg |', round(d['value'] / 1e6, 1), 'Msamples/s', round(d['rays_per_sec'] / 1e9, 3), 'Grays/s  e2e', round(d['e2e']['value'] / 1e6, 1),
      ' trace us', round(1e3 * r['kernel_ms_per_launch'], 1), 'share', round(r['kernel_share_of_step'], 3), 'shade share', round(r['shade_kernel_share_of_step'], 3))"
done
