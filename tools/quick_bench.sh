#!/bin/bash
# quick_bench.sh "<lanes list>" <mc-steps> [extra bench args]: value + ms/step per lane setting
lanes="$1"; steps="$2"; shift 2
for l in $lanes; do
  python bench.py --steps 2 --warmup 1 --mc-steps $steps --lanes $l --no-cpu-baseline --no-e2e "$@" | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lanes',d['lanes_per_replica'],'moves/s %.4g'%d['value'],'ms/step %.1f'%d['ms_per_step'])"
done
