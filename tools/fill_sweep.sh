#!/bin/bash
# fill_sweep.sh "<fills>" [extra bench args]: headline throughput per CP_WARP_FILL (replicas per warp)
fills="$1"; shift
for f in $fills; do
  CP_WARP_FILL=$f python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --configs "" "$@" 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fill $f moves/s %.4g'%d['value'],'ms/step %.1f'%d['ms_per_step'])"
done
