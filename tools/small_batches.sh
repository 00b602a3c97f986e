#!/bin/bash
# small_batches.sh: headline workload at small replica counts, automatic replicas-per-warp against 32
for n in 2048 8192 32768; do
  for f in 0 32; do
    CP_WARP_FILL=$f python bench.py --replicas $n --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --configs "" 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('replicas $n fill $f moves/s %.4g'%d['value'],'ms/step %.1f'%d['ms_per_step'])"
  done
done
