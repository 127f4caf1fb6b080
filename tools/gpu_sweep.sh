#!/bin/bash
# expand-kernel variant sweep (one GPU).  Usage: bash tools/gpu_sweep.sh "0 1 2 3 4"
for v in ${1:-0 1 2 3 4}; do
  echo "== variant $v"
  POB_EXPAND_VARIANT=$v python bench.py --no-cpu-baseline --batch 128 --steps 3 --warmup 2 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.1f e2e %.1f roofline %.1f GB/s frac %.4f eval %.2f ms/launch' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['config']['eval_kernel']['ms_per_launch']))
    elif 'rror' in l: print(l.strip())
"
done
