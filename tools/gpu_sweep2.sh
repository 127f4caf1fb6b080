#!/bin/bash
run() { echo "== $*"; env "$@" python bench.py --no-cpu-baseline --batch 128 --steps 3 --warmup 2 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.1f roofline %.1f GB/s frac %.4f eval %.2f ms/launch expand-share %.3f' % (d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['config']['eval_kernel']['ms_per_launch'], d['roofline']['expand_share_of_step']))
    elif 'rror' in l: print(l.strip())
"; }
run POB_SERIALIZE=1
run POB_EVAL_THREADS=256
run POB_EVAL_THREADS=512
run POB_EVAL_THREADS=256 POB_EVAL_CHUNK=64
run POB_EVAL_THREADS=512 POB_EVAL_CHUNK=16
