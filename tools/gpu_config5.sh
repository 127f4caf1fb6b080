#!/bin/bash
# BASELINE.json configs[4]: circuit-shape (maxNumLayers) x batch sweep on one GPU; CPU baseline skipped (see bench.py default run)
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/config5_sweep.log
for spec in "4 1024" "8 1024" "12 1024" "16 256" "16 1024" "16 4096"; do set -- $spec
  python bench.py --layers $1 --batch $2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('L=$1 batch=$2 n_signals %d slots %d | value %.1f wit/s  e2e %.1f  | expand %.0f GB/s (%.3f of measured peak) | %.2f TB/s of witness bytes' % (d['config']['n_signals'], d['config']['resident_slots'], d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['value'] * d['config']['witness_bytes'] / 1e12))
    elif 'rror' in l: print(l.strip())
" | tee -a $OUT/config5_sweep.log
done
