#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== write probe"; ./tools/probes/write_probe 2>&1 | tee $OUT/write_probe.log
echo "== extra gpu tests"; timeout 1200 python -m pytest tests/test_gpu_extra.py -m gpu -x -q 2>&1 | tail -8
echo "== layer sweep (config 5), batch 256"
for L in 4 8 12 16; do python bench.py --layers $L --batch 256 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('L=$L n_signals %d value %.1f e2e %.1f expand %.1f GB/s frac %.4f slots %d export %.1f GB/s' % (d['config']['n_signals'], d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['config']['resident_slots'], d['config']['witness_export_d2h_gbs'] or -1))
    elif 'rror' in l: print(l.strip())
" | tee -a $OUT/layer_sweep.log; done
