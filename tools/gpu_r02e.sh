#!/bin/bash
# round 2, GPU session E: does the write stream flush the eval working set out of L2?  streaming stores / persisting-L2 window sweep + launch list
TAG=${1:-r02e}; OUT=gpurun_out; mkdir -p $OUT
echo "== L2 sweep"; timeout 1200 python tools/coresidency_sweep.py 2>&1 | tee $OUT/l2_sweep_$TAG.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline --parity 0 --consume-batch 0 --export-sample 0 --reduced-batch 0 --no-selfcheck > $OUT/bench_under_ncu_$TAG.log 2>&1
grep -E "k_expand|k_eval" $OUT/launches_$TAG.csv | tail -6 | cut -d, -f5,9,15
