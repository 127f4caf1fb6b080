#!/bin/bash
# round 2, GPU session T: ncu evidence for the final k_eval (launch list of a short bench run, --set full of one 32-instance chunk), level profile
TAG=${1:-r02t}; OUT=gpurun_out; mkdir -p $OUT
NOEX="--no-cpu-baseline --parity 0 --consume-batch 0 --export-sample 0 --reduced-batch 0 --no-selfcheck"
echo "== level profile"; timeout 600 python tools/eval_levels.py $OUT 2>&1 | tee $OUT/eval_levels_$TAG.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --batch 64 --steps 1 --warmup 1 $NOEX > $OUT/bench_under_ncu_$TAG.log 2>&1
tail -4 $OUT/launches_$TAG.csv
echo "== ncu full k_eval (one chunk of 32 instances, 4-CTA clusters)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 1 -c 1 -o $OUT/prof_eval_$TAG -f \
    python bench.py --batch 32 --steps 1 --warmup 1 $NOEX > $OUT/ncu_eval_$TAG.log 2>&1; tail -2 $OUT/ncu_eval_$TAG.log
ncu -i $OUT/prof_eval_$TAG.ncu-rep --page raw --csv > $OUT/${TAG}_eval_ncu_raw.csv 2>/dev/null
echo "== bench (short)"; timeout 600 python bench.py --no-cpu-baseline --consume-batch 0 --export-sample 0 --reduced-batch 0 --no-selfcheck 2>/dev/null | tee $OUT/bench_$TAG.json | cut -c1-200
ls -la $OUT | tail -6
