#!/bin/bash
# One GPU session: parity tests, bench line, ncu launch list and one full capture of the expand kernel.
# Usage (under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_$TAG.log
echo "== bench"; timeout 600 python bench.py 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --batch 32 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_$TAG.log 2>&1
tail -5 $OUT/launches_$TAG.csv
echo "== ncu full (k_expand, 1 instance per launch)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_expand -s 1 -c 1 -o $OUT/prof_expand_$TAG -f \
    python bench.py --batch 1 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
tail -3 $OUT/ncu_full_$TAG.log
ls -la $OUT
