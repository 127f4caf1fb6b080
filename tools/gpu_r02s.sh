#!/bin/bash
# round 2, GPU session S: 256-bit value loads/stores in k_eval, register cap 104, SELSUM batch removed
TAG=${1:-r02s}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_$TAG.log
echo "== level profile"; timeout 600 python tools/eval_levels.py $OUT 2>&1 | tee $OUT/eval_levels_$TAG.log
echo "== bench default"; timeout 900 python bench.py --no-cpu-baseline 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
