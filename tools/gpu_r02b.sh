#!/bin/bash
# round 2, GPU session B: full GPU suite incl. pob_selfcheck + reduced witness, default bench (with the reduced figure)
TAG=${1:-r02b}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 | tee $OUT/pytest_gpu_$TAG.log
echo "== bench b1024"; timeout 900 python bench.py --batch 1024 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
