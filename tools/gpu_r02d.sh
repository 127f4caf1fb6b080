#!/bin/bash
# round 2, GPU session D: TMA-staged k_expand_codes + co-residency of k_eval with the expand CTAs
TAG=${1:-r02d}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu_$TAG.log
echo "== co-residency sweep"; timeout 1200 python tools/coresidency_sweep.py 2>&1 | tee $OUT/coresidency_$TAG.log
echo "== bench b1024"; timeout 900 python bench.py --batch 1024 --no-cpu-baseline 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
