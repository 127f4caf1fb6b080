#!/bin/bash
# round 2, GPU session P: inline-PTX Montgomery product (IMAD.WIDE pairs), inversion workers back on the last 8 warps, small-level path removed
TAG=${1:-r02p}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_$TAG.log
echo "== Montgomery product A/B"; timeout 900 python tools/mont_ab.py 2>&1 | grep "^{" | tee $OUT/mont_ab_$TAG.log
echo "== level profile"; timeout 600 python tools/eval_levels.py $OUT 2>&1 | tee $OUT/eval_levels_$TAG.log
echo "== k_eval sweep (reduced witness)"; SWEEP_OPT=1 timeout 900 python tools/eval_sweep.py 2>&1 | head -2 | tee $OUT/eval_sweep_o1_$TAG.log
echo "== bench default"; timeout 900 python bench.py --no-cpu-baseline 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
