#!/bin/bash
# round 2, GPU session B: full GPU suite incl. pob_selfcheck + reduced witness + TMA/cluster kernels, k_eval shape sweep, bench
TAG=${1:-r02b}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 | tee $OUT/pytest_gpu_$TAG.log
echo "== k_eval sweep (O0)"; timeout 900 python tools/eval_sweep.py 2>&1 | tee $OUT/eval_sweep_$TAG.log
echo "== k_eval sweep (reduced witness)"; SWEEP_OPT=1 timeout 900 python tools/eval_sweep.py 2>&1 | tee $OUT/eval_sweep_o1_$TAG.log
echo "== bench b1024"; timeout 900 python bench.py --batch 1024 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
