#!/bin/bash
# round 2, GPU session N: Kaliski almost-inverse for the deferred inverses, Keccak tables in constant memory, code-tile kernel next to the round kernel
TAG=${1:-r02n}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_$TAG.log
echo "== level profile"; timeout 600 python tools/eval_levels.py $OUT 2>&1 | tee $OUT/eval_levels_$TAG.log
echo "== code-tile kernel next to the round kernel"; timeout 900 python tools/codes_overlap_sweep.py 2>&1 | grep "^{" | tee $OUT/codes_overlap_$TAG.log
echo "== k_eval sweep (O0)"; timeout 900 python tools/eval_sweep.py 2>&1 | head -2 | tee $OUT/eval_sweep_$TAG.log
echo "== bench default"; timeout 900 python bench.py --no-cpu-baseline 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
