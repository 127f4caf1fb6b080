#!/bin/bash
# round 2, GPU session O: small levels on CTA 0 alone, Poseidon conversion per segment, branch-free inversion steps on every thread
# (the small-level path and tools/solo_sweep.py were removed after this session: measured slower, profiles/r02o_solo_sweep.log)
TAG=${1:-r02o}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_$TAG.log
echo "== level profile"; timeout 600 python tools/eval_levels.py $OUT 2>&1 | tee $OUT/eval_levels_$TAG.log
echo "== small levels on one CTA"; timeout 900 python tools/solo_sweep.py 2>&1 | grep "^{" | tee $OUT/solo_sweep_$TAG.log
echo "== k_eval sweep (reduced witness)"; SWEEP_OPT=1 timeout 900 python tools/eval_sweep.py 2>&1 | head -2 | tee $OUT/eval_sweep_o1_$TAG.log
echo "== bench default"; timeout 900 python bench.py --no-cpu-baseline 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
