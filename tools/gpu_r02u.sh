#!/bin/bash
# round 2, GPU session U: confirmation of the final commit of round 2 -- full GPU suite, smoke(), default bench (1 GPU)
TAG=${1:-r02u}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $OUT/smoke_$TAG.log
echo "== bench default"; timeout 900 python bench.py 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference 2>>$OUT/bench_$TAG.err | tee $OUT/bench_ref_$TAG.json
