#!/bin/bash
TAG=${1:-r02k}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python tools/eval_levels.py $OUT 2>&1 | tee $OUT/eval_levels_$TAG.log
