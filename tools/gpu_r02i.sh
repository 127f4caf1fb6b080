#!/bin/bash
# round 2, GPU session I: where do the last 4 % go?  expand-only vs the real pipeline (TUNING build)
TAG=${1:-r02i}; OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python tools/interference_probe.py 2>&1 | tee $OUT/interference_$TAG.log
