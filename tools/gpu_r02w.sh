#!/bin/bash
# round 2, GPU session W: compute-sanitizer memcheck over smoke() (Spend(31): k_eval cluster launch, TMA staging, 256-bit value accesses, both expand kernels)
TAG=${1:-r02w}; OUT=gpurun_out; mkdir -p $OUT
timeout 170 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py smoke > $OUT/memcheck_smoke_$TAG.log 2>&1; echo "rc=$?" >> $OUT/memcheck_smoke_$TAG.log
grep -E "ERROR SUMMARY|smoke ok|rc=|Invalid|misaligned" $OUT/memcheck_smoke_$TAG.log | head -8
