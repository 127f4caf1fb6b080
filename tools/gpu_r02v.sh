#!/bin/bash
# round 2, GPU session V (2 GPUs): smoke() and the 2-GPU line of the default bench on the final build
TAG=${1:-r02v}; OUT=gpurun_out; mkdir -p $OUT
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-160 | tee $OUT/smoke_$TAG.log
echo "== bench, 2 GPUs"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 \
    --no-cpu-baseline --consume-batch 0 --export-sample 0 --reduced-batch 0 --no-selfcheck 2>$OUT/bench_2gpu_$TAG.err | tee $OUT/bench_2gpu_$TAG.json | cut -c1-260
tail -2 $OUT/bench_2gpu_$TAG.err
