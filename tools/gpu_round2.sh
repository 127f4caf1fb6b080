#!/bin/bash
# profiles for the current build + multi-GPU scaling check.  Usage: bash tools/gpu_round2.sh <tag> [ngpu]
TAG=${1:-r01b}; NG=${2:-1}; OUT=gpurun_out; mkdir -p $OUT
if [ "$NG" -gt 1 ]; then
  echo "== torchrun x$NG"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $NG --steps 3 --warmup 3 2>$OUT/bench_${NG}gpu_$TAG.err | tee $OUT/bench_${NG}gpu_$TAG.json
  tail -5 $OUT/bench_${NG}gpu_$TAG.err
  exit 0
fi
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu_$TAG.log
echo "== bench b1024"; timeout 900 python bench.py --batch 1024 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | tee $OUT/bench_ref_$TAG.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_$TAG.log 2>&1
tail -4 $OUT/launches_$TAG.csv
echo "== ncu full k_expand (1 witness per launch)"
POB_EXPAND_GROUP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_expand_round -s 1 -c 1 -o $OUT/prof_expand_$TAG -f \
    python bench.py --batch 1 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_expand_$TAG.log 2>&1; tail -2 $OUT/ncu_expand_$TAG.log
echo "== ncu full k_eval (32 instances)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 1 -c 1 -o $OUT/prof_eval_$TAG -f \
    python bench.py --batch 32 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_eval_$TAG.log 2>&1; tail -2 $OUT/ncu_eval_$TAG.log
ls -la $OUT | tail -12
