#!/bin/bash
# round 2, GPU session C: evidence for the final build -- bench, reference arm, ncu launch list, ncu --set full of both expand kernels
# (1 witness per launch) and of k_eval, config-5 sweep on one GPU with in-run parity
TAG=${1:-r02c}; OUT=gpurun_out; mkdir -p $OUT
NOEX="--no-cpu-baseline --parity 0 --consume-batch 0 --export-sample 0 --reduced-batch 0 --no-selfcheck"
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu_$TAG.log
echo "== k_eval sweep (O0)"; timeout 900 python tools/eval_sweep.py 2>&1 | tee $OUT/eval_sweep_$TAG.log
echo "== k_eval sweep (reduced witness)"; SWEEP_OPT=1 timeout 900 python tools/eval_sweep.py 2>&1 | tee $OUT/eval_sweep_o1_$TAG.log
echo "== bench b1024"; timeout 900 python bench.py --batch 1024 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 1 | tee $OUT/bench_ref_$TAG.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --batch 64 --steps 1 --warmup 1 $NOEX > $OUT/bench_under_ncu_$TAG.log 2>&1
tail -4 $OUT/launches_$TAG.csv
for K in k_expand_round k_expand_codes; do
  echo "== ncu full $K (1 witness per launch)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o $OUT/prof_${K}_$TAG -f \
      python bench.py --batch 1 --steps 1 --warmup 1 $NOEX > $OUT/ncu_${K}_$TAG.log 2>&1; tail -2 $OUT/ncu_${K}_$TAG.log
  ncu -i $OUT/prof_${K}_$TAG.ncu-rep --page raw --csv > $OUT/${TAG}_${K}_ncu_raw.csv 2>/dev/null
done
echo "== ncu full k_expand_codes inside a 16-witness group"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_expand_codes -s 2 -c 1 -o $OUT/prof_k_expand_codes16_$TAG -f \
    python bench.py --batch 64 --steps 1 --warmup 1 $NOEX > $OUT/ncu_k_expand_codes16_$TAG.log 2>&1; tail -2 $OUT/ncu_k_expand_codes16_$TAG.log
ncu -i $OUT/prof_k_expand_codes16_$TAG.ncu-rep --page raw --csv > $OUT/${TAG}_k_expand_codes16_ncu_raw.csv 2>/dev/null
echo "== ncu full k_eval (one chunk)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 1 -c 1 -o $OUT/prof_eval_$TAG -f \
    python bench.py --batch 32 --steps 1 --warmup 1 $NOEX > $OUT/ncu_eval_$TAG.log 2>&1; tail -2 $OUT/ncu_eval_$TAG.log
ncu -i $OUT/prof_eval_$TAG.ncu-rep --page raw --csv > $OUT/${TAG}_eval_ncu_raw.csv 2>/dev/null
echo "== config 5 sweep, 1 GPU"
bash tools/gpu_config5.sh 1 full $TAG
ls -la $OUT | tail -8
