#!/bin/bash
# round 2, GPU session A: parity suite on the new boundary, bench with hand-off figures, reference arm, launch list,
# ncu --set full of BOTH shipped expand kernels (1 witness per launch) and of k_eval, GDS probe.
TAG=${1:-r02a}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_$TAG.log
echo "== bench b1024"; timeout 900 python bench.py --batch 1024 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 1 | tee $OUT/bench_ref_$TAG.json
echo "== gds probe"; timeout 300 tools/probes/gds_probe $OUT 2 2>&1 | tee $OUT/gds_probe_$TAG.log; timeout 120 tools/probes/gds_probe /dev/shm 2 2>&1 | tee -a $OUT/gds_probe_$TAG.log; df -h . /tmp /dev/shm | tee -a $OUT/gds_probe_$TAG.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline --parity 0 --consume-batch 0 --export-sample 0 --reduced-batch 0 --no-selfcheck > $OUT/bench_under_ncu_$TAG.log 2>&1
tail -4 $OUT/launches_$TAG.csv
for K in k_expand_round k_expand_codes; do
  echo "== ncu full $K (1 witness per launch)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o $OUT/prof_${K}_$TAG -f \
      python bench.py --batch 1 --steps 1 --warmup 1 --no-cpu-baseline --parity 0 --consume-batch 0 --export-sample 0 --reduced-batch 0 --no-selfcheck > $OUT/ncu_${K}_$TAG.log 2>&1; tail -2 $OUT/ncu_${K}_$TAG.log
  ncu -i $OUT/prof_${K}_$TAG.ncu-rep --page raw --csv > $OUT/${TAG}_${K}_ncu_raw.csv 2>/dev/null
done
echo "== ncu full k_eval (32 instances)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 1 -c 1 -o $OUT/prof_eval_$TAG -f \
    python bench.py --batch 32 --steps 1 --warmup 1 --no-cpu-baseline --parity 0 --consume-batch 0 --export-sample 0 --reduced-batch 0 --no-selfcheck > $OUT/ncu_eval_$TAG.log 2>&1; tail -2 $OUT/ncu_eval_$TAG.log
ncu -i $OUT/prof_eval_$TAG.ncu-rep --page raw --csv > $OUT/${TAG}_eval_ncu_raw.csv 2>/dev/null
ls -la $OUT | tail -15
