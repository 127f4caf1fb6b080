#!/bin/bash
# round 2, GPU session F: reduced-witness shape sweep; ncu of the TMA-staged k_expand_codes (1 witness, and inside a 16-witness group)
TAG=${1:-r02f}; OUT=gpurun_out; mkdir -p $OUT
NOEX="--no-cpu-baseline --parity 0 --consume-batch 0 --export-sample 0 --reduced-batch 0 --no-selfcheck"
echo "== O1 sweep"; timeout 1500 python tools/o1_sweep.py 2>&1 | tee $OUT/o1_sweep_$TAG.log
echo "== ncu full k_expand_codes (1 witness per launch)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_expand_codes -s 1 -c 1 -o $OUT/prof_k_expand_codes_$TAG -f \
    python bench.py --batch 1 --steps 1 --warmup 1 $NOEX > $OUT/ncu_k_expand_codes_$TAG.log 2>&1; tail -2 $OUT/ncu_k_expand_codes_$TAG.log
ncu -i $OUT/prof_k_expand_codes_$TAG.ncu-rep --page raw --csv > $OUT/${TAG}_k_expand_codes_ncu_raw.csv 2>/dev/null
echo "== ncu full k_expand_codes inside a 16-witness group"
timeout 900 ncu --set full --clock-control none -k regex:k_expand_codes -s 2 -c 1 -o $OUT/prof_k_expand_codes16_$TAG -f \
    python bench.py --batch 64 --steps 1 --warmup 1 $NOEX > $OUT/ncu_k_expand_codes16_$TAG.log 2>&1; tail -2 $OUT/ncu_k_expand_codes16_$TAG.log
ncu -i $OUT/prof_k_expand_codes16_$TAG.ncu-rep --page raw --csv > $OUT/${TAG}_k_expand_codes16_ncu_raw.csv 2>/dev/null
