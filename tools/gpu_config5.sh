#!/bin/bash
# BASELINE.json configs[4]: circuit-shape (maxNumLayers L) x batch sweep with the in-run parity check kept (whole-witness digests
# of 2 instances per run against the oracle).  Usage: bash tools/gpu_config5.sh <ngpu> [full|short] [tag]
#   1 GPU, full : L in {4, 8, 12, 16} x batch {256, 1024, 4096}, plus one CPU-reference number per L
#   N GPUs      : L in {4, 8, 12, 16} x batch 1024 per GPU (torchrun, one rank per GPU)
NG=${1:-1}; MODE=${2:-full}; TAG=${3:-r02}; OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/config5_${NG}gpu_$TAG.log; : > $LOG
COMMON="--steps 2 --warmup 3 --no-cpu-baseline --parity 2 --consume-batch 0 --export-sample 0 --reduced-batch 0 --no-selfcheck"
fmt() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); p = d.get('parity') or {}
        print('gpus=%d L=$1 batch/gpu=$2 n_signals %d | value %.1f wit/s  e2e %.1f | expand %.0f GB/s (%.3f of measured peak) | %.2f TB/s of witness bytes | parity %s on %s instances | clocks %s %s' % (d['n_gpus'], d['config']['n_signals'], d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['value'] * d['config']['witness_bytes'] / 1e12, p.get('digest_match'), p.get('instances'), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons')))
    elif 'rror' in l: print(l.strip())
"; }
if [ "$NG" -eq 1 ]; then
  BATCHES="256 1024 4096"; [ "$MODE" = short ] && BATCHES="1024"
  for L in 4 8 12 16; do
    for B in $BATCHES; do
      timeout 900 python bench.py --layers $L --batch $B $COMMON 2>>$OUT/config5_err_$TAG.log | fmt $L $B | tee -a $LOG
    done
    timeout 600 python bench.py --impl reference --layers $L --steps 1 --warmup 0 2>>$OUT/config5_err_$TAG.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('cpu reference (oracle port) L=$L: %.2f wit/s on %d host processes' % (d['value'], d['cpu_baseline']['cores']))
" | tee -a $LOG
  done
else
  for L in 4 8 12 16; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus $NG --layers $L --batch 1024 $COMMON 2>>$OUT/config5_err_$TAG.log | fmt $L 1024 | tee -a $LOG
  done
fi
