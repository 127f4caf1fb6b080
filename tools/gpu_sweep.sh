#!/bin/bash
# expand-kernel variant sweep + write-bandwidth ceiling probe (one GPU)
python - <<'PY'
import torch, time
x = torch.empty(8 << 30, dtype=torch.uint8, device="cuda")
for _ in range(2): x.zero_(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): x.zero_()
e1.record(); torch.cuda.synchronize()
print("write-only ceiling (torch zero_, 8 GiB x5): %.1f GB/s" % (5 * x.numel() / (e0.elapsed_time(e1) / 1e3) / 1e9))
PY
for v in 0 1 2 3; do
  echo "== variant $v"
  POB_EXPAND_VARIANT=$v python bench.py --no-cpu-baseline --batch 128 --steps 3 --warmup 2 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.1f e2e %.1f roofline %.1f GB/s frac %.4f eval %.2f ms/launch' % (d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['config']['eval_kernel']['ms_per_launch']))
    elif 'rror' in l: print(l.strip())
"
done
echo "== group 8 / 21 (variant 0)"
for g in 8 21; do POB_EXPAND_GROUP=$g python bench.py --no-cpu-baseline --batch 126 --steps 3 --warmup 2 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.1f roofline %.1f GB/s' % (d['value'], d['roofline']['achieved']))
"; done
