#!/usr/bin/env python3
"""Where do the last 4 % go?  Expand-only (k_eval skipped after the first two chunks, their stores re-expanded) vs the real pipeline, same
batch, same launches (TUNING build).  If the two agree, the gap between the kernels' stand-alone times and the pipelined pair time is not
the concurrent eval kernel."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proof-of-burn_b200"))
import pob_b200
from pob_b200 import synth
pob_b200.LIB_PATH = os.path.join(ROOT, "proof-of-burn_b200", "pob_b200", "libpob_b200_tuning.so")
shape = (16, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)
packed = synth.pack_instances(synth.make_batch(512, shape, seed=5), shape)
for skip, group in ((0, 0), (1, 0), (0, 0), (1, 0), (1, 8), (0, 8), (1, 21), (0, 21)):
    os.environ["POB_SKIP_EVAL"] = str(skip)
    if group:
        os.environ["POB_EXPAND_GROUP"] = str(group)
    else:
        os.environ.pop("POB_EXPAND_GROUP", None)
    c = pob_b200.Circuit(pob_b200.MAIN_PROOF_OF_BURN)
    c.stage(packed)
    c.run_packed(None, n=512, staged=True, discard=True)
    t = min((c.run_packed(None, n=512, staged=True, discard=True).timing for _ in range(3)), key=lambda q: q["total_ms"])
    print(json.dumps({"skip_eval": skip, "expand_group": c.desc["expand_group"], "wit_s": round(512 / (t["total_ms"] / 1e3), 1), "total_ms": round(t["total_ms"], 2),
                      "expand_ms_sum": round(t["expand_ms"], 2), "ms_per_16_witnesses": round(16 * t["expand_ms"] / 512, 3),
                      "expand_gbs": round(32.0 * c.n_signals * 512 / (t["expand_ms"] / 1e3) / 1e9, 1), "eval_ms_sum": round(t["eval_ms"], 2)}), flush=True)
    c.close()
