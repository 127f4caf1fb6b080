#!/usr/bin/env python3
"""k_eval shape sweep on one GPU (TUNING build: `make -C proof-of-burn_b200/csrc tuning`): threads per CTA x CTAs per instance
(thread-block cluster).  Prints, per shape: single-witness latency (host input -> witness in HBM), eval-only time of a 128-
instance batch, and generation throughput of a 256-instance batch.  Every shape is checked against the digests of the default
shape (and those against the oracle in the test-suite)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proof-of-burn_b200"))
import numpy as np
import pob_b200
from pob_b200 import synth

pob_b200.LIB_PATH = os.path.join(ROOT, "proof-of-burn_b200", "pob_b200", "libpob_b200_tuning.so")
shape = (16, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)
opt = int(os.environ.get("SWEEP_OPT", "0"))
packed = synth.pack_instances(synth.make_batch(256, shape, seed=5), shape)
ref = None
SHAPES = [(512, 0, 1), (512, 0, 0), (1024, 1, 1), (1024, 1, 0), (512, 4, 1), (512, 4, 0), (512, 2, 1), (256, 8, 1), (1024, 0, 1)]   # cluster 0 = chosen per launch
for threads, cluster, pf in SHAPES:
    os.environ["POB_EVAL_THREADS"], os.environ["POB_EVAL_CLUSTER"], os.environ["POB_EVAL_PREFETCH"] = str(threads), str(cluster), str(pf)
    try:
        c = pob_b200.Circuit(pob_b200.MAIN_PROOF_OF_BURN, opt=opt)
        one = packed[:1]
        c.run_packed(one)
        lat = sorted(c.run_packed(one).timing["total_ms"] for _ in range(7))[3]
        t1 = c.run_packed(one).timing
        c.run_packed(packed[:128], expand=False)
        ev = c.run_packed(packed[:128], expand=False).timing
        c.stage(packed)
        c.run_packed(None, n=256, staged=True, discard=True)
        th = c.run_packed(None, n=256, staged=True, discard=True).timing
        dg = c.run_packed(packed[:4], digest=True)
        ok = bool((dg.status == 0).all())
        if ref is None:
            ref = dg.digests.copy()
        same = bool(np.array_equal(dg.digests, ref))
        print(json.dumps({"threads": threads, "cluster": cluster, "prefetch": pf, "latency_ms": round(lat, 3), "lat_eval_ms": round(t1["eval_ms"], 3), "lat_expand_ms": round(t1["expand_ms"], 3),
                          "eval128_ms": round(ev["total_ms"], 3), "eval128_kernel_ms": round(ev["eval_ms"], 3), "batch256_wit_s": round(256 / (th["total_ms"] / 1e3), 1),
                          "batch256_eval_ms": round(th["eval_ms"], 2), "ok": ok, "digests_equal_default": same}), flush=True)
        c.close()
    except Exception as e:
        print(json.dumps({"threads": threads, "cluster": cluster, "error": str(e)}), flush=True)
