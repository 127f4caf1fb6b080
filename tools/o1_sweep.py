#!/usr/bin/env python3
"""Reduced-witness (POB_CREATE_O1) throughput vs k_eval shape and eval chunk (TUNING build).  With 10x fewer bytes to write the eval
kernel is a co-bottleneck; it is latency-bound, so several small CTAs (different instances) per SM may beat one large one."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proof-of-burn_b200"))
import numpy as np
import pob_b200
from pob_b200 import synth

pob_b200.LIB_PATH = os.path.join(ROOT, "proof-of-burn_b200", "pob_b200", "libpob_b200_tuning.so")
shape = (16, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)
N = 1184
packed = synth.pack_instances(synth.make_batch(N, shape, seed=5), shape)
ref = None
# (eval threads, cluster (0 = per launch), eval chunk)
for threads, cluster, chunk in [(512, 0, 128), (512, 1, 148), (256, 1, 148), (256, 1, 296), (256, 1, 444), (512, 1, 296), (256, 2, 148), (256, 1, 592)]:
    os.environ.update({"POB_EVAL_THREADS": str(threads), "POB_EVAL_CLUSTER": str(cluster), "POB_EVAL_CHUNK": str(chunk)})
    try:
        c = pob_b200.Circuit(pob_b200.MAIN_PROOF_OF_BURN, opt=1)
        c.stage(packed)
        c.run_packed(None, n=N, staged=True, discard=True)
        t = min((c.run_packed(None, n=N, staged=True, discard=True).timing for _ in range(2)), key=lambda q: q["total_ms"])
        ev = c.run_packed(None, n=N, staged=True, expand=False).timing
        dg = c.run_packed(packed[:3], digest=True)
        if ref is None:
            ref = dg.digests.copy()
        print(json.dumps({"eval_threads": threads, "eval_cluster": cluster, "eval_chunk": chunk, "slots": c.desc["n_slots"], "expand_group": c.desc["expand_group"],
                          "wit_s": round(N / (t["total_ms"] / 1e3), 1), "expand_gbs": round(32.0 * c.n_signals * N / (t["expand_ms"] / 1e3) / 1e9, 1),
                          "eval_only_wit_s": round(N / (ev["total_ms"] / 1e3), 1), "digests_ok": bool(np.array_equal(dg.digests, ref))}), flush=True)
        c.close()
    except Exception as e:
        print(json.dumps({"eval_threads": threads, "eval_cluster": cluster, "eval_chunk": chunk, "error": str(e)}), flush=True)
