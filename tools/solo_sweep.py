#!/usr/bin/env python3
"""k_eval: small levels run by CTA 0 of the cluster alone (POB_EVAL_SOLO = largest number of thread ops of such a level; 0 = every
level is spread over the whole cluster and ends with the cluster barrier).  TUNING build.  Single-witness latency, eval time of a
32-instance chunk, batch throughput; digests compared with the first row."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proof-of-burn_b200"))
import numpy as np
import pob_b200
from pob_b200 import synth

pob_b200.LIB_PATH = os.path.join(ROOT, "proof-of-burn_b200", "pob_b200", "libpob_b200_tuning.so")
shape = (16, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)
packed = synth.pack_instances(synth.make_batch(256, shape, seed=5), shape)
ref = None
for solo in [int(v) for v in sys.argv[1:]] or [0, 1024, 4096, 16384]:
    os.environ["POB_EVAL_SOLO"] = str(solo)
    c = pob_b200.Circuit(pob_b200.MAIN_PROOF_OF_BURN)
    one = packed[:1]
    c.run_packed(one)
    lat = sorted((c.run_packed(one).timing for _ in range(7)), key=lambda t: t["total_ms"])[3]
    c.run_packed(packed[:32], expand=False)
    ev = c.run_packed(packed[:32], expand=False).timing
    c.stage(packed)
    c.run_packed(None, n=256, staged=True, discard=True)
    th = c.run_packed(None, n=256, staged=True, discard=True).timing
    dg = c.run_packed(packed[:4], digest=True)
    if ref is None:
        ref = dg.digests.copy()
    print(json.dumps({"solo_ops": solo, "latency_ms": round(lat["total_ms"], 3), "lat_eval_ms": round(lat["eval_ms"], 3), "eval32_kernel_ms": round(ev["eval_ms"], 3),
                      "batch256_wit_s": round(256 / (th["total_ms"] / 1e3), 1), "ok": bool((dg.status == 0).all()), "digests_equal_first": bool(np.array_equal(dg.digests, ref))}), flush=True)
    c.close()
