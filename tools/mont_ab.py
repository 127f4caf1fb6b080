#!/usr/bin/env python3
"""A/B of the device Montgomery product (csrc/fr_hd.h): inline-PTX even/odd form (libpob_b200_tuning.so) against the portable C++
form compiled for the device (libpob_b200_portable.so, `make -C proof-of-burn_b200/csrc portable`).  Single-witness latency, eval
time of a 32-instance chunk, batch-256 generation throughput; digests must agree."""
import os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:                                  # child: one library per process (both export the same symbols)
    sys.path.insert(0, os.path.join(ROOT, "proof-of-burn_b200"))
    import numpy as np
    import pob_b200
    from pob_b200 import synth
    pob_b200.LIB_PATH = os.path.join(ROOT, "proof-of-burn_b200", "pob_b200", sys.argv[1])
    shape = (16, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)
    packed = synth.pack_instances(synth.make_batch(256, shape, seed=5), shape)
    c = pob_b200.Circuit(pob_b200.MAIN_PROOF_OF_BURN)
    one = packed[:1]
    c.run_packed(one)
    lat = sorted((c.run_packed(one).timing for _ in range(7)), key=lambda t: t["total_ms"])[3]
    c.run_packed(packed[:32], expand=False)
    ev = c.run_packed(packed[:32], expand=False).timing
    c.stage(packed)
    c.run_packed(None, n=256, staged=True, discard=True)
    th = min((c.run_packed(None, n=256, staged=True, discard=True).timing for _ in range(2)), key=lambda t: t["total_ms"])
    dg = c.run_packed(packed[:4], digest=True)
    print(json.dumps({"lib": sys.argv[1], "latency_ms": round(lat["total_ms"], 3), "lat_eval_ms": round(lat["eval_ms"], 3), "eval32_kernel_ms": round(ev["eval_ms"], 3),
                      "batch256_wit_s": round(256 / (th["total_ms"] / 1e3), 1), "ok": bool((dg.status == 0).all()), "digests": [int(v) for v in dg.digests]}), flush=True)
    c.close()
else:
    rows = []
    for lib in ("libpob_b200_portable.so", "libpob_b200_tuning.so"):
        out = subprocess.run([sys.executable, __file__, lib], capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(json.dumps({"lib": lib, "error": out.stderr[-400:]}))
            continue
        rows.append(json.loads(line[-1]))
    same = len(rows) == 2 and rows[0]["digests"] == rows[1]["digests"]
    for r in rows:
        r.pop("digests")
        r["digests_equal"] = same
        print(json.dumps(r), flush=True)
