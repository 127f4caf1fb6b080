#!/usr/bin/env python3
"""per-level clock profile of k_eval for one instance (TUNING build, POB_EVAL_PROFILE): where does the single-witness latency go?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proof-of-burn_b200"))
import pob_b200
from pob_b200 import synth
pob_b200.LIB_PATH = os.path.join(ROOT, "proof-of-burn_b200", "pob_b200", "libpob_b200_tuning.so")
shape = (16, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)
packed = synth.pack_instances(synth.make_batch(32, shape, seed=5), shape)
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
for n, tag in ((1, "single_c8"), (32, "chunk32_c4")):
    os.environ["POB_EVAL_PROFILE"] = os.path.join(out, "eval_levels_%s.txt" % tag)
    c = pob_b200.Circuit(pob_b200.MAIN_PROOF_OF_BURN, max_slots=2)
    c.run_packed(packed[:n], expand=False)
    r = c.run_packed(packed[:n], expand=False)
    print(tag, r.timing)
    c.close()
    rows = [l.split() for l in open(os.environ["POB_EVAL_PROFILE"])]
    tot = sum(int(r[-1]) for r in rows)
    top = sorted(rows, key=lambda r: -int(r[-1]))[:12]
    print("total cycles", tot)
    for r in top:
        print("  ", " ".join(r), "%.1f%%" % (100.0 * int(r[-1]) / tot))
