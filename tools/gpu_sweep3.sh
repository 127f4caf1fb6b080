#!/bin/bash
run() { echo "== $*"; env "$@" python - <<'PY'
import sys, os, time
sys.path.insert(0, "proof-of-burn_b200")
import numpy as np, pob_b200
from pob_b200 import synth
shape = (16, 4, 16, 50, 31, 2, 10**19, 10**20)
c = pob_b200.Circuit(pob_b200.MAIN_PROOF_OF_BURN)
insts = synth.make_batch(128, shape)
c.stage(synth.pack_instances(insts, shape))
for _ in range(2): r = c.run_packed(None, n=128, staged=True)
r = c.run_packed(None, n=128, staged=True)
t = r.timing
print("tiles %d  expand %.3f ms/launch (%d launches) = %.1f GB/s   total %.1f ms = %.1f wit/s  eval %.2f ms/launch" % (c.desc["n_tiles"], t["expand_ms"]/t["expand_launches"], t["expand_launches"], 32*c.n_signals*128/t["expand_ms"]/1e6, t["total_ms"], 128/t["total_ms"]*1e3, t["eval_ms"]/t["eval_launches"]))
PY
}
run "$@"
