#!/usr/bin/env python3
"""Pre-grind proof-of-work triples (burnKey, revealAmount, burnExtraCommitment) whose
keccak(burnKey|revealAmount|burnExtraCommitment|"EIP-7503") starts with two zero bytes
(circuits/utils/proof_of_work.circom:54-81, powMinimumZeroBytes = 2 in main_proof_of_burn.circom:27).
Grinding costs ~1 s per triple in numpy, so the bench/test input generator (pob_b200/synth.py) draws from
this committed pool instead of grinding inside a timed run.

    python tools/gen_pow_pool.py [count=128] [seed=7503]
"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "proof-of-burn_b200"))
from pob_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 7503
rng = np.random.default_rng(seed)
pool = []
for i in range(n):
    reveal, extra = int(rng.integers(0, 1 << 40)), int(rng.integers(0, 1 << 62))
    key = synth.find_burn_key(reveal, extra, 2, rng)
    pool.append([str(key), str(reveal), str(extra)])
out = os.path.join(os.path.dirname(os.path.abspath(synth.__file__)), "pow_pool.json")
json.dump(pool, open(out, "w"), indent=0)
print("wrote", out, len(pool))
