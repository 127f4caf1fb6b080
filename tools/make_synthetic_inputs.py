#!/usr/bin/env python3
"""Write synthetic, VALID proof-of-burn input.json files (the schema of reference tests/main.py:160-178) for a circuit shape.

    python tools/make_synthetic_inputs.py OUTDIR [count=4] [maxNumLayers=16] [seed=7503]

Each file loads unchanged into `python -m pob_b200 main_proof_of_burn OUTDIR/input_0000.json witness.wtns` (for
maxNumLayers = 16) or into the reference calculator built for the same shape."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "proof-of-burn_b200"))
from pob_b200 import synth

out = sys.argv[1]
count = int(sys.argv[2]) if len(sys.argv) > 2 else 4
layers = int(sys.argv[3]) if len(sys.argv) > 3 else 16
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 7503
shape = (layers, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)
os.makedirs(out, exist_ok=True)
for i, inst in enumerate(synth.make_batch(count, shape, seed=seed)):
    path = os.path.join(out, "input_%04d.json" % i)
    json.dump(synth.to_json(inst, shape), open(path, "w"))
    print(path, "numLayers", inst["numLayers"], "nibbles", inst["numLeafAddressNibbles"])
