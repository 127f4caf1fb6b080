#!/usr/bin/env python3
"""CPU-side tuning aid: what is in each level of the compiled k_eval program (opcode histogram, warp ops) -- read next to the
per-level clock profile of tools/eval_levels.py.  usage: level_detail.py [maxNumLayers=16] [opt=0]"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "proof-of-burn_b200")]
import numpy as np
import emu
from oracle import oracle
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
opt = int(sys.argv[2]) if len(sys.argv) > 2 else 0
name, params = oracle.parse_main("ProofOfBurn(%d, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)" % L)
prog = emu.EmuProgram(name, oracle.to_limbs(params), len(params), opt=opt)
out = np.zeros(24 * 256, dtype=np.uint32)
lib = emu.lib()
lib.pob_emu_level_detail.restype = ctypes.c_uint32
lib.pob_emu_level_detail.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
n = lib.pob_emu_level_detail(prog.h, out.ctypes.data, 256)
NAMES = {1: "fma", 2: "isz", 3: "inv", 4: "div", 5: "mod", 6: "pack8", 7: "chk_eq", 8: "chk_range", 9: "gtc", 10: "selsum"}
for i in range(n):
    o = out[24 * i:24 * i + 24]
    ops = " ".join("%s=%d" % (NAMES.get(k, "op%d" % k), o[k]) for k in range(16) if o[k])
    print("level %2d: %s | absorbs %d poseidon-segments %d (sum t %d) psums %d (longest %d, total %d)" % (i, ops, o[16], o[17], o[21], o[18], o[19], o[20]))
