import copy, json, os

GOLD_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_testcases.json")
_GOLD = None


def gold():
    global _GOLD
    if _GOLD is None:
        _GOLD = json.load(open(GOLD_PATH))
    return _GOLD


def suite(name):
    return next(s for s in gold() if s["suite"] == name)


def pob_fixture():
    """tests/test_pob_input.json of the reference (shape ProofOfBurn(4,4,5,...)), via the golden dump."""
    return copy.deepcopy(suite("test_proof_of_burn")["cases"][0]["input"])


def repad_pob(inp, max_layers, node_blocks, header_blocks):
    """Re-pad a ProofOfBurn input to another circuit shape: unused layers are zero with length 256
    (convention of reference tests/main.py:148-150), header zero-extended."""
    out = copy.deepcopy(inp)
    nb, hb = node_blocks * 136, header_blocks * 136
    layers = [list(l)[:nb] + ["0"] * (nb - len(l)) for l in out["layers"]]
    lens = list(out["layerLens"])
    while len(layers) < max_layers:
        layers.append(["0"] * nb)
        lens.append("256")
    out["layers"], out["layerLens"] = layers[:max_layers], lens[:max_layers]
    hdr = list(out["blockHeader"])
    out["blockHeader"] = hdr[:hb] + ["0"] * (hb - len(hdr))
    return out


_RT = None


def _cudart():
    global _RT
    if _RT is None:
        import ctypes, glob
        _RT = ctypes.CDLL(sorted(glob.glob("/usr/local/cuda/lib64/libcudart.so.*"))[-1])
    return _RT


def cuda_poke(dptr, signal, value):
    """Fault injection for the self-check tests: overwrite ONE 32-byte witness entry at device pointer `dptr`
    (pob_witness_device_ptr) through the CUDA runtime -- the product API has no write access to a witness."""
    import ctypes
    import numpy as np
    rt = _cudart()
    v = int(value)
    limbs = np.array([(v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)], dtype=np.uint64)
    rc = rt.cudaMemcpy(ctypes.c_void_p(dptr + 32 * int(signal)), ctypes.c_void_p(limbs.ctypes.data), ctypes.c_size_t(32), ctypes.c_int(1))
    assert rc == 0, "cudaMemcpy H2D failed: %d" % rc
