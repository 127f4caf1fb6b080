"""Seeded random inputs for gadget templates, including out-of-range / huge-field values that must be REJECTED the
same way by the oracle, the emulated VM and the CUDA path (status code equality) and bit-exact when accepted."""
import numpy as np

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _r(rng, hi):
    if hi <= (1 << 62):
        return int(rng.integers(0, hi))
    return int.from_bytes(rng.bytes(40), "big") % hi


def _weird(rng, hi):
    """mostly in-range values, sometimes boundary / huge ones"""
    k = _r(rng, 10)
    if k == 0: return hi
    if k == 1: return hi - 1 if hi > 0 else 0
    if k == 2: return P - 1 - _r(rng, 3)
    if k == 3: return (1 << 253) + _r(rng, 1 << 30)
    return _r(rng, max(1, hi))


def cases(seed=1234, n=6):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        out.append(("Selector(7)", {"vals": [_weird(rng, 1 << 20) for _ in range(7)], "select": _weird(rng, 7)}))
        out.append(("Pad(3, 4)", {"in": [_weird(rng, 256) for _ in range(12)], "inLen": _weird(rng, 12)}))
        out.append(("Divide(16)", {"a": _weird(rng, 1 << 16), "b": _weird(rng, 1 << 8) or 1}))
        out.append(("ShiftLeft(8)", {"in": [_r(rng, 256) for _ in range(8)], "count": _weird(rng, 9)}))
        out.append(("ShiftRight(8, 3)", {"in": [_r(rng, 256) for _ in range(8)], "count": _weird(rng, 4)}))
        out.append(("Concat(5,5)", {"a": [_r(rng, 256) for _ in range(5)], "aLen": _weird(rng, 6), "b": [_r(rng, 256) for _ in range(5)], "bLen": _weird(rng, 6)}))
        out.append(("SubstringCheck(10, 3)", {"mainInput": [_r(rng, 4) for _ in range(10)], "mainLen": _weird(rng, 11), "subInput": [_r(rng, 4) for _ in range(3)]}))
        out.append(("Num2BitsSafe(254)", {"in": _weird(rng, 1 << 200)}))
        out.append(("Num2BigEndianBytes(32)", {"in": _weird(rng, 1 << 250)}))
        out.append(("RlpInteger(3)", {"in": _weird(rng, 1 << 24)}))
        out.append(("RlpEmptyAccount(10)", {"balance": _weird(rng, 1 << 80)}))
        out.append(("LeafDetector(16)", {"layer": [_weird(rng, 256) for _ in range(16)], "layerLen": _weird(rng, 17)}))
        out.append(("IsInRange(16)", {"lower": _weird(rng, 1 << 16), "value": _weird(rng, 1 << 16), "upper": _weird(rng, 1 << 16)}))
        out.append(("AssertGreaterEqThan(3)", {"a": _weird(rng, 8), "b": _weird(rng, 8)}))
        out.append(("Filter(5)", {"in": _weird(rng, 6)}))
        out.append(("TruncatedAddressHash(3)", {"addressHashNibbles": [_weird(rng, 16) for _ in range(6)], "addressHashNibblesLen": _weird(rng, 7)}))
        out.append(("Poseidon(3)", {"inputs": [_weird(rng, 1 << 250) for _ in range(3)]}))
        out.append(("KeccakBytes(1)", {"in": [_r(rng, 256) for _ in range(136)], "inLen": _weird(rng, 137)}))
        out.append(("PublicCommitment(2)", {"in": [[_r(rng, 256) for _ in range(32)] for _ in range(2)]}))
        out.append(("Spend(31)", {"burnKey": _weird(rng, 1 << 250), "balance": _weird(rng, 1 << 60), "withdrawnBalance": _weird(rng, 1 << 60), "extraCommitment": _weird(rng, 1 << 62)}))
    return out
