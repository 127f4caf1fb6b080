#!/usr/bin/env python3
"""Dump the reference's own gadget test tables to tests/golden/reference_testcases.json.

The reference pins the hot path only through `tests/testcases/**` (tests/test.py:146-201):
each suite is (main-template-expression, [(input_dict, expected_outputs | None), ...]).
Several suites compute their expected values with web3 / rlp / eth_abi, which are not
installed here, so this script injects minimal stand-ins (keccak-256, RLP list encoding,
abi.encodePacked(uint256...)) into sys.modules and then imports the reference test modules
UNMODIFIED from /root/reference.  The result is committed as a fixture because the reference
tree does not exist on the GPU box.

    python tools/gen_golden.py [/root/reference]
"""
import importlib, json, os, sys, types

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                   "reference_testcases.json")

# ---- keccak-256 (FIPS-202 permutation, 0x01 domain byte) -------------------------------------
RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
      0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
      0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
      0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
      0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
      0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
M64 = (1 << 64) - 1
def rol(x, n): return ((x << n) | (x >> (64 - n))) & M64 if n else x
def keccak_f(A):
    for rnd in range(24):
        C = [A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20] for x in range(5)]
        D = [C[(x + 4) % 5] ^ rol(C[(x + 1) % 5], 1) for x in range(5)]
        A = [A[i] ^ D[i % 5] for i in range(25)]
        B = [0] * 25
        x, y = 1, 0
        B[0] = A[0]
        cur = A[1]
        for t in range(24):
            X, Y = y, (2 * x + 3 * y) % 5
            B[X + 5 * Y] = rol(A[x + 5 * y], ((t + 1) * (t + 2) // 2) % 64)
            x, y = X, Y
        A = [B[i] ^ ((~B[(i % 5 + 1) % 5 + 5 * (i // 5)]) & B[(i % 5 + 2) % 5 + 5 * (i // 5)]) for i in range(25)]
        A[0] ^= RC[rnd]
    return A
def keccak256(data: bytes) -> bytes:
    rate = 136
    p = bytearray(data) + b"\x01"
    p += b"\x00" * ((-len(p)) % rate)
    p[-1] |= 0x80
    A = [0] * 25
    for off in range(0, len(p), rate):
        for i in range(17):
            A[i] ^= int.from_bytes(p[off + 8 * i: off + 8 * i + 8], "little")
        A = keccak_f(A)
    return b"".join(a.to_bytes(8, "little") for a in A[:4])
assert keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"

# ---- stand-in modules --------------------------------------------------------------------------
class _Web3:
    @staticmethod
    def keccak(primitive=None, hexstr=None, text=None):
        if hexstr is not None: primitive = bytes.fromhex(hexstr[2:] if hexstr.startswith("0x") else hexstr)
        if text is not None: primitive = text.encode()
        return keccak256(bytes(primitive))
    @staticmethod
    def to_bytes(primitive=None, hexstr=None):
        if hexstr is not None: return bytes.fromhex(hexstr[2:] if hexstr.startswith("0x") else hexstr)
        if isinstance(primitive, int): return primitive.to_bytes(max(1, (primitive.bit_length() + 7) // 8), "big")
        return bytes(primitive)
web3 = types.ModuleType("web3"); web3.Web3 = _Web3

def _rlp_len(n, off):
    if n < 56: return bytes([off + n])
    b = n.to_bytes((n.bit_length() + 7) // 8, "big")
    return bytes([off + 55 + len(b)]) + b
def rlp_encode(x):
    if isinstance(x, int):
        x = b"" if x == 0 else x.to_bytes((x.bit_length() + 7) // 8, "big")
    if isinstance(x, (bytes, bytearray)):
        x = bytes(x)
        if len(x) == 1 and x[0] < 0x80: return x
        return _rlp_len(len(x), 0x80) + x
    body = b"".join(rlp_encode(e) for e in x)
    return _rlp_len(len(body), 0xC0) + body
rlp = types.ModuleType("rlp"); rlp.encode = rlp_encode

eth_abi = types.ModuleType("eth_abi"); packed = types.ModuleType("eth_abi.packed")
def encode_packed(types_, vals):
    assert all(t == "uint256" for t in types_)
    return b"".join(int(v).to_bytes(32, "big") for v in vals)
packed.encode_packed = encode_packed; eth_abi.packed = packed
sys.modules.update({"web3": web3, "rlp": rlp, "eth_abi": eth_abi, "eth_abi.packed": packed})

# ---- import the reference suites in tests/test.py order ------------------------------------------
SUITES = [  # (module, attribute) in the order tests/test.py:146-201 runs them
    ("spend", "test_spend"), ("proof_of_work", "test_pow_eip7503_postfix"),
    ("proof_of_work", "test_concat_fixed_4"), ("proof_of_work", "test_proof_of_work"),
    ("public_commitment", "test_public_commitment_1"), ("public_commitment", "test_public_commitment_2"),
    ("public_commitment", "test_public_commitment_6"), ("poseidon", "test_poseidon_2"),
    ("poseidon", "test_poseidon_3"), ("poseidon", "test_poseidon_4"), ("divide", "test_divide"),
    ("substring_check", "test_substring_check"), ("shift", "test_shift_left"), ("shift", "test_shift_right"),
    ("concat", "test_mask"), ("concat", "test_concat"), ("selector", "test_selector"),
    ("selector", "test_selector_array_1d"), ("selector", "test_selector_array_2d"),
    ("convert", "test_big_endian_bytes_2_num"), ("convert", "test_bytes_2_nibbles"),
    ("convert", "test_little_endian_bytes_2_num"), ("convert", "test_num_2_big_endian_bytes"),
    ("convert", "test_num_2_little_endian_bytes"), ("convert", "test_nibbles_2_bytes"),
    ("convert", "test_num_2_bits_safe_32"), ("convert", "test_num_2_bits_safe_254"),
    ("convert", "test_num_2_bits_safe_256"), ("keccak", "test_pad"), ("keccak", "test_keccak_1"),
    ("keccak", "test_keccak_2"), ("burn_address", "test_burn_address"),
    ("burn_address", "test_burn_address_hash"), ("assertion", "test_assert_bits"),
    ("assertion", "test_assert_byte_string"), ("assertion", "test_assert_less_eq_than"),
    ("assertion", "test_assert_less_than"), ("assertion", "test_assert_greater_eq_than"),
    ("proof_of_burn", "test_proof_of_burn"), ("array", "test_filter"), ("array", "test_fit_1"),
    ("array", "test_fit_2"), ("array", "test_reverse"), ("array", "test_flatten"), ("array", "test_reshape"),
    ("rlp.integer", "test_rlp_integer_1"), ("rlp.integer", "test_rlp_integer_2"),
    ("rlp.integer", "test_count_bytes"), ("rlp.empty_account", "test_rlp_empty_account_1"),
    ("rlp.empty_account", "test_rlp_empty_account_2"), ("rlp.empty_account", "test_rlp_empty_account_3"),
    ("rlp.merkle_patricia_trie_leaf", "test_truncated_address_hash"),
    ("rlp.merkle_patricia_trie_leaf", "test_is_in_range"),
    ("rlp.merkle_patricia_trie_leaf", "test_leaf_detector_1"),
    ("rlp.merkle_patricia_trie_leaf", "test_leaf_detector_2"),
    ("rlp.merkle_patricia_trie_leaf", "test_rlp_merkle_patricia_trie_leaf"),
]

def strify(x):
    if isinstance(x, bool): return int(x)
    if isinstance(x, int): return str(x)
    if isinstance(x, str): return str(int(x))
    if isinstance(x, (list, tuple)): return [strify(e) for e in x]
    if isinstance(x, dict): return {k: strify(v) for k, v in x.items()}
    if hasattr(x, "val"): return str(int(x.val))
    raise TypeError(type(x))

def main():
    os.chdir(REF)
    for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    # /root/reference/tests has no __init__.py and another `tests` package may shadow it:
    # register the namespace explicitly.
    pkg = types.ModuleType("tests"); pkg.__path__ = [os.path.join(REF, "tests")]
    sys.modules["tests"] = pkg
    out = []
    for mod, attr in SUITES:
        m = importlib.import_module("tests.testcases." + mod)
        main_expr, cases = getattr(m, attr)
        out.append({"suite": attr, "source": "tests/testcases/%s.py" % mod.replace(".", "/"),
                    "main": main_expr,
                    "cases": [{"input": strify(i), "expected": (None if e is None else strify(e))}
                              for (i, e) in cases]})
    with open(OUT, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", os.path.normpath(OUT), "suites:", len(out), "cases:", sum(len(s["cases"]) for s in out))

if __name__ == "__main__":
    main()
