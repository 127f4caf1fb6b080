"""Synthetic, VALID proof-of-burn inputs shaped like the reference's tests/test_pob_input.json.

The reference builds its input by burning ETH on a local chain and asking the node for an account proof
(tests/main.py:16-178); no chain exists here, so this module constructs a consistent world directly
(SURVEY.md 8(d), config 3): a burn key that satisfies the proof-of-work check, the burn address
Poseidon4(prefix, burnKey, revealAmount, burnExtraCommitment)[:20], an RLP leaf for that address with the
chosen balance, a chain of branch-shaped inner nodes each embedding the keccak of its child, and a block
header whose bytes 91..123 are the state root.  Everything the circuit checks holds, so every generated
instance is accepted (status 0) and exercises the whole hot path.

Pure Python/numpy; no dependence on the oracle.
"""
import json
import os
import re

import numpy as np

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
POSEIDON_PREFIX = 5265656504298861414514317065875120428884240036965045859626767452974705356670  # utils/constants.circom:4-5
_HERE = os.path.dirname(os.path.abspath(__file__))
M64 = (1 << 64) - 1

# ---- keccak-256 ---------------------------------------------------------------------------------------------
RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
      0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
      0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
      0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_PIL = []        # (src lane, dst lane, rotation) of the rho-pi walk
_x, _y = 1, 0
for _t in range(24):
    _X, _Y = _y, (2 * _x + 3 * _y) % 5
    _PIL.append((_x + 5 * _y, _X + 5 * _Y, ((_t + 1) * (_t + 2) // 2) % 64))
    _x, _y = _X, _Y


def keccak_f(A):
    for rnd in range(24):
        C = [A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20] for x in range(5)]
        D = [C[(x + 4) % 5] ^ (((C[(x + 1) % 5] << 1) | (C[(x + 1) % 5] >> 63)) & M64) for x in range(5)]
        A = [A[i] ^ D[i % 5] for i in range(25)]
        B = [0] * 25
        B[0] = A[0]
        for s, d, r in _PIL:
            v = A[s]
            B[d] = ((v << r) | (v >> (64 - r))) & M64
        A = [B[i] ^ ((~B[(i % 5 + 1) % 5 + 5 * (i // 5)]) & B[(i % 5 + 2) % 5 + 5 * (i // 5)] & M64) for i in range(25)]
        A[0] ^= RC[rnd]
    return A


def keccak256(data):
    p = bytearray(data) + b"\x01"
    p += b"\x00" * ((-len(p)) % 136)
    p[-1] |= 0x80
    A = [0] * 25
    for off in range(0, len(p), 136):
        for i in range(17):
            A[i] ^= int.from_bytes(p[off + 8 * i: off + 8 * i + 8], "little")
        A = keccak_f(A)
    return b"".join(a.to_bytes(8, "little") for a in A[:4])


def _keccak_f_np(A):
    """keccak-f over a list of 25 uint64 numpy arrays (vectorised over candidates)"""
    u = np.uint64
    for rnd in range(24):
        C = [A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20] for x in range(5)]
        D = [C[(x + 4) % 5] ^ ((C[(x + 1) % 5] << u(1)) | (C[(x + 1) % 5] >> u(63))) for x in range(5)]
        A = [A[i] ^ D[i % 5] for i in range(25)]
        B = [None] * 25
        B[0] = A[0]
        for s, d, r in _PIL:
            B[d] = (A[s] << u(r)) | (A[s] >> u(64 - r))
        A = [B[i] ^ (~B[(i % 5 + 1) % 5 + 5 * (i // 5)] & B[(i % 5 + 2) % 5 + 5 * (i // 5)]) for i in range(25)]
        A[0] = A[0] ^ u(RC[rnd])
    return A


def find_burn_key(reveal, burn_extra, zero_bytes, rng, span=1 << 17):
    """grind burnKey so that keccak(burnKey|reveal|extra|"EIP-7503") starts with `zero_bytes` zero bytes
    (reference tests/main.py:47-56, circuits/utils/proof_of_work.circom:54-81); vectorised with numpy."""
    postfix = int(reveal).to_bytes(32, "big") + int(burn_extra).to_bytes(32, "big") + b"EIP-7503"
    while True:
        hi_part = int.from_bytes(rng.bytes(24), "big") >> 3          # 189 bits => burnKey < 2^253 < p
        start = (hi_part << 64) | int(rng.integers(0, 1 << 62))      # low word cannot overflow while grinding
        hi = start.to_bytes(32, "big")[:24]
        low = np.uint64(start & M64) + np.arange(span, dtype=np.uint64)
        msg = bytearray(136)
        msg[0:24] = hi
        msg[32:32 + len(postfix)] = postfix
        msg[104] = 0x01
        msg[135] |= 0x80
        A = [np.full(span, int.from_bytes(msg[8 * i: 8 * i + 8], "little"), dtype=np.uint64) for i in range(17)]
        A += [np.zeros(span, dtype=np.uint64) for _ in range(8)]
        A[3] = low.byteswap()          # bytes 24..31 = low 64 bits, big-endian, loaded as a little-endian lane
        out0 = _keccak_f_np(A)[0]
        mask = np.uint64((1 << (8 * zero_bytes)) - 1)
        hits = np.nonzero((out0 & mask) == 0)[0]
        if len(hits):
            key = (start & ~M64) | int(low[hits[0]])
            assert not any(keccak256(key.to_bytes(32, "big") + postfix)[:zero_bytes])
            return key


# ---- Poseidon (circomlib optimised schedule, circomlib/circuits/poseidon.circom:67-196) ----------------------
_PCONST = None


def _poseidon_consts():
    global _PCONST
    if _PCONST is None:
        src = open(os.path.join(_HERE, "..", "csrc", "poseidon_constants_data.h")).read()
        tabs = {}
        for m in re.finditer(r"POSEIDON_([CSMP])_T(\d)\[\d+\]\[4\] = \{(.*?)\};", src, re.S):
            rows = re.findall(r"\{(0x[0-9a-f]+)ULL, (0x[0-9a-f]+)ULL, (0x[0-9a-f]+)ULL, (0x[0-9a-f]+)ULL\}", m.group(3))
            tabs[(m.group(1), int(m.group(2)))] = [int(a, 16) | int(b, 16) << 64 | int(c, 16) << 128 | int(d, 16) << 192 for a, b, c, d in rows]
        _PCONST = tabs
    return _PCONST


def poseidon(inputs):
    t = len(inputs) + 1
    K = _poseidon_consts()
    C, S, M, Pm = K[("C", t)], K[("S", t)], K[("M", t)], K[("P", t)]
    rp = {3: 57, 4: 56, 5: 60}[t]
    st = [0] + [int(v) % P for v in inputs]
    st = [(st[i] + C[i]) % P for i in range(t)]
    mix = lambda s, Mx: [sum(Mx[j * t + i] * s[j] for j in range(t)) % P for i in range(t)]
    for r in range(3):
        st = [pow(v, 5, P) for v in st]
        st = [(st[i] + C[(r + 1) * t + i]) % P for i in range(t)]
        st = mix(st, M)
    st = [pow(v, 5, P) for v in st]
    st = [(st[i] + C[4 * t + i]) % P for i in range(t)]
    st = mix(st, Pm)
    for r in range(rp):
        st[0] = (pow(st[0], 5, P) + C[5 * t + r]) % P
        s0 = sum(S[(2 * t - 1) * r + i] * st[i] for i in range(t)) % P
        st = [s0] + [(st[i] + st[0] * S[(2 * t - 1) * r + t + i - 1]) % P for i in range(1, t)]
    for r in range(3):
        st = [pow(v, 5, P) for v in st]
        st = [(st[i] + C[5 * t + rp + r * t + i]) % P for i in range(t)]
        st = mix(st, M)
    st = [pow(v, 5, P) for v in st]
    return sum(M[j * t] * st[j] for j in range(t)) % P


# ---- RLP -----------------------------------------------------------------------------------------------------
def _rlp_len(n, off):
    if n < 56:
        return bytes([off + n])
    b = n.to_bytes((n.bit_length() + 7) // 8, "big")
    return bytes([off + 55 + len(b)]) + b


def rlp_encode(x):
    if isinstance(x, int):
        x = b"" if x == 0 else x.to_bytes((x.bit_length() + 7) // 8, "big")
    if isinstance(x, (bytes, bytearray)):
        x = bytes(x)
        return x if (len(x) == 1 and x[0] < 0x80) else _rlp_len(len(x), 0x80) + x
    body = b"".join(rlp_encode(e) for e in x)
    return _rlp_len(len(body), 0xC0) + body


EMPTY_STORAGE = bytes.fromhex("56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421")   # rlp/empty_account.circom:9
EMPTY_CODE = bytes.fromhex("c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470")      # :10


def leaf_node(addr_hash, n_nibbles, balance):
    """RLP([hex-prefix(last n nibbles of keccak(address)), RLP([0, balance, EMPTY_STORAGE, EMPTY_CODE])])
    (rlp/merkle_patricia_trie_leaf.circom:102-189)"""
    nibs = [b >> 4 if k == 0 else b & 15 for b in addr_hash for k in (0, 1)][64 - n_nibbles:]
    if n_nibbles % 2:
        key = bytes([0x30 | nibs[0]] + [nibs[i] << 4 | nibs[i + 1] for i in range(1, n_nibbles, 2)])
    else:
        key = bytes([0x20] + [nibs[i] << 4 | nibs[i + 1] for i in range(0, n_nibbles, 2)])
    return rlp_encode([key, rlp_encode([0, balance, EMPTY_STORAGE, EMPTY_CODE])])


# ---- one synthetic instance ------------------------------------------------------------------------------------
def burn_address(burn_key, reveal, burn_extra):
    return poseidon([POSEIDON_PREFIX, burn_key, reveal, burn_extra]).to_bytes(32, "big")[:20]   # burn_address.circom:55-57


def load_pow_pool():
    path = os.path.join(_HERE, "pow_pool.json")
    return json.load(open(path)) if os.path.exists(path) else []


def make_instance(rng, shape, pow_triple=None, num_layers=None, nibbles=None):
    """shape = (maxNumLayers, maxNodeBlocks, maxHeaderBlocks, minLeafAddressNibbles, amountBytes, powMinimumZeroBytes,
    maxIntendedBalance, maxActualBalance).  Returns a dict with bytes/ints (see pack_instances / to_json)."""
    L, nbk, hbk, min_nib, _ab, pow_zero, max_intended, _max_actual = shape
    if pow_triple is None:
        reveal, extra = int(rng.integers(0, 1 << 40)), int(rng.integers(0, 1 << 62))
        burn_key = find_burn_key(reveal, extra, pow_zero, rng)
    else:
        burn_key, reveal, extra = (int(v) for v in pow_triple)
    balance = max(reveal + 1, int(rng.integers(1, min(max_intended, 1 << 62))))
    balance = min(balance, max_intended)
    n_layers = int(num_layers if num_layers is not None else rng.integers(min(8, L), min(10, L) + 1))
    n_nib = int(nibbles if nibbles is not None else rng.integers(max(min_nib, 54), 61))
    addr_hash = keccak256(burn_address(burn_key, reveal, extra))
    layers = [leaf_node(addr_hash, n_nib, balance)]
    node_cap = min(532, nbk * 136 - 1)
    for _ in range(n_layers - 1):
        child = keccak256(layers[0])
        body = bytearray(rng.integers(0, 256, node_cap, dtype=np.uint8).tobytes())
        body[0:3] = b"\xf9\x02\x11"
        off = int(rng.integers(3, node_cap - 32))
        body[off:off + 32] = child
        layers.insert(0, bytes(body))
    hlen = int(min(hbk * 136 - 1, max(124, rng.integers(603, 684))))
    header = bytearray(rng.integers(0, 256, hlen, dtype=np.uint8).tobytes())
    header[91:123] = keccak256(layers[0])
    return {"burnKey": burn_key, "actualBalance": balance, "intendedBalance": balance, "revealAmount": reveal,
            "burnExtraCommitment": extra, "numLeafAddressNibbles": n_nib, "layers": layers, "numLayers": n_layers,
            "blockHeader": bytes(header), "byteSecurityRelax": 0, "_proofExtraCommitment": int(rng.integers(0, 1 << 62))}


def to_json(inst, shape):
    """the reference's input.json schema (tests/main.py:160-178): decimal strings / ints, zero padding, unused layers len 256"""
    L, nbk, hbk = shape[0], shape[1], shape[2]
    layers = [list(l) + [0] * (nbk * 136 - len(l)) for l in inst["layers"]] + [[0] * (nbk * 136)] * (L - len(inst["layers"]))
    lens = [len(l) for l in inst["layers"]] + [256] * (L - len(inst["layers"]))
    hdr = list(inst["blockHeader"]) + [0] * (hbk * 136 - len(inst["blockHeader"]))
    return {"burnKey": str(inst["burnKey"]), "actualBalance": str(inst["actualBalance"]), "intendedBalance": str(inst["intendedBalance"]),
            "revealAmount": str(inst["revealAmount"]), "burnExtraCommitment": inst["burnExtraCommitment"],
            "numLeafAddressNibbles": str(inst["numLeafAddressNibbles"]), "layers": layers, "layerLens": lens,
            "numLayers": inst["numLayers"], "blockHeader": hdr, "blockHeaderLen": len(inst["blockHeader"]),
            "byteSecurityRelax": inst["byteSecurityRelax"], "_proofExtraCommitment": inst["_proofExtraCommitment"]}


def _put(arr, row, v):
    v = int(v) % P
    arr[row, 0] = v & M64
    if v >> 64:
        arr[row, 1] = (v >> 64) & M64; arr[row, 2] = (v >> 128) & M64; arr[row, 3] = (v >> 192) & M64


def pack_instances(insts, shape, out=None):
    """fast path of Circuit.pack for ProofOfBurn inputs: (n, n_inputs, 4) uint64 in declaration order
    (circuits/proof_of_burn.circom:43-72)"""
    L, nbk, hbk = shape[0], shape[1], shape[2]
    NB, HB = nbk * 136, hbk * 136
    n_in = 6 + L * NB + L + 1 + HB + 3
    arr = np.zeros((len(insts), n_in, 4), dtype=np.uint64) if out is None else out
    arr[...] = 0
    for i, it in enumerate(insts):
        a = arr[i]
        for k, name in enumerate(("burnKey", "actualBalance", "intendedBalance", "revealAmount", "burnExtraCommitment", "numLeafAddressNibbles")):
            _put(a, k, it[name])
        base = 6
        for j, l in enumerate(it["layers"]):
            a[base + j * NB: base + j * NB + len(l), 0] = np.frombuffer(l, dtype=np.uint8)
        lens = base + L * NB
        for j in range(L):
            a[lens + j, 0] = len(it["layers"][j]) if j < len(it["layers"]) else 256
        a[lens + L, 0] = it["numLayers"]
        hdr = lens + L + 1
        a[hdr: hdr + len(it["blockHeader"]), 0] = np.frombuffer(it["blockHeader"], dtype=np.uint8)
        a[hdr + HB, 0] = len(it["blockHeader"])
        _put(a, hdr + HB + 1, it["byteSecurityRelax"])
        _put(a, hdr + HB + 2, it["_proofExtraCommitment"])
    return arr


def _make_one(args):
    i, shape, seed, triple = args
    return make_instance(np.random.default_rng(seed + i), shape, triple)


def make_batch(n, shape, seed=7503, pool=None, workers=None):
    """n distinct valid instances (instance i depends on seed + i only); proof-of-work triples come from the pre-ground
    pool (pow_pool.json) when present.  Large batches are generated by a fork pool of host processes."""
    pool = load_pow_pool() if pool is None else pool
    jobs = [(i, shape, seed, pool[i % len(pool)] if pool else None) for i in range(n)]
    if workers is None:
        try:
            workers = min(32, len(os.sched_getaffinity(0)))
        except Exception:
            workers = 1
    if n < 128 or workers <= 1:
        return [_make_one(j) for j in jobs]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(workers) as p:
        return p.map(_make_one, jobs, chunksize=max(1, n // (4 * workers)))
