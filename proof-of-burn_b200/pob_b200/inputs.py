"""Packaging of real chain data into the circuit's input JSON -- the part of the reference input generator that
follows `eth_getProof` (reference tests/main.py:65-178), without the chain client.

    inp = build_pob_input(account_proof=[bytes, ...],      # proof.accountProof, root first, leaf last
                          header_rlp=bytes,                 # rlp(block header fields)  (tests/main.py:84-122), or
                          block={...},                      #   the eth_getBlockByNumber fields: header_rlp_from_block() assembles it
                          balance=proof.balance, burn_key=..., reveal_amount=..., burn_extra_commitment=...,
                          shape=(16, 4, 16))                # (maxNumLayers, maxNodeBlocks, maxHeaderBlocks)

Unlike tests/main.py:8-11 (which still pads to the (4,.,5) test shape) the padding follows the shape that is passed,
so the result loads into main_proof_of_burn = ProofOfBurn(16,4,16,...) unchanged."""


def _rlp_item(buf, pos):
    """decode one RLP item header at pos -> (is_list, payload_start, payload_len)"""
    b = buf[pos]
    if b < 0x80:
        return False, pos, 1
    if b < 0xB8:
        return False, pos + 1, b - 0x80
    if b < 0xC0:
        n = b - 0xB7
        return False, pos + 1 + n, int.from_bytes(buf[pos + 1: pos + 1 + n], "big")
    if b < 0xF8:
        return True, pos + 1, b - 0xC0
    n = b - 0xF7
    return True, pos + 1 + n, int.from_bytes(buf[pos + 1: pos + 1 + n], "big")


def _rlp_bytes(x):
    x = bytes(x)
    if len(x) == 1 and x[0] < 0x80:
        return x
    if len(x) < 56:
        return bytes([0x80 + len(x)]) + x
    n = len(x).to_bytes((len(x).bit_length() + 7) // 8, "big")
    return bytes([0xB7 + len(n)]) + n + x


def _rlp_list(items):
    body = b"".join(items)
    if len(body) < 56:
        return bytes([0xC0 + len(body)]) + body
    n = len(body).to_bytes((len(body).bit_length() + 7) // 8, "big")
    return bytes([0xF7 + len(n)]) + n + body


HEADER_FIELDS = ("parentHash", "sha3Uncles", "miner", "stateRoot", "transactionsRoot", "receiptsRoot", "logsBloom", "difficulty", "number",
                 "gasLimit", "gasUsed", "timestamp", "extraData", "mixHash", "nonce")                          # tests/main.py:84-100
OPTIONAL_HEADER_FIELDS = ("baseFeePerGas", "withdrawalsRoot", "blobGasUsed", "excessBlobGas", "parentBeaconBlockRoot", "requestsHash")   # :102-109


def _header_field(v):
    """one header field as the byte string the reference feeds to rlp.encode (tests/main.py:111-122): byte strings (HexBytes
    or 0x-hex text) verbatim -- fixed-width hashes, the 8-byte nonce, the 256-byte bloom keep their leading zeros -- and
    integers as minimal big-endian bytes, zero as the EMPTY string (`"0x" if h == "0x0"`)."""
    if isinstance(v, (bytes, bytearray)):
        return bytes(v)
    if isinstance(v, int):
        return b"" if v == 0 else v.to_bytes((v.bit_length() + 7) // 8, "big")
    if isinstance(v, str):
        t = v[2:] if v.lower().startswith("0x") else v
        if t in ("", "0"):
            return b""
        return bytes.fromhex(t if len(t) % 2 == 0 else "0" + t)
    raise TypeError("header field of type %s" % type(v).__name__)


def header_rlp_from_block(block):
    """RLP of an execution-layer block header from the fields `eth_getBlockByNumber` returns (a dict or an object with these
    attributes): the 15 fixed fields, then the post-London/Shanghai/Cancun/Prague ones that are present, in the reference's order
    (tests/main.py:84-122).  keccak256 of the result is the block hash the circuit commits to (proof_of_burn.circom:122)."""
    get = (lambda k: block.get(k)) if isinstance(block, dict) else (lambda k: getattr(block, k, None))
    items = []
    for k in HEADER_FIELDS:
        v = get(k)
        if v is None:
            raise KeyError("block header field %r missing" % k)
        items.append(_rlp_bytes(_header_field(v)))
    for k in OPTIONAL_HEADER_FIELDS:
        v = get(k)
        if v is not None:
            items.append(_rlp_bytes(_header_field(v)))
    return _rlp_list(items)


def leaf_address_nibbles(leaf):
    """number of address-hash nibbles stored in an MPT leaf node (tests/main.py:69-77: hex-prefix flag 0x2_ even,
    0x3_ odd)"""
    is_list, start, _ = _rlp_item(leaf, 0)
    if not is_list:
        raise ValueError("leaf node is not an RLP list")
    _, kstart, klen = _rlp_item(leaf, start)
    term = leaf[kstart: kstart + klen]
    if term[0] & 0xF0 == 0x20:
        return 2 * len(term) - 2
    if term[0] & 0xF0 == 0x30:
        return 2 * len(term) - 1
    raise ValueError("not a leaf node (hex-prefix flag 0x%02x)" % term[0])


def build_pob_input(account_proof, header_rlp, balance, burn_key, reveal_amount, burn_extra_commitment, shape=(16, 4, 16),
                    intended_balance=None, byte_security_relax=0, proof_extra_commitment=0, block=None):
    """header_rlp may be None when `block` (the eth_getBlockByNumber result) is given: the header is then assembled here"""
    if header_rlp is None:
        header_rlp = header_rlp_from_block(block)
    max_layers, node_blocks, header_blocks = shape
    nb, hb = node_blocks * 136, header_blocks * 136
    if len(account_proof) > max_layers:
        raise ValueError("account proof has %d nodes, the circuit supports %d" % (len(account_proof), max_layers))
    if any(len(n) >= nb for n in account_proof) or len(header_rlp) >= hb:
        raise ValueError("a proof node or the header does not fit the circuit shape")
    layers = [list(n) + [0] * (nb - len(n)) for n in account_proof]
    lens = [len(n) for n in account_proof]
    while len(layers) < max_layers:          # unused layers: zeros, length 256 (tests/main.py:148-150)
        layers.append([0] * nb)
        lens.append(256)
    return {"numLeafAddressNibbles": str(leaf_address_nibbles(bytes(account_proof[-1]))), "burnKey": str(burn_key),
            "burnExtraCommitment": burn_extra_commitment, "actualBalance": str(balance),
            "intendedBalance": str(balance if intended_balance is None else intended_balance), "revealAmount": str(reveal_amount),
            "numLayers": len(account_proof), "layers": layers, "layerLens": lens,
            "blockHeader": list(header_rlp) + [0] * (hb - len(header_rlp)), "blockHeaderLen": len(header_rlp),
            "byteSecurityRelax": byte_security_relax, "_proofExtraCommitment": proof_extra_commitment}
