"""Packaging of real chain data into the circuit's input JSON -- the part of the reference input generator that
follows `eth_getProof` (reference tests/main.py:65-178), without the chain client.

    inp = build_pob_input(account_proof=[bytes, ...],      # proof.accountProof, root first, leaf last
                          header_rlp=bytes,                 # rlp(block header fields)  (tests/main.py:84-122)
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
                    intended_balance=None, byte_security_relax=0, proof_extra_commitment=0):
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
