"""Host logic: synthetic input generator (valid instances), index sharding, and the N>1 reductions over gloo."""
import os, sys
import numpy as np
import pytest
import torch

from oracle import oracle


def test_synth_primitives_known_answers():
    from pob_b200 import synth
    assert synth.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"   # empty_account.circom:10
    assert synth.keccak256(b"\x80").hex() == "56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421"  # :9
    assert int.from_bytes(synth.keccak256(b"EIP-7503"), "big") % synth.P == synth.POSEIDON_PREFIX            # constants.circom:3-6
    assert synth.poseidon([1, 2]) == 7853200120776062878684798364095072458815029376092732009249414926327459813530   # poseidoncircuit.js:52
    assert synth.poseidon([3, 4]) == 14763215145315200506921711489642608356394854266165572616578112107564877678998  # :62


def test_pow_pool_entries_satisfy_the_check():
    from pob_b200 import synth
    pool = synth.load_pow_pool()
    assert len(pool) >= 16
    for key, reveal, extra in pool[:8]:
        h = synth.keccak256(int(key).to_bytes(32, "big") + int(reveal).to_bytes(32, "big") + int(extra).to_bytes(32, "big") + b"EIP-7503")
        assert h[:2] == b"\x00\x00"


def test_synthetic_instances_are_accepted_by_the_oracle():
    """test-shape circuit (64 M entries) so this stays fast; the generator is shape-generic"""
    from pob_b200 import synth
    shape = (4, 4, 5, 20, 31, 2, 10 ** 18, 10 ** 19)
    insts = synth.make_batch(2, shape, seed=11)
    packed = synth.pack_instances(insts, shape)
    for i, it in enumerate(insts):
        j = synth.to_json(it, shape)
        sch = oracle.schema("ProofOfBurn", list(shape))
        assert np.array_equal(packed[i], oracle.to_limbs(oracle.flatten_inputs(sch, j)))
        w = oracle.run_flat("ProofOfBurn", list(shape), packed[i])
        try:
            assert w.ok, "synthetic instance %d rejected: status %d" % (i, w.status)
        finally:
            w.free()
    assert insts[0]["burnKey"] != insts[1]["burnKey"] or insts[0]["blockHeader"] != insts[1]["blockHeader"]


def test_shard_ranges_cover_the_batch():
    from pob_b200 import shard
    for n, world in [(8192, 8), (1000, 3), (5, 8), (1, 1)]:
        spans = [shard.shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "proof-of-burn_b200"))
    from pob_b200 import shard
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lo, hi = shard.shard_range(100, rank, world)
    times, counts = shard.reduce_timing_and_counts(dist, "cpu", [10.0 + rank, 5.0 - rank], [hi - lo, rank])
    dist.barrier()
    if rank == 0:
        out.put((times, counts))
    dist.destroy_process_group()


def test_two_rank_reduction_over_gloo():
    """the N>1 path of bench.py on CPU: max-over-ranks timing and summed counts, world_size 2, gloo"""
    import torch.multiprocessing as tmp
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    times, counts = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert times == [11.0, 5.0] and counts == [100, 1]


def test_build_pob_input_packages_a_proof_like_the_reference_generator():
    """pob_b200.inputs mirrors tests/main.py:65-178 after eth_getProof: nibble count from the leaf, shape-correct padding.
    Checked on a synthetic world (the packaged JSON must equal the generator's own JSON and be accepted by the oracle)."""
    from pob_b200 import synth, inputs
    shape = (4, 4, 5, 20, 31, 2, 10 ** 18, 10 ** 19)
    inst = synth.make_batch(1, shape, seed=21)[0]
    j = inputs.build_pob_input(inst["layers"], inst["blockHeader"], inst["actualBalance"], inst["burnKey"], inst["revealAmount"],
                               inst["burnExtraCommitment"], shape=shape[:3], proof_extra_commitment=inst["_proofExtraCommitment"])
    ref = synth.to_json(inst, shape)
    assert int(j["numLeafAddressNibbles"]) == inst["numLeafAddressNibbles"]
    for k in ref:
        assert [int(v) for v in np.ravel(j[k])] == [int(v) for v in np.ravel(ref[k])], k
    w = oracle.run("ProofOfBurn(4, 4, 5, 20, 31, 2, 10 ** 18, 10 ** 19)", j)
    assert w.ok
    w.free()
    with pytest.raises(ValueError):
        inputs.build_pob_input(inst["layers"] * 3, inst["blockHeader"], 1, 1, 1, 1, shape=shape[:3])


def test_header_rlp_assembly_reproduces_the_reference_fixture_header():
    """pob_b200.inputs.header_rlp_from_block == the header assembly of the reference generator (tests/main.py:84-122).  Test
    vector: the 612-byte header of tests/test_pob_input.json (keccak e36499b5...368d, tests/testcases/proof_of_burn.py:22): its 21
    RLP fields are taken apart here, handed back as an eth_getBlockByNumber-style record (integers for the numeric fields, hex /
    bytes for the hashes) and must re-assemble to the identical bytes -- zero difficulty and zero blob-gas fields as EMPTY strings,
    the zero miner / nonce / bloom with their full width."""
    from helpers import pob_fixture
    from pob_b200 import inputs, synth
    f = pob_fixture()
    hdr = bytes(int(v) for v in f["blockHeader"][: int(f["blockHeaderLen"])])
    assert len(hdr) == 612 and synth.keccak256(hdr).hex() == "e36499b50da290131c3fa32d4f60717c8c529ae1bc3a216f32d05c05fe80368d"
    is_list, pos, n = inputs._rlp_item(hdr, 0)
    assert is_list and pos + n == len(hdr)
    items = []
    while pos < len(hdr):
        _, st, ln = inputs._rlp_item(hdr, pos)
        items.append(hdr[st: st + ln]); pos = st + ln
    names = inputs.HEADER_FIELDS + inputs.OPTIONAL_HEADER_FIELDS
    assert len(items) == len(names) == 21
    numeric = {"difficulty", "number", "gasLimit", "gasUsed", "timestamp", "baseFeePerGas", "blobGasUsed", "excessBlobGas"}
    block = {}
    for k, v in zip(names, items):
        if k in numeric:
            block[k] = int.from_bytes(v, "big")                         # what a JSON-RPC client hands back
        elif k in ("stateRoot", "miner"):
            block[k] = "0x" + v.hex()                                   # hex text is accepted too
        else:
            block[k] = v
    assert block["difficulty"] == 0 and block["number"] == 3 and block["blobGasUsed"] == 0
    assert inputs.header_rlp_from_block(block) == hdr
    class Obj:                                                           # attribute access like web3's AttributeDict
        pass
    o = Obj(); o.__dict__.update(block)
    assert inputs.header_rlp_from_block(o) == hdr
    pre_london = {k: block[k] for k in inputs.HEADER_FIELDS}
    short = inputs.header_rlp_from_block(pre_london)
    assert len(short) < len(hdr) and inputs._rlp_item(short, 0)[0]
    with pytest.raises(KeyError):
        inputs.header_rlp_from_block({k: block[k] for k in inputs.HEADER_FIELDS[:-1]})
    # and through the packager: block= instead of header_rlp=
    shape = (4, 4, 5)
    layers = [bytes(int(v) for v in f["layers"][i][: int(f["layerLens"][i])]) for i in range(int(f["numLayers"]))]
    j = inputs.build_pob_input(layers, None, int(f["actualBalance"]), int(f["burnKey"]), int(f["revealAmount"]), int(f["burnExtraCommitment"]),
                               shape=shape, block=block)
    assert [int(v) for v in j["blockHeader"]] == [int(v) for v in f["blockHeader"]] and int(j["blockHeaderLen"]) == 612
    assert [[int(v) for v in l] for l in j["layers"]] == [[int(v) for v in l] for l in f["layers"]]
    assert int(j["numLeafAddressNibbles"]) == int(f["numLeafAddressNibbles"])


def test_bench_helpers():
    """bench.py host logic: shape expression round-trips through the layout compiler; the CPU-baseline process count is
    bounded by available memory (a witness is 6.9 GB per oracle process)."""
    import importlib.util, pob_b200
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    expr = bench.shape_expr((8,) + bench.MAIN_SHAPE[1:])
    assert pob_b200.parse_main(expr) == ("ProofOfBurn", [8, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20])
    assert bench.shape_expr(bench.MAIN_SHAPE).replace(" ", "") == bench.MAIN_EXPR.replace(" ", "").replace("10**19", str(10 ** 19)).replace("10**20", str(10 ** 20))
    assert bench.N_SIGNALS_MAIN == pob_b200.layout_info(bench.MAIN_EXPR)["n_signals"]
    assert 1 <= bench.mem_limited_procs(64, 34 * bench.N_SIGNALS_MAIN) <= 64
    assert bench.mem_limited_procs(64, 1 << 60) == 1            # nothing fits: still one process


def test_even_odd_montgomery_model():
    """The device form of fr_mont (csrc/fr_hd.h, inline PTX: two accumulators E / O, every product on an aligned limb pair, one
    carry chain per row, the shift a renaming) modelled limb by limb with Python integers -- same rows, same carry hand-overs,
    same bounds (no O row carries out, E needs 9 limbs, result < 2p) -- against a*b*2^-256 mod p.  The PTX itself runs in the
    `-m gpu` parity tests (every Poseidon value of every witness goes through it)."""
    import random
    P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    M32 = (1 << 32) - 1
    N0 = (-pow(P, -1, 1 << 32)) % (1 << 32)
    assert N0 == 0xefffffff                                           # POB_N0
    pl = [(P >> (32 * i)) & M32 for i in range(8)]

    class Chain:                                                      # a PTX carry chain (add.cc / addc.cc / mad.lo.cc / madc.hi.cc)
        def __init__(self):
            self.c = 0

        def add(self, x, y, cin):
            s = x + y + (self.c if cin else 0)
            self.c = s >> 32
            return s & M32

    def row(T, src, w, ch, first_cin):                                # T[0..7] += sum_k src[k] * w << 64k: four lo/hi pairs on one chain
        cin = first_cin
        for k in range(4):
            T[2 * k] = ch.add(T[2 * k], (src[k] * w) & M32, cin)
            cin = True
            T[2 * k + 1] = ch.add(T[2 * k + 1], (src[k] * w) >> 32, True)

    def mont(a, b):
        al = [(a >> (32 * i)) & M32 for i in range(8)]
        bl = [(b >> (32 * i)) & M32 for i in range(8)]
        E, O, x = [0] * 9, [0] * 8, 0
        for i in range(8):
            ch = Chain()                                              # mont_row_o_carry
            E[0] = ch.add(E[0], x, False)
            row(O, al[1::2], bl[i], ch, True)
            assert ch.c == 0
            ch = Chain()                                              # mont_row_e
            row(E, al[0::2], bl[i], ch, False)
            E[8] = ch.add(E[8], 0, True)
            m = (E[0] * N0) & M32
            ch = Chain()                                              # mont_row_o
            row(O, pl[1::2], m, ch, False)
            assert ch.c == 0
            ch = Chain()                                              # mont_row_e
            row(E, pl[0::2], m, ch, False)
            E[8] = ch.add(E[8], 0, True)
            assert ch.c == 0 and E[0] == 0
            x, E, O = E[1], O[:] + [0], E[2:9] + [0]                  # t >>= 32
        ch = Chain()
        r = [ch.add(E[0], x, False)] + [0] * 7
        for k in range(1, 8):
            r[k] = ch.add(E[k], O[k - 1], True)
        assert ch.c == 0 and O[7] == 0 and E[8] == 0
        v = sum(r[k] << (32 * k) for k in range(8))
        assert v < 2 * P
        return v - P if v >= P else v

    rinv = pow(1 << 256, -1, P)
    rng = random.Random(7)
    vals = [0, 1, 2, P - 1, P - 2, 1 << 253, M32, (1 << 64) - 1, (1 << 224) - 1] + [rng.randrange(P) for _ in range(3000)]
    for i, a in enumerate(vals):
        b = vals[(i * 7 + 3) % len(vals)]
        assert mont(a, b) == a * b * rinv % P
    for _ in range(500):                                              # second operand only needs to be < 2^256 (inv_chain_result's 2^e)
        a, b = rng.randrange(P), rng.randrange(1 << 256)
        assert mont(a, b) == a * b * rinv % P
