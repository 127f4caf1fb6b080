"""The product's layout compiler, checked on the CPU: every compiled witness program is executed by the TEST-ONLY
host emulator (tests/emu/, the host instantiation of csrc/vm_exec.h) and its complete witness vector, status code
and outputs are compared entry by entry with the independent CPU oracle.  This validates layout + program on a box
without a GPU; the CUDA kernels themselves are covered by tests/test_gpu_parity.py (-m gpu)."""
import os, sys
import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
from helpers import gold, suite, pob_fixture
from oracle import oracle

SMALL = [s for s in gold() if s["suite"] != "test_proof_of_burn"]


def _compare(main, cases, hcreate=False):
    import emu
    name, params = oracle.parse_main(main)
    pl = oracle.to_limbs(params) if params else np.zeros((1, 4), dtype=np.uint64)
    prog = emu.EmuProgram(name, pl, len(params), hcreate)
    sch = oracle.schema(name, params)
    for k, inp in enumerate(cases):
        flat = oracle.to_limbs(oracle.flatten_inputs(sch, inp))
        w = oracle.run_flat(name, params, flat, hcreate)
        try:
            st, wit, outs = prog.run(flat)
            assert prog.stats["n_signals"] == w.n_signals
            assert st == w.status, "%s case %d: status %d vs oracle %d" % (main, k, st, w.status)
            if w.ok:
                neq = np.nonzero((wit != w.limbs).any(axis=1))[0]
                assert len(neq) == 0, "%s case %d: %d entries differ, first at %d" % (main, k, len(neq), neq[0])
                assert np.array_equal(outs, w.limbs[1:1 + w.n_outputs])
        finally:
            w.free()
    return prog


@pytest.mark.parametrize("s", SMALL, ids=[s["suite"] for s in SMALL])
def test_program_matches_oracle(s):
    _compare(s["main"], [c["input"] for c in s["cases"]])


def test_proof_of_burn_program_matches_oracle():
    """ProofOfBurn(4,4,5,...) on the reference fixture and one corrupted copy: 64.4 M entries each"""
    s = suite("test_proof_of_burn")
    prog = _compare(s["main"], [s["cases"][0]["input"], s["cases"][1]["input"]])
    assert prog.stats["n_absorbs"] == 25 and prog.stats["n_round_blocks"] == 25 * 24


def test_creation_order_policy_matches_oracle():
    s = suite("test_spend")
    _compare(s["main"], [s["cases"][0]["input"]], hcreate=True)
    s = suite("test_leaf_detector_2")
    _compare(s["main"], [s["cases"][0]["input"]], hcreate=True)


def test_round_table_is_shared():
    import emu
    name, params = oracle.parse_main("KeccakBytes(2)")
    prog = emu.EmuProgram(name, oracle.to_limbs(params), len(params))
    st = prog.stats
    assert st["n_round_blocks"] == 48
    # codes = one shared 102,656-entry KeccakfRound table + the flat (non-round) signals
    assert st["n_codes"] == 102656 + st["n_signals"] - 48 * 102656


def test_fuzz_programs_match_oracle():
    """seeded random / boundary / out-of-range inputs: status codes and accepted witnesses must agree"""
    import emu
    from fuzz_cases import cases
    progs = {}
    rejected = 0
    for main, inp in cases():
        name, params = oracle.parse_main(main)
        if main not in progs:
            progs[main] = emu.EmuProgram(name, oracle.to_limbs(params) if params else np.zeros((1, 4), dtype=np.uint64), len(params))
        flat = oracle.to_limbs(oracle.flatten_inputs(oracle.schema(name, params), inp))
        w = oracle.run_flat(name, params, flat)
        try:
            st, wit, _ = progs[main].run(flat)
            assert st == w.status, "%s %s: status %d vs oracle %d" % (main, inp, st, w.status)
            if w.ok:
                assert np.array_equal(wit, w.limbs), "%s %s" % (main, inp)
            else:
                rejected += 1
        finally:
            w.free()
    assert rejected > 10


def test_binary_eea_inverse_equals_fermat():
    """the VM's inversion routine (fr_inv_eea, csrc/fr_hd.h) against the Fermat ladder on 3000 field elements"""
    import ctypes, emu
    L = emu.lib()
    L.pob_emu_inv_selftest.restype = ctypes.c_uint32
    assert L.pob_emu_inv_selftest(ctypes.c_uint32(3000)) == 0


def test_montgomery_multiplication_against_shift_and_add():
    """fr_mont / fr_mul (csrc/fr_hd.h, column-wise Montgomery) against a bit-serial (a*b) mod p on 2000 pairs incl. p-k"""
    import ctypes, emu
    L = emu.lib()
    L.pob_emu_mul_selftest.restype = ctypes.c_uint32
    assert L.pob_emu_mul_selftest(ctypes.c_uint32(2000)) == 0


def test_main_shape_program_matches_oracle_entry_by_entry():
    """BASELINE.json configs[1] on the CPU: the compiled program of main_proof_of_burn = ProofOfBurn(16,4,16,50,31,2,
    10^19,10^20), run by the test-only emulator on the reference fixture re-padded to the main shape, equals the oracle on
    all 215,907,954 witness entries (needs ~14 GB of RAM, ~70 s)."""
    from helpers import repad_pob
    expr = "ProofOfBurn(16, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)"
    prog = _compare(expr, [repad_pob(pob_fixture(), 16, 4, 16)])
    assert prog.stats["n_signals"] == 215907954 and prog.stats["n_absorbs"] == 84 and prog.stats["n_round_blocks"] == 2016


@pytest.mark.slow
@pytest.mark.parametrize("layers", [8, 12])
def test_config5_shapes_program_matches_oracle(layers):
    """BASELINE.json configs[4] shapes ProofOfBurn(L,4,16,...) on a synthetic valid instance, all S(L) entries (POB_SLOW=1)"""
    from pob_b200 import synth
    shape = (layers, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)
    inst = synth.make_batch(1, shape, seed=layers)[0]
    expr = "ProofOfBurn(%d, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)" % layers
    prog = _compare(expr, [synth.to_json(inst, shape)])
    assert prog.stats["n_signals"] == 51277058 + 10289431 * layers


def test_inverse_classes_on_valid_inputs():
    """The layout compiler sorts IsZero's inverse hints into two classes: table-sized inputs (ordinary thread ops of their level, a
    2 MB table lookup on the GPU) and the ones that need a field inversion (deferred, batch-inverted by the state machine of
    k_eval).  On valid ProofOfBurn inputs no leveled inverse may miss the table (it would pay an inline inversion inside a level),
    the deferred ones must start before the last level, and nearly all of them must really need the inversion."""
    import ctypes, emu
    from pob_b200 import synth
    shape = (4, 4, 5, 50, 31, 2, 10 ** 19, 10 ** 20)
    name, params = oracle.parse_main("ProofOfBurn(4, 4, 5, 50, 31, 2, 10 ** 19, 10 ** 20)")
    prog = emu.EmuProgram(name, oracle.to_limbs(params), len(params))
    L = emu.lib()
    L.pob_emu_inv_miss_count.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    cases = [pob_fixture()] + [synth.to_json(inst, shape) for inst in synth.make_batch(3, shape, seed=11)]
    for inp in cases:
        flat = oracle.to_limbs(oracle.flatten_inputs(oracle.schema(name, params), inp))
        out = np.zeros(8, dtype=np.uint64)
        L.pob_emu_inv_miss_count(prog.h, flat.ctypes.data, out.ctypes.data)
        leveled, deferred, leveled_miss, deferred_hit, ginv_level, n_levels = (int(v) for v in out[:6])
        assert leveled > 1000 and deferred > 100
        assert leveled_miss == 0
        assert deferred_hit * 20 < deferred
        assert ginv_level < n_levels
