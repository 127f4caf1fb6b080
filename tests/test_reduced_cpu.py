"""The reduced (`--O1`-style) witness program on the CPU (test-only emulator): every retained entry equals the --O0 witness
of the ORACLE through the witness map, outputs and accept/reject are unchanged, main inputs/outputs stay, and everything the
map drops is a copy of an earlier entry or a constant."""
import os, sys
import numpy as np
import pytest

from helpers import gold, suite, pob_fixture

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

SMALL = [s for s in gold() if s["suite"] != "test_proof_of_burn"]


def _check(s, inputs):
    import emu
    from oracle import oracle
    name, params = oracle.parse_main(s["main"])
    pl = oracle.to_limbs(params) if params else np.zeros((1, 4), dtype=np.uint64)
    full = emu.EmuProgram(name, pl, len(params))
    red = emu.EmuProgram(name, pl, len(params), opt=1)
    m, n0 = red.witness_map()
    n_io = 1 + full.stats["n_outputs"] + full.stats["n_inputs"]
    assert n0 == full.stats["n_signals"] and len(m) == red.stats["n_signals"] <= n0
    assert np.array_equal(m[:n_io], np.arange(n_io)) and (np.diff(m.astype(np.int64)) > 0).all()
    sch = oracle.schema(name, params)
    not_copies, n_ok = {}, 0
    for inp in inputs:
        flat = oracle.to_limbs(oracle.flatten_inputs(sch, inp))
        w = oracle.run_flat(name, params, flat)
        try:
            st, wit, outs = red.run(flat)
            assert st == w.status
            if not w.ok:
                continue
            assert np.array_equal(wit, w.limbs[m]), "reduced entry differs from the --O0 witness"
            assert [oracle.from_limbs(r) for r in outs] == w.outputs()
            # what was dropped: equal to an earlier entry of the same witness, or a constant of the circuit (same value for every input)
            dropped = np.setdiff1d(np.arange(n0, dtype=np.int64), m.astype(np.int64))
            if len(dropped) and n0 < 3_000_000:
                keys = [bytes(r) for r in w.limbs.view(np.uint8).reshape(n0, 32)]
                first = {}
                for i, k in enumerate(keys):
                    first.setdefault(k, i)
                for i in dropped:
                    if first[keys[i]] == i:
                        not_copies.setdefault(int(i), set()).add(keys[i])
                n_ok += 1
        finally:
            w.free()
    varying = [i for i, vals in not_copies.items() if len(vals) > 1]
    assert not varying, "dropped entries that are neither copies of earlier entries nor constants: %s" % varying[:10]
    return red.stats["n_signals"], n0


@pytest.mark.parametrize("s", SMALL, ids=[s["suite"] for s in SMALL])
def test_reduced_witness_equals_o0_through_the_map(s):
    cases = [c["input"] for c in s["cases"]]
    big = oracle_size(s) > 1_000_000
    _check(s, cases[:2] if big else cases)


def oracle_size(s):
    import pob_b200
    return pob_b200.layout_info(s["main"])["n_signals"]


def test_reduced_proof_of_burn_fixture():
    s = suite("test_proof_of_burn")
    n1, n0 = _check(s, [s["cases"][0]["input"], s["cases"][1]["input"]])
    assert (n1, n0) == (6409863, 64355038)


def test_reduced_sizes():
    import pob_b200
    d = pob_b200.layout_info(pob_b200.MAIN_PROOF_OF_BURN, opt=1)
    assert (d["n_signals"], d["n_signals_o0"], d["opt_level"]) == (21454051, 215907954, 1)
    d = pob_b200.layout_info("Spend(31)", opt=1)
    assert (d["n_signals"], d["n_signals_o0"]) == (259945, 2603360)
