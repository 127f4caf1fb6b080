"""The constraint emitter (csrc/compiler.cpp, `want_constraints`) against ORACLE witnesses, on the CPU: every `<==` / `===`
of the circom sources, written over witness indices, must hold on the witness the independent oracle produced, must touch
every witness entry, and must notice a change of any entry the circuit pins.  (The on-GPU evaluation of the same records --
pob_selfcheck -- is covered by tests/test_gpu_selfcheck.py.)"""
import os, sys
import numpy as np
import pytest

from helpers import gold, suite, pob_fixture

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

SMALL = [s for s in gold() if s["suite"] != "test_proof_of_burn"]


def _limbs(params):
    from oracle import oracle
    return oracle.to_limbs(params) if params else np.zeros((1, 4), dtype=np.uint64)


@pytest.mark.parametrize("s", SMALL, ids=[s["suite"] for s in SMALL])
def test_oracle_witness_satisfies_every_constraint(s):
    import emu
    from oracle import oracle
    name, params = oracle.parse_main(s["main"])
    done = 0
    for case in s["cases"]:
        if case["expected"] is None:
            continue
        w = oracle.run(s["main"], case["input"])
        try:
            r = emu.check_constraints(name, _limbs(params), len(params), w.limbs)
            assert r["n_failed"] == 0 and r["n_hint_failed"] == 0, "%s: record %d fails" % (s["suite"], r["first_failed"])
            unused = {"test_fit_1": 2}.get(s["suite"], 0)          # Fit(5, 3) as a main component never reads in[3], in[4]
            assert r["signals_referenced"] == w.n_signals - unused, "%d of %d witness entries appear in no constraint" % (w.n_signals - r["signals_referenced"], w.n_signals)
        finally:
            w.free()
        done += 1
        if w.n_signals > 1_000_000 and done >= 2:
            break
    assert done


def test_proof_of_burn_fixture_satisfies_all_64m_constraints():
    import emu
    from oracle import oracle
    s = suite("test_proof_of_burn")
    name, params = oracle.parse_main(s["main"])
    w = oracle.run(s["main"], pob_fixture())
    try:
        r = emu.check_constraints(name, _limbs(params), len(params), w.limbs)
        assert r["n_failed"] == 0 and r["n_hint_failed"] == 0 and r["signals_referenced"] == w.n_signals == 64355038
        assert r["round_blocks"] == 600 and (r["round_eq"], r["round_kc"], r["round_r1"]) == (89216, 1920, 9920)
    finally:
        w.free()


POKE = [("test_leaf_detector_2", 150), ("test_rlp_empty_account_3", 150), ("test_truncated_address_hash", 712), ("test_substring_check", 659),
        ("test_poseidon_4", 300), ("test_num_2_bits_safe_256", 300), ("test_rlp_merkle_patricia_trie_leaf", 200), ("test_divide", 191),
        ("test_selector_array_2d", 149), ("test_concat", 530), ("test_keccak_1", 40)]


@pytest.mark.parametrize("sname,nprobe", POKE, ids=[p[0] for p in POKE])
def test_a_changed_entry_is_noticed(sname, nprobe):
    """fault injection on the CPU: add 1 to one witness entry at a time; some constraint (or hint record) must fail"""
    import emu
    from oracle import oracle
    s = suite(sname)
    name, params = oracle.parse_main(s["main"])
    case = [c for c in s["cases"] if c["expected"] is not None][-1]
    w = oracle.run(s["main"], case["input"])
    W = w.limbs.copy(); n = w.n_signals
    w.free()
    rng = np.random.default_rng(len(sname))
    missed = []
    for i in rng.choice(np.arange(1, n), size=min(nprobe, n - 1), replace=False):
        old = W[i].copy()
        W[i] = oracle.to_limbs([(oracle.from_limbs(old) + 1) % oracle.P])[0]
        r = emu.check_constraints(name, _limbs(params), len(params), W)
        if r["n_failed"] + r["n_hint_failed"] == 0:
            missed.append(int(i))
        W[i] = old
    assert not missed, "entries whose change no constraint notices: %s" % missed[:20]


def test_constraint_counts_of_the_main_shape():
    """size of the system a circom --O0 --r1cs run would report for main_proof_of_burn (for anyone with circom: compare)"""
    import pob_b200
    r = pob_b200.constraint_info(pob_b200.MAIN_PROOF_OF_BURN)
    assert r["signals_read"] == 215907954 and r["n_constraints"] == 215962293 and r["n_nonlinear"] == 17910859 and r["n_hints"] == 256010


def test_constraints_follow_the_numbering_policy():
    """the constraint records are stated over witness indices, so they hold on a witness of the SAME numbering policy only: the
    creation-order (hcreate) system accepts the oracle's hcreate witness, the default system rejects it"""
    import emu
    from oracle import oracle
    s = suite("test_num_2_bits_safe_256")
    name, params = oracle.parse_main(s["main"])
    w = oracle.run(s["main"], s["cases"][0]["input"], hcreate=True)
    try:
        same = emu.check_constraints(name, _limbs(params), len(params), w.limbs, hcreate=True)
        other = emu.check_constraints(name, _limbs(params), len(params), w.limbs, hcreate=False)
        assert same["n_failed"] == 0 and same["n_hint_failed"] == 0 and same["signals_referenced"] == w.n_signals
        assert other["n_failed"] > 0
    finally:
        w.free()
