"""Pins the CPU oracle against every test table the reference holds for this path.

tests/golden/reference_testcases.json is a dump of /root/reference/tests/testcases/** (56 suites,
342 cases: output signals and accept/reject), produced by tools/gen_golden.py.  This is the same
check tests/test.py:57-74 performs against the circom-generated calculator.
"""
import json, os
import pytest
from oracle import oracle

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_testcases.json")))


@pytest.mark.parametrize("suite", GOLD, ids=[s["suite"] for s in GOLD])
def test_reference_suite(suite):
    for k, case in enumerate(suite["cases"]):
        w = oracle.run(suite["main"], case["input"])
        try:
            if case["expected"] is None:
                assert not w.ok, "%s case %d must be rejected" % (suite["suite"], k)
            else:
                assert w.ok, "%s case %d rejected (status %d)" % (suite["suite"], k, w.status)
                assert w.outputs() == [int(e) for e in case["expected"]], "%s case %d" % (suite["suite"], k)
        finally:
            w.free()


# Witness sizes derived in SURVEY.md Appendix D from the circom --O0 numbering rules (signals + 1).
SIZES = {
    "Spend(31)": 2603360, "KeccakBytes(1)": 2580773, "KeccakBytes(2)": 5130757, "PublicCommitment(6)": 5136782,
    "ProofOfWorkChecker()": 2593232, "BurnAddressHash()": 2586877, "BurnAddress()": 5054,
    "LeafDetector(544)": 20516, "RlpEmptyAccount(31)": 15166, "Num2BigEndianBytes(32)": 3779,
    "Poseidon(2)": 768, "Poseidon(3)": 935, "Poseidon(4)": 1168,
    "ProofOfBurn(4, 4, 5, 20, 31, 2, 10 ** 18, 10 ** 19)": 64355038,
}


@pytest.mark.parametrize("main", sorted(SIZES))
def test_witness_sizes(main):
    suite = next(s for s in GOLD if s["main"] == main)
    w = oracle.run(main, suite["cases"][0]["input"])
    try:
        assert w.n_signals == SIZES[main]
        assert w.value(0) == 1
    finally:
        w.free()


def test_known_answers():
    # circomlib/test/poseidoncircuit.js:52,62
    w = oracle.run("Poseidon(2)", {"inputs": [1, 2]})
    assert w.outputs() == [7853200120776062878684798364095072458815029376092732009249414926327459813530]
    w = oracle.run("Poseidon(2)", {"inputs": [3, 4]})
    assert w.outputs() == [14763215145315200506921711489642608356394854266165572616578112107564877678998]
    # keccak("") literal, circuits/utils/rlp/empty_account.circom:10
    w = oracle.run("KeccakBytes(1)", {"in": [0] * 136, "inLen": 0})
    assert bytes(w.outputs()).hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    # keccak(0x80) literal, empty_account.circom:9
    w = oracle.run("KeccakBytes(1)", {"in": [0x80] + [0] * 135, "inLen": 1})
    assert bytes(w.outputs()).hex() == "56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421"


def test_ordering_policy_switch_only_moves_two_sites():
    """H-create vs H-complete must not change size, outputs or value multiset."""
    inp = next(s for s in GOLD if s["main"] == "Spend(31)")["cases"][0]["input"]
    a, b = oracle.run("Spend(31)", inp, hcreate=False), oracle.run("Spend(31)", inp, hcreate=True)
    assert a.n_signals == b.n_signals and a.outputs() == b.outputs() and a.ok and b.ok
    diff = (a.limbs != b.limbs).any(axis=1).sum()
    assert 0 < diff < 4 * 1283   # only inside the four Num2Bits_strict blocks
