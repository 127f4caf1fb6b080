"""GPU parity tests proper: the CUDA path, called through the C-ABI (ctypes), against the CPU oracle on the
same inputs.  Bit-exact (integer field work).  Run with `pytest -m gpu` on a B200."""
import numpy as np
import pytest

from helpers import gold, suite, pob_fixture, repad_pob

pytestmark = pytest.mark.gpu

FULL_COMPARE_LIMIT = 6_000_000     # entries; larger witnesses are compared by digest + sampled windows


def _check_suite(s, full_limit=FULL_COMPARE_LIMIT):
    import pob_b200
    from oracle import oracle
    c = pob_b200.Circuit(s["main"], max_slots=max(2, len(s["cases"])))
    try:
        res = c.run([k["input"] for k in s["cases"]], expand=True, digest=True)
        assert c.desc["n_slots"] >= len(s["cases"])
        for i, case in enumerate(s["cases"]):
            w = oracle.run(s["main"], case["input"])
            try:
                assert c.n_signals == w.n_signals
                assert int(res.status[i]) == w.status, "%s case %d: status %d vs oracle %d" % (s["suite"], i, res.status[i], w.status)
                if case["expected"] is None:
                    assert res.status[i] != 0
                    continue
                assert res.status[i] == 0
                assert res.outputs[i] == [int(e) for e in case["expected"]], "%s case %d outputs" % (s["suite"], i)
                assert int(res.digests[i]) == w.digest(), "%s case %d digest" % (s["suite"], i)
                if w.n_signals <= full_limit:
                    gw = c.witness(i)
                    neq = np.nonzero((gw != w.limbs).any(axis=1))[0]
                    assert len(neq) == 0, "%s case %d: %d entries differ, first at %d" % (s["suite"], i, len(neq), neq[0])
            finally:
                w.free()
    finally:
        c.close()


SMALL = [s for s in gold() if s["suite"] != "test_proof_of_burn"]


@pytest.mark.parametrize("s", SMALL, ids=[s["suite"] for s in SMALL])
def test_gadget_suite(s):
    """all 55 gadget suites of the reference (tests/test.py:146-201) incl. Spend(31), KeccakBytes, Poseidon, RLP/MPT"""
    _check_suite(s)


def test_smoke_entry():
    import __graft_entry__
    __graft_entry__.smoke()


def test_proof_of_burn_test_shape():
    """ProofOfBurn(4,4,5,...) on tests/test_pob_input.json and its four corruptions
    (tests/testcases/proof_of_burn.py:52-76): status, commitment, digest of all 64.4 M entries, and the full
    witness of case 0 entry by entry."""
    import pob_b200
    from oracle import oracle
    s = suite("test_proof_of_burn")
    c = pob_b200.Circuit(s["main"], max_slots=5)
    try:
        res = c.run([k["input"] for k in s["cases"]], expand=True, digest=True)
        for i, case in enumerate(s["cases"]):
            w = oracle.run(s["main"], case["input"])
            try:
                assert int(res.status[i]) == w.status
                if case["expected"] is None:
                    assert res.status[i] != 0
                    continue
                assert res.outputs[i] == [int(e) for e in case["expected"]]
                assert int(res.digests[i]) == w.digest()
                if i == 0:
                    step = 1 << 22
                    for first in range(0, w.n_signals, step):
                        cnt = min(step, w.n_signals - first)
                        assert np.array_equal(c.witness(0, first, cnt), w.limbs[first:first + cnt]), "window at %d" % first
            finally:
                w.free()
    finally:
        c.close()


def test_main_proof_of_burn_shape():
    """BASELINE.json configs[1]: main_proof_of_burn = ProofOfBurn(16,4,16,50,31,2,10^19,10^20), reference fixture
    re-padded to the main shape; 215,907,954 entries (6.9 GB).  Commitment (padding-independent, so equal to the
    pinned (4,4,5) value), whole-witness digest and sampled windows vs the oracle; plus a rejected instance."""
    import pob_b200
    from oracle import oracle
    inp = repad_pob(pob_fixture(), 16, 4, 16)
    bad = repad_pob(pob_fixture(), 16, 4, 16)
    bad["layers"][1][0] = str(int(bad["layers"][1][0]) + 1)
    expected = [int(e) for e in suite("test_proof_of_burn")["cases"][0]["expected"]]
    c = pob_b200.Circuit(pob_b200.MAIN_PROOF_OF_BURN, max_slots=2)
    try:
        assert c.n_signals == 215907954
        res = c.run([inp, bad], expand=True, digest=True)
        w = oracle.run(pob_b200.MAIN_PROOF_OF_BURN, inp)
        try:
            assert w.ok and w.n_signals == c.n_signals and w.outputs() == expected
            assert res.status[0] == 0 and res.outputs[0] == expected
            assert int(res.digests[0]) == w.digest()
            rng = np.random.default_rng(7503)
            for first in [0, c.n_signals - 4096] + [int(v) for v in rng.integers(0, c.n_signals - 65536, 24)]:
                cnt = min(65536, c.n_signals - first)
                assert np.array_equal(c.witness(0, first, cnt), w.limbs[first:first + cnt]), "window at %d" % first
        finally:
            w.free()
        wb = oracle.run(pob_b200.MAIN_PROOF_OF_BURN, bad)
        try:
            assert not wb.ok and int(res.status[1]) == wb.status
        finally:
            wb.free()
    finally:
        c.close()


def test_wtns_file_roundtrip(tmp_path):
    """The exported .wtns is byte-identical to the oracle's (SURVEY.md Appendix B layout)."""
    import pob_b200
    from oracle import oracle
    s = suite("test_spend")
    c = pob_b200.Circuit("Spend(31)", max_slots=1)
    try:
        res = c.run([s["cases"][0]["input"]])
        assert res.status[0] == 0
        a, b = str(tmp_path / "gpu.wtns"), str(tmp_path / "oracle.wtns")
        c.write_wtns(0, a)
        w = oracle.run("Spend(31)", s["cases"][0]["input"])
        w.write_wtns(b)
        w.free()
        da, db = open(a, "rb").read(), open(b, "rb").read()
        assert len(da) == 83307596 and da == db
        assert da[:4] == b"wtns" and int.from_bytes(da[4:8], "little") == 2
    finally:
        c.close()
