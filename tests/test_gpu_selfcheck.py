"""pob_selfcheck on the GPU: every constraint of the circuit evaluated against a materialised witness (SURVEY.md 8(f)
rank 4), with randomised fault injection."""
import numpy as np
import pytest

from helpers import suite, cuda_poke

pytestmark = pytest.mark.gpu


def test_selfcheck_passes_and_reads_every_entry():
    import pob_b200
    for main, sname, k in (("Spend(31)", "test_spend", 0), ("KeccakBytes(2)", "test_keccak_2", 3), ("LeafDetector(544)", "test_leaf_detector_2", 0)):
        s = suite(sname)
        c = pob_b200.Circuit(main, max_slots=1)
        try:
            res = c.run([s["cases"][k]["input"]])
            assert res.status[0] == 0
            r = c.selfcheck(0)
            assert r["n_failed"] == 0 and r["n_hint_failed"] == 0 and r["first_failed"] == 2 ** 64 - 1
            assert r["signals_read"] == c.n_signals
            assert r == dict(pob_b200.constraint_info(main), ms=r["ms"])
        finally:
            c.close()


def test_ten_thousand_random_pokes_are_all_noticed():
    """Spend(31): 2.6 M entries -- one Keccakf (24 round blocks), two Poseidons, range checks, byte decompositions, selectors.
    10,000 random entries, one at a time, get +1 (mod p): the check must fail every time and pass again after the restore."""
    import pob_b200
    s = suite("test_spend")
    c = pob_b200.Circuit("Spend(31)", max_slots=1)
    try:
        res = c.run([s["cases"][0]["input"]])
        assert res.status[0] == 0
        w = c.witness(0)
        dptr = c.witness_device_ptr(0)
        rng = np.random.default_rng(2024)
        missed = []
        for i in rng.choice(np.arange(1, c.n_signals), size=10000, replace=False):
            old = pob_b200.from_limbs(w[i])
            cuda_poke(dptr, i, (old + 1) % pob_b200.P)
            r = c.selfcheck(0)
            cuda_poke(dptr, i, old)
            if r["n_failed"] + r["n_hint_failed"] == 0:
                missed.append(int(i))
        assert not missed, "entries whose change no constraint notices: %s" % missed[:20]
        r = c.selfcheck(0)
        assert r["n_failed"] == 0 and r["n_hint_failed"] == 0 and np.array_equal(c.witness(0), w)
    finally:
        c.close()


def test_selfcheck_main_shape():
    """main_proof_of_burn: all 215,962,293 constraints hold on a synthetic instance; 150 random pokes anywhere in the
    6.9 GB witness are noticed, and the id of the first failing record is stable for a given poke."""
    import pob_b200
    from pob_b200 import synth
    shape = (16, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)
    insts = synth.make_batch(1, shape, seed=31337)
    c = pob_b200.Circuit(pob_b200.MAIN_PROOF_OF_BURN, max_slots=1)
    try:
        res = c.run_packed(synth.pack_instances(insts, shape))
        assert res.status[0] == 0
        r = c.selfcheck(0)
        assert (r["n_constraints"], r["n_nonlinear"], r["n_hints"], r["signals_read"]) == (215962293, 17910859, 256010, 215907954)
        assert r["n_failed"] == 0 and r["n_hint_failed"] == 0
        dptr = c.witness_device_ptr(0)
        rng = np.random.default_rng(7)
        missed = []
        for i in [int(v) for v in rng.integers(1, c.n_signals, 150)]:
            old = pob_b200.from_limbs(c.witness(0, i, 1)[0])
            cuda_poke(dptr, i, (old + 1) % pob_b200.P)
            r1 = c.selfcheck(0)
            cuda_poke(dptr, i, old)
            if r1["n_failed"] + r1["n_hint_failed"] == 0:
                missed.append(i)
        assert not missed, missed[:20]
        assert c.selfcheck(0)["n_failed"] == 0
        assert c.selfcheck_keccak(0) == (2016, 0)
    finally:
        c.close()


def test_reduced_witness_on_gpu_equals_o0_through_the_map(tmp_path):
    """POB_CREATE_O1: the reduced witness of Spend(31) and of main_proof_of_burn, entry by entry, equals
    the oracle's --O0 witness through pob_witness_map; outputs and rejection are unchanged; the .wtns header carries the
    reduced count."""
    import pob_b200
    from pob_b200 import synth
    from oracle import oracle
    s = suite("test_spend")
    c = pob_b200.Circuit("Spend(31)", max_slots=2, opt=1)
    try:
        assert (c.n_signals, c.desc["n_signals_o0"], c.desc["opt_level"]) == (259945, 2603360, 1)
        m = c.witness_map()
        res = c.run([s["cases"][0]["input"], s["cases"][1]["input"]])
        assert res.status[0] == 0 and res.status[1] != 0
        w = oracle.run("Spend(31)", s["cases"][0]["input"])
        assert res.outputs[0] == w.outputs() and np.array_equal(c.witness(0), w.limbs[m])
        w.free()
        f = str(tmp_path / "reduced.wtns")
        c.write_wtns(0, f)
        raw = open(f, "rb").read()
        assert len(raw) == 76 + 32 * 259945 and int.from_bytes(raw[60:64], "little") == 259945
        with pytest.raises(pob_b200.PobError):
            c.selfcheck(0)                                   # the constraint system is stated over the --O0 layout
    finally:
        c.close()
    shape = (16, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)
    insts = synth.make_batch(40, shape, seed=808)
    packed = synth.pack_instances(insts, shape)
    c = pob_b200.Circuit(pob_b200.MAIN_PROOF_OF_BURN, opt=1, max_slots=40)
    try:
        assert (c.n_signals, c.desc["n_signals_o0"]) == (21454051, 215907954) and c.desc["n_slots"] == 40
        m = c.witness_map()
        res = c.run_packed(packed)
        assert (res.status == 0).all()
        for i in (0, 39):
            w = oracle.run_flat(*oracle.parse_main(pob_b200.MAIN_PROOF_OF_BURN), packed[i])
            try:
                assert w.ok and res.outputs[i] == w.outputs()
                assert np.array_equal(c.witness(i), w.limbs[m]), "instance %d" % i
            finally:
                w.free()
    finally:
        c.close()
