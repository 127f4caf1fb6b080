"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/pob_b200.h
declares, its host-only layout compiler reports the witness shapes derived in SURVEY.md Appendix D, and the product
path fails loudly (no CPU fallback) when there is no GPU.  No compute calls."""
import ctypes, os, re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pob_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pob_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import pob_b200
    L = ctypes.CDLL(pob_b200.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "libpob_b200.so does not export %s" % n
    assert b"sm_100a" in pob_b200.lib().pob_version()


SHAPES = {   # main expression -> (n_signals, n_inputs, Keccak-f permutations)   SURVEY.md Appendix D / BASELINE.md section 2
    "Spend(31)": (2603360, 4, 1),
    "ProofOfBurn(4, 4, 5, 20, 31, 2, 10 ** 18, 10 ** 19)": (64355038, 4 * 544 + 4 + 1 + 680 + 9, 25),
    "ProofOfBurn(16, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)": (215907954, 10906, 84),
    "ProofOfBurn(8, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)": (51277058 + 10289431 * 8, None, 20 + 4 * 8),
    "KeccakBytes(1)": (2580773, 137, 1),
}


@pytest.mark.parametrize("main", sorted(SHAPES))
def test_layout_info_shapes(main):
    import pob_b200
    n_sig, n_in, n_perm = SHAPES[main]
    d = pob_b200.layout_info(main)
    assert d["n_signals"] == n_sig
    assert d["witness_bytes"] == 32 * n_sig and d["wtns_file_bytes"] == 76 + 32 * n_sig
    assert d["n_absorbs"] == n_perm
    if n_in is not None:
        assert d["n_inputs"] == n_in
    assert d["n_tiles"] >= n_sig // 8192
    assert d["store_bytes"] < 128 << 20


def test_ordering_policy_does_not_change_shape():
    import pob_b200
    a, b = pob_b200.layout_info("Spend(31)", hcreate=False), pob_b200.layout_info("Spend(31)", hcreate=True)
    assert a["n_signals"] == b["n_signals"] and a["n_ops"] == b["n_ops"]


def test_unknown_template_and_bad_shape_are_errors():
    import pob_b200
    with pytest.raises(pob_b200.PobError):
        pob_b200.layout_info("NoSuchTemplate(3)")
    with pytest.raises(pob_b200.PobError):
        pob_b200.layout_info("ProofOfBurn(4, 4)")


def test_no_cpu_fallback():
    """Without a CUDA device the product must refuse to run rather than compute on the host."""
    import torch
    import pob_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pob_b200.PobError) as e:
        pob_b200.Circuit("Spend(31)")
    assert e.value.code == -2 and "no CPU path" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """Nothing under proof-of-burn_b200/ may import, link or execute oracle/ (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "proof-of-burn_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", "Makefile")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "libpob_oracle" not in src and "pob_oracle_run" not in src, f


def test_input_schema_and_flatten():
    import pob_b200
    sch = pob_b200.input_schema("ProofOfBurn", [4, 4, 5, 20, 31, 2, 10 ** 18, 10 ** 19])
    assert [n for n, _ in sch] == ["burnKey", "actualBalance", "intendedBalance", "revealAmount", "burnExtraCommitment",
                                   "numLeafAddressNibbles", "layers", "layerLens", "numLayers", "blockHeader", "blockHeaderLen",
                                   "byteSecurityRelax", "_proofExtraCommitment"]     # circuits/proof_of_burn.circom:43-72
    assert dict(sch)["layers"] == [4, 544] and dict(sch)["blockHeader"] == [680]
    s2 = pob_b200.input_schema("Divide", [16])
    # scalars may arrive as 1-element arrays (reference tests/testcases/divide.py:4); strings and ints mix
    assert pob_b200.flatten_input(s2, {"a": [10], "b": "3"}) == [10, 3]
    with pytest.raises(ValueError):
        pob_b200.flatten_input(sch, {k: 0 for k, _ in sch})
    assert pob_b200.parse_main("component main = 0;".replace("component main = 0", "Spend(31)")) == ("Spend", [31])
    lim = pob_b200.to_limbs([-1, pob_b200.P + 5, 2 ** 200])
    assert pob_b200.from_limbs(lim[0]) == pob_b200.P - 1 and pob_b200.from_limbs(lim[1]) == 5 and pob_b200.from_limbs(lim[2]) == 2 ** 200


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    """The boundary is a C ABI: include/pob_b200.h must compile as C99 (no C++, no CUDA, no torch types) and a plain C host
    must link against the shared library and reach its host-only entry points (no GPU needed for these)."""
    import subprocess
    import pob_b200
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "pob_b200.h"
int main(void) {
    uint64_t p[1][4] = {{31, 0, 0, 0}};
    pob_desc d; pob_check_report r; int np = -1;
    if (pob_layout_info("Spend", &p[0][0], 1, 0, &d) != POB_OK) { printf("layout: %s\n", pob_last_error()); return 1; }
    if (pob_layout_info("Spend", &p[0][0], 1, POB_CREATE_O1, &d) != POB_OK || d.opt_level != 1 || d.n_signals_o0 != 2603360) return 2;
    if (pob_constraint_info("Spend", &p[0][0], 1, 0, &r) != POB_OK || r.signals_read != 2603360) return 3;
    if (!pob_input_schema("Spend", &np) || np != 1) return 4;
    if (pob_layout_info("NoSuch", 0, 0, 0, &d) != POB_E_COMPILE || strlen(pob_last_error()) == 0) return 5;
    printf("%llu %llu %s\n", (unsigned long long)d.n_signals, (unsigned long long)r.n_constraints, pob_version());
    return 0;
}
''')
    exe = str(tmp_path / "host")
    libdir = os.path.dirname(pob_b200.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe,
                           "-L", libdir, "-l:libpob_b200.so", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.split()[:2] == ["259945", "2605282"]
