#!/usr/bin/env python3
"""Order-pinning kit: settle with ONE command whether pob_b200's witness numbering equals circom's.

The reference pins no witness ORDER (it ships no .sym / .wtns / .r1cs; its calculator comes from the external circom compiler,
which is not available where this repo was built).  Anyone with circom >= 2.1 can close that gap:

    circom circuits/main_spend.circom --O0 --sym -o /tmp/out                 # in the reference checkout
    python tools/diff_sym.py /tmp/out/main_spend.sym main_spend              # -> "IDENTICAL" or the first divergence
    python tools/diff_sym.py --auto /path/to/proof-of-burn                   # runs circom itself when it is on PATH

What is compared: circom's `.sym` lists every signal as `#s,#w,#c,name` in numbering order; the own signals of one component
instance are contiguous, so the file collapses to a sequence of (first signal, number of own signals, component path).
pob_b200 writes the same sequence for its layout (pob_write_components: first signal, own signals, template name).  The two
sequences must agree element by element.  On a mismatch the tool prints both sides around the first divergence and repeats
the comparison with the other sub-component numbering policy (`hcreate`, SURVEY.md Appendix C rule R3: creation order instead
of completion order; the two sites where it matters are Num2Bits_strict and MultiAND(n >= 3)) and says which one fits.
"""
import argparse, os, shutil, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proof-of-burn_b200"))


def sym_components(path):
    """real circom .sym -> [(first_signal, n_own, component_path)]"""
    comps, cur, first, n = [], None, 0, 0
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            s, _w, _c, name = line.split(",", 3)
            comp = name.rsplit(".", 1)[0] if "." in name else ""
            if comp != cur:
                if cur is not None:
                    comps.append((first, n, cur))
                cur, first, n = comp, int(s), 0
            n += 1
    if cur is not None:
        comps.append((first, n, cur))
    return comps


def our_components(main_expr, hcreate, keep=None):
    import pob_b200
    path = keep or tempfile.mktemp(suffix=".comps")
    pob_b200.write_components(pob_b200.CIRCUIT_ALIASES.get(main_expr, main_expr), path, hcreate=hcreate)
    out = []
    with open(path) as f:
        for line in f:
            if line.startswith("#"):
                continue
            a, b, t = line.rstrip("\n").split(",", 2)
            out.append((int(a), int(b), t))
    if not keep:
        os.remove(path)
    return out


def first_divergence(real, ours):
    for i, (r, o) in enumerate(zip(real, ours)):
        if r[0] != o[0] or r[1] != o[1]:
            return i
    return None if len(real) == len(ours) else min(len(real), len(ours))


def show(real, ours, i, out):
    for k in range(max(0, i - 4), min(max(len(real), len(ours)), i + 5)):
        r = real[k] if k < len(real) else ("-", "-", "-")
        o = ours[k] if k < len(ours) else ("-", "-", "-")
        print("%s #%-8d circom: first %-10s own %-6s %-60s | pob_b200: first %-10s own %-6s %s" % (">>" if k == i else "  ", k, r[0], r[1], str(r[2])[-60:], o[0], o[1], o[2]), file=out)


def compare(sym_path, main_expr, out=sys.stdout):
    real = sym_components(sym_path)
    res = {}
    for hc in (False, True):
        ours = our_components(main_expr, hc)
        res[hc] = (first_divergence(real, ours), ours)
    d0, d1 = res[False][0], res[True][0]
    n_sig = real[-1][0] + real[-1][1] - 1 if real else 0
    print("circom .sym: %d components, %d signals; pob_b200: %d components (default policy), %d (hcreate)" % (len(real), n_sig, len(res[False][1]), len(res[True][1])), file=out)
    if d0 is None:
        print("IDENTICAL component structure with the default policy (completion order): the witness ORDER is pinned.", file=out)
        return 0
    if d1 is None:
        print("IDENTICAL component structure with hcreate=1 (creation order); the default policy diverges at component #%d:" % d0, file=out)
        show(real, res[False][1], d0, out)
        print("=> create the circuit with hcreate=1 (POB_CREATE_HCREATE) to match this circom build.", file=out)
        return 1
    best = False if d0 >= d1 else True
    print("DIVERGES under both policies: default at component #%d, hcreate at #%d.  Closest policy: %s" % (d0, d1, "hcreate" if best else "default"), file=out)
    show(real, res[best][1], res[best][0], out)
    return 2


def auto(ref, mains, out=sys.stdout):
    circom = shutil.which("circom")
    if not circom:
        print("circom is not on PATH: the witness ORDER stays unpinned (parity of values, outputs and accept/reject is pinned by the test-suite).", file=out)
        return 3
    rc = 0
    for m in mains:
        src = os.path.join(ref, "circuits", m + ".circom")
        tmp = tempfile.mkdtemp()
        subprocess.check_call([circom, src, "--O0", "--sym", "-o", tmp])
        rc |= compare(os.path.join(tmp, m + ".sym"), m, out)
        shutil.rmtree(tmp, ignore_errors=True)
    return rc


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("sym", nargs="?", help="circom .sym file (or, with --auto, the reference checkout)")
    ap.add_argument("main", nargs="?", default="main_spend", help="main_spend | main_proof_of_burn | 'Template(params)'")
    ap.add_argument("--auto", action="store_true", help="run circom --O0 --sym on <sym>/circuits/<main>.circom first")
    a = ap.parse_args()
    if a.auto:
        return auto(a.sym or "/root/reference", [a.main])
    if not a.sym:
        ap.error("a .sym file is required")
    return compare(a.sym, a.main)


if __name__ == "__main__":
    sys.exit(main())
