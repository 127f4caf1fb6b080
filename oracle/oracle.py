"""ctypes front-end of the CPU oracle (oracle/pob_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (proof-of-burn_b200/) never imports this module.

`run(main_expr, input_dict)` mirrors what the reference harness does for one case
(tests/test.py:57-74): build `component main = <main_expr>`, feed one input JSON, get the output
signals (witness[1..n_out]) or a failure.
"""
import ctypes, json, os, re, subprocess
import numpy as np

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libpob_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("pob_oracle.c", "fr.h", "poseidon_constants_data.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libpob_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.pob_oracle_schema.restype = ctypes.c_char_p
        L.pob_oracle_schema.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
        L.pob_oracle_run.restype = ctypes.c_int
        L.pob_oracle_run.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                     ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint64),
                                     ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint64)]
        L.pob_oracle_free.argtypes = [ctypes.c_void_p]
        L.pob_oracle_write_wtns.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint64]
        L.pob_oracle_digest.restype = ctypes.c_uint64
        L.pob_oracle_digest.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
        _LIB = L
    return _LIB


def _int_expr(text, env=None):
    """integer expressions of template parameters / schema dims (`10 ** 19`, `p1*136`): whitelisted AST walk"""
    import ast, operator
    ops = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.FloorDiv: operator.floordiv, ast.Pow: operator.pow}

    def ev(n):
        if isinstance(n, ast.Expression):
            return ev(n.body)
        if isinstance(n, ast.Constant) and type(n.value) is int:
            return n.value
        if isinstance(n, ast.Name) and env and n.id in env:
            return int(env[n.id])
        if isinstance(n, ast.UnaryOp) and isinstance(n.op, ast.USub):
            return -ev(n.operand)
        if isinstance(n, ast.BinOp) and type(n.op) in ops:
            a, b = ev(n.left), ev(n.right)
            if isinstance(n.op, ast.Pow) and not 0 <= b <= 4096:
                raise ValueError(text)
            return ops[type(n.op)](a, b)
        raise ValueError("unsupported expression %r" % text)
    return int(ev(ast.parse(text.strip(), mode="eval")))


def parse_main(expr):
    """'ProofOfBurn(4, 4, 5, 20, 31, 2, 10 ** 18, 10 ** 19)' -> ('ProofOfBurn', [4,4,5,20,31,2,10**18,10**19])"""
    m = re.match(r"\s*(\w+)\s*\((.*)\)\s*$", expr, re.S)
    name, args = m.group(1), m.group(2).strip()
    params = [_int_expr(a) for a in args.split(",")] if args else []
    return name, params


def to_limbs(vals):
    """list of python ints (any sign / size) -> (n,4) uint64 canonical limbs"""
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        v = int(v) % P
        if v < (1 << 64):
            out[i, 0] = v
        else:
            for k in range(4):
                out[i, k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def from_limbs(row):
    return int(row[0]) | (int(row[1]) << 64) | (int(row[2]) << 128) | (int(row[3]) << 192)


def schema(name, params):
    """[(input name, [dims...])] in declaration order"""
    n = ctypes.c_int(0)
    s = lib().pob_oracle_schema(name.encode(), ctypes.byref(n))
    if s is None:
        raise KeyError("oracle has no main template %r" % name)
    env = {"p%d" % i: v for i, v in enumerate(params)}
    out = []
    for item in [x for x in s.decode().split(",") if x]:
        m = re.match(r"(\w+)((?:\[[^\]]+\])*)$", item)
        dims = [_int_expr(d, env) for d in re.findall(r"\[([^\]]+)\]", m.group(2))]
        out.append((m.group(1), dims))
    return out


def _flatten(v):
    if isinstance(v, (list, tuple)):
        r = []
        for e in v:
            r.extend(_flatten(e))
        return r
    return [int(v)]


def flatten_inputs(sch, inp):
    """JSON dict -> flat list of ints.  Accepts ints or decimal strings, nested arrays, and scalars
    given as 1-element arrays (tests/testcases/divide.py:4) -- the leniency of the circom loader."""
    flat = []
    for name, dims in sch:
        vals = _flatten(inp[name])
        want = int(np.prod(dims)) if dims else 1
        if len(vals) != want:
            raise ValueError("input %s: expected %d values, got %d" % (name, want, len(vals)))
        flat.extend(vals)
    return flat


class Witness:
    def __init__(self, ptr, n_signals, n_outputs, status):
        self._ptr, self.n_signals, self.n_outputs, self.status = ptr, n_signals, n_outputs, status
        buf = (ctypes.c_uint64 * (4 * n_signals)).from_address(ptr)
        self.limbs = np.frombuffer(buf, dtype=np.uint64).reshape(n_signals, 4)

    @property
    def ok(self):
        return self.status == 0

    def outputs(self):
        return [from_limbs(self.limbs[1 + i]) for i in range(self.n_outputs)]

    def value(self, i):
        return from_limbs(self.limbs[i])

    def digest(self):
        return int(lib().pob_oracle_digest(self._ptr, self.n_signals))

    def write_wtns(self, path):
        if lib().pob_oracle_write_wtns(path.encode(), self._ptr, self.n_signals) != 0:
            raise IOError(path)

    def free(self):
        if self._ptr:
            self.limbs = None
            lib().pob_oracle_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def run_flat(name, params, flat_inputs, hcreate=False):
    pl = to_limbs(params) if params else np.zeros((1, 4), dtype=np.uint64)
    il = flat_inputs if isinstance(flat_inputs, np.ndarray) else to_limbs(flat_inputs)
    il = np.ascontiguousarray(il, dtype=np.uint64)
    if il.size == 0:
        il = np.zeros((1, 4), dtype=np.uint64)
    w, n, no, st = ctypes.c_void_p(), ctypes.c_uint64(), ctypes.c_uint32(), ctypes.c_uint64()
    rc = lib().pob_oracle_run(name.encode(), pl.ctypes.data, len(params), il.ctypes.data, il.shape[0], int(hcreate),
                              ctypes.byref(w), ctypes.byref(n), ctypes.byref(no), ctypes.byref(st))
    if rc != 0:
        raise RuntimeError("pob_oracle_run(%s) failed: %d" % (name, rc))
    return Witness(w.value, n.value, no.value, st.value)


def run(main_expr, inp, hcreate=False):
    name, params = parse_main(main_expr)
    return run_flat(name, params, flatten_inputs(schema(name, params), inp), hcreate)


if __name__ == "__main__":   # python -m oracle.oracle "Spend(31)" input.json witness.wtns
    import sys
    w = run(sys.argv[1], json.load(open(sys.argv[2])))
    print("signals", w.n_signals, "status", w.status, "outputs", w.outputs()[:4])
    if len(sys.argv) > 3 and w.ok:
        w.write_wtns(sys.argv[3])
    sys.exit(0 if w.ok else 1)
