"""pob_b200 -- Python host of the B200-native batched witness generator (ctypes over libpob_b200.so).

Mirrors the reference's interface for this path: the process CLI the circom toolchain emits,
`./<circuit> input.json witness.wtns` (reference Makefile:5-6, tests/test.py:60-63), with the input-JSON
schema tests/main.py:160-178 writes (keys = the main template's `signal input` names, values JSON ints or
decimal strings, nested arrays, scalars possibly wrapped in 1-element arrays).

    from pob_b200 import Circuit
    c = Circuit("ProofOfBurn(16, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)")       # == circom -c + make
    res = c.run([input_dict, ...])                                              # == N x ./main input.json w.wtns
    res.status[i] == 0, res.outputs[i] -> [commitment]; c.write_wtns(i, "witness.wtns")

All witness computation happens in hand-written CUDA (csrc/pob_b200.cu).  There is no CPU fallback: if the
extension is missing or no GPU is visible, construction fails loudly.
"""
import ctypes
import json
import os
import re

import numpy as np

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpob_b200.so")

RUN_EXPAND, RUN_DIGEST, RUN_INPUTS_STAGED, RUN_DISCARD = 1, 2, 4, 8
CREATE_HCREATE, CREATE_O1 = 1, 0x100
E_RANGE, E_REJECTED, E_BUSY, DONE = -5, -8, -9, 1
MAIN_PROOF_OF_BURN = "ProofOfBurn(16, 4, 16, 50, 31, 2, 10 ** 19, 10 ** 20)"   # circuits/main_proof_of_burn.circom:27
MAIN_SPEND = "Spend(31)"                                                       # circuits/main_spend.circom:6
TEST_PROOF_OF_BURN = "ProofOfBurn(4, 4, 5, 20, 31, 2, 10 ** 18, 10 ** 19)"     # tests/testcases/proof_of_burn.py:53


class PobError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pob_b200 error %d: %s" % (code, msg))
        self.code = code


class Desc(ctypes.Structure):
    _fields_ = [("n_signals", ctypes.c_uint64), ("n_outputs", ctypes.c_uint32), ("n_inputs", ctypes.c_uint32),
                ("witness_bytes", ctypes.c_uint64), ("wtns_file_bytes", ctypes.c_uint64), ("store_bytes", ctypes.c_uint64),
                ("n_ops", ctypes.c_uint64), ("n_absorbs", ctypes.c_uint32), ("n_levels", ctypes.c_uint32),
                ("n_tiles", ctypes.c_uint32), ("n_slots", ctypes.c_uint32), ("chunk", ctypes.c_uint32), ("expand_group", ctypes.c_uint32),
                ("opt_level", ctypes.c_uint32), ("n_signals_o0", ctypes.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class Timing(ctypes.Structure):
    _fields_ = [("total_ms", ctypes.c_float), ("expand_ms", ctypes.c_float), ("eval_ms", ctypes.c_float),
                ("expand_launches", ctypes.c_uint32), ("eval_launches", ctypes.c_uint32), ("other_launches", ctypes.c_uint32),
                ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64)]

    def as_dict(self):
        return {k: (float(getattr(self, k)) if k.endswith("_ms") else int(getattr(self, k))) for k, _ in self._fields_}


class ExportStats(ctypes.Structure):
    _fields_ = [("witnesses", ctypes.c_uint64), ("bytes", ctypes.c_uint64), ("total_ms", ctypes.c_float), ("d2h_gbs", ctypes.c_float)]

    def as_dict(self):
        return {"witnesses": int(self.witnesses), "bytes": int(self.bytes), "total_ms": float(self.total_ms), "d2h_gbs": float(self.d2h_gbs)}


class CheckReport(ctypes.Structure):
    _fields_ = [("n_constraints", ctypes.c_uint64), ("n_nonlinear", ctypes.c_uint64), ("n_hints", ctypes.c_uint64), ("n_failed", ctypes.c_uint64),
                ("n_hint_failed", ctypes.c_uint64), ("first_failed", ctypes.c_uint64), ("signals_read", ctypes.c_uint64), ("ms", ctypes.c_float)]

    def as_dict(self):
        return {k: (float(getattr(self, k)) if k == "ms" else int(getattr(self, k))) for k, _ in self._fields_}


_LIB = None


def lib():
    """Load the CUDA extension.  Fails loudly when it has not been built (python __graft_entry__.py build)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("pob_b200: %s is missing -- build it with `make -C proof-of-burn_b200/csrc` "
                              "(there is no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        vp, u32, u64, ci = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
        L.pob_create.restype = ci
        L.pob_create.argtypes = [ctypes.c_char_p, vp, ci, ci, ci, u32, ctypes.POINTER(vp)]
        L.pob_destroy.argtypes = [vp]
        L.pob_layout_info.restype = ci
        L.pob_layout_info.argtypes = [ctypes.c_char_p, vp, ci, ci, ctypes.POINTER(Desc)]
        L.pob_input_schema.restype = ctypes.c_char_p
        L.pob_input_schema.argtypes = [ctypes.c_char_p, ctypes.POINTER(ci)]
        L.pob_describe.restype = ci
        L.pob_describe.argtypes = [vp, ctypes.POINTER(Desc)]
        L.pob_alloc_pinned.restype = vp
        L.pob_alloc_pinned.argtypes = [u64]
        L.pob_free_pinned.argtypes = [vp]
        L.pob_stage_inputs.restype = ci
        L.pob_stage_inputs.argtypes = [vp, vp, u32]
        L.pob_run_batch.restype = ci
        L.pob_run_batch.argtypes = [vp, vp, u32, u32, vp, vp, vp]
        L.pob_last_timing.restype = ci
        L.pob_last_timing.argtypes = [vp, ctypes.POINTER(Timing)]
        L.pob_copy_witness.restype = ci
        L.pob_copy_witness.argtypes = [vp, u32, u64, u64, vp]
        L.pob_write_wtns.restype = ci
        L.pob_write_wtns.argtypes = [vp, u32, ctypes.c_char_p]
        L.pob_witness_device_ptr.restype = ci
        L.pob_witness_device_ptr.argtypes = [vp, u32, ctypes.POINTER(vp)]
        L.pob_selfcheck_keccak.restype = ci
        L.pob_selfcheck_keccak.argtypes = [vp, u32, ctypes.POINTER(u64), ctypes.POINTER(u64)]
        L.pob_selfcheck.restype = ci
        L.pob_selfcheck.argtypes = [vp, u32, ctypes.POINTER(CheckReport)]
        L.pob_constraint_info.restype = ci
        L.pob_constraint_info.argtypes = [ctypes.c_char_p, vp, ci, ci, ctypes.POINTER(CheckReport)]
        L.pob_write_components.restype = ci
        L.pob_write_components.argtypes = [ctypes.c_char_p, vp, ci, ci, ctypes.c_char_p, ctypes.POINTER(u64)]
        L.pob_witness_map.restype = ci
        L.pob_witness_map.argtypes = [vp, vp]
        L.pob_run_batch_retain.restype = ci
        L.pob_run_batch_retain.argtypes = [vp, vp, u32, u32, vp, u32, vp, vp, vp]
        L.pob_submit.restype = ci
        L.pob_submit.argtypes = [vp, vp, u32, u32]
        L.pob_acquire.restype = ci
        L.pob_acquire.argtypes = [vp, ctypes.POINTER(u32), ctypes.POINTER(vp), vp]
        L.pob_release.restype = ci
        L.pob_release.argtypes = [vp, u32, vp]
        L.pob_finish.restype = ci
        L.pob_finish.argtypes = [vp, vp, vp, vp]
        L.pob_export_batch.restype = ci
        L.pob_export_batch.argtypes = [vp, vp, u32, u32, vp, vp, vp, ctypes.POINTER(ExportStats)]
        L.pob_pow_grind.restype = ci
        L.pob_pow_grind.argtypes = [ci, vp, vp, vp, u32, u64, vp, ctypes.POINTER(u64)]
        L.pob_last_error.restype = ctypes.c_char_p
        L.pob_version.restype = ctypes.c_char_p
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise PobError(rc, lib().pob_last_error().decode())


# ---- field / schema helpers --------------------------------------------------------------------------------
def eval_int_expr(text, env=None):
    """Template parameters and schema dimensions are tiny integer expressions (`10 ** 19`, `p1*136`, `2*p0`): evaluate
    them with a whitelisted AST walk -- integers, names from `env`, + - * // ** and unary minus; nothing else."""
    import ast
    import operator
    ops = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.FloorDiv: operator.floordiv, ast.Pow: operator.pow}

    def ev(n):
        if isinstance(n, ast.Expression):
            return ev(n.body)
        if isinstance(n, ast.Constant) and isinstance(n.value, int) and not isinstance(n.value, bool):
            return n.value
        if isinstance(n, ast.Name) and env is not None and n.id in env:
            return int(env[n.id])
        if isinstance(n, ast.UnaryOp) and isinstance(n.op, ast.USub):
            return -ev(n.operand)
        if isinstance(n, ast.BinOp) and type(n.op) in ops:
            a, b = ev(n.left), ev(n.right)
            if isinstance(n.op, ast.Pow) and (b < 0 or b > 4096 or abs(a) > 1 << 64):
                raise ValueError("exponent out of range in %r" % text)
            return ops[type(n.op)](a, b)
        raise ValueError("unsupported expression %r" % text)
    return int(ev(ast.parse(text.strip(), mode="eval")))


def parse_int(v):
    """one input value: JSON int, decimal string, or 0x-prefixed hex string (some circom loaders accept it)"""
    if isinstance(v, str):
        t = v.strip()
        return int(t, 16) if t.lower().startswith(("0x", "-0x")) else int(t)
    return int(v)


def parse_main(expr):
    """'ProofOfBurn(4, 4, 5, 20, 31, 2, 10 ** 18, 10 ** 19)' -> ('ProofOfBurn', [4, 4, 5, 20, 31, 2, 10**18, 10**19])"""
    m = re.match(r"\s*(\w+)\s*(?:\((.*)\))?\s*;?\s*$", expr, re.S)
    if not m:
        raise ValueError("cannot parse main expression %r" % expr)
    args = (m.group(2) or "").strip()
    return m.group(1), ([eval_int_expr(a) for a in args.split(",")] if args else [])


def to_limbs(vals):
    """ints of any sign/size (reduced mod p like the circom loader) -> (n, 4) uint64 little-endian limbs"""
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        v = int(v) % P
        if v >> 64:
            for k in range(4):
                out[i, k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
        else:
            out[i, 0] = v
    return out


def from_limbs(row):
    return int(row[0]) | (int(row[1]) << 64) | (int(row[2]) << 128) | (int(row[3]) << 192)


def input_schema(name, params):
    """[(input name, [dims])] in declaration order, from the library (single source of truth)."""
    n = ctypes.c_int(0)
    s = lib().pob_input_schema(name.encode(), ctypes.byref(n))
    if s is None:
        raise PobError(-1, "unknown main template %r" % name)
    if len(params) < n.value:
        raise PobError(-1, "%s needs %d template parameters" % (name, n.value))
    env = {"p%d" % i: v for i, v in enumerate(params)}
    out = []
    for item in [x for x in s.decode().split(",") if x]:
        m = re.match(r"(\w+)((?:\[[^\]]+\])*)$", item)
        out.append((m.group(1), [eval_int_expr(d, env) for d in re.findall(r"\[([^\]]+)\]", m.group(2))]))
    return out


def _flatten(v, out):
    if isinstance(v, (list, tuple)):
        for e in v:
            _flatten(e, out)
    else:
        out.append(parse_int(v))


def flatten_input(schema, inp):
    """One input JSON object -> flat list of ints in declaration order.  Element counts must match the circuit
    exactly, as with the circom loader; unknown keys are ignored, missing keys are an error."""
    flat = []
    for name, dims in schema:
        if name not in inp:
            raise KeyError("input signal %r missing" % name)
        vals = []
        _flatten(inp[name], vals)
        want = int(np.prod(dims)) if dims else 1
        if len(vals) != want:
            raise ValueError("input %s: circuit expects %d values, got %d" % (name, want, len(vals)))
        flat.extend(vals)
    return flat


def layout_info(main_expr, hcreate=False, opt=0):
    """Shape of a circuit's witness program; runs the host-side layout compiler only (no GPU needed)."""
    name, params = parse_main(main_expr)
    pl = to_limbs(params) if params else np.zeros((1, 4), dtype=np.uint64)
    d = Desc()
    _check(lib().pob_layout_info(name.encode(), pl.ctypes.data, len(params), (CREATE_HCREATE if hcreate else 0) | (CREATE_O1 if opt else 0), ctypes.byref(d)))
    return d.as_dict()


def constraint_info(main_expr, hcreate=False):
    """size of a circuit shape's constraint system (host only): constraints, non-linear ones, hint records, signals covered"""
    name, params = parse_main(main_expr)
    pl = to_limbs(params) if params else np.zeros((1, 4), dtype=np.uint64)
    r = CheckReport()
    _check(lib().pob_constraint_info(name.encode(), pl.ctypes.data, len(params), int(hcreate), ctypes.byref(r)))
    return r.as_dict()


def write_components(main_expr, path, hcreate=False):
    """order-pinning kit: component list (`first_signal,n_own_signals,template` per line) of the --O0 layout; see tools/diff_sym.py"""
    name, params = parse_main(main_expr)
    pl = to_limbs(params) if params else np.zeros((1, 4), dtype=np.uint64)
    n = ctypes.c_uint64(0)
    _check(lib().pob_write_components(name.encode(), pl.ctypes.data, len(params), CREATE_HCREATE if hcreate else 0, os.fsencode(path), ctypes.byref(n)))
    return int(n.value)


class PinnedArray:
    """numpy view over cudaMallocHost memory (so H2D copies inside run() are asynchronous DMA)."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = lib().pob_alloc_pinned(max(1, self.nbytes))
        if not self.ptr:
            raise PobError(-3, "pob_alloc_pinned failed")
        buf = (ctypes.c_uint8 * max(1, self.nbytes)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().pob_free_pinned(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class BatchResult:
    def __init__(self, status, outputs, digests, timing):
        self.status, self.outputs_limbs, self.digests, self.timing = status, outputs, digests, timing

    @property
    def n_ok(self):
        return int((self.status == 0).sum())

    @property
    def outputs(self):
        return [[from_limbs(r) for r in inst] for inst in self.outputs_limbs]


class Circuit:
    """One compiled circuit shape bound to one GPU (== the executable `circom -c ... && make` produces)."""

    def __init__(self, main_expr, device=0, hcreate=False, max_slots=0, opt=0):
        """opt=1: the reduced (`--O1`-style) witness (pob_b200.h: POB_CREATE_O1); witness_map() gives the --O0 index of each entry"""
        self.main_expr = main_expr
        self.name, self.params = parse_main(main_expr)
        self.schema = input_schema(self.name, self.params)
        pl = to_limbs(self.params) if self.params else np.zeros((1, 4), dtype=np.uint64)
        h = ctypes.c_void_p()
        _check(lib().pob_create(self.name.encode(), pl.ctypes.data, len(self.params), (CREATE_HCREATE if hcreate else 0) | (CREATE_O1 if opt else 0),
                                int(device), int(max_slots), ctypes.byref(h)))
        self._h = h
        d = Desc()
        _check(lib().pob_describe(self._h, ctypes.byref(d)))
        self.desc = d.as_dict()
        self.n_signals, self.n_inputs, self.n_outputs = self.desc["n_signals"], self.desc["n_inputs"], self.desc["n_outputs"]

    def close(self):
        if getattr(self, "_h", None):
            lib().pob_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- inputs ----
    def pack(self, inputs, pinned=False):
        """list of input JSON objects (or one) -> (n, n_inputs, 4) uint64 array"""
        if isinstance(inputs, dict):
            inputs = [inputs]
        n = len(inputs)
        shape = (n, max(1, self.n_inputs), 4)
        holder = PinnedArray(shape, np.uint64) if pinned else None
        arr = holder.array if pinned else np.zeros(shape, dtype=np.uint64)
        for i, inp in enumerate(inputs):
            arr[i, : self.n_inputs] = to_limbs(flatten_input(self.schema, inp))
        return (arr, holder) if pinned else arr

    def stage(self, packed):
        packed = np.ascontiguousarray(packed, dtype=np.uint64)
        _check(lib().pob_stage_inputs(self._h, packed.ctypes.data, packed.shape[0]))

    # ---- run ----
    def _inputs_ptr(self, packed, n, staged):
        if staged:
            assert n is not None
            return None, n, None
        packed = np.ascontiguousarray(packed, dtype=np.uint64)
        return packed.ctypes.data, (packed.shape[0] if n is None else n), packed

    def run_packed(self, packed, n=None, expand=True, digest=False, staged=False, discard=False, retain=None):
        """One synchronous batch (pob_run_batch / pob_run_batch_retain).  With expand and n > n_slots the library wants
        to know what happens to the witnesses: digest=True (each one is consumed by the on-GPU digest), discard=True
        (generation-only measurement), retain=[indices] (only those are materialised) -- or use submit()/acquire()."""
        ptr, n, keep = self._inputs_ptr(packed, n, staged)
        flags = (RUN_EXPAND if expand else 0) | (RUN_DIGEST if digest else 0) | (RUN_INPUTS_STAGED if staged else 0) | (RUN_DISCARD if discard else 0)
        status = np.zeros(n, dtype=np.uint32)
        outputs = np.zeros((n, max(1, self.n_outputs), 4), dtype=np.uint64)
        digests = np.zeros(n, dtype=np.uint64)
        if retain is None:
            _check(lib().pob_run_batch(self._h, ptr, n, flags, status.ctypes.data, outputs.ctypes.data, digests.ctypes.data if digest else None))
        else:
            r = np.ascontiguousarray(retain, dtype=np.uint32)
            _check(lib().pob_run_batch_retain(self._h, ptr, n, flags, r.ctypes.data if len(r) else None, len(r), status.ctypes.data, outputs.ctypes.data,
                                              digests.ctypes.data if digest else None))
        return BatchResult(status, outputs[:, : self.n_outputs], digests if digest else None, self.last_timing())

    def run(self, inputs, expand=True, digest=False, discard=False, retain=None):
        return self.run_packed(self.pack(inputs), expand=expand, digest=digest, discard=discard, retain=retain)

    def last_timing(self):
        t = Timing()
        _check(lib().pob_last_timing(self._h, ctypes.byref(t)))
        return t.as_dict()

    # ---- consumer-paced hand-off (pob_submit / pob_acquire / pob_release / pob_finish) ----
    def submit(self, packed, n=None, staged=False, digest=False):
        ptr, n, keep = self._inputs_ptr(packed, n, staged)
        self._inflight = (keep, n, digest)                   # the input buffer must outlive the batch
        _check(lib().pob_submit(self._h, ptr, n, (RUN_DIGEST if digest else 0) | (RUN_INPUTS_STAGED if staged else 0)))

    def acquire(self, stream=None):
        """next witness of the submitted batch: (index, device pointer) ; (index, None) for a rejected instance ;
        None when the batch is exhausted.  Raises PobError(E_BUSY) when every slot is held."""
        idx, dptr = ctypes.c_uint32(0), ctypes.c_void_p()
        rc = lib().pob_acquire(self._h, ctypes.byref(idx), ctypes.byref(dptr), stream)
        if rc == DONE:
            return None
        if rc == E_REJECTED:
            return int(idx.value), None
        _check(rc)
        return int(idx.value), dptr.value

    def release(self, index, stream=None):
        _check(lib().pob_release(self._h, index, stream))

    def finish(self):
        keep, n, digest = self._inflight
        status = np.zeros(n, dtype=np.uint32)
        outputs = np.zeros((n, max(1, self.n_outputs), 4), dtype=np.uint64)
        digests = np.zeros(n, dtype=np.uint64)
        _check(lib().pob_finish(self._h, status.ctypes.data, outputs.ctypes.data, digests.ctypes.data if digest else None))
        self._inflight = None
        return BatchResult(status, outputs[:, : self.n_outputs], digests if digest else None, self.last_timing())

    def export_batch(self, packed, paths=None, n=None, staged=False):
        """== n runs of `./<circuit> input_i.json paths[i]` (reference Makefile:5-6): every accepted instance is exported
        while later ones are generated; paths=None (or a None entry) moves the .wtns image to host memory only.
        Returns (BatchResult, export stats)."""
        ptr, n, keep = self._inputs_ptr(packed, n, staged)
        arr = None
        if paths is not None:
            assert len(paths) == n
            arr = (ctypes.c_char_p * n)(*[None if q is None else os.fsencode(q) for q in paths])
        status = np.zeros(n, dtype=np.uint32)
        outputs = np.zeros((n, max(1, self.n_outputs), 4), dtype=np.uint64)
        st = ExportStats()
        _check(lib().pob_export_batch(self._h, ptr, n, RUN_INPUTS_STAGED if staged else 0, arr, status.ctypes.data, outputs.ctypes.data, ctypes.byref(st)))
        return BatchResult(status, outputs[:, : self.n_outputs], None, self.last_timing()), st.as_dict()

    # ---- witness access ----
    def witness(self, index, first=0, count=None):
        count = self.n_signals - first if count is None else count
        out = np.zeros((count, 4), dtype=np.uint64)
        _check(lib().pob_copy_witness(self._h, index, first, count, out.ctypes.data))
        return out

    def write_wtns(self, index, path):
        _check(lib().pob_write_wtns(self._h, index, os.fsencode(path)))

    def selfcheck_keccak(self, index):
        """on-GPU check that every KeccakfRound block of resident witness `index` satisfies out == KeccakRound(in);
        returns (blocks examined, blocks failing)"""
        nb, bad = ctypes.c_uint64(0), ctypes.c_uint64(0)
        _check(lib().pob_selfcheck_keccak(self._h, index, ctypes.byref(nb), ctypes.byref(bad)))
        return int(nb.value), int(bad.value)

    def selfcheck(self, index):
        """on-GPU evaluation of EVERY constraint of the circuit (all `<==` / `===` of the circom sources) against resident
        witness `index`; returns the report dict (n_constraints, n_failed, n_hint_failed, first_failed, signals_read, ms)"""
        r = CheckReport()
        _check(lib().pob_selfcheck(self._h, index, ctypes.byref(r)))
        return r.as_dict()

    def witness_map(self):
        m = np.zeros(self.n_signals, dtype=np.uint32)
        _check(lib().pob_witness_map(self._h, m.ctypes.data))
        return m

    def witness_device_ptr(self, index):
        p = ctypes.c_void_p()
        _check(lib().pob_witness_device_ptr(self._h, index, ctypes.byref(p)))
        return p.value


def pow_grind(start_key, reveal_amount, burn_extra_commitment, zero_bytes=2, max_tries=1 << 32, device=0):
    """GPU replacement of the reference's find_burn_key (tests/main.py:47-56): first burnKey >= start_key whose
    keccak(burnKey | revealAmount | burnExtraCommitment | "EIP-7503") starts with `zero_bytes` zero bytes.
    Returns (burn_key, tries)."""
    a, b, c = (to_limbs([v]) for v in (start_key, reveal_amount, burn_extra_commitment))
    out = np.zeros((1, 4), dtype=np.uint64)
    tries = ctypes.c_uint64(0)
    _check(lib().pob_pow_grind(int(device), a.ctypes.data, b.ctypes.data, c.ctypes.data, int(zero_bytes), int(max_tries), out.ctypes.data, ctypes.byref(tries)))
    return from_limbs(out[0]), int(tries.value)


def repad_pob_input(inp, max_layers, node_blocks, header_blocks):
    """Re-pad a ProofOfBurn input JSON to a circuit shape: unused layers are zero with length 256 (the convention of
    reference tests/main.py:148-150), the header is zero-extended.  The reference generator still pads to the (4,.,5)
    test shape (tests/main.py:8-11) although main_proof_of_burn.circom:27 needs (16,4,16)."""
    out = dict(inp)
    nb, hb = node_blocks * 136, header_blocks * 136
    layers = [list(l)[:nb] + [0] * (nb - len(l)) for l in inp["layers"]]
    lens = list(inp["layerLens"])
    while len(layers) < max_layers:
        layers.append([0] * nb)
        lens.append(256)
    out["layers"], out["layerLens"] = layers[:max_layers], lens[:max_layers]
    hdr = list(inp["blockHeader"])
    out["blockHeader"] = hdr[:hb] + [0] * (hb - len(hdr))
    return out


CIRCUIT_ALIASES = {"main_proof_of_burn": MAIN_PROOF_OF_BURN, "main_spend": MAIN_SPEND}


def main(argv=None):
    """CLI shim with the reference calculator's argv: `python -m pob_b200 <circuit> input.json witness.wtns`
    (reference Makefile:5-6); <circuit> is main_proof_of_burn, main_spend or a `Template(params)` expression.
    Batch form: `python -m pob_b200 <circuit> --batch in1.json in2.json ... --out DIR` evaluates all inputs in one
    pob_run_batch and writes DIR/<name>.wtns for every accepted instance."""
    import sys
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) >= 4 and argv[1] == "--batch" and "--out" in argv:
        k = argv.index("--out")
        files, outdir = argv[2:k], argv[k + 1]
        os.makedirs(outdir, exist_ok=True)
        c = Circuit(CIRCUIT_ALIASES.get(argv[0], argv[0]))
        paths = [os.path.join(outdir, os.path.splitext(os.path.basename(f))[0] + ".wtns") for f in files]
        res, _ = c.export_batch(c.pack([json.load(open(f)) for f in files]), paths)     # consumer-paced: any number of inputs
        rc = 0
        for i, f in enumerate(files):
            if res.status[i] != 0:
                print("%s: constraint failed in the component at witness index %d" % (f, int(res.status[i]) - 1), file=sys.stderr)
                rc = 1
        return rc
    if len(argv) != 3:
        print("usage: python -m pob_b200 <main_proof_of_burn|main_spend|Template(params)> input.json witness.wtns\n"
              "       python -m pob_b200 <circuit> --batch in1.json in2.json ... --out DIR", file=sys.stderr)
        return 2
    c = Circuit(CIRCUIT_ALIASES.get(argv[0], argv[0]), max_slots=1)
    res = c.run([json.load(open(argv[1]))])
    if res.status[0] != 0:
        print("Error: constraint failed in the component at witness index %d" % (int(res.status[0]) - 1), file=sys.stderr)
        return 1
    c.write_wtns(0, argv[2])
    return 0
