"""ctypes front-end of the TEST-ONLY host emulator (tests/emu/pob_emu.cpp)."""
import ctypes, os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "..", "..", "proof-of-burn_b200", "csrc")
_LIB = None


def build():
    so = os.path.join(_HERE, "libpob_emu.so")
    srcs = [os.path.join(_HERE, "pob_emu.cpp"), os.path.join(_HERE, "host_ref.h")] + [os.path.join(_CSRC, f) for f in
            ("compiler.cpp", "compiler.h", "program.h", "vm_exec.h", "fr_hd.h", "poseidon_constants_data.h", "cons_check.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-I", _CSRC,
                               "-I", _HERE, "-o", so, srcs[0], srcs[2]])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.pob_emu_compile.restype = ctypes.c_void_p
        L.pob_emu_compile.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
        L.pob_emu_free.argtypes = [ctypes.c_void_p]
        L.pob_emu_stats.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.pob_emu_schema.restype = ctypes.c_char_p
        L.pob_emu_schema.argtypes = [ctypes.c_void_p]
        L.pob_emu_run.restype = ctypes.c_uint64
        L.pob_emu_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.pob_emu_check_constraints.restype = ctypes.c_int
        L.pob_emu_check_constraints.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64,
                                                ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
        _LIB = L
    return _LIB


CHECK_NAMES = ["n_constraints", "n_hints", "n_failed", "n_hint_failed", "first_failed", "flat_eq", "flat_kc", "flat_r1",
               "round_eq", "round_kc", "round_r1", "round_blocks", "signals_referenced"]


def check_constraints(name, params_limbs, nparams, witness_limbs, hcreate=False):
    """compile `name(params)` WITH its constraint system and evaluate every record against `witness_limbs` (n x 4 uint64)"""
    out = np.zeros(13, dtype=np.uint64)
    err = ctypes.create_string_buffer(512)
    w = np.ascontiguousarray(witness_limbs, dtype=np.uint64)
    rc = lib().pob_emu_check_constraints(name.encode(), params_limbs.ctypes.data, nparams, int(hcreate), w.ctypes.data, w.shape[0], out.ctypes.data, err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return dict(zip(CHECK_NAMES, (int(v) for v in out)))


STAT_NAMES = ["n_signals", "n_outputs", "n_inputs", "n_words", "n_vals", "n_ops", "n_absorbs", "n_levels", "n_tiles",
              "n_codes", "n_konst", "n_round_blocks"]


class EmuProgram:
    def __init__(self, name, params_limbs, nparams, hcreate=False, opt=0):
        err = ctypes.create_string_buffer(512)
        self.h = lib().pob_emu_compile(name.encode(), params_limbs.ctypes.data, nparams, (1 if hcreate else 0) | (0x100 if opt else 0), err, 512)
        if not self.h:
            raise RuntimeError(err.value.decode())
        st = np.zeros(12, dtype=np.uint64)
        lib().pob_emu_stats(self.h, st.ctypes.data)
        self.stats = dict(zip(STAT_NAMES, (int(v) for v in st)))
        self.schema = lib().pob_emu_schema(self.h).decode()

    def witness_map(self):
        lib().pob_emu_witness_map.restype = ctypes.c_uint64
        lib().pob_emu_witness_map.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        m = np.zeros(self.stats["n_signals"], dtype=np.uint32)
        n0 = lib().pob_emu_witness_map(self.h, m.ctypes.data)
        return m, int(n0)

    def run(self, input_limbs, want_witness=True):
        n = self.stats["n_signals"]
        wit = np.zeros((n, 4), dtype=np.uint64) if want_witness else None
        outs = np.zeros((max(1, self.stats["n_outputs"]), 4), dtype=np.uint64)
        il = np.ascontiguousarray(input_limbs, dtype=np.uint64)
        if il.size == 0:
            il = np.zeros((1, 4), dtype=np.uint64)
        status = lib().pob_emu_run(self.h, il.ctypes.data, wit.ctypes.data if want_witness else None, outs.ctypes.data)
        return int(status), wit, outs[: self.stats["n_outputs"]]

    def __del__(self):
        try:
            if self.h:
                lib().pob_emu_free(self.h)
                self.h = None
        except Exception:
            pass
