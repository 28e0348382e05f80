"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/libcrf_decode_ref.so from oracle/csrc/crf_decode_ref.c (gcc + OpenMP)
and binds it with ctypes.  `oracle/_ref/` is git-ignored (it still travels to the GPU box with the gpurun snapshot).
The reference itself has no C/C++ sources to compile, so `_ref/` only ever holds this restatement.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "crf_decode_ref.c")
OUT = os.path.join(HERE, "_ref", "libcrf_decode_ref.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", OUT, SRC, "-lm"], check=True)
    return OUT


def load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.crf_decode_ref.restype = ctypes.c_int
        lib.crf_decode_ref.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                       ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib = lib
    return _lib


def decode(scores_ntc, state_len, blank_score=2.0, qscale=1.0, qbias=0.0):
    """C counterpart of crf_oracle.decode_native: (moves, sequence, qstring) uint8 [N, T]."""
    x = np.ascontiguousarray(scores_ntc, dtype=np.float32)
    n, t, c = x.shape
    assert c == 4 ** (state_len + 1)
    outs = [np.zeros((n, t), dtype=np.uint8) for _ in range(3)]
    rc = load().crf_decode_ref(x.ctypes.data, n, t, state_len, blank_score, qscale, qbias, *(o.ctypes.data for o in outs))
    if rc != 0:
        raise MemoryError("crf_decode_ref failed")
    return tuple(outs)
