"""ctypes loader of the C restatement (oracle/kmer_port.c).  TEST/BASELINE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle_np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_oracle_c.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        _lib = C.CDLL(_SO)
        _lib.oracle_count.restype = C.c_int
        _lib.oracle_count.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_int]
        _lib.oracle_am_clr.restype = C.c_int
        _lib.oracle_am_clr.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int]
    return _lib


def count(seq: np.ndarray, offsets: np.ndarray, k: int = 5, threads: int = 1) -> np.ndarray:
    names, lut = oracle_np.kmer_columns(k)
    lut = np.ascontiguousarray(lut, dtype=np.int32)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    out = np.empty((n, len(names)), dtype=np.int32)
    rc = lib().oracle_count(seq.ctypes.data, offsets.ctypes.data, n, k, lut.ctypes.data, len(names),
                            out.ctypes.data, threads)
    assert rc == 0
    return out


def am_clr(counts: np.ndarray, threads: int = 1) -> np.ndarray:
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    out = np.empty(counts.shape, dtype=np.float64)
    rc = lib().oracle_am_clr(counts.ctypes.data, counts.shape[0], counts.shape[1], out.ctypes.data, threads)
    assert rc == 0
    return out
