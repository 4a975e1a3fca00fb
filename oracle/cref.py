"""ctypes access to oracle/partition_ref.c (TEST INFRASTRUCTURE, see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def load(build: bool = True):
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "_build", "liboracle_partition.so")
    if not os.path.exists(path) and build:
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(path)
    p, i64, f64 = C.c_void_p, C.c_int64, C.c_double
    lib.oracle_clamp_round.argtypes = [p, i64, f64]
    lib.oracle_clamp_round.restype = None
    lib.oracle_part_from_values.argtypes = [p, i64, p]
    lib.oracle_part_from_values.restype = i64
    lib.oracle_sort_unique.argtypes = [p, i64]
    lib.oracle_sort_unique.restype = i64
    lib.oracle_refine.argtypes = [p, i64, p, i64]
    lib.oracle_refine.restype = i64
    lib.oracle_fill.argtypes = [p, p, p, i64]
    lib.oracle_fill.restype = None
    lib.oracle_round_refine.argtypes = [p, i64, p, i64, f64, p]
    lib.oracle_round_refine.restype = i64
    _LIB = lib
    return lib


def clamp_round(M: np.ndarray, atol: float) -> np.ndarray:
    out = np.array(M, dtype=np.float64, order="F", copy=True)
    load().oracle_clamp_round(out.ctypes.data, out.size, atol)
    return out


def part_from_values(M: np.ndarray):
    Mf = np.asfortranarray(M, dtype=np.float64)
    out = np.empty(Mf.shape, dtype=np.uint32, order="F")
    d = load().oracle_part_from_values(Mf.ctypes.data, Mf.size, out.ctypes.data)
    return int(d), out


def refine(labels: np.ndarray, dim1: int, p2: np.ndarray):
    l = np.array(labels, dtype=np.uint64, order="F", copy=True)
    p2 = np.asfortranarray(p2, dtype=np.uint32)
    d = load().oracle_refine(l.ctypes.data, dim1, p2.ctypes.data, l.size)
    return int(d), l


def round_refine(labels: np.ndarray, dim1: int, M: np.ndarray, atol: float):
    """refine!(S, Part(_clamp_round!(M))) in place on copies; returns (dim, labels)."""
    l = np.array(labels, dtype=np.uint64, order="F", copy=True)
    Mf = np.array(M, dtype=np.float64, order="F", copy=True)
    scratch = np.empty(Mf.size, dtype=np.uint32)
    d = load().oracle_round_refine(l.ctypes.data, dim1, Mf.ctypes.data, Mf.size, atol, scratch.ctypes.data)
    return int(d), l
