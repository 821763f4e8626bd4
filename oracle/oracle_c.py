"""ctypes binding of oracle/_build/liboracle.so (plain-C oracle).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build() -> str:
    """Compile oracle_c.c with gcc (idempotent)."""
    src = os.path.join(_HERE, "oracle_c.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        P = ctypes.c_void_p
        I = ctypes.c_int64
        L.oracle_cosine_pair.restype = ctypes.c_double
        L.oracle_cosine_pair.argtypes = [P, P, I]
        L.oracle_l2_normalize.restype = None
        L.oracle_l2_normalize.argtypes = [P, I, I, P, P]
        L.oracle_segment_mean.restype = None
        L.oracle_segment_mean.argtypes = [P, I, P, P, I, P]
        L.oracle_cosine_topk.restype = None
        L.oracle_cosine_topk.argtypes = [P, I, P, I, I, I, P, P, P]
        L.oracle_distance_topk.restype = None
        L.oracle_distance_topk.argtypes = [P, I, P, I, I, I, ctypes.c_int, P, P, P]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def cosine_pair(a, b) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return float(lib().oracle_cosine_pair(_p(a), _p(b), a.shape[0]))


def l2_normalize(x):
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
    out = np.empty_like(x)
    norms = np.empty(x.shape[0], dtype=np.float32)
    lib().oracle_l2_normalize(_p(x), x.shape[0], x.shape[1], _p(out), _p(norms))
    return out, norms


def segment_mean(stored, row_idx, offsets):
    stored = np.ascontiguousarray(stored, dtype=np.float32)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    ridx = None if row_idx is None else np.ascontiguousarray(row_idx, dtype=np.int64)
    C = len(offsets) - 1
    out = np.zeros((C, stored.shape[1]), dtype=np.float32)
    lib().oracle_segment_mean(_p(stored), stored.shape[1], _p(ridx), _p(offsets), C, _p(out))
    return out


def cosine_topk(queries, stored, k, row_allowed=None):
    q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
    g = np.ascontiguousarray(stored, dtype=np.float32)
    allowed = None if row_allowed is None else np.ascontiguousarray(row_allowed, dtype=np.uint8)
    out_s = np.empty((q.shape[0], k), dtype=np.float64)
    out_i = np.empty((q.shape[0], k), dtype=np.int64)
    lib().oracle_cosine_topk(_p(q), q.shape[0], _p(g), g.shape[0], g.shape[1], k, _p(allowed), _p(out_s), _p(out_i))
    return out_s, out_i


def distance_topk(queries, stored, k, metric, row_allowed=None):
    """(keys f64 [Q,k] descending, rows i64 [Q,k]); metric in {"dot", "euclid", "manhattan"}; key = q.g / -d^2 / -d."""
    q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
    g = np.ascontiguousarray(stored, dtype=np.float32)
    allowed = None if row_allowed is None else np.ascontiguousarray(row_allowed, dtype=np.uint8)
    out_s = np.empty((q.shape[0], k), dtype=np.float64)
    out_i = np.empty((q.shape[0], k), dtype=np.int64)
    lib().oracle_distance_topk(_p(q), q.shape[0], _p(g), g.shape[0], g.shape[1], k,
                               {"dot": 1, "euclid": 2, "manhattan": 3}[metric], _p(allowed), _p(out_s), _p(out_i))
    return out_s, out_i
