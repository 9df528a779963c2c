"""ctypes binding of oracle/liboracle_knn.so (test infrastructure, see oracle_knn.c header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .oracle_np import DMATCH_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle_knn.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liboracle_knn.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_pairs_unordered.restype = C.c_long
        _LIB.orc_pairs_video.restype = C.c_long
        _LIB.orc_pairs_grid.restype = C.c_long
    return _LIB


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def _pairs(fn, *args):
    n = fn(*args, None, C.c_long(0))
    if n < 0:
        raise ValueError("invalid pairing parameters")
    out = np.zeros((n, 2), np.int32)
    fn(*args, _p(out), C.c_long(n))
    return out


def pairs_unordered(n):
    return _pairs(lib().orc_pairs_unordered, C.c_int(n))


def pairs_video(n, seq):
    return _pairs(lib().orc_pairs_video, C.c_int(n), C.c_int(seq))


def pairs_grid(n, seq, rowlen):
    return _pairs(lib().orc_pairs_grid, C.c_int(n), C.c_int(seq), C.c_int(rowlen))


def knn2(q, t, norm):
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    idx = np.zeros((q.shape[0], 2), np.int32)
    dist = np.zeros((q.shape[0], 2), np.float32)
    fn = lib().orc_knn2_l2_u8 if norm == 4 else lib().orc_knn2_hamming
    fn(_p(q), C.c_int(q.shape[0]), _p(t), C.c_int(t.shape[0]), C.c_int(q.shape[1]), _p(idx), _p(dist))
    return idx, dist


def match_pairs(bank, pairs, norm, ratio=0.7, threads=0, want_matches=True):
    bank = [np.ascontiguousarray(b, np.uint8) for b in bank]
    pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
    n = len(pairs)
    rows = (C.c_void_p * len(bank))(*[b.ctypes.data for b in bank])
    nrows = np.array([b.shape[0] for b in bank], np.int32)
    counts = np.zeros(n, np.int32)
    stride = int(max([bank[l].shape[0] for l, _ in pairs] + [1]))
    out = np.zeros((n, stride), DMATCH_DTYPE) if want_matches else None
    err = lib().orc_match_pairs(rows, _p(nrows), C.c_int(bank[0].shape[1] if bank else 0), C.c_int(norm),
                                _p(pairs), C.c_long(n), C.c_double(ratio), _p(counts),
                                _p(out) if out is not None else None, C.c_long(stride), C.c_int(threads))
    if err:
        raise MemoryError
    if out is None:
        return counts
    return [out[p, :counts[p]].copy() for p in range(n)]


def max_threads():
    return int(lib().orc_max_threads())
