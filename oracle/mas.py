"""MAS oracle front-end (TEST INFRASTRUCTURE ONLY).

`mas_width1_numpy` is a pure-Python/numpy restatement of reference alignment.py:31-59 for small cases;
`binarize` / `mas_width1` call the C restatement in oracle/mas_oracle.c (built by oracle/Makefile).
Parity status: pinned against reference-generated golden vectors (tests/golden/mas_*.npz).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libmas_oracle.so")
    src = os.path.join(_HERE, "mas_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libmas_oracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        lib.mas_oracle_width1.restype = ctypes.c_int
        lib.mas_oracle_width1.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_long, ctypes.c_void_p, ctypes.c_long]
        lib.mas_oracle_binarize.restype = ctypes.c_int
        lib.mas_oracle_binarize.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.mas_oracle_logf.restype = None
        lib.mas_oracle_logf.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
        _LIB = lib
    return _LIB


def libm_logf(x):
    """float32 log through the host libm `logf` -- the function Numba's np.log(float32) lowers to."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    _lib().mas_oracle_logf(x.ctypes.data, y.ctypes.data, x.size)
    return y


def mas_width1(attn_map, is_prob=True):
    """(T1, T2) float32 -> (T1, T2) float32 {0,1}; see alignment.py:31-59."""
    a = np.ascontiguousarray(attn_map, dtype=np.float32)
    t1, t2 = a.shape
    opt = np.zeros_like(a)
    rc = _lib().mas_oracle_width1(a.ctypes.data, int(bool(is_prob)), t1, t2, t2, opt.ctypes.data, t2)
    if rc:
        raise MemoryError("mas_oracle_width1")
    return opt


def binarize(attn, in_lens, out_lens, is_prob=True):
    """(B,1,T1,T2) float32 + lens -> dense hard map, as RADTTS.binarize_attention (radtts.py:320-334)."""
    a = np.ascontiguousarray(attn, dtype=np.float32)
    b, one, t1, t2 = a.shape
    assert one == 1
    il = np.ascontiguousarray(in_lens, dtype=np.int64)
    ol = np.ascontiguousarray(out_lens, dtype=np.int64)
    out = np.empty_like(a)
    rc = _lib().mas_oracle_binarize(a.ctypes.data, int(bool(is_prob)), il.ctypes.data, ol.ctypes.data,
                                    b, t1, t2, out.ctypes.data)
    if rc:
        raise MemoryError("mas_oracle_binarize")
    return out


def mas_width1_numpy(logp):
    """Pure-Python restatement on LOG-probabilities (small cases only)."""
    a = np.array(logp, dtype=np.float32, copy=True)
    t1, t2 = a.shape
    a[0, 1:] = -np.inf
    score = a[0].copy()
    came_diag = np.zeros((t1, t2), dtype=np.int64)
    for i in range(1, t1):
        new = np.empty_like(score)
        for j in range(t2):
            src = j
            if j >= 1 and score[j - 1] >= score[j]:
                src = j - 1
            new[j] = np.float32(a[i, j]) + np.float32(score[src])
            came_diag[i, j] = j - src
        score = new
    opt = np.zeros((t1, t2), dtype=np.float32)
    j = t2 - 1
    for i in range(t1 - 1, -1, -1):
        opt[i, j] = 1
        j -= came_diag[i, j]
    opt[0, 0] = 1
    return opt
