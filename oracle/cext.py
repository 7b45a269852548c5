"""ctypes loader for oracle/oracle.c (TEST INFRASTRUCTURE, see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            try:
                build()
            except Exception:
                return None
        _lib = ctypes.CDLL(_SO)
        dp, ip, i64 = (ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64),
                       ctypes.c_int64)
        _lib.oracle_inverse_cdf.argtypes = [dp, i64, dp, i64, ip]
        _lib.oracle_searchsorted_left.argtypes = [dp, i64, dp, i64, ip]
        _lib.oracle_weights.argtypes = [dp, i64, dp, dp]
    return _lib


def available():
    return _load() is not None


def _d(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _i(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def inverse_cdf(su, W):
    su = np.ascontiguousarray(su, dtype=np.float64)
    W = np.ascontiguousarray(W, dtype=np.float64)
    A = np.empty(su.shape[0], dtype=np.int64)
    _load().oracle_inverse_cdf(_d(su), su.shape[0], _d(W), W.shape[0], _i(A))
    return A


def searchsorted_left(cdf, su):
    cdf = np.ascontiguousarray(cdf, dtype=np.float64)
    su = np.ascontiguousarray(su, dtype=np.float64)
    A = np.empty(su.shape[0], dtype=np.int64)
    _load().oracle_searchsorted_left(_d(cdf), cdf.shape[0], _d(su), su.shape[0], _i(A))
    return A


def weights(lw):
    lw = np.ascontiguousarray(lw, dtype=np.float64)
    W = np.empty_like(lw)
    out = np.empty(4)
    _load().oracle_weights(_d(lw), lw.shape[0], _d(W), _d(out))
    return W, out
