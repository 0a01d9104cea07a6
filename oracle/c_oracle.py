"""ctypes binding to oracle/libppcseq_oracle.so (the C restatement; ORACLE / CPU BASELINE ONLY)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libppcseq_oracle.so")
    src = os.path.join(_HERE, "ppcseq_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        L.oracle_log_prob_grad.argtypes = [ctypes.c_int] * 4 + [
            ctypes.POINTER(ctypes.c_int32), dp, dp, ctypes.POINTER(ctypes.c_uint8), ctypes.c_double,
            dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp]
        L.oracle_log_prob_grad.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def log_prob_grad(d, theta, propto=True, jacobian=True, n_shards=1):
    """d: oracle.model_np.ModelData.  Returns (lp, grad)."""
    L = lib()
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    grad = np.empty_like(theta)
    lp = ctypes.c_double()
    ex = None
    if d.exclude is not None:
        ex8 = np.ascontiguousarray(d.exclude, dtype=np.uint8)
        ex = _p(ex8, ctypes.c_uint8)
    rc = L.oracle_log_prob_grad(d.G, d.S, d.C, d.K, _p(d.counts, ctypes.c_int32), _p(d.X, ctypes.c_double),
                                _p(d.exposure, ctypes.c_double), ex, d.lambda_mu_mu,
                                _p(theta, ctypes.c_double), int(propto), int(jacobian), int(n_shards),
                                ctypes.byref(lp), _p(grad, ctypes.c_double))
    if rc != 0:
        raise RuntimeError(f"oracle_log_prob_grad rc={rc}")
    return lp.value, grad
