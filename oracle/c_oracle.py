"""ctypes loader of oracle/_build/libestep_oracle.so (TEST INFRASTRUCTURE, see estep_oracle.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libestep_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.oracle_estep.restype = ctypes.c_int
        _lib.oracle_estep.argtypes = [dp, ctypes.c_long, dp, ctypes.c_long, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                      ctypes.c_long, dp, dp, dp, dp]
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return _load().oracle_num_threads()


def expectation_step(t_source, target, sigma2, w=0.0, n_global=None):
    ts = np.ascontiguousarray(t_source, dtype=np.float64)
    tg = np.ascontiguousarray(target, dtype=np.float64)
    m, dim = ts.shape
    n = tg.shape[0]
    pt1, p1, px = np.empty(n), np.empty(m), np.empty((m, dim))
    n_p = ctypes.c_double()
    dp = ctypes.POINTER(ctypes.c_double)
    rc = _load().oracle_estep(ts.ctypes.data_as(dp), m, tg.ctypes.data_as(dp), n, dim, float(sigma2), float(w),
                              n if n_global is None else int(n_global), pt1.ctypes.data_as(dp), p1.ctypes.data_as(dp),
                              px.ctypes.data_as(dp), ctypes.byref(n_p))
    if rc != 0:
        raise MemoryError("oracle_estep failed")
    from .cpd_oracle import Estep
    return Estep(pt1, p1, px, n_p.value)
