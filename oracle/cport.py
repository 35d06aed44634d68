"""ctypes wrapper of the C restatement (oracle/irls_oracle.c).  Test / bench infrastructure only."""
import ctypes as C
import os

import numpy as np

from . import cbuild

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(cbuild.build())
        lib.ora_irls.restype = C.c_int
        lib.ora_irls.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.c_int32, C.c_double, C.c_int32, C.c_double, C.c_double,
                                 C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_double), C.POINTER(C.c_int32)]
        lib.ora_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def available() -> bool:
    try:
        _load()
        return True
    except Exception:
        return False


def max_threads() -> int:
    return int(_load().ora_max_threads())


def irls(QQ, I, cost, sigma, Q, f, max_iters, change_th, cg_rtol=1e-13, cg_max_iters=200000, threads=0):
    """Same contract as oracle.irls_oracle.irls(solver='pcg'); returns a dict."""
    lib = _load()
    QQc = np.ascontiguousarray(QQ, dtype=np.float64)
    Ic = np.ascontiguousarray(np.asarray(I, dtype=np.int32).reshape(-1, 2))
    Qc = np.array(Q, dtype=np.float64, order="C", copy=True)
    m, n = QQc.shape[0], Qc.shape[0]
    weights = np.empty(max(m, 1), dtype=np.float64)
    iters = C.c_int32(0)
    scores = np.zeros(max(max_iters, 1))
    cg = np.zeros(max(max_iters, 1), dtype=np.int32)
    pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    pi = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    rc = lib.ora_irls(m, n, int(f), pi(Ic), pd(QQc), pd(Qc), int(cost), float(sigma), int(max_iters), float(change_th),
                      float(cg_rtol), int(cg_max_iters), int(threads), pd(weights), C.byref(iters), pd(scores), pi(cg))
    if rc != 0:
        raise RuntimeError(f"ora_irls failed: {rc}")
    k = iters.value
    return {"Q": Qc, "weights": weights[:m], "iters": k, "scores": scores[:k].tolist(), "cg_iters": cg[:k].tolist()}
