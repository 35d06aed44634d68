"""ctypes access to oracle/_ref/libral_ref.so = the REFERENCE's own ral/l1_irls.cpp (compiled unmodified by
oracle/build_ref.py against the stand-in Eigen / SuiteSparse headers of oracle/ref_shim/).  TEST INFRASTRUCTURE ONLY.

Same calling conventions as oracle/irls_oracle.py (row-major (m, 4) / (n, 4) arrays, quaternions [x y z w])."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import build_ref

_lib = None


def available() -> bool:
    try:
        _load()
        return True
    except Exception:
        return False


def _load():
    global _lib
    if _lib is None:
        build_ref.build()
        lib = C.CDLL(build_ref.LIB)
        pd, pi, pl = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_long)
        lib.ref_irls.argtypes = [C.c_long, C.c_long, C.c_int, pi, pd, pd, C.c_int, C.c_double, C.c_int, C.c_double, pd, pi]
        lib.ref_l1ra.argtypes = [C.c_long, C.c_long, C.c_int, pi, pd, pd, C.c_int, C.c_double, pi]
        lib.ref_init_mst.argtypes = [C.c_long, C.c_long, C.c_int, pi, pd, pd]
        lib.ref_quat_normalised.argtypes = [C.c_long, C.c_int, pd]
        lib.ref_make_A.argtypes = [C.c_long, C.c_long, C.c_int, pi, pl, pl, pd]
        lib.ref_make_A.restype = C.c_long
        lib.ref_residual.argtypes = [C.c_long, C.c_long, pi, pd, pd, pd]
        lib.ref_exp_map.argtypes = [C.c_long, pd]
        lib.ref_quat_mult.argtypes = [pd, pd, pd]
        _lib = lib
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _prep(QQ, I, Q):
    return (np.ascontiguousarray(QQ, dtype=np.float64), np.ascontiguousarray(np.asarray(I, dtype=np.int32).reshape(-1, 2)),
            np.array(Q, dtype=np.float64, order="C", copy=True))


def irls(QQ, I, cost, sigma, Q, f, max_iters, change_th):
    """irotavg::irls (ral/l1_irls.cpp:559-752) after make_A.  Returns (Q, weights, iters)."""
    QQc, Ic, Qc = _prep(QQ, I, Q)
    m, n = QQc.shape[0], Qc.shape[0]
    w = np.zeros(max(m, 1))
    it = C.c_int(0)
    _load().ref_irls(m, n, int(f), _i(Ic), _d(QQc), _d(Qc), int(cost), float(sigma), int(max_iters), float(change_th),
                     _d(w), C.byref(it))
    return Qc, w[:m], it.value


def l1ra(QQ, I, Q, f, max_iters, change_th):
    """irotavg::l1ra (ral/l1_irls.cpp:851-912).  Returns (Q, iters)."""
    QQc, Ic, Qc = _prep(QQ, I, Q)
    it = C.c_int(0)
    _load().ref_l1ra(QQc.shape[0], Qc.shape[0], int(f), _i(Ic), _d(QQc), _d(Qc), int(max_iters), float(change_th), C.byref(it))
    return Qc, it.value


def init_mst(Q, QQ, I, f):
    """irotavg::init_mst (ral/l1_irls.cpp:915-979); exits the process when the edges do not span (as the reference)."""
    QQc, Ic, Qc = _prep(QQ, I, Q)
    _load().ref_init_mst(QQc.shape[0], Qc.shape[0], int(f), _i(Ic), _d(QQc), _d(Qc))
    return Qc


def quat_normalised(Q, f):
    Qc = np.array(Q, dtype=np.float64, order="C", copy=True)
    _load().ref_quat_normalised(Qc.shape[0], int(f), _d(Qc))
    return Qc


def make_A(n, f, I):
    """irotavg::make_A (ral/l1_irls.cpp:755-780) as a dense (m, n - f) array."""
    Ic = np.ascontiguousarray(np.asarray(I, dtype=np.int32).reshape(-1, 2))
    m = Ic.shape[0]
    lib = _load()
    nnz = lib.ref_make_A(m, n, int(f), _i(Ic), None, None, None)
    r = np.zeros(max(nnz, 1), dtype=np.int64); c = np.zeros(max(nnz, 1), dtype=np.int64); v = np.zeros(max(nnz, 1))
    lib.ref_make_A(m, n, int(f), _i(Ic), r.ctypes.data_as(C.POINTER(C.c_long)), c.ctypes.data_as(C.POINTER(C.c_long)), _d(v))
    A = np.zeros((m, n - f))
    A[r[:nnz], c[:nnz]] = v[:nnz]
    return A


def residual(I, QQ, Q):
    """log_map(delta_rel(I, QQ, Q)) (ral/l1_irls.cpp:592-593): (m, 4) rows [w_x w_y w_z theta]."""
    QQc, Ic, Qc = _prep(QQ, I, Q)
    out = np.zeros((QQc.shape[0], 4))
    _load().ref_residual(QQc.shape[0], Qc.shape[0], _i(Ic), _d(QQc), _d(Qc), _d(out))
    return out


def exp_map(W):
    Wc = np.array(W, dtype=np.float64, order="C", copy=True)
    _load().ref_exp_map(Wc.shape[0], _d(Wc))
    return Wc


def quat_mult(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    out = np.zeros(4)
    _load().ref_quat_mult(_d(a), _d(b), _d(out))
    return out


def cli(args, cwd=None, threads=None):
    """Runs the reference CLI (ral/test.cpp) with the given argument list; returns CompletedProcess."""
    build_ref.build()
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    return subprocess.run([build_ref.CLI] + [str(a) for a in args], capture_output=True, text=True, cwd=cwd, env=env)


def read_cli_output(path, n, m):
    """The CLI's output file (ral/test.cpp:314-326): n rows 'w x y z' then m weights -> (Q [x y z w], weights)."""
    vals = np.array(open(path).read().split(), dtype=np.float64)
    Q = vals[:4 * n].reshape(n, 4)[:, [1, 2, 3, 0]].copy()
    return Q, vals[4 * n:4 * n + m].copy()
