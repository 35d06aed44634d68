"""Host-side mirror of the reference's RAL interface (ral/l1_irls.hpp:89-112) over the C ABI.

Same names, argument order and meaning as the reference's free functions; numpy arrays stand in
for Eigen matrices (QQ: m x 4, Q: n x 4, rows [x y z w]; I: m x 2 int).  Every call goes through
include/ira.h into the sm_100a CUDA library - nothing here computes on the CPU except the two
trivial host helpers the C ABI itself exposes (make_A's index rule, quat_normalised).
The C++ twin of this file is irotavg_b200/host/l1_irls.hpp.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import IraError, MstStats, Options, Stats

# enum Cost (ral/l1_irls.hpp:56-57)
L2, L1, L15, L05, Geman_McClure, Huber, Pseudo_Huber, Andrews, Bisquare, Cauchy, Fair, Logistic, \
    Talwar, Welsch = range(14)


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _pi(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _colmajor(a, cols):
    """Eigen::MatrixXd memory: column-major float64.  Returns an F-ordered copy/view (rows x cols)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2 or a.shape[1] != cols:
        raise ValueError(f"expected (*, {cols}) array, got {a.shape}")
    return np.asfortranarray(a)


def _pairs(I):
    I = np.ascontiguousarray(np.asarray(I, dtype=np.int32).reshape(-1, 2))
    return I


@dataclass
class IrlsInfo:
    iters: int = 0
    runtime: float = 0.0
    scores: list = field(default_factory=list)
    cg_iters: list = field(default_factory=list)
    cg_relres: list = field(default_factory=list)
    cg_hit_max: int = 0
    kernel_launches: int = 0
    device_ms: float = 0.0
    upload_ms: float = 0.0
    download_ms: float = 0.0
    profile: dict = field(default_factory=dict)
    pcg_kernel: int = 0


def _info(st: Stats, iters: int, runtime: float) -> IrlsInfo:
    k = min(st.irls_iters, _lib.STATS_MAX_ITERS)
    prof = {}
    if st.pcg_spmv_phases:
        prof["pcg_phases"] = {"spmv_ms": st.pcg_spmv_ms, "update_ms": st.pcg_update_ms,
                              "kernel_ms": st.pcg_kernel_ms, "spmv_phases": st.pcg_spmv_phases,
                              "spmv_us_per_phase": 1000.0 * st.pcg_spmv_ms / st.pcg_spmv_phases}
    for name in ("residual", "rhs", "spmv", "cgvec", "weights", "update", "comm", "pcg"):
        cnt = getattr(st, "n_" + name)
        if cnt:
            prof[name] = {"ms": getattr(st, "t_" + name + "_ms"), "launches": cnt}
    return IrlsInfo(iters=iters, runtime=runtime, scores=list(st.score[:k]),
                    cg_iters=list(st.cg_iters[:k]), cg_relres=list(st.cg_relres[:k]),
                    cg_hit_max=st.cg_hit_max, kernel_launches=st.kernel_launches,
                    device_ms=st.t_total_ms, upload_ms=st.t_upload_ms, download_ms=st.t_download_ms,
                    profile=prof, pcg_kernel=st.pcg_kernel)


class Solver:
    """One ira_handle: a CUDA stream, cached HBM workspace and (optionally) an NCCL communicator."""

    def __init__(self, device: int = -1, cg_rtol: float | None = None, cg_max_iters: int | None = None,
                 cg_check_every: int | None = None, lanes_per_row: int = 0, world_size: int = 1,
                 rank: int = 0, profile: bool = False, solver: int = 0, spmv_variant: int = 0,
                 pair_theta: float | None = None, small_path: bool = True, shard_mode: int = 0,
                 pair_theta3: float | None = None, peer_min_rows: int | None = None):
        self._lib = _lib.load()
        opt = Options()
        self._check(self._lib.ira_options_default(C.byref(opt)), None)
        opt.device = device
        if cg_rtol is not None:
            opt.cg_rtol = cg_rtol
        if cg_max_iters is not None:
            opt.cg_max_iters = cg_max_iters
        if cg_check_every is not None:
            opt.cg_check_every = cg_check_every
        opt.lanes_per_row = lanes_per_row
        opt.world_size = world_size
        opt.rank = rank
        opt.profile = int(profile)
        opt.solver = solver
        opt.spmv_variant = spmv_variant
        if pair_theta is not None:
            opt.pair_theta = pair_theta
        opt.small_path = 0 if small_path else 1
        opt.shard_mode = shard_mode
        if pair_theta3 is not None:
            opt.pair_theta3 = pair_theta3
        if peer_min_rows is not None:
            opt.peer_min_rows = peer_min_rows
        self.options = opt
        self._h = C.c_void_p()
        self._check(self._lib.ira_create(C.byref(self._h), C.byref(opt)), None)
        self._keep = None

    # -- plumbing -----------------------------------------------------------------------------
    def _check(self, status, handle, allow=()):
        if status != 0 and status not in allow:
            text = self._lib.ira_status_string(status).decode()
            if handle:
                detail = self._lib.ira_last_error(handle).decode()
                if detail:
                    text += ": " + detail
            raise IraError(status, text)
        return status

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.ira_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def stream_ptr(self) -> int:
        """cudaStream_t of this handle as an integer (for torch.cuda.ExternalStream)."""
        p = C.c_void_p()
        self._check(self._lib.ira_get_stream(self._h, C.byref(p)), self._h)
        return p.value or 0

    # -- irotavg::irls ------------------------------------------------------------------------
    def irls_inplace(self, QQ_f, I_pairs, Q_f, weights, cost, sigma, f, max_iters, change_th) -> IrlsInfo:
        """ira_irls on caller-owned buffers without any copy: QQ_f (m x 4) and Q_f (n x 4) must be
        F-ordered float64, I_pairs C-ordered int32 (m x 2), weights float64 (m).  Q_f and weights
        are overwritten - exactly the in/out contract of the reference signature."""
        assert QQ_f.flags.f_contiguous and Q_f.flags.f_contiguous and I_pairs.flags.c_contiguous
        assert QQ_f.dtype == np.float64 and Q_f.dtype == np.float64 and I_pairs.dtype == np.int32
        assert weights.dtype == np.float64 and weights.flags.c_contiguous
        m, n = QQ_f.shape[0], Q_f.shape[0]
        iters = C.c_int32(0)
        runtime = C.c_double(0.0)
        st = Stats()
        self._check(self._lib.ira_irls(self._h, m, n, int(f), _pi(I_pairs), _pd(QQ_f), max(m, 1), _pd(Q_f),
                                       max(n, 1), int(cost), float(sigma), int(max_iters),
                                       float(change_th), _pd(weights), C.byref(iters), C.byref(runtime),
                                       C.byref(st)), self._h)
        return _info(st, iters.value, runtime.value)

    def irls(self, QQ, I, A, cost, sigma, Q, f, max_iters, change_th):
        """irotavg::irls (ral/l1_irls.hpp:103-106).  `A` is accepted for signature parity and
        ignored (pure function of (n, f, I)).  Returns (Q_new, weights, info); Q is not modified."""
        QQf = _colmajor(QQ, 4)
        Qf = np.array(_colmajor(Q, 4), order="F", copy=True)
        Ip = _pairs(I)
        m, n = QQf.shape[0], Qf.shape[0]
        if Ip.shape[0] != m:
            raise ValueError("I and QQ disagree on the number of edges")
        weights = np.empty(m, dtype=np.float64)
        iters = C.c_int32(0)
        runtime = C.c_double(0.0)
        st = Stats()
        self._check(self._lib.ira_irls(self._h, m, n, int(f), _pi(Ip), _pd(QQf), max(m, 1), _pd(Qf),
                                       max(n, 1), int(cost), float(sigma), int(max_iters),
                                       float(change_th), _pd(weights), C.byref(iters), C.byref(runtime),
                                       C.byref(st)), self._h)
        return np.ascontiguousarray(Qf), weights, _info(st, iters.value, runtime.value)

    # -- irotavg::l1ra ------------------------------------------------------------------------
    def l1ra(self, QQ, I, A, Q, f, max_iters, change_th):
        """irotavg::l1ra (ral/l1_irls.hpp:98-101).  Returns (Q_new, info); Q is not modified."""
        QQf = _colmajor(QQ, 4)
        Qf = np.array(_colmajor(Q, 4), order="F", copy=True)
        Ip = _pairs(I)
        m, n = QQf.shape[0], Qf.shape[0]
        iters = C.c_int32(0)
        runtime = C.c_double(0.0)
        st = Stats()
        self._check(self._lib.ira_l1ra(self._h, m, n, int(f), _pi(Ip), _pd(QQf), max(m, 1), _pd(Qf), max(n, 1),
                                       int(max_iters), float(change_th), C.byref(iters), C.byref(runtime),
                                       C.byref(st)), self._h)
        return np.ascontiguousarray(Qf), _info(st, iters.value, runtime.value)

    def l1ra_resident(self, max_iters, change_th) -> IrlsInfo:
        iters = C.c_int32(0)
        runtime = C.c_double(0.0)
        st = Stats()
        self._check(self._lib.ira_l1ra_resident(self._h, int(max_iters), float(change_th), C.byref(iters),
                                                C.byref(runtime), C.byref(st)), self._h)
        return _info(st, iters.value, runtime.value)

    def resident_start(self, from_current: bool):
        """False: resident calls restart from the uploaded Q0; True: continue from the current device Q."""
        self._check(self._lib.ira_resident_start(self._h, 1 if from_current else 0), self._h)

    # -- l1ra followed by irls on one upload (both callers' sequence) ---------------------------
    def l1ra_irls(self, QQ, I, Q, f, l1_iters, l1_th, cost, sigma, irls_iters, irls_th):
        """ira_l1ra_irls: src/ViewGraph.cpp:1402-1417 / ral/test.cpp:295-300.  Returns (Q, weights, l1_iters,
        irls info)."""
        QQf = _colmajor(QQ, 4)
        Qf = np.array(_colmajor(Q, 4), order="F", copy=True)
        Ip = _pairs(I)
        m, n = QQf.shape[0], Qf.shape[0]
        weights = np.zeros(m, dtype=np.float64)
        l1o, iro = C.c_int32(0), C.c_int32(0)
        runtime = C.c_double(0.0)
        st = Stats()
        self._check(self._lib.ira_l1ra_irls(self._h, m, n, int(f), _pi(Ip), _pd(QQf), max(m, 1), _pd(Qf), max(n, 1),
                                            int(l1_iters), float(l1_th), int(cost), float(sigma), int(irls_iters),
                                            float(irls_th), _pd(weights), C.byref(l1o), C.byref(iro), C.byref(runtime),
                                            C.byref(st)), self._h)
        return np.ascontiguousarray(Qf), weights, l1o.value, _info(st, iro.value, runtime.value)

    # -- irotavg::init_mst --------------------------------------------------------------------
    def init_mst(self, Q, QQ, I, f):
        """irotavg::init_mst (ral/l1_irls.hpp:89-90; argument order of the reference).  Returns
        (Q_new, stats dict); Q is not modified."""
        QQf = _colmajor(QQ, 4)
        Qf = np.array(_colmajor(Q, 4), order="F", copy=True)
        Ip = _pairs(I)
        m, n = QQf.shape[0], Qf.shape[0]
        st = MstStats()
        self._check(self._lib.ira_init_mst(self._h, m, n, int(f), _pi(Ip), _pd(QQf), max(m, 1), _pd(Qf), max(n, 1),
                                           C.byref(st)), self._h)
        return np.ascontiguousarray(Qf), {"passes_label": st.passes_label, "passes_propagate": st.passes_propagate,
                                          "unreached": st.unreached, "ms": st.t_ms}

    def init_mst_resident(self, f_init):
        st = MstStats()
        self._check(self._lib.ira_init_mst_resident(self._h, int(f_init), C.byref(st)), self._h)
        return {"passes_label": st.passes_label, "passes_propagate": st.passes_propagate,
                "unreached": st.unreached, "ms": st.t_ms}

    # -- device-resident variant --------------------------------------------------------------
    def upload(self, QQ, I, Q0, f):
        QQf = _colmajor(QQ, 4)
        Qf = _colmajor(Q0, 4)
        Ip = _pairs(I)
        self._shape = (QQf.shape[0], Qf.shape[0], int(f))
        self._check(self._lib.ira_problem_upload(self._h, QQf.shape[0], Qf.shape[0], int(f), _pi(Ip),
                                                 _pd(QQf), max(QQf.shape[0], 1), _pd(Qf),
                                                 max(Qf.shape[0], 1)), self._h)

    def irls_resident(self, cost, sigma, max_iters, change_th) -> IrlsInfo:
        iters = C.c_int32(0)
        runtime = C.c_double(0.0)
        st = Stats()
        self._check(self._lib.ira_irls_resident(self._h, int(cost), float(sigma), int(max_iters),
                                                float(change_th), C.byref(iters), C.byref(runtime),
                                                C.byref(st)), self._h)
        return _info(st, iters.value, runtime.value)

    def download(self):
        m, n, _ = self._shape
        Qf = np.empty((n, 4), dtype=np.float64, order="F")
        weights = np.empty(m, dtype=np.float64)
        self._check(self._lib.ira_problem_download(self._h, _pd(Qf), max(n, 1), _pd(weights)), self._h)
        return np.ascontiguousarray(Qf), weights

    # -- probes -------------------------------------------------------------------------------
    def probe_residual(self):
        m = self._shape[0]
        w = np.empty((m, 4), dtype=np.float64, order="F")
        self._check(self._lib.ira_probe_residual(self._h, _pd(w)), self._h)
        return np.ascontiguousarray(w)

    def probe_laplacian_apply(self, weights, X):
        m, n, f = self._shape
        nf = n - f
        wv = np.ascontiguousarray(weights, dtype=np.float64)
        Xf = _colmajor(X, 3)
        assert wv.shape[0] == m and Xf.shape[0] == nf
        Y = np.empty((nf, 3), dtype=np.float64, order="F")
        self._check(self._lib.ira_probe_laplacian_apply(self._h, _pd(wv), _pd(Xf), _pd(Y)), self._h)
        return np.ascontiguousarray(Y)

    def time_kernel(self, which: int, reps: int = 50, flush_l2: bool = False) -> float:
        us = C.c_double(0.0)
        self._check(self._lib.ira_probe_time_kernel(self._h, which, reps, int(flush_l2), C.byref(us)),
                    self._h)
        return us.value

    # -- communicator -------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        lib = _lib.load()
        buf = (C.c_uint8 * 128)()
        st = lib.ira_comm_unique_id(buf)
        if st != 0:
            raise IraError(st, lib.ira_status_string(st).decode())
        return bytes(buf)

    def comm_init(self, uid: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._check(self._lib.ira_comm_init(self._h, buf), self._h)


_default = None


def _solver() -> Solver:
    global _default
    if _default is None:
        _default = Solver()
    return _default


# ---- free functions with the reference's names ------------------------------------------------
def irls(QQ, I, A, cost, sigma, Q, f, max_iters, change_th):
    """irotavg::irls.  Returns (Q, weights, iters, runtime) - the reference's in/out parameters."""
    Qn, w, info = _solver().irls(QQ, I, A, cost, sigma, Q, f, max_iters, change_th)
    return Qn, w, info.iters, info.runtime


def l1ra(QQ, I, A, Q, f, max_iters, change_th):
    """irotavg::l1ra.  Returns (Q, iter, runtime)."""
    Qn, info = _solver().l1ra(QQ, I, A, Q, f, max_iters, change_th)
    return Qn, info.iters, info.runtime


def init_mst(Q, QQ, I, f):
    """irotavg::init_mst.  Returns the initialised Q."""
    return _solver().init_mst(Q, QQ, I, f)[0]


def make_A(n, f, I):
    """irotavg::make_A (ral/l1_irls.cpp:755-780) as a scipy CSC matrix, via ira_make_A."""
    import scipy.sparse as sp
    lib = _lib.load()
    Ip = _pairs(I)
    m = Ip.shape[0]
    cp = np.empty(m, dtype=np.int32)
    cm = np.empty(m, dtype=np.int32)
    st = lib.ira_make_A(m, int(n), int(f), _pi(Ip), _pi(cp), _pi(cm))
    if st != 0:
        raise IraError(st, lib.ira_status_string(st).decode())
    rows = np.concatenate([np.nonzero(cp >= 0)[0], np.nonzero(cm >= 0)[0]])
    cols = np.concatenate([cp[cp >= 0], cm[cm >= 0]])
    vals = np.concatenate([np.ones((cp >= 0).sum()), -np.ones((cm >= 0).sum())])
    return sp.csc_matrix((vals, (rows, cols)), shape=(m, n - f))


def quat_normalised(Q, f):
    """irotavg::quat_normalised (ral/l1_irls.cpp:982-991).  Returns the normalised copy."""
    lib = _lib.load()
    Qf = np.array(_colmajor(Q, 4), order="F", copy=True)
    st = lib.ira_quat_normalised(_pd(Qf), Qf.shape[0], max(Qf.shape[0], 1), int(f))
    if st != 0:
        raise IraError(st, lib.ira_status_string(st).decode())
    return np.ascontiguousarray(Qf)


def device_count() -> int:
    return int(_lib.load().ira_device_count())
