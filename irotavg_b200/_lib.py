"""ctypes binding of the C ABI in include/ira.h (irotavg_b200/lib/libira.so).

The library is hand-written sm_100a CUDA; there is no Python or CPU implementation behind these
calls.  Import fails loudly when the shared object has not been built (run
`python -c "import __graft_entry__ as g; g.build()"` or `python -m irotavg_b200.build`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libira.so")

STATS_MAX_ITERS = 256

# every symbol include/ira.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "ira_options_default", "ira_create", "ira_destroy", "ira_status_string", "ira_last_error",
    "ira_abi_version", "ira_device_count", "ira_get_stream", "ira_irls", "ira_problem_upload", "ira_irls_resident",
    "ira_problem_download", "ira_make_A", "ira_quat_normalised", "ira_probe_residual",
    "ira_probe_laplacian_apply", "ira_probe_time_kernel", "ira_comm_unique_id", "ira_comm_init",
    "ira_l1ra", "ira_l1ra_resident", "ira_resident_start", "ira_l1ra_irls", "ira_init_mst", "ira_init_mst_resident",
]


class Options(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("cg_max_iters", C.c_int32),
        ("cg_rtol", C.c_double),
        ("pair_theta", C.c_double),
        ("cg_check_every", C.c_int32),
        ("lanes_per_row", C.c_int32),
        ("world_size", C.c_int32),
        ("rank", C.c_int32),
        ("profile", C.c_int32),
        ("solver", C.c_int32),
        ("spmv_variant", C.c_int32),
        ("small_path", C.c_int32),
        ("shard_mode", C.c_int32),
        ("peer_min_rows", C.c_int32),
        ("reserved", C.c_int32 * 2),
        ("pair_theta3", C.c_double),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("irls_iters", C.c_int32),
        ("cg_iters_total", C.c_int32),
        ("kernel_launches", C.c_int32),
        ("cg_hit_max", C.c_int32),
        ("score", C.c_double * STATS_MAX_ITERS),
        ("cg_iters", C.c_int32 * STATS_MAX_ITERS),
        ("cg_relres", C.c_double * STATS_MAX_ITERS),
        ("t_total_ms", C.c_double),
        ("t_upload_ms", C.c_double),
        ("t_download_ms", C.c_double),
        ("t_residual_ms", C.c_double), ("t_rhs_ms", C.c_double), ("t_spmv_ms", C.c_double),
        ("t_cgvec_ms", C.c_double), ("t_weights_ms", C.c_double), ("t_update_ms", C.c_double),
        ("t_comm_ms", C.c_double),
        ("n_residual", C.c_int32), ("n_rhs", C.c_int32), ("n_spmv", C.c_int32),
        ("n_cgvec", C.c_int32), ("n_weights", C.c_int32), ("n_update", C.c_int32),
        ("n_comm", C.c_int32),
        ("n_pcg", C.c_int32), ("pcg_spmv_phases", C.c_int32),
        ("t_pcg_ms", C.c_double), ("pcg_spmv_ms", C.c_double), ("pcg_update_ms", C.c_double),
        ("pcg_kernel_ms", C.c_double),
        ("pcg_kernel", C.c_int32), ("reserved_stats", C.c_int32),
    ]


class MstStats(C.Structure):
    """ira_mst_stats of include/ira.h."""
    _fields_ = [("passes_label", C.c_int32), ("passes_propagate", C.c_int32), ("unreached", C.c_int32),
                ("t_ms", C.c_double)]


class IraError(RuntimeError):
    def __init__(self, status: int, text: str):
        super().__init__(f"ira status {status}: {text}")
        self.status = status


_lib = None


def load():
    """dlopen libira.so and declare the prototypes.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built and there is no fallback. "
            "Run `python -m irotavg_b200.build`.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    H = C.c_void_p
    i32, i64, f64 = C.c_int32, C.c_int64, C.c_double
    pi32, pf64, pu8 = C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_uint8)
    sig = {
        "ira_options_default": (i32, [C.POINTER(Options)]),
        "ira_create": (i32, [C.POINTER(H), C.POINTER(Options)]),
        "ira_destroy": (i32, [H]),
        "ira_status_string": (C.c_char_p, [i32]),
        "ira_last_error": (C.c_char_p, [H]),
        "ira_abi_version": (i32, []),
        "ira_device_count": (i32, []),
        "ira_get_stream": (i32, [H, C.POINTER(C.c_void_p)]),
        "ira_irls": (i32, [H, i64, i64, i32, pi32, pf64, i64, pf64, i64, i32, f64, i32, f64, pf64,
                           pi32, pf64, C.POINTER(Stats)]),
        "ira_problem_upload": (i32, [H, i64, i64, i32, pi32, pf64, i64, pf64, i64]),
        "ira_irls_resident": (i32, [H, i32, f64, i32, f64, pi32, pf64, C.POINTER(Stats)]),
        "ira_problem_download": (i32, [H, pf64, i64, pf64]),
        "ira_make_A": (i32, [i64, i32, i32, pi32, pi32, pi32]),
        "ira_quat_normalised": (i32, [pf64, i64, i64, i32]),
        "ira_probe_residual": (i32, [H, pf64]),
        "ira_probe_laplacian_apply": (i32, [H, pf64, pf64, pf64]),
        "ira_probe_time_kernel": (i32, [H, i32, i32, i32, pf64]),
        "ira_comm_unique_id": (i32, [pu8]),
        "ira_comm_init": (i32, [H, pu8]),
        "ira_l1ra": (i32, [H, i64, i64, i32, pi32, pf64, i64, pf64, i64, i32, f64, pi32, pf64, C.POINTER(Stats)]),
        "ira_l1ra_resident": (i32, [H, i32, f64, pi32, pf64, C.POINTER(Stats)]),
        "ira_resident_start": (i32, [H, i32]),
        "ira_l1ra_irls": (i32, [H, i64, i64, i32, pi32, pf64, i64, pf64, i64, i32, f64, i32, f64, i32, f64, pf64, pi32,
                                pi32, pf64, C.POINTER(Stats)]),
        "ira_init_mst": (i32, [H, i64, i64, i32, pi32, pf64, i64, pf64, i64, C.POINTER(MstStats)]),
        "ira_init_mst_resident": (i32, [H, i32, C.POINTER(MstStats)]),
    }
    for name, (res, args) in sig.items():
        if not hasattr(lib, name):
            continue        # test_abi reports missing symbols; keep the loader usable meanwhile
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
