"""irotavg-b200: the IRLS rotation-averaging core of ajparra/iRotAvg as hand-written sm_100a CUDA
behind a C ABI (include/ira.h).  `irotavg_b200.api` mirrors the reference's RAL interface
(ral/l1_irls.hpp:89-112); importing it requires the built CUDA library - there is no fallback."""
from . import _lib  # noqa: F401
from .api import (Solver, IraError, irls, l1ra, make_A, quat_normalised, device_count,  # noqa: F401
                  L2, L1, L15, L05, Geman_McClure, Huber, Pseudo_Huber, Andrews, Bisquare, Cauchy,
                  Fair, Logistic, Talwar, Welsch)

__all__ = ["Solver", "IraError", "irls", "l1ra", "make_A", "quat_normalised", "device_count"]
