#!/usr/bin/env python
"""Where a global rotAvg call goes: upload, l1ra (Newton solves, PCG iterations), irls, on a 9 501-view stream graph."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import irotavg_b200 as ira  # noqa: E402
from oracle import irls_oracle as O  # noqa: E402
from oracle import rotavg_stream as RS  # noqa: E402

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 9501
ops, Qgt = RS.make_stream(n_frames=nv, loop_every=500, min_loop_gap=500)
I = np.array([(op[1], op[2]) for op in ops if op[0] == "E"], dtype=np.int32)
QQ = np.array([O.rmat2quat(op[3]) for op in ops if op[0] == "E"])
Q0 = np.tile(np.array([0, 0, 0, 1.0]), (nv, 1))
Q0[:nv - 1] = Qgt[:nv - 1]
sigma = 5 * np.pi / 180
for sv in (0, 256):
    with ira.Solver(solver=sv, profile=1) as s:
        for rep in range(2):
            t0 = time.perf_counter()
            s.upload(QQ, I, Q0, 1)
            t1 = time.perf_counter()
            il = s.l1ra_resident(100, 1e-3)
            t2 = time.perf_counter()
            s.resident_start(True)
            ii = s.irls_resident(ira.Geman_McClure, sigma, 100, 1e-3)
            t3 = time.perf_counter()
        row = {"solver": sv, "views": nv, "upload_ms": 1e3 * (t1 - t0), "l1ra_ms": 1e3 * (t2 - t1), "irls_ms": 1e3 * (t3 - t2),
               "l1ra": {k: (v if not isinstance(v, np.ndarray) else v.tolist()) for k, v in vars(il).items() if k not in ("scores",)},
               "irls": {k: (v if not isinstance(v, np.ndarray) else v.tolist()) for k, v in vars(ii).items() if k not in ("scores",)}}
        print(json.dumps(row, default=str), flush=True)
