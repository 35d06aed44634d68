#!/usr/bin/env python
"""Bitwise repeatability and timing spread of l1ra + irls on a stream graph, per PCG variant (tridiagonal two-level,
dense two-level, one-level)."""
import hashlib
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

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 9001
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ops, Qgt = RS.make_stream(n_frames=nv, loop_every=500, min_loop_gap=500)
I = np.array([(op[1], op[2]) for op in ops if op[0] == "E"], dtype=np.int32)
QQ = np.array([O.rmat2quat(op[3]) for op in ops if op[0] == "E"])
rng = np.random.default_rng(3)
Q0 = O.quat_mult(Qgt, np.concatenate([rng.normal(0, 0.01, (nv, 3)), np.ones((nv, 1))], axis=1))
Q0 /= np.linalg.norm(Q0, axis=1, keepdims=True)
Q0[0] = Qgt[0]
sigma = 5 * np.pi / 180
for sv in (0, 256, 128):
    hs, ts, its = [], [], []
    for fresh in (True, False):
        s = ira.Solver(solver=sv)
        for r in range(reps):
            if fresh and r > 0:
                s.close()
                s = ira.Solver(solver=sv)
            t0 = time.perf_counter()
            Q, w, l1, info = s.l1ra_irls(QQ, I, Q0, 1, 100, 1e-3, ira.Geman_McClure, sigma, 100, 1e-3)
            ts.append(round(1e3 * (time.perf_counter() - t0), 2))
            hs.append(hashlib.sha1(np.ascontiguousarray(Q).tobytes() + np.ascontiguousarray(w).tobytes()).hexdigest()[:10])
            its.append(int(sum(info.cg_iters)))
        s.close()
    print(json.dumps({"solver": sv, "views": nv, "distinct_results": len(set(hs)), "hashes": hs, "ms": ts, "irls_pcg_iters": its}), flush=True)
