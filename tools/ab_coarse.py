#!/usr/bin/env python
"""A/B on one box: two-level PCG (ira_coarse.cuh, default for graphs of up to 32 768 nodes; tridiagonal coarse operator, and
the dense 64-block one with solver +256) against the one-level kernels
(solver +128) on the chain-like graphs: the bundled fixture's CLI flow, a global rotAvg call of the config-5 stream, config 2."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import irotavg_b200 as ira  # noqa: E402
from oracle import graphs as G  # noqa: E402
from oracle import irls_oracle as O  # noqa: E402
from oracle import rotavg_stream as RS  # noqa: E402

sigma = 5 * np.pi / 180
out = []
b = np.load(os.path.join(ROOT, "tests", "golden", "bundled_graph.npz"))
def stream_case(nv):
    ops, Qgt = RS.make_stream(n_frames=nv, loop_every=500, min_loop_gap=500)
    I = np.array([(op[1], op[2]) for op in ops if op[0] == "E"], dtype=np.int32)
    QQ = np.array([O.rmat2quat(op[3]) for op in ops if op[0] == "E"])
    Q0 = np.tile(np.array([0, 0, 0, 1.0]), (nv, 1))
    Q0[:nv - 1] = Qgt[:nv - 1]                             # a global call: the newest view enters at identity
    return QQ, I, Q0


QQ, I, Q0 = stream_case(4001)
QQ9, I9, Q09 = stream_case(9501)
gk = G.kitti_like_graph()
cases = [("bundled CLI flow l1ra(5)+irls(GM,50)", b["QQ"], b["I"], b["Q_mst"], int(b["f"]), 5, 50),
         ("stream global call, 4001 views", QQ, I, Q0, 1, 100, 100),
         ("stream global call, 9501 views", QQ9, I9, Q09, 1, 100, 100),
         ("config 2 l1ra(5)+irls(GM,30)", gk.QQ, gk.I, gk.Q0, gk.f, 5, 30)]
for name, qq, ii, q0, f, l1n, irn in cases:
    row = {"case": name, "n": int(q0.shape[0]), "m": int(len(ii))}
    Qs = {}
    for sv in (0, 256, 128):
        with ira.Solver(solver=sv) as s:
            s.l1ra_irls(qq, ii, q0, f, l1n, 1e-3, ira.Geman_McClure, sigma, irn, 1e-3)
            t0 = time.perf_counter()
            Q, w, l1_it, info = s.l1ra_irls(qq, ii, q0, f, l1n, 1e-3, ira.Geman_McClure, sigma, irn, 1e-3)
            dt = time.perf_counter() - t0
        Qs[sv] = Q
        row[{0: "two_level", 256: "two_level_dense64", 128: "one_level"}[sv]] = {"call_ms": 1e3 * dt, "l1_iters": l1_it, "irls_iters": info.iters,
                                                        "irls_pcg_iters": int(sum(info.cg_iters)), "pcg_kernel": info.pcg_kernel,
                                                        "cg_hit_max": info.cg_hit_max}
    row["rms_between_rad"] = O.geodesic_rms(Qs[0], Qs[128], f)
    row["rms_tri_vs_dense_rad"] = O.geodesic_rms(Qs[0], Qs[256], f)
    out.append(row)
    print(json.dumps(row), flush=True)
with open(os.path.join(ROOT, "gpurun_out", "r02_ab_coarse.json"), "w") as fh:
    json.dump(out, fh, indent=1)
