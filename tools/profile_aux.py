#!/usr/bin/env python
"""Commands profiled under ncu for the kernels outside the IRLS step: the single-block window solver
(k_small_l1ra_irls), init_mst (k_mst_labels / k_mst_propagate) on the 1M-edge graph.

    python tools/profile_aux.py [--reps 3]
Prints wall / device timings of each call."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import irotavg_b200 as ira  # noqa: E402
from oracle import graphs as G  # noqa: E402
from oracle import irls_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--skip-mst", action="store_true")
a = ap.parse_args()
sigma = 5 * np.pi / 180
s = ira.Solver()
out = {}
# a rotAvg(10) window: 15 views, 41 edges, 4 fixed (what ViewGraph::rotAvg hands over in config 5)
g = G.small_graph(n=15, extra=27, sigma_n=0.005, sigma_init=0.05, seed=1, f=4)
ts = []
for _ in range(a.reps + 2):
    t0 = time.perf_counter()
    Q, w, l1_it, info = s.l1ra_irls(g.QQ, g.I, g.Q0, g.f, 100, 1e-3, O.GEMAN_MCCLURE, sigma, 100, 1e-3)
    ts.append((time.perf_counter() - t0) * 1e6)
out["window_solver"] = {"n": g.n, "m": g.m, "f": g.f, "l1_iters": l1_it, "irls_iters": info.iters,
                        "wall_us_per_call_incl_python": ts[2:], "launches": info.kernel_launches}
if not a.skip_mst:
    g = G.random_graph()
    Q0 = np.zeros_like(g.Q0)
    Q0[0] = g.Q0[0]
    s.upload(g.QQ, g.I, Q0, 1)
    for _ in range(a.reps):
        st = s.init_mst_resident(1)
    out["init_mst_config3"] = st
    gk = G.kitti_like_graph()
    Q0 = np.zeros_like(gk.Q0)
    Q0[0] = gk.Q0[0]
    s.upload(gk.QQ, gk.I, Q0, 1)
    for _ in range(a.reps):
        st = s.init_mst_resident(1)
    out["init_mst_config2"] = st
print(json.dumps(out))
