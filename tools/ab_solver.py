#!/usr/bin/env python
"""A/B of the single-GPU PCG drivers on one box: solver 0 (two grid barriers per iteration) vs solver 8 (the
barrier-free kernel of the multi-GPU path, self-validating data) on config 2 and config 3."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import irotavg_b200 as ira  # noqa: E402
from oracle import graphs as G  # noqa: E402

sigma = 5 * np.pi / 180
for name, g in (("config 2", G.kitti_like_graph()), ("config 3", G.random_graph())):
    for cost, cname in ((4, "Geman-McClure"), (1, "L1")):
        res = {}
        for sv in (0, 8):
            with ira.Solver(solver=sv) as s:
                s.upload(g.QQ, g.I, g.Q0, g.f)
                s.irls_resident(cost, sigma, 30, -1.0)
                best = None
                for _ in range(3):
                    info = s.irls_resident(cost, sigma, 30, -1.0)
                    best = info.device_ms if best is None else min(best, info.device_ms)
                ph = info.profile.get("pcg_phases") or {}
                Q, w = s.download()
                res[sv] = (best, sum(info.cg_iters), ph.get("kernel_ms"), Q)
        d = float(np.max(np.abs(res[0][3] - res[8][3])))
        print(f"{name} {cname}: solver 0 {res[0][0]:.2f} ms ({res[0][1]} PCG its, kernels {res[0][2]:.2f} ms) | "
              f"solver 8 {res[8][0]:.2f} ms ({res[8][1]} its, kernels {res[8][2]:.2f} ms) | max |dQ| {d:.1e}", flush=True)
