#!/usr/bin/env python
"""Do the block-Jacobi thresholds found on config 3 / L1 (tools/sweep_theta.py) hold elsewhere?  Old defaults (0.2, 0.05)
against (0.5, 0.001) on: config 3 with L1.5 / L0.5 / Geman-McClure / Huber, config 2 with L1 / Geman-McClure, the x2 graph
(200k / 2M, HBM-vector kernel) with L1, a 20k / 200k graph with L1."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import irotavg_b200 as ira  # noqa: E402
from oracle import graphs as G  # noqa: E402
from oracle import irls_oracle as O  # noqa: E402

sigma = 5 * np.pi / 180
g3 = G.random_graph()
cases = [("config3 L1.5", g3, 2), ("config3 L0.5", g3, 3), ("config3 GM", g3, 4), ("config3 Huber", g3, 5),
         ("config2 L1", G.kitti_like_graph(), 1), ("config2 GM", G.kitti_like_graph(), 4),
         ("20k/200k L1", G.random_graph(n=20000, m=200000), 1), ("x2 200k/2M L1", G.random_graph(n=200000, m=2000000), 1)]
rows = []
for name, g, cost in cases:
    row = {"case": name}
    Qs = []
    for th, th3 in ((0.2, 0.05), (0.5, 0.001)):
        with ira.Solver(pair_theta=th, pair_theta3=th3) as s:
            s.upload(g.QQ, g.I, g.Q0, g.f)
            s.irls_resident(cost, sigma, 30, -1.0)
            info = s.irls_resident(cost, sigma, 30, -1.0)
            Q, w = s.download()
        Qs.append(Q)
        row[f"theta_{th}_{th3}"] = {"ms": info.device_ms, "pcg_iters": int(sum(info.cg_iters)), "hit_max": info.cg_hit_max}
    row["rms_between"] = O.geodesic_rms(Qs[0], Qs[1], g.f)
    rows.append(row)
    print(json.dumps(row), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r02_sweep_theta_broad.json"), "w"), indent=1)
