#!/usr/bin/env python
"""Sweep of the block-Jacobi thresholds (ira_options.pair_theta: mutual 2x2 pairs, pair_theta3: third member) on config 3,
L1, 30 IRLS iterations: PCG iterations per step, ms per step, RMS against the 30-iteration golden."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import irotavg_b200 as ira  # noqa: E402
from oracle import graphs as G  # noqa: E402
from oracle import irls_oracle as O  # noqa: E402

g = G.random_graph()
gold = np.load(os.path.join(ROOT, "tests", "golden", "cfg3_l1_30iters.npz"))
sigma = 5 * np.pi / 180
rows = []
for th, th3 in ((0.4, 0.001), (0.5, 0.001), (0.6, 0.001), (0.7, 0.001), (0.8, 0.001), (0.9, 0.001), (0.6, 0.01), (0.6, 0.05)):
    with ira.Solver(pair_theta=th, pair_theta3=th3) as s:
        s.upload(g.QQ, g.I, g.Q0, g.f)
        s.irls_resident(1, sigma, 30, -1.0)
        best = None
        for _ in range(2):
            info = s.irls_resident(1, sigma, 30, -1.0)
            best = info.device_ms if best is None else min(best, info.device_ms)
        Q, w = s.download()
    row = {"pair_theta": th, "pair_theta3": th3, "ms": best, "pcg_iters": int(sum(info.cg_iters)), "late": info.cg_iters[-3:],
           "rms_vs_golden": O.geodesic_rms(Q, gold["Q"], g.f), "hit_max": info.cg_hit_max}
    rows.append(row)
    print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r02_sweep_theta.json"), "w"), indent=1)
