#!/usr/bin/env python
"""One resident irls() step on a synthetic graph - the command profiled under ncu.

    python tools/profile_step.py [--cost L1] [--iters 4] [--n 100000] [--m 1000000] [--reps 1]
Prints the per-kernel-class device times (CUDA events on the launch stream; profile=1).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import irotavg_b200 as ira  # noqa: E402
from oracle import graphs as G  # noqa: E402

COSTS = {"L2": 0, "L1": 1, "Geman-McClure": 4, "Huber": 5}

ap = argparse.ArgumentParser()
ap.add_argument("--cost", default="L1")
ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--n", type=int, default=100_000)
ap.add_argument("--m", type=int, default=1_000_000)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--kitti", action="store_true")
ap.add_argument("--no-profile", action="store_true")
ap.add_argument("--lpr", type=int, default=0)
ap.add_argument("--solver", type=int, default=0)
ap.add_argument("--check-every", type=int, default=None)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--time-kernels", action="store_true")
ap.add_argument("--debug-phases", action="store_true", help="ira_options.profile = 2: the small-graph PCG kernel prints its phase split")
a = ap.parse_args()

g = G.kitti_like_graph() if a.kitti else G.random_graph(n=a.n, m=a.m)
s = ira.Solver(profile=2 if a.debug_phases else (not a.no_profile), lanes_per_row=a.lpr, solver=a.solver, cg_check_every=a.check_every,
               spmv_variant=a.variant)
s.upload(g.QQ, g.I, g.Q0, g.f)
for _ in range(a.reps):
    info = s.irls_resident(COSTS[a.cost], 5 * np.pi / 180, a.iters, -1.0)
out = {"graph": g.name, "cost": a.cost, "iters": info.iters, "device_ms": info.device_ms,
       "launches": info.kernel_launches, "cg_iters": info.cg_iters, "cg_total": sum(info.cg_iters),
       "scores": info.scores[-3:], "cg_hit_max": info.cg_hit_max}
for k, v in info.profile.items():
    if "launches" in v:
        out[k] = {"ms": round(v["ms"], 4), "launches": v["launches"], "us_per_launch": round(1000 * v["ms"] / v["launches"], 3)}
    else:
        out[k] = {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()}
if a.time_kernels:
    out["kernel_us_warm"] = {nm: round(s.time_kernel(k, 50, False), 3) for k, nm in enumerate(["residual", "spmv", "rhs", "weights", "update", "cg_update"])}
    out["kernel_us_cold"] = {nm: round(s.time_kernel(k, 20, True), 3) for k, nm in enumerate(["residual", "spmv", "rhs", "weights", "update", "cg_update"])}
print(json.dumps(out))
