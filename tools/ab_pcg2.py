#!/usr/bin/env python
"""A/B on one box: the matrix-in-shared-memory PCG kernel (ira_pcg2.cuh, default) against the register-resident
kernel it replaces (solver=16), config 3 and config 2, L1 and Geman-McClure; then a cg_rtol sweep of the default
kernel against the committed 30-iteration golden (tests/golden/cfg3_l1_30iters.npz)."""
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
out = {"ab": [], "rtol_sweep": []}
g3 = G.random_graph()
for name, g in (("config 3", g3), ("config 2", G.kitti_like_graph())):
    for cost, cname in ((1, "L1"), (4, "Geman-McClure")):
        res = {}
        for sv in (0, 16):
            with ira.Solver(solver=sv) as s:
                s.upload(g.QQ, g.I, g.Q0, g.f)
                s.irls_resident(cost, sigma, 30, -1.0)
                best = None
                for _ in range(3):
                    info = s.irls_resident(cost, sigma, 30, -1.0)
                    best = info.device_ms if best is None else min(best, info.device_ms)
                ph = info.profile.get("pcg_phases") or {}
                Q, w = s.download()
                res[sv] = dict(ms=best, its=int(sum(info.cg_iters)), kernel_ms=ph.get("kernel_ms"), spmv_ms=ph.get("spmv_ms"),
                               update_ms=ph.get("update_ms"), Q=Q)
        d = O.geodesic_rms(res[0]["Q"], res[16]["Q"], g.f)
        row = {"graph": name, "cost": cname, "rms_between": d}
        for sv in (0, 16):
            r = res[sv]
            row[f"solver{sv}"] = {k: r[k] for k in ("ms", "its", "kernel_ms", "spmv_ms", "update_ms")}
            row[f"solver{sv}"]["us_per_pcg_iter"] = 1e3 * r["kernel_ms"] / max(1, r["its"]) if r["kernel_ms"] else None
        out["ab"].append(row)
        print(json.dumps(row), flush=True)
gold = np.load(os.path.join(ROOT, "tests", "golden", "cfg3_l1_30iters.npz"))
for rtol in (1e-10, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5):
    with ira.Solver(cg_rtol=rtol) as s:
        s.upload(g3.QQ, g3.I, g3.Q0, g3.f)
        s.irls_resident(1, sigma, 30, -1.0)
        info = s.irls_resident(1, sigma, 30, -1.0)
        Q, w = s.download()
    row = {"cg_rtol": rtol, "ms": info.device_ms, "its": int(sum(info.cg_iters)),
           "rms_vs_golden_rad": O.geodesic_rms(Q, gold["Q"], g3.f),
           "max_score_rel_dev": float(np.abs(np.array(info.scores) / gold["scores"] - 1).max()),
           "weights_max_rel_dev_every97": float(np.abs(w[::97] / gold["weights_every97"] - 1).max())}
    out["rtol_sweep"].append(row)
    print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "r02_ab_pcg2.json"), "w") as fh:
    json.dump(out, fh, indent=1)
