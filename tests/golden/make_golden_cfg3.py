"""Writes tests/golden/cfg3_l1_30iters.npz: the FULL 30-iteration L1 run of configs[2] (n = 100 000,
m = 1 000 000, the configuration BASELINE.json's metric is quoted on) by the C restatement
(oracle/irls_oracle.c: plain Jacobi-PCG, rtol 1e-12 - a different preconditioner, a tighter tolerance and
a different summation order than the CUDA path, so agreement is not an artefact of shared code).

    python tests/golden/make_golden_cfg3.py [--threads 6] [--cost L1] [--scale 1]

~250 000 CG iterations: 10-20 minutes on 8 host cores.  The fixture holds the per-iteration scores and
CG counts, the final rotations as float64 (3.2 MB) and a digest of the final weights (quantiles + every
97th value).  tests/test_gpu_parity.py::test_config3_30_iterations_vs_golden and bench.py
(`geodesic_rms_vs_oracle_30iters_rad`) compare the CUDA path with it.
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import cport  # noqa: E402
from oracle import graphs as G  # noqa: E402

SIGMA = 5 * np.pi / 180.0
COSTS = {"L1": 1, "Geman-McClure": 4, "Huber": 5}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=6)
    ap.add_argument("--cost", default="L1", choices=sorted(COSTS))
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--rtol", type=float, default=1e-12)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    g = G.random_graph(n=100_000 * a.scale, m=1_000_000 * a.scale)
    t0 = time.perf_counter()
    r = cport.irls(g.QQ, g.I, COSTS[a.cost], SIGMA, g.Q0, g.f, a.iters, -1.0, cg_rtol=a.rtol, cg_max_iters=2_000_000,
                   threads=a.threads)
    dt = time.perf_counter() - t0
    w = r["weights"]
    out = dict(Q=r["Q"], scores=np.array(r["scores"]), cg_iters=np.array(r["cg_iters"], dtype=np.int64),
               weights_every97=w[::97].copy(), weights_quantiles=np.quantile(w, [0, .01, .1, .25, .5, .75, .9, .99, 1.0]),
               weights_sum=np.float64(w.sum()), n=np.int64(g.n), m=np.int64(g.m), f=np.int64(g.f),
               cost=np.int64(COSTS[a.cost]), rtol=np.float64(a.rtol), seconds=np.float64(dt), threads=np.int64(a.threads))
    tag = {"L1": "l1", "Geman-McClure": "gm", "Huber": "huber"}[a.cost]
    name = a.out or os.path.join(HERE, f"cfg3_{tag}_{a.iters}iters" + (f"_x{a.scale}" if a.scale > 1 else "") + ".npz")
    np.savez(name, **out)
    print(f"wrote {name}: {dt:.1f} s, {int(sum(r['cg_iters']))} CG iterations, scores {r['scores'][0]:.3e} .. {r['scores'][-1]:.3e}")
