#!/usr/bin/env python
"""A small persistent solve for compute-sanitizer (memcheck / racecheck / synccheck): 4 L1 iterations on a 3 000-node
graph (k_pcg_persistent_reg_mw), on a 30 000-node graph (k_pcg_smem) and an init_mst, so that every cooperative kernel
of the single-GPU path runs once under the tool.   compute-sanitizer --tool memcheck python tools/sanitize_step.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import irotavg_b200 as ira  # noqa: E402
from oracle import graphs as G  # noqa: E402

sigma = 5 * np.pi / 180
big = int(os.environ.get("SAN_N", "30000"))
for n, extra in ((3000, 30000), (big, big * 9)):
    g = G.small_graph(n=n, extra=extra, sigma_n=0.03, outlier_frac=0.1, seed=31, f=2, fixed_anywhere=True)
    with ira.Solver() as s:
        Q, w, info = s.irls(g.QQ, g.I, None, ira.L1, sigma, g.Q0, g.f, 4, -1.0)
        Q0 = np.zeros_like(g.Q0); Q0[:g.f] = g.Q0[:g.f]
        Qm, st = s.init_mst(Q0, g.QQ, g.I, g.f)
    print(f"n={n}: irls iters {info.iters}, PCG {info.cg_iters}, init_mst passes {st['passes_label']}/{st['passes_propagate']}", flush=True)
gw = G.small_graph(n=15, extra=27, sigma_n=0.005, sigma_init=0.05, seed=1, f=4)
with ira.Solver() as s:
    Q, w, l1, info = s.l1ra_irls(gw.QQ, gw.I, gw.Q0, gw.f, 100, 1e-3, ira.Geman_McClure, sigma, 100, 1e-3)
    print("window:", l1, info.iters, flush=True)
    g = G.small_graph(n=800, extra=6000, sigma_n=0.03, outlier_frac=0.1, seed=3)
    Q, info = s.l1ra(g.QQ, g.I, None, g.Q0, g.f, 2, 1e-3)
    print("l1ra:", info.iters, flush=True)
