#!/usr/bin/env python
"""Runs T IRLS iterations on config 3 on the GPU and dumps (Q, weights) for offline preconditioner studies."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import irotavg_b200 as ira
from oracle import graphs as G
T = int(sys.argv[1]) if len(sys.argv) > 1 else 30
g = G.random_graph()
with ira.Solver() as s:
    Q, w, info = s.irls(g.QQ, g.I, None, ira.L1, 5 * np.pi / 180, g.Q0, g.f, T, -1.0)
print(info.cg_iters)
np.savez_compressed(f"gpurun_out/state_T{T}.npz", Q=Q, weights=w.astype(np.float64))
