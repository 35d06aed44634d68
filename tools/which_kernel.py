import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import irotavg_b200 as ira
from oracle import graphs as G
g = G.random_graph()
for sv in (0, 64, 32):      # default (balanced slice deal), round-robin deal, matrix in shared memory
    with ira.Solver(solver=sv) as s:
        s.upload(g.QQ, g.I, g.Q0, g.f)
        info = s.irls_resident(1, 5*np.pi/180, 30, -1.0)
        info = s.irls_resident(1, 5*np.pi/180, 30, -1.0)
        Q, w = s.download()
        print("solver", sv, "pcg_kernel", info.pcg_kernel, "ms", info.device_ms, "its", sum(info.cg_iters), "Qsum", repr(float(Q.sum())), flush=True)
