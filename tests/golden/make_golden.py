"""Writes the committed golden fixtures under tests/golden/.  Run in the build container only
(it reads the reference's bundled input graph, which does not exist on the GPU box):

    python tests/golden/make_golden.py

bundled_graph.npz: the parsed reference fixture ral/data/ravg_input.txt (m=3655, n=1832, f=1) as
  ral/test.cpp:161-247 reads it, its init_mst start (ral/test.cpp:285-286), and the oracle's
  irls() outputs (ral/test.cpp:300 with the CLI defaults: 50 iterations max, change_th 1e-3,
  sigma 5 deg) for L2 / L1 / Geman-McClure / Huber, plus a 10-iteration change_th=-1 L1 run, plus the
  CLI's whole default flow init_mst -> l1ra(5 iterations, 1e-3) -> irls(Geman-McClure) -> quat_normalised.
small_costs.npz: a 60-node graph with outliers, f=3, edges in both orientations (exercises
  make_A's dropped-edge rule), oracle outputs after 6 iterations for all 14 costs (sigma 5 deg;
  20 deg for Talwar, whose hard threshold at 5 deg makes the system singular on this graph).
The reference ships no expected outputs (parity unpinned, see oracle/irls_oracle.py); these
goldens pin the ORACLE so that later edits to it cannot drift silently, and give the GPU tests
fixed vectors to hit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import graphs as G  # noqa: E402
from oracle import irls_oracle as O  # noqa: E402

REF_INPUT = "/root/reference/ral/data/ravg_input.txt"
SIGMA = 5 * np.pi / 180.0


def bundled():
    I, QQ, Q, f, n_given = G.read_ral_text(REF_INPUT)
    Qmst = O.init_mst(Q, QQ, I, max(n_given, f))
    out = dict(I=I, QQ=QQ, Q_file=Q, f=np.int32(f), n_given=np.int32(n_given), Q_mst=Qmst)
    for cost in (O.L2, O.L1, O.GEMAN_MCCLURE, O.HUBER):
        r = O.irls(QQ, I, None, cost, SIGMA, Qmst, f, 50, 1e-3, solver="direct")
        tag = f"c{cost}"
        out[tag + "_Q"] = r.Q
        out[tag + "_weights"] = r.weights
        out[tag + "_scores"] = np.array(r.scores)
        out[tag + "_iters"] = np.int32(r.iters)
    r = O.irls(QQ, I, None, O.L1, SIGMA, Qmst, f, 10, -1.0, solver="direct")
    out["l1x10_Q"] = r.Q
    out["l1x10_weights"] = r.weights
    out["l1x10_scores"] = np.array(r.scores)
    # the CLI's full flow with its defaults (ral/test.cpp:285-302): init_mst -> l1ra(5, 1e-3) -> irls(GM, 5 deg, 50, 1e-3)
    la = O.l1ra(QQ, I, None, Qmst, f, 5, 1e-3)
    out["l1ra_Q"] = la.Q
    out["l1ra_scores"] = np.array(la.scores)
    out["l1ra_iters"] = np.int32(la.iters)
    r = O.irls(QQ, I, None, O.GEMAN_MCCLURE, SIGMA, la.Q, f, 50, 1e-3, solver="direct")
    out["cli_Q"] = O.quat_normalised(r.Q.copy(), f)
    out["cli_weights"] = r.weights
    out["cli_irls_iters"] = np.int32(r.iters)
    np.savez_compressed(os.path.join(HERE, "bundled_graph.npz"), **out)
    print("bundled:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.endswith("scores")})


def small_costs():
    g = G.small_graph(n=60, extra=300, sigma_n=0.03, outlier_frac=0.1, sigma_init=0.3, seed=11, f=3,
                      fixed_anywhere=True)
    out = dict(I=g.I, QQ=g.QQ, Q0=g.Q0, Qgt=g.Qgt, f=np.int32(g.f))
    for cost in range(14):
        sg = 4 * SIGMA if cost == O.TALWAR else SIGMA   # Talwar at 5 deg makes A^T D^2 A singular here
        r = O.irls(g.QQ, g.I, None, cost, sg, g.Q0, g.f, 6, -1.0, solver="lstsq")
        out[f"c{cost}_Q"] = r.Q
        out[f"c{cost}_weights"] = r.weights
        out[f"c{cost}_scores"] = np.array(r.scores)
    np.savez_compressed(os.path.join(HERE, "small_costs.npz"), **out)
    print("small_costs: done")


if __name__ == "__main__":
    bundled()
    small_costs()
