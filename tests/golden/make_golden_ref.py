"""Golden vectors produced by the REFERENCE ITSELF (oracle/_ref: ral/l1_irls.cpp and ral/test.cpp compiled
unmodified from /root/reference by oracle/build_ref.py against the stand-in Eigen / SuiteSparse headers in
oracle/ref_shim/).  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_ref.py

ref_bundled_cli.npz   the reference CLI `l1_irls` on its own fixture ral/data/ravg_input.txt: default arguments
                      (init_mst -> l1ra(5, 1e-3) -> irls(Geman-McClure, 5 deg, 50, 1e-3) -> quat_normalised,
                      ral/test.cpp:250-302), and the same with cost L1, Huber, L2; Q as [x y z w], weights.
ref_small.npz         through the C entry points over irotavg::* on seeded synthetic graphs: all 14 costs (6 IRLS
                      iterations, f = 3, edges in both orientations -> make_A's dropped-edge rule), Talwar at 5 deg
                      (rank-deficient: every edge of two free nodes gets weight 0), make_A itself,
                      log_map(delta_rel) per edge incl. the wrap / s < EPS rows, exp_map incl. the NaN -> 0 row,
                      init_mst (forward and backward tree edges), l1ra alone, l1ra -> irls on a banded view graph,
                      an identity-start (large-angle) window.
tests/test_ref_pin.py holds the oracle (CPU) and the CUDA path (GPU) to these numbers.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import graphs as G  # noqa: E402
from oracle import refbin as R  # noqa: E402

REF_INPUT = "/root/reference/ral/data/ravg_input.txt"
SIGMA = 5 * np.pi / 180.0


def bundled_cli():
    out = {}
    n, m = 1832, 3655
    with tempfile.TemporaryDirectory() as td:
        for tag, extra in (("default", []), ("l1", ["L1"]), ("huber", ["Huber"]), ("l2", ["L2"])):
            path = os.path.join(td, f"out_{tag}.txt")
            res = R.cli([REF_INPUT, path] + extra)
            assert res.returncode == 0, res.stderr
            Q, w = R.read_cli_output(path, n, m)
            out[f"{tag}_Q"] = Q
            out[f"{tag}_weights"] = w
            out[f"{tag}_stdout_tail"] = np.array(res.stdout.splitlines()[-12:-3])
    np.savez_compressed(os.path.join(HERE, "ref_bundled_cli.npz"), **out)
    print("ref_bundled_cli:", sorted(out))


def graphs():
    gq = G.small_graph(n=60, extra=300, sigma_n=0.03, outlier_frac=0.1, sigma_init=0.3, seed=11, f=3, fixed_anywhere=True)
    gk = G.banded_graph()
    gw = G.small_graph(n=15, extra=27, sigma_n=0.005, sigma_init=0.05, seed=1, f=4)
    gi = G.small_graph(n=120, extra=500, sigma_n=0.02, seed=3)
    gi.Q0[gi.f:] = np.array([0, 0, 0, 1.0])      # every view starts at identity: full-range residuals
    gi.QQ[::5] *= -1.0                            # q / -q ambiguity of the measurements
    return gq, gk, gw, gi


def small():
    gq, gk, gw, gi = graphs()
    out = {}
    for cost in range(14):
        sg = 4 * SIGMA if cost == 12 else SIGMA   # Talwar at 5 deg zeroes every edge of some node on this graph
        Q, w, it = R.irls(gq.QQ, gq.I, cost, sg, gq.Q0, gq.f, 6, -1.0)
        out[f"quirk_c{cost}_Q"] = Q; out[f"quirk_c{cost}_weights"] = w; out[f"quirk_c{cost}_iters"] = np.int32(it)
    # Talwar at 5 deg zeroes every edge of two free nodes here: D A loses rank, SuiteSparseQR's rank detection declares
    # their columns dead and returns x = 0 for them (the basic solution, ral/l1_irls.cpp:550)
    Q, w, it = R.irls(gq.QQ, gq.I, 12, SIGMA, gq.Q0, gq.f, 6, -1.0)
    out["quirk_talwar5_Q"] = Q; out["quirk_talwar5_weights"] = w
    out["quirk_A"] = R.make_A(gq.n, gq.f, gq.I)
    out["quirk_residual"] = R.residual(gq.I, gq.QQ, gq.Q0)
    out["ident_residual"] = R.residual(gi.I, gi.QQ, gi.Q0)
    W = np.zeros((6, 4)); W[1, :3] = [0.3, -0.2, 0.1]; W[2, :3] = [1e-9, 0, 0]; W[3, :3] = [3.0, 0.5, -1.0]; W[4, :3] = [0, 0, 1e-300]
    out["expmap_in"] = W; out["expmap_out"] = R.exp_map(W)
    out["quirk_mst"] = R.init_mst(gq.Q0, gq.QQ, gq.I, gq.f)
    perm = np.random.default_rng(4).permutation(gk.m)
    out["kitti_perm"] = perm
    out["kitti_mst_shuffled"] = R.init_mst(gk.Q0, gk.QQ[perm], gk.I[perm], gk.f)
    Q, it = R.l1ra(gq.QQ, gq.I, gq.Q0, gq.f, 5, 1e-3)
    out["quirk_l1ra_Q"] = Q; out["quirk_l1ra_iters"] = np.int32(it)
    Q, it = R.l1ra(gk.QQ, gk.I, gk.Q0, gk.f, 5, 1e-3)
    out["kitti_l1ra_Q"] = Q; out["kitti_l1ra_iters"] = np.int32(it)
    Q2, w2, it2 = R.irls(gk.QQ, gk.I, 4, SIGMA, Q, gk.f, 50, 1e-3)
    out["kitti_flow_Q"] = R.quat_normalised(Q2, gk.f); out["kitti_flow_weights"] = w2; out["kitti_flow_iters"] = np.int32(it2)
    Q, w, it = R.irls(gk.QQ, gk.I, 1, SIGMA, gk.Q0, gk.f, 30, -1.0)
    out["kitti_l1x30_Q"] = Q; out["kitti_l1x30_weights"] = w
    Q, it = R.l1ra(gw.QQ, gw.I, gw.Q0, gw.f, 100, 1e-3)
    Q2, w2, it2 = R.irls(gw.QQ, gw.I, 4, SIGMA, Q, gw.f, 100, 1e-3)
    out["window_Q"] = Q2; out["window_weights"] = w2; out["window_iters"] = np.array([it, it2], dtype=np.int32)
    Q, w, it = R.irls(gi.QQ, gi.I, 4, SIGMA, gi.Q0, gi.f, 20, 1e-3)
    out["ident_gm_Q"] = Q; out["ident_gm_weights"] = w; out["ident_gm_iters"] = np.int32(it)
    np.savez_compressed(os.path.join(HERE, "ref_small.npz"), **out)
    print("ref_small:", len(out), "arrays")


if __name__ == "__main__":
    small()
    bundled_cli()
