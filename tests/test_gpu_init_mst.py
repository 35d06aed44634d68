"""Device init_mst (ira_init_mst, ira_mst.cuh) against the oracle's literal restatement of
ral/l1_irls.cpp:915-979.  The spanning tree depends on the edge order, so the check is on the rotations
themselves: every node must come out of the SAME single quaternion product as in the sequential sweeps
(max abs component difference <= 1e-12: only FMA contraction differs), including the reference's sign
convention for edges traversed backwards."""
import os

import numpy as np
import pytest

from oracle import graphs as G
from oracle import irls_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _check(solver, g_I, g_QQ, Q0, f):
    ref = O.init_mst(Q0, g_QQ, g_I, f)
    Q, st = solver.init_mst(Q0, g_QQ, g_I, f)
    assert st["unreached"] == 0
    assert np.max(np.abs(Q - ref)) <= 1e-12, np.max(np.abs(Q - ref))
    assert np.array_equal(Q[:f], np.asarray(Q0)[:f])
    return st


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_order(solver, seed):
    """Shuffled edge order and flipped orientations: many sweeps, both propagation directions."""
    g = G.small_graph(n=400, extra=1200, sigma_n=0.05, outlier_frac=0.1, sigma_init=0.3, seed=40 + seed)
    rng = np.random.default_rng(seed)
    perm = rng.permutation(g.m)
    I, QQ = g.I[perm].copy(), g.QQ[perm].copy()
    Q0 = np.zeros_like(g.Q0)
    Q0[0] = g.Q0[0]
    _check(solver, I, QQ, Q0, 1)


def test_reversed_path(solver):
    """A path listed backwards: the sequential sweeps flag one node per sweep (n-1 sweeps)."""
    n = 300
    g = G.small_graph(n=n, extra=0, sigma_n=0.02, seed=5)
    I, QQ = g.I[::-1].copy(), g.QQ[::-1].copy()
    Q0 = np.zeros((n, 4))
    Q0[0] = [0, 0, 0, 1]
    st = _check(solver, I, QQ, Q0, 1)
    assert st["passes_label"] >= 2


def test_given_rotations_are_kept(solver):
    """f_init > 1 (ral/test.cpp:285: max(#given, f)): given rows are not overwritten but still propagate."""
    g = G.small_graph(n=250, extra=900, sigma_n=0.05, sigma_init=0.4, seed=9)
    Q0 = g.Q0.copy()
    Q0[40:] = 0.0
    _check(solver, g.I, g.QQ, Q0, 40)


def test_bundled_graph(solver):
    """Config 1: the reference's only fixture, start of the CLI flow."""
    z = np.load(os.path.join(GOLD, "bundled_graph.npz"))
    I, QQ, f = z["I"], z["QQ"], int(z["f"])
    Q0 = np.zeros((int(I.max()) + 1, 4))
    Q0[:f] = z["Q_mst"][:f]
    Q, st = solver.init_mst(Q0, QQ, I, f)
    assert np.max(np.abs(Q - z["Q_mst"])) <= 1e-12


def test_not_spanning(solver):
    import irotavg_b200 as ira
    g = G.small_graph(n=50, extra=100, seed=3)
    keep = (g.I[:, 0] != 17) & (g.I[:, 1] != 17)            # isolate node 17
    Q0 = np.zeros((50, 4)); Q0[0] = [0, 0, 0, 1]
    with pytest.raises(ira.IraError) as ei:
        solver.init_mst(Q0, g.QQ[keep], g.I[keep], 1)
    assert ei.value.status == 8 and "DO NOT SPAN" in str(ei.value)


def test_config3_full_size(solver):
    """1M edges: the path edges come first, so the sweeps chain through all 100 000 nodes in one sweep."""
    g = G.random_graph()
    Q0 = np.zeros_like(g.Q0); Q0[0] = g.Q0[0]
    ref = O.init_mst(Q0, g.QQ, g.I, 1)
    Q, st = solver.init_mst(Q0, g.QQ, g.I, 1)
    # a 100 000-product chain: rounding differences (FMA) accumulate along the chain
    assert O.geodesic_rms(Q, ref, 1) <= 1e-10
