"""Device l1ra (ira_l1ra, ira_l1ra.cuh) against the oracle's restatement of ral/l1_irls.cpp:228-468,851-912.
Tolerance: geodesic RMS <= 1e-8 rad, outer scores to 1e-6 relative (Newton systems solved by PCG to 1e-10)."""
import os

import numpy as np
import pytest

from oracle import graphs as G
from oracle import irls_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
SIGMA = 5 * np.pi / 180.0


@pytest.mark.parametrize("maker", ["small", "quirk", "kitti_small", "large_angle"])
def test_l1ra_vs_oracle(solver, maker):
    if maker == "small":
        g = G.small_graph(n=300, extra=2500, sigma_n=0.03, outlier_frac=0.1, sigma_init=0.3, seed=21)
    elif maker == "quirk":
        g = G.small_graph(n=200, extra=1500, sigma_n=0.03, outlier_frac=0.1, sigma_init=0.3, seed=22, f=5, fixed_anywhere=True)
    elif maker == "kitti_small":
        g = G.kitti_like_graph(n=800, m=8500)
    else:
        g = G.small_graph(n=150, extra=800, sigma_n=0.01, seed=23)
        g.Q0[80:] = np.array([0, 0, 0, 1.0])
    ref = O.l1ra(g.QQ, g.I, None, g.Q0, g.f, 6, 1e-3)
    Q, info = solver.l1ra(g.QQ, g.I, None, g.Q0, g.f, 6, 1e-3)
    assert info.iters == ref.iters
    assert info.cg_hit_max == 0
    assert np.allclose(info.scores, ref.scores, rtol=1e-6, atol=1e-12)
    assert O.geodesic_rms(Q, ref.Q, g.f) <= 1e-8
    assert np.array_equal(Q[:g.f], g.Q0[:g.f])


def test_cli_flow_bundled_graph(solver):
    """Config 1 end to end as ral/test.cpp:285-302 runs it: (init_mst start) -> l1ra(5, 1e-3) -> irls(GM, 5 deg,
    50, 1e-3) -> quat_normalised, all on one uploaded graph (resident_start(True) chains the two stages)."""
    import irotavg_b200 as ira
    z = np.load(os.path.join(GOLD, "bundled_graph.npz"))
    I, QQ, Qm, f = z["I"], z["QQ"], z["Q_mst"], int(z["f"])
    Q1, info = solver.l1ra(QQ, I, None, Qm, f, 5, 1e-3)
    assert info.iters == int(z["l1ra_iters"])
    assert np.allclose(info.scores, z["l1ra_scores"], rtol=1e-6)
    assert O.geodesic_rms(Q1, z["l1ra_Q"], f) <= 1e-8
    solver.upload(QQ, I, Qm, f)
    la = solver.l1ra_resident(5, 1e-3)
    solver.resident_start(True)
    ir = solver.irls_resident(O.GEMAN_MCCLURE, SIGMA, 50, 1e-3)
    solver.resident_start(False)
    Q, w = solver.download()
    Q = ira.quat_normalised(Q, f)
    assert la.iters == int(z["l1ra_iters"]) and ir.iters == int(z["cli_irls_iters"])
    assert O.geodesic_rms(Q, z["cli_Q"], f) <= 1e-8
    assert np.allclose(w, z["cli_weights"], rtol=1e-5, atol=1e-8)


def test_l1ra_config2(solver):
    g = G.kitti_like_graph()
    ref = O.l1ra(g.QQ, g.I, None, g.Q0, g.f, 3, 1e-3, newton="pcg", pcg_rtol=1e-13)
    Q, info = solver.l1ra(g.QQ, g.I, None, g.Q0, g.f, 3, 1e-3)
    assert info.iters == ref.iters == 3 and info.cg_hit_max == 0
    assert np.allclose(info.scores, ref.scores, rtol=1e-6)
    assert O.geodesic_rms(Q, ref.Q, g.f) <= 1e-8


def test_l1ra_zero_iterations_and_errors(solver):
    import irotavg_b200 as ira
    g = G.small_graph(n=50, extra=200, sigma_n=0.01, seed=3)
    Q, info = solver.l1ra(g.QQ, g.I, None, g.Q0, g.f, 0, 1e-3)
    assert info.iters == 0 and np.array_equal(Q, g.Q0)
    with pytest.raises(ira.IraError):
        solver.l1ra(g.QQ, g.I, None, g.Q0, 0, 3, 1e-3)


def test_l1ra_on_a_hub_graph(built_lib):
    """A star-like view graph (one hub tied to every node): the SELL padding blow-up makes irls fall back to its
    CSR kernels; l1ra must still run (it used to return IRA_ERR_INVALID_ARG) and match the oracle."""
    import irotavg_b200 as ira
    rng = np.random.default_rng(9)
    n = 6000
    base = G.small_graph(n=n, extra=1500, sigma_n=0.01, sigma_init=0.05, seed=9)
    hub = np.stack([np.full(n - 2, 1), np.arange(2, n)], axis=1).astype(np.int32)       # node 1 sees everybody
    I = np.concatenate([base.I, hub])
    Qgt = base.Qgt
    QQh = O.quat_mult(Qgt[hub[:, 1]], Qgt[hub[:, 0]] * np.array([-1.0, -1.0, -1.0, 1.0]))
    QQ = np.concatenate([base.QQ, QQh])
    for lanes in (0, 8):                                                                # 8: irls on the CSR kernels by request
        with ira.Solver(lanes_per_row=lanes) as s:
            Q, info = s.l1ra(QQ, I, None, base.Q0, base.f, 3, 1e-6)
            Q2, w2, info2 = s.irls(QQ, I, None, O.GEMAN_MCCLURE, SIGMA, Q, base.f, 5, -1.0)
        ref = O.l1ra(QQ, I, None, base.Q0, base.f, 3, 1e-6)
        assert info.iters == ref.iters
        assert O.geodesic_rms(Q, ref.Q, base.f) <= 1e-8
        ref2 = O.irls(QQ, I, None, O.GEMAN_MCCLURE, SIGMA, ref.Q, base.f, 5, -1.0, solver="direct")
        assert O.geodesic_rms(Q2, ref2.Q, base.f) <= 1e-8
