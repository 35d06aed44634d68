"""Pins for the oracle's l1ra / l1decode_pd restatement (ral/l1_irls.cpp:228-468, 851-912)."""
import os

import numpy as np

from oracle import graphs as G
from oracle import irls_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_l1decode_pd_solves_l1_regression():
    """On a small dense problem the interior-point iterate approaches the LP optimum of min |Ax - y|_1
    (checked against scipy's linprog), which pins the restated update formulas."""
    import scipy.sparse as sp
    from scipy.optimize import linprog
    g = G.small_graph(n=25, extra=60, sigma_n=0.05, outlier_frac=0.2, seed=1)
    A = O.make_A(g.n, g.f, g.I)
    Ah = O.make_A_noquirk(g.n, g.f, g.I)
    y = O.log_map(O.delta_rel(g.I, g.QQ, g.Q0))[:, 0]
    x = O.l1decode_pd(np.zeros(g.n - g.f), A.tocsr(), y, 40, Ah.tocsr())
    m, n = A.shape
    c = np.concatenate([np.zeros(n), np.ones(m)])
    Aub = sp.vstack([sp.hstack([A, -sp.eye(m)]), sp.hstack([-A, -sp.eye(m)])]).tocsc()
    res = linprog(c, A_ub=Aub, b_ub=np.concatenate([y, -y]), bounds=[(None, None)] * n + [(0, None)] * m, method="highs")
    assert res.status == 0
    assert np.abs(A @ x - y).sum() <= res.fun * (1 + 1e-3) + 1e-9


def test_newton_formulations_agree_and_quirk_pattern():
    g = G.small_graph(n=120, extra=700, sigma_n=0.03, outlier_frac=0.1, sigma_init=0.3, seed=5, f=4, fixed_anywhere=True)
    a = O.l1ra(g.QQ, g.I, None, g.Q0, g.f, 6, 1e-3, newton="direct")
    b = O.l1ra(g.QQ, g.I, None, g.Q0, g.f, 6, 1e-3, newton="pcg")
    assert a.iters == b.iters and O.geodesic_rms(a.Q, b.Q, g.f) < 1e-11
    # make_AtA keeps (free i, fixed j) edges on the diagonal, make_A drops them
    dropped = (g.I[:, 1] < g.f) & (g.I[:, 0] >= g.f)
    assert dropped.any()
    assert O.make_A(g.n, g.f, g.I)[np.nonzero(dropped)[0]].nnz == 0
    assert O.make_A_noquirk(g.n, g.f, g.I)[np.nonzero(dropped)[0]].nnz == dropped.sum()


def test_l1ra_converges_towards_ground_truth():
    g = G.small_graph(n=150, extra=900, sigma_n=0.0, outlier_frac=0.0, sigma_init=0.3, seed=9)
    r = O.l1ra(g.QQ, g.I, None, g.Q0, g.f, 25, 1e-6)
    assert np.all(np.diff(r.scores) < 0)                                       # two damped Newton steps per outer
    assert O.geodesic_rms(r.Q, g.Qgt, g.f) < 0.01 * O.geodesic_rms(g.Q0, g.Qgt, g.f)   # iteration: slow but monotone
    g = G.small_graph(n=150, extra=900, sigma_n=0.0, outlier_frac=0.15, sigma_init=0.3, seed=9)
    r = O.l1ra(g.QQ, g.I, None, g.Q0, g.f, 25, 1e-6)
    assert r.scores[-1] < 1e-6 and O.geodesic_rms(r.Q, g.Qgt, g.f) < 0.2 * O.geodesic_rms(g.Q0, g.Qgt, g.f)
    assert all(len(t) <= 2 for t in r.trace)                                    # l1_step stays 2 (App. A.6.7)


def test_golden_cli_flow():
    z = np.load(os.path.join(GOLD, "bundled_graph.npz"))
    I, QQ, Qm, f = z["I"], z["QQ"], z["Q_mst"], int(z["f"])
    la = O.l1ra(QQ, I, None, Qm, f, 5, 1e-3)
    assert la.iters == int(z["l1ra_iters"]) == 1                                # survey probe: 1 outer iteration
    assert np.allclose(la.scores, z["l1ra_scores"], rtol=1e-9) and abs(la.scores[0] - 9.41e-4) < 1e-6
    assert O.geodesic_rms(la.Q, z["l1ra_Q"], f) < 1e-12
    r = O.irls(QQ, I, None, O.GEMAN_MCCLURE, 5 * np.pi / 180, la.Q, f, 50, 1e-3)
    assert r.iters == int(z["cli_irls_iters"])
    assert O.geodesic_rms(O.quat_normalised(r.Q.copy(), f), z["cli_Q"], f) < 1e-12
