"""The single-block window solver (ira_small.cuh: l1ra + irls in one launch, dense Cholesky solves) against the
oracle (direct solves) and against the general pipeline (PCG) on the same inputs.  Tolerances: geodesic RMS
<= 1e-9 rad vs the oracle (both solve exactly), identical iteration counts, weights to 1e-6 relative."""
import numpy as np
import pytest

from oracle import graphs as G
from oracle import irls_oracle as O

pytestmark = pytest.mark.gpu
SIGMA = 5 * np.pi / 180.0


def _ref(g, l1_iters, irls_iters, cost, th=1e-3):
    la = O.l1ra(g.QQ, g.I, None, g.Q0, g.f, l1_iters, th)
    r = O.irls(g.QQ, g.I, None, cost, SIGMA, la.Q, g.f, irls_iters, th, solver="direct")
    return la, r


CASES = {
    "window": dict(n=15, extra=26, sigma_n=0.005, sigma_init=0.05, seed=1, f=4),
    "outliers": dict(n=30, extra=150, sigma_n=0.03, outlier_frac=0.15, sigma_init=0.3, seed=2, f=1),
    "quirk": dict(n=40, extra=200, sigma_n=0.03, outlier_frac=0.1, sigma_init=0.3, seed=3, f=9, fixed_anywhere=True),
    "limit": dict(n=64, extra=192, sigma_n=0.02, outlier_frac=0.05, sigma_init=0.2, seed=4, f=32),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("cost", [O.GEMAN_MCCLURE, O.L1, O.HUBER])
def test_small_vs_oracle(solver, name, cost):
    g = G.small_graph(**CASES[name])
    assert g.n <= 64 and g.n - g.f <= 32 and g.m <= 256
    la, r = _ref(g, 100, 100, cost)
    Q, w, l1_it, info = solver.l1ra_irls(g.QQ, g.I, g.Q0, g.f, 100, 1e-3, cost, SIGMA, 100, 1e-3)
    assert info.kernel_launches == 1                       # the single-block path was taken
    assert l1_it == la.iters and info.iters == r.iters
    assert O.geodesic_rms(Q, r.Q, g.f) <= 1e-9
    assert np.allclose(w, r.weights, rtol=1e-6, atol=1e-9)
    assert np.array_equal(Q[:g.f], g.Q0[:g.f])
    k = min(len(r.scores), 8)
    assert np.allclose(info.scores[:k], r.scores[:k], rtol=1e-6, atol=1e-13)


def test_small_equals_general_pipeline(solver):
    import irotavg_b200 as ira
    g = G.small_graph(**CASES["outliers"])
    Qs, ws, l1s, infos = solver.l1ra_irls(g.QQ, g.I, g.Q0, g.f, 100, 1e-3, O.GEMAN_MCCLURE, SIGMA, 100, 1e-3)
    with ira.Solver(small_path=False) as big:
        Qb, wb, l1b, infob = big.l1ra_irls(g.QQ, g.I, g.Q0, g.f, 100, 1e-3, O.GEMAN_MCCLURE, SIGMA, 100, 1e-3)
    assert infob.kernel_launches > 1
    assert (l1s, infos.iters) == (l1b, infob.iters)
    assert O.geodesic_rms(Qs, Qb, g.f) <= 1e-8
    assert np.allclose(ws, wb, rtol=1e-5, atol=1e-8)


def test_small_is_bitwise_repeatable(solver):
    g = G.small_graph(**CASES["quirk"])
    a = solver.l1ra_irls(g.QQ, g.I, g.Q0, g.f, 100, 1e-3, O.L1, SIGMA, 30, -1.0)
    b = solver.l1ra_irls(g.QQ, g.I, g.Q0, g.f, 100, 1e-3, O.L1, SIGMA, 30, -1.0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_zero_iteration_caps(solver):
    """max_iters = 0 on both stages returns Q untouched and weights = 1 (weights.setOnes(), ral/l1_irls.cpp:577)."""
    g = G.small_graph(**CASES["window"])
    Q, w, l1_it, info = solver.l1ra_irls(g.QQ, g.I, g.Q0, g.f, 0, 1e-3, O.GEMAN_MCCLURE, SIGMA, 0, 1e-3)
    assert l1_it == 0 and info.iters == 0
    assert np.array_equal(Q, g.Q0) and np.all(w == 1.0)
