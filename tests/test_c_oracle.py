"""The independent plain-C restatement (oracle/irls_oracle.c) agrees with the numpy oracle."""
import numpy as np
import pytest

from oracle import graphs as G
from oracle import irls_oracle as O

SIGMA = 5 * np.pi / 180.0


@pytest.fixture(scope="module")
def cport():
    from oracle import cport as cp
    if not cp.available():
        pytest.fail("gcc could not build oracle/irls_oracle.c")
    return cp


@pytest.mark.parametrize("cost", range(14))
def test_c_port_matches_numpy_oracle(cport, cost):
    g = G.small_graph(n=80, extra=500, sigma_n=0.03, outlier_frac=0.1, sigma_init=0.3, seed=13, f=3, fixed_anywhere=True)
    sg = 4 * SIGMA if cost == O.TALWAR else SIGMA
    a = O.irls(g.QQ, g.I, None, cost, sg, g.Q0, g.f, 6, -1.0, solver="direct")
    for threads in (1, 4):
        b = cport.irls(g.QQ, g.I, cost, sg, g.Q0, g.f, 6, -1.0, threads=threads)
        assert b["iters"] == 6
        assert np.allclose(b["scores"], a.scores, rtol=1e-9, atol=1e-14)
        assert O.geodesic_rms(b["Q"], a.Q, g.f) < 1e-11
        assert np.allclose(b["weights"], a.weights, rtol=1e-7, atol=1e-9)


def test_c_port_stop_rule_and_bundled_golden(cport):
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "bundled_graph.npz"))
    r = cport.irls(z["QQ"], z["I"], O.L1, SIGMA, z["Q_mst"], int(z["f"]), 50, 1e-3)
    assert r["iters"] == int(z["c1_iters"])
    assert O.geodesic_rms(r["Q"], z["c1_Q"], 1) < 1e-10
    r0 = cport.irls(z["QQ"], z["I"], O.L1, SIGMA, z["Q_mst"], 1, 0, 1e-3)
    assert r0["iters"] == 0 and np.array_equal(r0["Q"], z["Q_mst"])
