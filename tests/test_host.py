"""Host helpers of the C ABI against the oracle: make_A's index rule and quat_normalised."""
import numpy as np
import pytest

from oracle import graphs as G
from oracle import irls_oracle as O


@pytest.mark.parametrize("f", [1, 2, 7])
def test_make_A_matches_oracle(built_lib, f):
    import irotavg_b200 as ira
    g = G.small_graph(n=40, extra=120, seed=f, f=f, fixed_anywhere=True)
    A = ira.make_A(g.n, f, g.I)
    B = O.make_A(g.n, f, g.I)
    assert A.shape == B.shape == (g.m, g.n - f)
    assert (A != B).nnz == 0
    dropped = (g.I[:, 1] < f) & (g.I[:, 0] >= f)
    assert dropped.any(), "fixture must contain (free i, fixed j) edges"
    assert A[np.nonzero(dropped)[0]].nnz == 0          # ral/l1_irls.cpp:770-771


def test_make_A_rejects_bad_index(built_lib):
    import irotavg_b200 as ira
    with pytest.raises(ira.IraError):
        ira.make_A(3, 1, [[0, 3]])


def test_quat_normalised(built_lib):
    import irotavg_b200 as ira
    rng = np.random.default_rng(0)
    Q = rng.standard_normal((20, 4)) * 3
    out = ira.quat_normalised(Q, 4)
    ref = O.quat_normalised(Q.copy(), 4)
    assert np.allclose(out, ref, atol=1e-15)
    assert np.array_equal(out[:4], Q[:4])             # fixed rows untouched
