"""Pins for the CPU oracle (oracle/irls_oracle.py).  The reference ships no expected outputs, so
the oracle is held to: exact known answers on noise-free graphs, agreement of three independent
formulations of the linear step, algebraic identities of the maps, and the committed goldens."""
import os

import numpy as np
import pytest

from oracle import graphs as G
from oracle import irls_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SIGMA = 5 * np.pi / 180.0


def sigma_for(cost):
    """Talwar's hard threshold at 5 deg zeroes every edge of some nodes of the small test graphs and
    makes A^T D^2 A singular (outside the parity contract); 20 deg keeps it regular."""
    return 4 * SIGMA if cost == O.TALWAR else SIGMA


def test_quat_mult_is_hamilton_product():
    rng = np.random.default_rng(0)
    a, b, c = rng.standard_normal((3, 50, 4))
    # associativity, norm multiplicativity, identity, and i*j = k in [x y z w] layout
    assert np.allclose(O.quat_mult(O.quat_mult(a, b), c), O.quat_mult(a, O.quat_mult(b, c)), atol=1e-12)
    assert np.allclose(np.linalg.norm(O.quat_mult(a, b), axis=1),
                       np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))
    e = np.array([0.0, 0, 0, 1])
    assert np.allclose(O.quat_mult(a, e), a) and np.allclose(O.quat_mult(e, a), a)
    i, j, k = np.eye(4)[0], np.eye(4)[1], np.eye(4)[2]
    assert np.allclose(O.quat_mult(i, j), k) and np.allclose(O.quat_mult(j, i), -k)


def test_log_exp_roundtrip_and_wrap():
    rng = np.random.default_rng(1)
    v = rng.standard_normal((200, 3))
    v *= (rng.uniform(0, np.pi * 0.999, 200) / np.linalg.norm(v, axis=1))[:, None]
    W = np.concatenate([v, np.zeros((200, 1))], axis=1)
    q = O.exp_map(W.copy())
    assert np.allclose(np.linalg.norm(q, axis=1), 1.0)
    back = O.log_map(q.copy())
    assert np.allclose(back[:, :3], v, atol=1e-13)
    # -q is the same rotation: atan2 gives theta in (pi, 2pi], the wrap brings it to [-pi, 0)
    back2 = O.log_map(-q.copy())
    assert np.allclose(back2[:, :3], v, atol=1e-12)
    assert np.all(back2[:, 3] < 0) and np.all(back2[:, 3] >= -np.pi)
    # identity / zero vector part: s < EPS -> 0, exp of 0 -> identity (NaN -> 0 rule)
    z = O.log_map(np.array([[0.0, 0, 0, 1.0], [0, 0, 0, -1.0]]))
    assert np.all(z[:, :3] == 0)
    assert np.array_equal(O.exp_map(np.zeros((1, 4))), np.array([[0.0, 0, 0, 1.0]]))


def test_residual_equals_rotation_matrix_log():
    """p = q~_j (x) QQ (x) Q_i is the rotation R_j^T R_ij R_i."""
    g = G.small_graph(n=40, extra=100, sigma_n=0.3, sigma_init=0.5, seed=3)
    w = O.log_map(O.delta_rel(g.I, g.QQ, g.Q0))
    for k in range(0, g.m, 7):
        i, j = g.I[k]
        R = O.quat2rmat(g.Q0[j]).T @ O.quat2rmat(g.QQ[k]) @ O.quat2rmat(g.Q0[i])
        ang = np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1))
        ax = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (2 * np.sin(ang))
        assert np.allclose(w[k, :3], ax * ang, atol=1e-10)


def test_make_A_rule_and_dropped_edges():
    I = np.array([[0, 1], [1, 2], [2, 0], [2, 3], [3, 1], [0, 0]], dtype=np.int32)
    A = O.make_A(4, 2, I).toarray()
    # f=2: columns are nodes 2,3.  (0,1): j fixed -> empty; (1,2): +1 at 2 only; (2,0): j fixed ->
    # dropped although i is free; (2,3): -1 at 2, +1 at 3; (3,1): j fixed -> dropped; (0,0): empty
    assert np.array_equal(A, np.array([[0, 0], [1, 0], [0, 0], [-1, 1], [0, 0], [0, 0]], dtype=float))


@pytest.mark.parametrize("cost", [O.L2, O.L1, O.GEMAN_MCCLURE, O.HUBER, O.CAUCHY, O.WELSCH])
def test_known_answer_noise_free(cost):
    g = G.small_graph()  # noise-free: ground truth is the unique zero-residual solution
    r = O.irls(g.QQ, g.I, None, cost, SIGMA, g.Q0, g.f, 20, 1e-12, solver="direct")
    assert O.geodesic_rms(r.Q, g.Qgt, g.f) < 1e-14
    assert r.iters <= 8


@pytest.mark.parametrize("cost", range(14))
def test_three_formulations_agree(cost):
    g = G.small_graph(n=50, extra=250, sigma_n=0.03, outlier_frac=0.1, sigma_init=0.3, seed=5, f=2,
                      fixed_anywhere=True)
    sg = sigma_for(cost)
    a = O.irls(g.QQ, g.I, None, cost, sg, g.Q0, g.f, 5, -1.0, solver="lstsq")
    b = O.irls(g.QQ, g.I, None, cost, sg, g.Q0, g.f, 5, -1.0, solver="direct")
    c = O.irls(g.QQ, g.I, None, cost, sg, g.Q0, g.f, 5, -1.0, solver="pcg")
    assert a.iters == b.iters == c.iters == 5
    assert O.geodesic_rms(a.Q, b.Q, g.f) < 1e-12
    assert O.geodesic_rms(a.Q, c.Q, g.f) < 1e-11
    assert np.allclose(a.weights, b.weights, rtol=1e-8, atol=1e-10)


def test_stop_rule_and_iteration_count():
    g = G.small_graph(sigma_n=0.01, seed=2)
    r = O.irls(g.QQ, g.I, None, O.L2, SIGMA, g.Q0, g.f, 50, 1e-3)
    assert 1 <= r.iters < 50 and r.scores[-1] <= 1e-3 and all(s > 1e-3 for s in r.scores[:-1])
    r0 = O.irls(g.QQ, g.I, None, O.L2, SIGMA, g.Q0, g.f, 0, 1e-3)
    assert r0.iters == 0 and np.array_equal(r0.Q, g.Q0)
    r3 = O.irls(g.QQ, g.I, None, O.L2, SIGMA, g.Q0, g.f, 3, -1.0)
    assert r3.iters == 3


def test_huber_is_sticky_and_l2_keeps_ones():
    g = G.small_graph(n=50, extra=200, sigma_n=0.05, outlier_frac=0.2, seed=9)
    r = O.irls(g.QQ, g.I, None, O.L2, SIGMA, g.Q0, g.f, 3, -1.0)
    assert np.all(r.weights == 1.0)
    E = np.zeros((3, 3))
    E[0, 0] = 10 * SIGMA
    w = O.update_weights(O.HUBER, SIGMA, E, np.array([0.5, 0.25, 0.125]))
    assert w[0] == pytest.approx(np.sqrt(1.345 / 10)) and w[1] == 0.25 and w[2] == 0.125


def test_init_mst_reproduces_noise_free_ground_truth():
    g = G.small_graph(n=80, extra=100)
    Q = np.zeros_like(g.Qgt)
    Q[0] = g.Qgt[0]
    Qm = O.init_mst(Q, g.QQ, g.I, 1)
    assert O.geodesic_rms(Qm, g.Qgt, 1) < 1e-13


def test_rmat2quat_roundtrip_all_branches():
    rng = np.random.default_rng(4)
    qs = np.concatenate([G._rand_quat(rng, 50),
                         [[1, 0, 0, 1e-9], [0, 1, 0, 1e-9], [0, 0, 1, 1e-9], [0.6, 0.8, 0, 0]]])
    for q in qs:
        q = q / np.linalg.norm(q)
        q2 = O.rmat2quat(O.quat2rmat(q))
        assert min(np.linalg.norm(q2 - q), np.linalg.norm(q2 + q)) < 1e-7


def test_golden_bundled_graph():
    z = np.load(os.path.join(GOLD, "bundled_graph.npz"))
    I, QQ, Qm, f = z["I"], z["QQ"], z["Q_mst"], int(z["f"])
    assert I.shape == (3655, 2) and Qm.shape == (1832, 4) and f == 1
    assert np.allclose(O.init_mst(z["Q_file"], QQ, I, max(int(z["n_given"]), f)), Qm, atol=1e-15)
    for cost in (O.L2, O.L1, O.GEMAN_MCCLURE, O.HUBER):
        r = O.irls(QQ, I, None, cost, SIGMA, Qm, f, 50, 1e-3, solver="direct")
        assert r.iters == int(z[f"c{cost}_iters"]) == 2
        assert np.allclose(r.scores, z[f"c{cost}_scores"], rtol=1e-9)
        assert O.geodesic_rms(r.Q, z[f"c{cost}_Q"], f) < 1e-12
        assert np.allclose(r.weights, z[f"c{cost}_weights"], rtol=1e-7, atol=1e-9)
    # the PCG formulation reproduces the direct-solve golden on this cond ~ 2.6e5 chain graph
    r = O.irls(QQ, I, None, O.L1, SIGMA, Qm, f, 10, -1.0, solver="pcg")
    assert O.geodesic_rms(r.Q, z["l1x10_Q"], f) < 1e-9
    assert np.allclose(r.scores, z["l1x10_scores"], rtol=1e-6)


def test_golden_small_costs():
    z = np.load(os.path.join(GOLD, "small_costs.npz"))
    for cost in range(14):
        r = O.irls(z["QQ"], z["I"], None, cost, sigma_for(cost), z["Q0"], int(z["f"]), 6, -1.0, solver="direct")
        assert O.geodesic_rms(r.Q, z[f"c{cost}_Q"], int(z["f"])) < 1e-11, O.COST_NAMES[cost]
        assert np.allclose(r.weights, z[f"c{cost}_weights"], rtol=1e-7, atol=1e-9), O.COST_NAMES[cost]


def test_text_format_roundtrip(tmp_path):
    g = G.small_graph(n=30, extra=40, sigma_n=0.01)
    p = tmp_path / "g.txt"
    G.write_ral_text(str(p), g.I + 1, g.QQ, g.Q0[:1], 1)      # ids start at 1 like the bundled file
    I, QQ, Q, f, ng = G.read_ral_text(str(p))
    assert f == 1 and ng == 1 and np.array_equal(I, g.I) and np.allclose(QQ, g.QQ) and np.allclose(Q[0], g.Q0[0])
