"""Parity of the sm_100a CUDA path (through the C ABI) against the CPU oracle.  -m gpu only.

Tolerances (FP64 everywhere; north_star: final rotations within 1e-6 rad RMS of the reference):
  * per-edge residuals              |dw| <= 1e-12 rad        (same arithmetic, different FMA contraction)
  * Laplacian apply                 relative 1e-12
  * irls() final rotations          geodesic RMS <= 1e-8 rad (asserted 100x tighter than the 1e-6 bar;
                                    the gap is the PCG tolerance cg_rtol = 1e-10 vs an exact solve)
  * per-iteration scores            relative 1e-6
"""
import os

import numpy as np
import pytest

from oracle import graphs as G
from oracle import irls_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
SIGMA = 5 * np.pi / 180.0
RMS_TOL = 1e-8


def sigma_for(cost):
    return 4 * SIGMA if cost == O.TALWAR else SIGMA


# ---------------------------------------------------------------------------- kernels
@pytest.mark.parametrize("maker", ["small", "quirk", "large_angle", "kitti"])
def test_residual_kernel_per_edge(solver, maker):
    if maker == "small":
        g = G.small_graph(n=300, extra=2000, sigma_n=0.1, sigma_init=0.3, seed=1)
    elif maker == "quirk":
        g = G.small_graph(n=100, extra=700, sigma_n=0.1, sigma_init=0.3, seed=2, f=5, fixed_anywhere=True)
    elif maker == "large_angle":
        g = G.small_graph(n=257, extra=1000, sigma_n=0.1, seed=3)
        g.Q0[g.f:] = np.array([0, 0, 0, 1.0])            # every new view starts at identity (App. A.6.9)
        g.QQ[::5] *= -1.0                                 # and q / -q sign ambiguity on the measurements
    else:
        g = G.kitti_like_graph()
    solver.upload(g.QQ, g.I, g.Q0, g.f)
    w = solver.probe_residual()
    ref = O.log_map(O.delta_rel(g.I, g.QQ, g.Q0))
    assert np.abs(w[:, :3] - ref[:, :3]).max() <= 1e-12
    assert np.abs(w[:, 3] - ref[:, 3]).max() <= 1e-12
    assert np.all(w[:, 3] >= -np.pi) and np.all(w[:, 3] < np.pi)


def test_residual_kernel_degenerate_rows(solver):
    """s < EPS rows give exactly 0 (ral/l1_irls.cpp:527-531), for p_w = +1 and p_w = -1."""
    n = 6
    Q0 = np.tile(np.array([0, 0, 0, 1.0]), (n, 1))
    I = np.array([[0, 1], [1, 2], [2, 3], [3, 4], [4, 5]], dtype=np.int32)
    QQ = np.tile(np.array([0, 0, 0, 1.0]), (5, 1))
    QQ[1] = [0, 0, 0, -1.0]
    QQ[2] = [1e-17, 0, 0, 1.0]
    solver.upload(QQ, I, Q0, 1)
    w = solver.probe_residual()
    ref = O.log_map(O.delta_rel(I, QQ, Q0))
    assert np.array_equal(w[:, :3], np.zeros((5, 3))) and np.array_equal(ref[:, :3], np.zeros((5, 3)))
    assert np.allclose(w[:, 3], ref[:, 3], atol=1e-15)


@pytest.mark.parametrize("f,flip", [(1, False), (4, True)])
def test_laplacian_apply(solver, f, flip):
    g = G.small_graph(n=500, extra=4000, sigma_n=0.1, seed=4, f=f, fixed_anywhere=flip)
    rng = np.random.default_rng(0)
    wts = rng.uniform(0.01, 30.0, g.m)
    X = rng.standard_normal((g.n - f, 3))
    solver.upload(g.QQ, g.I, g.Q0, g.f)
    Y = solver.probe_laplacian_apply(wts, X)
    A = O.make_A(g.n, f, g.I)
    ref = A.T @ ((wts * wts)[:, None] * (A @ X))
    assert np.abs(Y - ref).max() <= 1e-12 * np.abs(ref).max()


# ---------------------------------------------------------------------------- whole loop
@pytest.mark.parametrize("cost", range(14))
def test_irls_all_costs_vs_oracle(solver, cost):
    g = G.small_graph(n=300, extra=2500, sigma_n=0.03, outlier_frac=0.1, sigma_init=0.3, seed=21, f=3,
                      fixed_anywhere=True)
    sg = sigma_for(cost)
    ref = O.irls(g.QQ, g.I, None, cost, sg, g.Q0, g.f, 8, -1.0, solver="direct")
    Q, w, info = solver.irls(g.QQ, g.I, None, cost, sg, g.Q0, g.f, 8, -1.0)
    assert info.iters == ref.iters == 8
    assert info.cg_hit_max == 0
    assert np.allclose(info.scores, ref.scores, rtol=1e-6, atol=1e-12)
    assert O.geodesic_rms(Q, ref.Q, g.f) <= RMS_TOL
    assert np.array_equal(Q[:g.f], g.Q0[:g.f])                      # fixed rows untouched
    if cost == O.TALWAR:                                            # threshold cost: allow flips at the edge
        assert (w != ref.weights).mean() < 0.01
    else:
        assert np.allclose(w, ref.weights, rtol=1e-5, atol=1e-8)


def test_golden_small_costs(solver):
    z = np.load(os.path.join(GOLD, "small_costs.npz"))
    f = int(z["f"])
    for cost in range(14):
        Q, w, info = solver.irls(z["QQ"], z["I"], None, cost, sigma_for(cost), z["Q0"], f, 6, -1.0)
        assert O.geodesic_rms(Q, z[f"c{cost}_Q"], f) <= RMS_TOL, O.COST_NAMES[cost]
        assert np.allclose(info.scores, z[f"c{cost}_scores"], rtol=1e-6, atol=1e-12)


def test_golden_bundled_graph(solver):
    """Config 1: the reference's only fixture, from its init_mst start, CLI defaults."""
    z = np.load(os.path.join(GOLD, "bundled_graph.npz"))
    I, QQ, Qm, f = z["I"], z["QQ"], z["Q_mst"], int(z["f"])
    for cost in (O.L2, O.L1, O.GEMAN_MCCLURE, O.HUBER):
        Q, w, info = solver.irls(QQ, I, None, cost, SIGMA, Qm, f, 50, 1e-3)
        assert info.iters == int(z[f"c{cost}_iters"])
        assert np.allclose(info.scores, z[f"c{cost}_scores"], rtol=1e-6)
        assert O.geodesic_rms(Q, z[f"c{cost}_Q"], f) <= RMS_TOL
        assert np.allclose(w, z[f"c{cost}_weights"], rtol=1e-5, atol=1e-8)
    Q, w, info = solver.irls(QQ, I, None, O.L1, SIGMA, Qm, f, 10, -1.0)
    assert O.geodesic_rms(Q, z["l1x10_Q"], f) <= RMS_TOL
    assert np.allclose(info.scores, z["l1x10_scores"], rtol=1e-6)


@pytest.mark.parametrize("cost", [O.L1, O.GEMAN_MCCLURE, O.HUBER])
def test_known_answer_noise_free(solver, cost):
    g = G.small_graph(n=2000, extra=20000, seed=5)
    Q, w, info = solver.irls(g.QQ, g.I, None, cost, SIGMA, g.Q0, g.f, 20, 1e-13)
    assert O.geodesic_rms(Q, g.Qgt, g.f) <= 1e-9
    assert info.iters < 20


def test_config2_kitti_like(solver):
    """Config 2 (n=4541, m=50000), 6 iterations, oracle with PCG at rtol 1e-13 as the checker."""
    g = G.kitti_like_graph()
    for cost in (O.L1, O.GEMAN_MCCLURE):
        ref = O.irls(g.QQ, g.I, None, cost, SIGMA, g.Q0, g.f, 6, -1.0, solver="pcg", pcg_rtol=1e-13)
        Q, w, info = solver.irls(g.QQ, g.I, None, cost, SIGMA, g.Q0, g.f, 6, -1.0)
        assert info.cg_hit_max == 0
        assert np.allclose(info.scores, ref.scores, rtol=1e-6)
        assert O.geodesic_rms(Q, ref.Q, g.f) <= RMS_TOL
        assert np.allclose(w, ref.weights, rtol=1e-5, atol=1e-8)


def test_config3_full_size(solver):
    """Config 3 (n=100k, m=1M): 3 L1 iterations against the oracle (PCG, rtol 1e-13), then
    size-independent properties of the full 30-iteration run."""
    g = G.random_graph()
    ref = O.irls(g.QQ, g.I, None, O.L1, SIGMA, g.Q0, g.f, 3, -1.0, solver="pcg", pcg_rtol=1e-13)
    Q, w, info = solver.irls(g.QQ, g.I, None, O.L1, SIGMA, g.Q0, g.f, 3, -1.0)
    assert np.allclose(info.scores, ref.scores, rtol=1e-6)
    assert O.geodesic_rms(Q, ref.Q, g.f) <= RMS_TOL
    assert np.allclose(w, ref.weights, rtol=1e-5, atol=1e-8)
    # Geman-McClure, 30 iterations: monotone score decay, unit-norm drift stays O(eps * iters),
    # fixed node untouched, estimate far closer to ground truth than the start
    Q, w, info = solver.irls(g.QQ, g.I, None, O.GEMAN_MCCLURE, SIGMA, g.Q0, g.f, 30, -1.0)
    assert info.iters == 30 and info.cg_hit_max == 0
    s = np.array(info.scores)
    assert np.all(np.diff(s[2:]) < 0)
    assert np.abs(np.linalg.norm(Q, axis=1) - 1).max() < 1e-12
    assert np.array_equal(Q[0], g.Q0[0])
    assert O.geodesic_rms(Q, g.Qgt, g.f) < 0.5 * O.geodesic_rms(g.Q0, g.Qgt, g.f)
    # determinism: the same call twice is bitwise identical (atomic-free reductions)
    Q2, w2, _ = solver.irls(g.QQ, g.I, None, O.GEMAN_MCCLURE, SIGMA, g.Q0, g.f, 30, -1.0)
    assert np.array_equal(Q, Q2) and np.array_equal(w, w2)


def test_config3_30_iterations_vs_golden(solver):
    """The configuration BASELINE.json's metric is quoted on, for the FULL 30 L1 iterations the benchmark runs
    (iterations 10-30 are where the weights hit the 1e4 clamp and the system's condition number explodes):
    final rotations against the C restatement's run with plain Jacobi-PCG at rtol 1e-12
    (tests/golden/make_golden_cfg3.py; 291 568 CG iterations, a different preconditioner, tolerance and
    summation order).  north_star's bar is 1e-6 rad RMS; asserted 100x tighter."""
    gold = np.load(os.path.join(GOLD, "cfg3_l1_30iters.npz"))
    g = G.random_graph()
    assert (g.n, g.m, g.f) == (int(gold["n"]), int(gold["m"]), int(gold["f"]))
    Q, w, info = solver.irls(g.QQ, g.I, None, O.L1, SIGMA, g.Q0, g.f, 30, -1.0)
    assert info.iters == 30 and info.cg_hit_max == 0
    rms = O.geodesic_rms(Q, gold["Q"], g.f)
    dev = np.abs(np.array(info.scores) / gold["scores"] - 1).max()
    print(f"config 3, 30 L1 iterations: geodesic RMS vs golden {rms:.3e} rad, max relative score deviation {dev:.3e}")
    assert rms <= 1e-8
    assert dev <= 1e-6
    assert np.allclose(w[::97], gold["weights_every97"], rtol=1e-4, atol=1e-8)
    assert abs(w.sum() / float(gold["weights_sum"]) - 1) <= 1e-6


def test_known_answer_full_size(solver):
    """1M-edge noise-free graph: exact ground truth is recovered (implementation-independent)."""
    g = G.random_graph(sigma_n=0.0, outlier_frac=0.0, seed=77)
    Q, w, info = solver.irls(g.QQ, g.I, None, O.HUBER, SIGMA, g.Q0, g.f, 20, 1e-13)
    assert O.geodesic_rms(Q, g.Qgt, g.f) <= 1e-9


def test_resident_matches_host_call(solver):
    g = G.small_graph(n=400, extra=3000, sigma_n=0.02, outlier_frac=0.05, seed=8)
    Q, w, info = solver.irls(g.QQ, g.I, None, O.L1, SIGMA, g.Q0, g.f, 5, -1.0)
    solver.upload(g.QQ, g.I, g.Q0, g.f)
    for _ in range(2):                                  # restarts from the uploaded Q0 every call
        info2 = solver.irls_resident(O.L1, SIGMA, 5, -1.0)
        Q2, w2 = solver.download()
        assert np.array_equal(Q, Q2) and np.array_equal(w, w2) and info2.iters == 5
    assert info2.kernel_launches > 0


# ---------------------------------------------------------------------------- edge cases
def test_stop_rule(solver):
    g = G.small_graph(sigma_n=0.01, seed=2)
    ref = O.irls(g.QQ, g.I, None, O.L2, SIGMA, g.Q0, g.f, 50, 1e-3)
    Q, w, info = solver.irls(g.QQ, g.I, None, O.L2, SIGMA, g.Q0, g.f, 50, 1e-3)
    assert info.iters == ref.iters and np.all(w == 1.0)
    Q, w, info = solver.irls(g.QQ, g.I, None, O.L2, SIGMA, g.Q0, g.f, 0, 1e-3)
    assert info.iters == 0 and np.array_equal(Q, g.Q0)


def test_empty_and_degenerate_inputs(solver):
    import irotavg_b200 as ira
    # no edges: X = 0, one iteration, Q unchanged
    Q0 = G._rand_quat(np.random.default_rng(0), 5)
    Q, w, info = solver.irls(np.zeros((0, 4)), np.zeros((0, 2), dtype=np.int32), None, O.L1, SIGMA, Q0, 1, 10, 1e-3)
    assert info.iters == 1 and np.array_equal(Q, Q0) and w.shape == (0,)
    # every node fixed: nothing to optimise; the reference's mean over zero rows is NaN -> loop ends
    I = np.array([[0, 1], [1, 2]], dtype=np.int32)
    QQ = G._rand_quat(np.random.default_rng(1), 2)
    Q, w, info = solver.irls(QQ, I, None, O.L1, SIGMA, Q0[:3], 3, 10, 1e-3)
    assert info.iters == 1 and np.array_equal(Q, Q0[:3])
    # a free node with no edge at all stays where it is
    g = G.small_graph(n=50, extra=200, sigma_n=0.01, seed=3)
    Q0b = np.vstack([g.Q0, G._rand_quat(np.random.default_rng(2), 1)])
    Q, w, info = solver.irls(g.QQ, g.I, None, O.L2, SIGMA, Q0b, g.f, 3, -1.0)
    assert np.array_equal(Q[-1], Q0b[-1])
    ref = O.irls(g.QQ, g.I, None, O.L2, SIGMA, g.Q0, g.f, 3, -1.0)
    assert O.geodesic_rms(Q[:-1], ref.Q, g.f) <= RMS_TOL
    # errors: unknown cost, f = 0, out-of-range endpoint
    with pytest.raises(ira.IraError) as e:
        solver.irls(g.QQ, g.I, None, 14, SIGMA, g.Q0, g.f, 3, -1.0)
    assert e.value.status == 5
    with pytest.raises(ira.IraError):
        solver.irls(g.QQ, g.I, None, O.L1, SIGMA, g.Q0, 0, 3, -1.0)
    bad = g.I.copy()
    bad[3, 1] = g.n
    with pytest.raises(ira.IraError):
        solver.irls(g.QQ, bad, None, O.L1, SIGMA, g.Q0, g.f, 3, -1.0)
    # a self-loop (i == j) is refused (the reference's make_A would silently turn it into a single -1 entry),
    # by the general pipeline and by the single-block window solver alike
    loop = g.I.copy()
    loop[5] = [7, 7]
    with pytest.raises(ira.IraError) as e:
        solver.irls(g.QQ, loop, None, O.L1, SIGMA, g.Q0, g.f, 3, -1.0)
    assert e.value.status == 1 and "self-loop" in str(e.value)
    gw = G.small_graph(n=15, extra=27, sigma_n=0.005, sigma_init=0.05, seed=1, f=4)
    loopw = gw.I.copy()
    loopw[2] = [9, 9]
    with pytest.raises(ira.IraError) as e:
        solver.l1ra_irls(gw.QQ, loopw, gw.Q0, gw.f, 10, 1e-3, O.GEMAN_MCCLURE, SIGMA, 10, 1e-3)
    assert "self-loop" in str(e.value)
    # the handle is still usable afterwards
    Q, w, info = solver.irls(g.QQ, g.I, None, O.L2, SIGMA, g.Q0, g.f, 3, -1.0)
    assert O.geodesic_rms(Q, ref.Q, g.f) <= RMS_TOL


def test_large_angle_identity_start(solver):
    """Config-5 regime: new views enter at identity, residuals anywhere in [-pi, pi) (App. A.6.9)."""
    g = G.small_graph(n=120, extra=600, sigma_n=0.01, seed=6)
    Q0 = g.Q0.copy()
    Q0[60:] = np.array([0, 0, 0, 1.0])
    ref = O.irls(g.QQ, g.I, None, O.GEMAN_MCCLURE, SIGMA, Q0, g.f, 10, -1.0, solver="direct")
    Q, w, info = solver.irls(g.QQ, g.I, None, O.GEMAN_MCCLURE, SIGMA, Q0, g.f, 10, -1.0)
    assert np.allclose(info.scores, ref.scores, rtol=1e-6, atol=1e-12)
    assert O.geodesic_rms(Q, ref.Q, g.f) <= RMS_TOL


@pytest.mark.parametrize("solver_kind", [1, 2, 6, 8, 12, 32])
@pytest.mark.parametrize("maker", ["random", "kitti", "tiny"])
def test_solver_variants(built_lib, solver_kind, maker):
    """solver 1 = one kernel per CG step (SELL SpMV, host-polled), 2 = persistent cooperative PCG
    (register-resident state when one row per lane fits), 32 = the same with the matrix in shared memory
    (ira_pcg2.cuh), 6 = persistent with HBM-resident vectors, 8 / 12 = the
    barrier-free kernels of the multi-GPU path (self-validating data, ira_peer.cuh) on ONE GPU, register- /
    HBM-resident."""
    import irotavg_b200 as ira
    if maker == "random":
        g = G.small_graph(n=3000, extra=30000, sigma_n=0.03, outlier_frac=0.1, seed=31, f=2, fixed_anywhere=True)
    elif maker == "kitti":
        g = G.kitti_like_graph(n=1500, m=16000)
    else:
        g = G.small_graph(n=14, extra=26, sigma_n=0.01, seed=4, f=4)
    ref = O.irls(g.QQ, g.I, None, O.GEMAN_MCCLURE, SIGMA, g.Q0, g.f, 6, -1.0, solver="direct")
    with ira.Solver(solver=solver_kind) as s:
        Q, w, info = s.irls(g.QQ, g.I, None, O.GEMAN_MCCLURE, SIGMA, g.Q0, g.f, 6, -1.0)
        Q2, w2, _ = s.irls(g.QQ, g.I, None, O.GEMAN_MCCLURE, SIGMA, g.Q0, g.f, 6, -1.0)
    assert info.cg_hit_max == 0 and all(k > 0 for k in info.cg_iters)
    assert np.allclose(info.scores, ref.scores, rtol=1e-6, atol=1e-12)
    assert O.geodesic_rms(Q, ref.Q, g.f) <= RMS_TOL
    assert np.allclose(w, ref.weights, rtol=1e-5, atol=1e-8)
    assert np.array_equal(Q, Q2) and np.array_equal(w, w2)          # bitwise repeatable


def test_pair_preconditioner_l1_late_iterations(built_lib):
    """L1 weights reach the 1e4 clamp: the pairwise block-Jacobi correction must cut PCG iterations by
    a large factor and leave the result unchanged (same fixed point, tolerance of the two solves)."""
    import irotavg_b200 as ira
    g = G.random_graph(n=10000, m=100000)
    with ira.Solver(pair_theta=0.0) as s0:
        Q0, w0, i0 = s0.irls(g.QQ, g.I, None, O.L1, SIGMA, g.Q0, g.f, 16, -1.0)
    for kind in (0, 6, 1, 8, 12):             # persistent (registers / HBM state), one-kernel-per-step, barrier-free
        with ira.Solver(solver=kind) as s1:
            Q1, w1, i1 = s1.irls(g.QQ, g.I, None, O.L1, SIGMA, g.Q0, g.f, 16, -1.0)
        assert i0.cg_hit_max == 0 and i1.cg_hit_max == 0
        assert sum(i1.cg_iters) * 5 < sum(i0.cg_iters), (kind, i0.cg_iters, i1.cg_iters)
        assert np.allclose(i1.scores, i0.scores, rtol=1e-6)
        assert O.geodesic_rms(Q1, Q0, g.f) <= RMS_TOL


def test_three_by_three_blocks_vs_pairs_and_oracle(built_lib):
    """Late L1 iterations: letting single nodes join a stiff pair (3x3 blocks, pair_theta3) must cut the PCG
    iterations well below the pairs-only count, in every PCG variant, and the result must still be the exact
    solve's (direct oracle, 20 iterations, the regime where the blocks matter)."""
    import irotavg_b200 as ira
    g = G.small_graph(n=4000, extra=40000, sigma_n=0.05, outlier_frac=0.1, sigma_init=0.1, seed=77)
    ref = O.irls(g.QQ, g.I, None, O.L1, SIGMA, g.Q0, g.f, 20, -1.0, solver="direct")
    with ira.Solver(pair_theta3=0.0) as s2:
        Q2, w2, i2 = s2.irls(g.QQ, g.I, None, O.L1, SIGMA, g.Q0, g.f, 20, -1.0)
    assert O.geodesic_rms(Q2, ref.Q, g.f) <= RMS_TOL
    for kind in (0, 6, 1, 8, 12):
        with ira.Solver(solver=kind) as s3:
            Q3, w3, i3 = s3.irls(g.QQ, g.I, None, O.L1, SIGMA, g.Q0, g.f, 20, -1.0)
        assert i3.cg_hit_max == 0
        assert sum(i3.cg_iters[10:]) * 1.2 < sum(i2.cg_iters[10:]), (kind, i2.cg_iters, i3.cg_iters)
        assert np.allclose(i3.scores, ref.scores, rtol=1e-6, atol=1e-12)
        assert O.geodesic_rms(Q3, ref.Q, g.f) <= RMS_TOL
        assert np.allclose(w3, ref.weights, rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("lpr", [2, 4, 8, 16, 32])
def test_lanes_per_row_variants(built_lib, lpr):
    import irotavg_b200 as ira
    g = G.small_graph(n=300, extra=2500, sigma_n=0.03, outlier_frac=0.1, seed=21)
    ref = O.irls(g.QQ, g.I, None, O.L1, SIGMA, g.Q0, g.f, 4, -1.0, solver="direct")
    with ira.Solver(lanes_per_row=lpr) as s:
        Q, w, info = s.irls(g.QQ, g.I, None, O.L1, SIGMA, g.Q0, g.f, 4, -1.0)
    assert O.geodesic_rms(Q, ref.Q, g.f) <= RMS_TOL


def test_two_level_pcg_on_chain_graphs(built_lib):
    """View graphs are chains (cond ~ n^2): the two-level kernel (Jacobi + piecewise-constant coarse space over index
    blocks, ira_coarse.cuh; tridiagonal coarse operator of up to 1 024 blocks by default, dense 64-block one with
    solver +256) must give the same answer as the one-level kernels (solver +128) and the oracle, in several times
    fewer PCG iterations; L1 keeps the block-Jacobi kernels."""
    import irotavg_b200 as ira
    from oracle import rotavg_stream as RS
    ops, Qgt = RS.make_stream(n_frames=2500, loop_every=900, min_loop_gap=600, sigma_n=0.01)
    I = np.array([(op[1], op[2]) for op in ops if op[0] == "E"], dtype=np.int32)
    QQ = np.array([O.rmat2quat(op[3]) for op in ops if op[0] == "E"])
    rng = np.random.default_rng(5)
    Q0 = O.quat_mult(Qgt, G._exp_quat(rng.normal(0, 0.05, (Qgt.shape[0], 3))))
    Q0[0] = Qgt[0]
    res = {}
    for sv in (0, 256, 128):
        with ira.Solver(solver=sv) as s:
            Q, w, info = s.irls(QQ, I, None, O.GEMAN_MCCLURE, SIGMA, Q0, 1, 6, -1.0)
            Ql, il = s.l1ra(QQ, I, None, Q0, 1, 2, 1e-9)
            Q1, w1, i1 = s.irls(QQ, I, None, O.L1, SIGMA, Q0, 1, 3, -1.0)
        res[sv] = (Q, w, info, Ql, il, Q1, i1)
    ref = O.irls(QQ, I, None, O.GEMAN_MCCLURE, SIGMA, Q0, 1, 6, -1.0, solver="direct")
    refl = O.l1ra(QQ, I, None, Q0, 1, 2, 1e-9)
    for sv in (0, 256, 128):
        Q, w, info, Ql, il, Q1, i1 = res[sv]
        assert info.cg_hit_max == 0 and il.cg_hit_max == 0
        assert O.geodesic_rms(Q, ref.Q, 1) <= RMS_TOL
        assert np.allclose(w, ref.weights, rtol=1e-5, atol=1e-8)
        assert il.iters == refl.iters and O.geodesic_rms(Ql, refl.Q, 1) <= RMS_TOL
    assert res[0][2].pcg_kernel == 8 and res[256][2].pcg_kernel == 7 and res[128][2].pcg_kernel in (2, 3)
    assert res[0][6].pcg_kernel in (2, 3)                       # L1: stiff pairs need the exact blocks
    print("PCG iterations, tridiagonal / dense two-level vs one-level: irls", sum(res[0][2].cg_iters), sum(res[256][2].cg_iters),
          sum(res[128][2].cg_iters), "l1ra Newton", sum(res[0][4].cg_iters), sum(res[256][4].cg_iters), sum(res[128][4].cg_iters))
    assert sum(res[256][2].cg_iters) * 3 <= sum(res[128][2].cg_iters)
    assert sum(res[256][4].cg_iters) * 2 <= sum(res[128][4].cg_iters)
    assert sum(res[0][2].cg_iters) <= sum(res[256][2].cg_iters)          # 2 500 views: 313 blocks against 64
    assert sum(res[0][4].cg_iters) <= sum(res[256][4].cg_iters)


def test_two_level_tridiagonal_dead_pivots(built_lib):
    """Tridiagonal coarse operator on a graph whose partition has dead unknowns - a long all-fixed prefix leaves the
    first five blocks of the partition empty (zero pivots, switched off): same rotations as the one-level kernels and
    the oracle."""
    import irotavg_b200 as ira
    from oracle import rotavg_stream as RS
    ops, Qgt = RS.make_stream(n_frames=1500, loop_every=400, min_loop_gap=300, sigma_n=0.01)
    I = np.array([(op[1], op[2]) for op in ops if op[0] == "E"], dtype=np.int32)
    QQ = np.array([O.rmat2quat(op[3]) for op in ops if op[0] == "E"])
    rng = np.random.default_rng(9)
    Q0 = O.quat_mult(Qgt, G._exp_quat(rng.normal(0, 0.05, (Qgt.shape[0], 3))))
    f = 40                                                       # five whole blocks of 8 rows are fixed
    Q0[:f] = Qgt[:f]
    ref = O.irls(QQ, I, None, O.GEMAN_MCCLURE, SIGMA, Q0, f, 4, -1.0, solver="direct")
    out = {}
    for sv in (0, 128):
        with ira.Solver(solver=sv) as s:
            Q, w, info = s.irls(QQ, I, None, O.GEMAN_MCCLURE, SIGMA, Q0, f, 4, -1.0)
        assert info.cg_hit_max == 0
        assert O.geodesic_rms(Q, ref.Q, f) <= RMS_TOL
        out[sv] = info
    assert out[0].pcg_kernel == 8
    assert sum(out[0].cg_iters) * 2 <= sum(out[128].cg_iters)
