"""Parity pinned to the REFERENCE'S OWN CODE.

tests/golden/ref_*.npz were produced by oracle/_ref: ral/l1_irls.cpp and ral/test.cpp compiled UNMODIFIED from the
reference tree (oracle/build_ref.py; stand-in headers for the absent Eigen / SuiteSparse in oracle/ref_shim/) and run
on the reference's bundled fixture and on seeded synthetic graphs (tests/golden/make_golden_ref.py).

CPU (-m "not gpu"): the oracle restatement reproduces every golden; when oracle/_ref is built (this container; it
travels to the GPU box as a prebuilt file) the oracle is also compared with the live reference on fresh random graphs.
GPU (-m gpu): the CUDA path, through the C ABI / the CLI binary, reproduces the same goldens.
Tolerances: oracle vs reference 1e-12 rad RMS (exact solves on both sides, different summation order); CUDA vs
reference 1e-8 rad RMS (PCG at cg_rtol 1e-10 against the reference's QR / LU), weights 1e-5 relative.
"""
import os
import subprocess

import numpy as np
import pytest

from oracle import graphs as G
from oracle import irls_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SIGMA = 5 * np.pi / 180.0


def sigma_for(cost):
    return 4 * SIGMA if cost == O.TALWAR else SIGMA


def graphs():
    """Must match tests/golden/make_golden_ref.py::graphs()."""
    gq = G.small_graph(n=60, extra=300, sigma_n=0.03, outlier_frac=0.1, sigma_init=0.3, seed=11, f=3, fixed_anywhere=True)
    gk = G.banded_graph()
    gw = G.small_graph(n=15, extra=27, sigma_n=0.005, sigma_init=0.05, seed=1, f=4)
    gi = G.small_graph(n=120, extra=500, sigma_n=0.02, seed=3)
    gi.Q0[gi.f:] = np.array([0, 0, 0, 1.0])
    gi.QQ[::5] *= -1.0
    return gq, gk, gw, gi


@pytest.fixture(scope="module")
def ref_small():
    return np.load(os.path.join(GOLD, "ref_small.npz"))


@pytest.fixture(scope="module")
def ref_cli():
    return np.load(os.path.join(GOLD, "ref_bundled_cli.npz"))


@pytest.fixture(scope="module")
def bundled():
    return np.load(os.path.join(GOLD, "bundled_graph.npz"))


def wclose(a, b, rtol):
    return np.allclose(a, b, rtol=rtol, atol=rtol * 1e-3)


# ============================================================================== CPU: oracle == reference
@pytest.mark.parametrize("cost", range(14))
def test_oracle_irls_all_costs_vs_reference(ref_small, cost):
    gq = graphs()[0]
    r = O.irls(gq.QQ, gq.I, None, cost, sigma_for(cost), gq.Q0, gq.f, 6, -1.0, solver="lstsq")
    assert r.iters == int(ref_small[f"quirk_c{cost}_iters"])
    assert O.geodesic_rms(r.Q, ref_small[f"quirk_c{cost}_Q"], gq.f) <= 1e-12
    assert wclose(r.weights, ref_small[f"quirk_c{cost}_weights"], 1e-7)


def test_oracle_rank_deficient_talwar_vs_reference(ref_small):
    """Talwar at 5 deg gives every edge of two free nodes weight 0: the reference's SPQR call returns its basic
    solution (dead columns -> x = 0); the oracle's minimum-norm least squares agrees because the dead columns are
    exactly zero."""
    gq = graphs()[0]
    r = O.irls(gq.QQ, gq.I, None, O.TALWAR, SIGMA, gq.Q0, gq.f, 6, -1.0, solver="lstsq")
    w = ref_small["quirk_talwar5_weights"]
    live = np.zeros(gq.n)
    np.add.at(live, gq.I[:, 0], w > 0); np.add.at(live, gq.I[:, 1], w > 0)
    assert (live[gq.f:] == 0).sum() >= 1                                         # the case is really rank-deficient
    assert O.geodesic_rms(r.Q, ref_small["quirk_talwar5_Q"], gq.f) <= 1e-12
    assert np.array_equal(r.weights, w)


def test_oracle_kernels_vs_reference(ref_small):
    gq, gk, gw, gi = graphs()
    A = O.make_A(gq.n, gq.f, gq.I)
    assert np.array_equal(np.asarray(A.todense()), ref_small["quirk_A"])           # incl. the dropped-edge rule
    for g, key in ((gq, "quirk_residual"), (gi, "ident_residual")):
        w = O.log_map(O.delta_rel(g.I, g.QQ, g.Q0))
        assert np.abs(w - ref_small[key]).max() <= 1e-13
    th = ref_small["ident_residual"][:, 3]
    assert th.min() < -1.0 and th.max() > 1.0                                       # the wrap is exercised
    assert np.abs(O.exp_map(ref_small["expmap_in"].copy()) - ref_small["expmap_out"]).max() <= 1e-15
    assert np.array_equal(ref_small["expmap_out"][0], [0, 0, 0, 1.0])               # NaN -> 0 keeps cos(0) = 1


def test_oracle_init_mst_vs_reference(ref_small):
    gq, gk, _, _ = graphs()
    assert np.abs(O.init_mst(gq.Q0, gq.QQ, gq.I, gq.f) - ref_small["quirk_mst"]).max() <= 1e-13
    perm = ref_small["kitti_perm"]
    assert np.abs(O.init_mst(gk.Q0, gk.QQ[perm], gk.I[perm], gk.f) - ref_small["kitti_mst_shuffled"]).max() <= 1e-12


def test_oracle_l1ra_and_flow_vs_reference(ref_small):
    gq, gk, gw, gi = graphs()
    la = O.l1ra(gq.QQ, gq.I, None, gq.Q0, gq.f, 5, 1e-3)
    assert la.iters == int(ref_small["quirk_l1ra_iters"])
    assert O.geodesic_rms(la.Q, ref_small["quirk_l1ra_Q"], gq.f) <= 1e-11
    la = O.l1ra(gk.QQ, gk.I, None, gk.Q0, gk.f, 5, 1e-3)
    assert la.iters == int(ref_small["kitti_l1ra_iters"])
    assert O.geodesic_rms(la.Q, ref_small["kitti_l1ra_Q"], gk.f) <= 1e-11
    r = O.irls(gk.QQ, gk.I, None, O.GEMAN_MCCLURE, SIGMA, la.Q, gk.f, 50, 1e-3, solver="direct")
    assert r.iters == int(ref_small["kitti_flow_iters"])
    assert O.geodesic_rms(O.quat_normalised(r.Q.copy(), gk.f), ref_small["kitti_flow_Q"], gk.f) <= 1e-11
    assert wclose(r.weights, ref_small["kitti_flow_weights"], 1e-8)
    r = O.irls(gk.QQ, gk.I, None, O.L1, SIGMA, gk.Q0, gk.f, 30, -1.0, solver="direct")     # 30 L1 iterations
    assert O.geodesic_rms(r.Q, ref_small["kitti_l1x30_Q"], gk.f) <= 1e-10
    la = O.l1ra(gw.QQ, gw.I, None, gw.Q0, gw.f, 100, 1e-3)
    r = O.irls(gw.QQ, gw.I, None, O.GEMAN_MCCLURE, SIGMA, la.Q, gw.f, 100, 1e-3, solver="direct")
    assert [la.iters, r.iters] == list(ref_small["window_iters"])
    assert O.geodesic_rms(r.Q, ref_small["window_Q"], gw.f) <= 1e-11
    r = O.irls(gi.QQ, gi.I, None, O.GEMAN_MCCLURE, SIGMA, gi.Q0, gi.f, 20, 1e-3, solver="direct")   # identity start
    assert r.iters == int(ref_small["ident_gm_iters"])
    assert O.geodesic_rms(r.Q, ref_small["ident_gm_Q"], gi.f) <= 1e-11


def test_oracle_bundled_cli_flow_vs_reference(ref_cli, bundled):
    """The reference CLI on its own fixture (ral/test.cpp default flow and three other costs) == the oracle goldens
    that the GPU tests have been held to since round 1 (tests/golden/bundled_graph.npz)."""
    assert O.geodesic_rms(bundled["cli_Q"], ref_cli["default_Q"], 1) <= 1e-12
    assert wclose(bundled["cli_weights"], ref_cli["default_weights"], 1e-9)
    I, QQ, f = bundled["I"], bundled["QQ"], int(bundled["f"])
    la = O.l1ra(QQ, I, None, bundled["Q_mst"], f, 5, 1e-3)
    for tag, cost in (("l1", O.L1), ("huber", O.HUBER), ("l2", O.L2)):
        r = O.irls(QQ, I, None, cost, SIGMA, la.Q, f, 50, 1e-3, solver="direct")
        assert O.geodesic_rms(O.quat_normalised(r.Q.copy(), f), ref_cli[f"{tag}_Q"], f) <= 1e-11, tag
        assert wclose(r.weights, ref_cli[f"{tag}_weights"], 1e-8), tag


@pytest.mark.parametrize("seed", [101, 102, 103])
def test_oracle_vs_live_reference_random(seed):
    """Fresh graphs against the reference library itself (skipped where oracle/_ref cannot be built or found)."""
    from oracle import refbin as R
    if not R.available():
        pytest.skip("oracle/_ref not built (no /root/reference here and no prebuilt copy)")
    rng = np.random.default_rng(seed)
    f = int(rng.integers(1, 4))
    g = G.small_graph(n=int(rng.integers(20, 80)), extra=int(rng.integers(60, 300)), sigma_n=0.05, outlier_frac=0.15,
                      sigma_init=0.4, seed=seed, f=f, fixed_anywhere=bool(seed % 2))
    cost = int(rng.integers(0, 12))
    Qr, wr, it = R.irls(g.QQ, g.I, cost, SIGMA, g.Q0, g.f, 8, 1e-4)
    r = O.irls(g.QQ, g.I, None, cost, SIGMA, g.Q0, g.f, 8, 1e-4, solver="lstsq")
    assert it == r.iters and O.geodesic_rms(Qr, r.Q, g.f) <= 1e-12 and wclose(wr, r.weights, 1e-7)
    Ql, itl = R.l1ra(g.QQ, g.I, g.Q0, g.f, 4, 1e-3)
    la = O.l1ra(g.QQ, g.I, None, g.Q0, g.f, 4, 1e-3)
    assert itl == la.iters and O.geodesic_rms(Ql, la.Q, g.f) <= 1e-10
    assert np.abs(R.init_mst(g.Q0, g.QQ, g.I, g.f) - O.init_mst(g.Q0, g.QQ, g.I, g.f)).max() <= 1e-13


# ============================================================================== GPU: CUDA path == reference
@pytest.mark.gpu
@pytest.mark.parametrize("cost", range(14))
def test_gpu_irls_all_costs_vs_reference(solver, ref_small, cost):
    gq = graphs()[0]
    Q, w, info = solver.irls(gq.QQ, gq.I, None, cost, sigma_for(cost), gq.Q0, gq.f, 6, -1.0)
    assert info.iters == int(ref_small[f"quirk_c{cost}_iters"])
    assert O.geodesic_rms(Q, ref_small[f"quirk_c{cost}_Q"], gq.f) <= 1e-8
    assert wclose(w, ref_small[f"quirk_c{cost}_weights"], 1e-5)


@pytest.mark.gpu
def test_gpu_rank_deficient_talwar_vs_reference(solver, ref_small):
    """The same rank-deficient case on the device: rows whose edges all have weight 0 have a zero diagonal, PCG leaves
    their x at 0 - SPQR's basic solution."""
    gq = graphs()[0]
    Q, w, info = solver.irls(gq.QQ, gq.I, None, O.TALWAR, SIGMA, gq.Q0, gq.f, 6, -1.0)
    assert O.geodesic_rms(Q, ref_small["quirk_talwar5_Q"], gq.f) <= 1e-8
    assert np.array_equal(w, ref_small["quirk_talwar5_weights"])


@pytest.mark.gpu
def test_gpu_residual_and_mst_vs_reference(solver, ref_small):
    gq, gk, _, gi = graphs()
    for g, key in ((gq, "quirk_residual"), (gi, "ident_residual")):
        solver.upload(g.QQ, g.I, g.Q0, g.f)
        assert np.abs(solver.probe_residual() - ref_small[key]).max() <= 1e-12
    assert np.abs(solver.init_mst(gq.Q0, gq.QQ, gq.I, gq.f)[0] - ref_small["quirk_mst"]).max() <= 1e-12
    perm = ref_small["kitti_perm"]
    assert np.abs(solver.init_mst(gk.Q0, gk.QQ[perm], gk.I[perm], gk.f)[0] - ref_small["kitti_mst_shuffled"]).max() <= 1e-11


@pytest.mark.gpu
def test_gpu_l1ra_and_flow_vs_reference(solver, ref_small):
    import irotavg_b200 as ira
    gq, gk, gw, gi = graphs()
    Q, info = solver.l1ra(gq.QQ, gq.I, None, gq.Q0, gq.f, 5, 1e-3)
    assert info.iters == int(ref_small["quirk_l1ra_iters"]) and O.geodesic_rms(Q, ref_small["quirk_l1ra_Q"], gq.f) <= 1e-8
    Q, info = solver.l1ra(gk.QQ, gk.I, None, gk.Q0, gk.f, 5, 1e-3)
    assert info.iters == int(ref_small["kitti_l1ra_iters"]) and O.geodesic_rms(Q, ref_small["kitti_l1ra_Q"], gk.f) <= 1e-8
    Q, w, l1_it, info = solver.l1ra_irls(gk.QQ, gk.I, gk.Q0, gk.f, 5, 1e-3, ira.Geman_McClure, SIGMA, 50, 1e-3)
    assert info.iters == int(ref_small["kitti_flow_iters"])
    assert O.geodesic_rms(O.quat_normalised(Q.copy(), gk.f), ref_small["kitti_flow_Q"], gk.f) <= 1e-8
    assert wclose(w, ref_small["kitti_flow_weights"], 1e-5)
    Q, w, info = solver.irls(gk.QQ, gk.I, None, ira.L1, SIGMA, gk.Q0, gk.f, 30, -1.0)       # 30 L1 iterations
    assert O.geodesic_rms(Q, ref_small["kitti_l1x30_Q"], gk.f) <= 1e-8
    # a rotAvg window (single-block solver) and the same through the general pipeline
    for small_path in (True, False):
        with ira.Solver(small_path=small_path) as s:
            Q, w, l1_it, info = s.l1ra_irls(gw.QQ, gw.I, gw.Q0, gw.f, 100, 1e-3, ira.Geman_McClure, SIGMA, 100, 1e-3)
        assert [l1_it, info.iters] == list(ref_small["window_iters"])
        assert O.geodesic_rms(Q, ref_small["window_Q"], gw.f) <= 1e-8
        assert wclose(w, ref_small["window_weights"], 1e-5)
    Q, w, info = solver.irls(gi.QQ, gi.I, None, ira.Geman_McClure, SIGMA, gi.Q0, gi.f, 20, 1e-3)   # identity start
    assert info.iters == int(ref_small["ident_gm_iters"]) and O.geodesic_rms(Q, ref_small["ident_gm_Q"], gi.f) <= 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("tag,args", [("default", []), ("l1", ["L1"]), ("huber", ["Huber"]), ("l2", ["L2"])])
def test_gpu_cli_vs_reference_cli_on_bundled_fixture(built_lib, ref_cli, bundled, tmp_path, tag, args):
    """Our `l1_irls` binary and the reference's `l1_irls` (golden) on the reference's own input file."""
    from irotavg_b200 import build
    cli = build.build_cli()
    inp = tmp_path / "ravg_input.txt"
    G.write_ral_text(str(inp), bundled["I"] + 1, bundled["QQ"], bundled["Q_file"][: int(bundled["n_given"])], int(bundled["f"]))
    outp = tmp_path / "out.txt"
    r = subprocess.run([cli, str(inp), str(outp)] + args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    vals = np.array(outp.read_text().split(), dtype=np.float64)
    n, m = 1832, 3655
    Q = vals[:4 * n].reshape(n, 4)[:, [1, 2, 3, 0]]
    w = vals[4 * n:]
    assert w.size == m
    assert O.geodesic_rms(Q, ref_cli[f"{tag}_Q"], 1) <= 1e-8
    assert wclose(w, ref_cli[f"{tag}_weights"], 1e-5)
