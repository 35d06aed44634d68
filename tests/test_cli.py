"""The `l1_irls` executable (irotavg_b200/host/l1_irls_cli.cpp) against ral/test.cpp's behaviour:
argument handling and error exits on CPU; on the GPU box the whole default flow on the reference's
bundled graph (config 1) against the oracle golden, and the output file's Eigen IOFormat layout."""
import os
import subprocess

import numpy as np
import pytest

from oracle import graphs as G
from oracle import irls_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def cli(built_lib):
    from irotavg_b200 import build
    return build.build_cli()


def test_usage_and_bad_input_exit_codes(cli, tmp_path):
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 255 and "input_file" in r.stderr            # std::exit(-1), ral/test.cpp:136-150
    r = subprocess.run([cli, str(tmp_path / "missing.txt")], capture_output=True, text=True)
    assert r.returncode == 255 and "Unable to open file" in r.stderr   # ral/test.cpp:158-162
    bad = tmp_path / "bad.txt"
    bad.write_text("3 3 1\n0 1 1 0 0 0\n")
    r = subprocess.run([cli, str(bad)], capture_output=True, text=True)
    assert r.returncode == 255 and "inconsistent number of connections" in r.stderr   # ral/test.cpp:195-199
    few = tmp_path / "few.txt"
    few.write_text("1 2 1\n0 1 1 0 0 0\n")
    r = subprocess.run([cli, str(few)], capture_output=True, text=True)
    assert r.returncode == 255 and "Insuficient number of absolute rotations" in r.stderr   # ral/test.cpp:229-233
    r = subprocess.run([cli, str(few), str(tmp_path / "o.txt"), "nonsense"], capture_output=True, text=True)
    assert r.returncode == 255                                          # both exits come before any device call


def test_eigen_style_layout(tmp_path):
    """CPU: the output layout is Eigen's `operator<<` with IOFormat(FullPrecision) - %.15g entries, every entry
    right-aligned to the widest one of the whole matrix, one space between columns (ral/test.cpp:321-325)."""
    exe = str(tmp_path / "textio_main")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(root, "irotavg_b200", "host"),
                    os.path.join(root, "tests", "cpp", "textio_main.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    blocks = out.rstrip("\n").split("\n--\n")
    vals = [[1.0, 12345.678901234567], [-0.5, 2e-9], [1.0 / 3.0, -7.0]]
    cells = [[f"{v:.15g}" for v in row] for row in vals]
    width = max(len(c) for row in cells for c in row)
    assert blocks[0] == "\n".join(" ".join(c.rjust(width) for c in row) for row in cells)
    assert blocks[1] == "\n".join(" ".join(c.rjust(width) for c in row[::-1]) for row in cells)     # column order argument
    col = [f"{row[0]:.17g}" for row in vals]
    w17 = max(len(c) for c in col)
    assert blocks[2] == "\n".join(c.rjust(w17) for c in col)


def _parse_output(path, n, m):
    lines = open(path).read().split("\n")
    assert lines[-1] == "" and len(lines) == n + m + 1
    Qw = np.array([ln.split() for ln in lines[:n]], dtype=np.float64)
    w = np.array(lines[n:n + m], dtype=np.float64)
    return Qw[:, [1, 2, 3, 0]], w, lines


@pytest.mark.gpu
def test_default_flow_on_bundled_graph(cli, tmp_path):
    """Config 1: `l1_irls ravg_input.txt` with every default (Geman-McClure, 5 deg, 50 / 5 iterations, 1e-3)."""
    z = np.load(os.path.join(GOLD, "bundled_graph.npz"))
    I, QQ, f, ng = z["I"], z["QQ"], int(z["f"]), int(z["n_given"])
    n, m = int(I.max()) + 1, len(I)
    inp, outp = str(tmp_path / "in.txt"), str(tmp_path / "out.txt")
    G.write_ral_text(inp, I + 1, QQ, z["Q_file"][:ng], f)               # 1-based ids as in the reference file
    env = dict(os.environ, IRA_CLI_PRECISION="17")
    r = subprocess.run([cli, inp, outp], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert "cost: Geman-McClure" in r.stdout and "sigma [deg]: 5" in r.stdout
    assert f"L1-RA iterations = {int(z['l1ra_iters'])}" in r.stdout
    assert f"IRLS  iterations = {int(z['cli_irls_iters'])}" in r.stdout
    Q, w, _ = _parse_output(outp, n, m)
    assert O.geodesic_rms(Q, z["cli_Q"], f) <= 1e-8
    assert np.allclose(w, z["cli_weights"], rtol=1e-5, atol=1e-8)
    assert np.allclose(np.linalg.norm(Q[f:], axis=1), 1.0, atol=1e-14)

    # default precision: Eigen::FullPrecision = 15 significant digits, entries right-aligned to one width
    out15 = str(tmp_path / "out15.txt")
    r = subprocess.run([cli, inp, out15, "Geman-McClure", "5", "50", "5", "0.001"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    Q15, w15, lines = _parse_output(out15, n, m)
    assert np.allclose(Q15, Q, atol=1e-14) and np.allclose(w15, w, rtol=1e-14)
    assert len({len(ln) for ln in lines[:n]}) == 1                       # aligned columns: equal line lengths
    assert len({len(ln) for ln in lines[n:n + m]}) == 1
    width = (len(lines[0]) - 3) // 4
    assert lines[0] == " ".join(f"{float(t):.15g}".rjust(width) for t in lines[0].split())


@pytest.mark.gpu
def test_no_fixed_rotation_and_other_cost(cli, tmp_path):
    """f = 0 in the file: the first rotation is pinned to I (ral/test.cpp:277-282); Huber, 3 deg, explicit caps."""
    g = G.small_graph(n=120, extra=700, sigma_n=0.02, outlier_frac=0.1, sigma_init=0.3, seed=31)
    inp, outp = str(tmp_path / "in.txt"), str(tmp_path / "out.txt")
    G.write_ral_text(inp, g.I, g.QQ, np.zeros((0, 4)), 0)
    env = dict(os.environ, IRA_CLI_PRECISION="17")
    r = subprocess.run([cli, inp, outp, "huber", "3", "20", "4", "1e-4"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert "set first abs rot = I" in r.stdout
    Q0 = np.zeros((g.n, 4)); Q0[0] = [0, 0, 0, 1]
    Qm = O.init_mst(Q0, g.QQ, g.I, 1)
    la = O.l1ra(g.QQ, g.I, None, Qm, 1, 4, 1e-4)
    ref = O.irls(g.QQ, g.I, None, O.HUBER, 3 * np.pi / 180, la.Q, 1, 20, 1e-4, solver="direct")
    Q, w, _ = _parse_output(outp, g.n, g.m)
    assert O.geodesic_rms(Q, O.quat_normalised(ref.Q.copy(), 1), 1) <= 1e-8
    assert np.allclose(w, ref.weights, rtol=1e-5, atol=1e-8)
