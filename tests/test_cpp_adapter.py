"""The header-only C++ adapter (irotavg_b200/host/l1_irls.hpp) keeps the reference's signatures:
a C++ program written like ral/test.cpp's call sequence builds against it (CPU) and, on the GPU box,
reproduces the oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle import graphs as G
from oracle import irls_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "adapter_main.cpp")


def _build(tmp_path, built_lib):
    exe = str(tmp_path / "adapter_main")
    libdir = os.path.dirname(built_lib)
    subprocess.run(["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I",
                    os.path.join(ROOT, "irotavg_b200", "host"), SRC, "-o", exe, "-L", libdir, "-lira",
                    f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def test_adapter_compiles_as_cxx11(tmp_path, built_lib):
    assert os.path.exists(_build(tmp_path, built_lib))


@pytest.mark.gpu
def test_adapter_matches_oracle(tmp_path, built_lib):
    exe = _build(tmp_path, built_lib)
    g = G.small_graph(n=150, extra=900, sigma_n=0.02, outlier_frac=0.1, sigma_init=0.3, seed=17, f=2)
    inp, outp = str(tmp_path / "in.txt"), str(tmp_path / "out.txt")
    G.write_ral_text(inp, g.I, g.QQ, g.Q0, g.f)
    sigma = 5 * np.pi / 180
    subprocess.run([exe, inp, outp, str(O.GEMAN_MCCLURE), repr(sigma), "50", "1e-3"], check=True)
    tok = open(outp).read().split()
    iters = int(tok[0])
    vals = np.array(tok[1:], dtype=np.float64)
    Qw = vals[:4 * g.n].reshape(g.n, 4)
    Q = Qw[:, [1, 2, 3, 0]]
    w = vals[4 * g.n:]
    ref = O.irls(g.QQ, g.I, None, O.GEMAN_MCCLURE, sigma, g.Q0, g.f, 50, 1e-3, solver="direct")
    assert iters == ref.iters
    assert O.geodesic_rms(Q, O.quat_normalised(ref.Q.copy(), g.f), g.f) <= 1e-8
    assert np.allclose(w, ref.weights, rtol=1e-5, atol=1e-8)
    assert np.allclose(np.linalg.norm(Q[g.f:], axis=1), 1.0, atol=1e-14)
