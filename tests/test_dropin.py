"""The literal drop-in: the reference's OWN caller sources on top of this library.

tools/make_dropin.py compiles (where /root/reference exists; the binaries travel to the GPU box prebuilt)
  * ral/test.cpp, unmodified, against irotavg_b200/host/l1_irls.hpp instead of ral/l1_irls.hpp  -> l1_irls_refmain
  * the source text of ViewGraph::rotAvg / rmat2quat / savePoses / fixPose (src/ViewGraph.cpp:1175-1435) against the
    same header and OpenCV-free containers                                                       -> rotavg_refsrc
so `irotavg::Mat/Vec/SpMat/Quat/I_t/Cost`, `make_A`, `init_mst`, `l1ra`, `irls`, `quat_normalised` are exercised
through the exact expressions the reference's callers use.  CPU: they build, and the CLI's argument errors behave as
in the reference.  GPU: their results equal the reference's own goldens / the oracle's rotAvg replay."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import graphs as G
from oracle import irls_oracle as O
from oracle import rotavg_stream as RS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def dropin(built_lib):
    import make_dropin
    if not make_dropin.reference_present() and not make_dropin.built():
        pytest.skip("no /root/reference here and no prebuilt drop-in binaries")
    make_dropin.build()
    return make_dropin


def test_reference_callers_compile_against_the_adapter(dropin):
    assert os.path.exists(dropin.CLI) and os.path.exists(dropin.ROTAVG)
    r = subprocess.run([dropin.CLI], capture_output=True, text=True)
    assert r.returncode == 255 and "input_file" in r.stdout          # ral/test.cpp:136-140: usage, std::exit(-1)
    r = subprocess.run([dropin.CLI, "/nonexistent/input.txt"], capture_output=True, text=True)
    assert r.returncode == 255 and "Unable to open file" in r.stderr   # ral/test.cpp:158-162


def test_reference_rotavg_source_takes_the_early_returns(dropin, tmp_path):
    """No solve is reached (fewer than 2 views / too few edges): runs without a device."""
    ops = [("V",), ("A", 10), ("V",), ("A", 10), ("V",), ("E", 0, 1, np.eye(3)), ("A", 10)]
    inp, outp, poses = str(tmp_path / "ops.txt"), str(tmp_path / "out.txt"), str(tmp_path / "poses.txt")
    RS.write_ops(inp, ops)
    subprocess.run([dropin.ROTAVG, inp, outp, poses], check=True)
    tok = open(outp).read().split()
    assert int(tok[0]) == 3 and np.allclose(np.array(tok[2:29], dtype=float).reshape(3, 3, 3), np.eye(3))
    lines = open(poses).read().splitlines()                            # savePoses: id \t qw qx qy qz \t tx ty tz
    assert len(lines) == 3 and lines[1].split("\t")[:2] == ["1", "1.00000000000000000e+00"]
    # the host mirror's save_poses writes the same bytes as the reference's savePoses
    exe = str(tmp_path / "rotavg_main")
    libdir = os.path.join(ROOT, "irotavg_b200", "lib")
    subprocess.run(["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I",
                    os.path.join(ROOT, "irotavg_b200", "host"), "-I", os.path.join(ROOT, "tests", "cpp"),
                    os.path.join(ROOT, "tests", "cpp", "rotavg_main.cpp"), "-o", exe, "-L", libdir, "-lira",
                    f"-Wl,-rpath,{libdir}"], check=True)
    ops2 = ops + [("F", 1, O.quat2rmat(np.array([0.1, -0.7, 0.2, 0.68]) / np.linalg.norm([0.1, -0.7, 0.2, 0.68])))]
    RS.write_ops(inp, ops2)
    mine = str(tmp_path / "poses_mine.txt")
    subprocess.run([exe, inp, str(tmp_path / "o2.txt"), mine], check=True)
    subprocess.run([dropin.ROTAVG, inp, outp, poses], check=True)
    assert open(mine, "rb").read() == open(poses, "rb").read()


@pytest.mark.gpu
@pytest.mark.parametrize("tag,args", [("default", []), ("l1", ["L1"])])
def test_reference_cli_source_on_this_library_vs_reference_golden(dropin, tmp_path, tag, args):
    ref = np.load(os.path.join(GOLD, "ref_bundled_cli.npz"))
    b = np.load(os.path.join(GOLD, "bundled_graph.npz"))
    inp = tmp_path / "ravg_input.txt"
    G.write_ral_text(str(inp), b["I"] + 1, b["QQ"], b["Q_file"][: int(b["n_given"])], int(b["f"]))
    outp = tmp_path / "out.txt"
    r = subprocess.run([dropin.CLI, str(inp), str(outp)] + args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    vals = np.array(outp.read_text().split(), dtype=np.float64)
    n, m = 1832, 3655
    Q = vals[:4 * n].reshape(n, 4)[:, [1, 2, 3, 0]]
    assert O.geodesic_rms(Q, ref[f"{tag}_Q"], 1) <= 1e-8
    assert np.allclose(vals[4 * n:], ref[f"{tag}_weights"], rtol=1e-5, atol=1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["loops", "fixes"])
def test_reference_rotavg_source_on_this_library_vs_oracle(dropin, tmp_path, variant):
    if variant == "loops":
        ops, _ = RS.make_stream(n_frames=120, loop_every=40, min_loop_gap=20)
    else:
        ops, _ = RS.make_stream(n_frames=90, loop_every=0, fix_every=7, seed=5)
    Rref, reps = RS.replay(ops)
    inp, outp, poses = str(tmp_path / "ops.txt"), str(tmp_path / "out.txt"), str(tmp_path / "poses.txt")
    RS.write_ops(inp, ops)
    subprocess.run([dropin.ROTAVG, inp, outp, poses], check=True)
    tok = open(outp).read().split()
    nv = int(tok[0])
    R = np.array(tok[2:2 + 9 * nv], dtype=np.float64).reshape(nv, 3, 3)
    Q = np.array([O.rmat2quat(r) for r in R])
    Qr = np.array([O.rmat2quat(r) for r in Rref])
    assert O.geodesic_rms(Q, Qr, 0) <= 1e-8
    # savePoses text (src/ViewGraph.cpp:1206-1231) = rmat2quat of the same poses, 17 digits, scientific
    rows = np.array([l.split("\t")[1:5] for l in open(poses).read().splitlines()], dtype=np.float64)
    assert np.abs(rows[:, [1, 2, 3, 0]] - Q).max() <= 1e-15
