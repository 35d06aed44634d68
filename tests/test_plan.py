"""Host-side planning (irotavg_b200/csrc/ira_plan.hpp): which coarse space a graph gets and how SELL slices are dealt to
thread blocks.  Pure C++, compiled with g++ and run on the CPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_decisions(tmp_path):
    exe = str(tmp_path / "plan_test")
    subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "irotavg_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "plan_test.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for name in ("tri_block_rows: chains", "tri_block_rows: rejections", "dense_partition", "lpt_slice_map"):
        assert "ok " + name in r.stdout
