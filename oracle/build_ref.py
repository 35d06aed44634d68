"""Builds oracle/_ref/ from the reference's OWN sources where they lie under /root/reference:

    oracle/_ref/l1_irls_ref     ral/test.cpp + ral/l1_irls.cpp            (the reference CLI, unmodified)
    oracle/_ref/libral_ref.so   ral/l1_irls.cpp + oracle/ref_shim/ral_ref_capi.cpp (ctypes entry points)
    oracle/_ref/rotavg_reference  ral/l1_irls.cpp + the source text of ViewGraph::rotAvg / rmat2quat / savePoses /
                                fixPose (src/ViewGraph.cpp:1175-1435, extracted into oracle/_ref/ at build time) +
                                tests/cpp/rotavg_refsrc_main.cpp over OpenCV-free containers: the reference's whole
                                CPU path for the rotAvg stream (config 5), window-sized problems only

    python -m oracle.build_ref [--force]

The reference needs Eigen >= 3.3 and SuiteSparse (SPQR, UMFPACK, CHOLMOD types); neither exists in this image
and there is no network.  Instead of the reference's CMake build, g++ compiles the two reference files directly
against oracle/ref_shim/: an eager-evaluation stand-in for the ~60 Eigen operations those files use and dense
stand-ins for SuiteSparseQR (Householder QR least squares) and UMFPACK (LU with partial pivoting).  Everything
the reference itself wrote - delta_rel, log_map, exp_map, the 14-cost switch, make_A / make_AtA, l1decode_pd,
init_mst, the IRLS / L1RA loops, the CLI's parsing and output - runs as written; no reference source is copied
into this repository.  TEST INFRASTRUCTURE ONLY; outputs are git-ignored and travel to the GPU box with the tree.
The dense stand-ins limit it to graphs of a few thousand edges (the bundled fixture: 3 655 edges, ~3 s).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "ref_shim")
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("IROTAVG_REFERENCE", "/root/reference")
RAL = os.path.join(REF, "ral")
CLI = os.path.join(OUT, "l1_irls_ref")
LIB = os.path.join(OUT, "libral_ref.so")
ROTAVG = os.path.join(OUT, "rotavg_reference")
VG_RANGES = [(1175, 1203), (1206, 1231), (1234, 1260), (1263, 1435)]   # rmat2quat, savePoses, fixPose.., rotAvg
TESTS_CPP = os.path.join(os.path.dirname(HERE), "tests", "cpp")


def reference_present() -> bool:
    return os.path.exists(os.path.join(RAL, "l1_irls.cpp")) and os.path.exists(os.path.join(RAL, "test.cpp"))


def built() -> bool:
    return os.path.exists(CLI) and os.path.exists(LIB) and os.path.exists(ROTAVG)


def _run(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("build_ref: " + " ".join(cmd) + "\n" + (res.stdout + res.stderr)[-4000:])


def build(force: bool = False) -> str:
    """Returns OUT.  No-op when up to date; raises when the reference sources are absent and nothing is built."""
    if not reference_present():
        if built():
            return OUT          # the GPU box: prebuilt files travelled with the tree
        raise RuntimeError(f"reference sources not found under {RAL} and oracle/_ref is not prebuilt")
    srcs = [os.path.join(RAL, "l1_irls.cpp"), os.path.join(RAL, "l1_irls.hpp"), os.path.join(RAL, "test.cpp"),
            os.path.join(SHIM, "ral_ref_capi.cpp"), os.path.join(TESTS_CPP, "rotavg_refsrc_main.cpp"),
            os.path.join(TESTS_CPP, "viewgraph_decl.hpp"), os.path.join(TESTS_CPP, "view_shim.hpp")] + [
        os.path.join(SHIM, f) for f in ("mini_eigen.hpp", "SuiteSparseQR.hpp", "umfpack.h", "cholmod.h")]
    if not force and built() and all(os.path.getmtime(CLI) >= os.path.getmtime(s) and
                                     os.path.getmtime(LIB) >= os.path.getmtime(s) and
                                     os.path.getmtime(ROTAVG) >= os.path.getmtime(s) for s in srcs):
        return OUT
    os.makedirs(OUT, exist_ok=True)
    cxx = ["g++", "-std=c++11", "-O2", "-fopenmp", "-fPIC", "-w", "-I", SHIM, "-I", RAL]
    objs = {}
    for name, src in (("l1_irls", os.path.join(RAL, "l1_irls.cpp")), ("test", os.path.join(RAL, "test.cpp")),
                      ("capi", os.path.join(SHIM, "ral_ref_capi.cpp"))):
        objs[name] = os.path.join(OUT, name + ".o")
        _run(cxx + ["-c", src, "-o", objs[name]])
    _run(["g++", "-fopenmp", "-o", CLI, objs["test"], objs["l1_irls"]])
    _run(["g++", "-fopenmp", "-shared", "-o", LIB, objs["capi"], objs["l1_irls"]])
    vg = os.path.join(REF, "src", "ViewGraph.cpp")
    if os.path.exists(vg):
        with open(vg) as fh:
            lines = fh.readlines()
        frag = ['#include "viewgraph_decl.hpp"\nusing namespace irotavg;\n']
        for lo, hi in VG_RANGES:
            frag.append(f"// ---- src/ViewGraph.cpp:{lo}-{hi} (extracted at build time, unmodified) ----\n")
            frag.extend(lines[lo - 1:hi])
            frag.append("\n")
        fpath = os.path.join(OUT, "viewgraph_ref_fragment.cpp")
        with open(fpath, "w") as fh:
            fh.writelines(frag)
        # include order: the REFERENCE's l1_irls.hpp (RAL) is found before anything else
        _run(["g++", "-std=c++11", "-O2", "-fopenmp", "-w", "-I", RAL, "-I", SHIM, "-I", TESTS_CPP, fpath,
              os.path.join(TESTS_CPP, "rotavg_refsrc_main.cpp"), objs["l1_irls"], "-o", ROTAVG])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
