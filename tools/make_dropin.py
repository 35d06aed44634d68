#!/usr/bin/env python
"""Builds the two LITERAL drop-in binaries (only where /root/reference exists; outputs under irotavg_b200/lib/,
git-ignored, they travel to the GPU box with the tree):

  irotavg_b200/lib/l1_irls_refmain   the reference's ral/test.cpp, unmodified, compiled against
                                     irotavg_b200/host/l1_irls.hpp (instead of ral/l1_irls.hpp) + libira.so
  irotavg_b200/lib/rotavg_refsrc     the reference's own ViewGraph::rotAvg / rmat2quat / savePoses / fixPose source
                                     text (src/ViewGraph.cpp:1175-1435) over tests/cpp/viewgraph_decl.hpp, the same
                                     adapter header and libira.so, driven by tests/cpp/rotavg_refsrc_main.cpp

Eigen is not in this image: oracle/ref_shim's stand-in headers play <Eigen/Dense> for these two translation units
(types only - every solve goes to the CUDA library).  The reference text is copied to / extracted into the build
directory at build time and is never committed.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("IROTAVG_REFERENCE", "/root/reference")
LIBDIR = os.path.join(ROOT, "irotavg_b200", "lib")
WORK = os.path.join(LIBDIR, "dropin")
RANGES = [(1175, 1203), (1206, 1231), (1234, 1260), (1263, 1435)]   # rmat2quat, savePoses, fixPose.., rotAvg
CLI = os.path.join(LIBDIR, "l1_irls_refmain")
ROTAVG = os.path.join(LIBDIR, "rotavg_refsrc")


def reference_present():
    return os.path.exists(os.path.join(REF, "ral", "test.cpp")) and os.path.exists(os.path.join(REF, "src", "ViewGraph.cpp"))


def built():
    return os.path.exists(CLI) and os.path.exists(ROTAVG)


def _run(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("make_dropin: " + " ".join(cmd) + "\n" + (res.stdout + res.stderr)[-6000:])


def build(force=False):
    if not reference_present():
        if built():
            return LIBDIR
        raise RuntimeError("reference sources not found and the drop-in binaries are not prebuilt")
    if not os.path.exists(os.path.join(LIBDIR, "libira.so")):
        raise RuntimeError("build libira.so first (python -m irotavg_b200.build)")
    os.makedirs(WORK, exist_ok=True)
    shutil.copyfile(os.path.join(REF, "ral", "test.cpp"), os.path.join(WORK, "ref_test.cpp"))
    with open(os.path.join(REF, "src", "ViewGraph.cpp")) as fh:
        lines = fh.readlines()
    frag = ['#include "viewgraph_decl.hpp"\nusing namespace irotavg;\n']
    for lo, hi in RANGES:
        frag.append(f"// ---- src/ViewGraph.cpp:{lo}-{hi} (extracted, unmodified) ----\n")
        frag.extend(lines[lo - 1:hi])
        frag.append("\n")
    with open(os.path.join(WORK, "viewgraph_ref_fragment.cpp"), "w") as fh:
        fh.writelines(frag)
    inc = ["-I", os.path.join(ROOT, "irotavg_b200", "host"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "oracle", "ref_shim"), "-I", os.path.join(ROOT, "tests", "cpp")]
    link = ["-L", LIBDIR, "-lira", "-Wl,-rpath,$ORIGIN"]
    cxx = ["g++", "-std=c++11", "-O2", "-w"]
    _run(cxx + inc + [os.path.join(WORK, "ref_test.cpp"), "-o", CLI] + link)
    _run(cxx + inc + [os.path.join(WORK, "viewgraph_ref_fragment.cpp"), os.path.join(ROOT, "tests", "cpp", "rotavg_refsrc_main.cpp"),
                      "-o", ROTAVG] + link)
    return LIBDIR


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
