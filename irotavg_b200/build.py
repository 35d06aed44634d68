"""In-tree build of the sm_100a CUDA library (irotavg_b200/lib/libira.so) with plain nvcc.

`python -m irotavg_b200.build [--force] [--verbose]`.  nvcc cross-compiles without a GPU; the
resulting .so is git-ignored but travels with the tree to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libira.so")
STAMP = os.path.join(LIBDIR, "libira.stamp")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--shared",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "-Xptxas", "-v",
]


HOST = os.path.join(HERE, "host")
CLI = os.path.join(LIBDIR, "l1_irls")


def build_cli(force: bool = False) -> str:
    """The `l1_irls` executable (host/l1_irls_cli.cpp: the reference's ral/test.cpp flow over the adapter
    header), linked against libira.so with an rpath relative to the executable."""
    src = os.path.join(HOST, "l1_irls_cli.cpp")
    deps = [src, os.path.join(HOST, "l1_irls.hpp"), os.path.join(HOST, "ral_text_io.hpp"), os.path.join(INCLUDE, "ira.h")]
    if (not force and os.path.exists(CLI)
            and all(os.path.getmtime(CLI) >= os.path.getmtime(d) for d in deps + [LIB])):
        return CLI
    cmd = ["g++", "-std=c++11", "-O2", "-Wall", "-I", INCLUDE, "-I", HOST, src, "-o", CLI, "-L", LIBDIR, "-lira",
           "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + (res.stdout + res.stderr)[-4000:])
    return CLI


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    units = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    return [os.path.join(CSRC, u) for u in units], [os.path.join(CSRC, d) for d in deps] + [
        os.path.join(INCLUDE, "ira.h")]


def _digest(paths) -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in paths:
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    units, deps = _sources()
    digest = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                build_cli()
                return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-o", LIB, *units, "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(LIBDIR, "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    if verbose:
        print(log)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    build_cli(force=True)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
