"""Builds the C restatement (oracle/irls_oracle.c) into oracle/_build/libirls_oracle.so with gcc.
Test / bench infrastructure only (see oracle/irls_oracle.py).  The reference itself is built by
oracle/build_ref.py into oracle/_ref (small graphs only: its sparse solvers are dense stand-ins)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "irls_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libirls_oracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
