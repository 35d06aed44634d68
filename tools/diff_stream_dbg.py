#!/usr/bin/env python
"""First line at which two IRA_DEBUG logs of the same stream replay differ (timings stripped).

    DBG=2 SOLVERS=0 RUNS=3 FRAMES=4002 bash tools/stream_repeat.sh      # writes gpurun_out/stream_dbg_<solver>_<run>.err
    python tools/diff_stream_dbg.py gpurun_out/stream_dbg_0_1.err gpurun_out/stream_dbg_0_2.err

With DBG=2 every Newton solve of every l1ra iteration prints its PCG iteration count and residual norms, so the first
differing line names the solve at which two replays part (the open repeatability issue of DESIGN 4.5)."""
import re
import sys


def lines(path):
    out = []
    for ln in open(path, errors="replace"):
        if ln.startswith("[ira]"):
            out.append(re.sub(r"; [0-9.]+ ms\s*$", "", ln.rstrip()))
    return out


def main():
    a, b = lines(sys.argv[1]), lines(sys.argv[2])
    print(len(a), "and", len(b), "lines")
    for k, (x, y) in enumerate(zip(a, b)):
        if x != y:
            calls = sum(1 for ln in a[:k] if "l1ra_irls n" in ln)
            print(f"first difference at line {k} (after {calls} completed global calls):\n  {x}\n  {y}")
            return 1
    print("identical" if len(a) == len(b) else "one log is a prefix of the other")
    return 0


if __name__ == "__main__":
    sys.exit(main())
