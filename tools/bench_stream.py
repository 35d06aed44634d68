#!/usr/bin/env python
"""Config 5 (BASELINE.json configs[4]): sequential ViewGraph::rotAvg window calls on a growing graph.

    python tools/bench_stream.py [--frames 10000] [--loop-every 500] [--cpu-frames 150]

Generates the op list (oracle/rotavg_stream.py: per frame a view, edges to the previous <= 4 views, rotAvg(10);
every --loop-every frames a loop edge and rotAvg(5000000), src/IRotAvg.cpp:371-378), replays it through the C++
host mirror (tests/cpp/rotavg_main.cpp -> irotavg_b200/host/view_graph_rotavg.hpp -> libira.so) and prints one
JSON line: calls/s, p50 / p99 latency split local / global, plus the oracle's replay time on the first
--cpu-frames frames (CPU baseline, numpy/scipy restatement, 1 thread) and the rotation difference on that prefix."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(frames=10000, loop_every=500, cpu_frames=150):
    class A:
        pass
    args = A()
    args.frames, args.loop_every, args.cpu_frames = frames, loop_every, cpu_frames
    from irotavg_b200 import build
    from oracle import irls_oracle as O
    from oracle import rotavg_stream as RS
    lib = build.build()
    libdir = os.path.dirname(lib)
    tmp = tempfile.mkdtemp()
    exe = os.path.join(tmp, "rotavg_main")
    subprocess.run(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"), "-I",
                    os.path.join(ROOT, "irotavg_b200", "host"), "-I", os.path.join(ROOT, "tests", "cpp"),
                    os.path.join(ROOT, "tests", "cpp", "rotavg_main.cpp"), "-o", exe, "-L", libdir, "-lira",
                    f"-Wl,-rpath,{libdir}"], check=True)
    ops, Qgt = RS.make_stream(n_frames=args.frames, loop_every=args.loop_every, min_loop_gap=min(500, args.loop_every))
    inp, outp = os.path.join(tmp, "ops.txt"), os.path.join(tmp, "out.txt")
    RS.write_ops(inp, ops)
    san = os.environ.get("IRA_STREAM_SANITIZE")              # e.g. initcheck: replay under compute-sanitizer, print its report
    if san:
        r = subprocess.run(["compute-sanitizer", "--tool", san, "--print-limit", "30", exe, inp, outp], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
        print(r.stdout[-6000:])
        return {"sanitize": san, "returncode": r.returncode}
    t0 = time.perf_counter()
    subprocess.run([exe, inp, outp], check=True)
    wall = time.perf_counter() - t0
    tok = open(outp).read().split()
    nv, nc = int(tok[0]), int(tok[1])
    R = np.array(tok[2:2 + 9 * nv], dtype=np.float64).reshape(nv, 3, 3)
    calls = np.array(tok[2 + 9 * nv:], dtype=np.float64).reshape(nc, 8)
    solved = calls[:, 1] > 0
    glob = calls[:, 0] > 1000
    lat = calls[:, 7] * 1e3

    def stats(mask):
        v = lat[mask & solved]
        if v.size == 0:
            return None
        return {"calls": int(v.size), "p50_ms": float(np.percentile(v, 50)), "p99_ms": float(np.percentile(v, 99)),
                "mean_ms": float(v.mean()), "max_ms": float(v.max()), "calls_over_1ms": int((v > 1.0).sum()),
                "max_vertices": int(calls[mask & solved, 2].max()),
                "max_edges": int(calls[mask & solved, 3].max())}

    Q = np.array([O.rmat2quat(r) for r in R])
    import hashlib
    result_sha1 = hashlib.sha1(np.ascontiguousarray(R).tobytes()).hexdigest()[:12]
    line = {"metric": "rotAvg window calls/s on a growing graph (config 5)", "value": float(solved.sum() / lat[solved].sum() * 1e3),
            "unit": "calls/s", "frames": args.frames, "loop_every": args.loop_every, "wall_s_incl_parsing": wall,
            "local": stats(~glob), "global": stats(glob),
            "global_calls": [{"views": int(c[2]), "edges": int(c[3]), "l1_iters": int(c[5]), "irls_iters": int(c[6]),
                              "ms": float(c[7] * 1e3)} for c in calls[glob & solved]],
            "geodesic_rms_vs_ground_truth_rad": float(O.geodesic_rms(Q, Qgt, 1)), "result_sha1": result_sha1}
    # the reference's own CPU path on the same stream: oracle/_ref/rotavg_reference = the source text of
    # ViewGraph::rotAvg + ral/l1_irls.cpp compiled by oracle/build_ref.py (dense stand-in solvers: window-sized problems
    # only, so the prefix before the first global call)
    try:
        from oracle import build_ref
        if build_ref.built():
            sub = []
            for op in ops:
                if op[0] == "A" and op[1] > 1000:
                    break
                sub.append(op)
            inp2, out2 = os.path.join(tmp, "ops_ref.txt"), os.path.join(tmp, "out_ref.txt")
            RS.write_ops(inp2, sub)
            env = dict(os.environ, OMP_NUM_THREADS="1")
            subprocess.run([build_ref.ROTAVG, inp2, out2, os.path.join(tmp, "poses_ref.txt")], check=True, env=env,
                           stdout=subprocess.DEVNULL)
            tk = open(out2).read().split()
            nv2, nc2 = int(tk[0]), int(tk[1])
            c2 = np.array(tk[2 + 9 * nv2:], dtype=np.float64).reshape(nc2, 2)
            ms2 = c2[:, 1] * 1e3
            ours_same = lat[:nc2][solved[:nc2]]
            line["reference_build_local_calls"] = {
                "kind": "reference", "cores": 1, "calls": int(nc2), "p50_ms": float(np.percentile(ms2[2:], 50)),
                "p99_ms": float(np.percentile(ms2[2:], 99)),
                "ours_p50_ms_same_calls": float(np.percentile(ours_same, 50)) if ours_same.size else None,
                "sample": f"the {nc2} local rotAvg(10) calls before the first loop closure; the reference's rotAvg + ral sources "
                          "(oracle/_ref, -O2, dense stand-ins for SPQR / UMFPACK, eager stand-in for Eigen)",
                "parity": "tests/test_rotavg.py::test_oracle_stream_vs_reference_build and test_stream_matches_oracle hold both "
                          "sides to the same rotations"}
    except Exception as e:                                   # noqa: BLE001
        line["reference_build_local_calls"] = {"error": repr(e)[:200]}
    if args.cpu_frames > 0:
        k = 0
        sub = []
        for op in ops:                                   # prefix of the same stream
            sub.append(op)
            if op[0] == "A":
                k += 1
                if k >= args.cpu_frames:
                    break
        t0 = time.perf_counter()
        Rref, reps = RS.replay(sub)
        dt = time.perf_counter() - t0
        ns = sum(1 for r in reps if r["solved"])
        line["cpu_baseline"] = {"value": ns / dt, "unit": "calls/s", "cores": 1, "kind": "port",
                                "sample": f"first {args.cpu_frames} calls of the same stream, oracle/rotavg_stream.py"}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--loop-every", type=int, default=500)
    ap.add_argument("--cpu-frames", type=int, default=150)
    args = ap.parse_args()
    print(json.dumps(run(args.frames, args.loop_every, args.cpu_frames)))


if __name__ == "__main__":
    main()
