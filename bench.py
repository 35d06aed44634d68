#!/usr/bin/env python
"""bench.py - IRLS iters/sec on the 1M-edge SO(3) graph (BASELINE.json metric, configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cost L1]

One *step* = one irls() call of 30 IRLS iterations (max_iters = 30, change_th = -1 so that all 30 run;
ral/l1_irls.cpp:590) on the synthetic random SO(3) graph n = 100 000 / m = 1 000 000 (seed 20190319).
  value      IRLS iterations / second, graph already resident in HBM (ira_irls_resident), device time from CUDA
             events recorded on the library's launch stream, max over ranks.
  e2e        the same metric through the host-buffer C-ABI call ira_irls(): pinned host buffers in, H2D of
             (I, QQ, Q), CSR/SELL build, 30 iterations, D2H of (Q, weights) inside the timed region.
  roofline   the dominant kernel = the persistent PCG kernel (one launch = one linear solve, >85 % of the step):
             algorithmic bytes K (16 m + 224 n) per launch (SpMV 16m + 48n and CG vector work 176n per
             PCG iteration, K = PCG iterations of that solve) over its mean launch duration from CUDA events recorded
             on the launch stream around every launch of a profiled step, against MEASURED_PEAKS.json's HBM copy
             bandwidth; the SpMV-only figure and the residual kernel (L2-warm and after an L2 flush) beside it.
  cpu_baseline  the C restatement (oracle/irls_oracle.c) on the host cores, all cores and one core, on a bounded
             sample of the same call; `config1_bundled_cli.reference_build` times oracle/_ref (the reference's own
             ral/ sources) on the reference's bundled graph (configs[0]) beside this library on the same input.
  parity     geodesic RMS of the timed run's final rotations against the committed 30-iteration golden of the C
             restatement (tests/golden/cfg3_l1_30iters.npz) and of a 2-iteration run against the live oracle.
N > 1 (torchrun, one rank per GPU): WEAK scaling - the graph grows with the job, n = 100 000 N / m = 1 000 000 N
(same recipe and seed); rows of the normal equations partitioned over the ranks, one persistent kernel per rank
exchanging through NVLink peer memory.  `value` = IRLS iterations/s x N (units of the 1M-edge graph), plain IRLS
iterations/s on the N-times graph is `irls_iters_per_s_on_this_graph`.  Every N > 1 line carries the geodesic RMS of
the N-GPU result against ONE GPU on the same graph and against the oracle (2 iterations), the strong-scaling figure
of the 1M-edge graph and - at N = 8 - configs[3] at its stated size (n = 1M, m = 10M).
`--impl reference` times the CPU restatement alone on the same workload (the reference's own solver stack - SPQR -
cannot hold this graph: ~40 GB of fill; oracle/_ref, the reference's ral/ compiled against dense stand-ins, is limited
to a few thousand edges), with the SAME bounded sample at every N: the first 10 IRLS iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_NODES = 100_000
M_EDGES = 1_000_000
IRLS_ITERS = 30
SIGMA = 5 * np.pi / 180.0
COSTS = {"L2": 0, "L1": 1, "L1.5": 2, "L0.5": 3, "Geman-McClure": 4, "Huber": 5}
METRIC = "IRLS iters/sec on 1M-edge SO(3) graph"
UNIT = "irls_iters/s"
REF_SAMPLE_ITERS = 10           # --impl reference: the same iteration range at every N


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_graph(scale=1):
    from oracle import graphs as G
    return G.random_graph(n=N_NODES * scale, m=M_EDGES * scale)


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_port_run(g, cost, iters, threads=0):
    """C restatement on `threads` cores (0 = all).  Returns (irls_iters_per_s, cores, result dict)."""
    from oracle import cport
    threads = threads or (os.cpu_count() or 1)
    t0 = time.perf_counter()
    r = cport.irls(g.QQ, g.I, cost, SIGMA, g.Q0, g.f, iters, -1.0, cg_rtol=1e-10, threads=threads)
    dt = time.perf_counter() - t0
    return iters / dt, threads, r


def cpu_sample_text(iters, threads):
    return (f"first {iters} of the 30 IRLS iterations of the same call (the later, costlier ones are left out to bound the "
            f"run), C restatement oracle/irls_oracle.c, OpenMP {threads} thread(s), Jacobi-PCG rtol 1e-10; the reference's "
            "SuiteSparseQR solve cannot run this graph (~40 GB fill) and oracle/_ref is dense (few thousand edges)")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scale = max(1, args.gpus)                      # the same workload as our arm at N GPUs (weak scaling: graph x N)
    g = make_graph(scale)
    cost = COSTS[args.cost]
    for _ in range(min(args.warmup, 1)):
        cpu_port_run(g, cost, 1)
    cores = os.cpu_count() or 1
    cg = []
    t_all = time.perf_counter()
    for _ in range(args.steps):
        _, cores, r = cpu_port_run(g, cost, REF_SAMPLE_ITERS)
        cg = list(r["cg_iters"])
    dt = time.perf_counter() - t_all
    value = scale * REF_SAMPLE_ITERS * args.steps / dt          # units of the 1M-edge graph, like our arm
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, scale),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": cpu_sample_text(REF_SAMPLE_ITERS, cores) + "; the same 10-iteration sample at every N",
                         "cg_iters": cg},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, scale, world=None):
    world = scale if world is None else world
    tag = "configs[2]" if scale == 1 else f"configs[2] recipe scaled x{scale} (weak scaling; x10 on 8 GPUs = configs[3])"
    return {
        "workload": f"{tag}: synthetic random SO(3) graph n={N_NODES * scale} m={M_EDGES * scale} (path + uniform pairs, "
                    f"sigma_n=0.05 rad, 10% outliers, seed 20190319), {IRLS_ITERS} IRLS iters per step, cost {args.cost}, "
                    "sigma 5 deg, f=1",
        "cost": args.cost, "irls_iters_per_step": IRLS_ITERS, "cg_rtol": 1e-10,
        "parallelism": "single GPU" if world == 1 else (
            f"rows of A^T D^2 A partitioned over {world} ranks, one persistent PCG kernel per rank exchanging u slices / "
            "dot products through NVLink peer memory (CUDA IPC); edge kernels replicated"
            if getattr(args, "shard_mode", 0) == 1 else
            f"edges sharded over {world} ranks, NCCL all-reduce of node vectors per PCG iteration"),
        "l2": "512 MB memset between timed steps flushes L2 (working set ~150 MB also exceeds the 126 MB L2)",
    }


# ------------------------------------------------------------------------------------------------
# extra legs of the N = 1 line (each guarded: the headline line must never depend on them)
# ------------------------------------------------------------------------------------------------
def guarded(fn):
    try:
        return fn()
    except Exception as e:                                   # noqa: BLE001
        return {"error": repr(e)[:300]}


def leg_config2(ira, local_rank):
    """configs[1]: KITTI-00-scale view graph (n = 4 541, m = 50 000), 30 IRLS iterations, L1 and Geman-McClure,
    beside the oracle's sparse DIRECT solve of the normal equations on one host core ('restatement, direct' -
    the formulation the reference's SPQR is comparable with at this size)."""
    from oracle import graphs as G
    from oracle import irls_oracle as O
    g = G.kitti_like_graph()
    out = {"workload": f"configs[1]: kitti_like_graph n={g.n} m={g.m}, {IRLS_ITERS} IRLS iterations, change_th=-1"}
    with ira.Solver(device=local_rank) as s:
        s.upload(g.QQ, g.I, g.Q0, g.f)
        for nm in ("L1", "Geman-McClure"):
            s.irls_resident(COSTS[nm], SIGMA, IRLS_ITERS, -1.0)
            best, info = None, None
            for _ in range(3):
                info = s.irls_resident(COSTS[nm], SIGMA, IRLS_ITERS, -1.0)
                best = info.device_ms if best is None else min(best, info.device_ms)
            Q, _ = s.download()
            t0 = time.perf_counter()
            ref = O.irls(g.QQ, g.I, None, COSTS[nm], SIGMA, g.Q0, g.f, IRLS_ITERS, -1.0, solver="direct")
            cpu_s = time.perf_counter() - t0
            out[nm] = {"value": IRLS_ITERS / (best / 1e3), "unit": UNIT, "ms_per_step": best,
                       "cg_iters_per_step": int(sum(info.cg_iters)),
                       "cpu_baseline": {"value": IRLS_ITERS / cpu_s, "unit": UNIT, "cores": 1, "kind": "port",
                                        "sample": "all 30 iterations, oracle/irls_oracle.py with scipy's sparse LU of A^T D^2 A "
                                                  "(restatement, direct)"},
                       "geodesic_rms_vs_oracle_30iters_rad": float(O.geodesic_rms(Q, ref.Q, g.f))}
    return out


def leg_bundled(ira, local_rank):
    """configs[0]: the reference's own CLI flow on its bundled graph - oracle/_ref (the reference's ral/ sources) on the
    host cores beside this library's CLI binary on the GPU, same input file, outputs compared."""
    from oracle import build_ref, refbin
    from oracle import graphs as G
    from oracle import irls_oracle as O
    from irotavg_b200 import build
    b = np.load(os.path.join(ROOT, "tests", "golden", "bundled_graph.npz"))
    td = tempfile.mkdtemp()
    inp = os.path.join(td, "ravg_input.txt")
    G.write_ral_text(inp, b["I"] + 1, b["QQ"], b["Q_file"][: int(b["n_given"])], int(b["f"]))
    cli = build.build_cli()
    ours = []
    for _ in range(3):
        t0 = time.perf_counter()
        subprocess.run([cli, inp, os.path.join(td, "ours.txt")], check=True, capture_output=True)
        ours.append(time.perf_counter() - t0)
    n, m = 1832, 3655
    Qo, _ = refbin.read_cli_output(os.path.join(td, "ours.txt"), n, m)
    out = {"workload": "configs[0]: ral/data/ravg_input.txt (n=1832, m=3655), CLI default flow init_mst -> l1ra(5) -> "
                       "irls(Geman-McClure, 50) -> quat_normalised, whole process incl. file I/O and CUDA context creation",
           "ours_process_s": min(ours)}
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_bundled_cli.npz"))
    out["geodesic_rms_vs_reference_golden_rad"] = float(O.geodesic_rms(Qo, gold["default_Q"], 1))
    if build_ref.built():
        t0 = time.perf_counter()
        r = refbin.cli([inp, os.path.join(td, "ref.txt")])
        dt = time.perf_counter() - t0
        if r.returncode == 0:
            Qr, _ = refbin.read_cli_output(os.path.join(td, "ref.txt"), n, m)
            out["reference_build"] = {"process_s": dt, "kind": "reference", "cores": os.cpu_count(),
                                      "reported": [l for l in r.stdout.splitlines() if "runtime" in l],
                                      "note": "oracle/_ref = ral/test.cpp + ral/l1_irls.cpp compiled unmodified; its SPQR / "
                                              "UMFPACK are DENSE stand-ins (oracle/ref_shim), so this time is not SuiteSparse's",
                                      "geodesic_rms_ours_vs_live_reference_rad": float(O.geodesic_rms(Qo, Qr, 1))}
    with ira.Solver(device=local_rank) as s:             # the solve alone through the library call (host buffers)
        args = (b["QQ"], b["I"], b["Q_mst"], int(b["f"]), 5, 1e-3, COSTS["Geman-McClure"], SIGMA, 50, 1e-3)
        s.l1ra_irls(*args)
        t0 = time.perf_counter()
        s.l1ra_irls(*args)
        out["ours_l1ra_irls_call_ms"] = 1e3 * (time.perf_counter() - t0)
    return out


def leg_aux(ira, local_rank, g):
    """init_mst on the headline graph (device) beside one host core running the literal sweep (C restatement)."""
    from oracle import irls_oracle as O
    with ira.Solver(device=local_rank) as s:
        s.upload(g.QQ, g.I, g.Q0, g.f)
        s.init_mst_resident(g.f)
        st = s.init_mst_resident(g.f)
    t0 = time.perf_counter()
    O.init_mst(g.Q0, g.QQ, g.I, g.f)
    cpu = time.perf_counter() - t0
    return {"init_mst_device_ms": st["ms"], "passes_label": st["passes_label"], "passes_propagate": st["passes_propagate"],
            "init_mst_oracle_1core_ms": 1e3 * cpu}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import irotavg_b200 as ira

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the library has no CPU path (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    gscale = args.graph_scale or world
    g = make_graph(gscale)
    cost = COSTS[args.cost]
    m, n, f = g.m, g.n, g.f
    from irotavg_b200.sharding import broadcast_unique_id, edge_shard
    # N > 1: rows of the normal equations partitioned over the ranks, one persistent kernel per rank exchanging
    # through NVLink peer memory (shard_mode 1, every rank holds the graph).  If CUDA IPC between the ranks is not
    # available on the box, fall back - on all ranks together - to edge shards + NCCL all-reduce (shard_mode 0).
    shard_mode = 1 if world > 1 and not args.nccl_allreduce else 0
    s = None
    while True:
        if shard_mode == 1:
            lo, hi = 0, m
        else:
            lo, hi = edge_shard(m, world, rank)                      # this rank's edge shard
        I_loc = np.ascontiguousarray(g.I[lo:hi])
        QQ_loc = np.asfortranarray(g.QQ[lo:hi])
        m_loc = hi - lo
        s = ira.Solver(device=local_rank, world_size=world, rank=rank, shard_mode=shard_mode)
        if world == 1:
            break
        s.comm_init(broadcast_unique_id(dist, ira.Solver, rank, device="cuda"))
        bad = 0
        if shard_mode == 1:
            try:
                s.upload(QQ_loc, I_loc, g.Q0, f)
            except ira.IraError as e:
                print(f"[bench] rank {rank}: peer-memory set-up failed: {e}", file=sys.stderr, flush=True)
                bad = 1
        t = torch.tensor([bad], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if int(t.item()) == 0:
            break
        s.close()
        shard_mode = 0
    args.shard_mode = shard_mode
    ext = torch.cuda.ExternalStream(s.stream_ptr, device=torch.device("cuda", local_rank))
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_resident(c, reps):
        """reps timed resident steps of `s` (L2 flushed before each, CUDA events on the launch stream)."""
        ms, info = 0.0, None
        for _ in range(reps):
            with torch.cuda.stream(ext):
                flush.zero_()                                    # L2 flush, outside the event pair
                a = torch.cuda.Event(enable_timing=True)
                b = torch.cuda.Event(enable_timing=True)
                a.record(ext)
                info = s.irls_resident(c, SIGMA, IRLS_ITERS, -1.0)
                b.record(ext)
            b.synchronize()
            ms += a.elapsed_time(b)
        return ms, info

    # ---- resident arm (value) -----------------------------------------------------------------
    s.upload(QQ_loc, I_loc, g.Q0, f)
    for _ in range(args.warmup):
        timed_resident(cost, 1)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    dev_ms, launches, infos = 0.0, 0, []
    for _ in range(args.steps):
        ms, info = timed_resident(cost, 1)
        dev_ms += ms
        launches += info.kernel_launches
        infos.append(info)
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1000.0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(dev_ms)
    world_units = gscale                                               # m = gscale x 1M edges
    value = world_units * IRLS_ITERS * args.steps / (dev_ms / 1000.0)   # units of the 1M-edge graph
    Q_res, w_res = s.download()

    # ---- the callers' default cost and Huber on the same graph (1 warm-up + 2 timed steps each) ----
    other = {}
    for nm in ("Geman-McClure", "Huber"):
        if nm == args.cost:
            continue
        s.irls_resident(COSTS[nm], SIGMA, IRLS_ITERS, -1.0)
        barrier()
        ms, oi = timed_resident(COSTS[nm], 2)
        ms = max_over_ranks(ms)
        other[nm] = {"value": world_units * IRLS_ITERS * 2 / (ms / 1000.0), "unit": UNIT, "ms_per_step": ms / 2,
                     "cg_iters_per_step": int(sum(oi.cg_iters))}

    # ---- N > 1: parity of the N-GPU result, strong scaling, configs[3] at full size ---------------------------------
    multi = {}
    if world > 1:
        from oracle import irls_oracle as O

        def identical_on_all_ranks(Q):
            t = torch.from_numpy(np.ascontiguousarray(Q)).cuda()
            t0 = t.clone()
            dist.broadcast(t0, 0)
            return bool(max_over_ranks(0.0 if torch.equal(t, t0) else 1.0) == 0.0)
        multi["bitwise_identical_across_ranks"] = identical_on_all_ranks(Q_res)
        # (1) the same graph, the same 30 iterations, on ONE GPU (rank 0 only; the other ranks wait at the barrier)
        if rank == 0:
            try:
                with ira.Solver(device=local_rank) as s1:
                    s1.upload(g.QQ, g.I, g.Q0, f)
                    t1 = s1.irls_resident(cost, SIGMA, IRLS_ITERS, -1.0).device_ms
                    i1 = s1.irls_resident(cost, SIGMA, IRLS_ITERS, -1.0)
                    Q1, _ = s1.download()
                one_ms = min(t1, i1.device_ms)
                multi["geodesic_rms_vs_one_gpu_same_graph_30iters_rad"] = float(O.geodesic_rms(Q_res, Q1, f))
                multi["one_gpu_same_graph"] = {"ms_per_step": one_ms, "cg_iters_per_step": int(sum(i1.cg_iters)),
                                               "speedup_of_n_gpus": one_ms / (dev_ms / args.steps)}
            except Exception as e:                           # noqa: BLE001
                multi["one_gpu_same_graph"] = {"error": repr(e)[:300]}
        barrier()
        # (2) 2 iterations against the CPU oracle (C restatement, Jacobi-PCG rtol 1e-12) on rank 0
        try:
            s.irls_resident(cost, SIGMA, 2, -1.0)
            Q2, _ = s.download()
            if rank == 0:
                from oracle import cport
                r = cport.irls(g.QQ, g.I, cost, SIGMA, g.Q0, f, 2, -1.0, cg_rtol=1e-12, threads=os.cpu_count() or 1)
                multi["geodesic_rms_vs_oracle_2iters_rad"] = float(O.geodesic_rms(Q2, r["Q"], f))
        except Exception as e:                               # noqa: BLE001
            multi["geodesic_rms_vs_oracle_2iters_rad"] = repr(e)[:300]
        barrier()
        # (3) strong scaling: the SAME 1M-edge graph on N GPUs (below peer_min_rows: every rank solves it alone)
        try:
            g1 = make_graph(1)
            if shard_mode == 1:
                s.upload(g1.QQ, g1.I, g1.Q0, g1.f)
            else:
                lo1, hi1 = edge_shard(g1.m, world, rank)
                s.upload(np.asfortranarray(g1.QQ[lo1:hi1]), np.ascontiguousarray(g1.I[lo1:hi1]), g1.Q0, g1.f)
            s.irls_resident(cost, SIGMA, IRLS_ITERS, -1.0)
            barrier()
            ms, si = timed_resident(cost, 2)
            ms = max_over_ranks(ms)
            multi["strong_scaling_1M_edges"] = {
                "value": IRLS_ITERS * 2 / (ms / 1000.0), "unit": "irls_iters/s (n=100000, m=1000000)", "ms_per_step": ms / 2,
                "cg_iters_per_step": int(sum(si.cg_iters)),
                "note": "n < ira_options.peer_min_rows (120000): not partitioned, every rank runs the single-GPU kernels"}
            del g1
        except Exception as e:                               # noqa: BLE001
            multi["strong_scaling_1M_edges"] = {"error": repr(e)[:300]}
        # (4) N = 8: configs[3] at its stated size, n = 1 000 000 / m = 10 000 000
        if world == 8 and gscale == 8 and shard_mode == 1 and not args.no_config3:
            try:
                g10 = make_graph(10)
                s.upload(g10.QQ, g10.I, g10.Q0, g10.f)
                s.irls_resident(cost, SIGMA, IRLS_ITERS, -1.0)
                barrier()
                ms, ci = timed_resident(cost, 2)
                ms = max_over_ranks(ms)
                Q10, _ = s.download()
                c3 = {"workload": "configs[3]: n=1000000 m=10000000, 30 IRLS iterations, L1", "ms_per_step": ms / 2,
                      "irls_iters_per_s": IRLS_ITERS * 2 / (ms / 1000.0),
                      "value_in_1M_edge_units": 10 * IRLS_ITERS * 2 / (ms / 1000.0),
                      "cg_iters_per_step": int(sum(ci.cg_iters)), "cg_hit_max": int(ci.cg_hit_max),
                      "final_score": ci.scores[-1] if ci.scores else None,
                      "bitwise_identical_across_ranks": identical_on_all_ranks(Q10)}
                s.irls_resident(cost, SIGMA, 2, -1.0)
                Q2, _ = s.download()
                if rank == 0:
                    from oracle import cport
                    r = cport.irls(g10.QQ, g10.I, cost, SIGMA, g10.Q0, g10.f, 2, -1.0, cg_rtol=1e-12,
                                   threads=os.cpu_count() or 1)
                    c3["geodesic_rms_vs_oracle_2iters_rad"] = float(O.geodesic_rms(Q2, r["Q"], g10.f))
                    c3["geodesic_rms_vs_ground_truth_rad"] = float(O.geodesic_rms(Q10, g10.Qgt, g10.f))
                multi["configs3_full_size"] = c3
                del g10
            except Exception as e:                           # noqa: BLE001
                multi["configs3_full_size"] = {"error": repr(e)[:300]}
            barrier()
        s.upload(QQ_loc, I_loc, g.Q0, f)

    # ---- e2e arm: host-buffer C-ABI call, pinned buffers --------------------------------------
    def pinned(a, order):
        t = torch.empty(a.size, dtype=torch.float64 if a.dtype == np.float64 else torch.int32).pin_memory()
        v = t.numpy().reshape(a.shape, order=order)
        v[...] = a
        return t, v
    keep = []
    tI, pI = pinned(I_loc, "C"); keep.append(tI)
    tQQ, pQQ = pinned(QQ_loc, "F"); keep.append(tQQ)
    tQ, pQ = pinned(np.asfortranarray(g.Q0), "F"); keep.append(tQ)
    tW, pW = pinned(np.zeros(m_loc), "C"); keep.append(tW)
    Q0f = np.asfortranarray(g.Q0)

    def e2e_step():
        pQ[...] = Q0f                                       # host-side reset of the in/out buffer (untimed)
        with torch.cuda.stream(ext):
            flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        info = s.irls_inplace(pQQ, pI, pQ, pW, cost, SIGMA, f, IRLS_ITERS, -1.0)
        return (time.perf_counter() - t0) * 1000.0, info

    e2e_ms = 0.0
    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        ms, info_e = e2e_step()
        e2e_ms += ms
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_value = world_units * IRLS_ITERS * e2e_steps / (e2e_ms / 1000.0)
    h2d = I_loc.nbytes + QQ_loc.nbytes + Q0f.nbytes
    d2h = Q0f.nbytes + 8 * m_loc
    same = bool(np.array_equal(np.ascontiguousarray(pQ), Q_res))

    # ---- roofline of the dominant kernel, measured live ----------------------------------------
    peak, peak_src = measured_peak()
    roof = None
    prof_share = None
    if world == 1 and gscale == 1:
        info = infos[-1]
        ph = info.profile.get("pcg_phases") or {}
        # CUDA events around EVERY kernel launch of one more step of the same call (profile mode)
        sp = ira.Solver(device=local_rank, profile=True)
        sp.upload(QQ_loc, I_loc, g.Q0, f)
        sp.irls_resident(cost, SIGMA, IRLS_ITERS, -1.0)
        pinfo = sp.irls_resident(cost, SIGMA, IRLS_ITERS, -1.0)
        pr = pinfo.profile
        tot = sum(v["ms"] for v in pr.values() if "launches" in v)
        prof_share = {k: round(v["ms"] / tot, 4) for k, v in pr.items() if "launches" in v}
        pcg = pr.get("pcg", {"ms": 0.0, "launches": 1})
        K = sum(pinfo.cg_iters) / max(1, pcg["launches"])                      # PCG iterations per launch (mean)
        per_iter_bytes = 16 * m + 224 * n                                       # SpMV + CG vector work
        launch_us = 1e3 * pcg["ms"] / max(1, pcg["launches"])
        ach = K * per_iter_bytes / (launch_us * 1e-6) / 1e9 if launch_us > 0 else 0.0
        res_us = sp.time_kernel(0, 100, False)
        res_cold = sp.time_kernel(0, 20, True)
        sp.close()
        spmv_bytes = 16 * m + 48 * n
        res_bytes = 72 * m + 56 * n
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as fh:
                traffic = json.load(fh)
        tr_pcg = (traffic or {}).get("k_pcg_smem" if pinfo.pcg_kernel == 4 else "k_pcg_persistent_reg") or {}
        spmv_phase_us = ph.get("spmv_us_per_phase")
        kname = {1: "k_pcg_persistent (vectors in HBM)", 2: "k_pcg_persistent_reg (state in registers)",
                 3: "k_pcg_persistent_reg_mw", 4: "k_pcg_smem (state in registers, matrix in shared memory)",
                 5: "k_pcg_peer*"}.get(pinfo.pcg_kernel, f"driver {pinfo.pcg_kernel}")
        roof = {
            "kernel": kname + ": persistent PCG solve (SELL SpMV + Chronopoulos-Gear PCG, 3 RHS); one launch per IRLS iteration",
            "pcg_kernel_id": pinfo.pcg_kernel,
            "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "peak_source": peak_src, "traffic": tr_pcg.get("dram_bytes_per_launch"), "traffic_source": tr_pcg.get("source"),
            "algorithmic_bytes_per_launch": K * per_iter_bytes, "algorithmic_bytes_per_pcg_iteration": per_iter_bytes,
            "pcg_iterations_per_launch": K, "launch_us": launch_us, "launches_timed": pcg["launches"],
            "share_of_step": prof_share.get("pcg"),
            "us_per_pcg_iteration": launch_us / max(K, 1e-9),
            "note": "launch_us = CUDA events on the launch stream around each of the 30 PCG-kernel launches of one step "
                    "(profile mode, same call as the timed steps).  traffic (ncu, DRAM bytes per launch) is ~2 % of the "
                    "algorithmic bytes: the matrix and vectors come from HBM once per launch and stay in registers / L2 for "
                    "its ~48 PCG iterations, so the HBM roofline is a bound the kernel cannot reach: the SpMV phase "
                    "sits on the L2 sector bandwidth of the 2m random 32 B gathers (tools/microbench.cu: >= 10.4 us), the rest "
                    "is two grid barriers per iteration.",
            "spmv_phase": {"us_per_phase_incl_reduction": spmv_phase_us, "algorithmic_bytes": spmv_bytes,
                           "frac_spmv_bytes_only": (spmv_bytes / (spmv_phase_us * 1e-6) / 1e9 / peak) if spmv_phase_us else None,
                           "frac_of_l2_gather_ceiling_10p4us": (10.4 / spmv_phase_us) if spmv_phase_us else None},
            "residual_kernel": {"launch_us": res_us, "algorithmic_bytes_per_launch": res_bytes,
                                "achieved": res_bytes / (res_us * 1e-6) / 1e9,
                                "frac": res_bytes / (res_us * 1e-6) / 1e9 / peak,
                                "cold_l2_us": res_cold, "frac_cold": res_bytes / (res_cold * 1e-6) / 1e9 / peak,
                                "traffic": (traffic or {}).get("k_residual")},
        }

    # ---- parity against the oracle on the run itself ----------------------------------------------
    cpu = None
    rms = rms30 = dev30 = None
    if rank == 0 and world == 1 and gscale == 1:
        from oracle import irls_oracle as O
        v, cores, r = cpu_port_run(g, cost, args.ref_iters)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": cpu_sample_text(args.ref_iters, cores),
               "cg_iters": list(r["cg_iters"])}
        v1, _, _ = cpu_port_run(g, cost, 3, threads=1)
        cpu["single_thread"] = {"value": v1, "unit": UNIT, "cores": 1, "sample": cpu_sample_text(3, 1)}
        ref = O.irls(g.QQ, g.I, None, cost, SIGMA, g.Q0, f, 2, -1.0, solver="pcg", pcg_rtol=1e-13)
        s2 = ira.Solver(device=local_rank)
        Q2, _, _ = s2.irls(g.QQ, g.I, None, cost, SIGMA, g.Q0, f, 2, -1.0)
        s2.close()
        rms = O.geodesic_rms(Q2, ref.Q, f)
        gpath = os.path.join(ROOT, "tests", "golden", "cfg3_l1_30iters.npz")
        if args.cost == "L1" and os.path.exists(gpath):       # all 30 iterations of the timed run itself
            gold = np.load(gpath)
            rms30 = O.geodesic_rms(Q_res, gold["Q"], f)
            dev30 = float(np.abs(np.array(infos[-1].scores) / gold["scores"] - 1).max())

    # ---- the other configs (bounded), N = 1 only ------------------------------------------------------------
    extras = {}
    if rank == 0 and world == 1 and gscale == 1 and not args.no_extras:
        extras["config2_kitti_scale"] = guarded(lambda: leg_config2(ira, local_rank))
        extras["config1_bundled_cli"] = guarded(lambda: leg_bundled(ira, local_rank))
        extras["init_mst"] = guarded(lambda: leg_aux(ira, local_rank, g))
        if not args.no_stream:
            def stream():
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import bench_stream
                return bench_stream.run(frames=args.stream_frames, loop_every=500, cpu_frames=40)
            extras["config5_rotavg_stream"] = guarded(stream)

    if rank == 0:
        info = infos[-1]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, gscale, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "upload_ms": info_e.upload_ms, "download_ms": info_e.download_ms,
                    "bitwise_equal_to_resident": same},
            "gpu_launches": int(launches), "clocks": clocks,
            "wall_ms_per_step": wall_ms / args.steps,
            "irls_iters_per_s_on_this_graph": IRLS_ITERS * args.steps / (dev_ms / 1000.0),
            "cg_iters_per_step": int(sum(info.cg_iters)), "cg_hit_max": int(info.cg_hit_max),
            "final_score": info.scores[-1] if info.scores else None,
            "geodesic_rms_vs_oracle_2iters_rad": rms if world == 1 else multi.pop("geodesic_rms_vs_oracle_2iters_rad", None),
            "geodesic_rms_vs_oracle_30iters_rad": rms30, "max_score_rel_dev_vs_oracle_30iters": dev30,
        }
        if roof is not None:
            line["roofline"] = roof
            line["kernel_time_share"] = prof_share
        if other:
            line["other_costs"] = other
        if gscale > 1:
            line["value_note"] = (f"weak scaling: the graph is the configs[2] recipe x{gscale} ({m} edges); value = IRLS "
                                  f"iterations/s x {gscale}, i.e. in units of the 1M-edge graph")
        if world > 1:
            ph = info.profile.get("pcg_phases") or {}
            line["pcg_us_per_iteration"] = 1e3 * ph.get("kernel_ms", 0.0) / max(1, int(sum(info.cg_iters)))
            line.update(multi)
        line.update(extras)
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    s.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cost", default="L1", choices=sorted(COSTS))
    ap.add_argument("--nccl-allreduce", action="store_true",
                    help="N > 1: edge shards + NCCL all-reduce per PCG iteration instead of the peer-memory solve")
    ap.add_argument("--graph-scale", type=int, default=0,
                    help="graph = configs[2] recipe x this (default: the number of GPUs); e.g. --gpus 1 --graph-scale 8 "
                         "runs the 8-GPU workload on one GPU for comparison")
    ap.add_argument("--no-stream", action="store_true", help="skip the configs[4] rotAvg stream")
    ap.add_argument("--stream-frames", type=int, default=10000, help="frames of the configs[4] stream (all 10 000 by default)")
    ap.add_argument("--no-extras", action="store_true", help="N = 1: skip the configs[0] / [1] / [4] and init_mst legs")
    ap.add_argument("--no-config3", action="store_true", help="N = 8: skip configs[3] at full size")
    ap.add_argument("--ref-iters", type=int, default=10, help="IRLS iterations in the CPU port's bounded sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
