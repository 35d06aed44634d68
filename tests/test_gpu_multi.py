"""Multi-GPU (one process per GPU, NCCL): edge-sharded irls() equals the single-GPU / oracle result.
Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); skipped on a
1-GPU box."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, os.environ["IRA_ROOT"])
    import irotavg_b200 as ira
    from irotavg_b200.sharding import edge_shard, broadcast_unique_id
    from oracle import graphs as G, irls_oracle as O
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    g = G.small_graph(n=3000, extra=30000, sigma_n=0.03, outlier_frac=0.1, seed=41, f=2, fixed_anywhere=True)
    sigma = 5 * np.pi / 180
    lo, hi = edge_shard(g.m, world, rank)
    s = ira.Solver(device=lr, world_size=world, rank=rank)
    s.comm_init(broadcast_unique_id(dist, ira.Solver, rank, device="cuda"))
    ok = True
    for cost, its in ((O.L1, 14), (O.GEMAN_MCCLURE, 6)):      # 14 L1 iterations: stiff pairs appear
        Q, w, info = s.irls(g.QQ[lo:hi], g.I[lo:hi], None, cost, sigma, g.Q0, g.f, its, -1.0)
        ref = O.irls(g.QQ, g.I, None, cost, sigma, g.Q0, g.f, its, -1.0, solver="direct")
        rms = O.geodesic_rms(Q, ref.Q, g.f)
        wok = np.allclose(w, ref.weights[lo:hi], rtol=1e-5, atol=1e-8)
        t = torch.from_numpy(np.ascontiguousarray(Q)).cuda()
        t0 = t.clone(); dist.broadcast(t0, 0)
        same = bool(torch.equal(t, t0))                  # replicas are bitwise identical across ranks
        print(f"rank {rank} cost {cost} cg {info.cg_iters} rms {rms:.2e} weights {wok} identical {same} comm {info.profile.get('comm')}", flush=True)
        ok = ok and rms <= 1e-8 and wok and same and info.cg_hit_max == 0
    s.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)
""")


PEER_WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, os.environ["IRA_ROOT"])
    import irotavg_b200 as ira
    from irotavg_b200.sharding import broadcast_unique_id
    from oracle import graphs as G, irls_oracle as O
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    sigma = 5 * np.pi / 180
    s = ira.Solver(device=lr, world_size=world, rank=rank, shard_mode=int(os.environ["IRA_SHARD_MODE"]),
                   peer_min_rows=int(os.environ.get("IRA_PEER_MIN_ROWS", "0")))   # 0: partition even these small graphs
    s.comm_init(broadcast_unique_id(dist, ira.Solver, rank, device="cuda"))
    ref = np.load(os.environ["IRA_REF"])                # oracle results computed once by the parent process
    ok = True
    for gi, kw in enumerate(eval(os.environ["IRA_GRAPHS"])):
        g = G.small_graph(**kw)
        for cost, its in ((O.L1, 14), (O.GEMAN_MCCLURE, 6)):      # 14 L1 iterations: stiff 2x2 / 3x3 blocks across ranks
            Q, w, info = s.irls(g.QQ, g.I, None, cost, sigma, g.Q0, g.f, its, -1.0)
            rms = O.geodesic_rms(Q, ref[f"Q_{gi}_{cost}"], g.f)
            wok = np.allclose(w, ref[f"w_{gi}_{cost}"], rtol=1e-5, atol=1e-8)
            t = torch.from_numpy(np.ascontiguousarray(Q)).cuda()
            t0 = t.clone(); dist.broadcast(t0, 0)
            same = bool(torch.equal(t, t0))                  # replicas are bitwise identical across ranks
            ph = info.profile.get("pcg_phases", {})
            print(f"rank {rank}/{world} graph {gi} cost {cost} cg {sum(info.cg_iters)} rms {rms:.2e} weights {wok} "
                  f"identical {same} pcg kernel {ph.get('kernel_ms', 0):.2f} ms", flush=True)
            ok = ok and rms <= 1e-8 and wok and same and info.cg_hit_max == 0
    s.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)
""")

PEER_GRAPHS = [dict(n=3000, extra=30000, sigma_n=0.03, outlier_frac=0.1, seed=41, f=2, fixed_anywhere=True),
               dict(n=20000, extra=150000, sigma_n=0.05, outlier_frac=0.1, seed=42)]


def _torchrun(tmp_path, text, nproc, timeout=900, extra_env=None):
    script = tmp_path / "worker.py"
    script.write_text(text)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, IRA_ROOT=ROOT, **(extra_env or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script)]
    return subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("shard_mode", [1, 2])
def test_peer_memory_solve(tmp_path, built_lib, shard_mode):
    """shard_mode 1 / 2: rows partitioned over the ranks, one persistent kernel per rank, exchange through NVLink peer
    memory (ira_peer.cuh; 1 = barrier-free with self-validating data, 2 = two cross-GPU barriers per iteration).
    Every rank holds the whole graph; results must match the oracle and each other bitwise."""
    import irotavg_b200 as ira
    import numpy as np
    from oracle import graphs as G, irls_oracle as O
    nd = ira.device_count()
    if nd < 2:
        pytest.skip("needs 2 GPUs")
    sigma = 5 * np.pi / 180
    ref = {}
    for gi, kw in enumerate(PEER_GRAPHS):
        g = G.small_graph(**kw)
        for cost, its in ((O.L1, 14), (O.GEMAN_MCCLURE, 6)):
            r = (O.irls(g.QQ, g.I, None, cost, sigma, g.Q0, g.f, its, -1.0, solver="direct") if gi == 0 else
                 O.irls(g.QQ, g.I, None, cost, sigma, g.Q0, g.f, its, -1.0, solver="pcg", pcg_rtol=1e-12))
            ref[f"Q_{gi}_{cost}"], ref[f"w_{gi}_{cost}"] = r.Q, r.weights
    np.savez(tmp_path / "ref.npz", **ref)
    res = _torchrun(tmp_path, PEER_WORKER, min(nd, 8) if nd in (2, 4, 8) else 2,
                    extra_env={"IRA_REF": str(tmp_path / "ref.npz"), "IRA_GRAPHS": repr(PEER_GRAPHS),
                               "IRA_SHARD_MODE": str(shard_mode)})
    print(res.stdout[-3000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


def test_small_graph_dispatch_replicated(tmp_path, built_lib):
    """Default peer_min_rows: graphs this small are not partitioned - every rank solves the whole problem with the
    single-GPU kernels; same oracle parity, bitwise identical on all ranks, no exchange."""
    import irotavg_b200 as ira
    import numpy as np
    from oracle import graphs as G, irls_oracle as O
    nd = ira.device_count()
    if nd < 2:
        pytest.skip("needs 2 GPUs")
    sigma = 5 * np.pi / 180
    ref = {}
    for gi, kw in enumerate(PEER_GRAPHS[:1]):
        g = G.small_graph(**kw)
        for cost, its in ((O.L1, 14), (O.GEMAN_MCCLURE, 6)):
            r = O.irls(g.QQ, g.I, None, cost, sigma, g.Q0, g.f, its, -1.0, solver="direct")
            ref[f"Q_{gi}_{cost}"], ref[f"w_{gi}_{cost}"] = r.Q, r.weights
    np.savez(tmp_path / "ref.npz", **ref)
    res = _torchrun(tmp_path, PEER_WORKER, 2, extra_env={"IRA_REF": str(tmp_path / "ref.npz"), "IRA_GRAPHS": repr(PEER_GRAPHS[:1]),
                                                        "IRA_SHARD_MODE": "1", "IRA_PEER_MIN_ROWS": "120000"})
    print(res.stdout[-3000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


def test_two_rank_sharded_irls(tmp_path, built_lib):
    import irotavg_b200 as ira
    if ira.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, IRA_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
