"""N > 1 host logic on CPU: world_size-2 gloo run of the edge-sharded formulation.  Each rank builds
its shard's contribution to the right-hand side, the Jacobi diagonal and a Laplacian apply with the
oracle's make_A rule, all-reduces them, and both ranks must reproduce the unsharded operator."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_edge_shard_partition():
    from irotavg_b200.sharding import edge_shard
    for m in (0, 1, 7, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [edge_shard(m, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == m
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, os.environ["IRA_ROOT"])
    from oracle import graphs as G, irls_oracle as O
    from irotavg_b200.sharding import edge_shard
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    g = G.small_graph(n=120, extra=900, sigma_n=0.05, outlier_frac=0.1, seed=3, f=4, fixed_anywhere=True)
    rng = np.random.default_rng(5)
    wts = rng.uniform(0.1, 20.0, g.m); X = rng.standard_normal((g.n - g.f, 3))
    w3 = O.log_map(O.delta_rel(g.I, g.QQ, g.Q0))[:, :3]
    lo, hi = edge_shard(g.m, world, rank)
    A = O.make_A(g.n, g.f, g.I[lo:hi]).tocsr()
    w2 = wts[lo:hi] ** 2
    part = np.concatenate([(A.T @ (w2[:, None] * w3[lo:hi])).ravel(),
                           (A.T.multiply(A.T) @ w2).ravel(),
                           (A.T @ (w2[:, None] * (A @ X))).ravel()])
    t = torch.from_numpy(part.copy())
    dist.all_reduce(t)
    Af = O.make_A(g.n, g.f, g.I).tocsr(); w2f = wts ** 2
    full = np.concatenate([(Af.T @ (w2f[:, None] * w3)).ravel(), (Af.T.multiply(Af.T) @ w2f).ravel(),
                           (Af.T @ (w2f[:, None] * (Af @ X))).ravel()])
    err = np.abs(t.numpy() - full).max() / np.abs(full).max()
    # the unique-id broadcast path (payload only: NCCL itself is not initialised on CPU)
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0: buf.copy_(torch.arange(128, dtype=torch.uint8))
    dist.broadcast(buf, 0)
    ok = err < 1e-13 and bytes(buf.numpy().tobytes()) == bytes(range(128))
    print(f"rank {rank} err {err:.2e} ok {ok}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)
""")


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, IRA_ROOT=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok True") == 2
