"""Phase clocks of the peer-memory PCG kernel on N ranks (torchrun): tiny graph = fixed cost of the two cross-GPU
barriers per iteration, config 3 = with the u all-gather traffic.  python -m torch.distributed.run ... tools/peer_probe.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import irotavg_b200 as ira  # noqa: E402
from irotavg_b200.sharding import broadcast_unique_id  # noqa: E402
from oracle import graphs as G, irls_oracle as O  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
sigma = 5 * np.pi / 180
variant = int(os.environ.get("IRA_SPMV_VARIANT", "0"))
mode = int(os.environ.get("IRA_SHARD_MODE", "1"))      # 1 = barrier-free exchange, 2 = two cross-GPU barriers per iteration
s = ira.Solver(device=lr, world_size=world, rank=rank, shard_mode=mode, spmv_variant=variant)
s.comm_init(broadcast_unique_id(dist, ira.Solver, rank, device="cuda"))
cases = [("tiny n=3000", G.small_graph(n=3000, extra=30000, sigma_n=0.03, outlier_frac=0.1, seed=41)),
         ("config 3", G.random_graph())]
scale = int(os.environ.get("IRA_PROBE_SCALE", "0"))
if scale > 1:                                        # the weak-scaling workload of bench.py --gpus <scale>
    cases = [(f"config 3 x{scale}", G.random_graph(n=100_000 * scale, m=1_000_000 * scale))]
if os.environ.get("IRA_PROBE_ONLY"):
    cases = [c for c in cases if os.environ["IRA_PROBE_ONLY"] in c[0]]
for name, g in cases:
    s.upload(g.QQ, g.I, g.Q0, g.f)
    for cost, its in ((O.GEMAN_MCCLURE, 6), (O.L1, 12)):
        s.irls_resident(cost, sigma, its, -1.0)
        dist.barrier()
        info = s.irls_resident(cost, sigma, its, -1.0)
        ph = info.profile["pcg_phases"]
        n_it = sum(info.cg_iters)
        if rank == 0:
            print(f"[shard_mode {mode}, {world} ranks] {name} cost {cost}: cg {n_it} kernel {ph['kernel_ms']:.2f} ms -> {1e3 * ph['kernel_ms'] / max(n_it, 1):.1f} us/iter; "
                  f"phase A (SpMV+dots+barrier) {1e3 * ph['spmv_ms'] / max(ph['spmv_phases'], 1):.1f} us, "
                  f"phase B (update+all-gather+barrier) {1e3 * ph['update_ms'] / max(n_it, 1):.1f} us", flush=True)
s.close()
dist.destroy_process_group()
