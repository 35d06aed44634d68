"""Edge sharding for the multi-GPU path (one process per GPU).

Edges are split into `world` contiguous slices; every rank owns its slice of (I, QQ, weights) and a
full replica of the node arrays.  Because A^T D^2 A = sum over shards of A_s^T D_s^2 A_s (each edge
contributes to exactly one shard, and make_A's mask is a per-edge rule), the per-node partial sums
(rhs, Jacobi diagonal, every SpMV output) only need one all-reduce each; nothing else crosses GPUs.
"""
from __future__ import annotations


def edge_shard(m: int, world: int, rank: int):
    """[lo, hi) of rank's contiguous edge slice; slices differ by at most one edge."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (m * rank) // world, (m * (rank + 1)) // world


def broadcast_unique_id(dist, solver_cls, rank: int, device=None) -> bytes:
    """Rank 0 creates the NCCL unique id through the C ABI; everyone receives it via torch.distributed
    (works with the gloo and nccl backends)."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8, device=device if device is not None else "cpu")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(solver_cls.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())
