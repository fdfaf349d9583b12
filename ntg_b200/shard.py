"""Multi-GPU: the batch shards by problem index, one process per GPU.

Problems are independent (they share only read-only tables), so there is no
data-path collective: rank r of R owns the contiguous slice
[r*P/R, (r+1)*P/R).  The only exchange is the per-problem
result table (objective, max constraint violation) = 16 B/problem, gathered
with one all_gather over NCCL (NVLink/NVSwitch) -- or gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Tuple


def shard_range(P: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced, exhaustive: sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"rank {rank} / world {world}")
    base, rem = divmod(P, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(P: int, world: int) -> List[int]:
    return [hi - lo for lo, hi in (shard_range(P, r, world) for r in range(world))]


def gather_results(local, P: int, group=None):
    """all-gather the per-problem (objective, violation) rows of every rank into
    the full [P][2] table, in problem order.  `local` is this rank's
    [P_local][2] tensor (cuda for NCCL, cpu for gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(P, world)
    assert local.shape[0] == sizes[rank], "local shard does not match shard_range"
    if len(set(sizes)) == 1:
        out = torch.empty((P, local.shape[1]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # ragged: pad to the largest shard
    m = max(sizes)
    pad = torch.zeros((m, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    buf = torch.empty((world * m, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    return torch.cat([buf[r * m: r * m + sizes[r]] for r in range(world)], dim=0)
