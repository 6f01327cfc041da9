"""Data-parallel plumbing over the points: shard layout and the single all-reduce.

Every rank holds a contiguous shard of (X, Y) and replicas of all parameters; the per-shard sums produced by
`mgp_elbo_local` are combined with ONE all-reduce(sum) of the flat reduce buffer, after which every rank runs the
replicated `mgp_elbo_finish`.  The collective is torch.distributed's (NCCL over NVLink on the GPU box; gloo in
the CPU tests of this module's logic).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank`: the first n_total % world ranks get one extra row."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_count_and_offset(n_local: int, group=None, device: Optional[torch.device] = None) -> Tuple[int, int]:
    """(global number of points, global index of this rank's first point) from one small all-reduce."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[rank] = int(n_local)
    dist.all_reduce(counts, group=group)
    return int(counts.sum()), int(counts[:rank].sum())


def all_reduce_sum_(buf: torch.Tensor, group=None) -> torch.Tensor:
    """The step's single fused collective: in-place sum of the flat float64 reduce buffer across ranks."""
    import torch.distributed as dist
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf
