"""Guide sharding across the GPUs of one box.

Per-guide results depend only on that guide and the (replicated) index, so rank r of n takes the contiguous slice
[r*G/n, (r+1)*G/n) of the guides in aggregator order; hit lists stay with the owning rank.  The only collective of the
path is one all-gather of the per-guide total counts (int32) so that every rank holds the global vector.
torch.distributed is plumbing here: NCCL between GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n_guides: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced (sizes differ by at most one) slice of rank `rank`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world: %d/%d" % (rank, world))
    from . import api  # the C ABI's ff_shard_range: one definition for the one-process-per-GPU ranks and for ff_multi
    first, count = api.shard_range(n_guides, world, rank)
    return first, first + count


def all_gather_counts(local_counts, n_guides: int, group=None):
    """All-gather the per-guide totals of every rank's shard into the global [n_guides] vector (int32 tensor on the
    same device as `local_counts`).  Shards may differ by one element, so slices are padded to the largest shard."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(n_guides, rank, world)
    assert local_counts.numel() == hi - lo, "local_counts must hold exactly this rank's shard"
    width = -(-n_guides // world) if n_guides else 0
    padded = torch.zeros(max(width, 1), dtype=torch.int32, device=local_counts.device)
    padded[:hi - lo] = local_counts.to(torch.int32)
    out = torch.empty(world * max(width, 1), dtype=torch.int32, device=local_counts.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = []
    for r in range(world):
        a, b = shard_range(n_guides, r, world)
        parts.append(out[r * max(width, 1): r * max(width, 1) + (b - a)])
    return torch.cat(parts) if parts else out[:0]
