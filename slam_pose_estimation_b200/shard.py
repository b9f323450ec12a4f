"""Filter-index sharding across the GPUs of one box.

Filters share nothing (each reference filter object owns its own mu, sigma, Q and timestamp,
UnscentedKalmanFilter.hpp:150-154), so the batch is split into contiguous index ranges, one per rank, with no
collective on the step path.  The only cross-rank traffic is the final gather of the estimates.
"""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """[first, last) of rank's contiguous shard; the first (total % world) ranks hold one extra filter."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("shard_range: bad arguments")
    base, extra = divmod(total, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def shard_sizes(total: int, world: int) -> list[int]:
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def gather_estimates(local, total: int, group=None, dst: int = 0):
    """Gather per-rank estimate tensors (shard_rows x cols, any device the process group supports) to `dst`;
    returns the (total x cols) tensor there and None elsewhere.  Uneven shards are padded to the largest."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank}: shard has {local.shape[0]} rows, expected {sizes[rank]}")
    width = max(sizes)
    padded = local
    if local.shape[0] < width:
        pad = torch.zeros((width - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded.contiguous(), bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)
