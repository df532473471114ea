"""Batch sharding of independent audio streams over the GPUs of one box.

Every stream is independent through encode, conditioning, all N Euler steps (the three CFG rows of a stream stay
on one GPU) and decode, so the only multi-GPU step is ONE gather of the results at the end: no data-path
collective inside the loop.  Inputs are generated on the host from a global seed and sliced per rank, so the
result for stream b does not depend on the number of ranks.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_streams: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of ``n_streams`` for ``rank`` (first ``n % world`` ranks get one more)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    base, extra = divmod(n_streams, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(tensors: Sequence[torch.Tensor], world: int, rank: int):
    """Slice dim 0 of every tensor to this rank's shard."""
    n = tensors[0].shape[0]
    lo, hi = shard_bounds(n, world, rank)
    return [t[lo:hi].contiguous() for t in tensors]


def gather_streams(local: torch.Tensor, n_streams: int, group=None) -> torch.Tensor:
    """The single collective of the path: all-gather the per-rank results back into stream order.
    Ragged shards (n_streams % world != 0) are padded to the largest shard for the collective and trimmed."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = [shard_bounds(n_streams, world, r) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0], ) + tuple(local.shape[1:]))])
    out = local.new_empty((world * mx, ) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return torch.cat([out[r * mx:r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])
