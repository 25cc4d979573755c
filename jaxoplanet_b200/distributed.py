"""Set-sharding across the GPUs of one box (one process per GPU, torch.distributed).

The hot path shards with no data-path communication: every (parameter set, time) point is
independent and a set's log-likelihood only reduces over that set's own time axis
(SURVEY.md 8e).  Rank g owns a contiguous block of sets; `t`, `y`, `sigma` are replicated.
The only exchange is the all-gather of the per-set log-likelihoods (8 B per set).
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_sets: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; blocks differ by at most one set."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_sets, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_sets(local: torch.Tensor, n_sets: int) -> torch.Tensor:
    """All-gather per-set values (leading axis = this rank's block) into the full (n_sets, ...)
    tensor on every rank.  Ragged blocks are padded to the largest block for the collective."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(n_sets, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    parts = [out[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, 0)
