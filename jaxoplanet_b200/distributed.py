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


def reduce_time_shards(local: torch.Tensor) -> torch.Tensor:
    """Sum the per-set partial log-likelihoods of the time shards (all-reduce, 8 B per set).  The
    Gaussian log-likelihood is a sum over time samples, so the shards' values simply add."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    out = local.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


def _cuda_evaluate(params, u, t, y, sigma):
    """The fused CUDA path on this rank's block; the result STAYS on the device (the collective that follows
    runs over NCCL), whatever the inputs were."""
    from jaxoplanet_b200 import workloads
    from jaxoplanet_b200.light_curves.limb_dark import FusedSpec, _run_fused

    return _run_fused(FusedSpec(workloads.system_from(params), u), t, want="loglike", y=y, sigma=sigma)


def log_likelihood_distributed(params: dict, u, t, y, sigma, *, evaluate=None) -> torch.Tensor:
    """Per-set Gaussian log-likelihood of a batch of single- or multi-planet systems on every rank
    (SURVEY.md 8e).  `params`: dict of (n_sets,) or (n_sets, n_bodies) arrays of Body arguments, `u`:
    (n_sets, n_u) or (n_u,).

    * n_sets >= world size: the SET axis is sharded (rank g evaluates its contiguous block on all
      times), then one all-gather of 8 B per set -- no communication during compute.
    * n_sets < world size (e.g. one set with a very long time series): the TIME axis is sharded
      (rank g evaluates every set on its block of samples), then one all-reduce (sum).

    `evaluate(params, u, t, y, sigma)` -> (n_sets_local,) tensor; the default runs the fused CUDA path
    (the CPU tests pass the oracle to exercise this host logic under gloo)."""
    import numpy as np

    evaluate = evaluate or _cuda_evaluate
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n_sets = int(np.shape(params["period"])[0])
    u = np.asarray(u)
    sig = np.asarray(sigma)
    if n_sets >= world:
        lo, hi = shard_range(n_sets, rank, world)
        sub = {k: v[lo:hi] for k, v in params.items()}
        local = evaluate(sub, u[lo:hi] if u.ndim == 2 else u, t, y, sigma)
        return gather_sets(torch.as_tensor(local), n_sets)
    lo, hi = shard_range(len(t), rank, world)
    local = evaluate(params, u, t[lo:hi], y[lo:hi], sig[lo:hi] if sig.ndim == 1 and sig.shape[0] == len(t) else sigma)
    return reduce_time_shards(torch.as_tensor(local))
