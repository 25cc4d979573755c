"""Set-sharding across the GPUs of one box (one process per GPU, torch.distributed).

The hot path shards with no data-path communication: every (parameter set, time) point is
independent and a set's log-likelihood only reduces over that set's own time axis
(SURVEY.md 8e).  Rank g owns a contiguous block of sets; `t`, `y`, `sigma` are replicated.
The only exchange is the all-gather of the per-set log-likelihoods (8 B per set): at that size a
collective library is pure latency, so the ranks store their blocks straight into each other's
exchange buffers over NVLink (`PeerGather`, csrc/jxp_peer.cu); NCCL is the fallback.
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


class _RawCuda:
    """A device pointer dressed as a CUDA array (zero-copy torch view of memory this package allocated itself)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = dict(shape=(n,), typestr="<f8", data=(int(ptr), False), version=3, strides=None)


class PeerGather:
    """All-gather of float64 per-set values by peer stores over NVLink (jxp_peer_allgather_f64).

    Every rank allocates one exchange buffer (CUDA IPC, mapped into every process of the box); `gather` is
    ONE launch on the caller's stream: this rank's block is stored into all buffers, an epoch flag follows,
    and the kernel returns when the blocks of all ranks have arrived.  The returned view (n_total,) stays valid
    until the gather after the next one (the buffers are double-buffered by epoch parity).  Collective: every rank
    constructs it and calls `gather` the same number of times."""

    def __init__(self, n_total: int, device: torch.device):
        import ctypes as C

        from jaxoplanet_b200 import _lib

        self._lib, self._C = _lib, C
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.n_total, self.device, self.epoch = int(n_total), device, 0
        lib = _lib.load()
        nbytes = int(lib.jxp_peer_buffer_bytes(self.n_total))
        own, handle = C.c_void_p(), (C.c_ubyte * 64)()
        self._opened, self.own = [], None
        ok = True
        try:
            _lib.call("jxp_peer_alloc", nbytes, C.byref(own), C.addressof(handle))
            self.own = int(own.value)
        except Exception:  # noqa: BLE001 -- every rank must reach the collectives below
            ok = False
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle) if ok else None)
        self._bufs = (C.c_uint64 * self.world)()
        if ok and all(h is not None for h in handles):
            for r, h in enumerate(handles):
                if r == self.rank:
                    self._bufs[r] = self.own
                    continue
                try:
                    ptr, hb = C.c_void_p(), (C.c_ubyte * 64).from_buffer_copy(h)
                    _lib.call("jxp_peer_open", C.addressof(hb), C.byref(ptr))
                    self._opened.append(int(ptr.value))
                    self._bufs[r] = int(ptr.value)
                except Exception:  # noqa: BLE001
                    ok = False
        else:
            ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self.ok = bool(int(flag.item()))
        if not self.ok:
            self.close()
            return
        self._views = [torch.as_tensor(_RawCuda(self.own + 256 + k * self.n_total * 8, self.n_total), device=device)
                       for k in range(2)]

    def gather(self, local: torch.Tensor, offset: int, out: torch.Tensor | None = None) -> torch.Tensor:
        """This rank's block -> every rank.  Returns `out` (contiguous float64, n_total values, written by the same
        launch) if given, else a view of the exchange buffer that stays valid until the gather after next."""
        from jaxoplanet_b200 import _array as A

        if local.dtype != torch.float64 or local.dim() != 1 or not local.is_contiguous():
            raise ValueError("PeerGather.gather: contiguous 1-D float64 block expected")
        # (epoch argument 0: the kernel counts the epochs in its buffer -- the launch can be captured into a CUDA
        # graph and replayed; `self.epoch` mirrors the count for the calls made through here, see sync_epoch)
        self.epoch += 1
        self._lib.call("jxp_peer_allgather_f64", local.data_ptr(), local.numel(), int(offset), self.n_total,
                       self._C.addressof(self._bufs), self.rank, self.world, 0, A.ptr(out), A.stream_ptr())
        return out if out is not None else self._views[self.epoch & 1]

    def _status2(self):
        from jaxoplanet_b200 import _array as A

        st = (self._C.c_uint64 * 2)()
        self._lib.call("jxp_peer_status", self.own, self._C.addressof(st), A.stream_ptr())
        return int(st[0]), int(st[1])

    def status(self) -> int:
        """0, or 0x100 | peer if a wait for that peer's block timed out (the values are then incomplete)."""
        return self._status2()[0]

    def sync_epoch(self) -> int:
        """After replays of a CUDA graph that holds a gather (launches this object did not see): take the epoch
        count from the device.  Returns it; `latest()` is then the view of the last gather."""
        self.epoch = self._status2()[1]
        return self.epoch

    def latest(self) -> torch.Tensor:
        return self._views[self.epoch & 1]

    def close(self):
        """Collective: nobody unmaps or frees while a peer may still be storing."""
        if dist.is_initialized():
            torch.cuda.synchronize()
            dist.barrier()
        for p in self._opened:
            self._lib.load().jxp_peer_close(p)
        self._opened = []
        if dist.is_initialized():
            dist.barrier()
        if self.own:
            self._lib.load().jxp_peer_free(self.own)
        self.own = None
        self.ok = False


_peer_gathers: dict = {}


def peer_gather_for(n_total: int, device: torch.device):
    """The PeerGather of (n_total, device), set up on first use by all ranks together; None when the backend is not
    NCCL or the buffers could not be shared (the caller then takes the NCCL all-gather)."""
    if not dist.is_initialized() or dist.get_world_size() == 1 or dist.get_backend() != "nccl" or device.type != "cuda":
        return None
    key = (int(n_total), device.index)
    if key not in _peer_gathers:
        pg = PeerGather(n_total, device)
        _peer_gathers[key] = pg if pg.ok else None
    return _peer_gathers[key]


def gather_sets(local: torch.Tensor, n_sets: int) -> torch.Tensor:
    """All-gather per-set values (leading axis = this rank's block) into the full (n_sets, ...)
    tensor on every rank: peer stores over NVLink for float64 vectors on CUDA (`PeerGather`), else the NCCL /
    gloo collective (ragged blocks are padded to the largest block for it)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if local.is_cuda and local.dtype == torch.float64 and local.dim() == 1:
        pg = peer_gather_for(n_sets, local.device)
        if pg is not None:
            lo, _ = shard_range(n_sets, dist.get_rank(), world)
            return pg.gather(local.contiguous(), lo, out=torch.empty(n_sets, dtype=torch.float64, device=local.device))
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
