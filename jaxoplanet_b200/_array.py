"""Host <-> device plumbing for the facade (PyTorch is used for device memory and
streams only; all arithmetic of the hot path happens in the CUDA library)."""

from __future__ import annotations

import contextlib
import contextvars

import numpy as np
import torch

_BATCH_NDIM = contextvars.ContextVar("jxp_batch_ndim", default=0)


@contextlib.contextmanager
def batched():
    """Inside this context, orbit parameters may carry ONE leading axis (the parameter-set
    axis): the equivalent of writing the reference model under ``jax.vmap``.  Light curves
    evaluated from such orbits get a leading ``(n_sets,)`` axis."""
    tok = _BATCH_NDIM.set(1)
    try:
        yield
    finally:
        _BATCH_NDIM.reset(tok)


def batch_ndim() -> int:
    return _BATCH_NDIM.get()


def device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "jaxoplanet_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback"
        )
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def is_torch(x) -> bool:
    return isinstance(x, torch.Tensor)


def result_dtype(*xs) -> torch.dtype:
    """float32 only if every floating ARRAY input is float32 (python scalars are weak),
    else float64 -- the analogue of jax_enable_x64 being on unless the caller opted out."""
    seen32 = False
    for x in xs:
        if x is None:
            continue
        if isinstance(x, torch.Tensor):
            if x.dtype == torch.float32:
                seen32 = True
            elif x.dtype.is_floating_point:
                return torch.float64
        elif isinstance(x, np.ndarray):
            if x.dtype == np.float32:
                seen32 = True
            elif x.dtype.kind == "f":
                return torch.float64
        elif isinstance(x, np.floating):
            if x.dtype == np.float32:
                seen32 = True
            else:
                return torch.float64
        elif isinstance(x, (list, tuple)):
            for el in x:
                if isinstance(el, (torch.Tensor, np.ndarray, np.floating)):
                    d = result_dtype(el)
                    if d == torch.float64:
                        return d
                    seen32 = True
    return torch.float32 if seen32 else torch.float64


def to_dev(x, dtype: torch.dtype) -> torch.Tensor:
    """Contiguous tensor of `dtype` on the current CUDA device (no copy when possible)."""
    dev = device()
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=dtype).contiguous()
    arr = np.asarray(x)
    if arr.dtype.kind not in "fiub":
        raise TypeError(f"unsupported input dtype {arr.dtype}")
    if arr.size >= 4096 and arr.dtype.kind == "f":
        return host_to_dev(arr, dtype)
    return torch.as_tensor(np.ascontiguousarray(arr), device="cpu").to(device=dev, dtype=dtype)


_PINNED: dict = {}


def host_to_dev(arr, dtype: torch.dtype = torch.float64) -> torch.Tensor:
    """NumPy array -> new CUDA tensor through a cached PINNED staging buffer: one memcpy into page-locked
    memory and one asynchronous DMA, instead of the driver's pageable-memory path (which stages the same
    bytes itself, synchronously, at a fraction of the bandwidth).  The staging buffer is reused, so the copy
    is ordered behind the previous one by an event."""
    dev = device()
    a = np.ascontiguousarray(arr)
    src = torch.from_numpy(a)
    if src.dtype != dtype:
        src = src.to(dtype)
    n = src.numel()
    if n < 4096:  # small: the plain path is as fast
        return src.to(dev)
    key = (dtype, dev.index)
    slot = _PINNED.get(key)
    if slot is None or slot[0].numel() < n:
        buf = torch.empty(max(n, 1 << 16), dtype=dtype).pin_memory()
        slot = [buf, torch.cuda.Event()]
        slot[1].record()
        _PINNED[key] = slot
    buf, ev = slot
    ev.synchronize()  # the previous DMA out of this buffer has finished
    buf[:n].copy_(src.reshape(-1))
    out = torch.empty(src.shape, dtype=dtype, device=dev)
    out.reshape(-1).copy_(buf[:n], non_blocking=True)
    ev.record()
    return out


def like_input(out: torch.Tensor, *inputs):
    """Return torch tensors to torch callers and numpy arrays to everyone else."""
    if any(isinstance(x, torch.Tensor) for x in inputs):
        return out
    return out.detach().cpu().numpy()


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()
