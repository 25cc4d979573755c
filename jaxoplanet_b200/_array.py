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


_HAVE_CUDA = None
_DEVICES: dict = {}


def device() -> torch.device:
    """The current CUDA device (cached objects: this sits on the per-call path of the facade)."""
    global _HAVE_CUDA
    if _HAVE_CUDA is None:
        _HAVE_CUDA = bool(torch.cuda.is_available())
    if not _HAVE_CUDA:
        raise RuntimeError(
            "jaxoplanet_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback"
        )
    idx = torch._C._cuda_getDevice()
    dev = _DEVICES.get(idx)
    if dev is None:
        dev = _DEVICES[idx] = torch.device("cuda", idx)
    return dev


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream on the current device (the raw handle, without building a
    torch.cuda.Stream object per call)."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


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
    if arr.size >= 1024 and arr.dtype.kind == "f" and dtype == torch.float64:
        return host_to_dev(arr, dtype)
    return torch.as_tensor(np.ascontiguousarray(arr), device="cpu").to(device=dev, dtype=dtype)


# ---- pinned staging: host arrays reach the device through a small RING of page-locked arenas, one
# asynchronous DMA per call (the driver's pageable path stages the same bytes itself, synchronously, at a
# fraction of the bandwidth; a single reused buffer would make every upload wait for the previous one).
_RING_SLOTS = 4
_RINGS: dict = {}
_STREAMS: dict = {}


def _cur_stream():
    """torch.cuda.Stream object of the current stream (cached by raw handle: building one per call costs
    ~15 us, and Event.record() without an argument does just that)."""
    raw = stream_ptr()
    st = _STREAMS.get(raw)
    if st is None:
        st = _STREAMS[raw] = torch.cuda.current_stream()
    return st


class _Ring:
    def __init__(self, n_elems: int):
        self.cap = n_elems
        self.bufs = [torch.empty(n_elems, dtype=torch.float64).pin_memory() for _ in range(_RING_SLOTS)]
        self.views = [b.numpy() for b in self.bufs]
        self.events = [torch.cuda.Event() for _ in range(_RING_SLOTS)]
        self.used = [False] * _RING_SLOTS
        self.next = 0


def _ring(n_elems: int, dev_index: int) -> "_Ring":
    r = _RINGS.get(dev_index)
    if r is None or r.cap < n_elems:
        if r is not None:  # (outstanding DMAs out of the old arenas must finish before they are freed)
            torch.cuda.synchronize()
        r = _RINGS[dev_index] = _Ring(max(1 << 18, int(n_elems * 1.25)))
    return r


def stage_view(n_elems: int):
    """-> (float64 NumPy view of n_elems page-locked doubles, commit): fill the view in place, then
    ``commit()`` issues ONE asynchronous host-to-device copy and returns the float64 CUDA tensor."""
    dev = device()
    n = max(int(n_elems), 32)
    ring = _ring(n, dev.index)
    k = ring.next
    ring.next = (k + 1) % _RING_SLOTS
    if ring.used[k]:
        ring.events[k].synchronize()  # the DMA that last read this arena (four uploads ago) has finished

    def commit() -> torch.Tensor:
        darena = torch.empty(n, dtype=torch.float64, device=dev)
        darena.copy_(ring.bufs[k][:n], non_blocking=True)
        ring.events[k].record(_cur_stream())
        ring.used[k] = True
        return darena

    return ring.views[k][:n], commit


def stage_f64(arrays) -> list:
    """Several host float arrays -> float64 CUDA tensors of the same shapes, through ONE pinned arena and
    ONE host-to-device copy (each array starts on a 256-byte boundary of one device allocation)."""
    arrs = [np.asarray(a) for a in arrays]
    offs, n = [], 0
    for a in arrs:
        offs.append(n)
        n += (a.size + 31) & ~31
    view, commit = stage_view(n)
    for a, o in zip(arrs, offs):
        view[o:o + a.size] = a.reshape(-1)  # (converts any dtype to float64 on the way)
    darena = commit()
    return [darena[o:o + a.size].view(a.shape) for a, o in zip(arrs, offs)]


# ---- large host arrays that come back call after call (time stamps, observed flux): page-lock the caller's
# own memory once (cudaHostRegister) and DMA straight out of it -- no host-side copy at all, and always the
# array's CURRENT contents.  Only used where the call synchronises before it returns (NumPy in, NumPy out):
# the caller cannot touch the array while the copy is in flight.
_REGISTERED: dict = {}
_REGISTER_MIN_BYTES = 256 * 1024


def _unregister(ptr: int) -> None:
    if _REGISTERED.pop(ptr, None) is not None:
        try:
            torch.cuda.cudart().cudaHostUnregister(ptr)
        except Exception:  # interpreter shutdown
            pass


def registered_to_dev(arr: np.ndarray):
    """float64 C-contiguous NumPy array -> CUDA tensor by DMA out of the array's own (page-locked) memory;
    None when the array does not qualify (then the staging ring is used)."""
    if not (isinstance(arr, np.ndarray) and arr.dtype == np.float64 and arr.flags.c_contiguous
            and arr.nbytes >= _REGISTER_MIN_BYTES):
        return None
    ptr = arr.ctypes.data
    ent = _REGISTERED.get(ptr)
    if ent is None or ent < arr.nbytes:
        owner = arr
        while isinstance(owner.base, np.ndarray):
            owner = owner.base
        if owner.base is not None:  # memory owned by something else (mmap, bytes, another library): leave it alone
            return None
        if ent is not None:
            _unregister(ptr)
        try:
            rc = torch.cuda.cudart().cudaHostRegister(ptr, arr.nbytes, 0)
        except Exception:
            return None
        if int(rc) != 0:
            return None
        _REGISTERED[ptr] = arr.nbytes
        import weakref

        weakref.finalize(owner, _unregister, ptr)  # the pages are unpinned when the array is freed
    src = torch.from_numpy(arr)
    out = torch.empty(arr.shape, dtype=torch.float64, device=device())
    out.copy_(src, non_blocking=True)
    return out


def host_to_dev(arr, dtype: torch.dtype = torch.float64) -> torch.Tensor:
    """NumPy array -> new CUDA tensor through the pinned staging ring."""
    a = np.asarray(arr)
    if dtype == torch.float64 and a.size >= 1024:
        return stage_f64([a])[0]
    src = torch.from_numpy(np.ascontiguousarray(a))
    return src.to(device=device(), dtype=dtype)


def like_input(out: torch.Tensor, *inputs):
    """Return torch tensors to torch callers and numpy arrays to everyone else."""
    if any(isinstance(x, torch.Tensor) for x in inputs):
        return out
    return out.detach().cpu().numpy()


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()
