"""`jax.numpy` stand-in: NumPy itself, plus the `.at[...].set()` update syntax (see jax/__init__.py)."""

import numpy as _np
from numpy import *  # noqa: F401,F403
from numpy import linalg  # noqa: F401

pi = _np.pi
inf = _np.inf
ndarray = _np.ndarray
bool_ = _np.bool_


class _AtIndexer:
    def __init__(self, arr, idx=None):
        self._arr, self._idx = arr, idx

    def __getitem__(self, idx):
        return _AtIndexer(self._arr, idx)

    def _apply(self, op, value):
        out = _np.array(self._arr, copy=True).view(ShimArray)
        out[self._idx] = value if op is None else op(out[self._idx], value)
        return out

    def set(self, value):
        return self._apply(None, value)

    def add(self, value):
        return self._apply(_np.add, value)

    def multiply(self, value):
        return self._apply(_np.multiply, value)


class ShimArray(_np.ndarray):
    """ndarray with jax's functional-update property."""

    @property
    def at(self):
        return _AtIndexer(self)


def _wrap(fn):
    def wrapped(*args, **kwargs):
        out = fn(*args, **kwargs)
        return out.view(ShimArray) if isinstance(out, _np.ndarray) and out.ndim > 0 else out

    wrapped.__name__ = fn.__name__
    return wrapped


# creators hand back ShimArray so that `.at` is available on what the reference builds
for _name in ("array", "asarray", "ones", "zeros", "full", "ones_like", "zeros_like", "full_like", "linspace",
              "arange", "stack", "concatenate", "hstack", "vstack", "atleast_1d", "atleast_2d", "broadcast_to"):
    globals()[_name] = _wrap(getattr(_np, _name))


def vectorize(pyfunc, *, excluded=frozenset(), signature=None):
    return _np.vectorize(pyfunc, excluded=excluded, signature=signature)


def interp(x, xp, fp, left=None, right=None, period=None):
    return _np.interp(x, xp, fp, left=left, right=right, period=period)


def dtype(x):
    # jnp.dtype(array) returns the array's dtype
    return x.dtype if hasattr(x, "dtype") else _np.dtype(x)
