import numpy as _np


class Zero:
    """Symbolic zero tangent (never produced by the shim; kept so `type(tan) is ad.Zero` works)."""

    def __init__(self, aval=None):
        self.aval = aval


def zeros_like_aval(aval):
    return _np.zeros(getattr(aval, "shape", ()), dtype=getattr(aval, "dtype", _np.float64))
