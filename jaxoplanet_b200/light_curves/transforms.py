"""``integrate``: exposure-time integration, drop-in for
src/jaxoplanet/light_curves/transforms.py:21-116.

When ``func`` is a fused Keplerian light curve the sub-samples are folded into the same
kernel tile (no extra launches, no intermediate arrays); any other callable is evaluated
at ``t + dt_j`` and combined with the stencil, as the reference does.
"""

from __future__ import annotations

from functools import wraps

import numpy as np
import torch

from jaxoplanet_b200 import _array as A
from jaxoplanet_b200 import _lib

__all__ = ["integrate", "interpolate"]


def _stencil(exposure_time, order, num_samples):
    # transforms.py:67-89
    num_samples = int(num_samples)
    num_samples += 1 - num_samples % 2
    stencil = np.ones(num_samples)
    if order == 0:
        dt = np.linspace(-0.5, 0.5, 2 * num_samples + 1)[1:-1:2]
    elif order == 1:
        dt = np.linspace(-0.5, 0.5, num_samples)
        stencil = 2 * stencil
        stencil[0] = 1
        stencil[-1] = 1
    elif order == 2:
        dt = np.linspace(-0.5, 0.5, num_samples)
        stencil[1:-1:2] = 4
        stencil[2:-1:2] = 2
    else:
        raise ValueError("The parameter 'order' in 'integrate_exposure_time' must be 0, 1, or 2")
    return dt * exposure_time, stencil / np.sum(stencil)


def integrate(func, exposure_time=None, order: int = 0, num_samples: int = 7):
    """Transform a light curve function to apply exposure time integration.

    Args:
        func: A light curve function which takes a time as the first argument
        exposure_time: The exposure time (in days)
        order: 0 (resampling, Kipping 2010), 1 (trapezoid) or 2 (Simpson)
        num_samples: function evaluations per integral (made odd)
    """
    if exposure_time is None:
        return func

    if np.ndim(exposure_time) != 0:
        raise ValueError(
            "The exposure time passed to 'integrate_exposure_time' has shape "
            f"{np.shape(exposure_time)}, but a scalar was expected; "
            "To use exposure time integration with different exposures at different "
            "times, manually 'vmap' or 'vectorize' the function"
        )
    if order not in (0, 1, 2):
        raise ValueError("The parameter 'order' in 'integrate_exposure_time' must be 0, 1, or 2")

    spec = getattr(func, "_jxp_spec", None)
    if spec is not None and spec.exposure_time is None:
        # a fused light curve (or log-likelihood): the stencil goes into the same kernel launch; the wrapped
        # function keeps the signature of `func` ((t) or (t, y, sigma))
        from jaxoplanet_b200.light_curves.limb_dark import with_spec

        return with_spec(func, exposure_time=float(exposure_time), exp_order=int(order),
                         num_samples=int(num_samples))

    dt, stencil = _stencil(float(exposure_time), order, num_samples)

    @wraps(func)
    def wrapped(time, *args, **kwargs):
        res = None
        for d, w in zip(dt, stencil):
            term = func(time + float(d), *args, **kwargs)
            term = term * float(w)
            res = term if res is None else res + term
        return res

    return wrapped


def interpolate(func, *, period, time_transit, num_samples: int, duration=None, args=(), kwargs=None):
    """Pre-compute ``func`` on a grid around the transit and linearly interpolate it at the
    (period-wrapped) requested times: drop-in for src/jaxoplanet/light_curves/transforms.py:119-185
    (``jnp.interp(..., period=period)`` semantics).  The grid evaluation runs in the fused CUDA path;
    the table jnp.interp builds (grid reduced mod period, sorted, one wrapped point added on each side)
    is prepared once here, and every call evaluates it on the device (kernel K7,
    ``jxp_interp_periodic_*``)."""
    kwargs = kwargs or {}
    period = float(period)
    time_transit = float(time_transit)
    if duration is None:
        duration = period
    time_grid = time_transit + float(duration) * np.linspace(-0.5, 0.5, int(num_samples))
    flux_grid = func(time_grid, *args, **kwargs)
    # the table below takes axis 0 of func's output as the grid axis (what the reference's func returns for
    # a 1-D time array).  A fused light curve of a BATCHED orbit returns (n_sets, num_samples, ...): one
    # period / time_transit cannot serve a batch of ephemerides, and indexing the set axis with grid
    # indices would silently build a wrong table.
    spec_ = getattr(func, "_jxp_spec", None)
    if spec_ is not None:
        from jaxoplanet_b200.light_curves.limb_dark import pack_bodies

        if pack_bodies(spec_.orbit)[2]:
            raise ValueError("interpolate() takes the light curve of ONE parameter set (scalar period and "
                             "time_transit); evaluate batched orbits directly, or interpolate set by set")
    if tuple(np.shape(flux_grid))[:1] != (int(num_samples),):
        raise ValueError(f"func(time_grid) must have the grid axis first: expected leading size {int(num_samples)}, "
                         f"got shape {tuple(np.shape(flux_grid))}")
    grid_dtype = flux_grid.dtype if isinstance(flux_grid, torch.Tensor) else A.result_dtype(np.asarray(flux_grid))
    flux_np = flux_grid.detach().cpu().numpy() if isinstance(flux_grid, torch.Tensor) else np.asarray(flux_grid)
    cols_shape = flux_np.shape[1:]
    flat = flux_np.reshape(flux_np.shape[0], -1)
    xp = np.mod(time_grid, period)
    order = np.argsort(xp, kind="stable")
    xp = xp[order]
    fp = flat[order]
    xp = np.concatenate([xp[-1:] - period, xp, xp[:1] + period])
    fp = np.concatenate([fp[-1:], fp, fp[:1]])
    xp_d = A.to_dev(xp, torch.float64)
    fp_d = A.to_dev(fp, grid_dtype)

    @wraps(func)
    def wrapped(time, *a, **k):
        del a, k
        td = A.to_dev(time, torch.float64)
        out = torch.empty(tuple(td.shape) + (fp_d.shape[1],), dtype=grid_dtype, device=td.device)
        name = "jxp_interp_periodic_f64" if grid_dtype == torch.float64 else "jxp_interp_periodic_f32"
        _lib.call(name, xp_d.data_ptr(), fp_d.data_ptr(), int(xp_d.shape[0]), int(fp_d.shape[1]), period,
                  time_transit, td.data_ptr(), td.numel(), out.data_ptr(), A.stream_ptr())
        out = out.reshape(tuple(td.shape) + tuple(cols_shape))
        return A.like_input(out, time)

    return wrapped
