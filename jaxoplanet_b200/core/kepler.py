"""``kepler(M, ecc) -> (sin f, cos f)``: drop-in for src/jaxoplanet/core/kepler.py:14-79,
computed by the K1 CUDA kernel (jxp_kepler_f64/f32).  Differentiable through
``torch.autograd`` with the reference's closed-form rule (kepler.py:59-79), whose two
factors the kernel returns."""

from __future__ import annotations

import torch

from jaxoplanet_b200 import _array as A
from jaxoplanet_b200 import _lib

__all__ = ["kepler", "kepler_with_E"]


def _launch(M, e, want_E=False, want_jvp=False):
    """M, e: CUDA tensors of equal dtype, already broadcast-compatible (numel n or 1)."""
    dt = M.dtype
    n = max(M.numel(), e.numel())
    shape = torch.broadcast_shapes(M.shape, e.shape)
    if M.numel() not in (1, n):
        M = M.expand(shape).contiguous()
    if e.numel() not in (1, n):
        e = e.expand(shape).contiguous()
    ms = 0 if (M.numel() == 1 and n > 1) else 1
    es = 0 if (e.numel() == 1 and n > 1) else 1
    sinf = torch.empty(shape, dtype=dt, device=M.device)
    cosf = torch.empty_like(sinf)
    E = torch.empty_like(sinf) if want_E else None
    dfdM = torch.empty_like(sinf) if want_jvp else None
    dfde = torch.empty_like(sinf) if want_jvp else None
    name = "jxp_kepler_f64" if dt == torch.float64 else "jxp_kepler_f32"
    _lib.call(name, A.ptr(M), ms, A.ptr(e), es, n, A.ptr(sinf), A.ptr(cosf), A.ptr(E),
              A.ptr(dfdM), A.ptr(dfde), A.stream_ptr())
    return sinf, cosf, E, dfdM, dfde


class _KeplerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, M, e):
        sinf, cosf, _, dfdM, dfde = _launch(M.detach(), e.detach(), want_jvp=True)
        ctx.save_for_backward(sinf, cosf, dfdM, dfde)
        ctx.shapes = (M.shape, e.shape)
        return sinf, cosf

    @staticmethod
    def backward(ctx, gs, gc):
        sinf, cosf, dfdM, dfde = ctx.saved_tensors
        gf = gs * cosf - gc * sinf  # d sin f = cos f df, d cos f = -sin f df
        gM = (gf * dfdM).sum_to_size(ctx.shapes[0]) if ctx.needs_input_grad[0] else None
        ge = (gf * dfde).sum_to_size(ctx.shapes[1]) if ctx.needs_input_grad[1] else None
        return gM, ge


def kepler(M, ecc):
    """Solve Kepler's equation to compute the true anomaly.

    Args:
        M: Mean anomaly
        ecc: Eccentricity

    Returns:
        The sine and cosine of the true anomaly
    """
    dt = A.result_dtype(M, ecc)
    Mt = A.to_dev(M, dt)
    et = A.to_dev(ecc, dt)
    if Mt.requires_grad or et.requires_grad:
        sinf, cosf = _KeplerFn.apply(Mt, et)
    else:
        sinf, cosf, *_ = _launch(Mt, et)
    return A.like_input(sinf, M, ecc), A.like_input(cosf, M, ecc)


def kepler_with_E(M, ecc):
    """(sin f, cos f, E): also returns the eccentric anomaly the reference keeps internal
    (kepler.py:39-43); used by the parity tests for the 1e-13 rad bar."""
    dt = A.result_dtype(M, ecc)
    sinf, cosf, E, _, _ = _launch(A.to_dev(M, dt), A.to_dev(ecc, dt), want_E=True)
    return A.like_input(sinf, M, ecc), A.like_input(cosf, M, ecc), A.like_input(E, M, ecc)
