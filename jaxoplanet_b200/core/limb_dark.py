"""``light_curve(u, b, r, order=10) -> flux - 1``: drop-in for
src/jaxoplanet/core/limb_dark.py:19-47, computed by the K2 CUDA kernel
(jxp_limb_dark_f64/f32).  ``torch.autograd`` gradients w.r.t. (u, b, r) use the kernel's
forward-mode partials of the discretised formula (what jax.grad of the reference yields)."""

from __future__ import annotations

import math

import torch

from jaxoplanet_b200 import _array as A
from jaxoplanet_b200 import _lib

__all__ = ["light_curve"]


def greens_normalised(u: torch.Tensor) -> torch.Tensor:
    """g(u) / (pi (g0 + g1 / 1.5)); limb_dark.py:42-45, 87-99 (host side, tiny, differentiable)."""
    n = u.shape[0]
    if n == 0:
        return torch.full((1,), 1.0 / math.pi, dtype=u.dtype, device=u.device)
    ue = torch.cat((-torch.ones(1, dtype=u.dtype, device=u.device), u))
    size = n + 1
    B = torch.tensor([[math.comb(i, j) for i in range(size)] for j in range(size)], dtype=u.dtype,
                     device=u.device)
    sign = torch.tensor([(-1.0) ** (j + 1) for j in range(size)], dtype=u.dtype, device=u.device)
    p = sign * (B @ ue)
    g = [torch.zeros((), dtype=u.dtype, device=u.device) for _ in range(size + 2)]
    for k in range(size - 1, 1, -1):
        g[k] = p[k] / (k + 2) + g[k + 2]
    g[1] = p[1] + 3 * g[3]
    g[0] = p[0] + 2 * g[2]
    g = torch.stack(g[:-2])
    return g / (math.pi * (g[0] + g[1] / 1.5))


def _launch(u, b, r, order, want_grad):
    dt = b.dtype
    shape = torch.broadcast_shapes(b.shape, r.shape)
    n = 1
    for s in shape:
        n *= s
    if b.numel() not in (1, n):
        b = b.expand(shape).contiguous()
    if r.numel() not in (1, n):
        r = r.expand(shape).contiguous()
    bs = 0 if (b.numel() == 1 and n > 1) else 1
    rs = 0 if (r.numel() == 1 and n > 1) else 1
    n_u = u.shape[0]
    flux = torch.empty(shape, dtype=dt, device=b.device)
    db = torch.empty_like(flux) if want_grad else None
    dr = torch.empty_like(flux) if want_grad else None
    svec = torch.empty((n_u + 1,) + tuple(shape), dtype=dt, device=b.device) if want_grad else None
    name = "jxp_limb_dark_f64" if dt == torch.float64 else "jxp_limb_dark_f32"
    _lib.call(name, A.ptr(u), n_u, A.ptr(b), bs, A.ptr(r), rs, n, int(order), A.ptr(flux),
              A.ptr(db), A.ptr(dr), A.ptr(svec), A.stream_ptr())
    return flux, db, dr, svec


class _LimbDarkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, b, r, order):
        flux, db, dr, svec = _launch(u.detach(), b.detach(), r.detach(), order, True)
        ctx.save_for_backward(u.detach(), db, dr, svec)
        ctx.shapes = (b.shape, r.shape)
        return flux

    @staticmethod
    def backward(ctx, g):
        u, db, dr, svec = ctx.saved_tensors
        gu = gb = gr = None
        if ctx.needs_input_grad[0]:
            # flux = s . g(u) - 1  =>  d flux / d u = (sum_points g_out s_k) . dg_k/du
            w = (svec * g.unsqueeze(0)).reshape(svec.shape[0], -1).sum(1)
            with torch.enable_grad():
                uu = u.clone().requires_grad_(True)
                (gu,) = torch.autograd.grad((greens_normalised(uu) * w).sum(), uu)
        if ctx.needs_input_grad[1]:
            gb = (g * db).sum_to_size(ctx.shapes[0])
        if ctx.needs_input_grad[2]:
            gr = (g * dr).sum_to_size(ctx.shapes[1])
        return gu, gb, gr, None


def light_curve(u, b, r, *, order: int = 10):
    """Compute the light curve for polynomial limb darkening (flux - 1).

    Args:
        u: The coefficients of the polynomial limb darkening model (1-D, len <= 8)
        b: The center-to-center distance between the occultor and the occulted body
        r: The radius ratio between the occultor and the occulted body
        order: The quadrature order of the 1D line integral
    """
    dt = A.result_dtype(u, b, r)
    ut = A.to_dev(u, dt)
    assert ut.ndim == 1  # limb_dark.py:40
    bt = A.to_dev(b, dt)
    rt = A.to_dev(r, dt)
    if ut.requires_grad or bt.requires_grad or rt.requires_grad:
        flux = _LimbDarkFn.apply(ut, bt, rt, order)
    else:
        flux, *_ = _launch(ut, bt, rt, order, False)
    return A.like_input(flux, u, b, r)
