"""NumPy restatement of the jaxoplanet hot path.  TEST INFRASTRUCTURE ONLY.

Each function follows the cited lines of ``/root/reference`` operation by
operation (same association order, same guards).  Pinned (tests/test_reference_vectors.py)
against outputs of the reference's own source files executed on a NumPy-backed jax shim
(tests/golden/reference_v1.npz) and against the reference's known-answer tests
(tests/test_oracle_kats.py); only XLA's libm/fusion vs NumPy's remains unpinned.

Every numerical function takes an optional backend ``xp`` (``NP`` default, or
``TH`` = torch) so that the very same formulas are differentiated by torch-f64
autograd for the derivative oracle; the two reference custom-JVP rules
(``zero_safe_sqrt`` src/jaxoplanet/utils.py:16-24 and ``_kepler``
src/jaxoplanet/core/kepler.py:59-79) are mirrored as ``torch.autograd.Function``.
"""

from __future__ import annotations

import math

import numpy as np
from scipy.special import binom, roots_legendre

G_GRAV = 2942.2062175044193  # src/jaxoplanet/constants.py:2


# --------------------------------------------------------------------------
# backends
# --------------------------------------------------------------------------
class NP:
    name = "numpy"
    pi = np.pi
    where = staticmethod(np.where)
    sqrt = staticmethod(np.sqrt)
    sin = staticmethod(np.sin)
    cos = staticmethod(np.cos)
    tan = staticmethod(np.tan)
    arctan2 = staticmethod(np.arctan2)
    abs = staticmethod(np.abs)
    maximum = staticmethod(np.maximum)
    minimum = staticmethod(np.minimum)
    square = staticmethod(np.square)
    cbrt = staticmethod(np.cbrt)
    ones_like = staticmethod(np.ones_like)
    zeros_like = staticmethod(np.zeros_like)
    stack = staticmethod(np.stack)
    zero_safe_sqrt = staticmethod(np.sqrt)  # utils.py:11-13 (primal)

    @staticmethod
    def asarray(x, dtype=None):
        return np.asarray(x, dtype=dtype)

    @staticmethod
    def eps(x):
        return np.finfo(np.asarray(x).dtype).eps

    @staticmethod
    def remainder(x, y):
        return np.remainder(x, y)

    @staticmethod
    def kepler(M, ecc):
        return kepler(M, ecc)


def _make_torch_backend():
    import torch

    class _ZeroSafeSqrt(torch.autograd.Function):
        # src/jaxoplanet/utils.py:16-24
        @staticmethod
        def forward(ctx, x):
            y = torch.sqrt(x)
            ctx.save_for_backward(x, y)
            return y

        @staticmethod
        def backward(ctx, g):
            x, y = ctx.saved_tensors
            cond = x < 10 * torch.finfo(x.dtype).eps
            denom = torch.where(cond, torch.ones_like(x), x)
            return 0.5 * g * y / denom

    class _Kepler(torch.autograd.Function):
        # primal via the NumPy restatement, tangent rule kepler.py:59-79
        @staticmethod
        def forward(ctx, M, e):
            s, c = kepler(M.detach().numpy(), e.detach().numpy())
            s = torch.from_numpy(np.asarray(s))
            c = torch.from_numpy(np.asarray(c))
            ctx.save_for_backward(e, s, c)
            return s, c

        @staticmethod
        def backward(ctx, gs, gc):
            e, sinf, cosf = ctx.saved_tensors
            ecosf = e * cosf
            ome2 = 1 - e**2
            dfdM = (1 + ecosf) ** 2 / ome2**1.5
            dfde = (2 + ecosf) * sinf / ome2
            gf = gs * cosf - gc * sinf
            return gf * dfdM, gf * dfde

    class TH:
        name = "torch"
        pi = math.pi
        where = staticmethod(torch.where)
        sqrt = staticmethod(torch.sqrt)
        sin = staticmethod(torch.sin)
        cos = staticmethod(torch.cos)
        tan = staticmethod(torch.tan)
        arctan2 = staticmethod(torch.arctan2)
        abs = staticmethod(torch.abs)
        square = staticmethod(torch.square)
        ones_like = staticmethod(torch.ones_like)
        zeros_like = staticmethod(torch.zeros_like)
        zero_safe_sqrt = staticmethod(_ZeroSafeSqrt.apply)

        @staticmethod
        def stack(xs, axis=0):
            return torch.stack(list(xs), dim=axis)

        @staticmethod
        def _t(x, like):
            if torch.is_tensor(x):
                return x
            return torch.as_tensor(x, dtype=like.dtype)

        @staticmethod
        def maximum(a, b):
            if not torch.is_tensor(a):
                a = torch.as_tensor(a, dtype=b.dtype)
            if not torch.is_tensor(b):
                b = torch.as_tensor(b, dtype=a.dtype)
            return torch.maximum(a, b)

        @staticmethod
        def minimum(a, b):
            if not torch.is_tensor(a):
                a = torch.as_tensor(a, dtype=b.dtype)
            if not torch.is_tensor(b):
                b = torch.as_tensor(b, dtype=a.dtype)
            return torch.minimum(a, b)

        @staticmethod
        def cbrt(x):
            return torch.sign(x) * torch.abs(x) ** (1.0 / 3.0)

        @staticmethod
        def asarray(x, dtype=None):
            return torch.as_tensor(x, dtype=dtype)

        @staticmethod
        def eps(x):
            return torch.finfo(x.dtype).eps

        @staticmethod
        def remainder(x, y):
            return torch.remainder(x, y)

        @staticmethod
        def kepler(M, ecc):
            M, ecc = torch.broadcast_tensors(M, ecc)
            return _Kepler.apply(M.contiguous(), ecc.contiguous())

    return TH


_TH = None


def torch_backend():
    global _TH
    if _TH is None:
        _TH = _make_torch_backend()
    return _TH


# --------------------------------------------------------------------------
# Kepler solver  (src/jaxoplanet/core/kepler.py)
# --------------------------------------------------------------------------
def starter(M, ecc, ome):
    """Markley starter, kepler.py:82-92."""
    M2 = np.square(M)
    alpha = 3 * np.pi / (np.pi - 6 / np.pi)
    alpha = alpha + 1.6 / (np.pi - 6 / np.pi) * (np.pi - M) / (1 + ecc)
    d = 3 * ome + alpha * ecc
    alphad = alpha * d
    r = (3 * alphad * (d - ome) + M2) * M
    q = 2 * alphad * ome - M2
    q2 = np.square(q)
    w = np.square(np.cbrt(np.abs(r) + np.sqrt(q2 * q + r * r)))
    return (2 * r * w / (np.square(w) + w * q + q2) + M) / d


def refine(M, ecc, ome, E):
    """Single high-order correction, kepler.py:95-108."""
    sE = E - np.sin(E)
    cE = 1 - np.cos(E)

    f_0 = ecc * sE + E * ome - M
    f_1 = ecc * cE + ome
    f_2 = ecc * (E - sE)
    f_3 = 1 - f_1
    d_3 = -f_0 / (f_1 - 0.5 * f_0 * f_2 / f_1)
    d_4 = -f_0 / (f_1 + 0.5 * d_3 * f_2 + (d_3 * d_3) * f_3 / 6)
    d_42 = d_4 * d_4
    dE = -f_0 / (f_1 + 0.5 * d_4 * f_2 + d_4 * d_4 * f_3 / 6 - d_42 * d_4 * f_2 / 24)
    return E + dE


def kepler(M, ecc, return_E=False):
    """(M, e) -> (sin f, cos f); kepler.py:28-56.  dtype follows the inputs."""
    M = np.asarray(M)
    ecc = np.asarray(ecc)
    dt = np.result_type(M.dtype, ecc.dtype)
    if dt.kind != "f":
        dt = np.dtype(np.float64)
    M = M.astype(dt)
    ecc = ecc.astype(dt)
    two_pi = dt.type(2 * np.pi)
    pi = dt.type(np.pi)
    with np.errstate(all="ignore"):
        M = np.remainder(M, two_pi)  # :31
        high = M > pi  # :34
        M = np.where(high, two_pi - M, M)  # :35
        ome = 1 - ecc  # :38
        E = starter(M, ecc, ome)
        E = refine(M, ecc, ome, E)
        E = np.where(high, two_pi - E, E)  # :43
        tan_half_f = np.sqrt((1 + ecc) / (1 - ecc)) * np.tan(0.5 * E)  # :46
        tan2_half_f = np.square(tan_half_f)
        denom = 1 / (1 + tan2_half_f)  # :52
        sinf = 2 * tan_half_f * denom
        cosf = (1 - tan2_half_f) * denom
    sinf = sinf.astype(dt)
    cosf = cosf.astype(dt)
    if return_E:
        return sinf, cosf, E.astype(dt)
    return sinf, cosf


def kepler_jvp(M, ecc, M_dot, e_dot):
    """Closed-form tangents, kepler.py:59-79."""
    sinf, cosf = kepler(M, ecc)
    ecosf = ecc * cosf
    ome2 = 1 - ecc**2
    f_dot = M_dot * (1 + ecosf) ** 2 / ome2**1.5
    f_dot = f_dot + e_dot * (2 + ecosf) * sinf / ome2
    return (sinf, cosf), (cosf * f_dot, -sinf * f_dot)


# --------------------------------------------------------------------------
# Limb-darkened flux  (src/jaxoplanet/core/limb_dark.py)
# --------------------------------------------------------------------------
def greens_basis_transform(u, xp=NP):
    """u -> g, limb_dark.py:87-99."""
    if xp is NP:
        u = np.asarray(u)
        dtype = u.dtype if u.dtype.kind == "f" else np.float64
        u = np.concatenate((-np.ones(1, dtype=dtype), u.astype(dtype)))
        size = len(u)
        i = np.arange(size)
        arg = binom(i[None, :], i[:, None]).astype(dtype) @ u
        p = ((-1) ** (i + 1)).astype(dtype) * arg
        g = [np.zeros((), dtype=dtype) for _ in range(size + 2)]
        for n in range(size - 1, 1, -1):
            g[n] = p[n] / (n + 2) + g[n + 2]
        g[1] = p[1] + 3 * g[3]
        g[0] = p[0] + 2 * g[2]
        return np.stack(g[:-2])
    import torch

    u = torch.cat((-torch.ones(1, dtype=u.dtype), u))
    size = len(u)
    i = np.arange(size)
    Bm = torch.as_tensor(binom(i[None, :], i[:, None]), dtype=u.dtype)
    arg = Bm @ u
    sign = torch.as_tensor(((-1.0) ** (i + 1)), dtype=u.dtype)
    p = sign * arg
    g = [torch.zeros((), dtype=u.dtype) for _ in range(size + 2)]
    for n in range(size - 1, 1, -1):
        g[n] = p[n] / (n + 2) + g[n + 2]
    g[1] = p[1] + 3 * g[3]
    g[0] = p[0] + 2 * g[2]
    return torch.stack(g[:-2])


def kite_area(a, b, c, xp=NP):
    """limb_dark.py:187-196."""

    def sort2(a, b):
        return xp.minimum(a, b), xp.maximum(a, b)

    a, b = sort2(a, b)
    b, c = sort2(b, c)
    a, b = sort2(a, b)
    square_area = (a + (b + c)) * (c - (a - b)) * (c + (a - b)) * (a + (b - c))
    return xp.zero_safe_sqrt(xp.maximum(xp.zeros_like(square_area), square_area))


def kappas(b, r, xp=NP):
    """limb_dark.py:102-108."""
    b2 = xp.square(b)
    factor = (r - 1) * (r + 1)
    cond = (b > xp.abs(1 - r)) & (b < 1 + r)
    b_ = xp.where(cond, b, xp.ones_like(b))
    area = xp.where(cond, kite_area(r, b_, xp.ones_like(r), xp), xp.zeros_like(r))
    return area, xp.arctan2(area, b2 + factor), xp.arctan2(area, b2 - factor)


def s0s2(b, r, b2, r2, area, kappa0, kappa1, xp=NP):
    """limb_dark.py:111-137."""
    pi = xp.pi
    bpr = b + r
    onembpr2 = (1 + bpr) * (1 - bpr)
    eta2 = 0.5 * r2 * (r2 + 2 * b2)

    s0_lrg = pi * (1 - r2)
    s2_lrg = 2 * s0_lrg + 4 * pi * (eta2 - 0.5)

    Alens = kappa1 + r2 * kappa0 - area * 0.5
    s0_sml = pi - Alens
    s2_sml = 2 * s0_sml + 2 * (
        -(pi - kappa1) + 2 * eta2 * kappa0 - 0.25 * area * (1 + 5 * r2 + b2)
    )
    delta = 4 * b * r
    cond = (onembpr2 + delta) > delta
    return xp.where(cond, s0_lrg, s0_sml), xp.where(cond, s2_lrg, s2_sml)


def p_integral(order, l_max, b, r, b2, r2, kappa0, xp=NP):
    """limb_dark.py:140-184.  Returns an array with leading axis l_max-? terms:
    index 0 -> the l=1 term, indices 1.. -> n = 3..l_max."""
    eps = xp.eps(b)
    factor = 4 * b * r
    k2_cond = factor < 10 * eps
    factor = xp.where(k2_cond, xp.ones_like(factor), factor)
    k2 = xp.maximum(xp.zeros_like(factor), (1 - r2 - b2 + 2 * b * r) / factor)

    roots, weights = roots_legendre(order)
    roots = xp.asarray(roots, dtype=b.dtype)
    weights = xp.asarray(weights, dtype=b.dtype)
    shp = (order,) + (1,) * b.ndim
    roots = roots.reshape(shp)
    weights = weights.reshape(shp)

    rng = 0.5 * kappa0
    phi = rng * roots  # (order, ...)
    c = xp.cos(phi + 0.5 * kappa0)
    s = xp.sin((phi + rng) / 2)
    s2 = xp.square(s)

    terms = []
    if l_max >= 1:
        omz2 = xp.maximum(xp.zeros_like(c), r2 + b2 - 2 * b * r * c)
        z2 = 1 - omz2
        m = z2 < 10 * eps
        z2 = xp.where(m, xp.ones_like(z2), z2)
        z3 = z2 * xp.zero_safe_sqrt(z2)
        cond = omz2 < 10 * eps
        omz2 = xp.where(cond, xp.ones_like(z2), omz2)
        result = 2 * r * (r - b * c) * (1 - z3) / (3 * omz2)
        terms.append(xp.where(cond, xp.zeros_like(result), result))
    if l_max >= 3:
        f0 = xp.maximum(
            xp.zeros_like(s2), xp.where(k2_cond, 1 - r2 + 0 * s2, factor * (k2 - s2))
        )
        for n in range(3, l_max + 1):
            f = f0 ** (0.5 * n)
            f = f * (2 * r * (r - b + 2 * b * s2))
            terms.append(f)
    out = [rng * (t * weights).sum(0) for t in terms]
    return out


def solution_vector(l_max, b, r, order=10, xp=NP):
    """limb_dark.py:50-84; returns list [s_0 .. s_lmax] of arrays."""
    pi = xp.pi
    b = xp.abs(b)
    r = xp.abs(r)
    area, kappa0, kappa1 = kappas(b, r, xp)

    no_occ = b >= 1 + r
    full_occ = 1 + b <= r
    cond = no_occ | full_occ
    b_ = xp.where(cond, xp.ones_like(b), b)

    b2 = xp.square(b_)
    r2 = xp.square(r)

    s0, s2 = s0s2(b_, r, b2, r2, area, kappa0, kappa1, xp)
    s0 = xp.where(no_occ, pi * xp.ones_like(s0), s0)
    s0 = xp.where(full_occ, xp.zeros_like(s0), s0)
    s2 = xp.where(cond, xp.zeros_like(s2), s2)

    s = [s0]
    if l_max >= 1:
        P = p_integral(order, l_max, b_, r, b2, r2, kappa0, xp)
        P = [xp.where(cond, xp.zeros_like(p), p) for p in P]
        s.append(-P[0] - 2 * (kappa1 - pi) / 3)
        if l_max >= 2:
            s.append(s2)
        if l_max >= 3:
            s.extend(-p for p in P[1:])
    return s


def ld_light_curve(u, b, r, order=10, xp=NP):
    """flux - 1 for polynomial limb darkening; limb_dark.py:19-47."""
    if xp is NP:
        u = np.asarray(u)
        b = np.asarray(b)
        r = np.asarray(r)
        dt = np.result_type(b.dtype, r.dtype, u.dtype if u.size else b.dtype)
        if dt.kind != "f":
            dt = np.dtype(np.float64)
        u = u.astype(dt)
        b, r = np.broadcast_arrays(b.astype(dt), r.astype(dt))
    else:
        import torch

        b, r = torch.broadcast_tensors(b, r)
    assert u.ndim == 1
    if u.shape[0] == 0:
        g = [1.0 / xp.pi]
    else:
        g = greens_basis_transform(u, xp)
        g = g / (xp.pi * (g[0] + g[1] / 1.5))
    l_max = len(g) - 1
    with np.errstate(all="ignore"):
        s = solution_vector(l_max, b, r, order, xp)
        out = s[0] * g[0]
        for k in range(1, l_max + 1):
            out = out + s[k] * g[k]
    return out - 1


# --------------------------------------------------------------------------
# Keplerian bodies  (src/jaxoplanet/orbits/keplerian.py)
# --------------------------------------------------------------------------
def central(mass=None, radius=None, density=None):
    """Two-of-three rule, keplerian.py:30-70.  Returns (mass, radius, density)."""
    if radius is None and mass is None:
        radius = 1.0
        if density is None:
            mass = 1.0
    msg = "Values must be provided for exactly two of mass, radius, and density"
    if density is None:
        if mass is None or radius is None:
            raise ValueError(msg)
        density = 3 * mass / (4 * np.pi * radius**3)
    elif radius is None:
        if mass is None or density is None:
            raise ValueError(msg)
        radius = (3 * mass / (4 * np.pi * density)) ** (1 / 3)
    elif mass is None:
        mass = 4 * np.pi * radius**3 * density / 3.0
    return mass, radius, density


def orbital_body(
    *,
    central_mass=1.0,
    central_radius=1.0,
    period=None,
    semimajor=None,
    time_transit=None,
    time_peri=None,
    inclination=None,
    impact_param=None,
    eccentricity=None,
    omega_peri=None,
    sin_omega_peri=None,
    cos_omega_peri=None,
    asc_node=None,
    sin_asc_node=None,
    cos_asc_node=None,
    mass=None,
    radius=None,
    xp=NP,
):
    """Derived per-body constants, keplerian.py:262-340.  All inputs may be
    arrays (one entry per parameter set).  Returns a dict."""
    pi = xp.pi
    total_mass = central_mass if mass is None else mass + central_mass
    mass_factor = G_GRAV * total_mass
    if semimajor is None:
        semimajor = xp.cbrt(mass_factor * period**2 / (4 * pi**2))  # :275
    else:
        period = 2 * pi * semimajor * xp.sqrt(semimajor / mass_factor)  # :281

    if omega_peri is not None:
        sin_omega_peri = xp.sin(omega_peri)
        cos_omega_peri = xp.cos(omega_peri)
    if asc_node is not None:
        sin_asc_node = xp.sin(asc_node)
        cos_asc_node = xp.cos(asc_node)

    if eccentricity is None:
        M0 = 0.5 * pi + 0 * period  # :302
        incl_factor = 1
    else:
        opsw = 1 + sin_omega_peri
        E0 = 2 * xp.arctan2(
            xp.sqrt(1 - eccentricity) * cos_omega_peri,
            xp.sqrt(1 + eccentricity) * opsw,
        )  # :308-311
        M0 = E0 - eccentricity * xp.sin(E0)  # :312
        ome2 = 1 - eccentricity**2
        incl_factor = (1 + eccentricity * sin_omega_peri) / ome2  # :315

    dcosidb = incl_factor * central_radius / semimajor  # :318
    if impact_param is not None:
        cos_incl = dcosidb * impact_param
        sin_incl = xp.sqrt(1 - cos_incl**2)
    elif inclination is not None:
        cos_incl = xp.cos(inclination)
        sin_incl = xp.sin(inclination)
        impact_param = cos_incl / dcosidb
    else:
        impact_param = 0 * period
        cos_incl = 0 * period
        sin_incl = 1 + 0 * period

    time_ref = -M0 * period / (2 * pi)  # :334
    if time_transit is None:
        if time_peri is not None:
            time_transit = time_peri - time_ref
        else:
            time_transit = 0 * time_ref
    return dict(
        central_mass=central_mass,
        central_radius=central_radius,
        period=period,
        semimajor=semimajor,
        time_transit=time_transit,
        time_ref=time_ref,
        eccentricity=eccentricity,
        sin_omega_peri=sin_omega_peri,
        cos_omega_peri=cos_omega_peri,
        sin_asc_node=sin_asc_node,
        cos_asc_node=cos_asc_node,
        sin_inclination=sin_incl,
        cos_inclination=cos_incl,
        impact_param=impact_param,
        mass=mass,
        radius=radius,
        total_mass=total_mass,
    )


def true_anomaly(body, t, xp=NP):
    """keplerian.py:537-541."""
    M = 2 * xp.pi * ((t - body["time_transit"]) - body["time_ref"]) / body["period"]
    if body["eccentricity"] is None:
        return xp.sin(M), xp.cos(M)
    return xp.kepler(M, body["eccentricity"] + 0 * M)


def rotate_vector(body, x, y, include_inclination=True):
    """keplerian.py:543-588."""
    if body["eccentricity"] is None:
        x1, y1 = x, y
    else:
        x1 = body["cos_omega_peri"] * x - body["sin_omega_peri"] * y
        y1 = body["sin_omega_peri"] * x + body["cos_omega_peri"] * y
    if include_inclination:
        x2 = x1
        y2 = body["cos_inclination"] * y1
        Z = -body["sin_inclination"] * y1
    else:
        x2, y2, Z = x1, y1, -y1
    if body["cos_asc_node"] is None:
        return x2, y2, Z
    X = body["cos_asc_node"] * x2 - body["sin_asc_node"] * y2
    Y = body["sin_asc_node"] * x2 + body["cos_asc_node"] * y2
    return X, Y, Z


def position(body, t, semimajor, xp=NP):
    """Position half of keplerian.py:590-637 (parallax None)."""
    r0 = semimajor
    sinf, cosf = true_anomaly(body, t, xp)
    e = body["eccentricity"]
    if e is not None:
        r0 = r0 * ((1 - e**2) / (1 + e * cosf))  # :630
    return rotate_vector(body, r0 * cosf, r0 * sinf)


def relative_position(body, t, xp=NP):
    """keplerian.py:402-419."""
    return position(body, t, -body["semimajor"], xp)


def velocity(body, t, kind="relative", xp=NP):
    """Velocity half of keplerian.py:590-637 for semiamplitude None, parallax None.
    kind: '' (body, barycentric) | 'central' | 'relative'   (:439-506)."""
    if kind == "":
        m = -body["central_mass"]
    elif kind == "central":
        m = body["mass"]
    else:
        m = -body["total_mass"]
    k0 = 2 * xp.pi * body["semimajor"] * m / (body["total_mass"] * body["period"])
    e = body["eccentricity"]
    if e is not None:
        k0 = k0 / xp.sqrt(1 - e**2 + 0 * body["period"])
    sinf, cosf = true_anomaly(body, t, xp)
    if e is None:
        v1, v2 = -k0 * sinf, k0 * cosf
    else:
        v1, v2 = -k0 * sinf, k0 * (cosf + e)
    return rotate_vector(body, v1, v2)


# --------------------------------------------------------------------------
# Light curves  (src/jaxoplanet/light_curves/limb_dark.py, transforms.py)
# --------------------------------------------------------------------------
def body_light_curve(body, u, t, order=10, xp=NP):
    """One body's column of light_curves/limb_dark.py:49-60 (flux - 1, masked)."""
    r_star = body["central_radius"]
    x, y, z = relative_position(body, t, xp)
    b = xp.sqrt(x**2 + y**2) / r_star
    r = body["radius"] / r_star + 0 * b
    lc = ld_light_curve(u, b, r, order=order, xp=xp)
    return xp.where(z > 0, lc, xp.zeros_like(lc))


def system_light_curve(bodies, u, t, order=10, xp=NP):
    """t.shape + (n_bodies,), light_curves/limb_dark.py:15-62."""
    cols = [body_light_curve(bd, u, t, order, xp) for bd in bodies]
    return xp.stack(cols, axis=-1)


def emission_light_curve(bodies, star_u, surfaces, t, order=10, xp=NP):
    """src/jaxoplanet/light_curves/emission.py:11-52: [star: 1 + sum of transits, amplitude_i (1 +
    occultation_i) ...] per time; `surfaces` = [(u_i, amplitude_i)], bodies = orbital_body(...) dicts."""
    t = np.asarray(t, dtype=np.float64)
    star = system_light_curve(bodies, np.asarray(star_u, dtype=np.float64), t, order=order).sum(-1) + 1  # :31-38
    cols = [star]
    r_star = bodies[0]["central_radius"]
    for body, (u_b, amp) in zip(bodies, surfaces):
        x, y, z = relative_position(body, t)  # :41
        b = np.sqrt(x**2 + y**2) / body["radius"]  # :42
        r = r_star / body["radius"] * np.ones_like(b)  # :43
        lc = ld_light_curve(np.asarray(u_b, dtype=np.float64), b, r, order=order)  # :45-46
        cols.append(amp * (1 + np.where(z < 0, lc, 0.0)))  # :47
    return np.stack(cols, -1)  # :51


def integrate_stencil(exposure_time, order=0, num_samples=7, dtype=np.float64):
    """Offsets and weights of transforms.py:67-89."""
    num_samples = int(num_samples)
    num_samples += 1 - num_samples % 2
    stencil = np.ones(num_samples, dtype=dtype)
    if order == 0:
        dt = np.linspace(-0.5, 0.5, 2 * num_samples + 1, dtype=dtype)[1:-1:2]
    elif order == 1:
        dt = np.linspace(-0.5, 0.5, num_samples, dtype=dtype)
        stencil = 2 * stencil
        stencil[0] = 1
        stencil[-1] = 1
    elif order == 2:
        dt = np.linspace(-0.5, 0.5, num_samples, dtype=dtype)
        stencil[1:-1:2] = 4
        stencil[2:-1:2] = 2
    else:
        raise ValueError(
            "The parameter 'order' in 'integrate_exposure_time' must be 0, 1, or 2"
        )
    dt = dt * exposure_time
    stencil = stencil / np.sum(stencil)
    return dt, stencil


def integrate(func, exposure_time=None, order=0, num_samples=7):
    """transforms.py:21-116: returns time -> sum_j w_j func(t + dt_j)."""
    if exposure_time is None:
        return func
    dt, stencil = integrate_stencil(exposure_time, order, num_samples)

    def wrapped(time):
        time = np.asarray(time)
        res = None
        # jnp.dot(stencil, result): ordered accumulation over the sub-samples
        for d, w in zip(dt, stencil):
            term = w * np.asarray(func(time + d))
            res = term if res is None else res + term
        return res

    return wrapped


def gaussian_loglike(model_flux, y, sigma):
    """sum over time of Normal(model, sigma).log_prob(y); the NumPyro pattern of
    docs/tutorials/transit.ipynb cell 7 (user code)."""
    resid = (y - model_flux) / sigma
    return np.sum(-0.5 * resid**2 - np.log(sigma) - 0.5 * np.log(2 * np.pi), axis=-1)
