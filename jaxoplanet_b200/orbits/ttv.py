"""``TTVOrbit``: a ``TransitOrbit`` with transit timing variations; drop-in for
src/jaxoplanet/orbits/ttv.py:46-225 (same arguments, attributes and error messages).

The set-up (linear ephemeris fit, bin edges around each observed transit) is small host work on the
transit lists; per time sample the searchsorted time warp, the linear motion and the limb-darkened
flux run fused in the CUDA kernel ``k_transit_orbit`` (``limb_dark_light_curve(orbit, u)(t)``).
``relative_position`` is kept for the orbit protocol (device tensor ops).
"""

from __future__ import annotations

import numpy as np
import torch

from jaxoplanet_b200 import _array as A
from jaxoplanet_b200.orbits.transit import TransitOrbit

__all__ = ["TTVOrbit", "compute_expected_transit_times"]


def _np1(x):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.atleast_1d(np.asarray(x, dtype=np.float64))


def _linear_ephemeris(transit_times, indices):
    """ttv.py:8-17: least-squares line through (index, transit time)."""
    n = transit_times.shape[0]
    X = np.vstack([np.ones(n), np.asarray(indices, dtype=np.float64)])
    beta = np.linalg.solve(X @ X.T, X @ transit_times)
    intercept, slope = beta
    ttvs = transit_times - (intercept + slope * indices)
    return intercept, slope, ttvs


def _bin_edges_values(tt, period):
    """ttv.py:20-33."""
    midpoints = 0.5 * (tt[1:] + tt[:-1])
    edges = np.concatenate([[tt[0] - 0.5 * period], midpoints, [tt[-1] + 0.5 * period]])
    values = np.concatenate([[tt[0]], tt, [tt[-1]]])
    return edges, values


class TTVOrbit(TransitOrbit):
    def __init__(self, *, period=None, duration=None, speed=None, time_transit=None, impact_param=None,
                 radius_ratio=None, transit_times=None, transit_inds=None, ttvs=None, delta_log_period=None):
        if ttvs is not None and transit_times is not None:
            raise ValueError("Supply either ttvs or transit_times, not both.")
        if ttvs is None and transit_times is None:
            raise ValueError("You must supply either transit_times or ttvs.")

        if transit_times is not None:
            self.transit_times = tuple(_np1(tt) for tt in transit_times)
            if transit_inds is None:
                self.transit_inds = tuple(np.arange(tt.shape[0]) for tt in self.transit_times)
            else:
                self.transit_inds = tuple(np.asarray(i) for i in transit_inds)
            t0_list, period_list, ttvs_list = [], [], []
            for tt, inds in zip(self.transit_times, self.transit_inds):
                t0_i, period_i, ttv_i = _linear_ephemeris(tt, inds)
                t0_list.append(t0_i)
                period_list.append(period_i)
                ttvs_list.append(ttv_i)
            self.t0 = np.array(t0_list)
            self.ttv_period = np.array(period_list)
            self.ttvs = tuple(ttvs_list)
            if time_transit is None:
                time_transit = self.t0
            if period is None:
                if delta_log_period is not None:
                    period = np.exp(np.log(self.ttv_period) + delta_log_period)
                else:
                    period = self.ttv_period
        else:
            if period is None:
                raise ValueError("When supplying ttvs, period must be provided.")
            if time_transit is None:
                time_transit = 0.0
            self.ttvs = tuple(_np1(ttv) - np.mean(_np1(ttv)) for ttv in ttvs)
            if transit_inds is None:
                self.transit_inds = tuple(np.arange(ttv.shape[0]) for ttv in self.ttvs)
            else:
                self.transit_inds = tuple(np.asarray(i) for i in transit_inds)
            self.t0 = _np1(time_transit)
            self.ttv_period = _np1(period)
            tts = []
            for ttv, inds in zip(self.ttvs, self.transit_inds):
                if self.t0.shape[0] > 1:
                    t0_i = self.t0[len(tts)]
                    period_i = self.ttv_period[len(tts)]
                else:
                    t0_i = self.t0.item()
                    period_i = self.ttv_period.item()
                tts.append(t0_i + period_i * inds + ttv)
            self.transit_times = tuple(tts)

        super().__init__(period=period if period is not None else self.ttv_period, duration=duration, speed=speed,
                         time_transit=time_transit, impact_param=impact_param, radius_ratio=radius_ratio)

        edges, values = [], []
        for tt, per in zip(self.transit_times, self.ttv_period):
            e, v = _bin_edges_values(tt, per)
            edges.append(e)
            values.append(v)
        self._bin_edges = tuple(edges)
        self._bin_values = tuple(values)

    # -- tables for the fused kernel (include/jaxoplanet_b200.h, K6) ---------------------------------
    def _ttv_tables(self):
        offsets = np.zeros(len(self._bin_edges) + 1, dtype=np.int32)
        offsets[1:] = np.cumsum([e.shape[0] for e in self._bin_edges])
        t0 = np.broadcast_to(self.t0, (len(self._bin_edges),)).astype(np.float64)
        return (np.concatenate(self._bin_edges), np.concatenate(self._bin_values), offsets, np.ascontiguousarray(t0))

    def _get_model_dt(self, t: torch.Tensor) -> torch.Tensor:
        """ttv.py:194-206: per planet, the observed transit time of the bin each t falls in."""
        out = []
        for e, v in zip(self._bin_edges, self._bin_values):
            ed = torch.as_tensor(e, device=t.device)
            inds = torch.searchsorted(ed, t.contiguous())
            out.append(torch.as_tensor(v, device=t.device)[inds])
        return torch.stack(out)

    def _warp_times(self, t: torch.Tensor) -> torch.Tensor:
        dt = self._get_model_dt(t)
        t0 = torch.as_tensor(self.t0, device=t.device)
        return t - (dt - t0.reshape((-1,) + (1,) * (dt.ndim - 1)))

    def relative_position(self, t, parallax=None):
        t_in = t
        t = t if isinstance(t, torch.Tensor) else torch.as_tensor(np.asarray(t, dtype=np.float64), device=A.device())
        warped = self._warp_times(t)  # (n_planets, *t.shape)
        # TransitOrbit.relative_position broadcasts the planet axis last; warped carries it first
        n_p = warped.shape[0]
        outs = []
        for p in range(n_p):
            period = self.period.reshape(-1)[p if self.period.numel() > 1 else 0].to(t.device)
            half = 0.5 * period
            ref = self.time_transit.reshape(-1)[p if self.time_transit.numel() > 1 else 0].to(t.device) - half
            dt = torch.remainder(warped[p] - ref, period) - half
            x = self.speed.reshape(-1)[p if self.speed.numel() > 1 else 0].to(t.device) * dt
            y = torch.zeros_like(dt) + self.impact_param.reshape(-1)[p if self.impact_param.numel() > 1 else 0].to(t.device)
            m = torch.abs(dt) < 0.5 * self.duration.reshape(-1)[p if self.duration.numel() > 1 else 0].to(t.device)
            outs.append(torch.stack([x, y, torch.where(m, torch.ones_like(dt), -torch.ones_like(dt))]))
        out = torch.stack(outs, dim=-1)  # (3, *t.shape, n_planets)
        out = torch.squeeze(out, -1) if n_p == 1 else out
        res = tuple(out[k] for k in range(3))
        return tuple(A.like_input(r, t_in) for r in res)

    @property
    def linear_t0(self):
        return np.atleast_1d(self.t0)

    @property
    def linear_period(self):
        return np.atleast_1d(self.ttv_period)


def compute_expected_transit_times(min_time, max_time, period, t0):
    """ttv.py:240-272: expected (linear) transit times of each planet inside [min_time, max_time]."""
    period = np.atleast_1d(np.asarray(period, dtype=np.float64))
    t0 = np.atleast_1d(np.asarray(t0, dtype=np.float64))
    out = []
    for p, t0_val in zip(period, t0):
        i_min = int(np.ceil((min_time - t0_val) / p))
        i_max = int(np.floor((max_time - t0_val) / p))
        times = t0_val + p * np.arange(i_min, i_max + 1)
        out.append(times[(times >= min_time) & (times <= max_time)])
    return tuple(out)
