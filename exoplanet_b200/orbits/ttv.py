"""``TTVOrbit``: a Keplerian orbit whose time axis is re-zeroed on each (shifted) transit time,
so transit-timing variations move the transits without changing their shape.  Host-side mirror
of /root/reference/src/exoplanet/orbits/ttv.py:35-187 over torch; only ``_warp_times`` differs
from ``KeplerianOrbit`` (one ``searchsorted`` per cadence before the same Kepler solve), so light
curves of a TTVOrbit go through the same fused kernels with the transit-time table as an extra argument
(``ttv_table()`` -> ``xo_transit_lc_fwd_ttv`` / ``xo_transit_lc_vjp_ttv``); every other method uses the inherited
op-by-op code."""
import numpy as np
import torch

from .keplerian import KeplerianOrbit, _default_device, _tensor

__all__ = ["TTVOrbit", "compute_expected_transit_times"]


def compute_expected_transit_times(min_time, max_time, period, t0):
    """Per planet, the strictly periodic transit times t0 + k P inside [min_time, max_time]."""
    out = []
    for P, T0 in zip(np.atleast_1d(period), np.atleast_1d(t0)):
        k = np.arange(np.floor((min_time - T0) / P), np.ceil((max_time - T0) / P), 1)
        times = T0 + P * k
        out.append(times[(times >= min_time) & (times <= max_time)])
    return out


def _linear_fit(index, times):
    """Least-squares (t0, period) of times ~ t0 + period * index."""
    n = times.shape[0]
    sx, sxx, sy, sxy = index.sum(), (index * index).sum(), times.sum(), (index * times).sum()
    det = n * sxx - sx * sx
    return (sxx * sy - sx * sxy) / det, (n * sxy - sx * sy) / det


class TTVOrbit(KeplerianOrbit):
    """Give exactly one of ``ttvs`` (O-C per transit and planet) or ``transit_times``; optional
    ``transit_inds`` (zero-based transit numbers when some transits are missing) and, with
    ``transit_times``, ``period`` / ``delta_log_period`` to decouple the transit shape's period
    from the mean spacing ``ttv_period`` (ttv.py:44-71)."""

    def __init__(self, *args, ttvs=None, transit_times=None, transit_inds=None, **kwargs):
        if ttvs is None and transit_times is None:
            raise ValueError("one of 'ttvs' or 'transit_times' must be defined")
        dev = torch.device(kwargs["device"]) if kwargs.get("device") is not None else _default_device()
        as_index = lambda x: torch.as_tensor(np.asarray(x), device=dev).to(torch.int64).reshape(-1)
        if ttvs is not None:
            self.ttvs = [_tensor(v, dev).reshape(-1) for v in ttvs]
            self.transit_inds = [torch.arange(v.shape[0], device=dev) for v in self.ttvs] if transit_inds is None else [
                as_index(i) for i in transit_inds]
        else:
            self.transit_times = [_tensor(v, dev).reshape(-1) for v in transit_times]
            self.transit_inds = [torch.arange(v.shape[0], device=dev) for v in self.transit_times] if transit_inds is None else [
                as_index(i) for i in transit_inds]
            fits = [_linear_fit(i.to(torch.float64), tt) for i, tt in zip(self.transit_inds, self.transit_times)]
            self.ttvs = [tt - (t0 + P * i.to(torch.float64)) for (t0, P), i, tt in zip(fits, self.transit_inds, self.transit_times)]
            kwargs["t0"] = torch.stack([f[0] for f in fits])
            self.ttv_period = torch.stack([f[1] for f in fits])
            if "period" not in kwargs:
                dlp = kwargs.pop("delta_log_period", None)
                kwargs["period"] = self.ttv_period if dlp is None else torch.exp(torch.log(self.ttv_period) + _tensor(dlp, dev))
        kwargs.setdefault("device", dev)
        super().__init__(*args, **kwargs)
        if ttvs is not None:
            self.ttv_period = self.period
            self.transit_times = [self.t0[i] + self.period[i] * k.to(torch.float64) + v
                                  for i, (k, v) in enumerate(zip(self.transit_inds, self.ttvs))]
        # every transit number up to the last labelled one, unlabelled ones on the linear ephemeris
        self.all_transit_times = []
        for i, k in enumerate(self.transit_inds):
            full = self.t0[i] + self.period[i] * torch.arange(int(k.max()) + 1, device=dev, dtype=torch.float64)
            self.all_transit_times.append(full.index_put((k,), self.transit_times[i]))
        # nearest-transit lookup: bin edges half-way between consecutive transits
        half = 0.5 * self.ttv_period
        self._bin_edges = [torch.cat([(tt[:1] - half[i]), 0.5 * (tt[1:] + tt[:-1]), (tt[-1:] + half[i])])
                           for i, tt in enumerate(self.all_transit_times)]
        self._bin_values = [torch.cat([tt[:1], tt, tt[-1:]]) for tt in self.all_transit_times]

    def _nearest_transit(self, t):
        t = t.contiguous()
        cols = [vals[torch.searchsorted(edges.contiguous(), t)] for edges, vals in zip(self._bin_edges, self._bin_values)]
        return torch.stack(cols, dim=-1)

    def _warp_times(self, t, _pad=True):
        t = _tensor(t, self.device)
        return (t.unsqueeze(-1) if _pad else t) - self._nearest_transit(t)

    def ttv_table(self):
        """(times[P, NT], count[P] int32): the transit-time tables the fused kernels take (``xo_transit_*_ttv``) -- row p is
        ``all_transit_times[p]`` (differentiable), padded past ``count[p]`` by continuing the last transit with the period."""
        nt = max(int(tt.shape[0]) for tt in self.all_transit_times)
        rows = []
        for i, tt in enumerate(self.all_transit_times):
            n = int(tt.shape[0])
            if n < nt:
                extra = tt[-1] + self.period.reshape(-1)[i] * torch.arange(1, nt - n + 1, device=self.device, dtype=torch.float64)
                tt = torch.cat([tt, extra])
            rows.append(tt)
        count = torch.tensor([int(tt.shape[0]) for tt in self.all_transit_times], dtype=torch.int32, device=self.device)
        return torch.stack(rows), count

    # ``elements()`` is inherited: the element block's (t0, period) is the linear ephemeris transit j departs from
    # (transit j near t0 + j period), which is what the fused kernels' sweep tests against.
