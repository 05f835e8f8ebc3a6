"""``SimpleTransitOrbit``: a straight-line transit chord parameterised by the observables
(period, duration, t0, b, r/R*); no Kepler solve.  Host-side mirror of
/root/reference/src/exoplanet/orbits/simple.py:9-103 -- it feeds the same limb-darkening op
through ``LimbDarkLightCurve``'s generic (op-by-op) path."""
import torch

from .keplerian import _tensor

__all__ = ["SimpleTransitOrbit"]

_NO_DELAY = "Light travel time delay is not implemented for simple orbits"


class SimpleTransitOrbit:
    """Planets crossing the stellar disk at constant speed: the centre covers the chord
    2 R* sqrt((1 + ror)^2 - b^2) in ``duration`` (simple.py:31-33)."""

    def __init__(self, period, duration, t0=0.0, b=0.0, r_star=1.0, ror=0, device=None):
        self.period = _tensor(period, device)
        self.device = dev = self.period.device
        self.duration, self.t0, self.b = _tensor(duration, dev), _tensor(t0, dev), _tensor(b, dev)
        self.r_star = _tensor(r_star, dev)
        ror = _tensor(ror, dev)
        self._b_norm = self.b * self.r_star
        self.speed = 2.0 * self.r_star * torch.sqrt((1.0 + ror) ** 2 - self.b ** 2) / self.duration
        self._half_period = 0.5 * self.period
        self._ref_time = self.t0 - self._half_period

    def _time_since_transit(self, t):
        t = _tensor(t, self.device).unsqueeze(-1)
        return torch.remainder(t - self._ref_time, self.period) - self._half_period

    def get_star_position(self, t, light_delay=False):
        zero = torch.zeros_like(_tensor(t, self.device))
        return zero, zero, zero

    def get_relative_position(self, t, light_delay=False):
        """(x along the chord, y = b R*, z = +1 within half a duration of mid-transit else -1)."""
        if light_delay:
            raise NotImplementedError(_NO_DELAY)
        dt = self._time_since_transit(t)
        front = dt.abs() < 0.5 * self.duration
        z = torch.where(front, torch.ones_like(dt), -torch.ones_like(dt))
        return torch.squeeze(self.speed * dt), torch.squeeze(self._b_norm + torch.zeros_like(dt)), torch.squeeze(z)

    def get_planet_position(self, t, light_delay=False):
        return self.get_relative_position(t, light_delay=False)

    def _no_velocity(self, *args, **kwargs):
        raise NotImplementedError("a SimpleTransitOrbit has no velocity")

    get_planet_velocity = get_star_velocity = get_radial_velocity = _no_velocity

    def in_transit(self, t, r=None, texp=None, light_delay=False):
        """Indices with |dt| below half a duration (or the chord of radius r + R*), plus texp/2."""
        if light_delay:
            raise NotImplementedError(_NO_DELAY)
        dt = self._time_since_transit(t)
        if r is None:
            tol = 0.5 * self.duration
        else:
            tol = torch.sqrt((_tensor(r, self.device) + self.r_star) ** 2 - self._b_norm ** 2) / self.speed
        if texp is not None:
            tol = tol + 0.5 * _tensor(texp, self.device)
        mask = (dt.abs() < tol).reshape(dt.shape[0], -1).any(dim=-1)
        return torch.nonzero(mask).reshape(-1)
