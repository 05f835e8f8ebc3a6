"""``KeplerianOrbit`` over PyTorch CUDA tensors -- the host-side mirror of the reference
class of the same name (/root/reference/src/exoplanet/orbits/keplerian.py:26-281), so
models written against the reference read the same here.

Design (not a port): the constructor reduces whatever combination of inputs it is given
to one table of per-planet *derived elements* (``self.elements()``, the ``XO_EL_*`` block of
include/xo_b200.h) in differentiable torch; the position / velocity getters are thin
algebra on ``ops.kepler`` (the CUDA op with its VJP), and ``LimbDarkLightCurve`` hands the
element table straight to the fused kernels.  Units: R_sun, M_sun, days (the reference's
internal units, keplerian.py:30-34); astropy unit handling is out of scope.
"""
import warnings

import numpy as np
import torch

from .. import elements as _el
from .. import ops

__all__ = ["KeplerianOrbit", "get_true_anomaly", "get_aor_from_transit_duration"]

_TWO_PI = 2.0 * np.pi
#: 1 R_sun/day in m/s (IAU nominal solar radius 6.957e8 m), the default RV unit (keplerian.py:673-675)
M_PER_S_PER_RSUN_DAY = 6.957e8 / 86400.0


def _default_device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def _tensor(x, device=None):
    if isinstance(x, torch.Tensor):
        return x.to(dtype=torch.float64) if device is None else x.to(dtype=torch.float64, device=device)
    return torch.as_tensor(np.asarray(x, dtype=np.float64), device=device or _default_device())


def get_aor_from_transit_duration(duration, period, b, ror=None):
    """a/R_star implied by a circular orbit of given transit duration, and d(a/R)/d(duration)
    (same contract as keplerian.py:822-846)."""
    duration, period, b = _tensor(duration), _tensor(period), _tensor(b)
    k1 = 1.0 + (_tensor(0.0, duration.device) if ror is None else _tensor(ror, duration.device))
    phase = np.pi * duration / period
    sp, cp = torch.sin(phase), torch.cos(phase)
    root = torch.sqrt(k1 * k1 - (b * cp) ** 2)
    return root / sp, np.pi * cp * (b * b - k1 * k1) / (root * period * sp * sp)


def get_true_anomaly(M, e, **kwargs):
    """atan2 of the op's (sin f, cos f) (keplerian.py:807-819)."""
    sinf, cosf = ops.kepler(M, e)
    return torch.atan2(sinf, cosf)


def _stellar_and_orbital_scales(a, period, rho_star, r_star, m_star, m_planet, device):
    """Resolve (a, period, rho_star[g/cc], r_star, m_star, m_planet) from the combinations the
    reference accepts (keplerian.py:849-934): Kepler's third law ties a, P and the total mass."""
    if a is None and period is None:
        raise ValueError("values must be provided for at least one of a and period")
    T = lambda x: None if x is None else _tensor(x, device)
    a, period, rho_star, r_star, m_star = T(a), T(period), T(rho_star), T(r_star), T(m_star)
    m_planet = T(0.0 if m_planet is None else m_planet)
    four_pi2 = 4.0 * np.pi ** 2
    if a is not None and period is not None:
        if rho_star is not None or m_star is not None:
            raise ValueError("if both a and period are given, you can't also define rho_star or m_star")
        r_star = _tensor(1.0, device) if r_star is None else r_star
        m_star = four_pi2 * a ** 3 / (_el.G_GRAV * period ** 2) - m_planet
        rho = m_star / (4.0 * np.pi * r_star ** 3 / 3.0)
        return a, period, rho * _el.GCC_PER_SUN, r_star, m_star, m_planet
    if r_star is None and m_star is None:
        r_star = _tensor(1.0, device)
        if rho_star is None:
            m_star = _tensor(1.0, device)
    if sum(v is None for v in (rho_star, r_star, m_star)) != 1:
        raise ValueError("values must be provided for exactly two of rho_star, m_star, and r_star")
    rho = None if rho_star is None else rho_star / _el.GCC_PER_SUN  # g/cc -> M_sun / R_sun^3
    if rho is None:
        rho = 3.0 * m_star / (4.0 * np.pi * r_star ** 3)
    elif r_star is None:
        r_star = (3.0 * m_star / (4.0 * np.pi * rho)) ** (1.0 / 3.0)
    else:
        m_star = 4.0 * np.pi * r_star ** 3 * rho / 3.0
    mtot = m_star + m_planet
    if a is None:
        a = (_el.G_GRAV * mtot * period ** 2 / four_pi2) ** (1.0 / 3.0)
    else:
        period = _TWO_PI * a ** 1.5 / torch.sqrt(_el.G_GRAV * mtot)
    return a, period, rho * _el.GCC_PER_SUN, r_star, m_star, m_planet


class KeplerianOrbit:
    """A system of bodies on Keplerian orbits around a common central body.

    Accepted inputs are those of the reference (keplerian.py:36-71): ``period`` or ``a``;
    ``t0`` or ``t_periastron``; one of ``incl`` / ``b`` / ``duration``; ``ecc`` with ``omega``
    (or ``sin_omega`` & ``cos_omega``); optional ``Omega``, ``m_planet``, and two of
    ``m_star`` / ``r_star`` / ``rho_star``.  Every attribute the reference exposes
    (``a, period, n, a_star, a_planet, K0, M0, cos_omega, sin_omega, cos_incl, sin_incl, incl,
    b, t0, t_periastron, tref, dcosidb, jacobians``) is available with the same meaning."""

    def __init__(self, period=None, a=None, t0=None, t_periastron=None, incl=None, b=None, duration=None, ecc=None,
                 omega=None, sin_omega=None, cos_omega=None, Omega=None, m_planet=0.0, m_star=None, r_star=None,
                 rho_star=None, ror=None, model=None, device=None, **kwargs):
        if kwargs:
            warnings.warn("unit keyword arguments are ignored: inputs are taken in R_sun / M_sun / day / g cm^-3")
        dev = torch.device(device) if device is not None else next(
            (x.device for x in (period, a, t0, t_periastron, ecc, b, incl) if isinstance(x, torch.Tensor)), _default_device())
        self.device = dev
        self.jacobians = {"duration": {}, "b": {}}
        daordtau = None
        if ecc is None and duration is not None:  # circular orbit parameterised by duration (keplerian.py:113-131)
            if b is None:
                raise ValueError("'b' must be provided for a circular orbit with a 'duration'")
            if ror is None:
                warnings.warn("When using the 'duration' parameter in KeplerianOrbit, the 'ror' parameter should also be provided.",
                              UserWarning)
            r_star = _tensor(1.0, dev) if r_star is None else _tensor(r_star, dev)
            aor, daordtau = get_aor_from_transit_duration(_tensor(duration, dev), _tensor(period, dev), _tensor(b, dev), ror=ror)
            a, duration = r_star * aor, None

        (self.a, self.period, self.rho_star, self.r_star, self.m_star, self.m_planet) = _stellar_and_orbital_scales(
            a, period, rho_star, r_star, m_star, m_planet, dev)
        self.m_total = self.m_star + self.m_planet
        self.n = _TWO_PI / self.period
        self.a_star = self.a * self.m_planet / self.m_total
        self.a_planet = -self.a * self.m_star / self.m_total
        self.K0 = self.n * self.a / self.m_total
        if daordtau is not None:
            dadtau = self.r_star * daordtau
            J = self.jacobians["duration"]
            J["a"], J["a_star"], J["a_planet"] = dadtau, dadtau * self.m_planet / self.m_total, -dadtau * self.m_star / self.m_total
            J["rho_star"] = 9.0 * np.pi * (self.a / self.r_star) ** 2 * daordtau * _el.GCC_PER_SUN / (_el.G_GRAV * self.period ** 2)

        self.Omega = None if Omega is None else _tensor(Omega, dev)
        if self.Omega is not None:
            self.cos_Omega, self.sin_Omega = torch.cos(self.Omega), torch.sin(self.Omega)

        if ecc is None:
            self.ecc = None
            self.M0 = 0.5 * np.pi + torch.zeros_like(self.n)
            shape_factor = 1.0
        else:
            self.ecc = _tensor(ecc, dev)
            if omega is not None:
                if sin_omega is not None and cos_omega is not None:
                    raise ValueError("either 'omega' or 'sin_omega' and 'cos_omega' can be provided")
                self.omega = _tensor(omega, dev)
                self.cos_omega, self.sin_omega = torch.cos(self.omega), torch.sin(self.omega)
            elif sin_omega is not None and cos_omega is not None:
                self.cos_omega, self.sin_omega = _tensor(cos_omega, dev), _tensor(sin_omega, dev)
                self.omega = torch.atan2(self.sin_omega, self.cos_omega)
            else:
                raise ValueError("both e and omega must be provided")
            # eccentric anomaly of mid-transit (f = pi/2 - omega) and its mean anomaly (keplerian.py:205-210)
            E0 = 2.0 * torch.atan2(torch.sqrt(1.0 - self.ecc) * self.cos_omega, torch.sqrt(1.0 + self.ecc) * (1.0 + self.sin_omega))
            self.M0 = E0 - self.ecc * torch.sin(E0)
            one_minus_e2 = 1.0 - self.ecc ** 2
            self.K0 = self.K0 / torch.sqrt(one_minus_e2)
            shape_factor = (1.0 + self.ecc * self.sin_omega) / one_minus_e2

        self.dcosidb = self.jacobians["b"]["cos_incl"] = shape_factor * self.r_star / self.a
        given = [x is not None for x in (incl, b, duration)]
        if sum(given) > 1:
            raise ValueError("only one of 'incl', 'b', and 'duration' can be given")
        if b is not None:
            self.b = _tensor(b, dev)
            self.cos_incl = self.dcosidb * self.b
            self.incl = torch.arccos(self.cos_incl)
        elif incl is not None:
            self.incl = _tensor(incl, dev)
            self.cos_incl = torch.cos(self.incl)
            self.b = self.cos_incl / self.dcosidb
        elif duration is not None:  # eccentric orbit with a duration (keplerian.py:237-260)
            self.duration = _tensor(duration, dev)
            c2 = torch.sin(np.pi * self.duration * shape_factor / self.period) ** 2
            aor = self.a_planet / self.r_star
            esw = self.ecc * self.sin_omega
            e2 = self.ecc ** 2
            self.b = torch.sqrt((aor ** 2 * c2 - 1.0) / (c2 * esw ** 2 + 2.0 * c2 * esw + c2 - e2 ** 2 + 2.0 * e2 - 1.0)) * (1.0 - e2)
            self.cos_incl = self.dcosidb * self.b
            self.incl = torch.arccos(self.cos_incl)
        else:
            zero = torch.zeros_like(self.a)
            self.incl, self.cos_incl, self.b = 0.5 * np.pi + zero, zero, zero
        self.sin_incl = torch.sin(self.incl)

        if t0 is not None and t_periastron is not None:
            raise ValueError("you can't define both t0 and t_periastron")
        if t0 is None and t_periastron is None:
            t0 = torch.zeros_like(self.period)
        if t0 is None:
            self.t_periastron = _tensor(t_periastron, dev)
            self.t0 = self.t_periastron + self.M0 / self.n
        else:
            self.t0 = _tensor(t0, dev)
            self.t_periastron = self.t0 - self.M0 / self.n
        self.tref = self.t_periastron - self.t0

    # ------------------------------------------------------------------ element table
    def elements(self, r=None, texp=None, use_in_transit=True, K=None, light_delay=False):
        """[P, 16] element blocks for the fused kernels (include/xo_b200.h).  ``r`` is the planet
        radius in R_sun (same convention as ``get_light_curve``).  ``light_delay``: slot XO_EL_RSTAR carries R_star (the
        delay is dimensional) and the window is the ordinary one widened by the largest possible delay -- the reference
        evaluates every cadence in this mode (keplerian.py:720-723), which gives the same light curve because the flux is
        exactly 0 outside the contact points."""
        ror_in = None if r is None else _tensor(r, self.device).reshape(-1) / self.r_star
        z = torch.zeros_like(self.a + self.period + self.t0).reshape(-1)
        if ror_in is not None:
            z = z + torch.zeros_like(ror_in)
        ecc = z if self.ecc is None else self.ecc + z
        cw = z + 1.0 if self.ecc is None else self.cos_omega + z
        sw = z if self.ecc is None else self.sin_omega + z
        ror = z if ror_in is None else ror_in + z
        cosi, sini = self.cos_incl + z, self.sin_incl + z
        period, n, a = self.period + z, self.n + z, self.a + z
        if use_in_transit:
            if self.ecc is None:
                ts, te = _el.circular_window(a, period, self.b + z, ror, self.r_star, sini, texp)
            else:
                ts, te = _el.transit_window(a, ecc, cw, sw, cosi, sini, self.M0 + z, n, period, self.r_star * (1.0 + ror), texp)
        else:
            ts, te = torch.full_like(z, -float("inf")), torch.full_like(z, float("inf"))
        Kc = z if K is None else _tensor(K, self.device) + z
        rs = z
        if light_delay:
            rs = (self.r_star + z).detach()
            margin = (2.0 * a * (1.0 + ecc) / _el.C_LIGHT).detach()
            ts, te = ts - margin, te + margin
        cols = [self.t0 + z, self.tref + z, n, ecc, cw, sw, cosi, sini, a / self.r_star, ror, period, ts, te, Kc, rs, z]
        return torch.stack(cols, dim=-1).reshape(-1, 16)

    # ------------------------------------------------------------------ anomalies
    def _warp_times(self, t, _pad=True):
        t = _tensor(t, self.device)
        return (t.unsqueeze(-1) if _pad else t) - self.t0

    def _get_true_anomaly(self, t, _pad=True):
        M = (self._warp_times(t, _pad=_pad) - self.tref) * self.n  # order of operations as keplerian.py:324-330
        if self.ecc is None:
            return torch.sin(M), torch.cos(M)
        return ops.kepler(M, self.ecc + torch.zeros_like(M))

    def _to_observer(self, x, y):
        """orbital plane -> observer frame: rotate by omega about z, -incl about x, Omega about z
        (keplerian.py:283-322)."""
        if self.ecc is not None:
            x, y = self.cos_omega * x - self.sin_omega * y, self.sin_omega * x + self.cos_omega * y
        X, Y, Z = x, self.cos_incl * y, -self.sin_incl * y
        if self.Omega is not None:
            X, Y = self.cos_Omega * X - self.sin_Omega * Y, self.sin_Omega * X + self.cos_Omega * Y
        return X, Y, Z

    _rotate_vector = _to_observer

    def _radius(self, a, cosf):
        return a if self.ecc is None else a * (1.0 - self.ecc ** 2) / (1.0 + self.ecc * cosf)

    def _get_position(self, a, t, parallax=None, light_delay=False, _pad=True):
        if light_delay:
            return self._get_retarded_position(a, t, parallax=parallax, _pad=_pad)
        sinf, cosf = self._get_true_anomaly(t, _pad=_pad)
        r = self._radius(a, cosf)
        if parallax is not None:
            r = r * parallax * _el.AU_PER_R_SUN
        return self._to_observer(r * cosf, r * sinf)

    def _get_retarded_position(self, a, t, parallax=None, z0=0.0, _pad=True):
        """Position at the retarded time t - delay, delay from the quadratic light-travel
        equation along the line of sight (keplerian.py:411-470)."""
        sinf, cosf = self._get_true_anomaly(t, _pad=_pad)
        r = self._radius(a, cosf)
        if self.ecc is None:
            vz = self.n * a * self.sin_incl * cosf
        else:
            vz = self.n * a / torch.sqrt(1.0 - self.ecc ** 2) * self.sin_incl * (
                self.ecc * self.cos_omega + self.cos_omega * cosf - self.sin_omega * sinf)
        _, _, z = self._to_observer(r * cosf, r * sinf)
        az = -(self.n ** 2) * (a / r) ** 3 * z
        c = _el.C_LIGHT
        beta = 1.0 + vz / c
        quad = (c / az) * (beta - torch.sqrt(beta * beta - 2.0 * az * (z0 - z) / c ** 2))
        delay = torch.where(az.abs() < 1.0e-10, (z0 - z) / (c + vz), quad)
        tt = _tensor(t, self.device)
        new_t = (tt.unsqueeze(-1) if _pad else tt) - delay
        return self._get_position(a, new_t, parallax, _pad=False)

    @staticmethod
    def _squeeze(xs):
        return tuple(torch.squeeze(x) for x in xs)

    def get_planet_position(self, t, parallax=None, light_delay=False):
        return self._squeeze(self._get_position(self.a_planet, t, parallax, light_delay=light_delay))

    def get_star_position(self, t, parallax=None, light_delay=False):
        return self._squeeze(self._get_position(self.a_star, t, parallax, light_delay=light_delay))

    def get_relative_position(self, t, parallax=None, light_delay=False):
        """Planet positions relative to the star, each planet treated independently
        (keplerian.py:517-542); Z > 0 is towards the observer."""
        return self._squeeze(self._get_position(-self.a, t, parallax, light_delay=light_delay))

    def get_relative_angles(self, t, parallax=None, light_delay=False):
        X, Y, _ = self._get_position(-self.a, t, parallax, light_delay=light_delay)
        return torch.squeeze(torch.sqrt(X * X + Y * Y)), torch.squeeze(torch.atan2(Y, X))

    # ------------------------------------------------------------------ velocities
    def _get_velocity(self, m, t):
        sinf, cosf = self._get_true_anomaly(t)
        K = self.K0 * m
        return self._to_observer(-K * sinf, K * (cosf if self.ecc is None else cosf + self.ecc))

    def get_planet_velocity(self, t):
        return self._squeeze(self._get_velocity(-self.m_star, t))

    def get_star_velocity(self, t):
        return self._squeeze(self._get_velocity(self.m_planet, t))

    def get_relative_velocity(self, t):
        return self._squeeze(self._get_velocity(-self.m_total, t))

    def get_radial_velocity(self, t, K=None, output_units=None):
        """Stellar reflex radial velocity, positive = redshift (keplerian.py:633-677).  With ``K``
        the semi-amplitude form K (cos(omega + f) + e cos omega) is used; otherwise the full orbit
        in m/s (``output_units``: a float factor from R_sun/day, default m/s)."""
        if K is not None:
            sinf, cosf = self._get_true_anomaly(t)
            K = _tensor(K, self.device)
            if self.ecc is None:
                return torch.squeeze(K * cosf)
            return torch.squeeze(K * (self.cos_omega * cosf - self.sin_omega * sinf + self.ecc * self.cos_omega))
        conv = M_PER_S_PER_RSUN_DAY if output_units is None else float(output_units)
        return -conv * self.get_star_velocity(t)[2]

    # ------------------------------------------------------------------ accelerations
    def _get_acceleration(self, a, m, t):
        sinf, cosf = self._get_true_anomaly(t)
        K = self.K0 * m
        if self.ecc is None:
            f = -(K ** 2) / a
        else:
            f = K ** 2 * (self.ecc * cosf + 1.0) ** 2 / (a * (self.ecc ** 2 - 1.0))
        return self._to_observer(f * cosf, f * sinf)

    def get_planet_acceleration(self, t):
        return self._squeeze(self._get_acceleration(self.a_planet, -self.m_star, t))

    def get_star_acceleration(self, t):
        return self._squeeze(self._get_acceleration(self.a_star, self.m_planet, t))

    def get_relative_acceleration(self, t):
        return self._squeeze(self._get_acceleration(-self.a, -self.m_total, t))

    # ------------------------------------------------------------------ in-transit selection
    def transit_window(self, r=0.0, texp=None):
        """(t_start, t_end, flag) per planet on dt = mod(t - t0 + P/2, P) - P/2 (keplerian.py:729-767)."""
        z = torch.zeros_like(self.a + self.period)
        ror = _tensor(r, self.device) / self.r_star + z
        if texp is not None and isinstance(texp, torch.Tensor) and texp.dim() > 0:
            texp = texp.max()  # per-cadence exposure times: widen by the longest one (a superset)
        if self.ecc is None:
            ts, te = _el.circular_window(self.a + z, self.period + z, self.b + z, ror, self.r_star, self.sin_incl + z, texp)
        else:
            ts, te = _el.transit_window(self.a + z, self.ecc + z, self.cos_omega + z, self.sin_omega + z, self.cos_incl + z,
                                        self.sin_incl + z, self.M0 + z, self.n + z, self.period + z, self.r_star * (1.0 + ror), texp)
        return ts, te, torch.isinf(ts)

    def in_transit(self, t, r=0.0, texp=None, light_delay=False):
        """Indices of the time stamps that may be in transit for at least one planet; every index
        when the contact solver cannot bracket a transit (keplerian.py:708-777)."""
        if light_delay:
            raise NotImplementedError("Light travel time delay not yet implemented for `in_transit`")
        t = _tensor(t, self.device)
        ts, te, failed = self.transit_window(r=r, texp=texp)
        if bool(failed.any()):
            return torch.arange(t.shape[0], device=self.device)
        half = 0.5 * self.period
        dt = torch.remainder(self._warp_times(t) + half, self.period) - half
        mask = ((dt >= ts) & (dt <= te)).reshape(t.shape[0], -1).any(dim=-1)
        return torch.nonzero(mask).reshape(-1)

    # ------------------------------------------------------------------ star <-> planet swap
    def _flip(self, r_planet, model=None):
        """The orbit of the *star* about the planet (used for secondary eclipses, keplerian.py:779-804)."""
        common = dict(period=self.period, incl=self.incl, Omega=self.Omega, m_star=self.m_planet, m_planet=self.m_star,
                      r_star=r_planet, device=self.device)
        if self.ecc is None:
            return type(self)(t_periastron=self.t_periastron + 0.5 * self.period, **common)
        return type(self)(t_periastron=self.t_periastron, ecc=self.ecc, omega=self.omega - np.pi, **common)
