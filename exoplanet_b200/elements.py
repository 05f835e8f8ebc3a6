"""The O(P) prologue: orbital parameters -> element block, in differentiable torch.

What ``KeplerianOrbit.__init__`` derives per planet
(/root/reference/src/exoplanet/orbits/keplerian.py:133-281) plus the ``in_transit``
window (:729-767), packed as the ``XO_EL_*`` block the fused kernels read
(include/xo_b200.h).  The window comes from the GPU ``contact_points`` op and carries
no gradient (it only gates which cadences are evaluated, as in the reference).
"""
import numpy as np
import torch

from . import _lib, ops

G_GRAV = 2942.2062175044193  # R_sun^3 / M_sun / day^2   (orbits/constants.py:32)
GCC_PER_SUN = 5.905271918964842  # (orbits/constants.py:33)
AU_PER_R_SUN = 0.00465046726096215
C_LIGHT = 37231.66360672704  # R_sun / day


def _t(x, like):
    return torch.as_tensor(x, dtype=torch.float64, device=like.device)


def transit_window(a, ecc, cos_omega, sin_omega, cos_incl, sin_incl, M0, n, period, L, texp=None):
    """(t_start, t_end) on dt = mod(t - t0 + P/2, P) - P/2  (keplerian.py:744-767).
    Where the contact solver raises its flag the window is (-inf, +inf): every cadence
    is evaluated, which is the reference's ``ifelse`` fallback (:771-775)."""
    with torch.no_grad():
        Ml, Mr, flag = ops.contact_points(a, ecc, cos_omega, sin_omega, cos_incl, sin_incl, L)
        hp = 0.5 * period
        ts = torch.remainder((Ml - M0) / n + hp, period) - hp
        te = torch.remainder((Mr - M0) / n + hp, period) - hp
        ts = torch.where(ts > 0, ts - period, ts)
        te = torch.where(te < 0, te + period, te)
        if texp is not None:
            ts = ts - 0.5 * texp
            te = te + 0.5 * texp
        inf = torch.full_like(ts, float("inf"))
        ts = torch.where(flag != 0, -inf, ts)
        te = torch.where(flag != 0, inf, te)
    return ts, te


def circular_window(a, period, b, ror, r_star, sin_incl, texp=None):
    """Closed form for circular orbits, Winn (2010) eq. 14 (keplerian.py:733-741)."""
    with torch.no_grad():
        arg = (1.0 + ror) ** 2 - b ** 2
        hdur = 0.5 * period * torch.arcsin(r_star / (a * sin_incl) * torch.sqrt(arg)) / np.pi
        hdur = torch.where(torch.isfinite(hdur), hdur, torch.full_like(hdur, float("inf")))
        if texp is not None:
            hdur = hdur + 0.5 * texp
    return -hdur, hdur


def elements_from_params(period, t0, ecc, omega, b, ror, r_star=1.0, m_star=1.0, m_planet=0.0, K=0.0, texp=None,
                         use_in_transit=True):
    """Differentiable element blocks [..., 16] from (period, t0, ecc, omega, b, ror[, ...]).

    All inputs broadcast; they must live on one CUDA device.  ``ror`` is r_planet/R_star;
    ``K`` the RV semi-amplitude used by the RV kernel."""
    period = torch.as_tensor(period, dtype=torch.float64)
    dev = period.device
    t0, ecc, omega, b, ror, r_star, m_star, m_planet, K = [
        torch.as_tensor(x, dtype=torch.float64, device=dev) for x in (t0, ecc, omega, b, ror, r_star, m_star, m_planet, K)]
    period, t0, ecc, omega, b, ror, m_planet, K = torch.broadcast_tensors(period, t0, ecc, omega, b, ror, m_planet, K)
    a = (G_GRAV * (m_star + m_planet) * period ** 2 / (4 * np.pi ** 2)) ** (1.0 / 3)  # keplerian.py:924-927
    n = 2 * np.pi / period
    cw, sw = torch.cos(omega), torch.sin(omega)
    E0 = 2 * torch.atan2(torch.sqrt(1 - ecc) * cw, torch.sqrt(1 + ecc) * (1 + sw))  # :205-208
    M0 = E0 - ecc * torch.sin(E0)
    incl_factor = (1 + ecc * sw) / (1 - ecc ** 2)  # :214
    cosi = incl_factor * r_star / a * b  # :217-228
    sini = torch.sqrt((1 - cosi) * (1 + cosi))
    tref = -M0 / n  # t_periastron - t0   (:272-279)
    if use_in_transit:
        ts, te = transit_window(a, ecc, cw, sw, cosi, sini, M0, n, period, r_star * (1 + ror), texp)
    else:
        ts = torch.full_like(period, -float("inf"))
        te = torch.full_like(period, float("inf"))
    z = torch.zeros_like(period)
    cols = [t0, tref, n, ecc, cw, sw, cosi, sini, a / r_star, ror, period, ts, te, K, z, z]
    return torch.stack(cols, dim=-1)
