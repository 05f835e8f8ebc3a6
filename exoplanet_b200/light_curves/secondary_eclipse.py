"""``SecondaryEclipseLightCurve``: primary transit + secondary eclipse, a pure composition of
two ``LimbDarkLightCurve`` evaluations on the orbit and its star<->planet flip (mirror of
/root/reference/src/exoplanet/light_curves/secondary_eclipse.py:7-70)."""
from ..orbits.keplerian import _tensor
from .limb_dark import LimbDarkLightCurve

__all__ = ["SecondaryEclipseLightCurve"]


class SecondaryEclipseLightCurve:
    def __init__(self, u_primary, u_secondary, surface_brightness_ratio, model=None):
        self.primary = LimbDarkLightCurve(u_primary[0], u_primary[1])
        self.secondary = LimbDarkLightCurve(u_secondary[0], u_secondary[1])
        self.surface_brightness_ratio = _tensor(surface_brightness_ratio, self.primary.u1.device)

    def get_light_curve(self, orbit=None, r=None, t=None, **kwargs):
        r = _tensor(r, self.primary.u1.device)
        transit = self.primary.get_light_curve(orbit=orbit, r=r, t=t, **kwargs)
        eclipse = self.secondary.get_light_curve(orbit=orbit._flip(r), r=orbit.r_star, t=t, **kwargs)
        flux_ratio = self.surface_brightness_ratio * (r / orbit.r_star) ** 2
        return (transit + flux_ratio * eclipse) / (1.0 + flux_ratio)
