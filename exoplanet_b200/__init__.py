"""exoplanet_b200 -- a Blackwell (sm_100a) transit-model evaluation engine behind the
operator boundary of exoplanet-dev/exoplanet (``exoplanet.compat.ops``).

Only the hot path lives here: the three ops (``ops``), the fused light-curve /
log-likelihood kernels (``fused``), the celerite SHO scan (``gp``) and the host-side
mirror of the reference's ``KeplerianOrbit`` / ``LimbDarkLightCurve`` interface.
Everything computes on the GPU through include/xo_b200.h; there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401

__version__ = "0.1.0"
