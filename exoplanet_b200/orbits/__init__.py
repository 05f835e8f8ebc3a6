from .keplerian import KeplerianOrbit, get_aor_from_transit_duration, get_true_anomaly  # noqa: F401
from .simple import SimpleTransitOrbit  # noqa: F401
from .ttv import TTVOrbit, compute_expected_transit_times  # noqa: F401
