"""``exoplanet_core.pymc``: only ``ops`` exists (what /root/reference/src/exoplanet/compat.py:56 imports)."""
from . import ops  # noqa: F401
