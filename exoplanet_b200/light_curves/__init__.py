from .limb_dark import LimbDarkLightCurve  # noqa: F401
from .secondary_eclipse import SecondaryEclipseLightCurve  # noqa: F401
