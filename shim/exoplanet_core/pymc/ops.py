"""``exoplanet_core.pymc.ops`` served by the B200 engine: the three callables the reference touches --
``ops.kepler`` (/root/reference/src/exoplanet/orbits/keplerian.py:333, :818), ``ops.quad_solution_vector``
(light_curves/limb_dark.py:24) and ``ops.contact_points`` (orbits/keplerian.py:744-753) -- as PyTensor Ops whose
``perform`` calls the C ABI of libxo_b200.so (exoplanet_b200/pytensor_ops.py).  Needs PyTensor and a CUDA device;
there is no CPU fallback."""
from exoplanet_b200.pytensor_ops import contact_points, kepler, quad_solution_vector  # noqa: F401

__all__ = ["kepler", "quad_solution_vector", "contact_points"]
