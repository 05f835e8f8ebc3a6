"""Install mode (i) of SURVEY.md 8(b): a package importable as ``exoplanet_core`` whose ``pymc.ops`` module is the
B200 engine, so that an UNMODIFIED checkout of exoplanet-dev/exoplanet runs on it: the reference binds its three
external ops with ``from exoplanet_core.pymc import ops as ops`` (/root/reference/src/exoplanet/compat.py:56).

    PYTHONPATH=/path/to/repo/shim:/path/to/repo  python -c "import exoplanet; ..."

Nothing else of exoplanet-core is provided (its JAX / numpy front ends are not on the reference's path)."""
__version__ = "0.3.0+b200"
