"""CPU: the C-ABI library loads and exports every symbol include/xo_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "xo_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(xo_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    from exoplanet_b200.csrc import build as b

    path = b.build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export: " + n


def test_python_binding_covers_the_header():
    from exoplanet_b200 import _lib

    assert sorted(_lib.EXPORTS) == _declared()
    lib = _lib.load()
    assert lib.xo_abi_version() == 1
    assert b"sm_100a" in lib.xo_version()


def test_argument_errors_do_not_need_a_gpu():
    from exoplanet_b200 import _lib

    lib = _lib.load()
    assert lib.xo_kepler_fwd(None, None, None, None, -1, None) == 1  # XO_EINVAL
    assert lib.xo_transit_lc_fwd(None, 10, None, None, 1, 9, None, None, 1, None, None) == 2  # XO_ERANGE (P > 8)
    assert b"invalid" in lib.xo_strerror(1)
    # every family validates before it touches the device: negative sizes, null pointers, limits
    assert lib.xo_celerite_sho_loglike(None, None, None, 10, None, None, None, 1, 1e-5, None, None, 0, None) == 1
    assert lib.xo_celerite_sho_grad(None, None, None, -1, None, None, None, 1, 1e-5, None, None, None, None, None, 0, None) == 1
    assert lib.xo_celerite_sho_grad(None, None, None, 10, None, None, None, 0, 1e-5, None, None, None, None, None, 0, None) == 0  # no chains
    assert lib.xo_celerite_sho_grad(None, None, None, 10, None, None, None, 70000, 1e-5, None, None, None, None, None, 0, None) == 2
    assert lib.xo_celerite_sho_grad_workspace(0, 1) == 0 and lib.xo_celerite_sho_grad_workspace(1000, 2) > 0
    assert lib.xo_transit_data_moments(None, None, None, 0, None, None) == 1
    assert lib.xo_transit_model_logp_grad(None, None, None, 10, None, None, None, 0, 1, None, None, 1, 0.0, 1, 0.0, None, None, None,
                                          None, None, None, 0, None) == 0  # zero chains: nothing to do
    assert lib.xo_transit_model_logp_grad(None, None, None, 10, None, None, None, 4, 1, None, None, 1, 0.0, 1, 0.0, None, None, None,
                                          None, None, None, 0, None) == 1
    assert lib.xo_transit_model_workspace(1000, 4, 2) > 0 and lib.xo_transit_model_workspace(0, 4, 2) == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "exoplanet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "xo_oracle" not in src and "libxo_oracle" not in src, f


def test_ops_fail_loudly_without_cuda():
    import numpy as np
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from exoplanet_b200 import _lib, ops

    with pytest.raises(_lib.XoError):
        ops.kepler(np.zeros(4), np.zeros(4))
