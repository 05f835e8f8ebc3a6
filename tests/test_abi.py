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
