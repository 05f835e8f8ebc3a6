"""ctypes binding of libxo_b200.so (the C ABI declared in include/xo_b200.h).

There is no CPU arithmetic path in this package: if the CUDA library is missing
or no CUDA device is present, every op raises -- loudly -- instead of falling
back (the oracle under oracle/ is test infrastructure and is never imported here).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# XO_B200_LIB: an alternative build of the same library (tuning experiments); default = the in-tree build
LIB_PATH = os.environ.get("XO_B200_LIB") or os.path.join(_HERE, "libxo_b200.so")

NEL = 16
NGRAD = 10
NGRAD_RV = 7
NPAR = 8
NSTAR = 4
MOMENTS_LEN = 1024  # XO_MOMENTS_LEN: doubles in the buffer xo_transit_data_moments writes (8 results + scratch)
MAX_STENCIL = 33
MAX_PLANETS = 8

# element-block indices (include/xo_b200.h)
EL_T0, EL_TREF, EL_N, EL_ECC, EL_COSW, EL_SINW, EL_COSI, EL_SINI, EL_AOR, EL_ROR, EL_PERIOD, EL_TSTART, EL_TEND, EL_KRV, EL_RSTAR = range(15)


class XoError(RuntimeError):
    pass


_lib = None

_vp, _i64, _i32, _dbl, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_size_t

_SIGNATURES = {
    "xo_version": (C.c_char_p, []),
    "xo_abi_version": (_i32, []),
    "xo_device_count": (_i32, []),
    "xo_strerror": (C.c_char_p, [_i32]),
    "xo_kepler_fwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "xo_kepler_fwd_bcast": (_i32, [_vp, _vp, _i64, _vp, _vp, _i64, _vp]),
    "xo_kepler_vjp": (_i32, [_vp] * 7 + [_i64, _vp]),
    "xo_qsv_fwd": (_i32, [_vp] * 5 + [_i64, _vp]),
    "xo_qsv_vjp": (_i32, [_vp] * 5 + [_i64, _vp]),
    "xo_contact_points": (_i32, [_vp] * 10 + [_i64, _vp]),
    "xo_transit_lc_fwd": (_i32, [_vp, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _vp, _vp]),
    "xo_transit_lc_vjp": (_i32, [_vp, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "xo_transit_lc_fwd_ex": (_i32, [_vp, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "xo_transit_lc_vjp_ex": (_i32, [_vp, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                                    _vp]),
    "xo_transit_logp_grad_ttv": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _dbl, _vp, _vp, _vp, _i32, _vp, _vp,
                                        _vp, _vp, _vp, _vp, _sz, _vp]),
    "xo_transit_logp_grad_workspace": (_sz, [_i64, _i64, _i64]),
    "xo_transit_data_moments": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "xo_transit_moments_len": (_sz, [_i64]),
    "xo_transit_data_index": (_i32, [_vp, _i64, _vp, _sz, _vp]),
    "xo_transit_data_moments_index": (_i32, [_vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "xo_transit_logp_grad": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _dbl, _vp, _vp, _vp, _vp, _vp,
                                    _vp, _sz, _vp]),
    "xo_transit_model_workspace": (_sz, [_i64, _i64, _i64]),
    "xo_transit_model_logp_grad": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _dbl, _i32, _dbl, _vp,
                                          _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "xo_elements_from_params": (_i32, [_vp, _vp, _i64, _i64, _dbl, _i32, _vp, _vp, _vp]),
    "xo_orbit_eval": (_i32, [_vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "xo_rv_logp_grad_workspace": (_sz, [_i64, _i64, _i64]),
    "xo_rv_logp_grad": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _dbl, _vp, _vp, _vp, _vp, _sz, _vp]),
    "xo_celerite_sho_workspace": (_sz, [_i64, _i64]),
    "xo_celerite_sho_loglike": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _dbl, _vp, _vp, _sz, _vp]),
    "xo_celerite_sho_solve_workspace": (_sz, [_i64, _i64]),
    "xo_celerite_sho_solve": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _dbl, _vp, _vp, _vp, _sz, _vp]),
    "xo_celerite_sho_grad_workspace": (_sz, [_i64, _i64]),
    "xo_celerite_sho_grad": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _dbl, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "xo_transit_gp_workspace": (_sz, [_i64, _i64, _i64]),
    "xo_transit_gp_logp_grad": (_i32, [_vp, _vp, _vp, _i64] + [_vp] * 7 + [_i64, _i64, _vp, _vp, _i32, _dbl, _i32, _dbl] + [_vp] * 7 + [_sz, _vp]),
    "xo_transit_posterior_workspace": (_sz, [_i64, _i64, _i64]),
    "xo_transit_posterior_logp_grad": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _i32, _dbl, _i32, _dbl, _vp,
                                              _vp, _vp, _sz, _vp]),
    "xo_nuts_state_layout": (_sz, [_i64, _i32, _i32, _vp]),
    "xo_nuts_start": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "xo_nuts_leaf_pre": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "xo_nuts_leaf_post": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "xo_nuts_merge": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "xo_hmc_transition": (_i32, [_vp, _vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _i32, _dbl, _i32, _dbl, _vp, _vp, _vp, _vp, _vp,
                                 _i32, C.c_ulonglong, C.c_ulonglong, _vp, _vp, _vp, _sz, _vp]),
    "xo_kepler_host": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32]),
    "xo_qsv_host": (_i32, [_vp] * 5 + [_i64, _i32]),
    "xo_contact_points_host": (_i32, [_vp] * 10 + [_i64, _i32]),
    "xo_transit_lc_host": (_i32, [_vp, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _vp, _i32]),
    "xo_transit_logp_grad_host": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _dbl, _vp, _vp, _vp,
                                         _vp, _i32]),
    "xo_transit_model_logp_grad_host": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _dbl, _i32, _dbl,
                                               _vp, _vp, _vp, _vp, _i32]),
    "xo_diag_dfma_peak": (_i32, [_i32, C.POINTER(_dbl), C.POINTER(_dbl), _vp]),
    "xo_read_status": (_i32, [_vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), _vp]),
    "xo_diag_launch_count": (C.c_longlong, []),
    "xo_transit_data_create": (_i32, [_vp, _vp, _vp, _i64, _i32, C.POINTER(_vp)]),
    "xo_transit_data_destroy": (_i32, [_vp]),
    "xo_transit_model_logp_grad_resident": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _dbl, _i32, _dbl, _i32, _vp, _vp,
                                                   _vp, _vp]),
}

EXPORTS = tuple(sorted(_SIGNATURES))


def load():
    """Load the library (no GPU needed to load; needed to compute)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XoError(
            "exoplanet_b200: %s is missing -- build it with `python exoplanet_b200/csrc/build.py` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI drifted
        fn.restype = res
        fn.argtypes = args
    if lib.xo_abi_version() != 1:
        raise XoError("exoplanet_b200: ABI version mismatch")
    _lib = lib
    return lib


def require_cuda():
    lib = load()
    if lib.xo_device_count() < 1:
        raise XoError("exoplanet_b200: no CUDA device visible; this engine has no CPU path")
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().xo_strerror(int(rc)).decode()
        raise XoError("exoplanet_b200 %s failed: %s (status %d)" % (what, msg, rc))


def stream_ptr():
    """cudaStream_t of torch's current stream, as an integer for the C ABI."""
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
