"""celerite2-style SHO Gaussian-process log-likelihood as a parallel scan on the GPU
(BASELINE config 4).  celerite2 has no call site in the reference (setup.py:36 lists it
as an optional extra); the contract is ``celerite2.GaussianProcess(terms.SHOTerm(S0=,
w0=, Q=), mean=0).compute(t, diag=diag); gp.log_likelihood(y)``.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

__all__ = ["sho_loglike", "sho_params_from_sigma_rho"]


def _ptr(x):
    return C.c_void_p(x.data_ptr()) if x is not None else None


def sho_params_from_sigma_rho(sigma, rho, Q):
    """celerite2's alternative SHOTerm parameterisation: sigma^2 = S0 w0 Q, rho = 2 pi / w0."""
    w0 = 2.0 * np.pi / rho
    S0 = sigma ** 2 / (w0 * Q)
    return S0, w0, Q


def sho_loglike(t, y, diag, S0, w0, Q, eps=1e-5, workspace=None):
    """log-likelihood[B] of residuals y[B, N] (or [N]) under k_SHO + diag, zero mean.

    t[N] sorted, diag[N] = sigma^2 (+ jitter), S0/w0/Q scalars or [B] tensors."""
    lib = _lib.require_cuda()
    dev = t.device
    y2 = y if y.dim() == 2 else y[None, :]
    y2 = y2.contiguous()
    B, N = y2.shape

    def vec(x):
        x = torch.as_tensor(x, dtype=torch.float64, device=dev).detach()
        return x.expand(B).contiguous() if x.dim() == 0 else x.contiguous()

    S0, w0, Q = vec(S0), vec(w0), vec(Q)
    t = t.contiguous()
    diag = diag.contiguous()
    need = lib.xo_celerite_sho_workspace(N, B)
    if workspace is None or workspace.numel() * 8 < need:
        workspace = torch.empty(need // 8 + 1, dtype=torch.float64, device=dev)
    out = torch.empty(B, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.xo_celerite_sho_loglike(_ptr(t), _ptr(y2), _ptr(diag), N, _ptr(S0), _ptr(w0), _ptr(Q), B, float(eps),
                                               _ptr(out), _ptr(workspace), workspace.numel() * 8, _lib.stream_ptr()),
                   "celerite_sho_loglike")
    return out if y.dim() == 2 else out[0]
