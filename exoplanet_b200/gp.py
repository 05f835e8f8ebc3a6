"""celerite2-style SHO Gaussian-process log-likelihood as a parallel scan on the GPU
(BASELINE config 4).  celerite2 has no call site in the reference (setup.py:36 lists it
as an optional extra); the contract is ``celerite2.GaussianProcess(terms.SHOTerm(S0=,
w0=, Q=), mean=0).compute(t, diag=diag); gp.log_likelihood(y)``.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

__all__ = ["sho_loglike", "sho_solve", "sho_loglike_autograd", "sho_params_from_sigma_rho"]


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


def sho_solve(t, y, diag, S0, w0, Q, eps=1e-5, workspace=None):
    """(loglike[B], alpha[B, N]) with alpha = K^-1 y: the forward scan followed by the disturbance
    smoother run as a reverse affine scan (xo_celerite_sho_solve).  d loglike / d y = -alpha."""
    lib = _lib.require_cuda()
    dev = t.device
    y2 = (y if y.dim() == 2 else y[None, :]).contiguous()
    B, N = y2.shape

    def vec(x):
        x = torch.as_tensor(x, dtype=torch.float64, device=dev).detach()
        return x.expand(B).contiguous() if x.dim() == 0 else x.contiguous()

    S0, w0, Q = vec(S0), vec(w0), vec(Q)
    t = t.contiguous()
    diag = diag.contiguous()
    need = lib.xo_celerite_sho_solve_workspace(N, B)
    if workspace is None or workspace.numel() * 8 < need:
        workspace = torch.empty(need // 8 + 1, dtype=torch.float64, device=dev)
    ll = torch.empty(B, dtype=torch.float64, device=dev)
    alpha = torch.empty((B, N), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.xo_celerite_sho_solve(_ptr(t), _ptr(y2), _ptr(diag), N, _ptr(S0), _ptr(w0), _ptr(Q), B, float(eps), _ptr(ll),
                                             _ptr(alpha), _ptr(workspace), workspace.numel() * 8, _lib.stream_ptr()),
                   "celerite_sho_solve")
    return (ll, alpha) if y.dim() == 2 else (ll[0], alpha[0])


class _SHOLogLike(torch.autograd.Function):
    """log-likelihood[B] with gradients: exact w.r.t. y (= -K^-1 y, one reverse scan); w.r.t.
    (S0, w0, Q) by central differences of the forward scan, the 6 perturbed evaluations batched
    into one launch (relative step 1e-6 -- NOT an analytic VJP)."""

    REL_STEP = 1e-6

    @staticmethod
    def forward(ctx, y, S0, w0, Q, t, diag, eps):
        ll, alpha = sho_solve(t, y, diag, S0, w0, Q, eps)
        ctx.save_for_backward(y, S0, w0, Q, t, diag, alpha)
        ctx.eps = eps
        return ll

    @staticmethod
    def backward(ctx, g):
        y, S0, w0, Q, t, diag, alpha = ctx.saved_tensors
        B = y.shape[0]
        gy = -g[:, None] * alpha if ctx.needs_input_grad[0] else None
        gth = [None, None, None]
        if any(ctx.needs_input_grad[1:4]):
            th = [S0, w0, Q]
            rows = []
            for i in range(3):
                for sgn in (1.0, -1.0):
                    p = [x.clone() for x in th]
                    p[i] = p[i] * (1.0 + sgn * _SHOLogLike.REL_STEP)
                    rows.append(p)
            S = torch.cat([r[0] for r in rows])
            W = torch.cat([r[1] for r in rows])
            Qq = torch.cat([r[2] for r in rows])
            ll = sho_loglike(t, y.repeat(6, 1), diag, S, W, Qq, ctx.eps).reshape(6, B)
            for i in range(3):
                gth[i] = g * (ll[2 * i] - ll[2 * i + 1]) / (2.0 * _SHOLogLike.REL_STEP * th[i])
        return gy, gth[0], gth[1], gth[2], None, None, None


def sho_loglike_autograd(t, y, diag, S0, w0, Q, eps=1e-5):
    """Differentiable SHO GP log-likelihood: y[B, N] (residuals), S0/w0/Q [B] tensors."""
    y2 = y if y.dim() == 2 else y[None, :]
    B = y2.shape[0]
    dev = t.device

    def vec(x):
        x = torch.as_tensor(x, dtype=torch.float64, device=dev)
        return x.expand(B) if x.dim() == 0 else x

    out = _SHOLogLike.apply(y2, vec(S0), vec(w0), vec(Q), t, diag, float(eps))
    return out if y.dim() == 2 else out[0]
