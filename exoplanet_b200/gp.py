"""celerite2-style SHO Gaussian-process log-likelihood as a parallel scan on the GPU
(BASELINE config 4).  celerite2 has no call site in the reference (setup.py:36 lists it
as an optional extra); the contract is ``celerite2.GaussianProcess(terms.SHOTerm(S0=,
w0=, Q=), mean=0).compute(t, diag=diag); gp.log_likelihood(y)``.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

__all__ = ["read_status", "sho_loglike", "sho_solve", "sho_grad", "sho_loglike_autograd", "sho_params_from_sigma_rho"]


def _ptr(x):
    return C.c_void_p(x.data_ptr()) if x is not None else None


def _gp_args(t, y, diag, S0, w0, Q):
    """validated (t[N], y2[B, N], diag[N], S0[B], w0[B], Q[B]): float64 tensors on t's CUDA device; hyper-parameters
    may be python scalars / 0-d tensors (one value for every chain) or [B] tensors"""
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.dim() == 1):
        raise ValueError("t must be a 1-d float64 CUDA tensor")
    dev, N = t.device, t.numel()
    for name, x in (("y", y), ("diag", diag)):
        if not (isinstance(x, torch.Tensor) and x.dtype == torch.float64 and x.device == dev):
            raise ValueError("%s must be a float64 tensor on %s" % (name, dev))
    if y.dim() not in (1, 2) or y.shape[-1] != N:
        raise ValueError("y must have shape [N] or [B, N] with N = %d, got %s" % (N, tuple(y.shape)))
    if tuple(diag.shape) != (N,):
        raise ValueError("diag must have shape [%d], got %s" % (N, tuple(diag.shape)))
    y2 = (y if y.dim() == 2 else y[None, :]).detach().contiguous()
    B = y2.shape[0]

    def vec(x, name):
        x = torch.as_tensor(x, dtype=torch.float64, device=dev).detach()
        if x.dim() == 0:
            return x.expand(B).contiguous()
        if tuple(x.shape) != (B,):
            raise ValueError("%s must be a scalar or have shape [%d], got %s" % (name, B, tuple(x.shape)))
        return x.contiguous()

    return t.detach().contiguous(), y2, diag.detach().contiguous(), vec(S0, "S0"), vec(w0, "w0"), vec(Q, "Q")


def sho_params_from_sigma_rho(sigma, rho, Q):
    """celerite2's alternative SHOTerm parameterisation: sigma^2 = S0 w0 Q, rho = 2 pi / w0."""
    w0 = 2.0 * np.pi / rho
    S0 = sigma ** 2 / (w0 * Q)
    return S0, w0, Q


def read_status(workspace):
    """(bad_lanes, wait_failures) of the last celerite call that used ``workspace`` (xo_read_status); synchronises
    the current stream."""
    lib = _lib.require_cuda()
    bad, wf = C.c_ulonglong(0), C.c_ulonglong(0)
    with torch.cuda.device(workspace.device):
        _lib.check(lib.xo_read_status(_ptr(workspace), C.byref(bad), C.byref(wf), _lib.stream_ptr()), "read_status")
    return bad.value, wf.value


def sho_loglike(t, y, diag, S0, w0, Q, eps=1e-5, workspace=None):
    """log-likelihood[B] of residuals y[B, N] (or [N]) under k_SHO + diag, zero mean.

    t[N] sorted, diag[N] = sigma^2 (+ jitter), S0/w0/Q scalars or [B] tensors."""
    lib = _lib.require_cuda()
    t, y2, diag, S0, w0, Q = _gp_args(t, y, diag, S0, w0, Q)
    dev = t.device
    B, N = y2.shape
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
    t, y2, diag, S0, w0, Q = _gp_args(t, y, diag, S0, w0, Q)
    dev = t.device
    B, N = y2.shape
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


def sho_grad(t, y, diag, S0, w0, Q, eps=1e-5, want_y=True, want_diag=True, workspace=None):
    """(loglike[B], ghyper[B, 3], gy[B, N] | None, gdiag[B, N] | None): the log-likelihood and its
    exact reverse-mode gradient w.r.t. (S0, w0, Q), y and diag (xo_celerite_sho_grad): the forward
    scan, then the adjoint of the Kalman recursion as a reverse affine scan."""
    lib = _lib.require_cuda()
    t, y2, diag, S0, w0, Q = _gp_args(t, y, diag, S0, w0, Q)
    dev = t.device
    B, N = y2.shape
    need = lib.xo_celerite_sho_grad_workspace(N, B)
    if workspace is None or workspace.numel() * 8 < need:
        workspace = torch.empty(need // 8 + 1, dtype=torch.float64, device=dev)
    ll = torch.empty(B, dtype=torch.float64, device=dev)
    gh = torch.empty((B, 3), dtype=torch.float64, device=dev)
    gy = torch.empty((B, N), dtype=torch.float64, device=dev) if want_y else None
    gd = torch.empty((B, N), dtype=torch.float64, device=dev) if want_diag else None
    with torch.cuda.device(dev):
        _lib.check(lib.xo_celerite_sho_grad(_ptr(t), _ptr(y2), _ptr(diag), N, _ptr(S0), _ptr(w0), _ptr(Q), B, float(eps), _ptr(ll),
                                            _ptr(gh), _ptr(gy), _ptr(gd), _ptr(workspace), workspace.numel() * 8,
                                            _lib.stream_ptr()), "celerite_sho_grad")
    if y.dim() == 2:
        return ll, gh, gy, gd
    return ll[0], gh[0], (gy[0] if want_y else None), (gd[0] if want_diag else None)


class _SHOLogLike(torch.autograd.Function):
    """log-likelihood[B] with its exact gradient w.r.t. y, (S0, w0, Q) and diag: one forward scan and
    one reverse (adjoint) scan, both in forward() -- the reverse scan restarts from the states the
    forward one records, so the pair is cheaper than keeping them for backward()."""

    @staticmethod
    def forward(ctx, y, S0, w0, Q, t, diag, eps):
        need = ctx.needs_input_grad
        if not any(need):
            return sho_loglike(t, y, diag, S0, w0, Q, eps)
        ll, gh, gy, gd = sho_grad(t, y, diag, S0, w0, Q, eps, want_y=need[0], want_diag=need[5])
        ctx.save_for_backward(gh, gy, gd)
        return ll

    @staticmethod
    def backward(ctx, g):
        gh, gy, gd = ctx.saved_tensors
        need = ctx.needs_input_grad
        out_y = g[:, None] * gy if need[0] else None
        out_d = (g[:, None] * gd).sum(0) if need[5] else None
        return (out_y, g * gh[:, 0] if need[1] else None, g * gh[:, 1] if need[2] else None, g * gh[:, 2] if need[3] else None, None,
                out_d, None)


def sho_loglike_autograd(t, y, diag, S0, w0, Q, eps=1e-5):
    """Differentiable SHO GP log-likelihood: y[B, N] (residuals), S0/w0/Q [B] tensors, diag[N];
    exact gradients w.r.t. all of them."""
    y2 = y if y.dim() == 2 else y[None, :]
    B = y2.shape[0]
    dev = t.device

    def vec(x):
        x = torch.as_tensor(x, dtype=torch.float64, device=dev)
        return x.expand(B) if x.dim() == 0 else x

    out = _SHOLogLike.apply(y2, vec(S0), vec(w0), vec(Q), t, diag, float(eps))
    return out if y.dim() == 2 else out[0]
