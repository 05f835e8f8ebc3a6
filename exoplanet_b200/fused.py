"""Fused (chains x planets x cadences) kernels over PyTorch CUDA tensors.

These swallow the PyTensor graph the reference builds between its three ops:
``LimbDarkLightCurve.get_light_curve`` (/root/reference/src/exoplanet/light_curves/
limb_dark.py:99-232) over ``KeplerianOrbit.get_relative_position``
(orbits/keplerian.py:517-542), the in-transit selection (:708-777), the Gaussian
likelihood and its whole reverse pass.  Inputs are the per-planet *element block*
(include/xo_b200.h, ``XO_EL_*``) so the O(P) prologue stays in ordinary autograd.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import MOMENTS_LEN, NEL, NGRAD, NGRAD_RV

__all__ = ["exposure_stencil", "limb_dark_coeffs", "TransitData", "light_curve", "transit_logp_grad", "TransitLogp",
           "orbit_eval", "rv_logp_grad", "RvLogp", "TransitModel", "TransitGPModel", "pack_params", "pack_star"]

LOG_2PI = 1.8378770664093454835606594728112


def _ptr(x):
    return C.c_void_p(x.data_ptr()) if x is not None else None


def _arg(x, shape, name, dev):
    """Validated kernel argument: a float64 CUDA tensor on ``dev`` with exactly ``shape`` (the kernels index raw
    pointers, so a wrong shape would read out of bounds or give silently wrong results); returned detached and
    contiguous.  No implicit broadcasting: expand explicitly (``x.expand(C, 3)``) where a shared value is meant."""
    if not isinstance(x, torch.Tensor):
        raise TypeError("%s must be a torch tensor, got %s" % (name, type(x).__name__))
    if x.dtype != torch.float64:
        raise ValueError("%s must be float64, got %s" % (name, x.dtype))
    want = torch.device(dev)
    if not x.is_cuda or (want.index is not None and x.device.index != want.index):
        raise ValueError("%s lives on %s, expected %s" % (name, x.device, want))
    if tuple(x.shape) != tuple(shape):
        raise ValueError("%s has shape %s, expected %s" % (name, tuple(x.shape), tuple(shape)))
    return x.detach().contiguous()


def _elem_dims(elem, name="elem"):
    if not isinstance(elem, torch.Tensor) or elem.dim() != 3 or elem.shape[-1] != NEL:
        raise ValueError("%s must have shape [C, P, %d], got %s" % (name, NEL, tuple(getattr(elem, "shape", ()))))
    if elem.shape[1] > _lib.MAX_PLANETS:
        raise ValueError("at most %d planets per chain" % _lib.MAX_PLANETS)
    return elem.shape[0], elem.shape[1]


def exposure_stencil(texp=None, oversample=7, order=0):
    """(dt_k, w_k) of limb_dark.py:173-198: centred Riemann (0), trapezoid (1), Simpson (2).
    Returns float64 numpy arrays; dt_k already multiplied by ``texp``."""
    if texp is None:
        return np.zeros(1), np.ones(1)
    n = int(oversample)
    n += 1 - n % 2
    w = np.ones(n)
    if order == 0:
        x = (np.arange(n) + 0.5) / n - 0.5
    elif order == 1:
        x = np.linspace(-0.5, 0.5, n)
        w[1:-1] = 2.0
    elif order == 2:
        x = np.linspace(-0.5, 0.5, n)
        w[1:-1:2] = 4.0
        w[2:-1:2] = 2.0
    else:
        raise ValueError("order must be <= 2")
    if n > _lib.MAX_STENCIL:
        raise ValueError("oversample above the compiled maximum (%d)" % _lib.MAX_STENCIL)
    return float(texp) * x, w / w.sum()


def limb_dark_coeffs(u1, u2):
    """Normalised Green's-basis vector c (limb_dark.py:11-18); differentiable torch."""
    c0 = 1.0 - u1 - 1.5 * u2
    c1 = u1 + 2.0 * u2
    c2 = -0.25 * u2
    norm = np.pi * (c0 + c1 / 1.5)
    return torch.stack([c0, c1, c2], dim=-1) / norm.unsqueeze(-1)


class TransitData:
    """t, y, 1/sigma^2 resident in HBM (shared by every chain) + the data-only constant
    of the Gaussian log-density: -sum(log sigma) - N/2 log(2 pi)."""

    def __init__(self, t, y=None, yerr=None, device=None):
        dev = torch.device("cuda" if device is None else device)
        self.t = torch.as_tensor(np.asarray(t, dtype=np.float64)).to(dev).contiguous()
        n = self.t.numel()
        self.y = None if y is None else torch.as_tensor(np.asarray(y, dtype=np.float64)).to(dev).contiguous()
        if yerr is None:
            self.ivar = None
            self.lognorm = 0.0
        else:
            sig = np.broadcast_to(np.asarray(yerr, dtype=np.float64), (n,))
            self.ivar = torch.as_tensor(np.ascontiguousarray(1.0 / sig ** 2)).to(dev)
            self.lognorm = float(-np.sum(np.log(sig)) - 0.5 * n * LOG_2PI)
        self.n = n
        self.device = dev
        self._moments = None

    def moments(self, index=True):
        """Device tensor (y_ref, sum ivar, sum ivar (y - y_ref), sum ivar (y - y_ref)^2, t_is_sorted, n_bins, 1/w, ...):
        the data-only part of the Gaussian sums (xo_transit_data_moments) followed, for sorted t, by the
        time -> cadence index the sweep uses instead of reading time stamps (xo_transit_data_moments_index); computed once."""
        if self._moments is None and self.y is not None and self.ivar is not None:
            lib = _lib.require_cuda()
            n = lib.xo_transit_moments_len(self.n) if index else MOMENTS_LEN
            m = torch.zeros(n, dtype=torch.float64, device=self.device)
            with torch.cuda.device(self.device):
                if index:  # moments and time index in one pass over the data
                    _lib.check(lib.xo_transit_data_moments_index(_ptr(self.t), _ptr(self.y), _ptr(self.ivar), self.n, _ptr(m), n,
                                                                 _lib.stream_ptr()), "moments+index")
                else:
                    _lib.check(lib.xo_transit_data_moments(_ptr(self.t), _ptr(self.y), _ptr(self.ivar), self.n, _ptr(m), _lib.stream_ptr()),
                               "moments")
            self._moments = m
        return self._moments


def _stencil_dev(stencil, dev):
    dtk, wk = stencil
    return (torch.as_tensor(np.ascontiguousarray(dtk, dtype=np.float64)).to(dev),
            torch.as_tensor(np.ascontiguousarray(wk, dtype=np.float64)).to(dev))


def light_curve(t, elem, cl, stencil=None):
    """lc[C, N, P] for element blocks elem[C, P, 16] and limb-darkening vectors cl[C, 3]."""
    lib = _lib.require_cuda()
    Cc, P = _elem_dims(elem)
    t = _arg(t, (t.numel(),), "t", t.device)
    elem = _arg(elem, (Cc, P, NEL), "elem", t.device)
    cl = _arg(cl, (Cc, 3), "cl", t.device)
    dtk, wk = _stencil_dev(stencil or exposure_stencil(), t.device)
    lc = torch.empty((Cc, t.numel(), P), dtype=torch.float64, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(lib.xo_transit_lc_fwd(_ptr(t), t.numel(), _ptr(elem), _ptr(cl), Cc, P, _ptr(dtk), _ptr(wk), dtk.numel(),
                                         _ptr(lc), _lib.stream_ptr()), "transit_lc_fwd")
    return lc


def transit_logp_grad(data, elem, cl, mean, stencil=None, workspace=None, use_moments=True):
    """One launch: (logp[C], gel[C,P,10], gcl[C,3], gmean[C]).  ``use_moments``: take the
    out-of-transit part of the sums from the data moments (the sweep then reads t only)."""
    lib = _lib.require_cuda()
    Cc, P = _elem_dims(elem)
    dev = data.t.device
    if data.y is None or data.ivar is None:
        raise ValueError("TransitData needs y and yerr for a likelihood")
    elem = _arg(elem, (Cc, P, NEL), "elem", dev)
    cl = _arg(cl, (Cc, 3), "cl", dev)
    mean = _arg(mean, (Cc,), "mean", dev)
    dtk, wk = _stencil_dev(stencil or exposure_stencil(), dev)
    need = lib.xo_transit_logp_grad_workspace(data.n, Cc, P)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(max(need, 8) // 8 + 1, dtype=torch.float64, device=dev)
    logp = torch.empty(Cc, dtype=torch.float64, device=dev)
    gel = torch.empty((Cc, P, NGRAD), dtype=torch.float64, device=dev)
    gcl = torch.empty((Cc, 3), dtype=torch.float64, device=dev)
    gmean = torch.empty(Cc, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.xo_transit_logp_grad(_ptr(data.t), _ptr(data.y), _ptr(data.ivar), data.n, _ptr(elem), _ptr(cl),
                                            _ptr(mean), Cc, P, _ptr(dtk), _ptr(wk), dtk.numel(), data.lognorm,
                                            _ptr(data.moments() if use_moments else None), _ptr(logp),
                                            _ptr(gel), _ptr(gcl), _ptr(gmean), _ptr(workspace),
                                            workspace.numel() * workspace.element_size(), _lib.stream_ptr()),
                   "transit_logp_grad")
    return logp, gel, gcl, gmean


class _TransitLogp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, elem, cl, mean, data, stencil):
        logp, gel, gcl, gmean = transit_logp_grad(data, elem, cl, mean, stencil)
        full = torch.zeros_like(elem)
        full[..., :NGRAD] = gel
        ctx.save_for_backward(full, gcl, gmean)
        return logp

    @staticmethod
    def backward(ctx, g):
        full, gcl, gmean = ctx.saved_tensors
        return g[:, None, None] * full, g[:, None] * gcl, g * gmean, None, None


class TransitLogp:
    """Fused ``logp`` Op: parameters in (O(#params) per chain), data resident on the device."""

    def __init__(self, data, texp=None, oversample=7, order=0):
        self.data = data
        self.stencil = exposure_stencil(texp, oversample, order)

    def __call__(self, elem, cl, mean):
        return _TransitLogp.apply(elem, cl, mean, self.data, self.stencil)


def orbit_eval(t, elem, want=("X", "Y", "Z", "rv")):
    """Relative position / R_star and RV shape for every (chain, cadence, planet)."""
    lib = _lib.require_cuda()
    Cc, P = _elem_dims(elem)
    t = _arg(t, (t.numel(),), "t", t.device)
    elem = _arg(elem, (Cc, P, NEL), "elem", t.device)
    out = {k: torch.empty((Cc, t.numel(), P), dtype=torch.float64, device=t.device) for k in want}
    with torch.cuda.device(t.device):
        _lib.check(lib.xo_orbit_eval(_ptr(t), t.numel(), _ptr(elem), Cc, P, _ptr(out.get("X")), _ptr(out.get("Y")),
                                     _ptr(out.get("Z")), _ptr(out.get("rv")), _lib.stream_ptr()), "orbit_eval")
    return out


def rv_logp_grad(data, elem, gamma, workspace=None):
    """RV Gaussian log-likelihood + gradient: (logp[C], grv[C,P,7], ggamma[C])."""
    lib = _lib.require_cuda()
    Cc, P = _elem_dims(elem)
    dev = data.t.device
    if data.y is None or data.ivar is None:
        raise ValueError("TransitData needs y and yerr for a likelihood")
    elem = _arg(elem, (Cc, P, NEL), "elem", dev)
    gamma = _arg(gamma, (Cc,), "gamma", dev)
    need = lib.xo_rv_logp_grad_workspace(data.n, Cc, P)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(max(need, 8) // 8 + 1, dtype=torch.float64, device=dev)
    logp = torch.empty(Cc, dtype=torch.float64, device=dev)
    grv = torch.empty((Cc, P, NGRAD_RV), dtype=torch.float64, device=dev)
    gg = torch.empty(Cc, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.xo_rv_logp_grad(_ptr(data.t), _ptr(data.y), _ptr(data.ivar), data.n, _ptr(elem), _ptr(gamma), Cc, P,
                                       data.lognorm, _ptr(logp), _ptr(grv), _ptr(gg), _ptr(workspace),
                                       workspace.numel() * workspace.element_size(), _lib.stream_ptr()), "rv_logp_grad")
    return logp, grv, gg


class _RvLogp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, elem, gamma, data):
        logp, grv, gg = rv_logp_grad(data, elem, gamma)
        full = torch.zeros_like(elem)
        full[..., :6] = grv[..., :6]
        full[..., _lib.EL_KRV] = grv[..., 6]
        ctx.save_for_backward(full, gg)
        return logp

    @staticmethod
    def backward(ctx, g):
        full, gg = ctx.saved_tensors
        return g[:, None, None] * full, g * gg, None


class RvLogp:
    def __init__(self, data):
        self.data = data

    def __call__(self, elem, gamma):
        return _RvLogp.apply(elem, gamma, self.data)


# ---------------------------------------------------------------------------
# parameter-level model: prologue + logp + reverse pass in three launches
# ---------------------------------------------------------------------------
def pack_params(period, t0, ecc, omega, b, ror, m_planet=0.0, K=0.0):
    """[..., P] tensors -> par[..., P, 8] in the order the C ABI expects (XO_NPAR)."""
    period = torch.as_tensor(period, dtype=torch.float64)
    cols = [torch.as_tensor(x, dtype=torch.float64, device=period.device) for x in (period, t0, ecc, omega, b, ror, m_planet, K)]
    return torch.stack(torch.broadcast_tensors(*cols), dim=-1).contiguous()


def pack_star(r_star, m_star, u1, u2):
    """[C] tensors -> star[C, 4] (XO_NSTAR)."""
    u1 = torch.as_tensor(u1, dtype=torch.float64)
    cols = [torch.as_tensor(x, dtype=torch.float64, device=u1.device) for x in (r_star, m_star, u1, u2)]
    return torch.stack(torch.broadcast_tensors(*cols), dim=-1).contiguous()


def _par_dims(par):
    if not isinstance(par, torch.Tensor) or par.dim() != 3 or par.shape[-1] != _lib.NPAR:
        raise ValueError("par must have shape [C, P, %d] (pack_params), got %s" % (_lib.NPAR, tuple(getattr(par, "shape", ()))))
    if par.shape[1] > _lib.MAX_PLANETS:
        raise ValueError("at most %d planets per chain" % _lib.MAX_PLANETS)
    return par.shape[0], par.shape[1]


class TransitModel:
    """``logp, d logp/d params`` of a transit light-curve model with Gaussian noise, from
    sampler-side parameters, entirely on the device (the fused counterpart of the PyMC model
    in /root/reference/docs/tutorials/data-and-models.md:458-509)."""

    def __init__(self, data, texp=None, oversample=7, order=0, use_in_transit=True, use_moments=True, cadence_group=None):
        """``cadence_group``: a torch.distributed process group (or True for the default group) when
        ``data`` holds only this rank's block of the light curve (parallel.cadence_bounds): logp and the
        gradients are then summed over the group with one packed all-reduce per evaluation."""
        self.data = data
        self.cadence_group = cadence_group
        self.texp = 0.0 if texp is None else float(texp)
        self.use_in_transit = bool(use_in_transit)
        self.use_moments = bool(use_moments)
        st = exposure_stencil(texp, oversample, order)
        self.dtk, self.wk = _stencil_dev(st, data.device)
        self._ws = None

    def logp_grad(self, par, star, mean):
        lib = _lib.require_cuda()
        d = self.data
        Cc, P = _par_dims(par)
        par = _arg(par, (Cc, P, _lib.NPAR), "par", d.device)
        star = _arg(star, (Cc, _lib.NSTAR), "star", d.device)
        mean = _arg(mean, (Cc,), "mean", d.device)
        need = lib.xo_transit_model_workspace(d.n, Cc, P)
        if self._ws is None or self._ws.numel() * 8 < need:
            self._ws = torch.empty(need // 8 + 1, dtype=torch.float64, device=d.device)
        logp = torch.empty(Cc, dtype=torch.float64, device=d.device)
        gpar = torch.empty_like(par)
        gstar = torch.empty_like(star)
        gmean = torch.empty_like(mean)
        with torch.cuda.device(d.device):
            _lib.check(lib.xo_transit_model_logp_grad(_ptr(d.t), _ptr(d.y), _ptr(d.ivar), d.n, _ptr(par), _ptr(star), _ptr(mean), Cc, P,
                                                      _ptr(self.dtk), _ptr(self.wk), self.dtk.numel(), self.texp,
                                                      int(self.use_in_transit), d.lognorm,
                                                      _ptr(d.moments() if self.use_moments else None), _ptr(logp), _ptr(gpar), _ptr(gstar),
                                                      _ptr(gmean), _ptr(self._ws), self._ws.numel() * 8, _lib.stream_ptr()),
                       "transit_model_logp_grad")
        if self.cadence_group is not None:
            from . import parallel

            parallel.allreduce_sum_([logp, gpar, gstar, gmean], None if self.cadence_group is True else self.cadence_group)
        return logp, gpar, gstar, gmean

    def __call__(self, par, star, mean):
        return _TransitModelLogp.apply(par, star, mean, self)


class _TransitModelLogp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, par, star, mean, model):
        logp, gpar, gstar, gmean = model.logp_grad(par, star, mean)
        ctx.save_for_backward(gpar, gstar, gmean)
        return logp

    @staticmethod
    def backward(ctx, g):
        gpar, gstar, gmean = ctx.saved_tensors
        return g[:, None, None] * gpar, g[:, None] * gstar, g * gmean, None


class TransitGPModel:
    """Transit mean model + celerite SHO GP noise model (BASELINE configs[3]): ``loglike[C]`` and its gradient w.r.t.
    the orbital / stellar parameters, the mean flux, the GP hyper-parameters (S0, w0, Q) and an optional per-chain
    jitter term added to ``diag``, in one stream-ordered C call (``xo_transit_gp_logp_grad``): no torch op, no host
    round trip between the light curve, the residuals, the GP scans and the light curve's reverse pass.

    ``t, y, diag``: [N] (diag = sigma^2, the argument of celerite2's ``gp.compute(t, diag=...)``)."""

    def __init__(self, t, y, diag, texp=None, oversample=7, order=0, use_in_transit=True, device=None, eps=1e-5):
        dev = torch.device("cuda" if device is None else device)
        as_dev = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float64)).to(dev).contiguous()  # noqa: E731
        self.t, self.y, self.diag = as_dev(t), as_dev(y), as_dev(diag)
        self.n = self.t.numel()
        if self.y.shape != (self.n,) or self.diag.shape != (self.n,):
            raise ValueError("t, y, diag must have the same length")
        self.device = dev
        self.texp = 0.0 if texp is None else float(texp)
        self.use_in_transit = bool(use_in_transit)
        self.eps = float(eps)
        self.dtk, self.wk = _stencil_dev(exposure_stencil(texp, oversample, order), dev)
        self._ws = None

    def logp_grad(self, par, star, mean, S0, w0, Q, jitter=None):
        """-> dict(loglike[C], gpar[C,P,8], gstar[C,4], gmean[C], ghyper[C,3], gjitter[C] | None)"""
        lib = _lib.require_cuda()
        dev = self.device
        Cc, P = _par_dims(par)
        par = _arg(par, (Cc, P, _lib.NPAR), "par", dev)
        star = _arg(star, (Cc, _lib.NSTAR), "star", dev)
        mean = _arg(mean, (Cc,), "mean", dev)
        S0, w0, Q = _arg(S0, (Cc,), "S0", dev), _arg(w0, (Cc,), "w0", dev), _arg(Q, (Cc,), "Q", dev)
        jit = None if jitter is None else _arg(jitter, (Cc,), "jitter", dev)
        need = lib.xo_transit_gp_workspace(self.n, Cc, P)
        if self._ws is None or self._ws.numel() * 8 < need:
            self._ws = torch.empty(need // 8 + 1, dtype=torch.float64, device=dev)
        new = lambda *shape: torch.empty(shape, dtype=torch.float64, device=dev)  # noqa: E731
        out = dict(loglike=new(Cc), gpar=new(Cc, P, _lib.NPAR), gstar=new(Cc, _lib.NSTAR), gmean=new(Cc), ghyper=new(Cc, 3),
                   gjitter=None if jit is None else new(Cc))
        with torch.cuda.device(dev):
            _lib.check(lib.xo_transit_gp_logp_grad(_ptr(self.t), _ptr(self.y), _ptr(self.diag), self.n, _ptr(par), _ptr(star), _ptr(mean),
                                                   _ptr(S0), _ptr(w0), _ptr(Q), _ptr(jit), Cc, P, _ptr(self.dtk), _ptr(self.wk),
                                                   self.dtk.numel(), self.texp, int(self.use_in_transit), self.eps, _ptr(out["loglike"]),
                                                   _ptr(out["gpar"]), _ptr(out["gstar"]), _ptr(out["gmean"]), _ptr(out["ghyper"]),
                                                   _ptr(out["gjitter"]), _ptr(self._ws), self._ws.numel() * 8, _lib.stream_ptr()),
                       "transit_gp_logp_grad")
        return out

    def status(self):
        """(bad_lanes, wait_failures) of the last evaluation (xo_read_status); synchronises the current stream"""
        from . import gp

        return gp.read_status(self._ws)
