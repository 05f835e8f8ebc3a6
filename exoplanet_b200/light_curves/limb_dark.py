"""``LimbDarkLightCurve`` over PyTorch CUDA tensors -- host-side mirror of
/root/reference/src/exoplanet/light_curves/limb_dark.py:27-252.

Two execution paths behind one interface:
 * fused (default for a ``KeplerianOrbit`` without light delay): the orbit's element table
   goes to ``xo_transit_lc_fwd``; the backward pass is ``xo_transit_lc_vjp`` -- one launch
   each instead of the dozens of elementwise sweeps the reference graph makes;
 * generic (any object with ``get_relative_position`` / ``in_transit``, or light delay): the
   reference's own sequence gather -> exposure grid -> position -> ``ops.quad_solution_vector``
   -> stencil sum -> scatter, each step a CUDA op with autograd.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib, fused, ops
from ..orbits.keplerian import KeplerianOrbit, _tensor

__all__ = ["LimbDarkLightCurve"]


def _ptr(x):
    return C.c_void_p(x.data_ptr()) if x is not None else None


class _FusedLightCurve(torch.autograd.Function):
    """lc[N, P] = f(elem[P, 16], cl[3]); gradients w.r.t. the element table and cl."""

    @staticmethod
    def forward(ctx, elem, cl, t, dtk, wk):
        lib = _lib.require_cuda()
        e3 = elem.detach().contiguous().unsqueeze(0)
        c2 = cl.detach().contiguous().unsqueeze(0)
        P = e3.shape[1]
        lc = torch.empty((1, t.numel(), P), dtype=torch.float64, device=t.device)
        with torch.cuda.device(t.device):
            _lib.check(lib.xo_transit_lc_fwd(_ptr(t), t.numel(), _ptr(e3), _ptr(c2), 1, P, _ptr(dtk), _ptr(wk), dtk.numel(), _ptr(lc),
                                             _lib.stream_ptr()), "transit_lc_fwd")
        ctx.save_for_backward(e3, c2, t, dtk, wk)
        return lc[0]

    @staticmethod
    def backward(ctx, bar):
        lib = _lib.require_cuda()
        e3, c2, t, dtk, wk = ctx.saved_tensors
        P = e3.shape[1]
        bar = bar.contiguous().unsqueeze(0)
        need = lib.xo_transit_logp_grad_workspace(t.numel(), 1, P) + 16
        ws = torch.empty(need // 8 + 2, dtype=torch.float64, device=t.device)
        gel = torch.empty((1, P, _lib.NGRAD), dtype=torch.float64, device=t.device)
        gcl = torch.empty((1, 3), dtype=torch.float64, device=t.device)
        with torch.cuda.device(t.device):
            _lib.check(lib.xo_transit_lc_vjp(_ptr(t), t.numel(), _ptr(e3), _ptr(c2), 1, P, _ptr(dtk), _ptr(wk), dtk.numel(), _ptr(bar),
                                             _ptr(gel), _ptr(gcl), _ptr(ws), ws.numel() * 8, _lib.stream_ptr()), "transit_lc_vjp")
        full = torch.zeros_like(e3[0])
        full[:, :_lib.NGRAD] = gel[0]
        return full, gcl[0], None, None, None


class _FusedLightCurveDelay(torch.autograd.Function):
    """lc[N, P] with light-travel delay (keplerian.py:411-470): every sample at its retarded time (a second Kepler solve in
    the kernel); gradients w.r.t. the element table (sin i included) and cl.  elem[:, 14] = R_star."""

    @staticmethod
    def forward(ctx, elem, cl, t, dtk, wk):
        lib = _lib.require_cuda()
        e3 = elem.detach().contiguous().unsqueeze(0)
        c2 = cl.detach().contiguous().unsqueeze(0)
        P = e3.shape[1]
        lc = torch.empty((1, t.numel(), P), dtype=torch.float64, device=t.device)
        with torch.cuda.device(t.device):
            _lib.check(lib.xo_transit_lc_fwd_ex(_ptr(t), t.numel(), _ptr(e3), _ptr(c2), 1, P, _ptr(dtk), _ptr(wk), dtk.numel(), None, None, 0,
                                                1, None, _ptr(lc), _lib.stream_ptr()), "transit_lc_fwd_ex")
        ctx.save_for_backward(e3, c2, t, dtk, wk)
        return lc[0]

    @staticmethod
    def backward(ctx, bar):
        lib = _lib.require_cuda()
        e3, c2, t, dtk, wk = ctx.saved_tensors
        P = e3.shape[1]
        bar = bar.contiguous().unsqueeze(0)
        need = lib.xo_transit_logp_grad_workspace(t.numel(), 1, P) + 16
        ws = torch.empty(need // 8 + 2, dtype=torch.float64, device=t.device)
        gel = torch.empty((1, P, _lib.NGRAD), dtype=torch.float64, device=t.device)
        gcl = torch.empty((1, 3), dtype=torch.float64, device=t.device)
        with torch.cuda.device(t.device):
            _lib.check(lib.xo_transit_lc_vjp_ex(_ptr(t), t.numel(), _ptr(e3), _ptr(c2), 1, P, _ptr(dtk), _ptr(wk), dtk.numel(), None, None, 0,
                                                1, None, _ptr(bar), _ptr(gel), _ptr(gcl), None, _ptr(ws), ws.numel() * 8, _lib.stream_ptr()),
                       "transit_lc_vjp_ex")
        full = torch.zeros_like(e3[0])
        full[:, :_lib.NGRAD] = gel[0]
        return full, gcl[0], None, None, None


class _FusedLightCurveTTV(torch.autograd.Function):
    """lc[N, P] = f(elem[P, 16], cl[3], transit times[P, NT]) for a TTVOrbit (ttv.py:175-187: every cadence is measured
    from its nearest transit); gradients w.r.t. the element table, cl and every transit time."""

    @staticmethod
    def forward(ctx, elem, cl, times, t, dtk, wk, count):
        lib = _lib.require_cuda()
        e3 = elem.detach().contiguous().unsqueeze(0)
        c2 = cl.detach().contiguous().unsqueeze(0)
        tt = times.detach().contiguous()
        P, nt = tt.shape
        lc = torch.empty((1, t.numel(), P), dtype=torch.float64, device=t.device)
        with torch.cuda.device(t.device):
            _lib.check(lib.xo_transit_lc_fwd_ex(_ptr(t), t.numel(), _ptr(e3), _ptr(c2), 1, P, _ptr(dtk), _ptr(wk), dtk.numel(), _ptr(tt),
                                                _ptr(count), nt, 0, None, _ptr(lc), _lib.stream_ptr()), "transit_lc_fwd_ex")
        ctx.save_for_backward(e3, c2, tt, t, dtk, wk, count)
        return lc[0]

    @staticmethod
    def backward(ctx, bar):
        lib = _lib.require_cuda()
        e3, c2, tt, t, dtk, wk, count = ctx.saved_tensors
        P, nt = tt.shape
        bar = bar.contiguous().unsqueeze(0)
        need = lib.xo_transit_logp_grad_workspace(t.numel(), 1, P) + 16
        ws = torch.empty(need // 8 + 2, dtype=torch.float64, device=t.device)
        gel = torch.empty((1, P, _lib.NGRAD), dtype=torch.float64, device=t.device)
        gcl = torch.empty((1, 3), dtype=torch.float64, device=t.device)
        gtt = torch.empty((1, P, nt), dtype=torch.float64, device=t.device)
        with torch.cuda.device(t.device):
            _lib.check(lib.xo_transit_lc_vjp_ex(_ptr(t), t.numel(), _ptr(e3), _ptr(c2), 1, P, _ptr(dtk), _ptr(wk), dtk.numel(), _ptr(tt),
                                                _ptr(count), nt, 0, None, _ptr(bar), _ptr(gel), _ptr(gcl), _ptr(gtt), _ptr(ws), ws.numel() * 8,
                                                _lib.stream_ptr()), "transit_lc_vjp_ex")
        full = torch.zeros_like(e3[0])
        full[:, :_lib.NGRAD] = gel[0]
        full[:, 0] = 0.0  # the time axis is anchored on the transit times, not on elem's t0 (its derivative is in gtt)
        valid = torch.arange(nt, device=t.device)[None, :] < count[:, None]
        return full, gcl[0], torch.where(valid, gtt[0], torch.zeros_like(gtt[0])), None, None, None, None


class _FusedLightCurveTexp(torch.autograd.Function):
    """lc[N, P] with one exposure time per cadence (limb_dark.py:199-205): sample k of cadence i sits at t_i + texp[i] dt_k;
    ``dtk`` is the unit stencil.  Gradients w.r.t. the element table and cl (exposure times are data)."""

    @staticmethod
    def forward(ctx, elem, cl, t, dtk, wk, texp_vec):
        lib = _lib.require_cuda()
        e3 = elem.detach().contiguous().unsqueeze(0)
        c2 = cl.detach().contiguous().unsqueeze(0)
        P = e3.shape[1]
        lc = torch.empty((1, t.numel(), P), dtype=torch.float64, device=t.device)
        with torch.cuda.device(t.device):
            _lib.check(lib.xo_transit_lc_fwd_ex(_ptr(t), t.numel(), _ptr(e3), _ptr(c2), 1, P, _ptr(dtk), _ptr(wk), dtk.numel(), None, None, 0,
                                                0, _ptr(texp_vec), _ptr(lc), _lib.stream_ptr()), "transit_lc_fwd_ex")
        ctx.save_for_backward(e3, c2, t, dtk, wk, texp_vec)
        return lc[0]

    @staticmethod
    def backward(ctx, bar):
        lib = _lib.require_cuda()
        e3, c2, t, dtk, wk, texp_vec = ctx.saved_tensors
        P = e3.shape[1]
        bar = bar.contiguous().unsqueeze(0)
        need = lib.xo_transit_logp_grad_workspace(t.numel(), 1, P) + 16
        ws = torch.empty(need // 8 + 2, dtype=torch.float64, device=t.device)
        gel = torch.empty((1, P, _lib.NGRAD), dtype=torch.float64, device=t.device)
        gcl = torch.empty((1, 3), dtype=torch.float64, device=t.device)
        with torch.cuda.device(t.device):
            _lib.check(lib.xo_transit_lc_vjp_ex(_ptr(t), t.numel(), _ptr(e3), _ptr(c2), 1, P, _ptr(dtk), _ptr(wk), dtk.numel(), None, None, 0,
                                                0, _ptr(texp_vec), _ptr(bar), _ptr(gel), _ptr(gcl), None, _ptr(ws), ws.numel() * 8,
                                                _lib.stream_ptr()), "transit_lc_vjp_ex")
        full = torch.zeros_like(e3[0])
        full[:, :_lib.NGRAD] = gel[0]
        return full, gcl[0], None, None, None, None


class LimbDarkLightCurve:
    """A quadratically limb-darkened light curve (coefficients ``u1``, ``u2``)."""

    def __init__(self, u1, u2=None, model=None, device=None):
        if u2 is None:  # deprecated vector form (limb_dark.py:43-65)
            u = _tensor(u1, device)
            if u.dim() != 1 or u.shape[0] != 2:
                raise AssertionError("only quadratic limb darkening is supported; use `starry` for more flexibility")
            u1, u2 = u[0], u[1]
        self.u1 = _tensor(u1, device)
        self.u2 = _tensor(u2, self.u1.device)
        self.c = fused.limb_dark_coeffs(self.u1, self.u2)

    def get_ror_from_approx_transit_depth(self, delta, b, jac=False):
        """Radius ratio for an approximate depth at impact parameter |b| < 1 (small-planet
        approximation, limb_dark.py:70-97); optionally d ror / d delta."""
        b = _tensor(b, self.u1.device)
        delta = _tensor(delta, self.u1.device)
        f0 = 1.0 - 2.0 * self.u1 / 6.0 - 2.0 * self.u2 / 12.0
        arg = 1.0 - torch.sqrt(1.0 - b ** 2)
        factor = f0 / (1.0 - self.u1 * arg - self.u2 * arg ** 2)
        ror = torch.sqrt(delta * factor)
        if not jac:
            return ror.reshape(b.shape)
        return ror.reshape(b.shape), (0.5 * factor / ror).reshape(b.shape)

    def get_light_curve(self, orbit=None, r=None, t=None, texp=None, oversample=7, order=0, use_in_transit=None,
                        light_delay=False):
        """Relative flux [n_times, n_planets] (0 out of transit) -- limb_dark.py:99-232.

        ``r``: planet radii in the units of ``orbit.r_star``; ``texp``: exposure time (scalar, or
        per-cadence vector on the generic path); ``order`` 0/1/2 = Riemann/trapezoid/Simpson."""
        for name, val in (("orbit", orbit), ("r", r), ("t", t)):
            if val is None:
                raise ValueError("missing required argument '%s'" % name)
        if order not in (0, 1, 2):
            raise ValueError("order must be <= 2")
        use_in_transit = (not light_delay) if use_in_transit is None else use_in_transit
        dev = self.u1.device
        r = _tensor(r, dev).reshape(-1)
        t = _tensor(t, dev)
        scalar_texp = texp is None or np.ndim(texp) == 0 or (isinstance(texp, torch.Tensor) and texp.dim() == 0)
        if type(orbit) is KeplerianOrbit and not light_delay and scalar_texp and t.dim() == 1 and t.is_cuda:
            tex = None if texp is None else float(texp)
            elem = orbit.elements(r=r, texp=tex, use_in_transit=use_in_transit)
            dtk, wk = fused._stencil_dev(fused.exposure_stencil(tex, oversample, order), dev)
            return _FusedLightCurve.apply(elem, self.c, t.contiguous(), dtk, wk)
        if type(orbit) is KeplerianOrbit and light_delay and scalar_texp and t.dim() == 1 and t.is_cuda and not (
                isinstance(orbit.r_star, torch.Tensor) and orbit.r_star.requires_grad):
            # the fused delay kernels do not return d / d R_star: an R_star that is being differentiated takes the op-by-op path
            tex = None if texp is None else float(texp)
            elem = orbit.elements(r=r, texp=tex, use_in_transit=bool(use_in_transit), light_delay=True)
            dtk, wk = fused._stencil_dev(fused.exposure_stencil(tex, oversample, order), dev)
            return _FusedLightCurveDelay.apply(elem, self.c, t.contiguous(), dtk, wk)
        if type(orbit) is KeplerianOrbit and not light_delay and not scalar_texp and t.dim() == 1 and t.is_cuda:
            # one exposure time per cadence (limb_dark.py:199-205): unit stencil, scaled per cadence in the kernel; the window
            # carries the longest exposure (what the reference's broadcast of t_start - texp/2 amounts to, keplerian.py:766-770)
            tv = _tensor(texp, dev).to(torch.float64).reshape(-1).contiguous()
            if tv.numel() == t.numel():
                elem = orbit.elements(r=r, texp=float(tv.max()), use_in_transit=use_in_transit)
                dtk, wk = fused._stencil_dev(fused.exposure_stencil(1.0, oversample, order), dev)
                return _FusedLightCurveTexp.apply(elem, self.c, t.contiguous(), dtk, wk, tv)
        from ..orbits.ttv import TTVOrbit

        if type(orbit) is TTVOrbit and not light_delay and scalar_texp and t.dim() == 1 and t.is_cuda:
            tex = None if texp is None else float(texp)
            elem = orbit.elements(r=r, texp=tex, use_in_transit=use_in_transit)
            times, count = orbit.ttv_table()
            dtk, wk = fused._stencil_dev(fused.exposure_stencil(tex, oversample, order), dev)
            return _FusedLightCurveTTV.apply(elem, self.c, times, t.contiguous(), dtk, wk, count)
        return self._generic_light_curve(orbit, r, t, texp, oversample, order, use_in_transit, light_delay)

    def _generic_light_curve(self, orbit, r, t, texp, oversample, order, use_in_transit, light_delay):
        dev = t.device
        full_shape = tuple(t.shape) + (r.numel(),)
        inds = None
        if use_in_transit:
            inds = orbit.in_transit(t, r=r, texp=texp, light_delay=light_delay)
            t = t[inds]
        if texp is None:
            tgrid, weights = t, None
        else:
            dtk, wk = fused.exposure_stencil(1.0, oversample, order)
            texp_t = _tensor(texp, dev)
            if texp_t.dim() > 0 and inds is not None:
                texp_t = texp_t[inds]
            step = texp_t.unsqueeze(-1) * _tensor(dtk, dev) if texp_t.dim() > 0 else texp_t * _tensor(dtk, dev)
            tgrid = t.unsqueeze(-1) + step
            weights = _tensor(wk, dev)
        x, y, z = orbit.get_relative_position(tgrid, light_delay=light_delay)
        shape = tuple(tgrid.shape) + (r.numel(),)
        b = torch.sqrt(x * x + y * y).reshape(shape)
        z = z.reshape(shape)
        rr = r.expand(shape)
        lc = self._compute_light_curve(b / orbit.r_star, rr / orbit.r_star, z / orbit.r_star)
        if weights is not None:
            lc = (lc * weights.reshape((1,) * t.dim() + (-1, 1))).sum(dim=t.dim())
        if inds is None:
            return lc
        out = torch.zeros(full_shape, dtype=torch.float64, device=dev)
        return out.index_put((inds,), lc)

    def _compute_light_curve(self, b, r, los=None):
        """``s(b, r) . c - 1`` where ``los > 0``, else 0 (limb_dark.py:234-252); b, r in stellar radii."""
        dev = self.u1.device
        b = _tensor(b, dev)
        r = _tensor(r, dev)
        s = ops.quad_solution_vector(b, r + torch.zeros_like(b))
        lc = (s * self.c).sum(dim=-1) - 1.0
        if los is None:
            return lc
        return torch.where(_tensor(los, dev) > 0, lc, torch.zeros_like(lc))
