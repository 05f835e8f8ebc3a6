"""The three-name operator namespace the reference consumes as ``exoplanet.compat.ops``
(/root/reference/src/exoplanet/compat.py:56), over PyTorch CUDA tensors.

    sinf, cosf = ops.kepler(M, e)                         # keplerian.py:333
    s          = ops.quad_solution_vector(b, r)           # limb_dark.py:24, shape b.shape + (3,)
    Ml, Mr, fl = ops.contact_points(a, e, cosw, sinw, cosi, sini, L)   # keplerian.py:744-753

CUDA fp64 tensors in -> CUDA fp64 tensors out with reverse-mode gradients
(torch.autograd.Function whose backward launches the VJP kernels).  CPU torch tensors are moved to the
current CUDA device, go through the same autograd Functions (so gradients are kept) and the results come
back on the CPU.  numpy arrays in -> numpy arrays out through the host-buffer C entry points (the shape
PyTensor's ``Op.perform`` needs); still computed on the GPU, no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

__all__ = ["kepler", "quad_solution_vector", "quad_solution_vector_with_jac", "contact_points"]


def _ptr(x):
    return C.c_void_p(x.data_ptr()) if x is not None else None


def _nptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _is_cuda(*xs):
    """torch path: any torch tensor among the inputs (numpy-only calls take the host-buffer entry points)"""
    return any(isinstance(x, torch.Tensor) for x in xs)


def _torch_device(*xs):
    """(device the kernels run on, device the results are returned on)"""
    for x in xs:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device, x.device
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device()), torch.device("cpu")


def _as_cuda(x, device):
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.asarray(x, dtype=np.float64))
    return x.to(device=device, dtype=torch.float64)


def _device_index(device=None):
    if device is None:
        return torch.cuda.current_device() if torch.cuda.is_available() else 0
    return torch.device(device).index or 0


# --------------------------------------------------------------------------- kepler
class _Kepler(torch.autograd.Function):
    @staticmethod
    def forward(ctx, M, e):
        lib = _lib.require_cuda()
        M = M.contiguous()
        e = e.contiguous()
        sinf = torch.empty_like(M)
        cosf = torch.empty_like(M)
        with torch.cuda.device(M.device):
            _lib.check(lib.xo_kepler_fwd(_ptr(M), _ptr(e), _ptr(sinf), _ptr(cosf), M.numel(), _lib.stream_ptr()), "kepler")
        ctx.save_for_backward(sinf, cosf, e)
        return sinf, cosf

    @staticmethod
    def backward(ctx, bs, bc):
        lib = _lib.require_cuda()
        sinf, cosf, e = ctx.saved_tensors
        bs = torch.zeros_like(sinf) if bs is None else bs.contiguous()
        bc = torch.zeros_like(sinf) if bc is None else bc.contiguous()
        bM = torch.empty_like(sinf)
        be = torch.empty_like(sinf)
        with torch.cuda.device(sinf.device):
            _lib.check(lib.xo_kepler_vjp(_ptr(sinf), _ptr(cosf), _ptr(e), _ptr(bs), _ptr(bc), _ptr(bM), _ptr(be), sinf.numel(),
                                         _lib.stream_ptr()), "kepler_vjp")
        return bM, be


def kepler(M, e):
    """Solve Kepler's equation; returns (sin f, cos f) of the true anomaly.

    ``e`` is broadcast against ``M`` (the reference materialises ``ecc + zeros_like(M)``,
    keplerian.py:333)."""
    if _is_cuda(M, e):
        dev, out = _torch_device(M, e)
        M = _as_cuda(M, dev)
        e = _as_cuda(e, dev)
        M, e = torch.broadcast_tensors(M, e)
        sf, cf = _Kepler.apply(M, e)
        return sf.to(out), cf.to(out)
    lib = _lib.require_cuda()
    M = np.asarray(M, dtype=np.float64)
    e = np.asarray(e, dtype=np.float64)
    M, e = [np.ascontiguousarray(x) for x in np.broadcast_arrays(M, e)]
    sinf = np.empty_like(M)
    cosf = np.empty_like(M)
    _lib.check(lib.xo_kepler_host(_nptr(M), _nptr(e), _nptr(sinf), _nptr(cosf), M.size, _device_index()), "kepler_host")
    return sinf, cosf


# --------------------------------------------------------------------------- limb darkening
class _QuadSolutionVector(torch.autograd.Function):
    @staticmethod
    def forward(ctx, b, r):
        lib = _lib.require_cuda()
        b = b.contiguous()
        r = r.contiguous()
        s = torch.empty(b.shape + (3,), dtype=torch.float64, device=b.device)
        need = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        dsdb = torch.empty_like(s) if need else None
        dsdr = torch.empty_like(s) if need else None
        with torch.cuda.device(b.device):
            _lib.check(lib.xo_qsv_fwd(_ptr(b), _ptr(r), _ptr(s), _ptr(dsdb), _ptr(dsdr), b.numel(), _lib.stream_ptr()), "qsv")
        if need:
            ctx.save_for_backward(dsdb, dsdr)
        ctx.need = need
        return s

    @staticmethod
    def backward(ctx, bs):
        lib = _lib.require_cuda()
        dsdb, dsdr = ctx.saved_tensors
        bs = bs.contiguous()
        bb = torch.empty(dsdb.shape[:-1], dtype=torch.float64, device=dsdb.device)
        br = torch.empty_like(bb)
        with torch.cuda.device(dsdb.device):
            _lib.check(lib.xo_qsv_vjp(_ptr(bs), _ptr(dsdb), _ptr(dsdr), _ptr(bb), _ptr(br), bb.numel(), _lib.stream_ptr()),
                       "qsv_vjp")
        return bb, br


def quad_solution_vector(b, r):
    """Agol-Luger-Foreman-Mackey solution vector for quadratic limb darkening:
    shape ``b.shape + (3,)``; ``lc = s @ c - 1`` (limb_dark.py:21-24)."""
    if _is_cuda(b, r):
        dev, out = _torch_device(b, r)
        b = _as_cuda(b, dev)
        r = _as_cuda(r, dev)
        b, r = torch.broadcast_tensors(b, r)
        return _QuadSolutionVector.apply(b, r).to(out)
    return quad_solution_vector_with_jac(b, r, jac=False)


def quad_solution_vector_with_jac(b, r, jac=True):
    """Host-array variant: numpy in, numpy (s, dsdb, dsdr) out."""
    lib = _lib.require_cuda()
    b = np.asarray(b, dtype=np.float64)
    r = np.asarray(r, dtype=np.float64)
    b, r = [np.ascontiguousarray(x) for x in np.broadcast_arrays(b, r)]
    s = np.empty(b.shape + (3,))
    dsdb = np.empty_like(s) if jac else None
    dsdr = np.empty_like(s) if jac else None
    _lib.check(lib.xo_qsv_host(_nptr(b), _nptr(r), _nptr(s), _nptr(dsdb), _nptr(dsdr), b.size, _device_index()), "qsv_host")
    return (s, dsdb, dsdr) if jac else s


# --------------------------------------------------------------------------- contact points
def contact_points(a, e, cosw, sinw, cosi, sini, L):
    """Mean anomalies of first/last contact with a circle of radius L; no gradient
    (outputs feed comparisons only, keplerian.py:753-769).  flag != 0 => no window."""
    args = (a, e, cosw, sinw, cosi, sini, L)
    if _is_cuda(*args):
        lib = _lib.require_cuda()
        dev, out = _torch_device(*args)
        ts = torch.broadcast_tensors(*[_as_cuda(x, dev).detach() for x in args])
        ts = [x.contiguous() for x in ts]
        Ml = torch.empty_like(ts[0])
        Mr = torch.empty_like(ts[0])
        flag = torch.empty(ts[0].shape, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.xo_contact_points(*[_ptr(x) for x in ts], _ptr(Ml), _ptr(Mr), _ptr(flag), ts[0].numel(),
                                             _lib.stream_ptr()), "contact_points")
        return Ml.to(out), Mr.to(out), flag.to(out)
    lib = _lib.require_cuda()
    arrs = [np.ascontiguousarray(x) for x in np.broadcast_arrays(*[np.asarray(x, dtype=np.float64) for x in args])]
    Ml = np.empty_like(arrs[0])
    Mr = np.empty_like(arrs[0])
    flag = np.empty(arrs[0].shape, dtype=np.int32)
    _lib.check(lib.xo_contact_points_host(*[_nptr(x) for x in arrs], _nptr(Ml), _nptr(Mr), _nptr(flag), arrs[0].size,
                                          _device_index()), "contact_points_host")
    return Ml, Mr, flag
