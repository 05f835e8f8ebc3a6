"""PyTensor ``Op`` wrappers: the drop-in for ``exoplanet_core.pymc.ops`` (the object the
reference binds at /root/reference/src/exoplanet/compat.py:56; ``shim/exoplanet_core`` re-exports this module under
that name), plus one fused ``Op`` for a whole transit likelihood.

PyTensor / PyMC are not installable in the build image: the module is written against the public
``pytensor.graph.op.Op`` contract (``make_node`` / ``perform`` / ``infer_shape`` / ``grad``), kept structurally identical
to the torch layer in ``ops.py`` / ``fused.py``, and executed in the test-suite on a minimal stand-in for PyTensor
(tests/fake_pytensor.py: ``make_node``, ``perform`` and ``grad`` graphs evaluated eagerly on numpy).  Importing it
without PyTensor raises ImportError.

Per-op wrappers receive HOST numpy arrays in ``perform`` (PyTensor has no GPU backend), so they go through the
``*_host`` C entry points and pay H2D + D2H per call; ``TransitLogpOp`` keeps t, y, sigma on the device and moves only
O(#params) per evaluation -- use it for sampling.  The CUDA device is ``XO_B200_DEVICE`` (default 0) for the per-op
wrappers and the device of the model's data for ``TransitLogpOp``.
"""
import ctypes as C
import os

import numpy as np

try:  # pragma: no cover - pytensor is absent from the build image
    import pytensor
    import pytensor.tensor as pt
    from pytensor.graph.basic import Apply
    from pytensor.graph.op import Op
except ImportError as exc:  # pragma: no cover
    raise ImportError("exoplanet_b200.pytensor_ops needs pytensor (pymc >= 5)") from exc

from . import _lib  # noqa: E402

__all__ = ["kepler", "quad_solution_vector", "contact_points", "TransitLogpOp"]


def _device():
    return int(os.environ.get("XO_B200_DEVICE", "0"))


def _np(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(x):
    return pt.as_tensor_variable(x).astype("float64")


def _bcast(*xs):
    """inputs broadcast to one common shape (the reference passes ``ecc + zeros_like(M)``, keplerian.py:333, but a
    scalar against an array must work too), so that the output type can be taken from the first one"""
    return list(pt.broadcast_arrays(*[_f64(x) for x in xs]))


def _disconnected(g):
    return isinstance(g.type, pytensor.gradient.DisconnectedType)


class Kepler(Op):
    """(M, e) -> (sin f, cos f); grad from the outputs alone, no second solve."""

    __props__ = ()

    def make_node(self, M, e):
        M, e = _bcast(M, e)
        return Apply(self, [M, e], [M.type(), M.type()])

    def infer_shape(self, fgraph, node, shapes):
        return [shapes[0], shapes[0]]

    def perform(self, node, inputs, outputs):
        M = _np(inputs[0])
        e = _np(np.broadcast_to(inputs[1], M.shape))
        sf, cf = np.empty_like(M), np.empty_like(M)
        _lib.check(_lib.require_cuda().xo_kepler_host(_p(M), _p(e), _p(sf), _p(cf), M.size, _device()), "kepler_host")
        outputs[0][0], outputs[1][0] = sf, cf

    def grad(self, inputs, gout):
        M, e = inputs
        sinf, cosf = self(M, e)
        zero = pt.zeros_like(M)
        bs = zero if _disconnected(gout[0]) else gout[0]
        bc = zero if _disconnected(gout[1]) else gout[1]
        bf = bs * cosf - bc * sinf
        ome2 = 1.0 - e ** 2
        return [bf * (1.0 + e * cosf) ** 2 / ome2 ** 1.5, bf * (2.0 + e * cosf) * sinf / ome2]


class QuadSolutionVector(Op):
    """(b, r) -> (s, ds/db, ds/dr), each shape + (3,); the public function returns s only."""

    __props__ = ()

    def make_node(self, b, r):
        b, r = _bcast(b, r)
        out = pt.TensorType("float64", shape=(None,) * (b.ndim + 1))
        return Apply(self, [b, r], [out(), out(), out()])

    def infer_shape(self, fgraph, node, shapes):
        s = tuple(shapes[0]) + (3,)
        return [s, s, s]

    def perform(self, node, inputs, outputs):
        b = _np(inputs[0])
        r = _np(np.broadcast_to(inputs[1], b.shape))
        s = np.empty(b.shape + (3,))
        dsdb, dsdr = np.empty_like(s), np.empty_like(s)
        _lib.check(_lib.require_cuda().xo_qsv_host(_p(b), _p(r), _p(s), _p(dsdb), _p(dsdr), b.size, _device()), "qsv_host")
        outputs[0][0], outputs[1][0], outputs[2][0] = s, dsdb, dsdr

    def grad(self, inputs, gout):
        b, r = inputs
        s, dsdb, dsdr = self(b, r)
        if _disconnected(gout[0]):
            return [pt.zeros_like(b), pt.zeros_like(r)]
        bs = gout[0]
        return [pt.sum(bs * dsdb, axis=-1), pt.sum(bs * dsdr, axis=-1)]


class ContactPoints(Op):
    """7 same-shape inputs -> (M_left, M_right, flag:int32); no gradient (keplerian.py:753-769)."""

    __props__ = ()

    def make_node(self, *args):
        args = _bcast(*args)
        return Apply(self, args, [args[0].type(), args[0].type(), pt.TensorType("int32", shape=args[0].type.shape)()])

    def infer_shape(self, fgraph, node, shapes):
        return [shapes[0]] * 3

    def perform(self, node, inputs, outputs):
        arrs = [_np(x) for x in np.broadcast_arrays(*inputs)]
        Ml, Mr = np.empty_like(arrs[0]), np.empty_like(arrs[0])
        flag = np.empty(arrs[0].shape, dtype=np.int32)
        _lib.check(_lib.require_cuda().xo_contact_points_host(*[_p(a) for a in arrs], _p(Ml), _p(Mr), _p(flag), arrs[0].size, _device()),
                   "contact_points_host")
        outputs[0][0], outputs[1][0], outputs[2][0] = Ml, Mr, flag

    def grad(self, inputs, gout):
        return [pytensor.gradient.grad_undefined(self, i, x) for i, x in enumerate(inputs)]


_kepler, _qsv, _contact = Kepler(), QuadSolutionVector(), ContactPoints()


def kepler(M, e):
    return _kepler(M, e)


def quad_solution_vector(b, r):
    return _qsv(b, r)[0]


def contact_points(a, e, cosw, sinw, cosi, sini, L):
    return _contact(a, e, cosw, sinw, cosi, sini, L)


# Models are registered by name so that the fused Ops hash, compare and pickle by that name (PyTensor merges equal
# Ops and pickles compiled functions); the model itself (device tensors) is looked up at run time.
_MODELS = {}


def register_model(name, model):
    """make ``model`` (a fused.TransitModel) available to ``TransitLogpOp(name)``"""
    _MODELS[str(name)] = model
    return str(name)


def _model(name):
    try:
        return _MODELS[name]
    except KeyError:
        raise KeyError("no transit model registered under %r in this process: call pytensor_ops.register_model first" % name) from None


class _TransitLogpGrad(Op):
    """(par[C,P,8], star[C,4], mean[C]) -> (logp[C], gpar, gstar, gmean) in one device call."""

    __props__ = ("name",)

    def __init__(self, name):
        self.name = str(name)

    def make_node(self, par, star, mean):
        par, star, mean = _f64(par), _f64(star), _f64(mean)
        return Apply(self, [par, star, mean], [mean.type(), par.type(), star.type(), mean.type()])

    def infer_shape(self, fgraph, node, shapes):
        return [shapes[2], shapes[0], shapes[1], shapes[2]]

    def perform(self, node, inputs, outputs):
        import torch

        model = _model(self.name)
        dev = model.data.device
        par, star, mean = [torch.as_tensor(_np(x), device=dev) for x in inputs]
        res = model.logp_grad(par, star, mean)
        for slot, val in zip(outputs, res):
            slot[0] = val.cpu().numpy()

    def grad(self, inputs, gout):  # second derivatives of the likelihood are not provided
        return [pytensor.gradient.grad_undefined(self, i, x, "the fused transit likelihood provides first derivatives only")
                for i, x in enumerate(inputs)]


class TransitLogpOp(Op):
    """Fused Gaussian transit log-likelihood: data on the device, parameters in, logp[C] out;
    ``grad`` re-uses the gradients the same device call produced.  ``model``: a fused.TransitModel or the name it
    was registered under."""

    __props__ = ("name",)

    def __init__(self, model, name=None):
        if isinstance(model, str):
            self.name = model
            _model(model)
        else:
            self.name = register_model(name if name is not None else "transit_model_%d" % len(_MODELS), model)
        self._both = _TransitLogpGrad(self.name)

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._both = _TransitLogpGrad(self.name)

    def make_node(self, par, star, mean):
        par, star, mean = _f64(par), _f64(star), _f64(mean)
        return Apply(self, [par, star, mean], [mean.type()])

    def infer_shape(self, fgraph, node, shapes):
        return [shapes[2]]

    def perform(self, node, inputs, outputs):
        tmp = [[None], [None], [None], [None]]
        self._both.perform(node, inputs, tmp)
        outputs[0][0] = tmp[0][0]

    def grad(self, inputs, gout):
        _, gpar, gstar, gmean = self._both(*inputs)
        g = gout[0]
        return [g[:, None, None] * gpar, g[:, None] * gstar, g * gmean]
