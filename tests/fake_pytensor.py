"""TEST INFRASTRUCTURE: a minimal eager stand-in for the slice of PyTensor that exoplanet_b200/pytensor_ops.py touches
(``Op`` / ``Apply`` / a numpy-backed tensor variable), so that the module's ``make_node``, ``perform``, ``infer_shape``
and ``grad`` code paths execute in an image where PyTensor cannot be installed.  Graphs are evaluated eagerly: calling
an Op runs ``make_node`` and then ``perform`` on the inputs' values.  It checks the wrappers' own logic and the numbers
they return; it does not check PyTensor's graph machinery (rewrites, merging, C linker)."""
import sys
import types

import numpy as np


class TensorType:
    def __init__(self, dtype, shape=()):
        self.dtype, self.shape = dtype, tuple(shape)
        self.ndim = len(self.shape)

    def __call__(self):
        return Var(np.zeros((0,) * self.ndim, dtype=self.dtype))

    def __eq__(self, other):
        return isinstance(other, TensorType) and (self.dtype, self.ndim) == (other.dtype, other.ndim)

    def __hash__(self):
        return hash((self.dtype, self.ndim))


class Var(np.ndarray):
    def __new__(cls, a):
        return np.asarray(a).view(cls)

    @property
    def type(self):
        return TensorType(str(self.dtype), (None,) * self.ndim)

    def astype(self, dtype, **kw):
        return Var(np.asarray(self).astype(dtype))


class DisconnectedType:
    pass


class _Disconnected:
    type = DisconnectedType()


class Apply:
    def __init__(self, op, inputs, outputs):
        self.op, self.inputs, self.outputs = op, list(inputs), list(outputs)


class Op:
    __props__ = ()

    def __call__(self, *inputs):
        node = self.make_node(*inputs)
        storage = [[None] for _ in node.outputs]
        self.perform(node, [np.asarray(v) for v in node.inputs], storage)
        outs = [Var(s[0]) for s in storage]
        shapes = self.infer_shape(None, node, [np.shape(v) for v in node.inputs])
        assert [tuple(s) for s in shapes] == [o.shape for o in outs], (shapes, [o.shape for o in outs])
        for o, proto in zip(outs, node.outputs):
            assert o.ndim == proto.type.ndim and str(o.dtype) == proto.type.dtype, (o.dtype, o.ndim, proto.type.dtype, proto.type.ndim)
        return outs[0] if len(outs) == 1 else outs

    def _props(self):
        return tuple(getattr(self, k) for k in self.__props__)

    def __eq__(self, other):
        return type(self) is type(other) and self._props() == other._props()

    def __hash__(self):
        return hash((type(self), self._props()))


def install():
    """register the stand-in as ``pytensor`` (only when the real one is absent) and return the module"""
    try:
        import pytensor  # noqa: F401

        return None
    except ImportError:
        pass
    pytensor = types.ModuleType("pytensor")
    pt = types.ModuleType("pytensor.tensor")
    pt.as_tensor_variable = lambda x, **kw: x if isinstance(x, Var) else Var(np.asarray(x))
    pt.TensorType = TensorType
    pt.zeros_like = lambda x: Var(np.zeros_like(np.asarray(x)))
    pt.sum = lambda x, axis=None: Var(np.sum(np.asarray(x), axis=axis))
    pt.broadcast_arrays = lambda *xs: [Var(np.array(a)) for a in np.broadcast_arrays(*[np.asarray(x) for x in xs])]
    grad = types.ModuleType("pytensor.gradient")
    grad.DisconnectedType = DisconnectedType
    grad.disconnected = _Disconnected()

    class _Undefined:
        def __init__(self, op, i, x, comment=""):
            self.op, self.index, self.comment = op, i, comment

    grad.grad_undefined = _Undefined
    graph = types.ModuleType("pytensor.graph")
    basic = types.ModuleType("pytensor.graph.basic")
    basic.Apply = Apply
    opmod = types.ModuleType("pytensor.graph.op")
    opmod.Op = Op
    pytensor.tensor, pytensor.gradient, pytensor.graph = pt, grad, graph
    graph.basic, graph.op = basic, opmod
    pytensor.__fake__ = True
    sys.modules.update({"pytensor": pytensor, "pytensor.tensor": pt, "pytensor.gradient": grad, "pytensor.graph": graph,
                        "pytensor.graph.basic": basic, "pytensor.graph.op": opmod})
    return pytensor
