"""exoplanet_b200/pytensor_ops.py (the drop-in for ``exoplanet_core.pymc.ops``, /root/reference/src/exoplanet/compat.py:56)
executed on tests/fake_pytensor.py where the real PyTensor is absent: Op construction, broadcasting, shapes,
equality / pickling on the CPU; values and gradients against the oracle on the GPU; the ``shim/exoplanet_core`` package."""
import importlib
import os
import pickle
import sys

import numpy as np
import pytest

import fake_pytensor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def xops():
    fake_pytensor.install()
    sys.modules.pop("exoplanet_b200.pytensor_ops", None)
    return importlib.import_module("exoplanet_b200.pytensor_ops")


def test_ops_construct_broadcast_and_compare(xops):
    M = np.linspace(0, 6, 7)
    node = xops.Kepler().make_node(M, 0.3)  # a scalar e against an array M: broadcast in make_node
    assert [np.shape(v) for v in node.inputs] == [(7,), (7,)] and len(node.outputs) == 2
    node = xops.QuadSolutionVector().make_node(np.zeros((4, 3)), 0.1)
    assert node.outputs[0].type.ndim == 3 and xops.QuadSolutionVector().infer_shape(None, node, [(4, 3), (4, 3)]) == [(4, 3, 3)] * 3
    node = xops.ContactPoints().make_node(*[np.ones(2)] * 6, 1.0)
    assert node.outputs[2].type.dtype == "int32" and all(np.shape(v) == (2,) for v in node.inputs)
    assert xops.Kepler() == xops.Kepler() and hash(xops.Kepler()) == hash(xops.Kepler())
    # fused Op: identified by the registered name, not by id(model); survives pickling
    name = xops.register_model("unit-test-model", object())
    a, b = xops.TransitLogpOp(name), xops.TransitLogpOp(name)
    assert a == b and hash(a) == hash(b) and a.name == "unit-test-model"
    c = pickle.loads(pickle.dumps(a))
    assert c == a and c._both == a._both
    with pytest.raises(KeyError):
        xops.TransitLogpOp("never registered")


def test_shim_package_exposes_the_three_ops(xops):
    sys.path.insert(0, os.path.join(ROOT, "shim"))
    try:
        for m in [k for k in sys.modules if k.startswith("exoplanet_core")]:
            sys.modules.pop(m)
        ops = importlib.import_module("exoplanet_core.pymc").ops  # what compat.py:56 does
        assert ops.kepler is xops.kepler and ops.quad_solution_vector is xops.quad_solution_vector
        assert ops.contact_points is xops.contact_points
    finally:
        sys.path.remove(os.path.join(ROOT, "shim"))


@pytest.mark.gpu
def test_op_values_and_gradients_match_the_oracle(xops):
    import xo_oracle as O
    from fake_pytensor import Var

    rng = np.random.default_rng(3)
    M, e = rng.uniform(-20, 20, (5, 40)), rng.uniform(0, 0.9, (5, 40))
    sf, cf = xops.kepler(M, e)
    so, co = O.kepler(M, e)
    assert np.max(np.abs(sf - so)) < 1e-13 and np.max(np.abs(cf - co)) < 1e-13
    gs, gc = rng.normal(size=M.shape), rng.normal(size=M.shape)
    gM, ge = xops.Kepler().grad([Var(M), Var(e)], [Var(gs), Var(gc)])
    bM, be = O.kepler_vjp(so, co, e, gs, gc)
    assert np.max(np.abs(gM - bM)) < 1e-10 * np.abs(bM).max() and np.max(np.abs(ge - be)) < 1e-10 * np.abs(be).max()
    b, r = rng.uniform(0, 1.3, 500), rng.uniform(0.01, 0.2, 500)
    s = xops.quad_solution_vector(b, r)
    s_o, db_o, dr_o = O.quad_solution_vector(b, r)
    assert s.shape == (500, 3) and np.max(np.abs(s - s_o)) < 1e-13
    bs = rng.normal(size=(500, 3))
    gb, gr = xops.QuadSolutionVector().grad([Var(b), Var(r)], [Var(bs), None, None])
    far = np.abs(b - (1 - r)) > 1e-3  # away from the square-root singular contact curve
    assert np.max(np.abs(gb - (bs * db_o).sum(-1))[far]) < 1e-9 and np.max(np.abs(gr - (bs * dr_o).sum(-1))[far]) < 1e-9
    args = [np.array([10.0, 8.0]), np.array([0.1, 0.6]), np.array([0.87, 0.2]), np.array([0.48, -0.98]), np.array([0.02, 0.05]),
            np.array([0.9998, 0.9987]), np.array([1.1, 1.05])]
    Ml, Mr, flag = xops.contact_points(*args)
    Mlo, Mro, fo = O.contact_points(*args)
    assert flag.dtype == np.int32 and np.array_equal(flag, fo) and np.allclose(Ml, Mlo, atol=1e-10) and np.allclose(Mr, Mro, atol=1e-10)


@pytest.mark.gpu
def test_fused_transit_op_perform_and_grad(xops):
    import torch
    import xo_oracle as O
    from fake_pytensor import Var

    from exoplanet_b200 import fused, synthetic

    N, C, P = 20000, 3, 2
    p = synthetic.transit_params(C, P, seed=5)
    t = synthetic.cadence_times(N)
    data = fused.TransitData(t, synthetic.noise(N), np.full(N, 5e-4), device="cuda")
    op = xops.TransitLogpOp(fused.TransitModel(data), name="gpu-test-model")
    par = torch.stack([torch.as_tensor(p[k]) for k in ("period", "t0", "ecc", "omega", "b", "ror", "m_planet", "K")], dim=-1).numpy()
    star = np.stack([np.ones(C), np.ones(C), p["u1"], p["u2"]], axis=-1)
    logp = op(par, star, p["mean"])
    el = O.elements_from_params(p["period"], p["t0"], p["ecc"], p["omega"], p["b"], p["ror"], m_planet=p["m_planet"])
    lo = O.transit_logp_grad(t, synthetic.noise(N), np.full(N, 1 / 5e-4 ** 2), el, O.get_cl(p["u1"], p["u2"]), p["mean"], np.zeros(1),
                             np.ones(1), lognorm=data.lognorm)[0]
    assert logp.shape == (C,) and np.max(np.abs(logp - lo) / np.abs(lo)) < 1e-11
    g = op.grad([Var(par), Var(star), Var(p["mean"])], [Var(np.array([1.0, 2.0, 0.5]))])
    ref = fused.TransitModel(data).logp_grad(torch.as_tensor(par).cuda(), torch.as_tensor(star).cuda(), torch.as_tensor(p["mean"]).cuda())
    assert np.allclose(g[0], np.array([1.0, 2.0, 0.5])[:, None, None] * ref[1].cpu().numpy(), rtol=1e-13)
    assert np.allclose(g[2], np.array([1.0, 2.0, 0.5]) * ref[3].cpu().numpy(), rtol=1e-13)
