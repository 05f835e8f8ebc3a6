"""GPU parity tests for the three reference ops, through the C ABI (ctypes) and the
torch op layer.  Comparator: the CPU oracle + committed golden vectors."""
import numpy as np
import pytest
import torch

import xo_oracle as O
from exoplanet_b200 import ops, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=DEV)


def test_kepler_golden(golden):
    g = golden("kepler")
    R = np.array(g["rows"])
    sf, cf = ops.kepler(dev(R[:, 0]), dev(R[:, 1]))
    # tolerance stated by BASELINE.md section 6: 1e-12; achieved ~1e-15
    assert np.max(np.abs(sf.cpu().numpy() - R[:, 2])) < 5e-15
    assert np.max(np.abs(cf.cpu().numpy() - R[:, 3])) < 5e-15
    for M, e, E, s, c in g["known_answers"]:
        sf, cf = ops.kepler(dev([M]), dev([e]))
        assert abs(sf.item() - s) < 2e-15 and abs(cf.item() - c) < 2e-15


def test_kepler_vs_oracle_and_residual():
    n = 1 << 20
    M, e = synthetic.kepler_inputs(n, e_mode="uniform")
    sf, cf = ops.kepler(dev(M), dev(e))
    sf, cf = sf.cpu().numpy(), cf.cpu().numpy()
    so, co = O.kepler(M, e)
    assert np.max(np.abs(sf - so)) < 5e-15 and np.max(np.abs(cf - co)) < 5e-15
    # Kepler residual |M - E + e sin E| < 1e-12 with E rebuilt from the outputs (BASELINE.json)
    den = 1 + e * cf
    cE, sE = (e + cf) / den, np.sqrt(1 - e * e) * sf / den
    E = np.arctan2(sE, cE)
    k = np.rint(M / (2 * np.pi))
    Mr = (M - k * 6.283185307179586232) - k * 2.449293598294706414e-16
    res = np.abs((Mr - E + e * sE + np.pi) % (2 * np.pi) - np.pi)
    assert res.max() < 1e-12


@pytest.mark.parametrize("ecc", [0.0, 0.1, 0.5, 0.9, 0.99])
def test_kepler_fixed_e(ecc):
    M, e = synthetic.kepler_inputs(1 << 16, e_mode="fixed", e_value=ecc)
    sf, cf = ops.kepler(dev(M), dev(e))
    so, co = O.kepler(M, e)
    tol = 5e-15 if ecc < 0.95 else 2e-13
    assert np.max(np.abs(sf.cpu().numpy() - so)) < tol and np.max(np.abs(cf.cpu().numpy() - co)) < tol


def test_kepler_edge_cases():
    # empty, odd length, unaligned views, broadcasting of e, NaN in -> NaN out
    sf, cf = ops.kepler(dev(np.zeros(0)), dev(np.zeros(0)))
    assert sf.numel() == 0
    M = np.linspace(-7, 9, 1001)
    base = dev(np.concatenate([[0.0], M]))
    view = base[1:]  # 8-byte aligned only
    sf, cf = ops.kepler(view, torch.tensor(0.3, dtype=torch.float64, device=DEV))
    so, co = O.kepler(M, np.full_like(M, 0.3))
    assert np.max(np.abs(sf.cpu().numpy() - so)) < 5e-15
    sf, cf = ops.kepler(dev([np.nan, 1.0]), dev([0.1, 0.2]))
    assert np.isnan(sf[0].item()) and np.isfinite(sf[1].item())
    # huge |M|: reduction stays accurate
    M = np.array([1e5 + 0.123, -3.3e5, 6283.2])
    sf, cf = ops.kepler(dev(M), dev(np.full(3, 0.4)))
    so, co = O.kepler(M, np.full(3, 0.4), long_double=True)
    assert np.max(np.abs(sf.cpu().numpy() - so)) < 1e-13


def test_kepler_autograd_matches_oracle_vjp_and_fd():
    rng = np.random.default_rng(3)
    M = rng.uniform(-10, 10, 257)
    e = rng.uniform(0.0, 0.9, 257)
    Mt = dev(M).requires_grad_(True)
    et = dev(e).requires_grad_(True)
    sf, cf = ops.kepler(Mt, et)
    bs, bc = rng.normal(size=257), rng.normal(size=257)
    (sf * dev(bs) + cf * dev(bc)).sum().backward()
    bM, be = O.kepler_vjp(sf.detach().cpu().numpy(), cf.detach().cpu().numpy(), e, bs, bc)
    assert np.allclose(Mt.grad.cpu().numpy(), bM, rtol=1e-12, atol=1e-12)
    assert np.allclose(et.grad.cpu().numpy(), be, rtol=1e-12, atol=1e-12)
    assert torch.autograd.gradcheck(lambda a, b: ops.kepler(a, b), (dev(M[:8]).requires_grad_(True), dev(e[:8]).requires_grad_(True)),
                                    eps=1e-6, atol=1e-6, rtol=1e-6)


def test_qsv_golden(golden):
    R = np.array(golden("qsv")["rows"])
    b, r = R[:, 0], R[:, 1]
    bt = dev(b).requires_grad_(True)
    rt = dev(r).requires_grad_(True)
    s = ops.quad_solution_vector(bt, rt)
    assert s.shape == (len(b), 3)
    assert np.max(np.abs(s.detach().cpu().numpy() - R[:, 2:5])) < 1e-14
    contact = (np.abs(np.abs(b) + r - 1) < 1e-7) | (np.abs(np.abs(b) - r - 1) < 1e-7)
    for k in range(3):
        gb, gr = torch.autograd.grad(s[:, k].sum(), (bt, rt), retain_graph=True)
        assert np.max(np.abs(gb.cpu().numpy() - R[:, 5 + k])[~contact]) < 1e-13
        assert np.max(np.abs(gr.cpu().numpy() - R[:, 8 + k])[~contact]) < 1e-13


def test_qsv_vs_oracle_dense_and_edges():
    rng = np.random.default_rng(4)
    n = 200001  # odd, not a multiple of the tile
    r = np.concatenate([10 ** rng.uniform(-4, 0, n // 2), rng.uniform(0, 1.5, n - n // 2)])
    b = rng.uniform(-1.0, 1.0, n) * (1 + r) * 1.05
    s = ops.quad_solution_vector(dev(b), dev(r)).cpu().numpy()
    so, _, _ = O.quad_solution_vector(b, r)
    assert np.max(np.abs(s - so)) < 1e-13
    # unocculted / fully occulted / r = 0 exact values
    s = ops.quad_solution_vector(dev([1.5, 0.2, 0.3]), dev([0.1, 1.5, 0.0])).cpu().numpy()
    assert np.array_equal(s[0], [np.pi, 2 * np.pi / 3, 0.0]) and np.array_equal(s[1], [0, 0, 0]) and np.array_equal(s[2], s[0])
    assert ops.quad_solution_vector(dev(np.zeros(0)), dev(np.zeros(0))).shape == (0, 3)


def test_qsv_singular_points_are_continuous():
    # the reference's continuity test, tests/light_curves_test.py:220-254
    eps = 1e-8
    pts = [(0.9, 0.1), (0.5, 0.5), (0.0, 0.1), (0.0, 1.0), (1.1, 0.1)]
    cl = O.get_cl(np.array(0.2), np.array(0.3))
    for b0, r0 in pts:
        for db in (-eps, 0.0, eps):
            for dr in (-eps, 0.0, eps):
                b, r = abs(b0 + db), r0 + dr
                s = ops.quad_solution_vector(dev([b]), dev([r])).cpu().numpy()[0]
                s0 = ops.quad_solution_vector(dev([b0]), dev([r0])).cpu().numpy()[0]
                assert abs((s - s0) @ cl) < 1e-6


def test_contact_points_golden(golden):
    R = np.array(golden("contact")["rows"])
    Ml, Mr, flag = ops.contact_points(*[dev(R[:, k]) for k in range(7)])
    assert np.all(flag.cpu().numpy() == R[:, 9].astype(int))
    twopi = 2 * np.pi
    dl = np.abs((Ml.cpu().numpy() - R[:, 7] + np.pi) % twopi - np.pi)
    dr = np.abs((Mr.cpu().numpy() - R[:, 8] + np.pi) % twopi - np.pi)
    assert dl.max() < 1e-12 and dr.max() < 1e-12
    # no-transit geometry raises the flag (keplerian.py:753,771-775)
    _, _, fl = ops.contact_points(dev([10.0]), dev([0.1]), dev([1.0]), dev([0.0]), dev([0.5]), dev([np.sqrt(0.75)]), dev([1.1]))
    assert fl.item() != 0


def test_host_buffer_entry_points_match_device_path(golden):
    R = np.array(golden("kepler")["rows"])
    sf, cf = ops.kepler(R[:, 0], R[:, 1])  # numpy in -> xo_kepler_host
    assert isinstance(sf, np.ndarray)
    assert np.max(np.abs(sf - R[:, 2])) < 5e-15 and np.max(np.abs(cf - R[:, 3])) < 5e-15
    Q = np.array(golden("qsv")["rows"])
    s, dsdb, dsdr = ops.quad_solution_vector_with_jac(Q[:, 0], Q[:, 1])
    sd = ops.quad_solution_vector(dev(Q[:, 0]), dev(Q[:, 1])).cpu().numpy()
    assert np.array_equal(s, sd)
    Cn = np.array(golden("contact")["rows"])
    Ml, Mr, flag = ops.contact_points(*[Cn[:, k] for k in range(7)])
    Mld, _, _ = ops.contact_points(*[dev(Cn[:, k]) for k in range(7)])
    assert np.array_equal(Ml, Mld.cpu().numpy())


def test_kepler_broadcast_e_entry_point():
    """xo_kepler_fwd_bcast: one e per `inner` consecutive M (per-planet scalar e, keplerian.py:333)."""
    import ctypes as C

    from exoplanet_b200 import _lib

    lib = _lib.require_cuda()
    rng = np.random.default_rng(2)
    inner, blocks = 1001, 7
    M = rng.uniform(-50, 50, inner * blocks)
    e = rng.uniform(0, 0.9, blocks)
    Md, ed = dev(M), dev(e)
    sf, cf = torch.empty_like(Md), torch.empty_like(Md)
    p = lambda x: C.c_void_p(x.data_ptr())
    assert lib.xo_kepler_fwd_bcast(p(Md), p(ed), inner, p(sf), p(cf), Md.numel(), None) == 0
    torch.cuda.synchronize()
    so, co = O.kepler(M, np.repeat(e, inner))
    assert np.max(np.abs(sf.cpu().numpy() - so)) < 5e-15 and np.max(np.abs(cf.cpu().numpy() - co)) < 5e-15


def test_error_codes_on_device():
    import ctypes as C

    from exoplanet_b200 import _lib

    lib = _lib.require_cuda()
    x = dev(np.zeros(8))
    p = C.c_void_p(x.data_ptr())
    assert lib.xo_transit_logp_grad(p, p, p, 8, p, p, p, 1, 1, p, p, 1, 0.0, None, p, p, p, p, p, 8, None) == 3  # XO_EWORKSPACE
    assert lib.xo_transit_lc_fwd(p, 8, p, p, 1, 1, p, p, 99, p, None) == 2  # stencil above XO_MAX_STENCIL
    with pytest.raises(_lib.XoError):
        _lib.check(3, "demo")


def test_cpu_torch_tensors_keep_their_gradients():
    """a CPU tensor that requires grad must not be detached into the numpy path: it is moved to the device,
    goes through the autograd Function and the result comes back on the CPU with its graph"""
    M = torch.linspace(-3.0, 3.0, 50, dtype=torch.float64, requires_grad=True)
    e = torch.full((50,), 0.3, dtype=torch.float64, requires_grad=True)
    sf, cf = ops.kepler(M, e)
    assert sf.device.type == "cpu" and sf.requires_grad
    gM, ge = torch.autograd.grad((sf * 0.7 + cf * 0.2).sum(), [M, e])
    so, co = O.kepler(M.detach().numpy(), e.detach().numpy())
    bM, be = O.kepler_vjp(so, co, e.detach().numpy(), np.full(50, 0.7), np.full(50, 0.2))
    assert np.allclose(gM.numpy(), bM, rtol=1e-10, atol=1e-12) and np.allclose(ge.numpy(), be, rtol=1e-10, atol=1e-12)
    b = torch.linspace(0.0, 1.2, 40, dtype=torch.float64, requires_grad=True)
    s = ops.quad_solution_vector(b, torch.tensor(0.1, dtype=torch.float64))
    assert s.device.type == "cpu" and s.shape == (40, 3) and s.requires_grad
    # numpy in -> numpy out is unchanged
    sn, cn = ops.kepler(M.detach().numpy(), 0.3)
    assert isinstance(sn, np.ndarray) and np.allclose(sn, so, atol=1e-13)
