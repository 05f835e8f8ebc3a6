"""GPU parity tests of the fused L3 kernels against the CPU oracle on identical seeded inputs."""
import numpy as np
import pytest
import torch

import xo_oracle as O
from exoplanet_b200 import elements, fused, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=DEV)


def problem(N, C, P, ecc=None, texp=None, seed=11, use_in_transit=True):
    p = synthetic.transit_params(C, P, seed=seed, ecc=ecc)
    t = synthetic.cadence_times(N)
    el_o = O.elements_from_params(p["period"], p["t0"], p["ecc"], p["omega"], p["b"], p["ror"], texp=texp, K=p["K"],
                                  use_in_transit=use_in_transit)
    el_g = elements.elements_from_params(dev(p["period"]), dev(p["t0"]), dev(p["ecc"]), dev(p["omega"]), dev(p["b"]),
                                         dev(p["ror"]), K=dev(p["K"]), texp=texp, use_in_transit=use_in_transit)
    cl = O.get_cl(p["u1"], p["u2"])
    return p, t, el_o, el_g, cl


def test_prologue_elements_match_oracle():
    p, t, el_o, el_g, cl = problem(100, 16, 3)
    eg = el_g.cpu().numpy()
    assert np.allclose(eg[..., :11], el_o[..., :11], rtol=1e-13, atol=1e-13)
    assert np.allclose(eg[..., 11:13], el_o[..., 11:13], rtol=0, atol=1e-10)  # contact times
    clg = fused.limb_dark_coeffs(dev(p["u1"]), dev(p["u2"])).cpu().numpy()
    assert np.allclose(clg, cl, rtol=1e-14, atol=1e-15)


@pytest.mark.parametrize("P,texp,order", [(1, None, 0), (2, 0.02, 0), (3, 0.0204, 2), (1, 0.02, 1)])
def test_light_curve_matches_oracle(P, texp, order):
    N, C = 20000, 4
    p, t, el_o, el_g, cl = problem(N, C, P, texp=texp)
    st = O.stencil(texp, 7, order)
    lc_o = O.transit_lc(t, el_o, cl, *st)
    lc = fused.light_curve(dev(t), dev(el_o), dev(cl), fused.exposure_stencil(texp, 7, order)).cpu().numpy()
    assert np.count_nonzero(lc_o) > 50
    # the oracle evaluates every planet where ANY planet is in window (reference semantics);
    # outside a planet's own window its column is exactly 0 in both unless b < 1 + r there
    assert np.max(np.abs(lc - lc_o)) < 1e-14
    assert np.array_equal(lc == 0.0, lc_o == 0.0)


def test_light_curve_exact_zeros_and_use_in_transit_equivalence():
    # tests/light_curves_test.py:75-102: use_in_transit True == False
    N = 30000
    p, t, el_o, el_g, cl = problem(N, 3, 2, ecc=0.5)
    p2, _, el_all, _, _ = problem(N, 3, 2, ecc=0.5, use_in_transit=False)
    a = fused.light_curve(dev(t), dev(el_o), dev(cl)).cpu().numpy()
    b = fused.light_curve(dev(t), dev(el_all), dev(cl)).cpu().numpy()
    assert np.array_equal(a, b)


@pytest.mark.parametrize("P,texp,C", [(1, None, 5), (1, 0.02, 3), (3, None, 4), (2, 0.0204, 2)])
def test_logp_grad_matches_oracle(P, texp, C):
    N = 50000
    p, t, el_o, el_g, cl = problem(N, C, P, texp=texp, seed=5 + P)
    if P > 1:  # force overlapping transits of two planets in chain 0
        p["period"][0, 1] = p["period"][0, 0]
        p["t0"][0, 1] = p["t0"][0, 0] + 0.01
        el_o = O.elements_from_params(p["period"], p["t0"], p["ecc"], p["omega"], p["b"], p["ror"], texp=texp, K=p["K"])
    st = O.stencil(texp, 7, 0)
    truth = O.transit_lc(t, el_o, cl, *st).sum(-1)
    y = truth[0] + synthetic.noise(N)  # every chain sees the same data: chain 0's truth + noise
    sigma = 5e-4
    data = fused.TransitData(t, y, np.full(N, sigma), device=DEV)
    logp, gel, gcl, gmean = fused.transit_logp_grad(data, dev(el_o), dev(cl), dev(p["mean"]), fused.exposure_stencil(texp))
    lo, geo, gco, gmo = O.transit_logp_grad(t, y, np.full(N, 1 / sigma ** 2), el_o, cl, p["mean"], *st, lognorm=data.lognorm)
    # logp rel err <= 1e-10, gradient <= 1e-8
    assert np.max(np.abs(logp.cpu().numpy() - lo) / np.abs(lo)) < 1e-12
    scale = np.abs(geo).max(axis=(0, 1), keepdims=True) + 1e-300
    assert np.max(np.abs(gel.cpu().numpy() - geo) / scale) < 1e-10
    assert np.max(np.abs(gcl.cpu().numpy() - gco) / (np.abs(gco).max() + 1e-300)) < 1e-10
    assert np.max(np.abs(gmean.cpu().numpy() - gmo) / np.abs(gmo).max()) < 1e-10


@pytest.mark.parametrize("N", [1, 31, 1000, 100000, 1234567])
def test_data_moments_kernel(N):
    """one-pass multi-CTA moments: y_ref inside the data's range, shifted sums vs numpy (long double),
    sortedness flag, bit-reproducible (the last-arriving CTA folds the partials in CTA order)"""
    rng = np.random.default_rng(N)
    t = np.cumsum(rng.uniform(0.01, 0.03, N))
    y = 1.0 + 5e-4 * rng.normal(size=N)
    ivar = rng.uniform(0.5, 2.0, N) / 5e-4 ** 2
    d = fused.TransitData(t, y, 1 / np.sqrt(ivar), device=DEV)
    m = d.moments().cpu().numpy()
    iv = d.ivar.cpu().numpy().astype(np.longdouble)
    assert y.min() <= m[0] <= y.max()
    dy = y.astype(np.longdouble) - np.longdouble(m[0])
    ref = [iv.sum(), (iv * dy).sum(), (iv * dy * dy).sum()]
    for k in range(3):
        assert abs(m[1 + k] - float(ref[k])) <= 1e-13 * float((iv * np.abs(dy) ** k).sum()), (k, m[1 + k], ref[k])
    assert m[4] == 1.0 and m[7] == 0.0
    # the time index behind the moments: table[j] = #{t < t[0] + j w}, only for sorted, non-degenerate time stamps
    nb = int(m[5])
    if N > 1:
        assert nb == max(N // 2, 1) and len(m) >= 1024 + (nb + 2) // 2
        w = (t[-1] - t[0]) / nb
        assert abs(m[6] * w - 1.0) < 1e-12
        table = d.moments().cpu().numpy()[1024:].view(np.uint32)[: nb + 1]
        grid = np.array([np.float64(j) * w + t[0] for j in range(nb)]) if nb < 2000 else t[0] + np.arange(nb) * w
        ref_tab = np.searchsorted(t, grid, side="left")
        # a grid point within rounding of a time stamp may fall on either side: compare where the bin edge is well separated
        sep = np.min(np.abs(grid[:, None] - t[None, np.clip(ref_tab, 0, N - 1)]), axis=1) > 1e-9 if nb < 2000 else np.ones(nb, bool)
        assert np.array_equal(table[:nb][sep], ref_tab[sep]) or nb >= 2000
        assert table[nb] == N and np.all(np.diff(table.astype(np.int64)) >= 0)
    else:
        assert nb == 0
    assert np.all(fused.TransitData(t, y, 1 / np.sqrt(ivar), device=DEV).moments(index=False).cpu().numpy()[5:8] == 0.0)
    d2 = fused.TransitData(t, y, 1 / np.sqrt(ivar), device=DEV)
    assert np.array_equal(d2.moments().cpu().numpy(), d.moments().cpu().numpy())
    if N > 2:
        tu = t.copy()
        tu[[N // 2, N // 2 + 1]] = tu[[N // 2 + 1, N // 2]]
        mu = fused.TransitData(tu, y, 1 / np.sqrt(ivar), device=DEV).moments().cpu().numpy()
        assert mu[4] == 0.0 and mu[5] == 0.0  # unsorted: no index, the stepping sweep reads every time stamp


def test_moments_path_equals_full_sweep():
    """closed-form out-of-transit sums (data moments) vs summing every cadence in the sweep"""
    N, C = 100000, 9
    p, t, el_o, el_g, cl = problem(N, C, 2, seed=31)
    y = 1.0 + synthetic.noise(N)  # a large offset: the moments are taken about y_ref, so no cancellation
    data = fused.TransitData(t, y, np.full(N, 5e-4), device=DEV)
    mean = dev(1.0 + p["mean"])
    a = fused.transit_logp_grad(data, dev(el_o), dev(cl), mean, use_moments=True)
    b = fused.transit_logp_grad(data, dev(el_o), dev(cl), mean, use_moments=False)
    assert torch.allclose(a[0], b[0], rtol=1e-12, atol=0)
    assert torch.allclose(a[3], b[3], rtol=1e-9, atol=1e-9 * b[3].abs().max().item())
    # element / limb-darkening gradients: the same samples, enumerated in a different order (time index vs stepping)
    assert torch.allclose(a[1], b[1], rtol=1e-11, atol=1e-11 * b[1].abs().max().item())
    assert torch.allclose(a[2], b[2], rtol=1e-11, atol=1e-11 * b[2].abs().max().item())
    # without the index the stepping sweep over the same moments visits the same samples in the same order as ... itself
    d2 = fused.TransitData(t, y, np.full(N, 5e-4), device=DEV)
    d2.moments(index=False)
    c = fused.transit_logp_grad(d2, dev(el_o), dev(cl), mean, use_moments=True)
    assert torch.allclose(a[0], c[0], rtol=1e-12, atol=0) and torch.allclose(a[1], c[1], rtol=1e-11, atol=1e-11 * b[1].abs().max().item())
    # bit-reproducible run to run
    a2 = fused.transit_logp_grad(data, dev(el_o), dev(cl), mean, use_moments=True)
    assert all(torch.equal(x, z) for x, z in zip(a, a2))
    m = data.moments().cpu().numpy()
    assert m[4] == 1.0  # t is sorted -> step skipping is on; an unsorted copy must give the same answer
    perm = np.random.default_rng(3).permutation(N)
    data_u = fused.TransitData(t[perm], y[perm], np.full(N, 5e-4), device=DEV)
    assert data_u.moments().cpu().numpy()[4] == 0.0
    c = fused.transit_logp_grad(data_u, dev(el_o), dev(cl), mean, use_moments=True)
    assert torch.allclose(a[0], c[0], rtol=1e-12, atol=0)
    assert torch.allclose(a[1], c[1], rtol=1e-9, atol=1e-9 * a[1].abs().max().item())
    iv = 1 / 5e-4 ** 2
    yref = m[0]  # any value inside the data's range (test_data_moments_kernel checks the sums about it)
    assert y.min() <= yref <= y.max() and abs(m[1] - iv * N) < 1e-6 * iv
    assert abs(m[3] - np.sum(iv * (y - yref) ** 2)) < 1e-9 * m[3]


def test_logp_is_deterministic_and_chain_independent():
    N, C = 100000, 64
    p, t, el_o, el_g, cl = problem(N, C, 1, ecc=0.5)
    y = synthetic.noise(N)
    data = fused.TransitData(t, y, np.full(N, 5e-4), device=DEV)
    a = fused.transit_logp_grad(data, dev(el_o), dev(cl), dev(p["mean"]))
    b = fused.transit_logp_grad(data, dev(el_o), dev(cl), dev(p["mean"]))
    for x, z in zip(a, b):
        assert torch.equal(x, z)  # bit-reproducible
    # a chain's result does not depend on its batch: rows of a sub-batch agree bitwise
    # only if the work split is equal, so compare to tolerance instead
    c = fused.transit_logp_grad(data, dev(el_o[:7]), dev(cl[:7]), dev(p["mean"][:7]))
    assert torch.allclose(a[0][:7], c[0], rtol=1e-13, atol=0)
    assert torch.allclose(a[1][:7], c[1], rtol=1e-9, atol=1e-9)
    # spot-check 3 chains of the full-size batch against the oracle
    sel = [0, 31, 63]
    lo, geo, _, _ = O.transit_logp_grad(t, y, np.full(N, 1 / 5e-4 ** 2), el_o[sel], cl[sel], p["mean"][sel], np.zeros(1),
                                        np.ones(1), lognorm=data.lognorm)
    assert np.max(np.abs(a[0].cpu().numpy()[sel] - lo) / np.abs(lo)) < 1e-12
    assert np.allclose(a[1].cpu().numpy()[sel], geo, rtol=1e-8, atol=1e-6)


def test_logp_autograd_through_prologue():
    N = 20000
    p = synthetic.transit_params(3, 1, seed=9, ecc=0.3)
    t = synthetic.cadence_times(N)
    data = fused.TransitData(t, synthetic.noise(N), np.full(N, 5e-4), device=DEV)
    names = ["period", "t0", "ecc", "omega", "b", "ror"]
    ts = {k: dev(p[k]).requires_grad_(True) for k in names}
    u1 = dev(p["u1"]).requires_grad_(True)
    u2 = dev(p["u2"]).requires_grad_(True)
    mean = dev(p["mean"]).requires_grad_(True)
    op = fused.TransitLogp(data)

    def f():
        el = elements.elements_from_params(ts["period"], ts["t0"], ts["ecc"], ts["omega"], ts["b"], ts["ror"])
        return op(el, fused.limb_dark_coeffs(u1, u2), mean).sum()

    val = f()
    grads = torch.autograd.grad(val, [ts[k] for k in names] + [u1, u2, mean])
    # central differences on a few parameters (float64, so FD is meaningful)
    for k, h in [("b", 1e-7), ("ror", 1e-8), ("ecc", 1e-8), ("omega", 1e-8)]:
        base = ts[k].detach().clone()
        with torch.no_grad():
            ts[k].copy_(base + h)
            fp = f().item()
            ts[k].copy_(base - h)
            fm = f().item()
            ts[k].copy_(base)
        fd = (fp - fm) / (2 * h)
        g = grads[names.index(k)].sum().item()
        assert abs(fd - g) <= 1e-4 * abs(fd) + 1e-2, (k, fd, g)


def test_orbit_eval_and_rv_match_oracle():
    N, C, P = 4000, 3, 5
    p, t, el_o, el_g, cl = problem(N, C, P, use_in_transit=False)
    out = fused.orbit_eval(dev(t), dev(el_o))
    Xo, Yo, Zo, rvo = O.orbit_eval(t, el_o)
    for k, ref in zip(("X", "Y", "Z", "rv"), (Xo, Yo, Zo, rvo)):
        assert np.max(np.abs(out[k].cpu().numpy() - ref) / (1 + np.abs(ref))) < 1e-13
    rv_obs = rvo[0].sum(-1) + np.random.default_rng(0).normal(size=N)
    data = fused.TransitData(t, rv_obs, np.full(N, 1.0), device=DEV)
    logp, grv, gg = fused.rv_logp_grad(data, dev(el_o), dev(p["gamma"]))
    lo, go, ggo = O.rv_logp_grad(t, rv_obs, np.ones(N), el_o, p["gamma"], lognorm=data.lognorm)
    assert np.max(np.abs(logp.cpu().numpy() - lo) / np.abs(lo)) < 1e-12
    scale = np.abs(go).max(axis=(0, 1), keepdims=True)
    assert np.max(np.abs(grv.cpu().numpy() - go) / scale) < 1e-10
    assert np.allclose(gg.cpu().numpy(), ggo, rtol=1e-10)


def test_known_answer_B6(golden):
    b6 = golden("end_to_end")["B6"]
    el = elements.elements_from_params(dev(b6["period"]), dev(0.0), dev(b6["ecc"]), dev(b6["omega"]), dev(0.0), dev(b6["r"]))
    el = el.reshape(16)
    assert abs(el[8].item() - b6["a"]) < 1e-12
    assert abs(el[11].item() - b6["t_start"]) < 1e-10 and abs(el[12].item() - b6["t_end"]) < 1e-10
    cl = fused.limb_dark_coeffs(dev(b6["u"][0]), dev(b6["u"][1])).reshape(3)
    lc = fused.light_curve(dev(b6["t"]), el[None, None, :], cl[None, :])[0, :, 0].cpu().numpy()
    assert np.max(np.abs(lc - np.array(b6["lc"]))) < 5e-15


def test_host_buffer_logp_entry_point():
    import ctypes as C

    from exoplanet_b200 import _lib

    N, Cn, P = 10000, 3, 2
    p, t, el_o, el_g, cl = problem(N, Cn, P)
    y = synthetic.noise(N)
    ivar = np.full(N, 4e6)
    dtk, wk = fused.exposure_stencil()
    logp, gel, gcl, gm = np.empty(Cn), np.empty((Cn, P, 10)), np.empty((Cn, 3)), np.empty(Cn)
    f = lambda a: a.ctypes.data_as(C.c_void_p)
    el_c = np.ascontiguousarray(el_o)
    cl_c = np.ascontiguousarray(cl)
    rc = _lib.require_cuda().xo_transit_logp_grad_host(f(t), f(y), f(ivar), N, f(el_c), f(cl_c), f(p["mean"]), Cn, P, f(dtk),
                                                       f(wk), 1, 0.0, f(logp), f(gel), f(gcl), f(gm), 0)
    assert rc == 0
    lo, geo, gco, gmo = O.transit_logp_grad(t, y, ivar, el_o, cl, p["mean"], dtk, wk)
    assert np.allclose(logp, lo, rtol=1e-12) and np.allclose(gel, geo, rtol=1e-8, atol=1e-6)
    lc = np.empty((Cn, N, P))
    rc = _lib.load().xo_transit_lc_host(f(t), N, f(el_c), f(cl_c), Cn, P, f(dtk), f(wk), 1, f(lc), 0)
    assert rc == 0 and np.max(np.abs(lc - O.transit_lc(t, el_o, cl, dtk, wk))) < 1e-14


@pytest.mark.parametrize("P,texp", [(1, None), (3, 0.02)])
def test_parameter_level_model_matches_autograd_prologue_and_oracle(P, texp):
    N, C = 30000, 6
    p = synthetic.transit_params(C, P, seed=21)
    t = synthetic.cadence_times(N)
    y = synthetic.noise(N)
    data = fused.TransitData(t, y, np.full(N, 5e-4), device=DEV)
    rs = np.linspace(0.9, 1.2, C)
    ms = np.linspace(0.8, 1.3, C)
    par = fused.pack_params(dev(p["period"]), dev(p["t0"]), dev(p["ecc"]), dev(p["omega"]), dev(p["b"]), dev(p["ror"]),
                            dev(p["m_planet"]), dev(p["K"]))
    star = fused.pack_star(dev(rs), dev(ms), dev(p["u1"]), dev(p["u2"]))
    mean = dev(p["mean"])
    model = fused.TransitModel(data, texp=texp)
    logp, gpar, gstar, gmean = model.logp_grad(par, star, mean)
    # (a) same thing through the differentiable torch prologue + element-level kernel
    par_t = par.clone().requires_grad_(True)
    star_t = star.clone().requires_grad_(True)
    mean_t = mean.clone().requires_grad_(True)
    el = elements.elements_from_params(par_t[..., 0], par_t[..., 1], par_t[..., 2], par_t[..., 3], par_t[..., 4], par_t[..., 5],
                                       r_star=star_t[:, None, 0], m_star=star_t[:, None, 1], m_planet=par_t[..., 6], texp=texp)
    op = fused.TransitLogp(data, texp=texp)
    lp2 = op(el, fused.limb_dark_coeffs(star_t[:, 2], star_t[:, 3]), mean_t)
    lp2.sum().backward()
    # the two prologues differ by a few ulp in the elements (reciprocal-multiply vs divide, sin E0 from
    # tan(E0/2) vs a second sine); logp is ill-conditioned in them for these random, badly fitting
    # parameters (|d logp / d t0| ~ 1e8 against |logp| ~ 1e5), hence 1e-10 and not 1e-15
    assert torch.allclose(logp, lp2.detach(), rtol=1e-10, atol=0), ((logp - lp2.detach()).abs() / lp2.detach().abs()).max()
    sc = par_t.grad.abs().amax(dim=(0, 1), keepdim=True) + 1e-300
    assert ((gpar - par_t.grad).abs() / sc).max().item() < 1e-9
    assert torch.allclose(gstar, star_t.grad, rtol=1e-8, atol=1e-6 * star_t.grad.abs().max().item())
    assert torch.allclose(gmean, mean_t.grad, rtol=1e-10)
    # (b) oracle logp on the same parameters
    el_o = O.elements_from_params(p["period"], p["t0"], p["ecc"], p["omega"], p["b"], p["ror"], r_star=rs[:, None], m_star=ms[:, None],
                                  m_planet=p["m_planet"], texp=texp)
    lo = O.transit_logp_grad(t, y, np.full(N, 1 / 5e-4 ** 2), el_o, O.get_cl(p["u1"], p["u2"]), p["mean"], *O.stencil(texp),
                             lognorm=data.lognorm)[0]
    assert np.max(np.abs(logp.cpu().numpy() - lo) / np.abs(lo)) < 1e-10
    # elements produced by the prologue kernel
    import ctypes as Ct
    from exoplanet_b200 import _lib
    elem = torch.empty((C, P, 16), dtype=torch.float64, device=DEV)
    clk = torch.empty((C, 3), dtype=torch.float64, device=DEV)
    rc = _lib.load().xo_elements_from_params(Ct.c_void_p(par.data_ptr()), Ct.c_void_p(star.data_ptr()), C, P, 0.0 if texp is None else texp,
                                             1, Ct.c_void_p(elem.data_ptr()), Ct.c_void_p(clk.data_ptr()), None)
    assert rc == 0
    torch.cuda.synchronize()
    assert np.allclose(elem.cpu().numpy()[..., :11], el_o[..., :11], rtol=1e-13, atol=1e-13)
    assert np.allclose(elem.cpu().numpy()[..., 11:13], el_o[..., 11:13], rtol=0, atol=1e-10)


def test_resident_data_handle_matches_device_path():
    import ctypes as C

    from exoplanet_b200 import _lib

    lib = _lib.require_cuda()
    N, Cn, P = 30000, 5, 2
    p = synthetic.transit_params(Cn, P, seed=77)
    t = synthetic.cadence_times(N)
    y = 1.0 + synthetic.noise(N)
    sig = np.full(N, 5e-4)
    data = fused.TransitData(t, y, sig, device=DEV)
    par = fused.pack_params(dev(p["period"]), dev(p["t0"]), dev(p["ecc"]), dev(p["omega"]), dev(p["b"]), dev(p["ror"]), dev(p["m_planet"]),
                            dev(p["K"]))
    star = fused.pack_star(torch.ones(Cn, dtype=torch.float64, device=DEV), torch.ones(Cn, dtype=torch.float64, device=DEV), dev(p["u1"]),
                           dev(p["u2"]))
    mean = dev(1.0 + p["mean"])
    ref = fused.TransitModel(data, texp=0.02).logp_grad(par, star, mean)
    f = lambda a: a.ctypes.data_as(C.c_void_p)
    ivar = np.ascontiguousarray(1 / sig ** 2)
    h = C.c_void_p()
    assert lib.xo_transit_data_create(f(t), f(y), f(ivar), N, 0, C.byref(h)) == 0
    dtk, wk = fused.exposure_stencil(0.02)
    parh, starh, meanh = par.cpu().numpy(), star.cpu().numpy(), mean.cpu().numpy()
    out = [np.empty(Cn), np.empty((Cn, P, 8)), np.empty((Cn, 4)), np.empty(Cn)]
    for use_moments in (1, 0):
        rc = lib.xo_transit_model_logp_grad_resident(h, f(parh), f(starh), f(meanh), Cn, P, f(dtk), f(wk), dtk.size, 0.02, 1, data.lognorm,
                                                     use_moments, *[f(x) for x in out])
        assert rc == 0
        assert np.allclose(out[0], ref[0].cpu().numpy(), rtol=1e-12)
        assert np.allclose(out[1], ref[1].cpu().numpy(), rtol=1e-9, atol=1e-9 * np.abs(out[1]).max())
        assert np.allclose(out[2], ref[2].cpu().numpy(), rtol=1e-9, atol=1e-9 * np.abs(out[2]).max())
    assert lib.xo_transit_data_destroy(h) == 0
    # the fully stateless host call (light curve uploaded inside the call) gives the same numbers
    out2 = [np.empty(Cn), np.empty((Cn, P, 8)), np.empty((Cn, 4)), np.empty(Cn)]
    rc = lib.xo_transit_model_logp_grad_host(f(t), f(y), f(ivar), N, f(parh), f(starh), f(meanh), Cn, P, f(dtk), f(wk), dtk.size, 0.02, 1,
                                             data.lognorm, *[f(x) for x in out2], 0)
    assert rc == 0
    for a_, b_ in zip(out, out2):
        assert np.allclose(a_, b_, rtol=1e-9, atol=1e-9 * np.abs(a_).max())


def test_host_entry_point_from_two_threads():
    """the host-buffer call keeps per-thread streams, events and pinned staging: two sampler threads
    driving it concurrently (ctypes drops the GIL) get exactly what each gets alone"""
    import ctypes as C
    import threading

    from exoplanet_b200 import _lib

    lib = _lib.require_cuda()
    N, Cn, P = 20000, 4, 2
    t = synthetic.cadence_times(N)
    y = synthetic.noise(N)
    ivar = np.full(N, 1 / 5e-4 ** 2)
    dtk, wk = np.zeros(1), np.ones(1)
    f = lambda a: a.ctypes.data_as(C.c_void_p)

    def problem_for(seed):
        p = synthetic.transit_params(Cn, P, seed=seed)
        par = np.ascontiguousarray(np.stack([p["period"], p["t0"], p["ecc"], p["omega"], p["b"], p["ror"], p["m_planet"], p["K"]], axis=-1))
        star = np.ascontiguousarray(np.stack([np.ones(Cn), np.ones(Cn), p["u1"], p["u2"]], axis=-1))
        return par, star, np.ascontiguousarray(p["mean"])

    def call(args):
        par, star, mean = args
        out = [np.empty(Cn), np.empty((Cn, P, 8)), np.empty((Cn, 4)), np.empty(Cn)]
        rc = lib.xo_transit_model_logp_grad_host(f(t), f(y), f(ivar), N, f(par), f(star), f(mean), Cn, P, f(dtk), f(wk), 1, 0.0, 1, 0.0,
                                                 *[f(x) for x in out], 0)
        assert rc == 0
        return out

    probs = [problem_for(101), problem_for(202)]
    alone = [call(pr) for pr in probs]
    results = [None, None]
    errors = []

    def worker(k):
        try:
            for _ in range(25):
                results[k] = call(probs[k])
        except Exception as exc:  # surfaced below: an assert inside a thread would otherwise be lost
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for k in range(2):
        for a, b in zip(alone[k], results[k]):
            assert np.array_equal(a, b)


def test_one_pass_index_equals_the_standalone_index_kernel():
    """the time index built inside the moments pass (xo_transit_data_moments_index) == the one the standalone kernel builds
    by one binary search per bin (xo_transit_data_index), entry for entry: long gaps (one thread owns thousands of edges),
    repeated time stamps, two cadences"""
    import ctypes as C

    from exoplanet_b200 import _lib

    lib = _lib.require_cuda()
    rng = np.random.default_rng(5)
    for N in (2, 3, 4097, 60000):
        t = np.cumsum(rng.uniform(0.01, 0.03, N))
        if N > 1000:
            t[N // 3:] += 400.0   # a long gap
            t[N // 2:] += 0.5
            t[100:110] = t[100]   # repeated stamps
        y = 1.0 + 5e-4 * rng.normal(size=N)
        ivar = np.full(N, 4e6)
        tt, yy, vv = dev(t), dev(y), dev(ivar)
        n = lib.xo_transit_moments_len(N)
        a = torch.zeros(n, dtype=torch.float64, device=DEV)
        b_ = torch.zeros(n, dtype=torch.float64, device=DEV)
        ptr = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
        assert lib.xo_transit_data_moments_index(ptr(tt), ptr(yy), ptr(vv), N, ptr(a), n, _lib.stream_ptr()) == 0
        assert lib.xo_transit_data_moments(ptr(tt), ptr(yy), ptr(vv), N, ptr(b_), _lib.stream_ptr()) == 0
        assert lib.xo_transit_data_index(ptr(tt), N, ptr(b_), n, _lib.stream_ptr()) == 0
        a, b_ = a.cpu().numpy(), b_.cpu().numpy()
        nb = int(a[5])
        assert nb == int(b_[5]) == max(N // 2, 1) and a[6] == b_[6]
        assert np.array_equal(a[:7], b_[:7])
        ta, tb = a[1024:].view(np.uint32)[: nb + 1], b_[1024:].view(np.uint32)[: nb + 1]
        assert np.array_equal(ta, tb), (N, np.flatnonzero(ta != tb)[:10])
        # a block too short for the table is refused (XO_EWORKSPACE = 3), nothing is written
        assert lib.xo_transit_data_moments_index(ptr(tt), ptr(yy), ptr(vv), N, ptr(torch.zeros(1024, dtype=torch.float64, device=DEV)), 1024,
                                                 _lib.stream_ptr()) == 3


@pytest.mark.parametrize("case", ["gaps", "short_period", "flag_window", "wide_window", "three_planets", "tiny", "unsorted", "texp"])
def test_indexed_sweep_edge_cases(case):
    """the epoch-enumerating sweep (time index) against the oracle where its bookkeeping could go wrong: irregular
    sampling with long gaps, periods shorter than the cadence (dense fallback), contact-solver flag windows
    (-inf, +inf), windows clipped at half a period, overlapping planets, slots with a handful of cadences, unsorted
    time stamps (stepping fallback) and exposure-widened windows"""
    rng = np.random.default_rng(abs(hash(case)) % 1000)
    N, C, P, texp = 30000, 6, 1, None
    t = synthetic.cadence_times(N)
    if case == "gaps":
        keep = np.ones(N, bool)
        for a in rng.integers(0, N - 3000, 6):
            keep[a:a + int(rng.integers(200, 3000))] = False
        t = t[keep] + 1e-4 * rng.uniform(size=keep.sum())
        t.sort()
        N = t.size
    if case == "tiny":
        N = 700
        t = t[:N]
    if case == "three_planets":
        P = 3
    if case == "texp":
        texp, P = 0.0204, 2
    p = synthetic.transit_params(C, P, seed=77 + len(case))
    if case == "short_period":
        p["period"][1, 0] = 0.31
        p["b"][:2] = 0.2
    if case == "three_planets":
        p["period"][0, 1] = p["period"][0, 0]
        p["t0"][0, 1] = p["t0"][0, 0] + 0.02
        p["period"][1, 2] = 2.0 * p["period"][1, 0]
        p["t0"][1, 2] = p["t0"][1, 0]
    el = O.elements_from_params(p["period"], p["t0"], p["ecc"], p["omega"], p["b"], p["ror"], texp=texp)
    if case == "short_period":  # a period shorter than the cadence: more epochs than cadences in a slot -> dense fallback
        el[0, 0, 10] = 0.011
        el[0, 0, 2] = 2 * np.pi / 0.011
        el[0, 0, 11], el[0, 0, 12] = -0.003, 0.003
    if case == "flag_window":
        el[0, 0, 11], el[0, 0, 12] = -np.inf, np.inf  # what the contact solver's flag produces
        el[1, 0, 11], el[1, 0, 12] = -np.inf, 0.3
    if case == "wide_window":
        el[0, 0, 11] = -0.7 * el[0, 0, 10]  # beyond half a period on both sides: clipped, every cadence evaluated once
        el[0, 0, 12] = 0.8 * el[0, 0, 10]
        el[1, 0, 11] = -0.5 * el[1, 0, 10]
        el[2, 0, 12] = 0.5 * el[2, 0, 10]
    cl = O.get_cl(p["u1"], p["u2"])
    st = O.stencil(texp)
    y = O.transit_lc(t, el[:1], cl[:1], *st).sum(-1)[0] + synthetic.noise(N)
    if case == "unsorted":
        perm = rng.permutation(N)
        t, y = t[perm], y[perm]
    sigma = 5e-4
    data = fused.TransitData(t, y, np.full(N, sigma), device=DEV)
    m = data.moments().cpu().numpy()
    assert (m[5] > 0) == (case != "unsorted")
    logp, gel, gcl, gmean = fused.transit_logp_grad(data, dev(el), dev(cl), dev(p["mean"]), fused.exposure_stencil(texp))
    lo, geo, gco, gmo = O.transit_logp_grad(t, y, np.full(N, 1 / sigma ** 2), el, cl, p["mean"], *st, lognorm=data.lognorm)
    assert np.max(np.abs(logp.cpu().numpy() - lo) / np.abs(lo)) < 1e-11
    scale = np.abs(geo).max(axis=(0, 1), keepdims=True) + 1e-300
    assert np.max(np.abs(gel.cpu().numpy() - geo) / scale) < 1e-10
    assert np.max(np.abs(gcl.cpu().numpy() - gco) / (np.abs(gco).max() + 1e-300)) < 1e-10
    assert np.max(np.abs(gmean.cpu().numpy() - gmo) / np.abs(gmo).max()) < 1e-10


@pytest.mark.parametrize("C,P,N", [(2, 1, 40000), (300, 1, 20000), (512, 8, 3000), (513, 1, 6000), (40, 3, 100000)])
def test_work_partition_shapes(C, P, N):
    """the weighted partition of the logp launch (CTAs dealt to the chains by expected in-transit work, xo_transit.cu
    weighted_partition) in its different regimes -- two chains sharing the whole device very unevenly, more than one wave,
    four chains per thread of the table with eight planets each, one chain too many for the table (uniform split), many CTAs
    per chain on the full-size series -- against the element-level launch (uniform split) of the same elements, and the
    oracle on a few chains"""
    import ctypes as Ct

    from exoplanet_b200 import _lib

    p = synthetic.transit_params(C, P, seed=300 + C)
    if C == 2:  # one chain in transit almost always, the other almost never
        p["period"][0, 0], p["b"][0, 0], p["ror"][0, 0], p["ecc"][0, 0] = 3.0, 0.0, 0.12, 0.0
        p["period"][1, 0], p["b"][1, 0], p["ror"][1, 0] = 29.0, 0.85, 0.03
    t = synthetic.cadence_times(N)
    y = synthetic.noise(N)
    data = fused.TransitData(t, y, np.full(N, 5e-4), device=DEV)
    par = fused.pack_params(dev(p["period"]), dev(p["t0"]), dev(p["ecc"]), dev(p["omega"]), dev(p["b"]), dev(p["ror"]),
                            dev(p["m_planet"]), dev(p["K"]))
    star = fused.pack_star(torch.ones(C, dtype=torch.float64, device=DEV), torch.ones(C, dtype=torch.float64, device=DEV), dev(p["u1"]),
                           dev(p["u2"]))
    mean = dev(p["mean"])
    logp, gpar, gstar, gmean = fused.TransitModel(data).logp_grad(par, star, mean)
    elem = torch.empty((C, P, 16), dtype=torch.float64, device=DEV)
    clk = torch.empty((C, 3), dtype=torch.float64, device=DEV)
    assert _lib.load().xo_elements_from_params(Ct.c_void_p(par.data_ptr()), Ct.c_void_p(star.data_ptr()), C, P, 0.0, 1,
                                               Ct.c_void_p(elem.data_ptr()), Ct.c_void_p(clk.data_ptr()), None) == 0
    lp2, gel2, gcl2, gm2 = fused.transit_logp_grad(data, elem, clk, mean, fused.exposure_stencil(None))
    assert torch.allclose(logp, lp2, rtol=1e-12, atol=0), ((logp - lp2).abs() / lp2.abs()).max()
    assert torch.allclose(gmean, gm2, rtol=1e-9, atol=1e-9 * gm2.abs().max().item())
    again = fused.TransitModel(data).logp_grad(par, star, mean)
    for a, b in zip((logp, gpar, gstar, gmean), again):
        assert torch.equal(a, b)  # the partition is a deterministic function of the parameters
    sel = np.unique(np.linspace(0, C - 1, 3).astype(int))
    el_o = O.elements_from_params(p["period"][sel], p["t0"][sel], p["ecc"][sel], p["omega"][sel], p["b"][sel], p["ror"][sel],
                                  m_planet=p["m_planet"][sel])
    lo, _, _, gmo = O.transit_logp_grad(t, y, np.full(N, 1 / 5e-4 ** 2), el_o, O.get_cl(p["u1"][sel], p["u2"][sel]), p["mean"][sel],
                                        *O.stencil(None), lognorm=data.lognorm)
    assert np.max(np.abs(logp.cpu().numpy()[sel] - lo) / np.abs(lo)) < 1e-10
    assert np.max(np.abs(gmean.cpu().numpy()[sel] - gmo) / np.abs(gmo).max()) < 1e-9


@pytest.mark.parametrize("use_moments,use_in_transit", [(False, True), (True, False), (False, False)])
def test_parameter_level_model_options_agree(use_moments, use_in_transit):
    """the parameter-level model with the data moments switched off (every cadence read and summed in the sweep: the
    stepping producer under the weighted partition, whose weights then include the sweep itself) and / or with the
    in-transit window switched off (every cadence evaluated, keplerian.py:155-157) == the default evaluation"""
    N, C, P = 20000, 12, 2
    p = synthetic.transit_params(C, P, seed=91)
    t = synthetic.cadence_times(N)
    y = synthetic.noise(N)
    data = fused.TransitData(t, y, np.full(N, 5e-4), device=DEV)
    par = fused.pack_params(dev(p["period"]), dev(p["t0"]), dev(p["ecc"]), dev(p["omega"]), dev(p["b"]), dev(p["ror"]),
                            dev(p["m_planet"]), dev(p["K"]))
    one = torch.ones(C, dtype=torch.float64, device=DEV)
    star = fused.pack_star(one, one, dev(p["u1"]), dev(p["u2"]))
    mean = dev(p["mean"])
    ref = fused.TransitModel(data).logp_grad(par, star, mean)
    got = fused.TransitModel(data, use_moments=use_moments, use_in_transit=use_in_transit).logp_grad(par, star, mean)
    assert torch.allclose(got[0], ref[0], rtol=1e-11, atol=0)
    for a, b in zip(got[1:], ref[1:]):
        assert torch.allclose(a, b, rtol=1e-8, atol=1e-9 * b.abs().max().item())


def test_parameter_level_model_random_shapes():
    """the whole parameter-level evaluation (prologue -> partition -> indexed / stepping sweep -> epilogue) on two dozen random
    problem shapes -- chains, planets, cadences, exposure time, gaps in the sampling, a short-period planet among long ones --
    against the oracle: logp, d/d mean and the element gradients pushed through the oracle's own elements"""
    rng = np.random.default_rng(2024)
    for trial in range(24):
        C = int(rng.integers(1, 48))
        P = int(rng.integers(1, 5))
        N = int(rng.integers(300, 24000))
        texp = None if trial % 3 else float(rng.uniform(0.005, 0.05))
        t = synthetic.cadence_times(N)
        if trial % 4 == 1:  # gaps
            keep = np.ones(N, bool)
            for a in rng.integers(0, max(N - 400, 1), 3):
                keep[a:a + int(rng.integers(20, 400))] = False
            t = t[keep]
            N = t.size
        p = synthetic.transit_params(C, P, seed=1000 + trial)
        if trial % 5 == 2:
            p["period"][0, 0] = 0.7  # many epochs of one planet
        y = synthetic.noise(N, seed=trial)
        data = fused.TransitData(t, y, np.full(N, 5e-4), device=DEV)
        par = fused.pack_params(dev(p["period"]), dev(p["t0"]), dev(p["ecc"]), dev(p["omega"]), dev(p["b"]), dev(p["ror"]),
                                dev(p["m_planet"]), dev(p["K"]))
        one = torch.ones(C, dtype=torch.float64, device=DEV)
        logp, gpar, gstar, gmean = fused.TransitModel(data, texp=texp).logp_grad(par, fused.pack_star(one, one, dev(p["u1"]), dev(p["u2"])),
                                                                               dev(p["mean"]))
        el_o = O.elements_from_params(p["period"], p["t0"], p["ecc"], p["omega"], p["b"], p["ror"], m_planet=p["m_planet"], texp=texp)
        lo, geo, gco, gmo = O.transit_logp_grad(t, y, np.full(N, 1 / 5e-4 ** 2), el_o, O.get_cl(p["u1"], p["u2"]), p["mean"], *O.stencil(texp),
                                                lognorm=data.lognorm)
        tag = (trial, C, P, N, texp)
        assert np.max(np.abs(logp.cpu().numpy() - lo) / np.abs(lo)) < 1e-10, tag
        assert np.max(np.abs(gmean.cpu().numpy() - gmo) / (np.abs(gmo).max() + 1e-300)) < 1e-9, tag
        # d logp / d (r/R*) = the element gradient's ror slot (nothing else depends on it but the window): index 5 of par
        g_ror = geo[..., 9]
        assert np.max(np.abs(gpar.cpu().numpy()[..., 5] - g_ror) / (np.abs(g_ror).max() + 1e-300)) < 1e-7, tag
        assert bool(torch.isfinite(gpar).all()) and bool(torch.isfinite(gstar).all()), tag
