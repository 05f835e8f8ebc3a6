"""GPU parity at BASELINE.json's own configurations and sizes ("what's weak" #1):

  configs[2]  5 planets, joint transit + RV, N = 1e5, every gradient against the oracle
  configs[3]  celerite SHO log-likelihood AND its exact gradient at N = 1e6 against the O(N) C oracle
  configs[4]  the 1024-chain step: sampled chains against the oracle (logp and all parameter gradients)
  f3          TTVOrbit against the fixture made by the reference's own ttv.py (oracle/run_reference_glue.py)
  f4          BatchedNUTS on the device: Gaussian target moments; transit posterior against BatchedHMC
"""
import numpy as np
import pytest
import torch

import xo_oracle as O
from exoplanet_b200 import elements, fused, gp, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"
SIGMA = 5e-4


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=DEV)


def _rel(a, b, axis=None):
    scale = np.abs(b).max(axis=axis, keepdims=axis is not None) + 1e-300
    return float(np.max(np.abs(a - b) / scale))


# ------------------------------------------------------------------------------------------ configs[2]
def test_config2_five_planet_transit_plus_rv_all_gradients():
    """5-planet KeplerianOrbit joint transit + RV model, N = 1e5 cadences, batched VJP through all ops:
    element-level gradients of both likelihoods against the oracle for 8 chains, and the parameter-level
    evaluation (prologue + kernels + reverse prologue) against the torch twin of the prologue chained with the
    oracle's element gradients."""
    N, C, P = 100000, 8, 5
    p = synthetic.transit_params(C, P, seed=synthetic.SEED + 5)
    t = synthetic.cadence_times(N)
    el_o = O.elements_from_params(p["period"], p["t0"], p["ecc"], p["omega"], p["b"], p["ror"], m_planet=p["m_planet"], K=p["K"])
    cl = O.get_cl(p["u1"], p["u2"])
    dtk, wk = O.stencil(None)
    truth = O.transit_lc(t, el_o[:1], cl[:1], dtk, wk).sum(-1)[0]
    y = truth + synthetic.noise(N, SIGMA)
    ivar = np.full(N, 1 / SIGMA ** 2)
    data = fused.TransitData(t, y, np.full(N, SIGMA), device=DEV)
    logp, gel, gcl, gmean = fused.transit_logp_grad(data, dev(el_o), dev(cl), dev(p["mean"]))
    lo, geo, gco, gmo = O.transit_logp_grad(t, y, ivar, el_o, cl, p["mean"], dtk, wk, lognorm=data.lognorm)
    assert _rel(logp.cpu().numpy(), lo) < 1e-12
    assert _rel(gel.cpu().numpy(), geo, axis=(0, 1)) < 1e-10
    assert _rel(gcl.cpu().numpy(), gco) < 1e-10
    assert _rel(gmean.cpu().numpy(), gmo) < 1e-10
    # RV at the same time stamps
    rv_obs = 3.0 * np.random.default_rng(7).normal(size=N)
    rvdata = fused.TransitData(t, rv_obs, np.full(N, 3.0), device=DEV)
    lr, grv, gg = fused.rv_logp_grad(rvdata, dev(el_o), dev(p["gamma"]))
    lro, grvo, ggo = O.rv_logp_grad(t, rv_obs, np.full(N, 1 / 9.0), el_o, p["gamma"], lognorm=rvdata.lognorm)
    assert _rel(lr.cpu().numpy(), lro) < 1e-12
    assert _rel(grv.cpu().numpy(), grvo, axis=(0, 1)) < 1e-10
    assert _rel(gg.cpu().numpy(), ggo) < 1e-10
    # parameter level: xo_transit_model_logp_grad == chain rule of the differentiable torch prologue on the oracle's gel
    par = fused.pack_params(dev(p["period"]), dev(p["t0"]), dev(p["ecc"]), dev(p["omega"]), dev(p["b"]), dev(p["ror"]), dev(p["m_planet"]),
                            dev(p["K"]))
    one = torch.ones(C, dtype=torch.float64, device=DEV)
    star = fused.pack_star(one, one, dev(p["u1"]), dev(p["u2"]))
    lm, gpar, gstar, gm = fused.TransitModel(data).logp_grad(par, star, dev(p["mean"]))
    assert _rel(lm.cpu().numpy(), lo) < 1e-11
    leaves = [dev(p[k]).requires_grad_(True) for k in ("period", "t0", "ecc", "omega", "b", "ror")]
    el_t = elements.elements_from_params(*leaves, m_planet=dev(p["m_planet"]))
    full = torch.zeros_like(el_t)
    full[..., : geo.shape[-1]] = dev(geo)
    # the contact-point window is piecewise constant in the likelihood: no gradient flows through it
    ref = torch.autograd.grad((el_t * full).sum(), leaves, allow_unused=True)
    for j, g in enumerate(ref):
        got = gpar[..., j].cpu().numpy()
        assert _rel(got, g.cpu().numpy()) < 1e-8, (j, got, g)


# ------------------------------------------------------------------------------------------ configs[3]
@pytest.mark.parametrize("Q", [3.0, 0.70710678, 0.3])
def test_config3_celerite_gradient_at_1e6(Q):
    """transit + celerite2 SHOTerm GP, N = 1e6: log-likelihood and the exact gradient w.r.t. (S0, w0, Q), y and
    diag against the sequential adjoint of the C oracle (itself pinned to the 40-digit dense identities)."""
    N = 1000000
    pr = synthetic.gp_problem(N, Q=Q)
    ll, gh, gy, gd = gp.sho_grad(dev(pr["t"]), dev(pr["y"]), dev(pr["diag"]), pr["S0"], pr["w0"], pr["Q"])
    lo, gho, gyo, gdo = O.celerite_sho_grad_c(pr["t"], pr["y"], pr["diag"], pr["S0"], pr["w0"], pr["Q"])
    assert abs(ll.item() - lo) < 1e-10 * abs(lo)
    assert np.all(np.abs(gh.cpu().numpy() - gho) < 1e-8 * (np.abs(gho) + 1e-6 * np.abs(gho).max())), (gh, gho)
    assert np.max(np.abs(gy.cpu().numpy() - gyo)) < 1e-9 * np.abs(gyo).max()
    assert np.max(np.abs(gd.cpu().numpy() - gdo)) < 1e-9 * np.abs(gdo).max()
    # the light-weight entry points agree with it
    ll1 = gp.sho_loglike(dev(pr["t"]), dev(pr["y"]), dev(pr["diag"]), pr["S0"], pr["w0"], pr["Q"]).item()
    assert abs(ll1 - lo) < 1e-10 * abs(lo)
    ll2, alpha = gp.sho_solve(dev(pr["t"]), dev(pr["y"]), dev(pr["diag"]), pr["S0"], pr["w0"], pr["Q"])
    assert abs(ll2.item() - lo) < 1e-10 * abs(lo)
    assert np.max(np.abs(alpha.cpu().numpy() + gyo)) < 1e-9 * np.abs(gyo).max()


def test_celerite_extreme_variances_do_not_overflow_the_log_determinant():
    """sum log D is accumulated as the log of a renormalised running product: variances of 1e+-120 must not
    take the product out of the double range."""
    rng = np.random.default_rng(3)
    N = 4099
    t = np.sort(rng.uniform(0, 400, N))
    for scale in (1e-120, 1e-60, 1e60, 1e120):
        diag = np.full(N, 0.01 * scale)
        y = np.sqrt(scale) * rng.normal(size=N)
        S0 = 1.3 * scale
        ll = gp.sho_loglike(dev(t), dev(y), dev(diag), S0, 2.1, 3.0).item()
        lo = O.celerite_sho_loglike(t, y, diag, S0, 2.1, 3.0)
        assert np.isfinite(ll) and abs(ll - lo) < 1e-9 * abs(lo), (scale, ll, lo)
        llg = gp.sho_grad(dev(t), dev(y), dev(diag), S0, 2.1, 3.0)[0].item()
        assert abs(llg - lo) < 1e-9 * abs(lo), (scale, llg, lo)


# ------------------------------------------------------------------------------------------ configs[4]
def test_config4_1024_chain_step_sampled_chains():
    """the 1024-chain step of configs[4] on one GPU: logp and every parameter gradient of 16 sampled chains
    against the oracle (element gradients chained through the torch twin of the prologue)."""
    N, C = 100000, 1024
    p = synthetic.transit_params(C, 1, seed=synthetic.SEED + 11, ecc=0.5)
    t = synthetic.cadence_times(N)
    truth = synthetic.transit_params(1, 1, seed=synthetic.SEED - 1, ecc=0.5)
    el_t = O.elements_from_params(truth["period"], truth["t0"], truth["ecc"], truth["omega"], truth["b"], truth["ror"])
    dtk, wk = O.stencil(None)
    y = O.transit_lc(t, el_t, O.get_cl(truth["u1"], truth["u2"]), dtk, wk)[0, :, 0] + synthetic.noise(N, SIGMA)
    data = fused.TransitData(t, y, np.full(N, SIGMA), device=DEV)
    par = fused.pack_params(dev(p["period"]), dev(p["t0"]), dev(p["ecc"]), dev(p["omega"]), dev(p["b"]), dev(p["ror"]), dev(p["m_planet"]),
                            dev(p["K"]))
    one = torch.ones(C, dtype=torch.float64, device=DEV)
    star = fused.pack_star(one, one, dev(p["u1"]), dev(p["u2"]))
    logp, gpar, gstar, gmean = fused.TransitModel(data).logp_grad(par, star, dev(p["mean"]))
    sel = np.linspace(0, C - 1, 16).astype(int)
    el_o = O.elements_from_params(p["period"][sel], p["t0"][sel], p["ecc"][sel], p["omega"][sel], p["b"][sel], p["ror"][sel],
                                  m_planet=p["m_planet"][sel])
    cl = O.get_cl(p["u1"][sel], p["u2"][sel])
    lo, geo, gco, gmo = O.transit_logp_grad(t, y, np.full(N, 1 / SIGMA ** 2), el_o, cl, p["mean"][sel], dtk, wk, lognorm=data.lognorm)
    assert _rel(logp.cpu().numpy()[sel], lo) < 1e-11
    assert _rel(gmean.cpu().numpy()[sel], gmo) < 1e-10
    leaves = [dev(p[k][sel]).requires_grad_(True) for k in ("period", "t0", "ecc", "omega", "b", "ror")]
    twin = elements.elements_from_params(*leaves, m_planet=dev(p["m_planet"][sel]))
    full = torch.zeros_like(twin)
    full[..., : geo.shape[-1]] = dev(geo)
    ref = torch.autograd.grad((twin * full).sum(), leaves, allow_unused=True)
    for j, g in enumerate(ref):
        assert _rel(gpar[..., j].cpu().numpy()[sel], g.cpu().numpy()) < 1e-8, j
    u = [dev(p[k][sel]).requires_grad_(True) for k in ("u1", "u2")]
    gu = torch.autograd.grad((fused.limb_dark_coeffs(*u) * dev(gco)).sum(), u)
    assert _rel(gstar[:, 2].cpu().numpy()[sel], gu[0].cpu().numpy()) < 1e-8
    assert _rel(gstar[:, 3].cpu().numpy()[sel], gu[1].cpu().numpy()) < 1e-8


# ------------------------------------------------------------------------------------------ f3: TTV
def _ttv_orbit(s, **kw):
    from exoplanet_b200.orbits import TTVOrbit

    sh = s["shape"]
    return TTVOrbit(period=dev(s["periods"]), t0=dev(s["t0"]), ttvs=[np.array(v) for v in s["ttvs"]],
                    transit_inds=[np.array(i) for i in s["transit_inds"]], ecc=dev(sh["ecc"]), omega=dev(sh["omega"]), b=dev(sh["b"]),
                    r_star=sh["r_star"], m_star=sh["m_star"], device=DEV, **kw)


def test_ttv_orbit_matches_reference_fixture(golden):
    """TTVOrbit against the outputs of /root/reference/src/exoplanet/orbits/ttv.py:35-187 (run unmodified by
    oracle/run_reference_glue.py): labelled and filled-in transit times, bin edges, warped times, positions,
    RV, in_transit and the light curve with and without exposure-time integration, for both constructors."""
    from exoplanet_b200.light_curves import LimbDarkLightCurve
    from exoplanet_b200.orbits import TTVOrbit

    s = golden("reference_glue")["ttv"]
    orbit = _ttv_orbit(s)
    t = dev(s["t"])
    for i in range(2):
        assert np.allclose(orbit.transit_times[i].cpu().numpy(), s["transit_times"][i], rtol=0, atol=1e-12)
        assert np.allclose(orbit.all_transit_times[i].cpu().numpy(), s["all_transit_times"][i], rtol=0, atol=1e-12)
        assert np.allclose(orbit._bin_edges[i].cpu().numpy(), s["bin_edges"][i], rtol=0, atol=1e-12)
    assert np.max(np.abs(orbit._warp_times(t).cpu().numpy() - np.array(s["warp_times"]))) < 1e-12
    pos = np.array([c.cpu().numpy() for c in orbit.get_relative_position(t)])
    assert np.max(np.abs(pos - np.array(s["relpos"]))) < 1e-10
    rv = orbit.get_radial_velocity(t, K=dev([5.0, 2.0])).cpu().numpy()
    assert np.max(np.abs(rv - np.array(s["rv"]))) < 1e-11
    r = dev(s["r"])
    assert orbit.in_transit(t, r=r).cpu().tolist() == s["in_transit"]
    lc = LimbDarkLightCurve(s["u"][0], s["u"][1], device=DEV)
    got = lc.get_light_curve(orbit=orbit, r=r, t=t).cpu().numpy()
    ref = np.array(s["lc"])
    # (for these limb-darkening coefficients the reference's un-occulted value s.c - 1 is -1.1e-16, not 0, wherever the OTHER
    # planet is in transit; the fused kernels return an exact 0 there)
    assert np.max(np.abs(got - ref)) < 1e-13 and np.array_equal(np.abs(got) < 1e-15, np.abs(ref) < 1e-15)
    got = lc.get_light_curve(orbit=orbit, r=r, t=t, texp=0.05, oversample=5).cpu().numpy()
    assert np.max(np.abs(got - np.array(s["lc_texp"]))) < 1e-13
    # the other constructor: least-squares ephemeris of the transit times, shape period offset by delta_log_period
    sh = s["shape"]
    ob = TTVOrbit(transit_times=[np.array(v) for v in s["transit_times"]], transit_inds=[np.array(i) for i in s["transit_inds"]],
                  delta_log_period=dev([1e-4, -2e-4]), ecc=dev(sh["ecc"]), omega=dev(sh["omega"]), b=dev(sh["b"]), r_star=sh["r_star"],
                  m_star=sh["m_star"], device=DEV)
    f = s["fit"]
    assert np.allclose(ob.t0.cpu().numpy(), f["t0"], rtol=0, atol=1e-10) and np.allclose(ob.ttv_period.cpu().numpy(), f["ttv_period"], rtol=1e-13)
    assert np.allclose(ob.period.cpu().numpy(), f["period"], rtol=1e-13)
    for i in range(2):
        assert np.allclose(ob.ttvs[i].cpu().numpy(), f["ttvs"][i], rtol=0, atol=1e-10)
    got = lc.get_light_curve(orbit=ob, r=r, t=t).cpu().numpy()
    assert np.max(np.abs(got - np.array(f["lc"]))) < 1e-12


def test_ttv_fused_kernels_match_generic_path_values_and_gradients(golden):
    """the fused TTV kernels (transit-time table as a kernel argument: light curve forward, its reverse pass with one
    gradient per transit time, and the logp+grad kernel with both sweeps) against the op-by-op torch path, which itself
    reproduces the reference fixture above"""
    import ctypes as C

    from exoplanet_b200 import _lib
    from exoplanet_b200.light_curves import LimbDarkLightCurve
    from exoplanet_b200.orbits import TTVOrbit

    s = golden("reference_glue")["ttv"]
    sh = s["shape"]
    t = dev(s["t"])
    r = dev(s["r"])
    leaves = dict(period=dev(s["periods"]).requires_grad_(True), t0=dev(s["t0"]).requires_grad_(True),
                  ecc=dev(sh["ecc"]).requires_grad_(True), omega=dev(sh["omega"]).requires_grad_(True), b=dev(sh["b"]).requires_grad_(True))
    ttvs = [dev(v).requires_grad_(True) for v in s["ttvs"]]

    def orbit():
        return TTVOrbit(period=leaves["period"], t0=leaves["t0"], ttvs=ttvs, transit_inds=[np.array(i) for i in s["transit_inds"]],
                        ecc=leaves["ecc"], omega=leaves["omega"], b=leaves["b"], r_star=sh["r_star"], m_star=sh["m_star"], device=DEV)

    lc = LimbDarkLightCurve(s["u"][0], s["u"][1], device=DEV)
    w = torch.linspace(0.5, 1.5, t.numel(), dtype=torch.float64, device=DEV)[:, None] * dev([1.0, -0.7])[None, :]
    params = list(leaves.values()) + ttvs
    for kw in ({}, dict(texp=0.05, oversample=5)):
        fused_lc = lc.get_light_curve(orbit=orbit(), r=r, t=t, **kw)
        o = orbit()
        generic = lc._generic_light_curve(o, r, t, kw.get("texp"), kw.get("oversample", 7), 0, True, False)
        assert torch.max(torch.abs(fused_lc - generic)).item() < 1e-13
        assert torch.equal(fused_lc.abs() < 1e-15, generic.abs() < 1e-15)
        ref = np.array(s["lc_texp" if kw else "lc"])
        assert np.max(np.abs(fused_lc.detach().cpu().numpy() - ref)) < 1e-13
        ga = torch.autograd.grad((fused_lc * w).sum(), params, allow_unused=True)
        gb = torch.autograd.grad((generic * w).sum(), params, allow_unused=True)
        for a, b_, name in zip(ga, gb, list(leaves) + ["ttv0", "ttv1"]):
            assert a is not None and b_ is not None, name
            assert torch.max(torch.abs(a - b_)).item() < 1e-8 * (b_.abs().max().item() + 1e-12), (name, a, b_)
    # logp + grad with the table: indexed sweep (moments + time index) == stepping sweep == light-curve composition
    o = orbit()
    elem = o.elements(r=r).detach()[None].contiguous()
    times, count = o.ttv_table()
    times = times.detach()[None].contiguous()
    cl = lc.c.detach()[None].contiguous()
    N = t.numel()
    truth = lc.get_light_curve(orbit=o, r=r, t=t).detach().sum(-1)
    y = truth + dev(synthetic.noise(N, SIGMA))
    data = fused.TransitData(np.array(s["t"]), y.cpu().numpy(), np.full(N, SIGMA), device=DEV)
    lib = _lib.require_cuda()
    ptr = lambda x: C.c_void_p(x.data_ptr()) if x is not None else None  # noqa: E731
    mean = dev([1e-4])
    dtk, wk = fused._stencil_dev(fused.exposure_stencil(None), torch.device(DEV))
    res = {}
    for mode in ("indexed", "stepping"):
        ws = torch.empty(lib.xo_transit_logp_grad_workspace(N, 1, 2) // 8 + 1, dtype=torch.float64, device=DEV)
        out = [torch.empty(1, dtype=torch.float64, device=DEV), torch.empty((1, 2, 10), dtype=torch.float64, device=DEV),
               torch.empty((1, 3), dtype=torch.float64, device=DEV), torch.empty(1, dtype=torch.float64, device=DEV),
               torch.empty_like(times)]
        mom = data.moments() if mode == "indexed" else None
        _lib.check(lib.xo_transit_logp_grad_ttv(ptr(data.t), ptr(data.y), ptr(data.ivar), N, ptr(elem), ptr(cl), ptr(mean), 1, 2, ptr(dtk), ptr(wk),
                                                1, data.lognorm, ptr(mom), ptr(times), ptr(count), times.shape[-1], *[ptr(x) for x in out],
                                                ptr(ws), ws.numel() * 8, _lib.stream_ptr()), "logp_grad_ttv")
        res[mode] = [x.clone() for x in out]
    for a, b_ in zip(res["indexed"], res["stepping"]):
        assert torch.allclose(a, b_, rtol=1e-10, atol=1e-10 * b_.abs().max().item())
    # against autograd through the fused light curve: logp = lognorm - 1/2 sum ivar (y - mean - sum_p lc)^2
    lcv = lc.get_light_curve(orbit=orbit(), r=r, t=t)
    logp = data.lognorm - 0.5 * (data.ivar * (data.y - mean - lcv.sum(-1)) ** 2).sum()
    assert abs(res["indexed"][0].item() - logp.item()) < 1e-11 * abs(logp.item())
    gt = torch.autograd.grad(logp, ttvs)
    for p_, (g_, inds) in enumerate(zip(gt, s["transit_inds"])):
        got = res["indexed"][4][0, p_, torch.as_tensor(inds, device=DEV)]
        assert torch.max(torch.abs(got - g_)).item() < 1e-8 * g_.abs().max().item(), (p_, got, g_)


def test_light_delay_fused_kernels_match_reference_fixture_and_generic_gradients(golden):
    """light-travel delay (keplerian.py:411-470) inside the fused kernels: the light curve against the reference's own
    get_light_curve(light_delay=True) (fixture from oracle/run_reference_glue.py), and the reverse pass -- second Kepler
    solve, d delay / d elements with sin i -- against autograd through the op-by-op path"""
    from exoplanet_b200.light_curves import LimbDarkLightCurve
    from exoplanet_b200.light_curves.limb_dark import _FusedLightCurveDelay
    from exoplanet_b200.orbits import KeplerianOrbit

    s = golden("reference_glue")["two_planets"]
    p = s["params"]
    keys = ("period", "t0", "ecc", "omega", "b", "m_planet")
    leaves = {k: dev(p[k]).requires_grad_(True) for k in keys}

    def orbit():
        return KeplerianOrbit(m_star=p["m_star"], r_star=p["r_star"], Omega=dev(p["Omega"]), device=DEV, **leaves)

    lc = LimbDarkLightCurve(s["u"][0], s["u"][1], device=DEV)
    t, r = dev(s["t_lc"]), dev(s["r_pl"])
    w = torch.linspace(0.5, 1.5, t.numel(), dtype=torch.float64, device=DEV)[:, None] * dev([1.0, -0.7])[None, :]
    for key, kw in (("lc_light_delay", {}), ("lc_light_delay_texp", dict(texp=0.1, oversample=7))):
        got = lc.get_light_curve(orbit=orbit(), r=r, t=t, light_delay=True, **kw)
        assert isinstance(got.grad_fn, _FusedLightCurveDelay._backward_cls)
        ref = np.array(s[key])
        assert np.max(np.abs(got.detach().cpu().numpy() - ref)) < 2e-13
        assert np.max(np.abs(ref - np.array(s["lc_notexp_all"]))) > 1e-5 or kw  # the delay is visible in this fixture
        o = orbit()
        generic = lc._generic_light_curve(o, r, t, kw.get("texp"), kw.get("oversample", 7), 0, False, True)
        assert torch.max(torch.abs(got - generic)).item() < 2e-13
        ga = torch.autograd.grad((got * w).sum(), list(leaves.values()), allow_unused=True)
        gb = torch.autograd.grad((generic * w).sum(), list(leaves.values()), allow_unused=True)
        for name, a, b_ in zip(keys, ga, gb):
            assert torch.max(torch.abs(a - b_)).item() < 1e-7 * (b_.abs().max().item() + 1e-12), (name, a, b_)


# ------------------------------------------------------------------------------------------ f4: NUTS on the device
def test_batched_nuts_on_device_gaussian_target():
    """lock-step NUTS on cuda: correlated-scale Gaussian, 256 chains; means and variances recovered"""
    from exoplanet_b200.sampling import BatchedNUTS

    D, C = 6, 256
    sd = torch.tensor([0.1, 0.5, 1.0, 2.0, 5.0, 0.02], dtype=torch.float64, device=DEV)
    mu = torch.tensor([1.0, -2.0, 0.0, 3.0, 10.0, 0.5], dtype=torch.float64, device=DEV)

    def target(q):
        z = (q - mu) / sd
        return -0.5 * (z * z).sum(-1), -z / sd

    g = torch.Generator(device=DEV).manual_seed(3)
    q0 = mu + sd * torch.randn((C, D), dtype=torch.float64, device=DEV, generator=g)
    nuts = BatchedNUTS(target, q0, step_size=0.05, max_depth=7, seed=5)
    nuts.warmup(150)
    out = nuts.sample(150)
    torch.cuda.synchronize()
    flat = out["q"].reshape(-1, D)
    n_eff = flat.shape[0] / 4.0  # conservative
    assert torch.all((flat.mean(0) - mu).abs() < 5.0 * sd / np.sqrt(n_eff) + 1e-12), (flat.mean(0), mu)
    assert torch.all((flat.var(0) / sd ** 2 - 1.0).abs() < 0.1), flat.var(0) / sd ** 2
    assert 0.6 < out["accept_prob"].mean().item() <= 1.0
    assert not bool(nuts.diverging.any())
    assert int(nuts.depth.max()) >= 1 and nuts.n_leapfrog > 0


@pytest.mark.parametrize("use_graph", [False, True])
def test_nuts_tree_kernels_equal_the_tensor_implementation(use_graph):
    """the library's NUTS kernels (xo_nuts_start / leaf_pre / leaf_post / merge) are deterministic functions of the same
    inputs as the tensor code of BatchedNUTS._doubling, uniforms included: from the same seed the two samplers visit the
    same states -- positions, log-densities, tree depths, divergence flags, acceptance statistics -- transition after
    transition, eagerly and as CUDA graphs (a banana-shaped target so that U-turns happen at every depth)"""
    from exoplanet_b200.sampling import BatchedNUTS

    D, C = 5, 96
    sd = torch.tensor([0.3, 1.0, 2.0, 0.05, 4.0], dtype=torch.float64, device=DEV)

    def target(q):
        x = q.clone()
        x[:, 1] = q[:, 1] - 0.3 * q[:, 0] ** 2  # curved: the step size that suits one region turns early in another
        z = x / sd
        g = -z / sd
        g0 = g[:, 0] + g[:, 1] * (-0.6 * q[:, 0])
        gq = g.clone()
        gq[:, 0] = g0
        lp = -0.5 * (z * z).sum(-1)
        # a wall far out: divergences (energy error > 1000) must be handled identically too
        far = q[:, 2].abs() > 9.0
        return torch.where(far, torch.full_like(lp, -1e6), lp), gq

    gen = torch.Generator(device=DEV).manual_seed(11)
    q0 = sd * torch.randn((C, D), dtype=torch.float64, device=DEV, generator=gen)
    depths, diverged = set(), 0
    for step, n_it in ((0.07, 10), (0.12, 6)):  # well-tuned: depths 1..6; too large: divergences on the narrow coordinate
        a = BatchedNUTS(target, q0, step_size=step, max_depth=6, seed=21, use_graph=use_graph, fused_tree=True)
        b_ = BatchedNUTS(target, q0, step_size=step, max_depth=6, seed=21, use_graph=use_graph, fused_tree=False)
        for it in range(n_it):
            a.transition()
            b_.transition()
            if it == 4:  # the adaptation path too
                a._adapt_step()
                b_._adapt_step()
            assert torch.equal(a.depth, b_.depth), (step, it, a.depth, b_.depth)
            assert torch.equal(a.diverging, b_.diverging)
            assert torch.allclose(a.q, b_.q, rtol=1e-10, atol=1e-12), (step, it, (a.q - b_.q).abs().max())
            assert torch.allclose(a.logp, b_.logp, rtol=1e-10, atol=1e-10)
            assert torch.allclose(a.accept_prob, b_.accept_prob, rtol=1e-9, atol=1e-12)
            assert torch.allclose(a.energy_error, b_.energy_error, rtol=1e-8, atol=1e-10)
            assert a.n_leapfrog == b_.n_leapfrog
            depths |= set(a.depth.tolist())
            diverged += int(a.diverging.sum())
    assert len(depths) >= 5 and diverged > 0, (depths, diverged)  # trees of every depth were built, and some diverged


def test_batched_nuts_on_transit_posterior_agrees_with_hmc():
    """NUTS and fixed-length HMC target the same fused transit posterior: their pooled posterior means agree
    within the Monte-Carlo error and their widths within 35 %"""
    from test_gpu_sampling import _problem

    from exoplanet_b200.sampling import BatchedHMC, BatchedNUTS

    post, q0, q_true, _ = _problem(C=64, N=20000)
    hmc = BatchedHMC(post, q0, step_size=2e-3, n_leapfrog=8, seed=11, use_graph=True)
    hmc.warmup(160)
    a = hmc.sample(150)["q"].reshape(-1, q0.shape[1])
    nuts = BatchedNUTS(post, q0, step_size=2e-3, max_depth=6, seed=12)
    nuts.warmup(120)
    o = nuts.sample(100)
    torch.cuda.synchronize()
    b = o["q"].reshape(-1, q0.shape[1])
    assert torch.isfinite(o["logp"]).all()
    assert 0.5 < o["accept_prob"].mean().item() <= 1.0
    sa, sb = a.std(0), b.std(0)
    for j in (0, 1, 5, 8):  # log period, t0, log r/R*, mean flux: the well-constrained coordinates
        assert abs(a[:, j].mean() - b[:, j].mean()).item() < 0.5 * max(sa[j].item(), sb[j].item()), (j, a[:, j].mean(), b[:, j].mean(), sa[j], sb[j])
        assert 0.65 < (sb[j] / sa[j]).item() < 1.55, (j, sa[j], sb[j])


def test_per_cadence_exposure_times_fused_kernels(golden):
    """one exposure time per cadence (limb_dark.py:199-205) inside the fused kernels: against the reference's own
    get_light_curve(texp=vector, use_in_transit=False) (fixture from oracle/run_reference_glue.py), the reference's
    test_variable_texp property (constant vector == scalar, tests/light_curves_test.py:105-146), and the reverse pass
    against autograd through the op-by-op path"""
    from exoplanet_b200.light_curves import LimbDarkLightCurve
    from exoplanet_b200.light_curves.limb_dark import _FusedLightCurveTexp
    from exoplanet_b200.orbits import KeplerianOrbit

    s = golden("reference_glue")["two_planets"]
    p = s["params"]
    keys = ("period", "t0", "ecc", "omega", "b", "m_planet")
    leaves = {k: dev(p[k]).requires_grad_(True) for k in keys}

    def orbit():
        return KeplerianOrbit(m_star=p["m_star"], r_star=p["r_star"], Omega=dev(p["Omega"]), device=DEV, **leaves)

    lc = LimbDarkLightCurve(s["u"][0], s["u"][1], device=DEV)
    t, r = dev(s["t_lc"]), dev(s["r_pl"])
    tv = dev(s["texp_vec"])
    got = lc.get_light_curve(orbit=orbit(), r=r, t=t, texp=tv, oversample=7, order=1, use_in_transit=False)
    assert isinstance(got.grad_fn, _FusedLightCurveTexp._backward_cls)
    ref = np.array(s["lc_texp_vec"])
    assert np.max(np.abs(got.detach().cpu().numpy() - ref)) < 2e-13
    # with the in-transit window (widened by the longest exposure) the same values, exact zeros elsewhere
    win = lc.get_light_curve(orbit=orbit(), r=r, t=t, texp=tv, oversample=7, order=1)
    assert torch.max(torch.abs(win - got)).item() < 1e-15
    # constant vector == scalar, both ways round (the reference's test_variable_texp)
    const = torch.full_like(t, 0.1)
    for uit in (False, None):
        a = lc.get_light_curve(orbit=orbit(), r=r, t=t, texp=0.1, use_in_transit=uit)
        b_ = lc.get_light_curve(orbit=orbit(), r=r, t=t, texp=const, use_in_transit=False)
        assert torch.max(torch.abs(a - b_)).item() < 1e-14
    # reverse pass
    w = torch.linspace(0.5, 1.5, t.numel(), dtype=torch.float64, device=DEV)[:, None] * dev([1.0, -0.7])[None, :]
    generic = lc._generic_light_curve(orbit(), r, t, tv, 7, 1, False, False)
    assert torch.max(torch.abs(generic - got)).item() < 1e-13
    ga = torch.autograd.grad((got * w).sum(), list(leaves.values()), allow_unused=True)
    gb = torch.autograd.grad((generic * w).sum(), list(leaves.values()), allow_unused=True)
    for a, b_, name in zip(ga, gb, keys):
        if b_ is None:
            continue
        assert a is not None, name
        assert torch.max(torch.abs(a - b_)).item() < 1e-8 * (b_.abs().max().item() + 1e-12), (name, a, b_)
