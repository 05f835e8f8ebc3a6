"""Fixtures produced by the REFERENCE'S OWN L3 code (oracle/run_reference_glue.py imports
/root/reference/src unmodified onto a numpy stand-in for PyTensor; only the three external ops
come from the oracle).  CPU: the oracle's restatement of that glue reproduces them.
GPU: this engine's KeplerianOrbit / LimbDarkLightCurve mirror and fused kernels reproduce them."""
import numpy as np
import pytest
import torch

import xo_oracle as O


@pytest.fixture(scope="module")
def glue(golden):
    return golden("reference_glue")


def _elements_two_planets(s, texp=None, use_in_transit=True):
    p = s["params"]
    r_star = p["r_star"]
    ror = np.array(s["r_pl"]) / r_star
    return O.elements_from_params(np.array(p["period"]), np.array(p["t0"]), np.array(p["ecc"]), np.array(p["omega"]), np.array(p["b"]),
                                  ror, r_star=r_star, m_star=p["m_star"], m_planet=np.array(p["m_planet"]), texp=texp,
                                  use_in_transit=use_in_transit)


# ------------------------------------------------------------------ CPU: oracle vs the reference's glue
def test_oracle_elements_match_reference_attributes(glue):
    s = glue["two_planets"]
    el = _elements_two_planets(s)
    r_star = s["params"]["r_star"]
    assert np.allclose(el[:, 8] * r_star, s["a"], rtol=1e-14)
    assert np.allclose(el[:, 2], s["n"], rtol=1e-15)
    assert np.allclose(el[:, 1], s["tref"], rtol=1e-13, atol=1e-14)
    assert np.allclose(el[:, 6], s["cos_incl"], rtol=1e-13) and np.allclose(el[:, 7], s["sin_incl"], rtol=1e-14)
    assert np.allclose(el[:, 4], s["cos_omega"]) and np.allclose(el[:, 5], s["sin_omega"])


def test_oracle_positions_and_rv_match_reference(glue):
    s = glue["two_planets"]
    el = _elements_two_planets(s, use_in_transit=False)
    el[:, 13] = [12.0, 3.5]
    t = np.array(s["t"])
    X, Y, Z, rv = O.orbit_eval(t, el[None])
    r_star = s["params"]["r_star"]
    ref = np.array(s["get_relative_position"])  # Omega rotation is applied by the reference; undo it for X, Y
    cO, sO = np.cos(np.array(s["params"]["Omega"])), np.sin(np.array(s["params"]["Omega"]))
    Xr, Yr = X[0] * r_star, Y[0] * r_star
    assert np.allclose(cO * Xr - sO * Yr, ref[0], rtol=0, atol=1e-11)
    assert np.allclose(sO * Xr + cO * Yr, ref[1], rtol=0, atol=1e-11)
    assert np.allclose(Z[0] * r_star, ref[2], rtol=0, atol=1e-11)
    assert np.allclose(rv[0], s["rv_K"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("key,texp,order,os_", [("lc_notexp", None, 0, 7), ("lc_texp_order0", 0.1, 0, 7), ("lc_texp_order1", 0.1, 1, 7),
                                                ("lc_texp_order2", 0.1, 2, 7), ("lc_texp_os4", 0.05, 0, 4)])
def test_oracle_light_curve_matches_reference(glue, key, texp, order, os_):
    s = glue["two_planets"]
    el = _elements_two_planets(s, texp=texp)
    cl = O.get_cl(np.array(s["u"][0]), np.array(s["u"][1]))
    assert np.allclose(cl, s["cl"], rtol=1e-15)
    lc = O.transit_lc(np.array(s["t_lc"]), el[None], cl[None], *O.stencil(texp, os_, order))[0]
    ref = np.array(s[key])
    assert ref.shape == lc.shape and np.count_nonzero(ref) > 20
    assert np.max(np.abs(lc - ref)) < 1e-14
    assert np.array_equal(ref == 0, lc == 0)  # exact zeros out of transit, like the reference's scatter


def test_oracle_circular_and_contact_bug_match_reference(glue):
    s = glue["circular"]
    p = s["params"]
    el = O.elements_from_params(p["period"], p["t0"], 0.0, 0.0, p["b"], s["r"] / p["r_star"], r_star=p["r_star"], m_star=p["m_star"])
    cl = O.get_cl(np.array(s["u"][0]), np.array(s["u"][1]))
    lc = O.transit_lc(np.array(s["t"]), el[None, None], cl[None], np.zeros(1), np.ones(1))[0, :, 0]
    assert np.max(np.abs(lc - np.array(s["lc"])[:, 0])) < 1e-14
    el = O.elements_from_params(p["period"], p["t0"], 0.0, 0.0, p["b"], s["r"] / p["r_star"], r_star=p["r_star"], m_star=p["m_star"], texp=0.02)
    lc = O.transit_lc(np.array(s["t"]), el[None, None], cl[None], *O.stencil(0.02, 7, 0))[0, :, 0]
    assert np.max(np.abs(lc - np.array(s["lc_texp"])[:, 0])) < 1e-14
    c = glue["contact_bug"]
    el = O.elements_from_params(3.456, 0.0, 0.6, -1.5, 0.0, 0.1, texp=0.02)
    cl = O.get_cl(np.array(0.3), np.array(0.2))
    lc = O.transit_lc(np.array(c["t"]), el[None, None], cl[None], *O.stencil(0.02, 7, 0))[0, :, 0]
    assert np.max(np.abs(lc - np.array(c["lc"])[:, 0])) < 1e-14 and np.max(np.abs(lc - np.array(c["lc_all"])[:, 0])) < 1e-14


# ------------------------------------------------------------------ GPU: the engine vs the reference's glue
gpu = pytest.mark.gpu


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device="cuda")


def _orbit(s, **extra):
    from exoplanet_b200.orbits import KeplerianOrbit

    p = s["params"]
    kw = {k: (dev(v) if isinstance(v, list) else v) for k, v in p.items()}
    kw.update(extra)
    if "device" not in kw and not any(isinstance(v, torch.Tensor) for v in kw.values()):
        kw["device"] = "cuda"
    return KeplerianOrbit(**kw)


@gpu
def test_engine_orbit_getters_match_reference(glue):
    s = glue["two_planets"]
    orbit = _orbit(s)
    for name in ("a", "period", "n", "a_star", "a_planet", "K0", "M0", "cos_incl", "sin_incl", "incl", "b", "t_periastron", "tref",
                 "dcosidb", "rho_star", "m_total"):
        got = torch.broadcast_to(getattr(orbit, name), (2,)).cpu().numpy()
        assert np.allclose(got, s[name], rtol=1e-13, atol=1e-14), name
    t = dev(s["t"])
    for getter in ("get_relative_position", "get_planet_position", "get_star_position", "get_planet_velocity", "get_star_velocity",
                   "get_relative_velocity", "get_planet_acceleration", "get_star_acceleration", "get_relative_acceleration"):
        got = np.array([c.cpu().numpy() for c in getattr(orbit, getter)(t)])
        ref = np.array(s[getter])
        tol = 1e-10 if "acceleration" in getter else 1e-12  # accelerations amplify 1 ulp of (sin f, cos f) at e = 0.8
        assert np.max(np.abs(got - ref) / (1 + np.abs(ref))) < tol, getter
    assert np.allclose(orbit.get_radial_velocity(t, K=dev([12.0, 3.5])).cpu().numpy(), s["rv_K"], rtol=1e-12, atol=1e-12)
    assert np.allclose(orbit.get_radial_velocity(t, output_units=1.0).cpu().numpy(), s["rv_rsun_day"], rtol=1e-12, atol=1e-12)
    rho, theta = orbit.get_relative_angles(t, parallax=0.05)
    assert np.allclose(rho.cpu().numpy(), s["relative_angles"][0], rtol=1e-12) and np.allclose(theta.cpu().numpy(), s["relative_angles"][1], atol=1e-12)
    ld = np.array([c.cpu().numpy() for c in orbit.get_relative_position(t, light_delay=True)])
    assert np.max(np.abs(ld - np.array(s["light_delay_position"]))) < 1e-10
    r_pl = dev(s["r_pl"])
    assert orbit.in_transit(t, r=r_pl).cpu().tolist() == s["in_transit"]
    assert orbit.in_transit(t, r=r_pl, texp=0.1).cpu().tolist() == s["in_transit_texp"]


@gpu
@pytest.mark.parametrize("key,kw", [("lc_notexp", {}), ("lc_notexp_all", dict(use_in_transit=False)), ("lc_texp_order0", dict(texp=0.1, order=0)),
                                    ("lc_texp_order1", dict(texp=0.1, order=1)), ("lc_texp_order2", dict(texp=0.1, order=2)),
                                    ("lc_texp_os4", dict(texp=0.05, oversample=4))])
def test_engine_light_curves_match_reference(glue, key, kw):
    from exoplanet_b200.light_curves import LimbDarkLightCurve

    s = glue["two_planets"]
    orbit = _orbit(s)
    lc = LimbDarkLightCurve(s["u"][0], s["u"][1], device="cuda")
    assert np.allclose(lc.c.cpu().numpy(), s["cl"], rtol=1e-15)
    got = lc.get_light_curve(orbit=orbit, r=dev(s["r_pl"]), t=dev(s["t_lc"]), **kw).cpu().numpy()
    ref = np.array(s[key])
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) < 1e-14
    generic = lc._generic_light_curve(orbit, dev(s["r_pl"]), dev(s["t_lc"]), kw.get("texp"), kw.get("oversample", 7), kw.get("order", 0),
                                      kw.get("use_in_transit", True), False).cpu().numpy()
    assert np.max(np.abs(generic - ref)) < 1e-14


@gpu
def test_engine_circular_secondary_and_simple_match_reference(glue):
    from exoplanet_b200.light_curves import LimbDarkLightCurve, SecondaryEclipseLightCurve
    from exoplanet_b200.orbits import KeplerianOrbit, SimpleTransitOrbit

    s = glue["circular"]
    orbit = _orbit(s)
    assert abs(orbit.a.item() - s["a"]) < 1e-12 and abs(orbit.incl.item() - s["incl"]) < 1e-13
    lc = LimbDarkLightCurve(s["u"][0], s["u"][1], device="cuda")
    t = dev(s["t"])
    assert np.max(np.abs(lc.get_light_curve(orbit=orbit, r=s["r"], t=t).cpu().numpy() - np.array(s["lc"]))) < 1e-14
    assert np.max(np.abs(lc.get_light_curve(orbit=orbit, r=s["r"], t=t, texp=0.02).cpu().numpy() - np.array(s["lc_texp"]))) < 1e-14
    assert orbit.in_transit(t, r=s["r"]).cpu().tolist() == s["in_transit"]
    ror = lc.get_ror_from_approx_transit_depth(dev([0.005, 0.01]), dev([0.1, 0.35])).cpu().numpy()
    assert np.allclose(ror, s["ror_from_depth"], rtol=1e-14)
    c = glue["contact_bug"]
    o3 = KeplerianOrbit(period=3.456, ecc=0.6, omega=-1.5, device="cuda")
    got = LimbDarkLightCurve(0.3, 0.2, device="cuda").get_light_curve(orbit=o3, r=0.1, t=dev(c["t"]), texp=0.02).cpu().numpy()
    assert np.max(np.abs(got - np.array(c["lc"]))) < 1e-14
    # secondary eclipse + _flip
    q = glue["secondary"]
    o4 = KeplerianOrbit(period=1.543, t0=-0.123, ecc=0.2, omega=0.7, m_star=1.0, r_star=1.0, m_planet=0.01, b=0.1, device="cuda")
    se = SecondaryEclipseLightCurve(dev([0.3, 0.2]), dev([0.4, 0.1]), 0.3)
    got = se.get_light_curve(orbit=o4, r=0.08, t=dev(q["t"])).cpu().numpy()
    assert np.max(np.abs(got.reshape(-1) - np.array(q["lc"]).reshape(-1))) < 1e-13
    fl = o4._flip(0.08)
    assert abs(fl.a.item() - q["flip_a"]) < 1e-12 and abs(fl.t_periastron.item() - q["flip_t_periastron"]) < 1e-12
    pos = np.array([x.cpu().numpy() for x in fl.get_planet_position(dev(q["t"][:100]))])
    assert np.max(np.abs(pos - np.array(q["flip_planet_position"]))) < 1e-11
    # simple orbit
    m = glue["simple"]
    so = SimpleTransitOrbit(period=3.456, duration=0.12, t0=1.45, b=0.5, r_star=1.2, ror=0.04, device="cuda")
    t5 = dev(m["t"])
    pos = np.array([x.cpu().numpy() for x in so.get_relative_position(t5)])
    assert np.max(np.abs(pos - np.array(m["relpos"]))) < 1e-13
    assert so.in_transit(t5, r=0.048).cpu().tolist() == m["in_transit"]
    got = LimbDarkLightCurve(0.3, 0.2, device="cuda").get_light_curve(orbit=so, r=0.048, t=t5).cpu().numpy()
    assert np.max(np.abs(got - np.array(m["lc"]))) < 1e-14
