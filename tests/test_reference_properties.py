"""The reference's own property tests for the hot path, re-stated against this engine's
mirror of its interface (KeplerianOrbit / LimbDarkLightCurve over torch CUDA tensors).
Each test cites the reference test it follows (/root/reference/tests/...)."""
import numpy as np
import pytest
import torch

import xo_oracle as O
from exoplanet_b200.light_curves import LimbDarkLightCurve, SecondaryEclipseLightCurve
from exoplanet_b200.orbits import KeplerianOrbit, SimpleTransitOrbit

pytestmark = pytest.mark.gpu
DEV = "cuda"


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=DEV)


def npy(xs):
    return np.array([x.detach().cpu().numpy() for x in xs])


def two_planets(**kw):
    return KeplerianOrbit(m_star=1.45, r_star=1.0, t0=dev([0.5, 17.4]), period=dev([100.0, 37.3]), ecc=dev([0.1, 0.8]),
                          omega=dev([0.5, 1.3]), Omega=dev([0.0, 1.0]), incl=dev([0.25 * np.pi, 0.3 * np.pi]),
                          m_planet=dev([0.5, 0.1]), **kw)


def one_planet(**kw):
    args = dict(m_star=1.3, r_star=1.0, t0=0.5, period=100.0, ecc=0.1, omega=0.5, Omega=1.0, incl=0.25 * np.pi, m_planet=0.1,
                device=DEV)
    args.update(kw)
    return KeplerianOrbit(**args)


def test_center_of_mass():  # orbits/keplerian_test.py:61-88
    t = dev(np.linspace(0, 100, 1000))
    orbit = two_planets()
    m_planet, m_star = np.array([0.5, 0.1]), 1.45
    planet, star = npy(orbit.get_planet_position(t)), npy(orbit.get_star_position(t))
    com = np.sum((m_planet[None, :] * planet + m_star * star) / (m_star + m_planet)[None, :], axis=0)
    assert np.allclose(com, 0.0)


@pytest.mark.parametrize("which", ["star", "planet", "relative"])
def test_velocity_is_time_derivative_of_position(which):  # orbits/keplerian_test.py:91-131 (pins kepler.grad wrt M)
    orbit = one_planet()
    tn = np.linspace(0, 100, 1000)
    pos_fn = getattr(orbit, "get_%s_position" % which)
    vel = npy(getattr(orbit, "get_%s_velocity" % which)(dev(tn)))
    for i in range(3):
        t = dev(tn).requires_grad_(True)
        (g,) = torch.autograd.grad(pos_fn(t)[i].sum(), t)
        assert np.allclose(vel[i], g.cpu().numpy())


@pytest.mark.parametrize("which", ["star", "planet", "relative"])
def test_acceleration_is_time_derivative_of_velocity(which):  # orbits/keplerian_test.py:157-196
    orbit = one_planet(Omega=None)
    tn = np.linspace(0, 100, 1000)
    acc = npy(getattr(orbit, "get_%s_acceleration" % which)(dev(tn)))
    for i in range(3):
        t = dev(tn).requires_grad_(True)
        (g,) = torch.autograd.grad(getattr(orbit, "get_%s_velocity" % which)(t)[i].sum(), t)
        assert np.allclose(acc[i], g.cpu().numpy())


def test_radial_velocity_forms_agree(golden):  # orbits/keplerian_test.py:134-154 + SURVEY B.7
    t = dev(np.linspace(0, 100, 1000))
    orbit = one_planet()
    rv1 = orbit.get_radial_velocity(t, output_units=1.0)  # R_sun / day
    rv2 = orbit.get_radial_velocity(t, K=orbit.K0 * orbit.m_planet * orbit.sin_incl)
    assert np.allclose(rv1.cpu().numpy(), rv2.cpu().numpy())
    b7 = golden("end_to_end")["B7"]
    rv = orbit.get_radial_velocity(dev(b7["t"]), output_units=1.0).cpu().numpy()
    assert np.allclose(rv, b7["rv_rsun_day"], rtol=0, atol=1e-12)
    x, y, z = [c.cpu().numpy() for c in orbit.get_relative_position(dev([0.0, 10.0]))]
    assert np.allclose([x[0], y[0], z[0]], b7["relpos_t0"], rtol=0, atol=1e-9)
    assert np.allclose([x[1], y[1], z[1]], b7["relpos_t10"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("ecc", [0.3, None])
def test_flip(ecc):  # orbits/keplerian_test.py:199-254
    t = dev(np.linspace(0, 100, 1000))
    kw = dict(r_star=1.1, ecc=ecc, omega=0.5) if ecc is not None else dict(r_star=1.1, ecc=None, omega=None)
    orbit1 = one_planet(**kw)
    orbit2 = orbit1._flip(0.7)
    assert np.allclose(npy(orbit1.get_star_position(t)), npy(orbit2.get_planet_position(t)), atol=1e-5)
    assert np.allclose(npy(orbit1.get_planet_position(t)), npy(orbit2.get_star_position(t)), atol=1e-5)


def _in_transit_orbit(ecc, omega):
    return KeplerianOrbit(m_star=1.45, r_star=1.5, t0=dev([0.5, 17.4]), period=dev([10.0, 5.3]), ecc=ecc, omega=omega,
                          Omega=dev([0.0, 1.0]), m_planet=dev([0.3, 0.5]))


def test_in_transit_brackets_the_true_set():  # orbits/keplerian_test.py:257-285 (pins contact_points)
    tn = np.linspace(-20, 20, 1000)
    orbit = _in_transit_orbit(dev([0.1, 0.8]), dev([0.5, 1.3]))
    r_pl, r_star = np.array([0.1, 0.03]), 1.5
    coords = npy(orbit.get_relative_position(dev(tn)))
    r2 = coords[0] ** 2 + coords[1] ** 2
    inds = orbit.in_transit(dev(tn), r=dev(r_pl)).cpu().numpy()
    m = np.isin(np.arange(len(tn)), inds)
    inside = (r2[inds] <= ((r_star + r_pl) ** 2)[None, :]) & (coords[2][inds] > 0)
    assert np.all(np.any(inside, axis=1))
    outside = (r2[~m] > ((r_star + r_pl) ** 2)[None, :]) | (coords[2][~m] <= 0)
    assert np.all(outside)


def test_in_transit_circular_branch_matches_eccentric_solver():  # orbits/keplerian_test.py:288-313
    tn = dev(np.linspace(-20, 20, 1000))
    a = _in_transit_orbit(dev([0.0, 0.0]), dev([0.0, 0.0]))
    b = KeplerianOrbit(m_star=1.45, r_star=1.5, t0=dev([0.5, 17.4]), period=dev([10.0, 5.3]), m_planet=dev([0.3, 0.5]))
    r_pl = dev([0.1, 0.03])
    assert torch.equal(a.in_transit(tn, r=r_pl), b.in_transit(tn, r=r_pl))


def test_impact_parameter_is_recovered_at_t0():  # orbits/keplerian_test.py:352-374
    orbit = KeplerianOrbit(r_star=0.189, m_star=0.151, period=0.4626413, t0=0.2, b=0.5, ecc=0.8, omega=0.1, device=DEV)
    x, y, z = orbit.get_relative_position(dev([0.2]))
    assert np.allclose((torch.sqrt(x * x + y * y) / 0.189).item(), 0.5)
    assert z.item() > 0


def test_light_curve_gradient():  # light_curves_test.py:42-53 (pins ds/db, ds/dr)
    def lc(u1, u2, b, r):
        return LimbDarkLightCurve(u1, u2)._compute_light_curve(b, r)

    b = dev(np.linspace(-1.5, 1.5, 20)).requires_grad_(True)
    r = (0.1 + torch.zeros_like(b)).detach().requires_grad_(True)
    u1 = dev(0.2).reshape(()).requires_grad_(True)
    u2 = dev(0.3).reshape(()).requires_grad_(True)
    assert torch.autograd.gradcheck(lc, (u1, u2, b, r), eps=1e-6, atol=1e-6, rtol=1e-5)


def test_vector_params_deprecated_form():  # light_curves_test.py:56-72
    u = np.array([0.3, 0.2])
    b = dev(np.linspace(-1.5, 1.5, 20))
    a = LimbDarkLightCurve(dev(u))._compute_light_curve(b, 0.1 + torch.zeros_like(b))
    c = LimbDarkLightCurve(u[0], u[1], device=DEV)._compute_light_curve(b, 0.1 + torch.zeros_like(b))
    assert torch.equal(a, c)
    with pytest.raises(AssertionError):
        LimbDarkLightCurve(dev([0.3, 0.2, 0.1]))


def _lc_orbit():
    return KeplerianOrbit(m_star=1.45, r_star=1.5, t0=dev([0.5, 17.4]), period=dev([10.0, 5.3]), ecc=dev([0.1, 0.8]),
                          omega=dev([0.5, 1.3]), m_planet=dev([0.3, 0.5]))


@pytest.mark.parametrize("texp", [None, 0.1])
def test_use_in_transit_does_not_change_the_model(texp):  # light_curves_test.py:75-102
    t = dev(np.linspace(-20, 20, 1000))
    orbit, r = _lc_orbit(), dev([0.1, 0.03])
    lc = LimbDarkLightCurve(0.3, 0.2, device=DEV)
    a = lc.get_light_curve(orbit=orbit, r=r, t=t, texp=texp, use_in_transit=False)
    b = lc.get_light_curve(orbit=orbit, r=r, t=t, texp=texp, use_in_transit=True)
    assert a.shape == (1000, 2)
    assert np.allclose(a.cpu().numpy(), b.cpu().numpy())
    # fused path == generic (op-by-op) path
    g = lc._generic_light_curve(orbit, r, t, texp, 7, 0, True, False)
    assert np.allclose(a.cpu().numpy(), g.cpu().numpy(), rtol=0, atol=1e-14)


def test_variable_texp():  # light_curves_test.py:105-145
    t = dev(np.linspace(-20, 20, 1000))
    orbit, r = _lc_orbit(), dev([0.1, 0.03])
    lc = LimbDarkLightCurve(0.3, 0.2, device=DEV)
    texp0 = 0.1
    vec = lc.get_light_curve(orbit=orbit, r=r, t=t, texp=texp0 + torch.zeros_like(t), use_in_transit=False)
    for uit in (False, True):
        scalar = lc.get_light_curve(orbit=orbit, r=r, t=t, texp=texp0, use_in_transit=uit)
        assert np.allclose(scalar.cpu().numpy(), vec.cpu().numpy())
    # extension: per-cadence texp together with the in-transit selection (window widened by max texp)
    vec2 = lc.get_light_curve(orbit=orbit, r=r, t=t, texp=texp0 + torch.zeros_like(t), use_in_transit=True)
    assert np.allclose(vec.cpu().numpy(), vec2.cpu().numpy())


def test_contact_bug_regression():  # light_curves_test.py:148-164
    orbit = KeplerianOrbit(period=3.456, ecc=0.6, omega=-1.5, device=DEV)
    t = dev(np.linspace(-0.1, 0.1, 1000))
    lc = LimbDarkLightCurve(0.3, 0.2, device=DEV)
    a = lc.get_light_curve(orbit=orbit, r=0.1, t=t, texp=0.02)
    b = lc.get_light_curve(orbit=orbit, r=0.1, t=t, texp=0.02, use_in_transit=False)
    assert np.allclose(a.cpu().numpy(), b.cpu().numpy())


def test_light_curve_autograd_through_fused_path():
    """fused forward/VJP kernels vs the generic op-by-op path: same gradients w.r.t. orbit params."""
    t = dev(np.linspace(-20, 20, 4000))
    names = ["period", "t0", "ecc", "omega", "b"]
    base = dict(period=[10.0, 5.3], t0=[0.5, 17.4], ecc=[0.1, 0.6], omega=[0.5, 1.3], b=[0.2, 0.5])

    def run(generic):
        ps = {k: dev(v).requires_grad_(True) for k, v in base.items()}
        r = dev([0.1, 0.05]).requires_grad_(True)
        u1 = dev(0.3).reshape(()).requires_grad_(True)
        orbit = KeplerianOrbit(m_star=1.2, r_star=1.1, **ps)
        lc = LimbDarkLightCurve(u1, dev(0.2).reshape(()))
        out = lc._generic_light_curve(orbit, r, t, 0.05, 7, 0, True, False) if generic else lc.get_light_curve(orbit=orbit, r=r, t=t,
                                                                                                              texp=0.05)
        w = torch.sin(torch.arange(out.numel(), device=DEV, dtype=torch.float64)).reshape(out.shape)
        gs = torch.autograd.grad((out * w).sum(), [ps[k] for k in names] + [r, u1])
        return out.detach(), gs

    o1, g1 = run(False)
    o2, g2 = run(True)
    assert torch.allclose(o1, o2, rtol=0, atol=1e-14)
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, rtol=1e-8, atol=1e-10 * (1 + b.abs().max().item())), (a, b)


def test_approx_transit_depth():  # light_curves_test.py:270-282
    lc = LimbDarkLightCurve(0.3, 0.2, device=DEV)
    for b in [0.0, 0.3]:
        ror = lc.get_ror_from_approx_transit_depth(dev(0.01).reshape(()), dev(b).reshape(()))
        depth = -lc._compute_light_curve(dev([b]), ror.reshape(1)).item()
        assert abs(depth - 0.01) < 2e-4
    ror, jac = lc.get_ror_from_approx_transit_depth(dev([0.01, 0.02]), dev([0.1, 0.2]), jac=True)
    d = dev([0.01, 0.02]).requires_grad_(True)
    (g,) = torch.autograd.grad(lc.get_ror_from_approx_transit_depth(d, dev([0.1, 0.2])).sum(), d)
    assert torch.allclose(jac, g)


def test_secondary_eclipse():  # light_curves_test.py:285-311
    u1, u2, s, ror = np.array([0.3, 0.2]), np.array([0.4, 0.1]), 0.3, 0.08
    f = ror ** 2 * s
    lc = SecondaryEclipseLightCurve(dev(u1), dev(u2), s)
    t = dev(np.linspace(-6.435, 10.4934, 5000))
    orbit1 = KeplerianOrbit(period=1.543, t0=-0.123, device=DEV)
    orbit2 = KeplerianOrbit(period=orbit1.period, t0=orbit1.t0 + 0.5 * orbit1.period, r_star=ror, m_star=1.0)
    y1 = LimbDarkLightCurve(u1[0], u1[1], device=DEV).get_light_curve(orbit=orbit1, r=ror, t=t)
    y2 = LimbDarkLightCurve(u2[0], u2[1], device=DEV).get_light_curve(orbit=orbit2, r=1.0, t=t)
    model = lc.get_light_curve(orbit=orbit1, r=ror, t=t)
    expect = (y1 + f * y2) / (1 + f)
    assert np.allclose(model.cpu().numpy(), expect.cpu().numpy(), atol=5e-6)


def test_simple_orbit_matches_keplerian():  # orbits/simple_test.py:51-81
    t = dev(np.linspace(0, 10, 5000))
    lc = LimbDarkLightCurve(0.3, 0.2, device=DEV)
    kep = KeplerianOrbit(period=3.456, t0=1.45, b=0.5, duration=0.12, device=DEV, ror=0.04)
    simple = SimpleTransitOrbit(period=3.456, t0=1.45, b=0.5, duration=0.12, r_star=kep.r_star, ror=0.04)
    a = lc.get_light_curve(orbit=kep, r=0.04 * kep.r_star, t=t).cpu().numpy()
    b = lc.get_light_curve(orbit=simple, r=0.04 * kep.r_star, t=t).cpu().numpy()
    assert a.min() < -1e-3
    assert np.allclose(a, b, rtol=0, atol=2e-5)


def test_ttv_orbit_without_ttvs_equals_keplerian():  # orbits/ttv_test.py:49-83
    from exoplanet_b200.orbits import TTVOrbit, compute_expected_transit_times

    periods, t0s, lo, hi = [10.5, 56.34], [45.3, 48.1], 456.023, 595.23
    time = dev(np.linspace(lo, hi, 5000))
    expected = compute_expected_transit_times(lo, hi, periods, t0s)
    for p_, t0_, tt in zip(periods, t0s, expected):  # orbits/ttv_test.py:8-20
        assert np.all(tt >= lo) and np.all(tt <= hi) and tt[0] - p_ < lo and tt[-1] + p_ > hi
    orbit0 = KeplerianOrbit(period=dev(periods), t0=dev(t0s))
    orbit1 = TTVOrbit(period=dev(periods), t0=dev([t[0] for t in expected]), ttvs=[np.zeros_like(t) for t in expected])
    orbit2 = TTVOrbit(transit_times=expected, device=DEV)
    for name in ("get_relative_position", "get_star_position", "get_planet_position", "get_relative_velocity", "get_star_velocity",
                 "get_planet_velocity"):
        ref = npy(getattr(orbit0, name)(time))
        assert np.allclose(ref, npy(getattr(orbit1, name)(time))), name
        assert np.allclose(ref, npy(getattr(orbit2, name)(time))), name
    rv0 = orbit0.get_radial_velocity(time).cpu().numpy()
    assert np.allclose(rv0, orbit1.get_radial_velocity(time).cpu().numpy())
    # consistency of the two constructors (orbits/ttv_test.py:23-46)
    rng = np.random.default_rng(6934104)
    ttvs = [0.01 * rng.normal(size=len(t)) for t in expected]
    a = TTVOrbit(period=dev(periods), t0=dev([t[0] for t in expected]), ttvs=ttvs)
    for i in range(2):
        assert np.allclose(a.transit_times[i].cpu().numpy(), expected[i] + ttvs[i])
    b = TTVOrbit(transit_times=[x for x in a.transit_times])
    c = TTVOrbit(period=b.period, t0=b.t0, ttvs=b.ttvs)
    for i in range(2):
        assert np.allclose(b.transit_times[i].cpu().numpy(), c.transit_times[i].cpu().numpy())
    # a light curve through the generic path moves with the TTVs
    lc = LimbDarkLightCurve(0.3, 0.2, device=DEV)
    shifted = lc.get_light_curve(orbit=a, r=dev([0.05, 0.08]), t=time)
    assert shifted.shape == (5000, 2) and shifted.min().item() < -1e-3
