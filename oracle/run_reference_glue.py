"""TEST INFRASTRUCTURE: run the REFERENCE'S OWN L3 Python (KeplerianOrbit, LimbDarkLightCurve,
SecondaryEclipseLightCurve, SimpleTransitOrbit -- imported unmodified from
/root/reference/src) on a small numpy stand-in for PyTensor, with the three external ops
(exoplanet-core, not installable here) supplied by the CPU oracle, and store inputs + outputs
as tests/golden/reference_glue.json.

    python oracle/run_reference_glue.py        # only works where /root/reference exists

What this pins: every formula of the path that IS in /root/reference (mean anomaly, element
algebra, rotations, RV, in_transit post-processing, exposure stencils, gather/scatter,
limb-darkening coefficients, secondary-eclipse composition) -- SURVEY.md 8a rows a1, a2, a4,
a5, a6, a7, a8, a10.  What it cannot pin: the arithmetic inside the three ops themselves
(those come from the oracle here; see oracle/mp_truth.py for their definition-level pins).
The generated fixture travels to the GPU box; this script does not.
"""
import importlib
import json
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
OUT = os.path.join(HERE, "..", "tests", "golden", "reference_glue.json")
sys.path.insert(0, HERE)
import xo_oracle as O  # noqa: E402


# ----------------------------------------------------------------------------- numpy "tensor"
class T(np.ndarray):
    """ndarray that remembers (parent, index) of its last fancy-index so that
    ``set_subtensor(x[inds], v)`` can be emulated."""

    def __new__(cls, a):
        return np.asarray(a).view(cls)

    def __array_finalize__(self, obj):
        self._parent = None
        self._index = None

    def __getitem__(self, idx):
        out = super().__getitem__(idx)
        if isinstance(out, np.ndarray):
            out = out.view(T)
            out._parent, out._index = self, idx
        return out

    def eval(self):
        return np.asarray(self)


def _t(x):
    return x if isinstance(x, T) else T(np.asarray(x))


def _wrap(f):
    def g(*a, **k):
        out = f(*[np.asarray(v) if isinstance(v, T) else v for v in a], **k)
        return T(out) if isinstance(out, np.ndarray) or np.isscalar(out) else out

    return g


def make_tensor_module():
    pt = types.ModuleType("pytensor.tensor")
    pt.as_tensor_variable = lambda x, **kw: _t(np.asarray(x, dtype=np.float64) if not isinstance(x, np.ndarray) or x.dtype != np.int64 else x)
    for name in ["sqrt", "sin", "cos", "arctan2", "arccos", "arcsin", "zeros_like", "ones_like", "square", "squeeze", "concatenate",
                 "dot", "arange", "mod", "reshape", "any", "all", "abs", "sum", "log", "exp"]:
        setattr(pt, name, _wrap(getattr(np, name)))
    pt.cast = lambda x, dtype: T(np.asarray(x).astype(dtype))
    pt.extra_ops = types.SimpleNamespace(searchsorted=_wrap(np.searchsorted))  # orbits/ttv.py:178
    pt.stack = lambda xs, axis=0: T(np.stack([np.asarray(x) for x in xs], axis=axis))
    pt.switch = _wrap(np.where)
    pt.gt, pt.lt, pt.eq, pt.ge, pt.le = _wrap(np.greater), _wrap(np.less), _wrap(np.equal), _wrap(np.greater_equal), _wrap(np.less_equal)
    pt.and_ = _wrap(np.logical_and)
    pt.shape_padright = lambda x, n_ones=1: T(np.asarray(x).reshape(np.shape(x) + (1,) * n_ones))
    pt.shape_padleft = lambda x, n_ones=1: T(np.asarray(x).reshape((1,) * n_ones + np.shape(x)))

    def set_subtensor(sub, val):
        out = np.array(sub._parent, copy=True)
        out[sub._index] = np.asarray(val)
        return T(out)

    pt.set_subtensor = set_subtensor
    return pt


# ----------------------------------------------------------------------------- fake dependencies
class _U:
    """SI-scale-only stand-in for the astropy units the reference touches."""

    def __init__(self, scale):
        self.scale = float(scale)

    def __mul__(self, o):
        return _U(self.scale * (o.scale if isinstance(o, _U) else o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        return _U(self.scale / (o.scale if isinstance(o, _U) else o))

    def __pow__(self, p):
        return _U(self.scale ** p)

    def to(self, other):
        return _U(self.scale / other.scale)

    @property
    def value(self):
        return self.scale


def install_fakes():
    pt = make_tensor_module()
    pytensor = types.ModuleType("pytensor")
    pytensor.tensor = pt
    pytensor.function = lambda ins, outs, **k: (lambda *a: outs)
    pytensor.grad = None
    pytensor.config = types.SimpleNamespace(change_flags=lambda **k: None)
    grad_mod = types.ModuleType("pytensor.gradient")
    grad_mod.verify_grad = None
    ifelse_mod = types.ModuleType("pytensor.ifelse")
    ifelse_mod.ifelse = lambda c, a, b: a if bool(np.asarray(c)) else b
    raise_mod = types.ModuleType("pytensor.raise_op")

    class Assert:
        def __init__(self, msg=""):
            self.msg = msg

        def __call__(self, x, *conds):
            assert all(bool(np.all(np.asarray(c))) for c in conds), self.msg
            return x

    raise_mod.Assert = Assert
    pm = types.ModuleType("pymc")

    def modelcontext(model):
        raise TypeError("no model on the context stack")

    pm.modelcontext = modelcontext
    # the three external ops: supplied by the oracle (exoplanet-core is not installable here)
    core = types.ModuleType("exoplanet_core")
    core_pymc = types.ModuleType("exoplanet_core.pymc")
    ops = types.ModuleType("exoplanet_core.pymc.ops")

    def kepler(M, e):
        sf, cf = O.kepler(np.asarray(M), np.asarray(e))
        return T(sf), T(cf)

    def quad_solution_vector(b, r):
        return T(O.quad_solution_vector(np.asarray(b), np.asarray(r), jac=False))

    def contact_points(*args):
        Ml, Mr, flag = O.contact_points(*[np.asarray(a) for a in args])
        return [T(Ml), T(Mr), T(flag)]

    ops.kepler, ops.quad_solution_vector, ops.contact_points = kepler, quad_solution_vector, contact_points
    core_pymc.ops = ops
    # astropy: make constants.py take its documented numeric fall-backs (constants.py:26-36)
    astropy = types.ModuleType("astropy")
    units = types.ModuleType("astropy.units")
    units.R_sun, units.M_sun, units.day, units.au = _U(6.957e8), _U(1.98840987e30), _U(86400.0), _U(1.495978707e11)
    units.m, units.s, units.g, units.cm = _U(1.0), _U(1.0), _U(1e-3), _U(1e-2)
    consts = types.ModuleType("astropy.constants")

    class _Boom:
        def to(self, *a, **k):
            raise TypeError("astropy is not available in this image")

    consts.G = _Boom()
    consts.c = _Boom()
    astropy.units, astropy.constants = units, consts
    mods = {"pytensor": pytensor, "pytensor.tensor": pt, "pytensor.gradient": grad_mod, "pytensor.ifelse": ifelse_mod,
            "pytensor.raise_op": raise_mod, "pymc": pm, "exoplanet_core": core, "exoplanet_core.pymc": core_pymc,
            "exoplanet_core.pymc.ops": ops, "astropy": astropy, "astropy.units": units, "astropy.constants": consts}
    sys.modules.update(mods)
    # bare package shells so the reference's __init__ files (which import PyMC distributions,
    # estimators, ...) are not executed -- only the hot-path modules are imported, unmodified
    for name, sub in (("exoplanet", ""), ("exoplanet.orbits", "orbits"), ("exoplanet.light_curves", "light_curves")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF_SRC, "exoplanet", sub)]
        sys.modules[name] = m


def arr(x):
    return np.asarray(x, dtype=np.float64).tolist()


def main():
    if not os.path.isdir(REF_SRC):
        raise SystemExit("/root/reference is not present here")
    install_fakes()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kep = importlib.import_module("exoplanet.orbits.keplerian")
        simple = importlib.import_module("exoplanet.orbits.simple")
        ttv = importlib.import_module("exoplanet.orbits.ttv")
        ld = importlib.import_module("exoplanet.light_curves.limb_dark")
        sec = importlib.import_module("exoplanet.light_curves.secondary_eclipse")
    assert kep.G_grav == 2942.2062175044193  # the fall-back constants of constants.py:32
    out = {"_comment": "generated by oracle/run_reference_glue.py from /root/reference @ 6161d2b; ops from the CPU oracle"}
    warnings.simplefilter("ignore")

    # ---- scenario 1: two eccentric planets with Omega, masses (tests/orbits/keplerian_test.py:61-88,257-271)
    p1 = dict(m_star=1.45, r_star=1.5, t0=[0.5, 17.4], period=[10.0, 5.3], ecc=[0.1, 0.8], omega=[0.5, 1.3], Omega=[0.0, 1.0],
              m_planet=[0.3, 0.5], b=[0.2, 0.4])
    orbit = kep.KeplerianOrbit(**{k: np.array(v) if isinstance(v, list) else v for k, v in p1.items()})
    t = np.linspace(-20, 20, 401)
    s1 = {"params": p1, "t": arr(t)}
    for name in ("a", "period", "n", "a_star", "a_planet", "K0", "M0", "cos_omega", "sin_omega", "cos_incl", "sin_incl", "incl", "b",
                 "t0", "t_periastron", "tref", "dcosidb", "rho_star", "m_total"):
        s1[name] = arr(np.broadcast_to(getattr(orbit, name), (2,)))
    for getter in ("get_relative_position", "get_planet_position", "get_star_position", "get_planet_velocity", "get_star_velocity",
                   "get_relative_velocity", "get_planet_acceleration", "get_star_acceleration", "get_relative_acceleration"):
        s1[getter] = [arr(c) for c in getattr(orbit, getter)(t)]
    s1["rv_K"] = arr(orbit.get_radial_velocity(t, K=np.array([12.0, 3.5])))
    s1["rv_rsun_day"] = arr(-orbit.get_star_velocity(t)[2])
    r_pl = np.array([0.1, 0.03])
    s1["r_pl"] = arr(r_pl)
    s1["in_transit"] = np.asarray(orbit.in_transit(t, r=r_pl)).astype(int).tolist()
    s1["in_transit_texp"] = np.asarray(orbit.in_transit(t, r=r_pl, texp=0.1)).astype(int).tolist()
    s1["relative_angles"] = [arr(c) for c in orbit.get_relative_angles(t, parallax=0.05)]
    s1["light_delay_position"] = [arr(c) for c in orbit.get_relative_position(t, light_delay=True)]
    lc = ld.LimbDarkLightCurve(0.3, 0.2)
    s1["u"] = [0.3, 0.2]
    s1["cl"] = arr(lc.c)
    tt = np.linspace(-20, 20, 2001)
    s1["t_lc"] = arr(tt)
    s1["lc_notexp"] = arr(lc.get_light_curve(orbit=orbit, r=r_pl, t=tt))
    s1["lc_notexp_all"] = arr(lc.get_light_curve(orbit=orbit, r=r_pl, t=tt, use_in_transit=False))
    for order in (0, 1, 2):
        s1["lc_texp_order%d" % order] = arr(lc.get_light_curve(orbit=orbit, r=r_pl, t=tt, texp=0.1, oversample=7, order=order))
    s1["lc_texp_os4"] = arr(lc.get_light_curve(orbit=orbit, r=r_pl, t=tt, texp=0.05, oversample=4))
    # light-travel delay (keplerian.py:411-470): every cadence evaluated at its retarded position (limb_dark.py:156,215)
    # one exposure time per cadence (limb_dark.py:199-205; the reference's own test uses use_in_transit=False)
    texp_vec = 0.02 + 0.08 * (0.5 + 0.5 * np.sin(0.37 * np.arange(len(np.asarray(tt)))))
    s1["texp_vec"] = arr(texp_vec)
    s1["lc_texp_vec"] = arr(lc.get_light_curve(orbit=orbit, r=r_pl, t=tt, texp=texp_vec, oversample=7, order=1, use_in_transit=False))
    s1["lc_light_delay"] = arr(lc.get_light_curve(orbit=orbit, r=r_pl, t=tt, light_delay=True))
    s1["lc_light_delay_texp"] = arr(lc.get_light_curve(orbit=orbit, r=r_pl, t=tt, texp=0.1, oversample=7, light_delay=True))
    out["two_planets"] = s1

    # ---- scenario 2: circular, single planet, impact parameter (data-and-models tutorial shape)
    p2 = dict(period=3.456, t0=1.45, b=0.35, r_star=1.1, m_star=0.9)
    orbit2 = kep.KeplerianOrbit(**p2)
    t2 = np.arange(0, 35, 0.02)
    lc2 = ld.LimbDarkLightCurve(0.45, 0.15)
    s2 = {"params": p2, "t": arr(t2), "u": [0.45, 0.15], "r": 0.08}
    s2["lc"] = arr(lc2.get_light_curve(orbit=orbit2, r=0.08, t=t2))
    s2["lc_texp"] = arr(lc2.get_light_curve(orbit=orbit2, r=0.08, t=t2, texp=0.02))
    s2["in_transit"] = np.asarray(orbit2.in_transit(t2, r=0.08)).astype(int).tolist()
    s2["a"], s2["incl"] = float(orbit2.a), float(orbit2.incl)
    s2["relpos"] = [arr(c) for c in orbit2.get_relative_position(t2[:200])]
    s2["ror_from_depth"] = arr(lc2.get_ror_from_approx_transit_depth(np.array([0.005, 0.01]), np.array([0.1, 0.35])))
    out["circular"] = s2

    # ---- scenario 3: the contact-bug orbit (tests/light_curves_test.py:148-164)
    orbit3 = kep.KeplerianOrbit(period=3.456, ecc=0.6, omega=-1.5)
    t3 = np.linspace(-0.2, 0.2, 801)
    lc3 = ld.LimbDarkLightCurve(0.3, 0.2)
    out["contact_bug"] = {"t": arr(t3), "lc": arr(lc3.get_light_curve(orbit=orbit3, r=0.1, t=t3, texp=0.02)),
                          "lc_all": arr(lc3.get_light_curve(orbit=orbit3, r=0.1, t=t3, texp=0.02, use_in_transit=False))}

    # ---- scenario 4: flip + secondary eclipse (tests/light_curves_test.py:285-311)
    orbit4 = kep.KeplerianOrbit(period=1.543, t0=-0.123, ecc=0.2, omega=0.7, m_star=1.0, r_star=1.0, m_planet=0.01, b=0.1)
    t4 = np.linspace(-1.0, 2.5, 1401)
    se = sec.SecondaryEclipseLightCurve(np.array([0.3, 0.2]), np.array([0.4, 0.1]), 0.3)
    flipped = orbit4._flip(0.08)
    out["secondary"] = {"t": arr(t4), "lc": arr(se.get_light_curve(orbit=orbit4, r=0.08, t=t4)),
                        "flip_planet_position": [arr(c) for c in flipped.get_planet_position(t4[:100])],
                        "flip_a": float(flipped.a), "flip_t_periastron": float(flipped.t_periastron)}

    # ---- scenario 5: SimpleTransitOrbit (tests/orbits/simple_test.py)
    so = simple.SimpleTransitOrbit(period=3.456, duration=0.12, t0=1.45, b=0.5, r_star=1.2, ror=0.04)
    t5 = np.linspace(0, 10, 1001)
    lc5 = ld.LimbDarkLightCurve(0.3, 0.2)
    out["simple"] = {"t": arr(t5), "relpos": [arr(c) for c in so.get_relative_position(t5)],
                     "in_transit": np.asarray(so.in_transit(t5, r=0.048)).astype(int).tolist(),
                     "lc": arr(lc5.get_light_curve(orbit=so, r=0.048, t=t5))}

    # ---- scenario 6: TTVOrbit (orbits/ttv.py:35-187; tests/orbits/ttv_test.py:23-83): both constructors, a
    # missing transit (transit_inds), eccentric shape, light curve with and without exposure integration
    rng = np.random.default_rng(6934104)
    periods, t0s = [10.5, 56.34], [45.3, 48.1]
    tmin, tmax = 30.0, 330.0
    expected = ttv.compute_expected_transit_times(tmin, tmax, periods, t0s)
    inds = [np.delete(np.arange(len(expected[0])), 3), np.arange(len(expected[1]))]  # planet 0 misses transit #3
    ttvs = [0.02 * rng.normal(size=len(i)) for i in inds]
    shape = dict(ecc=np.array([0.3, 0.05]), omega=np.array([0.7, -1.1]), b=np.array([0.3, 0.5]), r_star=1.1, m_star=0.95)
    oa = ttv.TTVOrbit(period=np.array(periods), t0=np.array([e[0] for e in expected]), ttvs=ttvs, transit_inds=inds, **shape)
    t6 = np.linspace(tmin, tmax, 6001)
    r6 = np.array([0.07, 0.11])
    lc6 = ld.LimbDarkLightCurve(0.35, 0.2)
    s6 = {"periods": periods, "t0": arr([e[0] for e in expected]), "ttvs": [arr(v) for v in ttvs], "transit_inds": [i.tolist() for i in inds],
          "shape": {k: (arr(v) if isinstance(v, np.ndarray) else v) for k, v in shape.items()}, "t": arr(t6), "r": arr(r6), "u": [0.35, 0.2],
          "transit_times": [arr(v) for v in oa.transit_times], "all_transit_times": [arr(v) for v in oa.all_transit_times],
          "bin_edges": [arr(v) for v in oa._bin_edges], "warp_times": arr(oa._warp_times(_t(t6))),
          "relpos": [arr(c) for c in oa.get_relative_position(t6)], "rv": arr(oa.get_radial_velocity(t6, K=np.array([5.0, 2.0]))),
          "in_transit": np.asarray(oa.in_transit(t6, r=r6)).astype(int).tolist(),
          "lc": arr(lc6.get_light_curve(orbit=oa, r=r6, t=t6)), "lc_texp": arr(lc6.get_light_curve(orbit=oa, r=r6, t=t6, texp=0.05, oversample=5))}
    ob = ttv.TTVOrbit(transit_times=[np.asarray(v) for v in oa.transit_times], transit_inds=inds, delta_log_period=np.array([1e-4, -2e-4]), **shape)
    s6["fit"] = {"t0": arr(ob.t0), "ttv_period": arr(ob.ttv_period), "period": arr(ob.period), "ttvs": [arr(v) for v in ob.ttvs],
                 "lc": arr(lc6.get_light_curve(orbit=ob, r=r6, t=t6))}
    out["ttv"] = s6

    with open(OUT, "w") as fh:
        json.dump(out, fh)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
