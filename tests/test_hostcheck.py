"""CPU: the kernels' own __host__ __device__ arithmetic (exoplanet_b200/csrc/xo_math.cuh)
compiled for the host by tests/hostcheck and compared with the golden vectors and the
oracle.  This is a pre-GPU logic check of the device code, not a product path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import xo_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "hostcheck", "libhostcheck.so")


@pytest.fixture(scope="module")
def hc():
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    csrc = os.path.join(HERE, "..", "exoplanet_b200", "csrc")
    hdrs = [os.path.join(csrc, h) for h in ("xo_math.cuh", "xo_contact.cuh", "xo_prologue.cuh", "xo_common.cuh", "xo_transit.cuh")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")  # cuda_runtime.h types only, nothing is linked
        subprocess.check_call(["/usr/bin/g++", "-O2", "-march=x86-64-v3", "-ffp-contract=fast", "-shared", "-fPIC", "-I", cuda_inc, "-o", SO,
                               src])
    return C.CDLL(SO)


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_kepler_math(hc, golden):
    R = np.array(golden("kepler")["rows"])
    M, e = np.ascontiguousarray(R[:, 0]), np.ascontiguousarray(R[:, 1])
    s, c = np.empty_like(M), np.empty_like(M)
    hc.hc_kepler(p(M), p(e), p(s), p(c), C.c_int64(M.size))
    assert np.max(np.abs(s - R[:, 2])) < 2e-15 and np.max(np.abs(c - R[:, 3])) < 2e-15
    rng = np.random.default_rng(0)
    n = 200000
    M = rng.uniform(-2000 * np.pi, 2000 * np.pi, n)
    e = rng.uniform(0, 0.99, n)
    M[:2000] = 10.0 ** rng.uniform(-12, 0, 2000)
    e[:2000] = 1 - 10.0 ** rng.uniform(-6, -1, 2000)
    s, c = np.empty(n), np.empty(n)
    hc.hc_kepler(p(M), p(e), p(s), p(c), C.c_int64(n))
    so, co = O.kepler(M, e)
    assert np.max(np.abs(s - so)) < 3e-15 and np.max(np.abs(c - co)) < 3e-15
    assert np.max(np.abs(s * s + c * c - 1)) < 1e-14


def test_qsv_math(hc, golden):
    R = np.array(golden("qsv")["rows"])
    b, r = np.ascontiguousarray(R[:, 0]), np.ascontiguousarray(R[:, 1])
    n = b.size
    S, B, Rr = np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 3))
    hc.hc_qsv(p(b), p(r), p(S), p(B), p(Rr), C.c_int64(n))
    assert np.max(np.abs(S - R[:, 2:5])) < 1e-14
    contact = (np.abs(np.abs(b) + r - 1) < 1e-7) | (np.abs(np.abs(b) - r - 1) < 1e-7)
    assert np.max(np.abs(B - R[:, 5:8])[~contact]) < 1e-13
    assert np.max(np.abs(Rr - R[:, 8:11])[~contact]) < 1e-13
    S2 = np.empty((n, 3))
    hc.hc_qsv_nojac(p(b), p(r), p(S2), C.c_int64(n))
    assert np.array_equal(S, S2)


def test_lc_point_math(hc):
    N = 20000
    t = np.arange(N) * 0.0204 + 0.0137
    el = O.elements_from_params(7.3, 2.1, 0.5, 0.7, 0.35, 0.08)
    cl = O.get_cl(np.array(0.3), np.array(0.2))
    dtk, wk = O.stencil(None)
    lc_o = O.transit_lc(t, el[None, None, :], cl[None, :], dtk, wk)[0, :, 0]
    lc = np.empty(N)
    g = np.empty((N, 10))
    gs = np.empty((N, 3))
    win = np.empty(N, dtype=np.int32)
    hc.hc_lc_point(p(el), p(cl), p(t), p(lc), p(g), p(gs), C.c_int64(N))
    hc.hc_in_window(p(el), p(t), p(win), C.c_int64(N))
    assert np.count_nonzero(lc_o) > 100
    assert np.max(np.abs(lc * win - lc_o)) < 1e-15
    y = np.zeros(N)
    ivar = np.ones(N)
    _, gel, gcl, _ = O.transit_logp_grad(t, y, ivar, el[None, None, :], cl[None, :], np.zeros(1), dtk, wk)
    wres = -(lc * win)
    assert np.allclose((wres[:, None] * g * win[:, None]).sum(0), gel[0, 0], rtol=1e-9, atol=1e-12)
    assert np.allclose((wres[:, None] * gs * win[:, None]).sum(0), gcl[0], rtol=1e-9, atol=1e-12)


def _random_geometry(n, seed):
    rng = np.random.default_rng(seed)
    a = np.exp(rng.uniform(np.log(3.0), np.log(60.0), n))
    e = rng.uniform(0.0, 0.9, n)
    w = rng.uniform(-np.pi, np.pi, n)
    ror = np.exp(rng.uniform(np.log(0.01), np.log(0.2), n))
    b = rng.uniform(0.0, 1.3, n) * (1.0 + ror)  # some geometries graze or miss the star
    cosi = (1.0 + e * np.sin(w)) / ((1.0 - e) * (1.0 + e)) / a * b
    cosi = np.clip(cosi, 0.0, 0.999)
    return a, e, np.cos(w), np.sin(w), cosi, np.sqrt((1.0 - cosi) * (1.0 + cosi)), 1.0 + ror


@pytest.mark.parametrize("fast", [0, 1])
def test_contact_solver_logic(hc, fast):
    """bracketed Newton of csrc/xo_contact.cuh (fast = the fused prologue's variant: series seed, Newton
    stopped after a step below 3e-8, reciprocal-multiply) vs the oracle's bisection on random geometries,
    transiting, grazing and missing"""
    n = 4000
    a, e, cw, sw, ci, si, L = _random_geometry(n, 99)
    Ml, Mr = np.empty(n), np.empty(n)
    flag = np.empty(n, dtype=np.int32)
    hc.hc_contact(p(a), p(e), p(cw), p(sw), p(ci), p(L), C.c_int(fast), p(Ml), p(Mr), p(flag), C.c_int64(n))
    Mlo, Mro, fo = O.contact_points(a, e, cw, sw, ci, si, L)
    # orbits whose periastron dips inside the contact radius itself have no bracketed root in general;
    # either solver may then report "no window" (the caller evaluates every cadence) -- leave them out
    phys = a * (1.0 - e) > 1.5 * L
    assert np.array_equal(flag[phys] != 0, fo[phys] != 0)
    ok = (fo == 0) & (flag == 0)
    assert ok.sum() > 2000 and (~ok).sum() > 200

    def wrap(d):
        return np.abs((d + np.pi) % (2 * np.pi) - np.pi)

    err = np.maximum(wrap(Ml - Mlo), wrap(Mr - Mro))
    b = ci * a * (1.0 - e) * (1.0 + e) / (1.0 + e * sw)
    well = ok & (b < 0.9 * L)  # away from grazing incidence, where the root is well conditioned
    assert well.sum() > 1500 and err[well].max() < 1e-11
    # near grazing |g''/g'| grows; the fast variant stops after a step below 3e-8, i.e. within ~1e-15 |g''/g'|
    assert err[ok].max() < 1e-10


@pytest.mark.parametrize("texp", [0.0, 0.02])
def test_model_prologue_logic(hc, texp):
    """prologue_pair (KeplerianOrbit.__init__ + the in-transit window in its instruction-lean forms) for
    both lanes of a pair vs the numpy restatement of the oracle"""
    n = 3000
    rng = np.random.default_rng(7)
    period = np.exp(rng.uniform(np.log(0.7), np.log(80.0), n))
    t0 = rng.uniform(0, 1, n) * period
    ecc = rng.uniform(0, 0.85, n)
    ecc[:50] = 0.0
    omega = rng.uniform(-np.pi, np.pi, n)
    omega[50:60] = [-np.pi, np.pi, -0.5 * np.pi, 0.5 * np.pi, 0.0, 1e-9, -1e-9, 3.0, -3.0, 1.0]
    b = rng.uniform(0, 1.2, n)
    ror = np.exp(rng.uniform(np.log(0.02), np.log(0.15), n))
    rs, ms = rng.uniform(0.7, 1.5, n), rng.uniform(0.6, 1.4, n)
    mpl = np.exp(rng.uniform(np.log(1e-6), np.log(1e-3), n))
    par = np.ascontiguousarray(np.stack([period, t0, ecc, omega, b, ror, mpl, rng.uniform(1, 30, n)], axis=-1))
    star = np.ascontiguousarray(np.stack([rs, ms, np.full(n, 0.3), np.full(n, 0.2)], axis=-1))
    el = np.empty((n, 16))
    hc.hc_prologue(p(par), p(star), C.c_double(texp), C.c_int(1), p(el), C.c_int64(n))
    el_o = O.elements_from_params(period, t0, ecc, omega, b, ror, r_star=rs, m_star=ms, m_planet=mpl, texp=texp or None, K=par[:, 7])
    phys = np.isfinite(el_o[:, :11]).all(axis=1)  # cos i <= 1 (a random b can ask for more than the orbit allows)
    assert phys.sum() > 2500
    assert np.allclose(el[phys, :11], el_o[phys, :11], rtol=1e-13, atol=1e-13)  # same bar as the GPU test of this kernel
    fin = np.isfinite(el_o[:, 11]) & phys
    grazing = b > 0.9 * (1.0 + ror)
    clear = phys & ~grazing & (el_o[:, 8] * (1.0 - ecc) > 1.5 * (1.0 + ror))  # periastron well outside the contact radius
    assert clear.sum() > 1500 and np.array_equal(np.isfinite(el[clear, 11]), np.isfinite(el_o[clear, 11]))
    assert fin.sum() > 2000 and (phys & ~np.isfinite(el_o[:, 11])).sum() > 50
    both = fin & np.isfinite(el[:, 11])
    assert np.max(np.abs(el[both & ~grazing, 11:13] - el_o[both & ~grazing, 11:13])) < 1e-10
    assert np.max(np.abs(el[both, 11:13] - el_o[both, 11:13])) < 1e-6
    assert np.array_equal(el[:, 13], par[:, 7])
