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
    hdr = os.path.join(HERE, "..", "exoplanet_b200", "csrc", "xo_math.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-march=x86-64-v3", "-ffp-contract=fast", "-shared", "-fPIC", "-o", SO, src])
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
