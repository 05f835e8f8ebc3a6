"""CPU: the oracle (oracle/xo_oracle.c, a restatement of the path) against the committed
golden vectors (tests/golden/*.json, 40-digit mpmath definitions)."""
import numpy as np
import pytest

import xo_oracle as O


def _rows(g):
    return np.array(g["rows"], dtype=np.float64)


def test_kepler_golden(golden):
    g = golden("kepler")
    R = _rows(g)
    sf, cf = O.kepler(R[:, 0], R[:, 1])
    assert np.max(np.abs(sf - R[:, 2])) < 2e-15
    assert np.max(np.abs(cf - R[:, 3])) < 2e-15
    for M, e, E, s, c in g["known_answers"]:
        sf, cf = O.kepler(np.array([M]), np.array([e]))
        assert abs(sf[0] - s) < 1e-15 and abs(cf[0] - c) < 1e-15


def test_kepler_vjp_golden(golden):
    R = _rows(golden("kepler"))
    bs = np.cos(R[:, 0]) + 0.3
    bc = np.sin(R[:, 0]) - 0.2
    bM, be = O.kepler_vjp(R[:, 2], R[:, 3], R[:, 1], bs, bc)
    bf = bs * R[:, 3] - bc * R[:, 2]
    assert np.allclose(bM, bf * R[:, 4], rtol=1e-12, atol=1e-12)
    assert np.allclose(be, bf * R[:, 5], rtol=1e-12, atol=1e-12)


def test_qsv_golden(golden):
    g = golden("qsv")
    R = _rows(g)
    s, dsdb, dsdr = O.quad_solution_vector(R[:, 0], R[:, 1])
    assert np.max(np.abs(s - R[:, 2:5])) < 1e-14
    b, r = np.abs(R[:, 0]), R[:, 1]
    contact = (np.abs(b + r - 1) < 1e-7) | (np.abs(b - r - 1) < 1e-7)  # sqrt-singular derivative
    assert np.max(np.abs(dsdb - R[:, 5:8])[~contact]) < 1e-13
    assert np.max(np.abs(dsdr - R[:, 8:11])[~contact]) < 1e-13
    assert np.max(np.abs(dsdb - R[:, 5:8])[contact]) < 1e-6
    Q = np.array(g["quadrature"])  # definition by quadrature, SURVEY B.2
    s = O.quad_solution_vector(Q[:, 0], Q[:, 1], jac=False)
    assert np.max(np.abs(s - Q[:, 2:5])) < 1e-14


def test_qsv_jacobian_matches_finite_differences():
    rng = np.random.default_rng(0)
    b = rng.uniform(0.01, 1.2, 200)
    r = rng.uniform(0.02, 0.4, 200)
    ok = (np.abs(b + r - 1) > 1e-3) & (np.abs(b - r - 1) > 1e-3) & (np.abs(b - r) > 1e-3)
    b, r = b[ok], r[ok]
    h = 1e-6
    s, dsdb, dsdr = O.quad_solution_vector(b, r)
    fb = (O.quad_solution_vector(b + h, r, jac=False) - O.quad_solution_vector(b - h, r, jac=False)) / (2 * h)
    fr = (O.quad_solution_vector(b, r + h, jac=False) - O.quad_solution_vector(b, r - h, jac=False)) / (2 * h)
    assert np.max(np.abs(fb - dsdb)) < 2e-7
    assert np.max(np.abs(fr - dsdr)) < 2e-7


def test_contact_golden(golden):
    R = _rows(golden("contact"))
    Ml, Mr, flag = O.contact_points(*[R[:, k] for k in range(7)])
    assert np.all(flag == R[:, 9].astype(int))
    twopi = 2 * np.pi
    dl = np.abs((Ml - R[:, 7] + np.pi) % twopi - np.pi)
    dr = np.abs((Mr - R[:, 8] + np.pi) % twopi - np.pi)
    assert dl.max() < 1e-12 and dr.max() < 1e-12


def test_celerite_golden(golden):
    g = golden("celerite")
    t, y, diag = np.array(g["t"]), np.array(g["y"]), np.array(g["diag"])
    for c in g["cases"]:
        n = c["n"]
        ll = O.celerite_sho_loglike(t[:n], y[:n], diag[:n], c["S0"], c["w0"], c["Q"])
        assert abs(ll - c["loglike"]) < 1e-8 * max(1.0, abs(c["loglike"]))
    assert abs(O.celerite_sho_loglike(t, y, diag, 1.3, 2.1, 3.0) - g["appendix_B9"]) < 1e-9


def test_celerite_gradient_golden(golden):
    """the Kalman-filter adjoint restated in oracle/xo_oracle.py vs the dense identities at 40 digits
    (dL/dtheta = 1/2 a^T dK a - 1/2 tr(K^-1 dK), dL/dy = -a, dL/d diag = 1/2 a^2 - 1/2 diag(K^-1))"""
    g = golden("celerite")
    t, y, diag = np.array(g["t"]), np.array(g["y"]), np.array(g["diag"])
    seen = 0
    for c in g["cases"]:
        if "grad" not in c:
            continue
        seen += 1
        gr = c["grad"]
        n = gr["n"]
        for impl in (O.celerite_sho_grad, O.celerite_sho_grad_c):  # pure-Python and C restatements
            ll, gth, gy, gd = impl(t[:n], y[:n], diag[:n], c["S0"], c["w0"], c["Q"], eps=0.0)
            assert abs(ll - gr["loglike"]) < 1e-10 * max(1.0, abs(gr["loglike"]))
            ref = np.array(gr["hyper"])
            assert np.all(np.abs(gth - ref) < 1e-9 * (1.0 + np.abs(ref))), (c["Q"], gth, ref)
            assert np.max(np.abs(gy - np.array(gr["gy"]))) < 1e-9 * (1.0 + np.max(np.abs(gr["gy"])))
            assert np.max(np.abs(gd - np.array(gr["gdiag"]))) < 1e-9 * (1.0 + np.max(np.abs(gr["gdiag"])))
    assert seen >= 3  # under-damped, critically-ish damped (Q = 1/sqrt 2), over-damped


def test_celerite_gradient_c_matches_python_and_finite_differences():
    rng = np.random.default_rng(21)
    N = 3000
    t = np.sort(rng.uniform(0, 400, N))
    diag = (0.1 + 0.05 * rng.uniform(size=N)) ** 2
    y = np.sin(t) + 0.1 * rng.normal(size=N)
    for Q, eps in ((3.0, 1e-5), (0.3, 1e-5), (0.5, 1e-5)):
        a = O.celerite_sho_grad(t[:500], y[:500], diag[:500], 1.3, 2.1, Q, eps=eps)
        b = O.celerite_sho_grad_c(t[:500], y[:500], diag[:500], 1.3, 2.1, Q, eps=eps)
        assert abs(a[0] - b[0]) < 1e-11 * (1 + abs(a[0]))
        for u, v in zip(a[1:], b[1:]):
            assert np.max(np.abs(u - v)) < 1e-10 * (1 + np.max(np.abs(u)))
        # the likelihood the C gradient belongs to is the celerite recurrence of this same file
        ll, gh, _, _ = O.celerite_sho_grad_c(t, y, diag, 1.3, 2.1, Q, eps=eps)
        assert abs(ll - O.celerite_sho_loglike(t, y, diag, 1.3, 2.1, Q, eps=eps)) < 1e-9 * (1 + abs(ll))
        th = np.array([1.3, 2.1, Q])
        for i in range(3):
            if Q == 0.5:
                continue  # clamped: non-smooth in Q, and the recurrence's two nearly equal exponents make its value too noisy for differences
            hh = 1e-6 * th[i]
            p, m = th.copy(), th.copy()
            p[i] += hh
            m[i] -= hh
            fd = (O.celerite_sho_loglike(t, y, diag, *p, eps=eps) - O.celerite_sho_loglike(t, y, diag, *m, eps=eps)) / (2 * hh)
            assert abs(gh[i] - fd) < 2e-6 * (1 + abs(fd)), (Q, i, gh[i], fd)


def test_end_to_end_known_answers(golden):
    g = golden("end_to_end")
    b5 = g["B5"]
    cl = O.get_cl(np.array(b5["u"][0]), np.array(b5["u"][1]))
    assert np.allclose(cl, b5["cl"], rtol=0, atol=1e-15)
    s = O.quad_solution_vector(np.array([b5["b"]]), np.array([b5["r"]]), jac=False)
    assert abs(s[0] @ cl - 1 - b5["lc"]) < 1e-15
    b6 = g["B6"]
    el = O.elements_from_params(b6["period"], 0.0, b6["ecc"], b6["omega"], 0.0, b6["r"])
    assert abs(el[8] - b6["a"]) < 1e-12 and abs(el[2] - b6["n"]) < 1e-13
    assert abs(el[1] - b6["t_periastron"]) < 1e-13
    assert abs(el[11] - b6["t_start"]) < 1e-11 and abs(el[12] - b6["t_end"]) < 1e-11
    cl = O.get_cl(np.array(b6["u"][0]), np.array(b6["u"][1]))
    lc = O.transit_lc(np.array(b6["t"]), el[None, None, :], cl[None, :], np.zeros(1), np.ones(1))[0, :, 0]
    assert np.max(np.abs(lc - np.array(b6["lc"]))) < 2e-15
    X, Y, Z, _ = O.orbit_eval(np.array(b6["t"]), el[None, None, :])
    bb = np.sqrt(X ** 2 + Y ** 2)[0, :, 0]
    assert np.max(np.abs(bb[:5] - np.array(b6["b"][:5]))) < 1e-12
    b8 = g["B8"]
    el = O.elements_from_params(b8["period"], 0.0, b8["ecc"], b8["omega"], 0.0, 0.03 / 1.5, r_star=1.5, m_star=b8["m_star"],
                                m_planet=b8["m_planet"])
    assert abs(el[8] * 1.5 - b8["a"]) < 1e-12
    assert abs(el[11] - b8["t_start"]) < 1e-12 and abs(el[12] - b8["t_end"]) < 1e-12


def test_logp_grad_matches_finite_differences():
    rng = np.random.default_rng(1)
    N = 6000
    t = np.arange(N) * 0.0204 + 0.0137
    el = O.elements_from_params(np.array([7.3, 11.9]), np.array([2.1, 4.4]), np.array([0.5, 0.2]), np.array([0.7, -2.0]),
                                np.array([0.35, 0.6]), np.array([0.08, 0.05]), texp=0.02)
    cl = O.get_cl(np.array(0.3), np.array(0.2))
    y = 1e-3 * rng.normal(size=N)
    ivar = np.full(N, 1 / (5e-4) ** 2)
    mean = np.array([1e-4])
    dtk, wk = O.stencil(0.02, 7, 0)

    def f(elv, clv=cl, mv=mean):
        return O.transit_logp_grad(t, y, ivar, elv[None], clv[None, :], mv, dtk, wk)[0][0]

    logp, gel, gcl, gm = O.transit_logp_grad(t, y, ivar, el[None], cl[None, :], mean, dtk, wk)
    for p in range(2):
        for j, h in [(1, 1e-8), (2, 1e-10), (3, 1e-7), (4, 1e-7), (5, 1e-7), (6, 1e-7), (8, 1e-6), (9, 1e-8)]:
            ep, em = el.copy(), el.copy()
            ep[p, j] += h
            em[p, j] -= h
            fd = (f(ep) - f(em)) / (2 * h)
            assert abs(fd - gel[0, p, j]) <= 2e-5 * abs(fd) + 1e-3, (p, j, fd, gel[0, p, j])
    h = 1e-7
    fd = (f(el, mv=mean + h) - f(el, mv=mean - h)) / (2 * h)
    assert abs(fd - gm[0]) < 1e-5 * abs(fd)
    for j in range(3):
        cp, cm = cl.copy(), cl.copy()
        cp[j] += 1e-8
        cm[j] -= 1e-8
        fd = (f(el, clv=cp) - f(el, clv=cm)) / 2e-8
        assert abs(fd - gcl[0, j]) <= 1e-5 * abs(fd) + 1e-2


def test_rv_grad_matches_finite_differences():
    rng = np.random.default_rng(2)
    N = 500
    t = np.sort(rng.uniform(0, 200, N))
    el = O.elements_from_params(np.array([7.3, 31.9]), np.array([2.1, 4.4]), np.array([0.5, 0.2]), np.array([0.7, -2.0]),
                                np.array([0.35, 0.6]), np.array([0.08, 0.05]), K=np.array([12.0, 3.0]), use_in_transit=False)
    y = rng.normal(size=N) * 3
    ivar = np.full(N, 0.25)
    gamma = np.array([0.3])
    logp, grv, gg = O.rv_logp_grad(t, y, ivar, el[None], gamma)
    idx = [0, 1, 2, 3, 4, 5, 13]
    for p in range(2):
        for k, j in enumerate(idx):
            h = 1e-7 * max(1.0, abs(el[p, j])) if j != 2 else 1e-9
            ep, em = el.copy(), el.copy()
            ep[p, j] += h
            em[p, j] -= h
            fd = (O.rv_logp_grad(t, y, ivar, ep[None], gamma)[0][0] - O.rv_logp_grad(t, y, ivar, em[None], gamma)[0][0]) / (2 * h)
            assert abs(fd - grv[0, p, k]) <= 1e-5 * abs(fd) + 1e-4, (p, j, fd, grv[0, p, k])


def _series(N, C=3, P=2, seed=5):
    from exoplanet_b200 import synthetic

    p = synthetic.transit_params(C, P, seed=seed)
    t = synthetic.cadence_times(N)
    el = O.elements_from_params(p["period"], p["t0"], p["ecc"], p["omega"], p["b"], p["ror"])
    cl = O.get_cl(p["u1"], p["u2"])
    y = synthetic.noise(N)
    return t, y, np.full(N, 4.0e6), el, cl, p["mean"]


def test_logp_grad_thread_count_and_block_additivity():
    t, y, iv, el, cl, mean = _series(20000)
    dtk, wk = O.stencil(None)
    threads = O.max_threads()
    O.set_threads(1)
    a = O.transit_logp_grad(t, y, iv, el, cl, mean, dtk, wk)
    O.set_threads(threads)
    b = O.transit_logp_grad(t, y, iv, el, cl, mean, dtk, wk)
    for x, z in zip(a, b):
        assert np.array_equal(x, z)
    # additive over cadences: the two halves (each below the blocking threshold: plain cadence-order sums) add up to the whole
    h = 8000
    lo = O.transit_logp_grad(t[:h], y[:h], iv[:h], el, cl, mean, dtk, wk)
    hi = O.transit_logp_grad(t[h:16000], y[h:16000], iv[h:16000], el, cl, mean, dtk, wk)
    whole = O.transit_logp_grad(t[:16000], y[:16000], iv[:16000], el, cl, mean, dtk, wk)
    for x, u, v in zip(whole, lo, hi):
        assert np.allclose(x, u + v, rtol=1e-12, atol=1e-9 * np.abs(x).max())
