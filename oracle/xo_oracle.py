"""TEST INFRASTRUCTURE ONLY: ctypes front-end of oracle/libxo_oracle.so.

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package exoplanet_b200 never imports this module.
See oracle/xo_oracle.c for provenance and parity status ("parity unpinned" at
the exoplanet-core binary level; pinned by mpmath definitions + goldens).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libxo_oracle.so")
NEL = 16
NGRAD = 10
NGRAD_RV = 7


def build(force=False):
    src = os.path.join(_HERE, "xo_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def set_threads(n):
    lib().xo_oracle_set_threads(C.c_int(int(n)))


def max_threads():
    return int(lib().xo_oracle_max_threads())


def kepler(M, e, long_double=False):
    M = _f64(M)
    e = _f64(np.broadcast_to(e, M.shape))
    sf = np.empty_like(M)
    cf = np.empty_like(M)
    if long_double:
        lib().xo_oracle_kepler_ld(_p(M), _p(e), _p(sf), _p(cf), None, C.c_int64(M.size))
    else:
        lib().xo_oracle_kepler(_p(M), _p(e), _p(sf), _p(cf), C.c_int64(M.size))
    return sf, cf


def kepler_vjp(sinf, cosf, e, bs, bc):
    sinf, cosf, e, bs, bc = [_f64(x) for x in (sinf, cosf, e, bs, bc)]
    bM = np.empty_like(sinf)
    be = np.empty_like(sinf)
    lib().xo_oracle_kepler_vjp(_p(sinf), _p(cosf), _p(e), _p(bs), _p(bc), _p(bM), _p(be), C.c_int64(sinf.size))
    return bM, be


def quad_solution_vector(b, r, jac=True):
    b = _f64(b)
    r = _f64(np.broadcast_to(r, b.shape))
    s = np.empty(b.shape + (3,))
    dsdb = np.empty(b.shape + (3,))
    dsdr = np.empty(b.shape + (3,))
    lib().xo_oracle_qsv(_p(b), _p(r), _p(s), _p(dsdb), _p(dsdr), C.c_int64(b.size))
    return (s, dsdb, dsdr) if jac else s


def contact_points(a, e, cosw, sinw, cosi, sini, L):
    arrs = [_f64(np.atleast_1d(x)) for x in (a, e, cosw, sinw, cosi, sini, L)]
    arrs = [_f64(x) for x in np.broadcast_arrays(*arrs)]
    n = arrs[0].size
    Ml = np.empty(n)
    Mr = np.empty(n)
    flag = np.empty(n, dtype=np.int32)
    lib().xo_oracle_contact_points(*[_p(x) for x in arrs], _p(Ml), _p(Mr), _p(flag), C.c_int64(n))
    return Ml, Mr, flag


def transit_lc(t, elem, cl, dtk, wk):
    t = _f64(t)
    elem = _f64(elem)
    cl = _f64(cl)
    dtk = _f64(dtk)
    wk = _f64(wk)
    Cc, P = elem.shape[0], elem.shape[1]
    lc = np.empty((Cc, t.size, P))
    lib().xo_oracle_transit_lc(_p(t), C.c_int64(t.size), _p(elem), _p(cl), C.c_int64(Cc), C.c_int64(P), _p(dtk), _p(wk),
                               C.c_int(dtk.size), _p(lc))
    return lc


def transit_logp_grad(t, y, ivar, elem, cl, mean, dtk, wk, lognorm=0.0):
    t, y, ivar, elem, cl, mean, dtk, wk = [_f64(x) for x in (t, y, ivar, elem, cl, mean, dtk, wk)]
    Cc, P = elem.shape[0], elem.shape[1]
    logp = np.empty(Cc)
    gel = np.empty((Cc, P, NGRAD))
    gcl = np.empty((Cc, 3))
    gmean = np.empty(Cc)
    lib().xo_oracle_transit_logp_grad(_p(t), _p(y), _p(ivar), C.c_int64(t.size), _p(elem), _p(cl), _p(mean), C.c_int64(Cc),
                                      C.c_int64(P), _p(dtk), _p(wk), C.c_int(dtk.size), C.c_double(lognorm), _p(logp),
                                      _p(gel), _p(gcl), _p(gmean))
    return logp, gel, gcl, gmean


def orbit_eval(t, elem):
    t = _f64(t)
    elem = _f64(elem)
    Cc, P = elem.shape[0], elem.shape[1]
    out = [np.empty((Cc, t.size, P)) for _ in range(4)]
    lib().xo_oracle_orbit_eval(_p(t), C.c_int64(t.size), _p(elem), C.c_int64(Cc), C.c_int64(P), *[_p(o) for o in out])
    return out


def rv_logp_grad(t, y, ivar, elem, gamma, lognorm=0.0):
    t, y, ivar, elem, gamma = [_f64(x) for x in (t, y, ivar, elem, gamma)]
    Cc, P = elem.shape[0], elem.shape[1]
    logp = np.empty(Cc)
    grv = np.empty((Cc, P, NGRAD_RV))
    gg = np.empty(Cc)
    lib().xo_oracle_rv_logp_grad(_p(t), _p(y), _p(ivar), C.c_int64(t.size), _p(elem), _p(gamma), C.c_int64(Cc), C.c_int64(P),
                                 C.c_double(lognorm), _p(logp), _p(grv), _p(gg))
    return logp, grv, gg


def celerite_sho_loglike(t, y, diag, S0, w0, Q, eps=1e-5):
    t, y, diag = [_f64(x) for x in (t, y, diag)]
    out = C.c_double(0.0)
    lib().xo_oracle_celerite_sho_loglike(_p(t), _p(y), _p(diag), C.c_int64(t.size), C.c_double(S0), C.c_double(w0),
                                         C.c_double(Q), C.c_double(eps), C.byref(out))
    return out.value


def celerite_sho_grad_c(t, y, diag, S0, w0, Q, eps=1e-5):
    """the same quantities as celerite_sho_grad from the C restatement (xo_oracle.c): O(N), fast enough for
    N = 1e6; checked against the pure-Python version and the 40-digit goldens in tests/test_oracle.py"""
    t, y, diag = [_f64(x) for x in (t, y, diag)]
    n = t.size
    ll = C.c_double(0.0)
    gh, gy, gd = np.empty(3), np.empty(n), np.empty(n)
    rc = lib().xo_oracle_celerite_sho_grad(_p(t), _p(y), _p(diag), C.c_int64(n), C.c_double(S0), C.c_double(w0), C.c_double(Q),
                                           C.c_double(eps), C.byref(ll), _p(gh), _p(gy), _p(gd))
    if rc != 0:
        raise RuntimeError("xo_oracle_celerite_sho_grad failed: %d" % rc)
    return ll.value, gh, gy, gd


def celerite_sho_grad(t, y, diag, S0, w0, Q, eps=1e-5):
    """(loglike, [dS0, dw0, dQ], gy[N], gdiag[N]): the SHO GP log-likelihood and its gradient by a
    sequential 2-state Kalman filter followed by its hand-written adjoint (pure-Python loops: small N
    only).  Same contract as celerite2's reverse pass (GaussianProcess.log_likelihood under autodiff);
    pinned against the 40-digit dense identities of mp_truth.sho_loglike_dense_grad
    (tests/golden/celerite.json, tests/test_oracle.py)."""
    import math
    t, y, diag = [_f64(x) for x in (t, y, diag)]
    N = t.size
    a, c = S0 * w0 * Q, 0.5 * w0 / Q
    under = Q >= 0.5
    sgn = 1.0 if under else -1.0
    arg = sgn * (4.0 * Q * Q - 1.0)
    f = math.sqrt(max(arg, eps))
    h = c * f
    w2 = c * c + sgn * h * h

    def phi(dt):
        if under:
            ex = math.exp(-c * dt)
            Cf, Sf = ex * math.cos(h * dt), ex * math.sin(h * dt) / h
        else:
            ep, en = math.exp(-(c - h) * dt), math.exp(-(c + h) * dt)
            Cf, Sf = 0.5 * (ep + en), 0.5 * (ep - en) / h
        return Cf, Sf, np.array([[Cf + c * Sf, Sf], [-w2 * Sf, Cf - c * Sf]])

    # forward: state (m, Z) with Z = Sigma_inf - P, Sigma_inf = diag(a, a w2); record what the adjoint needs
    m, Z = np.zeros(2), np.zeros((2, 2))
    rec, acc = [], 0.0
    for i in range(N):
        dt = t[i] - t[i - 1] if i > 0 else 0.0
        Cf, Sf, F = phi(dt)
        cv, W = F @ m, F @ Z
        Pz = W @ F.T
        k = np.array([a - Pz[0, 0], -Pz[0, 1]])
        D = k[0] + diag[i]
        v = y[i] - cv[0]
        acc += v * v / D + math.log(D)
        rec.append((dt, Cf, Sf, F, m.copy(), W, k, D, v))
        m = cv + k * v / D
        Z = Pz + np.outer(k, k) / D
    ll = -0.5 * acc - 0.5 * N * math.log(2.0 * math.pi)
    # adjoint, last cadence first: mb = dL/dm, Zb = dL/dZ (full symmetric matrix)
    mb, Zb = np.zeros(2), np.zeros((2, 2))
    ga = gc = gh = gw2 = 0.0
    gy, gd = np.zeros(N), np.zeros(N)
    for i in range(N - 1, -1, -1):
        dt, Cf, Sf, F, m0, W, k, D, v = rec[i]
        w = v / D
        vbar = (k @ mb) / D - w
        Dbar = -(k @ mb) * w / D - (k @ Zb @ k) / (D * D) + 0.5 * (w * w - 1.0 / D)
        kb = mb * w + 2.0 * (Zb @ k) / D
        cb = mb.copy()
        cb[0] -= vbar
        Pb = Zb.copy()
        Pb[0, 0] -= kb[0] + Dbar
        Pb[0, 1] -= 0.5 * kb[1]
        Pb[1, 0] -= 0.5 * kb[1]
        ga += kb[0] + Dbar
        gy[i], gd[i] = vbar, Dbar
        Fb = np.outer(cb, m0) + 2.0 * Pb @ W
        mb, Zb = F.T @ cb, F.T @ Pb @ F
        Cb = Fb[0, 0] + Fb[1, 1]
        Sb = c * (Fb[0, 0] - Fb[1, 1]) + Fb[0, 1] - w2 * Fb[1, 0]
        gc += Sf * (Fb[0, 0] - Fb[1, 1]) - dt * (Cf * Cb + Sf * Sb)
        gw2 -= Sf * Fb[1, 0]
        gh += -sgn * dt * h * Sf * Cb + (dt * Cf - Sf) / h * Sb
    gc += 2.0 * c * gw2
    gh += sgn * 2.0 * h * gw2
    dfdQ = sgn * 4.0 * Q / f if arg > eps else 0.0
    dcdQ = -c / Q
    g = [ga * w0 * Q, ga * S0 * Q + gc * 0.5 / Q + gh * 0.5 / Q * f, ga * S0 * w0 + gc * dcdQ + gh * (dcdQ * f + c * dfdQ)]
    return ll, np.array(g), gy, gd


# ---------------------------------------------------------------------------
# host-side helpers shared by tests/bench (pure numpy restatement of the O(P)
# prologue, /root/reference/src/exoplanet/orbits/keplerian.py:133-281)
# ---------------------------------------------------------------------------
G_GRAV = 2942.2062175044193  # orbits/constants.py:32


def elements_from_params(period, t0, ecc, omega, b, ror, r_star=1.0, m_star=1.0, m_planet=0.0, texp=None, K=0.0,
                         use_in_transit=True):
    """(…,) arrays -> element block (…, NEL).  Mirrors KeplerianOrbit.__init__ +
    in_transit() window (keplerian.py:729-767)."""
    period, t0, ecc, omega, b, ror, m_planet, K = np.broadcast_arrays(
        *[np.asarray(x, dtype=np.float64) for x in (period, t0, ecc, omega, b, ror, m_planet, K)])
    a = (G_GRAV * (m_star + m_planet) * period ** 2 / (4 * np.pi ** 2)) ** (1.0 / 3)
    n = 2 * np.pi / period
    cw, sw = np.cos(omega), np.sin(omega)
    E0 = 2 * np.arctan2(np.sqrt(1 - ecc) * cw, np.sqrt(1 + ecc) * (1 + sw))
    M0 = E0 - ecc * np.sin(E0)
    incl_factor = (1 + ecc * sw) / (1 - ecc ** 2)
    cosi = incl_factor * r_star / a * b
    sini = np.sin(np.arccos(cosi))
    tperi = t0 - M0 / n
    tref = tperi - t0
    el = np.zeros(period.shape + (NEL,))
    el[..., 0] = t0
    el[..., 1] = tref
    el[..., 2] = n
    el[..., 3] = ecc
    el[..., 4] = cw
    el[..., 5] = sw
    el[..., 6] = cosi
    el[..., 7] = sini
    el[..., 8] = a / r_star
    el[..., 9] = ror
    el[..., 10] = period
    el[..., 13] = K
    if use_in_transit:
        Ml, Mr, flag = contact_points(a.ravel(), ecc.ravel(), cw.ravel(), sw.ravel(), cosi.ravel(), sini.ravel(),
                                      (r_star * (1 + ror)).ravel())
        Ml = Ml.reshape(period.shape)
        Mr = Mr.reshape(period.shape)
        flag = flag.reshape(period.shape)
        hp = 0.5 * period
        ts = np.mod((Ml - M0) / n + hp, period) - hp
        te = np.mod((Mr - M0) / n + hp, period) - hp
        ts = np.where(ts > 0, ts - period, ts)
        te = np.where(te < 0, te + period, te)
        if texp is not None:
            ts = ts - 0.5 * texp
            te = te + 0.5 * texp
        bad = flag != 0
        el[..., 11] = np.where(bad, -np.inf, ts)
        el[..., 12] = np.where(bad, np.inf, te)
    else:
        el[..., 11] = -np.inf
        el[..., 12] = np.inf
    return el


def elements_from_params_c(period, t0, ecc, omega, b, ror, r_star=1.0, m_star=1.0, m_planet=0.0, texp=None, K=0.0,
                           use_in_transit=True):
    """the same element blocks from the C restatement (xo_oracle_elements_from_params): [C, P] inputs, r_star /
    m_star scalars or [C]; used by the CPU-baseline legs so that their per-evaluation prologue is not Python"""
    arrs = np.broadcast_arrays(*[np.asarray(x, dtype=np.float64) for x in (period, t0, ecc, omega, b, ror, m_planet, K)])
    shape = arrs[0].shape
    planets = shape[-1] if len(shape) > 1 else 1
    par = _f64(np.stack(arrs, axis=-1).reshape(-1, 8))
    n = par.shape[0]
    star = _f64(np.stack(np.broadcast_arrays(np.asarray(r_star, dtype=np.float64), np.asarray(m_star, dtype=np.float64),
                                             np.zeros(max(n // planets, 1))), axis=-1)[:, :2])
    el = np.empty((n, NEL))
    lib().xo_oracle_elements_from_params(_p(par), _p(star), C.c_int64(n), C.c_int64(planets), C.c_double(0.0 if texp is None else texp),
                                         C.c_int(int(use_in_transit)), _p(el))
    return el.reshape(shape + (NEL,))


def stencil(texp, oversample=7, order=0):
    """limb_dark.py:173-198 -> (dt_k * texp, w_k)."""
    if texp is None:
        return np.zeros(1), np.ones(1)
    oversample = int(oversample)
    oversample += 1 - oversample % 2
    w = np.ones(oversample)
    if order == 0:
        dt = np.linspace(-0.5, 0.5, 2 * oversample + 1)[1:-1:2]
    elif order == 1:
        dt = np.linspace(-0.5, 0.5, oversample)
        w[1:-1] = 2
    elif order == 2:
        dt = np.linspace(-0.5, 0.5, oversample)
        w[1:-1:2] = 4
        w[2:-1:2] = 2
    else:
        raise ValueError("order must be <= 2")
    w /= np.sum(w)
    return texp * dt, w


def get_cl(u1, u2):
    """limb_dark.py:11-18"""
    c0 = 1 - u1 - 1.5 * u2
    c1 = u1 + 2 * u2
    c2 = -0.25 * u2
    norm = np.pi * (c0 + c1 / 1.5)
    return np.stack([c0, c1, c2], axis=-1) / np.asarray(norm)[..., None]
