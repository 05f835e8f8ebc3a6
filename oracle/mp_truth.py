"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Tier-(i) oracle: 40-digit mpmath *definitions* of every op on the hot path.
Nothing here restates an algorithm: each function is the mathematical
definition the reference's external ops implement (exoplanet-core, pinned
``>=0.3.0`` at /root/reference/setup.py:35, NOT vendored under
/root/reference), anchored on the reference's call sites:

* ``kepler(M, e) -> (sin f, cos f)``      /root/reference/src/exoplanet/orbits/keplerian.py:333
* ``quad_solution_vector(b, r) -> s[3]``  /root/reference/src/exoplanet/light_curves/limb_dark.py:24
* ``contact_points(a,e,cosw,sinw,cosi,sini,L)`` /root/reference/src/exoplanet/orbits/keplerian.py:744-753
* celerite SHO log-likelihood (celerite2>=0.3.1, setup.py:36; no call site)

Parity status: the reference holds NO golden vectors for this path and its
arithmetic cannot be imported here -> these definitions
are pinned by (1) agreement of independent formulations (root find vs series,
1-D quadrature vs closed forms, dense Cholesky vs recurrence) and (2) the
reference's own *property* tests re-run against them (tests/test_reference_properties.py).
"parity unpinned" at the exoplanet-core binary level; see DESIGN.md.
"""
import mpmath as mp

mp.mp.dps = 40


# ----------------------------------------------------------------------------
# A.1 Kepler
# ----------------------------------------------------------------------------
def kepler_E(M, e):
    """Root E of E - e sin E = M (unique, since d/dE = 1 - e cos E > 0)."""
    M = mp.mpf(M)
    e = mp.mpf(e)
    twopi = 2 * mp.pi
    k = mp.floor((M + mp.pi) / twopi)
    Mr = M - k * twopi  # in [-pi, pi)
    if e == 0:
        return Mr + k * twopi
    sgn = 1 if Mr >= 0 else -1
    Ma = abs(Mr)
    # bisection bracket [Ma, Ma + e] then Newton in high precision
    E = Ma + e * mp.sin(Ma) / (1 - mp.sin(Ma + e) + mp.sin(Ma)) if Ma > 0 else mp.mpf(0)
    if Ma == 0:
        return k * twopi
    for _ in range(200):
        f = E - e * mp.sin(E) - Ma
        d = f / (1 - e * mp.cos(E))
        E = E - d
        if abs(d) < mp.mpf(10) ** (-mp.mp.dps + 3):
            break
    return sgn * E + k * twopi


def kepler(M, e):
    """(sin f, cos f) as the reference's op returns them (keplerian.py:333)."""
    e = mp.mpf(e)
    E = kepler_E(M, e)
    den = 1 - e * mp.cos(E)
    sinf = mp.sqrt(1 - e * e) * mp.sin(E) / den
    cosf = (mp.cos(E) - e) / den
    return sinf, cosf


def kepler_jac(M, e):
    """(df/dM, df/de) -- SURVEY Appendix A.1."""
    e = mp.mpf(e)
    sinf, cosf = kepler(M, e)
    ome2 = 1 - e * e
    dfdM = (1 + e * cosf) ** 2 / ome2 ** mp.mpf(1.5)
    dfde = (2 + e * cosf) * sinf / ome2
    return dfdM, dfde


# ----------------------------------------------------------------------------
# A.2 quadratic limb-darkening solution vector: surface integrals over the
# visible part of the unit disk, as a 1-D quadrature in the radial coordinate.
# ----------------------------------------------------------------------------
def _phi(rho, b, r):
    """Half-angle of the circle of radius rho (centred on the star) covered by
    the occultor (disk radius r centred at distance b)."""
    if rho + b <= r:
        return mp.pi
    if rho <= b - r or rho >= b + r or rho <= r - b - 0:  # outside annulus
        if rho <= r - b:
            return mp.pi
        return mp.mpf(0)
    if b == 0:
        return mp.pi if rho < r else mp.mpf(0)
    c = (rho * rho + b * b - r * r) / (2 * b * rho)
    c = max(min(c, mp.mpf(1)), mp.mpf(-1))
    return mp.acos(c)


def quad_solution_vector(b, r):
    """s_n = int_V g_n dA, g = {1, mu, 4 mu^2 - 2}, mu = sqrt(1-rho^2)."""
    b = abs(mp.mpf(b))
    r = mp.mpf(r)
    if r == 0 or b >= 1 + r:
        return [mp.pi, 2 * mp.pi / 3, mp.mpf(0)]
    if b <= r - 1:
        return [mp.mpf(0)] * 3
    pts = sorted(set([mp.mpf(0), min(mp.mpf(1), abs(b - r)), min(mp.mpf(1), b + r), mp.mpf(1)]))
    g = [
        lambda rho: mp.mpf(1),
        lambda rho: mp.sqrt(max(mp.mpf(0), 1 - rho * rho)),
        lambda rho: 4 * (1 - rho * rho) - 2,
    ]
    out = []
    for gn in g:
        tot = mp.mpf(0)
        for lo, hi in zip(pts[:-1], pts[1:]):
            if hi <= lo:
                continue
            tot += mp.quad(lambda rho: gn(rho) * rho * (2 * mp.pi - 2 * _phi(rho, b, r)), [lo, hi])
        out.append(tot)
    return out


def quad_solution_vector_jac(b, r, h=None):
    """Central differences of the quadrature definition (slow; spot checks)."""
    b = mp.mpf(b)
    r = mp.mpf(r)
    h = mp.mpf(10) ** (-12) if h is None else mp.mpf(h)
    sp = quad_solution_vector(b + h, r)
    sm = quad_solution_vector(b - h, r)
    dsdb = [(p - m) / (2 * h) for p, m in zip(sp, sm)]
    sp = quad_solution_vector(b, r + h)
    sm = quad_solution_vector(b, r - h)
    dsdr = [(p - m) / (2 * h) for p, m in zip(sp, sm)]
    return dsdb, dsdr


# closed forms in elliptic integrals (parameter m), used for *dense* truth grids
# once validated against the quadrature above (tests/test_oracle.py does so).
def _Lambda(b, r):
    b = mp.mpf(b)
    r = mp.mpf(r)
    if r == 0 or b >= 1 + r:
        return mp.mpf(0)
    if b == 0:
        return -mp.mpf(2) / 3 * (1 - r * r) ** mp.mpf(1.5)
    a = (b - r) ** 2
    beta = (b + r) ** 2
    q = r * r - b * b
    if b == r:
        if r == mp.mpf(1) / 2:
            return mp.mpf(1) / 3 - 4 / (9 * mp.pi)
        if r < mp.mpf(1) / 2:
            m = 4 * r * r
            return mp.mpf(1) / 3 + 2 / (9 * mp.pi) * (4 * (2 * r * r - 1) * mp.ellipe(m) + (1 - 4 * r * r) * mp.ellipk(m))
        m = 1 / (4 * r * r)
        return (
            mp.mpf(1) / 3
            + 16 * r / (9 * mp.pi) * (2 * r * r - 1) * mp.ellipe(m)
            - (1 - 4 * r * r) * (3 - 8 * r * r) / (9 * mp.pi * r) * mp.ellipk(m)
        )
    if b == 1 - r:
        return (
            2 / (3 * mp.pi) * mp.acos(1 - 2 * r)
            - 4 / (9 * mp.pi) * (3 + 2 * r - 8 * r * r) * mp.sqrt(r * (1 - r))
            - mp.mpf(2) / 3 * (1 if r > mp.mpf(1) / 2 else 0)
        )
    k2 = (1 - a) / (4 * b * r)
    if k2 < 1:
        return (
            ((1 - beta) * (2 * beta + a - 3) - 3 * q * (beta - 2)) * mp.ellipk(k2)
            + 4 * b * r * (b * b + 7 * r * r - 4) * mp.ellipe(k2)
            - 3 * (q / a) * mp.ellippi((a - 1) / a, k2)
        ) / (9 * mp.pi * mp.sqrt(b * r))
    m = 1 / k2
    return (
        2
        / (9 * mp.pi * mp.sqrt(1 - a))
        * (
            (1 - 5 * b * b + r * r + q * q) * mp.ellipk(m)
            + (1 - a) * (b * b + 7 * r * r - 4) * mp.ellipe(m)
            - 3 * (q / a) * mp.ellippi((a - beta) / a, m)
        )
    )


def quad_solution_vector_closed(b, r):
    b = abs(mp.mpf(b))
    r = mp.mpf(r)
    if r == 0 or b >= 1 + r:
        return [mp.pi, 2 * mp.pi / 3, mp.mpf(0)]
    if b <= r - 1:
        return [mp.mpf(0)] * 3
    if b <= 1 - r:
        s0 = mp.pi * (1 - r * r)
        eta = r * r * (r * r + 2 * b * b) / 2
    else:
        D = mp.sqrt((1 + b + r) * (1 + b - r) * (1 - b + r) * (-1 + b + r))
        k0 = mp.atan2(D, (r - 1) * (r + 1) + b * b)
        k1 = mp.atan2(D, (1 - r) * (1 + r) + b * b)
        s0 = mp.pi - (k1 + r * r * k0 - D / 2)
        eta = (k1 + r * r * (r * r + 2 * b * b) * k0 - (1 + 5 * r * r + b * b) * D / 4) / (2 * mp.pi)
    s2 = 2 * s0 + 4 * mp.pi * eta - 2 * mp.pi
    s1 = 2 * mp.pi / 3 * (0 if r > b else 1) - mp.pi * _Lambda(b, r)
    return [s0, s1, s2]


def quad_solution_vector_jac_closed(b, r):
    """Analytic Jacobians, K and E only."""
    b = abs(mp.mpf(b))
    r = mp.mpf(r)
    z = [mp.mpf(0)] * 3
    if r == 0 or b >= 1 + r or b <= r - 1:
        return z, z
    A = 1 - (b - r) ** 2
    B = 4 * b * r
    if b <= 1 - r:
        k0 = mp.pi
        sk0 = mp.mpf(0)
        ck0 = mp.mpf(-1)
    else:
        D = mp.sqrt((1 + b + r) * (1 + b - r) * (1 - b + r) * (-1 + b + r))
        k0 = mp.atan2(D, (r - 1) * (r + 1) + b * b)
        sk0 = D / (2 * b * r)
        ck0 = mp.cos(k0)
    ds0db = 2 * r * sk0
    ds0dr = -2 * r * k0
    g = 2 - 4 * b * b - 4 * r * r
    ds2dr = -r * (2 * g * k0 + 16 * b * r * sk0)
    ds2db = r * (2 * g * sk0 + 8 * b * r * (k0 + sk0 * ck0))
    if b == 0:
        ds1db = mp.mpf(0)
        ds1dr = -2 * mp.pi * r * mp.sqrt(1 - r * r)
    elif A == B:  # b = 1 - r exactly: E(1) = 1, (1 - m) K(m) -> 0
        ds1dr = -4 * r * mp.sqrt(A)
        ds1db = 4 * r * mp.sqrt(A) / 3
    elif A > B:  # interior, m = B/A < 1
        m = B / A
        K = mp.ellipk(m)
        E = mp.ellipe(m)
        ds1dr = -4 * r * mp.sqrt(A) * E
        ds1db = 4 * r * mp.sqrt(A) * ((2 - m) * E - 2 * (1 - m) * K) / (3 * m)
    else:
        k2 = A / B
        K = mp.ellipk(k2)
        E = mp.ellipe(k2)
        ds1dr = -4 * r * mp.sqrt(B) * (E - (1 - k2) * K)
        ds1db = 4 * r * mp.sqrt(B) / 3 * ((2 * k2 - 1) * E + (1 - k2) * K)
    return [ds0db, ds1db, ds2db], [ds0dr, ds1dr, ds2dr]


# ----------------------------------------------------------------------------
# A.3 contact points
# ----------------------------------------------------------------------------
def _f_to_M(f, e):
    E = 2 * mp.atan(mp.sqrt((1 - e) / (1 + e)) * mp.tan(f / 2))
    return E - e * mp.sin(E)


def contact_points(a, e, cosw, sinw, cosi, sini, L):
    """Mean anomalies (M_left, M_right, flag) of the two roots of
    rho(f)^2 = L^2 adjacent to mid-transit f0 = pi/2 - omega, where
    rho(f)^2 = [a(1-e^2)/(1+e cos f)]^2 [1 - sin^2 i sin^2(omega+f)].
    M values are returned on the branch continuous with M0 (|M - M0| < pi)."""
    a, e, cosw, sinw, cosi, sini, L = [mp.mpf(x) for x in (a, e, cosw, sinw, cosi, sini, L)]
    w = mp.atan2(sinw, cosw)
    f0 = mp.pi / 2 - w

    def g(f):
        rr = a * (1 - e * e) / (1 + e * mp.cos(f))
        return rr * rr * (1 - sini * sini * mp.sin(w + f) ** 2) - L * L

    if g(f0) >= 0:
        return mp.mpf(0), mp.mpf(0), 1
    # march outwards from f0 until sign change (|df| < pi/2 keeps z>0)
    def root(direction):
        lo = f0
        step = mp.mpf("1e-3")
        hi = f0
        while True:
            hi = hi + direction * step
            if abs(hi - f0) >= mp.pi / 2:
                return None
            if g(hi) > 0:
                break
            lo = hi
            step *= 1.5
        return mp.findroot(g, (lo, hi), solver="anderson", tol=mp.mpf(10) ** (-35))

    fl = root(-1)
    fr = root(+1)
    if fl is None or fr is None:
        return mp.mpf(0), mp.mpf(0), 1
    M0 = _f_to_M(f0, e)

    def near(M):
        d = M - M0
        d = d - 2 * mp.pi * mp.floor((d + mp.pi) / (2 * mp.pi))
        return M0 + d

    return near(_f_to_M(fl, e)), near(_f_to_M(fr, e)), 0


# ----------------------------------------------------------------------------
# A.4 celerite SHO log-likelihood: dense definition
# ----------------------------------------------------------------------------
def sho_kernel(tau, S0, w0, Q):
    tau = abs(tau)
    if Q > mp.mpf(1) / 2:
        eta = mp.sqrt(1 - 1 / (4 * Q * Q))
        return S0 * w0 * Q * mp.exp(-w0 * tau / (2 * Q)) * (mp.cos(eta * w0 * tau) + mp.sin(eta * w0 * tau) / (2 * eta * Q))
    if Q == mp.mpf(1) / 2:
        return S0 * w0 * Q * mp.exp(-w0 * tau) * (1 + w0 * tau)
    eta = mp.sqrt(1 / (4 * Q * Q) - 1)
    return S0 * w0 * Q * mp.exp(-w0 * tau / (2 * Q)) * (mp.cosh(eta * w0 * tau) + mp.sinh(eta * w0 * tau) / (2 * eta * Q))


def sho_loglike_dense(t, y, diag, S0, w0, Q):
    n = len(t)
    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            K[i, j] = sho_kernel(mp.mpf(t[i]) - mp.mpf(t[j]), mp.mpf(S0), mp.mpf(w0), mp.mpf(Q))
        K[i, i] += mp.mpf(diag[i])
    L = mp.cholesky(K)
    yv = mp.matrix([mp.mpf(v) for v in y])
    z = mp.lu_solve(L, yv)  # L is triangular; lu_solve is fine for small n
    logdet = 2 * sum(mp.log(L[i, i]) for i in range(n))
    return -(z.T * z)[0] / 2 - logdet / 2 - n * mp.log(2 * mp.pi) / 2


def sho_loglike_dense_grad(t, y, diag, S0, w0, Q):
    """Gradient of the dense log-likelihood by the textbook identities (Rasmussen & Williams eq. 5.9):
       dL/dtheta = 1/2 alpha^T (dK/dtheta) alpha - 1/2 tr(K^-1 dK/dtheta),   alpha = K^-1 y,
       dL/dy = -alpha,   dL/d diag_k = 1/2 alpha_k^2 - 1/2 (K^-1)_kk,
    with dK/dtheta taken by mp.diff of the kernel definition (no hand-derived kernel derivatives).
    Returns (loglike, [dS0, dw0, dQ], gy[n], gdiag[n])."""
    n = len(t)
    S0, w0, Q = mp.mpf(S0), mp.mpf(w0), mp.mpf(Q)
    K = mp.matrix(n, n)
    dK = [mp.matrix(n, n) for _ in range(3)]
    for i in range(n):
        for j in range(i, n):
            tau = mp.mpf(t[i]) - mp.mpf(t[j])
            K[i, j] = K[j, i] = sho_kernel(tau, S0, w0, Q)
            d = [mp.diff(lambda x: sho_kernel(tau, x, w0, Q), S0), mp.diff(lambda x: sho_kernel(tau, S0, x, Q), w0),
                 mp.diff(lambda x: sho_kernel(tau, S0, w0, x), Q)]
            for k in range(3):
                dK[k][i, j] = dK[k][j, i] = d[k]
        K[i, i] += mp.mpf(diag[i])
    Kinv = K ** -1
    yv = mp.matrix([mp.mpf(v) for v in y])
    alpha = Kinv * yv
    L = mp.cholesky(K)
    logdet = 2 * sum(mp.log(L[i, i]) for i in range(n))
    ll = -(yv.T * alpha)[0] / 2 - logdet / 2 - n * mp.log(2 * mp.pi) / 2
    grads = []
    for k in range(3):
        quad = (alpha.T * dK[k] * alpha)[0]
        tr = sum(Kinv[i, j] * dK[k][j, i] for i in range(n) for j in range(n))
        grads.append(quad / 2 - tr / 2)
    gy = [-alpha[i] for i in range(n)]
    gdiag = [alpha[i] ** 2 / 2 - Kinv[i, i] / 2 for i in range(n)]
    return ll, grads, gy, gdiag
