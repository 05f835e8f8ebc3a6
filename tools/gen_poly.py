"""Generate the sin/cos kernel polynomials used by exoplanet_b200/csrc/xo_math.cuh.

sin x = x + x^3 * S(x^2),  cos x = 1 - x^2/2 + x^4 * C(x^2)  on |x| <= pi/4 (+2%),
by weighted least squares at Chebyshev nodes in 60-digit arithmetic (close to
minimax; the printed max errors are what matter).  Run: python tools/gen_poly.py
"""
import mpmath as mp

mp.mp.dps = 60


def fit(fn, deg, xmax, npts=400):
    # unknown polynomial in u = x^2, fitted at Chebyshev nodes of u in [0, xmax^2]
    umax = xmax * xmax
    us = [umax * (1 - mp.cos(mp.pi * (k + mp.mpf(1) / 2) / npts)) / 2 for k in range(npts)]
    A = mp.matrix(npts, deg + 1)
    y = mp.matrix(npts, 1)
    for i, u in enumerate(us):
        for j in range(deg + 1):
            A[i, j] = u ** j
        y[i] = fn(u)
    return mp.lu_solve(A.T * A, A.T * y)


def S(u):
    x = mp.sqrt(u)
    return (mp.sin(x) - x) / (x ** 3) if u > 0 else -mp.mpf(1) / 6


def Cc(u):
    x = mp.sqrt(u)
    return (mp.cos(x) - 1 + u / 2) / (u * u) if u > 0 else mp.mpf(1) / 24


xmax = mp.pi / 4 * mp.mpf("1.02")
for name, fn, deg in [("S", S, 6), ("C", Cc, 6)]:
    c = fit(fn, deg, xmax)
    print(name, "= {" + ", ".join(mp.nstr(v, 21) for v in c) + "}")
    # max abs error of the reconstructed function on a fine grid, with fp64-rounded coefficients
    cf = [mp.mpf(float(v)) for v in c]
    worst = 0
    for k in range(2001):
        x = xmax * k / 2000
        u = x * x
        p = sum(cf[j] * u ** j for j in range(deg + 1))
        if name == "S":
            err = abs(x + x ** 3 * p - mp.sin(x))
        else:
            err = abs(1 - u / 2 + u * u * p - mp.cos(x))
        worst = max(worst, err)
    print("  max abs err with fp64 coefficients:", mp.nstr(worst, 3))


# ---- atan x = x + x^3 A(x^2) on |x| <= tan(pi/8) (after the usual two reductions) -------------
def A(u):
    x = mp.sqrt(u)
    return (mp.atan(x) - x) / (x ** 3) if u > 0 else -mp.mpf(1) / 3


def report_atan(deg):
    xm = mp.tan(mp.pi / 8) * mp.mpf("1.01")
    c = fit(A, deg, xm)
    cf = [mp.mpf(float(v)) for v in c]
    worst = 0
    for k in range(2001):
        x = xm * k / 2000
        u = x * x
        p = sum(cf[j] * u ** j for j in range(deg + 1))
        worst = max(worst, abs(x + x ** 3 * p - mp.atan(x)))
    print("atan deg", deg, "max abs err", mp.nstr(worst, 3))
    return c


if __name__ == "__main__":
    for deg in (9, 10, 11, 12):
        c = report_atan(deg)
    print("A = {" + ", ".join(mp.nstr(v, 21) for v in c) + "}")
