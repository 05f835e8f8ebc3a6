"""TEST INFRASTRUCTURE: regenerate tests/golden/*.json from the 40-digit
mpmath definitions in oracle/mp_truth.py.

    python oracle/make_golden.py            # ~2-3 min

The reference (/root/reference) stores NO golden vectors for this path and its
arithmetic (exoplanet-core) cannot be imported in this image, so these vectors
are computed from the mathematical definition of each op;
inputs are exact fp64 values (json round-trips repr(float)), outputs are the
40-digit results rounded to fp64.
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mp_truth as T  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def f(x):
    return float(x)


def golden_kepler():
    rng = np.random.default_rng(20240924)
    rows = []
    Ms = list(np.linspace(-np.pi, np.pi, 33)) + list(10.0 ** np.linspace(-12, 0, 13)) + [-2.5, 7.0, 100.3, -1234.567, 6283.2,
                                                                                          1.0e5 + 0.123]
    Ms += list(rng.uniform(-2000 * np.pi, 2000 * np.pi, 40))
    for e in [0.0, 1e-8, 0.1, 0.3, 0.5, 0.7, 0.9, 0.95, 0.99, 0.999]:
        for M in Ms:
            M = float(M)
            sf, cf = T.kepler(mp.mpf(M), mp.mpf(e))
            dM, de = T.kepler_jac(mp.mpf(M), mp.mpf(e))
            rows.append([M, e, f(sf), f(cf), f(dM), f(de)])
    # Appendix B.1 known answers
    kat = [[1.0, 0.5], [0.3, 0.9], [2.5, 0.1], [5.5, 0.7]]
    katrows = []
    for M, e in kat:
        E = T.kepler_E(mp.mpf(M), mp.mpf(e))
        sf, cf = T.kepler(mp.mpf(M), mp.mpf(e))
        katrows.append([M, e, f(E), f(sf), f(cf)])
    return {"columns": ["M", "e", "sinf", "cosf", "dfdM", "dfde"], "rows": rows, "known_answers": katrows}


def golden_qsv():
    rng = np.random.default_rng(7)
    pts = []
    for r in [1e-4, 1e-3, 0.01, 0.03, 0.1, 0.3, 0.5, 0.7, 0.9, 1.0, 1.2]:
        bs = list(np.linspace(0, 1 + r + 0.05, 15))
        bs += [r, r * (1 + 1e-9), r * (1 - 1e-9), abs(1 - r), abs(1 - r) + 1e-9, max(abs(1 - r) - 1e-9, 0.0), 1 + r - 1e-9,
               1e-9, 1e-5]
        pts += [(abs(float(b)), float(r)) for b in bs]
    pts += [(float(rng.uniform(0, 1.3)), float(rng.uniform(0, 0.3))) for _ in range(150)]
    pts += [(-0.5, 0.1), (-0.95, 0.1), (-1.05, 0.1)]  # negative b (tests/light_curves_test.py:24,44)
    # the reference's singular-point set (tests/light_curves_test.py:220-254)
    pts += [(0.9, 0.1), (0.5, 0.5), (0.0, 0.1), (0.0, 1.0), (1.1, 0.1)]
    rows = []
    for b, r in pts:
        bb, rr = mp.mpf(b), mp.mpf(r)
        s = T.quad_solution_vector_closed(bb, rr)
        jb, jr = T.quad_solution_vector_jac_closed(bb, rr)
        sg = -1 if b < 0 else 1
        rows.append([b, r] + [f(x) for x in s] + [f(sg * x) for x in jb] + [f(x) for x in jr])
    # quadrature cross-check subset (definition, not closed form)
    quad = []
    for b, r in [(0.5, 0.1), (0.95, 0.1), (1.05, 0.1), (0.0, 0.1), (0.1, 0.9), (0.5, 0.5), (0.3, 1.2), (0.05, 0.3), (1.5, 0.1),
                 (0.2, 1.5), (0.3, 0.3), (0.8, 0.2)]:
        s = T.quad_solution_vector(b, r)
        quad.append([b, r] + [f(x) for x in s])
    return {"columns": ["b", "r", "s0", "s1", "s2", "ds0db", "ds1db", "ds2db", "ds0dr", "ds1dr", "ds2dr"], "rows": rows,
            "quadrature": quad,
            "note": "Jacobian rows within 1e-7 of b=1-r or b=1+r are sqrt-singular; compare those with a loose tolerance"}


def golden_contact():
    rows = []
    G = 2942.2062175044193
    rng = np.random.default_rng(11)
    cases = [
        # (period, ecc, omega, b, ror, r_star, m_star, m_planet)
        (3.456, 0.6, -1.5, 0.0, 0.1, 1.0, 1.0, 0.0),  # light_curves_test.py:149
        (5.3, 0.8, 1.3, 0.0, 0.02, 1.5, 1.45, 0.5),  # keplerian_test.py:257-271 planet 2
        (2.3, 0.1, 0.4, 0.0, 0.02, 1.5, 1.45, 0.01),
        (10.0, 0.0, 0.0, 0.3, 0.05, 1.0, 1.0, 0.0),
    ]
    for _ in range(24):
        cases.append((float(np.exp(rng.uniform(np.log(3), np.log(30)))), float(rng.uniform(0, 0.7)), float(rng.uniform(-np.pi, np.pi)),
                      float(rng.uniform(0, 0.9)), float(np.exp(rng.uniform(np.log(0.03), np.log(0.12)))), 1.0, 1.0, 0.0))
    for (P, e, w, b, ror, rs, ms, mpl) in cases:
        a = (G * (ms + mpl) * P ** 2 / (4 * np.pi ** 2)) ** (1.0 / 3)
        cw, sw = float(np.cos(w)), float(np.sin(w))
        incl_factor = (1 + e * sw) / (1 - e ** 2)
        cosi = incl_factor * rs / a * b
        sini = float(np.sin(np.arccos(cosi)))
        L = rs * (1 + ror)
        Ml, Mr, flag = T.contact_points(a, e, cw, sw, cosi, sini, L)
        n = 2 * np.pi / P
        E0 = 2 * np.arctan2(np.sqrt(1 - e) * cw, np.sqrt(1 + e) * (1 + sw))
        M0 = E0 - e * np.sin(E0)
        ts = (float(Ml) - M0) / n
        te = (float(Mr) - M0) / n
        rows.append([float(a), e, cw, sw, float(cosi), sini, float(L), f(Ml), f(Mr), int(flag), float(M0), float(n), ts, te])
    return {"columns": ["a", "e", "cosw", "sinw", "cosi", "sini", "L", "M_left", "M_right", "flag", "M0", "n", "t_start", "t_end"],
            "rows": rows, "note": "M_left/M_right on the branch continuous with M0; compare modulo 2 pi"}


def golden_celerite():
    out = []
    rng = np.random.default_rng(42)
    N = 300
    t = np.sort(rng.uniform(0, 30, N))
    sig = 0.1 + 0.05 * rng.uniform(size=N)
    y = np.sin(t) + sig * rng.normal(size=N)
    for (S0, w0, Q, n) in [(1.3, 2.1, 3.0, 120), (1.3, 2.1, 1 / np.sqrt(2), 120), (0.7, 1.1, 0.3, 120), (1.3, 2.1, 3.0, 300)]:
        ll = T.sho_loglike_dense(t[:n], y[:n], (sig ** 2)[:n], S0, w0, Q)
        case = {"S0": S0, "w0": w0, "Q": float(Q), "n": n, "loglike": f(ll)}
        if n <= 120:  # gradient goldens (dense identities at 40 digits) on a 60-point prefix
            m = 60
            llg, gth, gy, gd = T.sho_loglike_dense_grad(t[:m], y[:m], (sig ** 2)[:m], S0, w0, Q)
            case["grad"] = {"n": m, "loglike": f(llg), "hyper": [f(v) for v in gth], "gy": [f(v) for v in gy],
                            "gdiag": [f(v) for v in gd]}
        out.append(case)
    return {"seed": 42, "N": N, "t": t.tolist(), "y": y.tolist(), "diag": (sig ** 2).tolist(), "cases": out,
            "appendix_B9": 2.5488694069}


def golden_end_to_end():
    """SURVEY Appendix B.5-B.8 constants (independently computed at 30-40 digits)."""
    return {
        "B5": {"u": [0.2, 0.3], "cl": [0.12612278509169064, 0.28828065163815004, -0.027026311091076566], "b": 0.5, "r": 0.1,
               "lc": -0.0109428062581156},
        "B6": {"period": 3.456, "ecc": 0.6, "omega": -1.5, "u": [0.3, 0.2], "r": 0.1, "a": 9.6195246813286831,
               "M0": 2.9156100962788655, "t_periastron": -1.6037006709361017, "n": 1.81805130416076,
               "t": [-0.10, -0.05, 0.0, 0.03, 0.12, 1.70],
               "b": [0.876518258356369, 0.438712619646992, 0.0, 0.263285406785418, 1.0512243309427, 4.3857],
               "lc": [-0.00901008729818381, -0.0111501469127744, -0.0115297739116249, -0.0114032405215315, -0.00130762597469262,
                      0.0],
               "t_start": -0.12559732590425, "t_end": 0.12559110744303, "M_left": 2.6872677141195399},
        "B7": {"m_star": 1.3, "r_star": 1.0, "t0": 0.5, "period": 100.0, "ecc": 0.1, "omega": 0.5, "Omega": 1.0,
               "incl": 0.7853981633974483, "m_planet": 0.1, "a": 101.4254950069737, "K0": 4.5748975330525786,
               "t": [0.0, 10.0, 50.0, 99.5],
               "rv_rsun_day": [0.0397502766686905, -0.161586016836695, -0.0799786329269859, 0.0511628047858637],
               "relpos_t0": [54.9988275930975, -39.3017670823021, 67.5147529973038],
               "relpos_t10": [81.2528749574821, 18.8086252500341, 58.2095931161422]},
        "B8": {"m_star": 1.45, "r_star": 1.5, "period": 5.3, "ecc": 0.8, "omega": 1.3, "m_planet": 0.5, "b": 0.0, "L": 1.53,
               "a": 15.982077816794024, "M0": 0.018251391858714681, "M_left": -0.014934500002072788,
               "M_right": 0.051317758316664052, "t_start": -0.0279930032719543, "t_end": 0.0278921810609147},
    }


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name, fn in [("kepler", golden_kepler), ("qsv", golden_qsv), ("contact", golden_contact), ("celerite", golden_celerite),
                     ("end_to_end", golden_end_to_end)]:
        with open(os.path.join(OUT, name + ".json"), "w") as fh:
            json.dump(fn(), fh)
        print("wrote", name)
