"""GPU: celerite SHO scan kernel vs dense-Cholesky goldens and the sequential oracle."""
import numpy as np
import pytest
import torch

import xo_oracle as O
from exoplanet_b200 import gp, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=DEV)


def test_golden_dense(golden):
    g = golden("celerite")
    t, y, diag = np.array(g["t"]), np.array(g["y"]), np.array(g["diag"])
    for c in g["cases"]:
        n = c["n"]
        ll = gp.sho_loglike(dev(t[:n]), dev(y[:n]), dev(diag[:n]), c["S0"], c["w0"], c["Q"]).item()
        # <= 1e-9 relative vs dense Cholesky
        assert abs(ll - c["loglike"]) < 1e-9 * max(1.0, abs(c["loglike"])), c
    assert abs(gp.sho_loglike(dev(t), dev(y), dev(diag), 1.3, 2.1, 3.0).item() - g["appendix_B9"]) < 1e-9


@pytest.mark.parametrize("N", [1, 15, 16, 17, 2047, 2049, 100000, 1000000])
@pytest.mark.parametrize("Q", [3.0, 0.70710678, 0.3])
def test_matches_sequential_oracle(N, Q):
    if N > 200000 and Q != 3.0:
        pytest.skip("one full-size case is enough")
    pr = synthetic.gp_problem(N, Q=Q)
    ll = gp.sho_loglike(dev(pr["t"]), dev(pr["y"]), dev(pr["diag"]), pr["S0"], pr["w0"], pr["Q"]).item()
    lo = O.celerite_sho_loglike(pr["t"], pr["y"], pr["diag"], pr["S0"], pr["w0"], pr["Q"])
    assert abs(ll - lo) < 1e-9 * max(1.0, abs(lo)), (ll, lo)


def test_batch_of_chains_and_irregular_sampling():
    rng = np.random.default_rng(5)
    N, B = 30000, 5
    t = np.sort(rng.uniform(0, 300, N))
    diag = (0.1 + 0.05 * rng.uniform(size=N)) ** 2
    y = rng.normal(size=(B, N)) * 0.3
    S0 = rng.uniform(0.5, 2, B)
    w0 = rng.uniform(0.5, 3, B)
    Q = np.array([3.0, 0.2, 1.0, 0.6, 10.0])
    ll = gp.sho_loglike(dev(t), dev(y), dev(diag), dev(S0), dev(w0), dev(Q)).cpu().numpy()
    for k in range(B):
        lo = O.celerite_sho_loglike(t, y[k], diag, S0[k], w0[k], Q[k])
        assert abs(ll[k] - lo) < 1e-9 * abs(lo)
    ll2 = gp.sho_loglike(dev(t), dev(y), dev(diag), dev(S0), dev(w0), dev(Q)).cpu().numpy()
    assert np.array_equal(ll, ll2)  # deterministic


def _dense_sho(t, diag, S0, w0, Q):
    a, c = S0 * w0 * Q, 0.5 * w0 / Q
    tau = np.abs(t[:, None] - t[None, :])
    if Q > 0.5:
        h = c * np.sqrt(4 * Q * Q - 1)
        k = a * np.exp(-c * tau) * (np.cos(h * tau) + c / h * np.sin(h * tau))
    else:
        h = c * np.sqrt(1 - 4 * Q * Q)  # two decaying exponentials (cosh/sinh would overflow at large tau)
        k = 0.5 * a * ((1 + c / h) * np.exp(-(c - h) * tau) + (1 - c / h) * np.exp(-(c + h) * tau))
    return k + np.diag(diag)


def _dense_loglike(t, y, diag, S0, w0, Q):
    K = _dense_sho(t, diag, S0, w0, Q)
    L = np.linalg.cholesky(K)
    z = np.linalg.solve(L, y)
    return -0.5 * z @ z - np.log(np.diag(L)).sum() - 0.5 * len(t) * np.log(2 * np.pi)


@pytest.mark.parametrize("Q", [3.0, 0.3])
@pytest.mark.parametrize("N", [1, 17, 300, 2500])
def test_solve_matches_dense(N, Q):
    """alpha = K^-1 y from the reverse affine scan vs a dense solve (definition)."""
    rng = np.random.default_rng(N)
    t = np.sort(rng.uniform(0, 0.15 * N + 1, N))
    diag = (0.1 + 0.05 * rng.uniform(size=N)) ** 2
    y = np.sin(t) + 0.1 * rng.normal(size=N)
    S0, w0 = 1.3, 2.1
    ll, alpha = gp.sho_solve(dev(t), dev(y), dev(diag), S0, w0, Q, eps=0.0)
    ref = np.linalg.solve(_dense_sho(t, diag, S0, w0, Q), y)
    assert np.max(np.abs(alpha.cpu().numpy() - ref)) < 1e-9 * (1 + np.abs(ref).max())
    assert abs(ll.item() - _dense_loglike(t, y, diag, S0, w0, Q)) < 1e-9 * (1 + abs(ll.item()))


def test_solve_large_batch_is_consistent():
    """N = 1e5, several chains: K alpha = y checked through a second, independent property --
    d loglike / d y = -alpha verified by a directional finite difference of the forward scan."""
    pr = synthetic.gp_problem(100000)
    t, diag = dev(pr["t"]), dev(pr["diag"])
    rng = np.random.default_rng(8)
    B = 3
    y = dev(np.stack([pr["y"] * (1 + 0.1 * k) for k in range(B)]))
    S0 = dev([pr["S0"], 2 * pr["S0"], 0.5 * pr["S0"]])
    w0 = dev([pr["w0"], 1.5 * pr["w0"], pr["w0"]])
    Q = dev([3.0, 0.4, 1.0])
    ll, alpha = gp.sho_solve(t, y, diag, S0, w0, Q)
    d = dev(rng.normal(size=(B, 100000)) * 1e-4)
    h = 1e-3
    lp = gp.sho_loglike(t, y + h * d, diag, S0, w0, Q)
    lm = gp.sho_loglike(t, y - h * d, diag, S0, w0, Q)
    fd = (lp - lm) / (2 * h)
    an = -(alpha * d).sum(dim=1)
    assert torch.allclose(fd, an, rtol=1e-7, atol=1e-6)
    ll2, alpha2 = gp.sho_solve(t, y, diag, S0, w0, Q)
    assert torch.equal(alpha, alpha2) and torch.equal(ll, ll2)  # deterministic
    assert torch.allclose(ll, gp.sho_loglike(t, y, diag, S0, w0, Q), rtol=1e-14)


def test_gradient_golden(golden):
    """xo_celerite_sho_grad vs the dense gradient identities evaluated at 40 digits"""
    g = golden("celerite")
    t, y, diag = np.array(g["t"]), np.array(g["y"]), np.array(g["diag"])
    for c in g["cases"]:
        if "grad" not in c:
            continue
        gr = c["grad"]
        n = gr["n"]
        ll, gh, gy, gd = gp.sho_grad(dev(t[:n]), dev(y[:n]), dev(diag[:n]), c["S0"], c["w0"], c["Q"], eps=0.0)
        assert abs(ll.item() - gr["loglike"]) < 1e-10 * max(1.0, abs(gr["loglike"]))
        ref = np.array(gr["hyper"])
        assert np.all(np.abs(gh.cpu().numpy() - ref) < 1e-9 * (1.0 + np.abs(ref))), (c["Q"], gh, ref)
        assert np.max(np.abs(gy.cpu().numpy() - np.array(gr["gy"]))) < 1e-9 * (1.0 + np.max(np.abs(gr["gy"])))
        assert np.max(np.abs(gd.cpu().numpy() - np.array(gr["gdiag"]))) < 1e-9 * (1.0 + np.max(np.abs(gr["gdiag"])))


@pytest.mark.parametrize("Q", [3.0, 0.70710678, 0.3, 0.5])
@pytest.mark.parametrize("N", [1, 2, 15, 16, 17, 300, 2049, 4100])
def test_gradient_matches_oracle(N, Q):
    """the reverse affine scan (runs of 16, CTA boundaries at 2048) vs the sequential adjoint of the oracle;
    Q = 0.5 with eps = 1e-5 exercises the clamped (zero-derivative) branch of celerite2's SHOTerm"""
    rng = np.random.default_rng(100 + N)
    t = np.sort(rng.uniform(0, 0.15 * N + 1, N))
    diag = (0.1 + 0.05 * rng.uniform(size=N)) ** 2
    y = np.sin(t) + 0.1 * rng.normal(size=N)
    S0, w0 = 1.3, 2.1
    eps = 1e-5 if Q == 0.5 else 0.0
    ll, gh, gy, gd = gp.sho_grad(dev(t), dev(y), dev(diag), S0, w0, Q, eps=eps)
    lo, go, gyo, gdo = O.celerite_sho_grad(t, y, diag, S0, w0, Q, eps=eps)
    assert abs(ll.item() - lo) < 1e-10 * (1 + abs(lo))
    assert np.all(np.abs(gh.cpu().numpy() - go) < 1e-9 * (1 + np.abs(go))), (gh, go)
    assert np.max(np.abs(gy.cpu().numpy() - gyo)) < 1e-9 * (1 + np.abs(gyo).max())
    assert np.max(np.abs(gd.cpu().numpy() - gdo)) < 1e-9 * (1 + np.abs(gdo).max())


def test_gradient_large_batch():
    """N = 1e5, three chains: d/dy equals -alpha of the disturbance smoother (independent kernel), the
    hyper-parameter and diag gradients agree with central differences of the forward scan, and the
    result is bit-reproducible."""
    N = 100000
    pr = synthetic.gp_problem(N)
    t, diag = dev(pr["t"]), dev(pr["diag"])
    B = 3
    y = dev(np.stack([pr["y"] * (1 + 0.1 * k) for k in range(B)]))
    th = [dev([pr["S0"], 2 * pr["S0"], 0.5 * pr["S0"]]), dev([pr["w0"], 1.5 * pr["w0"], pr["w0"]]), dev([3.0, 0.4, 1.0])]
    ll, gh, gy, gd = gp.sho_grad(t, y, diag, *th)
    ll_s, alpha = gp.sho_solve(t, y, diag, *th)
    assert torch.allclose(ll, ll_s, rtol=1e-14)
    assert torch.allclose(gy, -alpha, rtol=1e-9, atol=1e-9 * alpha.abs().max().item())
    for i in range(3):
        h = 1e-6 * th[i]
        p = [x.clone() for x in th]
        m = [x.clone() for x in th]
        p[i] = p[i] + h
        m[i] = m[i] - h
        fd = (gp.sho_loglike(t, y, diag, *p) - gp.sho_loglike(t, y, diag, *m)) / (2 * h)
        assert torch.allclose(gh[:, i], fd, rtol=2e-5, atol=1e-6 * gh[:, i].abs().max().item()), (i, gh[:, i], fd)
    d = dev(np.random.default_rng(3).uniform(0.5, 1.5, N)) * diag

    def cd(hh):
        return (gp.sho_loglike(t, y, diag + hh * d, *th) - gp.sho_loglike(t, y, diag - hh * d, *th)) / (2 * hh)

    fd = (4.0 * cd(1e-4) - cd(2e-4)) / 3.0  # Richardson: the h^2 term is ~1e-6 relative here, the sum cancels 100:1
    assert torch.allclose((gd * d).sum(1), fd, rtol=1e-6, atol=2e-3)
    ll2, gh2, gy2, gd2 = gp.sho_grad(t, y, diag, *th)
    assert torch.equal(gh, gh2) and torch.equal(gy, gy2) and torch.equal(gd, gd2)


def test_autograd_gradients_match_dense():
    """torch autograd through sho_loglike_autograd: y, (S0, w0, Q) and diag (a shared jitter term)"""
    rng = np.random.default_rng(4)
    N = 200
    t = np.sort(rng.uniform(0, 30, N))
    sig2 = (0.1 + 0.05 * rng.uniform(size=N)) ** 2
    ynp = np.sin(t) + 0.1 * rng.normal(size=N)
    th = np.array([1.3, 2.1, 3.0])
    y = dev(ynp).requires_grad_(True)
    S0, w0, Q = [dev(v).reshape(()).requires_grad_(True) for v in th]
    logj = dev(np.log(0.05)).reshape(()).requires_grad_(True)
    diag = dev(sig2) + torch.exp(2.0 * logj)
    ll = gp.sho_loglike_autograd(dev(t), y, diag, S0, w0, Q, eps=0.0)
    gy, g0, g1, g2, gj = torch.autograd.grad(ll, (y, S0, w0, Q, logj))
    dnp = sig2 + 0.05 ** 2
    lo, go, gyo, gdo = O.celerite_sho_grad(t, ynp, dnp, *th, eps=0.0)
    assert abs(ll.item() - _dense_loglike(t, ynp, dnp, *th)) < 1e-9 * (1 + abs(lo))
    ref_alpha = np.linalg.solve(_dense_sho(t, dnp, *th), ynp)
    assert np.max(np.abs(gy.cpu().numpy() + ref_alpha)) < 1e-9 * (1 + np.abs(ref_alpha).max())
    for i, g in enumerate((g0, g1, g2)):
        assert abs(g.item() - go[i]) < 1e-9 * (1 + abs(go[i])), (i, g.item(), go[i])
        hh = 1e-6 * th[i]
        p, m = th.copy(), th.copy()
        p[i] += hh
        m[i] -= hh
        fd = (_dense_loglike(t, ynp, dnp, *p) - _dense_loglike(t, ynp, dnp, *m)) / (2 * hh)
        assert abs(g.item() - fd) < 1e-6 * (1 + abs(fd)), (i, g.item(), fd)
    assert abs(gj.item() - gdo.sum() * 2 * 0.05 ** 2) < 1e-9 * (1 + abs(gj.item()))


def test_transit_plus_gp_joint_gradient():
    """BASELINE configs[3]: transit mean model + SHO GP noise.  d logL / d (orbital parameters)
    = transit-light-curve VJP with cotangent alpha = K^-1 (y - mu); checked by finite differences."""
    from exoplanet_b200.light_curves import LimbDarkLightCurve
    from exoplanet_b200.orbits import KeplerianOrbit

    N = 20000
    t_np = synthetic.cadence_times(N)
    t = dev(t_np)
    rng = np.random.default_rng(12)
    diag = dev(np.full(N, 5e-4 ** 2))
    y = dev(1e-3 * np.sin(2 * np.pi * t_np / 5.0) + 5e-4 * rng.normal(size=N))
    S0, w0, Q = gp.sho_params_from_sigma_rho(1e-3, 5.0, 3.0)

    def loglike(period, b, ror):
        orbit = KeplerianOrbit(period=period, t0=dev([1.3]), ecc=dev([0.3]), omega=dev([0.4]), b=b)
        mu = LimbDarkLightCurve(0.3, 0.2, device=DEV).get_light_curve(orbit=orbit, r=ror, t=t)[:, 0]
        return gp.sho_loglike_autograd(t, y - mu, diag, S0, w0, Q)

    ps = [dev([7.31]).requires_grad_(True), dev([0.3]).requires_grad_(True), dev([0.08]).requires_grad_(True)]
    grads = torch.autograd.grad(loglike(*ps), ps)
    for i, h in enumerate([1e-7, 1e-6, 1e-7]):
        base = ps[i].detach().clone()
        with torch.no_grad():
            ps[i].copy_(base + h)
            fp = loglike(*ps).item()
            ps[i].copy_(base - h)
            fm = loglike(*ps).item()
            ps[i].copy_(base)
        fd = (fp - fm) / (2 * h)
        assert abs(grads[i].item() - fd) <= 2e-4 * abs(fd) + 1e-3, (i, grads[i].item(), fd)


def test_status_word_counts_non_finite_lanes_and_no_expired_waits():
    """the device status word (xo_read_status): clean input -> (0, 0); a NaN in y or a non-positive variance is
    counted, NaN in -> NaN out, and no inter-CTA wait of the one-pass scans ever expires"""
    from exoplanet_b200 import _lib

    pr = synthetic.gp_problem(300000)
    lib = _lib.require_cuda()
    t, y, diag = dev(pr["t"]), dev(pr["y"]), dev(pr["diag"])
    ws = torch.empty(lib.xo_celerite_sho_grad_workspace(t.numel(), 2) // 8 + 1, dtype=torch.float64, device=DEV)
    yb = torch.stack([y, y])
    ll, gh, gy, gd = gp.sho_grad(t, yb, diag, pr["S0"], pr["w0"], pr["Q"], workspace=ws)
    assert gp.read_status(ws) == (0, 0)
    assert torch.isfinite(ll).all()
    yb[1, 12345] = float("nan")
    ll, gh, gy, gd = gp.sho_grad(t, yb, diag, pr["S0"], pr["w0"], pr["Q"], workspace=ws)
    bad, waits = gp.read_status(ws)
    assert waits == 0 and bad > 0
    assert torch.isfinite(ll[0]) and torch.isnan(ll[1])
    d2 = diag.clone()
    d2[777] = -1.0  # D = a - p00 + diag < 0 at that cadence: log of a negative number
    ws1 = torch.empty(lib.xo_celerite_sho_workspace(t.numel(), 1) // 8 + 1, dtype=torch.float64, device=DEV)
    l1 = gp.sho_loglike(t, y, d2, pr["S0"], pr["w0"], pr["Q"], workspace=ws1)
    bad, waits = gp.read_status(ws1)
    assert waits == 0 and bad > 0 and not torch.isfinite(l1)


def test_many_tiles_and_chains_exercise_the_inter_cta_tree():
    """N = 3e6 (1465 forward tiles = 46 groups: two windows of group totals; 2930 reverse tiles) x 3 chains in all
    damping regimes against the sequential oracle, gradient included; bit-reproducible"""
    N = 3000000
    pr = synthetic.gp_problem(N)
    t, diag = dev(pr["t"]), dev(pr["diag"])
    Qs = [3.0, 0.4, 1.0]
    y = dev(np.stack([pr["y"] * (1 + 0.1 * k) for k in range(3)]))
    S0 = dev([pr["S0"], 2 * pr["S0"], 0.5 * pr["S0"]])
    w0 = dev([pr["w0"], 1.5 * pr["w0"], pr["w0"]])
    ll, gh, gy, gd = gp.sho_grad(t, y, diag, S0, w0, dev(Qs))
    ll2, gh2, gy2, gd2 = gp.sho_grad(t, y, diag, S0, w0, dev(Qs))
    assert torch.equal(ll, ll2) and torch.equal(gh, gh2) and torch.equal(gy, gy2) and torch.equal(gd, gd2)
    for k in range(3):
        lo, gho, gyo, gdo = O.celerite_sho_grad_c(pr["t"], pr["y"] * (1 + 0.1 * k), pr["diag"], S0[k].item(), w0[k].item(), Qs[k])
        assert abs(ll[k].item() - lo) < 1e-10 * abs(lo)
        assert np.all(np.abs(gh[k].cpu().numpy() - gho) < 1e-8 * (np.abs(gho) + 1e-6 * np.abs(gho).max())), (k, gh[k], gho)
        assert np.max(np.abs(gy[k].cpu().numpy() - gyo)) < 1e-9 * np.abs(gyo).max()
        assert np.max(np.abs(gd[k].cpu().numpy() - gdo)) < 1e-9 * np.abs(gdo).max()


def _joint_oracle(t, y, diag, p, hyper, jit, texp=None):
    """log-likelihood of the joint transit + SHO GP model composed from the oracle's pieces"""
    el = O.elements_from_params(p["period"], p["t0"], p["ecc"], p["omega"], p["b"], p["ror"], texp=texp)
    lc = O.transit_lc(t, el, O.get_cl(p["u1"], p["u2"]), *O.stencil(texp)).sum(-1)
    out = []
    for c in range(lc.shape[0]):
        r = y - p["mean"][c] - lc[c]
        out.append(O.celerite_sho_grad_c(t, r, diag + jit[c], hyper[0][c], hyper[1][c], hyper[2][c]))
    return out


@pytest.mark.parametrize("P,texp", [(1, None), (2, 0.02)])
def test_fused_transit_gp_entry_point(P, texp):
    """xo_transit_gp_logp_grad: logL, d/d(S0, w0, Q), d/d mean, d/d jitter against the oracle composition; orbital and
    limb-darkening gradients against central differences of the oracle composition and against the torch-level
    composition (light curve -> residual -> sho_loglike_autograd) on the same engine"""
    from exoplanet_b200 import fused

    N, C = 40000, 3
    rng = np.random.default_rng(17 + P)
    t = np.arange(N) * 0.0204 + 0.0137
    p = synthetic.transit_params(C, P, seed=23 + P, ecc_max=0.4)
    el = O.elements_from_params(p["period"][:1], p["t0"][:1], p["ecc"][:1], p["omega"][:1], p["b"][:1], p["ror"][:1], texp=texp)
    truth = O.transit_lc(t, el, O.get_cl(p["u1"][:1], p["u2"][:1]), *O.stencil(texp)).sum(-1)[0]
    y = truth + 5e-4 * rng.normal(size=N) + 1e-3 * np.sin(2 * np.pi * t / 5.0)
    diag = np.full(N, 5e-4 ** 2)
    hyper = [np.array([2e-7, 3e-7, 1e-7]), np.array([1.2, 0.9, 2.0]), np.array([3.0, 0.4, 1.0])]
    jit = np.array([1e-8, 0.0, 5e-8])
    model = fused.TransitGPModel(t, y, diag, texp=texp, device=DEV)
    par = fused.pack_params(dev(p["period"]), dev(p["t0"]), dev(p["ecc"]), dev(p["omega"]), dev(p["b"]), dev(p["ror"]))
    one = torch.ones(C, dtype=torch.float64, device=DEV)
    star = fused.pack_star(one, one, dev(p["u1"]), dev(p["u2"]))
    out = model.logp_grad(par, star, dev(p["mean"]), dev(hyper[0]), dev(hyper[1]), dev(hyper[2]), jitter=dev(jit))
    assert model.status() == (0, 0)
    ref = _joint_oracle(t, y, diag, p, hyper, jit, texp)
    for c in range(C):
        lo, gho, gyo, gdo = ref[c]
        assert abs(out["loglike"][c].item() - lo) < 1e-10 * abs(lo)
        assert np.all(np.abs(out["ghyper"][c].cpu().numpy() - gho) < 1e-8 * (np.abs(gho) + 1e-6 * np.abs(gho).max()))
        assert abs(out["gmean"][c].item() + gyo.sum()) < 1e-8 * np.abs(gyo).sum()
        assert abs(out["gjitter"][c].item() - gdo.sum()) < 1e-8 * np.abs(gdo).sum()
    # orbital parameters: central differences of the oracle composition (chain 0, planet 0)
    # (a coarse pin: the differences of a ~1e5 log-likelihood carry ~1e-4 of relative noise; the exact pin follows)
    for j, (key, h) in enumerate([("period", 1e-7), ("t0", 1e-6), ("ecc", 1e-5), ("omega", 1e-5), ("b", 1e-5), ("ror", 1e-6)]):
        vals = []
        for sgn in (+1, -1):
            q = {k: np.array(v, copy=True) for k, v in p.items()}
            q[key][0, 0] += sgn * h
            vals.append(_joint_oracle(t, y, diag, {k: v[:1] for k, v in q.items()}, [x[:1] for x in hyper], jit[:1], texp)[0][0])
        fd = (vals[0] - vals[1]) / (2 * h)
        got = out["gpar"][0, 0, j].item()
        assert abs(got - fd) < 2e-3 * max(abs(fd), 1.0) + 1e-3 * np.abs(out["gpar"][0, :, j].cpu().numpy()).max(), (key, got, fd)
    # every gradient against the torch-level composition on the same engine
    from exoplanet_b200 import elements

    leaves = [dev(p[k]).requires_grad_(True) for k in ("period", "t0", "ecc", "omega", "b", "ror")]
    u = [dev(p[k]).requires_grad_(True) for k in ("u1", "u2")]
    el_t = elements.elements_from_params(*leaves, texp=texp)
    from exoplanet_b200.light_curves.limb_dark import _FusedLightCurve

    dtk, wk = fused._stencil_dev(fused.exposure_stencil(texp), torch.device(DEV))
    tt = dev(t)
    total = 0.0
    for c in range(C):
        mu = _FusedLightCurve.apply(el_t[c], fused.limb_dark_coeffs(u[0][c:c + 1], u[1][c:c + 1])[0], tt, dtk, wk).sum(-1)
        total = total + gp.sho_loglike_autograd(tt, dev(y) - p["mean"][c] - mu, dev(diag) + jit[c], hyper[0][c], hyper[1][c], hyper[2][c])
    g = torch.autograd.grad(total, leaves + u)
    for j in range(6):
        a, b = out["gpar"][..., j].cpu().numpy(), g[j].cpu().numpy()
        assert np.max(np.abs(a - b)) < 1e-8 * np.abs(b).max() + 1e-12, (j, a, b)
    assert np.max(np.abs(out["gstar"][:, 2].cpu().numpy() - g[6].cpu().numpy())) < 1e-8 * np.abs(g[6].cpu().numpy()).max()
    assert np.max(np.abs(out["gstar"][:, 3].cpu().numpy() - g[7].cpu().numpy())) < 1e-8 * np.abs(g[7].cpu().numpy()).max()
