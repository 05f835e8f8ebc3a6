"""CPU: host-side logic (constructor algebra, stencils, sharding) -- no GPU arithmetic."""
import numpy as np
import pytest
import torch

import xo_oracle as O
from exoplanet_b200 import fused, synthetic
from exoplanet_b200.orbits import KeplerianOrbit, SimpleTransitOrbit, get_aor_from_transit_duration


def test_constructor_known_answers(golden):
    g = golden("end_to_end")
    b6 = g["B6"]
    o = KeplerianOrbit(period=b6["period"], ecc=b6["ecc"], omega=b6["omega"], device="cpu")
    assert abs(o.a.item() - b6["a"]) < 1e-12 and abs(o.M0.item() - b6["M0"]) < 1e-13
    assert abs(o.t_periastron.item() - b6["t_periastron"]) < 1e-13 and abs(o.n.item() - b6["n"]) < 1e-13
    b7 = g["B7"]
    o = KeplerianOrbit(m_star=b7["m_star"], r_star=b7["r_star"], t0=b7["t0"], period=b7["period"], ecc=b7["ecc"], omega=b7["omega"],
                       Omega=b7["Omega"], incl=b7["incl"], m_planet=b7["m_planet"], device="cpu")
    assert abs(o.a.item() - b7["a"]) < 1e-10 and abs(o.K0.item() - b7["K0"]) < 1e-12


def test_consistent_inputs_roundtrip():  # orbits/keplerian_test.py:377-501 (without astropy units)
    o1 = KeplerianOrbit(period=10.0, m_star=1.1, r_star=0.9, device="cpu")
    o2 = KeplerianOrbit(a=o1.a, m_star=1.1, r_star=0.9, device="cpu")
    assert abs(o2.period.item() - 10.0) < 1e-12
    o3 = KeplerianOrbit(a=o1.a, period=10.0, r_star=0.9, device="cpu")
    assert abs(o3.m_star.item() - 1.1) < 1e-12 and abs(o3.rho_star.item() - o1.rho_star.item()) < 1e-12
    o4 = KeplerianOrbit(period=10.0, rho_star=o1.rho_star, r_star=0.9, device="cpu")
    assert abs(o4.m_star.item() - 1.1) < 1e-12
    o5 = KeplerianOrbit(period=10.0, rho_star=o1.rho_star, m_star=1.1, device="cpu")
    assert abs(o5.r_star.item() - 0.9) < 1e-12
    for bad in (dict(), dict(period=1.0, a=2.0, m_star=1.0), dict(period=1.0, m_star=1.0, r_star=1.0, rho_star=1.0)):
        with pytest.raises(ValueError):
            KeplerianOrbit(device="cpu", **bad)
    with pytest.raises(ValueError):
        KeplerianOrbit(period=1.0, ecc=0.1, device="cpu")  # omega missing
    with pytest.raises(ValueError):
        KeplerianOrbit(period=1.0, b=0.1, incl=1.0, device="cpu")
    with pytest.raises(ValueError):
        KeplerianOrbit(period=1.0, t0=0.0, t_periastron=0.0, device="cpu")


def test_incl_b_duality_and_duration():
    o = KeplerianOrbit(period=5.0, b=0.4, ecc=0.3, omega=0.7, device="cpu")
    o2 = KeplerianOrbit(period=5.0, incl=o.incl, ecc=0.3, omega=0.7, device="cpu")
    assert abs(o2.b.item() - 0.4) < 1e-12
    aor, jac = get_aor_from_transit_duration(torch.tensor(0.12, dtype=torch.float64), torch.tensor(3.4, dtype=torch.float64),
                                             torch.tensor(0.3, dtype=torch.float64), ror=0.05)
    d = torch.tensor(0.12, dtype=torch.float64, requires_grad=True)
    a2, _ = get_aor_from_transit_duration(d, torch.tensor(3.4, dtype=torch.float64), torch.tensor(0.3, dtype=torch.float64), ror=0.05)
    (g,) = torch.autograd.grad(a2, d)
    assert abs(g.item() - jac.item()) < 1e-10 * abs(jac.item())
    with pytest.warns(UserWarning):
        KeplerianOrbit(period=3.4, b=0.3, duration=0.12, device="cpu")


def test_exposure_stencils_match_reference_construction():
    for order in (0, 1, 2):
        for n in (1, 3, 7, 8):
            dt, w = fused.exposure_stencil(0.02, n, order)
            dto, wo = O.stencil(0.02, n, order)  # literal limb_dark.py:181-198 construction
            assert np.allclose(dt, dto, rtol=0, atol=1e-17) and np.allclose(w, wo, rtol=0, atol=1e-16)
    with pytest.raises(ValueError):
        fused.exposure_stencil(0.02, 7, 3)
    dt, w = fused.exposure_stencil(None)
    assert dt.tolist() == [0.0] and w.tolist() == [1.0]


def test_limb_dark_coeffs_match_oracle():
    u1, u2 = np.array([0.3, 0.1]), np.array([0.2, 0.5])
    c = fused.limb_dark_coeffs(torch.tensor(u1), torch.tensor(u2)).numpy()
    assert np.allclose(c, O.get_cl(u1, u2), rtol=1e-15)


def test_simple_orbit_geometry():
    o = SimpleTransitOrbit(period=3.0, duration=0.2, t0=1.0, b=0.3, r_star=1.1, ror=0.1, device="cpu")
    x, y, z = o.get_relative_position(np.array([1.0, 1.05, 2.5, 4.0]))
    assert abs(x[0].item()) < 1e-12 and abs(x[3].item()) < 1e-12
    assert np.allclose(y.numpy(), 0.33)
    assert z[0].item() > 0 and z[2].item() < 0
    inds = o.in_transit(np.linspace(0, 6, 601))
    assert set(inds.tolist()) <= set(np.nonzero(np.abs((np.linspace(0, 6, 601) - 1.0 + 1.5) % 3.0 - 1.5) < 0.1 + 1e-9)[0].tolist())
    with pytest.raises(NotImplementedError):
        o.get_planet_velocity(0.0)


def test_synthetic_is_seeded():
    a = synthetic.transit_params(4, 2, seed=5)
    b = synthetic.transit_params(4, 2, seed=5)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert np.array_equal(synthetic.cadence_times(5), np.arange(5) * 0.0204 + 0.0137)


def test_batched_hmc_gaussian_target_cpu():
    """the lock-step HMC driver on a known target (independent Gaussians, different scales per
    coordinate): step-size and mass adaptation reach the target acceptance, moments are recovered"""
    import torch

    from exoplanet_b200.sampling import BatchedHMC

    torch.manual_seed(0)
    C, D = 256, 3
    mu = torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
    sd = torch.tensor([0.1, 1.0, 10.0], dtype=torch.float64)

    def logp_grad(q):
        z = (q - mu) / sd
        return -0.5 * (z * z).sum(-1), -z / sd

    q0 = mu + sd * torch.randn(C, D, dtype=torch.float64)
    hmc = BatchedHMC(logp_grad, q0, step_size=0.05, n_leapfrog=8, seed=3, use_graph=False)
    hmc.warmup(200)
    out = hmc.sample(200)
    acc = out["accept_prob"].mean().item()
    assert 0.65 < acc < 0.95, acc
    flat = out["q"].reshape(-1, D)
    assert torch.all((flat.mean(0) - mu).abs() < 0.05 * sd), flat.mean(0)
    assert torch.all((flat.std(0) / sd - 1.0).abs() < 0.1), flat.std(0)
    # the mass matrix picked up the scales (variance per coordinate)
    assert torch.all((hmc.inv_mass.median(0).values / sd ** 2 - 1.0).abs() < 0.6)


def test_batched_nuts_gaussian_targets_cpu():
    """lock-step NUTS (tree doubling, multinomial sampling, generalised U-turn checks on every completed
    sub-tree): leaves a product of Gaussians with very different scales and a strongly correlated
    Gaussian invariant -- the second cannot be fixed by the diagonal mass matrix, so it exercises deep trees"""
    import torch

    from exoplanet_b200.sampling import BatchedNUTS

    torch.manual_seed(0)
    C = 192
    mu = torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
    sd = torch.tensor([0.1, 1.0, 10.0], dtype=torch.float64)

    def diag_target(q):
        z = (q - mu) / sd
        return -0.5 * (z * z).sum(-1), -z / sd

    nuts = BatchedNUTS(diag_target, mu + sd * torch.randn(C, 3, dtype=torch.float64), step_size=0.05, max_depth=8, seed=3)
    nuts.warmup(120)
    out = nuts.sample(120)
    flat = out["q"].reshape(-1, 3)
    assert 0.7 < out["accept_prob"].mean().item() < 0.98
    assert torch.all((flat.mean(0) - mu).abs() < 0.06 * sd), flat.mean(0)
    assert torch.all((flat.std(0) / sd - 1.0).abs() < 0.08), flat.std(0)
    assert not nuts.diverging.any() and 1 <= nuts.depth.float().mean().item() <= 5

    cov = torch.tensor([[1.0, 1.9], [1.9, 4.0]], dtype=torch.float64)  # correlation 0.95
    prec = torch.linalg.inv(cov)

    def corr_target(q):
        g = -(q @ prec)
        return 0.5 * (q * g).sum(-1), g

    nuts = BatchedNUTS(corr_target, torch.randn(C, 2, dtype=torch.float64), step_size=0.1, max_depth=8, seed=5)
    nuts.warmup(120)
    out = nuts.sample(150)
    flat = out["q"].reshape(-1, 2)
    est = torch.cov(flat.T)
    assert torch.all((est - cov).abs() < 0.12 * torch.sqrt(torch.outer(torch.diag(cov), torch.diag(cov)))), est
    assert flat.mean(0).abs().max().item() < 0.1
    assert nuts.depth.float().mean().item() > 2.0  # the ridge needs longer trajectories than the isotropic case
    # a chain count of one and a tree capped at depth 1 still work (edge cases of the masking)
    one = BatchedNUTS(corr_target, torch.zeros(1, 2, dtype=torch.float64), step_size=0.3, max_depth=1, seed=1)
    one.warmup(10, adapt_mass=False)
    assert one.sample(5)["q"].shape == (5, 1, 2) and int(one.depth.max()) <= 1


def test_bench_reads_the_container_cpu_quota(monkeypatch):
    """bench.py sizes the CPU arm's OpenMP team to the cgroup's CFS quota (the GPU boxes give a container 16 of 128 CPUs;
    a 64-thread team inside that quota is what made round 1's CPU numbers irreproducible): both cgroup layouts, no limit"""
    import builtins
    import importlib.util
    import io
    import os

    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    real_open = builtins.open

    def fake(files):
        def _open(path, *a, **k):
            if str(path).startswith("/sys/fs/cgroup"):
                if path in files:
                    return io.StringIO(files[path])
                raise OSError(path)
            return real_open(path, *a, **k)
        return _open

    monkeypatch.setattr(builtins, "open", fake({"/sys/fs/cgroup/cpu.max": "1600000 100000\n"}))
    assert bench._cpu_quota() == 16.0
    monkeypatch.setattr(builtins, "open", fake({"/sys/fs/cgroup/cpu.max": "max 100000\n"}))
    assert bench._cpu_quota() is None
    monkeypatch.setattr(builtins, "open", fake({"/sys/fs/cgroup/cpu/cpu.cfs_quota_us": "-1\n", "/sys/fs/cgroup/cpu/cpu.cfs_period_us": "100000\n"}))
    assert bench._cpu_quota() is None
    monkeypatch.setattr(builtins, "open", fake({"/sys/fs/cgroup/cpu/cpu.cfs_quota_us": "350000\n", "/sys/fs/cgroup/cpu/cpu.cfs_period_us": "100000\n"}))
    assert bench._cpu_quota() == 3.5
    monkeypatch.setattr(builtins, "open", fake({}))
    assert bench._cpu_quota() is None
    assert "cpu_quota" in bench._host_state()


def test_partition_apportionment_algorithm():
    """the rule the logp kernel uses to deal its CTAs out to the chains (xo_transit.cu weighted_partition: n_c = ceil(w_c r)
    with the largest r from five bisection steps between the always-feasible (G - C) / W and G / W), restated in numpy:
    never more CTAs than the grid has, at least one per chain, and a heaviest CTA within 25 % of the best possible
    apportionment (greedy minimax) -- on the bench's own weight distribution and on adversarial ones"""
    rng = np.random.default_rng(8)

    def deal(w, G):
        W, C = int(w.sum()), w.size
        lo, hi = (G - C) / W * (1.0 - 2.0 ** -40), G / W
        for _ in range(5):
            r = 0.5 * (lo + hi)
            if np.maximum(1, np.ceil(w * r)).sum() <= G:
                lo = r
            else:
                hi = r
        return np.maximum(1, np.ceil(w * lo)).astype(int)

    def best(w, G):
        n = np.ones(w.size)
        for _ in range(G - w.size):
            n[np.argmax(w / n)] += 1
        return n

    cases = []
    p = synthetic.transit_params(128, 1, seed=synthetic.SEED, ecc=0.5)
    el = O.elements_from_params(p["period"], p["t0"], p["ecc"], p["omega"], p["b"], p["ror"])[:, 0]
    duty = np.clip((el[:, 12] - el[:, 11]) / el[:, 10], 0.0, 1.0)
    cases.append((1 + np.floor(duty * 100000 / 32), 592))
    cases.append((rng.integers(1, 4000, 300).astype(float), 1184))
    cases.append((np.full(64, 7.0), 592))                                   # all equal
    cases.append((np.r_[1e6, np.ones(99)], 592))                            # one chain has all the work
    cases.append((rng.integers(1, 50, 512).astype(float), 1184))            # about two CTAs per chain
    for w, G in cases:
        n = deal(w, G)
        assert n.min() >= 1 and n.sum() <= G, (n.sum(), G)
        heavy, opt = (w / n).max(), (w / best(w, G)).max()
        assert heavy <= 1.25 * opt + 1e-9, (heavy, opt, G, w.size)
    # the bench's distribution: the heaviest CTA within 15 % of the average load (the round's 6-CTAs-per-chain split: 3x)
    w, G = cases[0]
    assert (w / deal(w, G)).max() <= 1.15 * w.sum() / G
