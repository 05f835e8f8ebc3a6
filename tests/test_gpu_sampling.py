"""GPU: the device-resident batched HMC driver on the fused transit posterior."""
import numpy as np
import pytest
import torch

import xo_oracle as O
from exoplanet_b200 import fused, synthetic
from exoplanet_b200.sampling import BatchedHMC, TransitPosterior

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _problem(C=64, N=20000, seed=5):
    rng = np.random.default_rng(seed)
    truth = dict(period=np.array([[4.7]]), t0=np.array([[1.3]]), ecc=np.array([[0.2]]), omega=np.array([[0.6]]), b=np.array([[0.35]]),
                 ror=np.array([[0.08]]), u1=np.array([0.4]), u2=np.array([0.2]), mean=np.array([0.0]))
    t = synthetic.cadence_times(N)
    el = O.elements_from_params(truth["period"], truth["t0"], truth["ecc"], truth["omega"], truth["b"], truth["ror"])
    lc = O.transit_lc(t, el, O.get_cl(truth["u1"], truth["u2"]), np.zeros(1), np.ones(1))[0, :, 0]
    sigma = 3e-4
    y = lc + sigma * rng.normal(size=N)
    data = fused.TransitData(t, y, np.full(N, sigma), device=DEV)
    q_true = TransitPosterior.pack(*[np.repeat(truth[k], C, axis=0) for k in ("period", "t0", "ecc", "omega", "b", "ror", "u1", "u2", "mean")])
    D = q_true.shape[1]
    prior_sigma = torch.tensor([1e-3, 0.02, 0.3, 0.3, 0.3, 0.3, 1.5, 1.5, 1e-3], dtype=torch.float64)
    post = TransitPosterior(fused.TransitModel(data), 1, q_true[0], prior_sigma)
    scatter = torch.tensor([1e-5, 2e-4, 0.02, 0.02, 0.02, 0.005, 0.1, 0.1, 1e-5], dtype=torch.float64)
    q0 = (q_true + scatter * torch.as_tensor(rng.normal(size=(C, D)))).to(DEV)
    return post, q0, q_true[0].to(DEV), scatter


def test_posterior_gradient_matches_finite_differences():
    post, q0, _, scale = _problem(C=4)
    lp, g = post(q0)
    for j in range(q0.shape[1]):
        h = 1e-3 * scale[j].item()  # a small fraction of the posterior width of that coordinate (log period: ~1e-5)
        qp, qm = q0.clone(), q0.clone()
        qp[:, j] += h
        qm[:, j] -= h
        fd = (post(qp)[0] - post(qm)[0]) / (2 * h)
        assert torch.allclose(g[:, j], fd, rtol=2e-5, atol=1e-6 * g[:, j].abs().max().item() + 1e-4), (j, g[:, j], fd)


def test_fused_posterior_equals_autograd_composition():
    """xo_transit_posterior_logp_grad (unpack kernel, fused model, chain rule + prior kernel) against the same posterior
    assembled from torch ops and torch.autograd around the fused model"""
    post, q0, _, _ = _problem(C=16)
    assert post.fused_args(16) is not None
    lp, g = post(q0)
    lp2, g2 = post.autograd_call(q0)
    assert torch.allclose(lp, lp2, rtol=1e-13, atol=0)
    assert torch.allclose(g, g2, rtol=1e-10, atol=1e-10 * g2.abs().max().item())
    # per-chain priors ([C, D] rows) take the same entry point with a chain stride
    from exoplanet_b200.sampling import TransitPosterior

    mu_c = post.mu.expand(16, -1).clone()
    mu_c[3:] += 0.01
    per_chain = TransitPosterior(post.model, 1, mu_c, post.sigma)
    assert per_chain.fused_args(16) is not None
    lp3, g3 = per_chain(q0)
    lp4, g4 = per_chain.autograd_call(q0)
    assert torch.allclose(lp3, lp4, rtol=1e-13, atol=0) and torch.allclose(g3, g4, rtol=1e-10, atol=1e-10 * g4.abs().max().item())
    assert torch.allclose(lp3[:3], lp2[:3], rtol=1e-13, atol=0) and not torch.allclose(lp3[3:], lp2[3:], rtol=1e-9, atol=0)


@pytest.mark.parametrize("mode", ["fused", "graph", "eager"])
def test_batched_hmc_recovers_transit_parameters(mode):
    """64 chains in lock-step: one C call per transition (xo_hmc_transition), the torch transition replayed from a CUDA
    graph, or eager torch launches"""
    post, q0, q_true, _ = _problem()
    hmc = BatchedHMC(post, q0, step_size=2e-3, n_leapfrog=8, seed=11, use_graph=(mode == "graph"), fused=None if mode == "fused" else False)
    assert (hmc._fused is not None) == (mode == "fused")
    hmc.warmup(160)
    out = hmc.sample(120)
    torch.cuda.synchronize()
    acc = out["accept_prob"].mean().item()
    assert 0.55 < acc <= 1.0, acc
    assert torch.isfinite(out["logp"]).all()
    assert out["energy_error"].abs().median().item() < 1.0
    flat = out["q"].reshape(-1, q0.shape[1])
    mean, sd = flat.mean(0), flat.std(0)
    # log(r/R*) and log(period) are tightly constrained by the data: recovered within the posterior width
    for j in (0, 5):
        assert abs(mean[j] - q_true[j]).item() < 4.0 * sd[j].item() + 1e-9, (j, mean[j], q_true[j], sd[j])
    assert sd[5].item() < 0.05  # the radius ratio is measured to a few per cent
    # the chains moved (not stuck at their starting points)
    assert (out["q"][-1] - q0).abs().max().item() > 0
