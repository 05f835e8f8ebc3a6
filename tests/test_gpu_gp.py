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
