"""Seeded synthetic workloads: Kepler long-cadence time stamps, random-init
orbital parameters per chain.  numpy only; shared by tests/ and bench.py so both sides of a
parity test see identical inputs.
"""
import numpy as np

SEED = 20240924


def cadence_times(N, dt=0.0204, t_first=0.0137):
    return np.arange(N, dtype=np.float64) * dt + t_first


def transit_params(C, P=1, seed=SEED, ecc=None, ecc_max=0.7):
    """Per (chain, planet) parameters: log P ~ U(log 3, log 30), t0 ~ U(0, P), e fixed or
    U(0, ecc_max), omega ~ U(-pi, pi), b ~ U(0, 0.9), r/R* ~ logU(0.03, 0.12); per chain
    Kipping (q1, q2) -> (u1, u2) (distributions.py:98-102) and a mean-flux offset."""
    rng = np.random.default_rng(seed)
    period = np.exp(rng.uniform(np.log(3.0), np.log(30.0), (C, P)))
    t0 = rng.uniform(0.0, 1.0, (C, P)) * period
    e = np.full((C, P), float(ecc)) if ecc is not None else rng.uniform(0.0, ecc_max, (C, P))
    omega = rng.uniform(-np.pi, np.pi, (C, P))
    b = rng.uniform(0.0, 0.9, (C, P))
    ror = np.exp(rng.uniform(np.log(0.03), np.log(0.12), (C, P)))
    q1 = rng.uniform(0.0, 1.0, C)
    q2 = rng.uniform(0.0, 1.0, C)
    u1 = 2.0 * np.sqrt(q1) * q2
    u2 = np.sqrt(q1) * (1.0 - 2.0 * q2)
    mean = 1e-4 * rng.normal(size=C)
    m_planet = np.exp(rng.uniform(np.log(1e-5), np.log(1e-3), (C, P)))
    K = rng.uniform(1.0, 30.0, (C, P))  # RV semi-amplitudes, m/s
    gamma = rng.normal(size=C)
    return dict(period=period, t0=t0, ecc=e, omega=omega, b=b, ror=ror, u1=u1, u2=u2, mean=mean, m_planet=m_planet, K=K,
                gamma=gamma)


def noise(N, sigma=5e-4, seed=SEED + 1):
    rng = np.random.default_rng(seed)
    return sigma * rng.normal(size=N)


def kepler_inputs(n, seed=SEED, e_mode="uniform", e_value=0.5):
    """Kepler micro-benchmark inputs: M ~ U(-1e3 2pi, 1e3 2pi); e fixed or ~ U(0, 0.95)."""
    rng = np.random.default_rng(seed)
    M = rng.uniform(-2000.0 * np.pi, 2000.0 * np.pi, n)
    e = rng.uniform(0.0, 0.95, n) if e_mode == "uniform" else np.full(n, float(e_value))
    return M, e


def gp_problem(N, seed=SEED, dt=0.00139, S0=None, w0=None, Q=3.0, sigma_gp=1e-3, rho=5.0, jitter=5e-4):
    """SHOTerm(sigma=1e-3, rho=5 d, Q) + white jitter on a 2-minute grid (config 4)."""
    rng = np.random.default_rng(seed)
    t = np.arange(N, dtype=np.float64) * dt + 0.0137
    if w0 is None:
        w0 = 2.0 * np.pi / rho
    if S0 is None:
        S0 = sigma_gp ** 2 / (w0 * Q)
    y = sigma_gp * np.sin(w0 * t + 0.3) + jitter * rng.normal(size=N)
    diag = np.full(N, jitter ** 2)
    return dict(t=t, y=y, diag=diag, S0=S0, w0=w0, Q=Q)
