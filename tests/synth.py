"""Seeded synthetic inputs of SURVEY.md section 8(d) (shared by tests and bench)."""
import numpy as np


def make_inputs(N, M, Q, seed=20240607, n_control=0, ard=True):
    rng = np.random.default_rng(seed)
    mu = rng.normal(size=(N, Q))
    S = rng.uniform(0.01, 0.5, size=(N, Q))
    if n_control:
        S[:, Q - n_control:] = 1e-10            # control inputs carry variance 1e-10 (model.py:65)
    Z = rng.normal(size=(M, Q))
    ell = np.sqrt(Q) * rng.uniform(0.7, 1.4, size=Q if ard else 1)
    variance = 1.3
    return variance, ell, Z, mu, S


def make_upstream(N, M, seed=7):
    rng = np.random.default_rng(seed)
    dL0 = np.full(N, -0.5)
    dL1 = rng.normal(size=(N, M)) / M
    dL2 = rng.normal(size=(M, M)) / (M * M)
    dL2 = 0.5 * (dL2 + dL2.T)
    return dL0, dL1, dL2


def relerr(a, b):
    """max|a-b| / max|b| per array ('Hard parts')."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.abs(b).max()
    return float(np.abs(a - b).max() / (denom if denom > 0 else 1.0))
