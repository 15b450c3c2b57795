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


def make_deep_model(seed=3, wins=(0, 2, 3), nDims=(2, 1, 2), seq_lens=(9, 7), U_win=2, U_dim=1,
                    M=5, control=True, svi=False):
    """A small DeepAutoreg_new instance (autoreg/model.py:95-156): observations, hidden
    latents per level (wins[i] + T_s steps), optional controls with variance 1e-10, and one
    parameter dict per layer.  Level 0 is the observed layer (window 0)."""
    rng = np.random.default_rng(seed)
    L = len(wins)
    Ys = [rng.normal(size=(T, nDims[0])) for T in seq_lens]
    Us = None
    if control:        # model.py:57-66: the control series is U_win-1 steps longer than Y
        Us = [(rng.normal(size=(T + U_win - 1, U_dim)), np.full((T + U_win - 1, U_dim), 1e-10)) for T in seq_lens]
    latents = []
    for i in range(1, L):
        latents.append([(rng.normal(size=(wins[i] + T, nDims[i])) * 0.7,
                         rng.uniform(0.02, 0.3, size=(wins[i] + T, nDims[i]))) for T in seq_lens])
    params = []
    for i in range(L):
        top = i == L - 1
        Q = wins[i] * nDims[i] if i > 0 else 0
        Q += (U_win * U_dim if control else 0) if top else wins[i + 1] * nDims[i + 1]
        D = nDims[i]
        p = dict(variance=float(rng.uniform(0.8, 1.6)), lengthscale=np.sqrt(Q) * rng.uniform(0.7, 1.4, size=Q),
                 Z=rng.normal(size=(M, Q)), noise_variance=float(rng.uniform(0.05, 0.3)))
        if svi:
            p.update(qU_mean=rng.normal(size=(M, D)), qU_W=rng.normal(size=(M, M)) * 0.3,
                     qU_a=float(rng.uniform(0.2, 0.6)), qU_ratio=1.0)
        params.append(p)
    return dict(wins=list(wins), Ys=Ys, Us=Us, U_win=U_win, latents=latents, params=params, svi=svi)


def load_actuator_config1(path=None):
    """BASELINE.json config 1 on the real Actuator data (tests/golden/actuator_config1.npz, made
    by tests/golden/make_actuator_config1.py): a make_deep_model-style dict plus the oracle's
    bound and gradients."""
    import os
    if path is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "actuator_config1.npz")
    g = np.load(path)
    params = [dict(variance=float(g["p%d_variance" % i]), lengthscale=g["p%d_lengthscale" % i].copy(),
                   Z=g["p%d_Z" % i].copy(), noise_variance=float(g["p%d_noise_variance" % i])) for i in range(2)]
    m = dict(wins=[0, 10], Ys=[g["Y"].copy()], Us=[(g["U"].copy(), np.full(g["U"].shape, 1e-10))], U_win=10,
             latents=[[(g["lat_mean"].copy(), g["lat_var"].copy())]], params=params, svi=False)
    return m, g


def stack_model(m, to=None):
    """Stacked tensors of a make_deep_model instance (``to``: ndarray -> tensor; default CPU torch)."""
    if to is None:
        import torch
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    Y = to(np.vstack(m["Ys"]))
    latents = [(to(np.vstack([s[0] for s in lvl])), to(np.vstack([s[1] for s in lvl]))) for lvl in m["latents"]]
    controls = None
    if m["Us"] is not None:
        controls = (to(np.vstack([u[0] for u in m["Us"]])), to(np.vstack([u[1] for u in m["Us"]])))
    params = []
    for p in m["params"]:
        params.append({k: (to(np.asarray(v)) if isinstance(v, np.ndarray) else v) for k, v in p.items()})
    return Y, latents, controls, params
