"""Lag-window gather / scatter: CPU tests pin the oracle restatement of
layers.py:510-571 by its adjoint identity and a hand-checked tiny case; the GPU test compares
the CUDA kernels with it bit for bit (pure data movement / ordered sums)."""
import numpy as np
import pytest

from oracle.lag_oracle import build_rows, get_conv_1D, scatter_rows


def _case(seed=0, lens=(9, 6, 12), X_win=3, X_dim=2, U_win=2, U_dim=3, extra_ctl=2):
    rng = np.random.default_rng(seed)
    Xs = [rng.normal(size=(T, X_dim)) for T in lens]
    Us = [rng.normal(size=(T - X_win + U_win - 1 + extra_ctl, U_dim)) for T in lens] if U_win else None
    return Xs, Us


def test_window_layout_matches_reference_definition():
    x = np.arange(10.0).reshape(5, 2)                 # steps 0..4, dim 2
    w = get_conv_1D(x[:-1], 2)                        # layers.py:519 passes arr[:-1]
    assert w.shape == (3, 2, 2)
    np.testing.assert_array_equal(w.reshape(3, -1), [[0, 1, 2, 3], [2, 3, 4, 5], [4, 5, 6, 7]])
    Xs, Us = [x], [np.arange(12.0).reshape(6, 2)]
    rows = build_rows(Xs, Us, X_win=2, U_win=2)       # N = 3, controls use the last N+U_win-1 = 4 steps
    np.testing.assert_array_equal(rows[0], [0, 1, 2, 3, 4, 5, 6, 7])
    np.testing.assert_array_equal(rows[2], [4, 5, 6, 7, 8, 9, 10, 11])


@pytest.mark.parametrize("U_win", [0, 2])
def test_scatter_is_the_adjoint_of_gather(U_win):
    Xs, Us = _case(U_win=U_win)
    X = build_rows(Xs, Us, 3, U_win)
    g = np.random.default_rng(1).normal(size=X.shape)
    gX, gU = scatter_rows(g, [x.shape for x in Xs], [u.shape for u in Us] if Us else None, 3, U_win, 2, 3)
    lhs = (X * g).sum()
    rhs = sum((a * b).sum() for a, b in zip(Xs, gX)) + (sum((a * b).sum() for a, b in zip(Us, gU)) if Us else 0.0)
    np.testing.assert_allclose(lhs, rhs, rtol=1e-13)
    assert all(np.all(gx[-1] == 0) for gx in gX)      # the last latent step only appears in Y


@pytest.mark.gpu
@pytest.mark.parametrize("lens,X_win,X_dim,U_win,U_dim", [((9, 6, 12), 3, 2, 2, 3), ((512,), 10, 1, 10, 1),
                                                          ((408, 100, 77), 20, 2, 0, 0), ((30, 31), 0, 0, 4, 2)])
def test_cuda_gather_scatter_match_oracle(lens, X_win, X_dim, U_win, U_dim):
    import torch
    from rgp_b200._lib import Handle
    from rgp_b200.lagwindow import LagWindow
    rng = np.random.default_rng(2)
    Xs = [rng.normal(size=(T, max(X_dim, 1))) for T in lens]
    Us = [rng.normal(size=(T - X_win + U_win - 1 + 3, U_dim)) for T in lens] if U_win else None
    X = build_rows([x[:, :X_dim] for x in Xs] if X_dim else Xs, Us, X_win, U_win) if X_win else \
        build_rows(Xs, Us, 0, U_win)
    lw = LagWindow(Handle(0), lens, X_win, X_dim, [u.shape[0] for u in Us] if Us else None, U_win, U_dim)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    lat = t(np.vstack(Xs)[:, :max(X_dim, 1)])
    ctl = t(np.vstack(Us)) if Us else None
    Xd = lw.gather(lat, ctl)
    assert Xd.shape == X.shape
    np.testing.assert_array_equal(Xd.cpu().numpy(), X)
    g = rng.normal(size=X.shape)
    gX, gU = scatter_rows(g, [x.shape for x in Xs], [u.shape for u in Us] if Us else None, X_win, U_win,
                          max(X_dim, 1), U_dim)
    dlat, dctl = lw.scatter_add(t(g))
    if X_win:
        np.testing.assert_allclose(dlat.cpu().numpy(), np.vstack(gX), rtol=1e-15, atol=1e-15)
    if U_win:
        np.testing.assert_allclose(dctl.cpu().numpy(), np.vstack(gU), rtol=1e-15, atol=1e-15)
