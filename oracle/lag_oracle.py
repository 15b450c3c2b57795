"""CPU oracle for the lag-window build / gradient scatter.
TEST INFRASTRUCTURE ONLY (see oracle/psi_oracle.py).

Restates autoreg/util.py:6-12 (get_conv_1D), autoreg/layers.py:510-526 (_update_conv),
:475-489 (_init_XY stacking) and :552-571 (update_latent_gradients) for one layer, with
plain numpy and the same loops.  ``get_conv_1D`` is pinned bit for bit against the reference's own function executed in
the build container (tests/golden/ref_conv.npz, tests/test_reference_golden.py); the stacking and
the scatter loops by the adjoint identity <gather(x), g> == <x, scatter(g)> (tests/test_lagwindow.py).
"""
import numpy as np


def get_conv_1D(arr, win):
    """util.py:6-12 (as_strided window view), copied out."""
    assert win > 0
    n = arr.shape[0] - win + 1
    return np.stack([arr[i:i + n] for i in range(win)], axis=1)      # (n, win, dim)


def build_rows(Xs, Us, X_win, U_win):
    """layers.py:510-526 + :481-489: list of per-sequence latents (T_s x Dx) and controls
    (T_u x Du) -> stacked row matrix (N x Q)."""
    rows = []
    for i, x in enumerate(Xs):
        N = x.shape[0] - X_win
        parts = []
        if X_win > 0:
            parts.append(get_conv_1D(x[:-1], X_win).reshape(N, -1))
        if Us is not None:
            u = Us[i]
            parts.append(get_conv_1D(u[-N - U_win + 1:], U_win).reshape(N, -1))
        rows.append(np.hstack(parts))
    return np.vstack(rows)


def scatter_rows_into(dX, gX, gU, X_win, U_win, X_dim, U_dim):
    """layers.py:552-571 verbatim in effect: ``+=`` of every X-row gradient onto the latent /
    control gradient arrays it was built from, in place, rows in increasing n."""
    X_offset = 0
    Qx = X_win * X_dim
    for i in range(len(gX)):
        N = gX[i].shape[0] - X_win
        if gU is not None:
            U_offset = -N - U_win + 1 + gU[i].shape[0]
        for n in range(N):
            if X_win > 0:
                gX[i][n:n + X_win] += dX[X_offset + n, :Qx].reshape(-1, X_dim)
            if gU is not None:
                gU[i][U_offset + n:U_offset + n + U_win] += dX[X_offset + n, Qx:].reshape(-1, U_dim)
        X_offset += N


def scatter_rows(dX, Xs_shapes, Us_shapes, X_win, U_win, X_dim, U_dim):
    """``scatter_rows_into`` onto zero-initialised gradients."""
    gX = [np.zeros(s) for s in Xs_shapes]
    gU = [np.zeros(s) for s in Us_shapes] if Us_shapes is not None else None
    scatter_rows_into(dX, gX, gU, X_win, U_win, X_dim, U_dim)
    return gX, gU
