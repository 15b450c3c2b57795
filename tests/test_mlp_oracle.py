"""Pins oracle/mlp_oracle.py (MLP back-constraint: free-run recurrence and its back-propagation,
autoreg/layers.py:623-715, autoreg/mlp.py) by finite differences of a scalar functional of the
free-run output - the reference evaluates the network with theano, which is absent."""
import numpy as np
import pytest

from oracle import mlp_oracle as mo


def make_case(seed=0, X_win=3, X_dim=2, U_win=2, U_dim=1, n_steps=(7, 5), control=True, units=None):
    rng = np.random.default_rng(seed)
    Q = X_win * X_dim + (U_win * U_dim if control else 0)
    units = units or mo.default_units(Q, X_dim)
    assert units[0] == Q and units[-1] == X_dim
    params = [(rng.normal(size=(units[i + 1], units[i])) * np.sqrt(2.0 / (units[i] + units[i + 1])),
               rng.normal(size=units[i + 1]) * 0.1) for i in range(len(units) - 1)]
    init = [rng.normal(size=(X_win, X_dim)) * 0.5 for _ in n_steps]
    ctl = [rng.normal(size=(N + U_win - 1 + 2, U_dim)) for N in n_steps] if control else None
    weights = [rng.normal(size=(X_win + N, X_dim)) for N in n_steps]      # F = sum(weights * means)
    return dict(params=params, init=init, ctl=ctl, n_steps=list(n_steps), X_win=X_win, U_win=U_win, weights=weights)


def functional(c):
    X = mo.freerun(c["params"], c["init"], c["ctl"], c["n_steps"], c["X_win"], c["U_win"])
    return sum((w * x).sum() for w, x in zip(c["weights"], X)), X


@pytest.mark.parametrize("control", [True, False])
def test_backprop_through_the_recurrence_matches_finite_differences(control):
    c = make_case(control=control)
    F, X = functional(c)
    g = [w.copy() for w in c["weights"]]
    pg, cg = mo.freerun_backward(c["params"], X, c["ctl"], g, c["X_win"], c["U_win"])
    rng = np.random.default_rng(1)

    def fd(arr, idx, h=1e-6):
        old = arr[idx]
        arr[idx] = old + h; fp = functional(c)[0]
        arr[idx] = old - h; fm = functional(c)[0]
        arr[idx] = old
        return (fp - fm) / (2 * h)

    for l, (W, b) in enumerate(c["params"]):
        for _ in range(3):
            idx = tuple(rng.integers(0, n) for n in W.shape)
            assert abs(fd(W, idx) - pg[l][0][idx]) <= 1e-7 * max(1.0, abs(pg[l][0][idx]))
        j = int(rng.integers(0, b.size))
        assert abs(fd(b, (j,)) - pg[l][1][j]) <= 1e-7 * max(1.0, abs(pg[l][1][j]))
    for s in range(len(c["init"])):
        idx = (int(rng.integers(0, c["X_win"])), 0)
        assert abs(fd(c["init"][s], idx) - g[s][idx]) <= 1e-7 * max(1.0, abs(g[s][idx]))
        if control:
            idx = (c["ctl"][s].shape[0] - 2, 0)
            assert abs(fd(c["ctl"][s], idx) - cg[s][idx]) <= 1e-7 * max(1.0, abs(cg[s][idx]))
            assert np.all(cg[s][:2] == 0)               # unused head of the control series


def test_default_widths_follow_the_reference():
    assert mo.default_units(20, 1) == [20, 40, 20, 1]
    assert mo.default_units(8, 3) == [8, 16, 9, 3]
