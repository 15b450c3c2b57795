"""CPU tests of the oracle itself (no GPU): the known-answer tests of SURVEY.md 8(c).
The reference ships no golden vectors for this path, so the oracle is pinned by the
definition (quadrature), by finite differences, by limits and by the relational
properties the reference's own tests assert (testing/minibatch_tests.py:98-100,281-296)."""
import os

import numpy as np
import pytest

from oracle import bound_oracle as bo
from oracle.psi_oracle import (linear_functional, psi_backward, psi_backward_rowlocal,
                               psi_forward, psi_quadrature)
from synth import make_inputs, make_upstream, relerr

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_kat1_quadrature_matches_closed_forms():
    var, ell, Z, mu, S = make_inputs(7, 5, 3, seed=1)
    _, p1, p2 = psi_forward(var, ell, Z, mu, S)
    q1, q2n = psi_quadrature(var, ell, Z, mu, S, nodes=80)
    assert relerr(p1, q1) < 1e-13
    assert relerr(p2, q2n.sum(axis=0)) < 1e-13


def test_kat1_quadrature_with_tiny_variance_columns():
    var, ell, Z, mu, S = make_inputs(5, 4, 4, seed=2, n_control=2)
    _, p1, p2 = psi_forward(var, ell, Z, mu, S)
    q1, q2n = psi_quadrature(var, ell, Z, mu, S, nodes=60)
    assert relerr(p1, q1) < 1e-12
    assert relerr(p2, q2n.sum(axis=0)) < 1e-12


@pytest.mark.parametrize("ard", [True, False])
def test_kat2_finite_differences_of_linear_functional(ard):
    N, M, Q = 6, 4, 3
    var, ell, Z, mu, S = make_inputs(N, M, Q, seed=3, ard=ard)
    rng = np.random.default_rng(5)
    dL0, dL1, dL2 = rng.normal(size=N), rng.normal(size=(N, M)), rng.normal(size=(M, M))  # dL2 not symmetric
    g = psi_backward(dL0, dL1, dL2, var, ell, Z, mu, S)
    h = 1e-6
    params = [np.array([var]), ell.copy(), Z.copy(), mu.copy(), S.copy()]

    def F(p):
        return linear_functional(dL0, dL1, dL2, float(p[0][0]), p[1], p[2], p[3], p[4])

    for k, gk in enumerate(g):
        gk = np.atleast_1d(np.asarray(gk, dtype=float))
        assert gk.shape == params[k].shape
        fd = np.zeros_like(gk)
        for idx in np.ndindex(gk.shape):
            pp = [a.copy() for a in params]
            pm = [a.copy() for a in params]
            pp[k][idx] += h
            pm[k][idx] -= h
            fd[idx] = (F(pp) - F(pm)) / (2 * h)
        assert relerr(gk, fd) < 1e-7, k


def test_kat3_zero_variance_limit_is_the_plain_kernel():
    # S -> 0: Psi1 -> K(mu,Z), Psi2 -> K^T K  (the observed-input branch, vardtc.py:63-65)
    var, ell, Z, mu, _ = make_inputs(9, 5, 4, seed=4)
    S = np.full_like(mu, 1e-14)
    _, p1, p2 = psi_forward(var, ell, Z, mu, S)
    d2 = (np.square(mu[:, None, :] - Z[None, :, :]) / ell ** 2).sum(-1)
    K = var * np.exp(-0.5 * d2)
    assert relerr(p1, K) < 1e-10
    assert relerr(p2, K.T @ K) < 1e-10


def test_kat4_row_additivity_and_permutation():
    var, ell, Z, mu, S = make_inputs(64, 6, 5, seed=6)
    dL0, dL1, dL2 = make_upstream(64, 6)
    full_f = psi_forward(var, ell, Z, mu, S)
    full_b = psi_backward(dL0, dL1, dL2, var, ell, Z, mu, S)
    for k in (2, 4, 8):
        parts = np.array_split(np.arange(64), k)
        p2 = sum(psi_forward(var, ell, Z, mu[i], S[i])[2] for i in parts)
        assert relerr(p2, full_f[2]) < 1e-13
        bs = [psi_backward(dL0[i], dL1[i], dL2, var, ell, Z, mu[i], S[i]) for i in parts]
        assert abs(sum(b[0] for b in bs) - full_b[0]) < 1e-12 * max(1, abs(full_b[0]))
        assert relerr(sum(b[1] for b in bs), full_b[1]) < 1e-12
        assert relerr(sum(b[2] for b in bs), full_b[2]) < 1e-12
        assert relerr(np.vstack([b[3] for b in bs]), full_b[3]) < 1e-13
    perm = np.random.default_rng(0).permutation(64)
    pf = psi_forward(var, ell, Z, mu[perm], S[perm])
    assert relerr(pf[2], full_f[2]) < 1e-13
    assert relerr(pf[1], full_f[1][perm]) == 0.0


def test_rowlocal_identities_equal_gpy_form():
    var, ell, Z, mu, S = make_inputs(11, 7, 4, seed=8, n_control=1)
    dL0, dL1, dL2 = make_upstream(11, 7)
    a = psi_backward(dL0, dL1, dL2, var, ell, Z, mu, S)
    b = psi_backward_rowlocal(dL0, dL1, dL2, var, ell, Z, mu, S)
    for x, y in zip(a, b):
        assert relerr(x, y) < 1e-13


def test_chunked_forward_equals_unchunked():
    var, ell, Z, mu, S = make_inputs(50, 6, 3, seed=9)
    a = psi_forward(var, ell, Z, mu, S)
    b = psi_forward(var, ell, Z, mu, S, budget_bytes=8 * 36 * 4 * 3)
    for x, y in zip(a, b):
        assert relerr(x, y) < 1e-14


@pytest.mark.parametrize("svi", [False, True])
def test_kat5_checkgrad_through_the_restated_bound(svi):
    N, M, Q, D = 30, 5, 3, 2
    var, ell, Z, mu, S = make_inputs(N, M, Q, seed=10)
    rng = np.random.default_rng(3)
    Y = rng.normal(size=(N, D))
    mode = None
    if svi:
        W = rng.normal(size=(M, M)) * 0.1
        mode = dict(qU_mean=rng.normal(size=(M, D)), qU_var=W @ W.T + 0.5 * np.eye(M), qU_ratio=0.7)

    def f(p):
        return bo.layer_bound_and_grads(float(p[0][0]), p[1], p[2], p[3], p[4], Y, 0.1,
                                        psi_forward, psi_backward, svi=mode)

    params = [np.array([var]), ell, Z, mu, S]
    _, g = f(params)
    names = ["variance", "lengthscale", "Z", "mu", "S"]
    h = 1e-6
    for k, nm in enumerate(names):
        gk = np.atleast_1d(np.asarray(g[nm], dtype=float))
        idxs = list(np.ndindex(gk.shape))[:12]
        for idx in idxs:
            pp = [a.copy() for a in params]
            pm = [a.copy() for a in params]
            pp[k][idx] += h
            pm[k][idx] -= h
            fd = (f(pp)[0] - f(pm)[0]) / (2 * h)
            assert abs(fd - gk[idx]) < 2e-6 * max(1.0, np.abs(gk).max()), (nm, idx)


def test_bound_row_additivity_svi():
    # ELBO of two half minibatches sums to the full one (minibatch_tests.py:288-296)
    N, M, Q, D = 40, 4, 3, 2
    var, ell, Z, mu, S = make_inputs(N, M, Q, seed=12)
    rng = np.random.default_rng(4)
    Y = rng.normal(size=(N, D))
    qm, qv = rng.normal(size=(M, D)), np.eye(M) * 0.3
    Kuu = bo.rbf_K(var, ell, Z)
    p = psi_forward(var, ell, Z, mu, S)
    full, _, _ = bo.svi_vardtc_inference(p[0], p[1], p[2], Kuu, Y, 0.2, qm, qv)
    halves = 0.0
    for sl in (slice(0, 20), slice(20, 40)):
        ph = psi_forward(var, ell, Z, mu[sl], S[sl])
        halves += bo.svi_vardtc_inference(ph[0], ph[1], ph[2], Kuu, Y[sl], 0.2, qm, qv)[0]
    np.testing.assert_allclose(halves, full, rtol=1e-12)


@pytest.mark.parametrize("name", ["tiny_ragged", "actuator_hidden", "actuator_output"])
def test_oracle_reproduces_golden_fixtures(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    p0, p1, p2 = psi_forward(float(g["variance"]), g["ell"], g["Z"], g["mu"], g["S"])
    assert relerr(p1, g["psi1"]) < 1e-13 and relerr(p2, g["psi2"]) < 1e-13
    b = psi_backward(g["dL0"], g["dL1"], g["dL2"], float(g["variance"]), g["ell"], g["Z"], g["mu"], g["S"])
    for x, key in zip(b, ["dvar", "dl", "dZ", "dmu", "dS"]):
        assert relerr(x, g[key]) < 1e-12, key
