"""CPU oracle for the RBF-ARD psi-statistics hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product (``rgp_b200``) never
does, and fails loudly when its CUDA library is missing.

PARITY UNPINNED (versus GPy binaries).  The arithmetic of this path lives in the
third-party package GPy (module ``GPy.kern.src.psi_comp.rbf_psi_comp``; not vendored
by, pinned by, or present beside ``/root/reference``; GPy 1.x era per the reference's
``paramz`` import at ``autoreg/util.py:3``).  GPy cannot be imported in this image and
the reference ships no golden vectors (its tests use unseeded random data,
``testing/minibatch_tests.py:16-33``).  This file therefore *restates* GPy's published
closed forms and pins them two independent ways that
do not depend on any recollection of GPy's code:

* ``psi_quadrature``      - Gauss-Hermite expectation of the RBF kernel, the
                            definition of Psi1/Psi2 (known-answer test, <=1e-13);
* finite differences     - ``tests/test_oracle.py`` differentiates the linear
                            functional F = <dL0,psi0> + <dL1,Psi1> + <dL2,Psi2>.

Reference call sites the restated functions stand behind:
  forward   autoreg/inference/vardtc.py:59-61, autoreg/inference/svi_vardtc.py:48-50
  backward  autoreg/layers.py:98-102 (variance, lengthscale), :127-132 (Z),
            :574-580 (q(X) mean / variance)

Two formulations are given on purpose:

``psi_forward`` / ``psi_backward``
    GPy's own structure: a chunk x M x M tensor is materialised and contracted with
    einsum / GEMM.  This is "the reference CPU path" and is what ``bench.py`` times as
    the CPU baseline.
``psi_backward_rowlocal``
    the factorised row-local identities of SURVEY.md section 8(a5) that the CUDA
    kernels implement; kept here so a test proves the two agree to ~1e-15.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "psi_forward", "psi_backward", "psi_backward_rowlocal", "psi_quadrature",
    "psi1_closed", "psi2n_closed", "linear_functional",
]


def _as2d(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ell(lengthscale, Q):
    ell = np.asarray(lengthscale, dtype=np.float64).reshape(-1)
    if ell.size == 1:
        ell = np.full(Q, float(ell[0]))
    assert ell.size == Q
    return ell


# --------------------------------------------------------------------------- forward
def psi1_closed(variance, lengthscale, Z, mu, S):
    """Psi1[n,m] = s2 * exp(-1/2 sum_q log(S/l^2+1) - 1/2 sum_q (mu-Z)^2/(S+l^2)).

    SURVEY.md section 8 row a2 (GPy ``_psi1computations``), reached from
    autoreg/inference/vardtc.py:60.
    """
    Z, mu, S = _as2d(Z), _as2d(mu), _as2d(S)
    ell = _ell(lengthscale, mu.shape[1])
    l2 = ell * ell
    logdenom = np.log(S / l2 + 1.0).sum(axis=1)                      # N
    diff2 = np.square(mu[:, None, :] - Z[None, :, :])                # N x M x Q
    quad = np.einsum("nmq,nq->nm", diff2, 1.0 / (S + l2))
    return float(variance) * np.exp(-0.5 * (logdenom[:, None] + quad))


def psi2n_closed(variance, lengthscale, Z, mu, S):
    """P_n[m,m'] (N x M x M, not summed).  SURVEY.md section 8 row a3.

    GPy ``_psi2computations`` builds the exponent from three pieces:
      -1/2 sum_q log(2S/l^2+1)                      (row term)
      -sum_q (Z_m - Z_m')^2 / (4 l^2)               (M x M, row independent)
      -sum_q (mu - (Z_m+Z_m')/2)^2 / (2S + l^2)     (expanded as two GEMMs against Zhat)
    """
    Z, mu, S = _as2d(Z), _as2d(mu), _as2d(S)
    N, Q = mu.shape
    M = Z.shape[0]
    ell = _ell(lengthscale, Q)
    l2 = ell * ell
    row = -0.5 * np.log(2.0 * S / l2 + 1.0).sum(axis=1)              # N
    zz = -0.25 * (np.square(Z[:, None, :] - Z[None, :, :]) / l2).sum(axis=2)   # M x M
    zhat = 0.5 * (Z[:, None, :] + Z[None, :, :]).reshape(M * M, Q)   # M^2 x Q
    d = 1.0 / (2.0 * S + l2)                                         # N x Q
    cross = (2.0 * (mu * d) @ zhat.T - d @ np.square(zhat).T).reshape(N, M, M)
    mu2 = (np.square(mu) * d).sum(axis=1)
    expo = row[:, None, None] + zz[None, :, :] + cross - mu2[:, None, None]
    return float(variance) ** 2 * np.exp(expo)


def _row_chunks(N, M, budget_bytes=256 << 20):
    step = max(1, int(budget_bytes // max(1, 8 * M * M * 4)))
    for s in range(0, N, step):
        yield s, min(N, s + step)


def psi_forward(variance, lengthscale, Z, mu, S, budget_bytes=256 << 20):
    """(psi0[N], Psi1[N,M], Psi2[M,M]) - the ``psicomputations`` contract (row a1-a3).

    Rows are chunked so the chunk x M x M tensor stays under ``budget_bytes`` (GPy's
    form materialises N x M x M at once, which is 4.4 TB at the headline shape).
    """
    Z, mu, S = _as2d(Z), _as2d(mu), _as2d(S)
    N = mu.shape[0]
    M = Z.shape[0]
    psi0 = np.full(N, float(variance))
    psi1 = np.empty((N, M))
    psi2 = np.zeros((M, M))
    for s, e in _row_chunks(N, M, budget_bytes):
        psi1[s:e] = psi1_closed(variance, lengthscale, Z, mu[s:e], S[s:e])
        psi2 += psi2n_closed(variance, lengthscale, Z, mu[s:e], S[s:e]).sum(axis=0)
    return psi0, psi1, psi2


# -------------------------------------------------------------------------- backward
def _psi1_grads(dL_dpsi1, variance, ell, Z, mu, S):
    """SURVEY.md section 8 row a4 (GPy ``_psi1compDer``), GPy sign convention Z - mu."""
    l2 = ell * ell
    L1 = dL_dpsi1 * psi1_closed(variance, ell, Z, mu, S)             # N x M
    zmu = Z[None, :, :] - mu[:, None, :]                             # N x M x Q
    e = 1.0 / (S + l2)                                               # N x Q
    v = np.square(zmu) * e[:, None, :]
    dvar = L1.sum() / variance
    dmu = np.einsum("nm,nmq,nq->nq", L1, zmu, e)
    dS = 0.5 * np.einsum("nm,nmq,nq->nq", L1, v - 1.0, e)
    dZ = -np.einsum("nm,nmq,nq->mq", L1, zmu, e)
    dl = np.einsum("nm,nmq,nq->q", L1, v + (S / l2)[:, None, :], e * ell)
    return dvar, dl, dZ, dmu, dS


def _psi2_grads(dL_dpsi2, variance, ell, Z, mu, S):
    """SURVEY.md section 8 row a5 in GPy's structure (``_psi2compDer``): contract the
    stored chunk x M x M tensor against Z and Z^2."""
    N, Q = mu.shape
    M = Z.shape[0]
    l2 = ell * ell
    d = 1.0 / (2.0 * S + l2)
    d2 = d * d
    dLs = 0.5 * (dL_dpsi2 + dL_dpsi2.T)
    L = dLs[None, :, :] * psi2n_closed(variance, ell, Z, mu, S)      # N x M x M
    Lsum = L.reshape(N, M * M).sum(axis=1)                           # N
    LZ_m = (L.reshape(N * M, M) @ Z).reshape(N, M, Q)                # sum_m' L[n,m,m'] Z[m',q]
    LZ = LZ_m.sum(axis=1)                                            # N x Q
    LZ2 = (L.reshape(N * M, M) @ np.square(Z)).reshape(N, M, Q).sum(axis=1)
    LZZ = (LZ_m * Z[None, :, :]).sum(axis=1)                         # sum_mm' L Z_m Z_m'
    LZh2 = 0.5 * (LZ2 + LZZ)                                         # sum L Zhat^2

    dvar = 2.0 * Lsum.sum() / variance
    dmu = (-2.0 * d) * (mu * Lsum[:, None] - LZ)
    dS = 2.0 * d2 * (np.square(mu) * Lsum[:, None] - 2.0 * mu * LZ + LZh2) - d * Lsum[:, None]
    Lm = L.sum(axis=0)                                               # M x M (= dLs * Psi2)
    rows = L.sum(axis=2)                                             # N x M  (lambda_nm)
    dZ = (-(Lm.sum(axis=0)[:, None] * Z) + Lm @ Z) / l2 \
        + 2.0 * np.einsum("nm,nq->mq", rows, mu * d) \
        - np.einsum("nm,nq->mq", rows, d) * Z \
        - np.einsum("nmq,nq->mq", LZ_m, d)
    dl = 2.0 * ell * ((S / l2 * d + np.square(mu * d)) * Lsum[:, None]
                      + (LZ2 - LZZ) / (2.0 * l2 * l2)
                      - 2.0 * mu * d2 * LZ + d2 * LZh2).sum(axis=0)
    return dvar, dl, dZ, dmu, dS


def psi_backward(dL_dpsi0, dL_dpsi1, dL_dpsi2, variance, lengthscale, Z, mu, S,
                 budget_bytes=256 << 20):
    """(dL_dvar, dL_dlengthscale, dL_dZ, dL_dmu, dL_dS) - the
    ``psiDerivativecomputations`` contract.

    A size-1 ``lengthscale`` (non-ARD kernel) gets its gradient summed to a scalar,
    as GPy does.
    """
    Z, mu, S = _as2d(Z), _as2d(mu), _as2d(S)
    N, Q = mu.shape
    M = Z.shape[0]
    ard = np.asarray(lengthscale).size != 1
    ell = _ell(lengthscale, Q)
    dL_dpsi0 = np.broadcast_to(np.asarray(dL_dpsi0, dtype=np.float64), (N,))
    dL_dpsi1 = _as2d(dL_dpsi1)
    dL_dpsi2 = _as2d(dL_dpsi2)
    variance = float(variance)

    dvar = float(dL_dpsi0.sum())
    dl = np.zeros(Q)
    dZ = np.zeros((M, Q))
    dmu = np.empty((N, Q))
    dS = np.empty((N, Q))
    for s, e in _row_chunks(N, M, budget_bytes):
        a = _psi1_grads(dL_dpsi1[s:e], variance, ell, Z, mu[s:e], S[s:e])
        b = _psi2_grads(dL_dpsi2, variance, ell, Z, mu[s:e], S[s:e])
        dvar += a[0] + b[0]
        dl += a[1] + b[1]
        dZ += a[2] + b[2]
        dmu[s:e] = a[3] + b[3]
        dS[s:e] = a[4] + b[4]
    if not ard:
        dl = np.array([dl.sum()])
    return dvar, dl, dZ, dmu, dS


def psi_backward_rowlocal(dL_dpsi0, dL_dpsi1, dL_dpsi2, variance, lengthscale, Z, mu, S):
    """Same outputs as ``psi_backward`` through the factorised row-local identities of
    SURVEY.md section 8(a5) - the algebra the CUDA kernels implement (one row at a
    time, nothing of size N x M x M).  Python loop over rows: small cases only.
    """
    Z, mu, S = _as2d(Z), _as2d(mu), _as2d(S)
    N, Q = mu.shape
    M = Z.shape[0]
    ard = np.asarray(lengthscale).size != 1
    ell = _ell(lengthscale, Q)
    l2 = ell * ell
    s2 = float(variance)
    dL_dpsi0 = np.broadcast_to(np.asarray(dL_dpsi0, dtype=np.float64), (N,))
    dL1 = _as2d(dL_dpsi1)
    dLs = 0.5 * (_as2d(dL_dpsi2) + _as2d(dL_dpsi2).T)
    E1 = -0.25 * (np.square(Z[:, None, :] - Z[None, :, :]) / l2).sum(axis=2)
    C = dLs * (s2 * s2) * np.exp(E1)                                 # row independent

    dvar = float(dL_dpsi0.sum())
    dl = np.zeros(Q)
    dZ = np.zeros((M, Q))
    dmu = np.zeros((N, Q))
    dS = np.zeros((N, Q))
    psi2 = np.zeros((M, M))
    for n in range(N):
        a = mu[n][None, :] - Z                                       # M x Q
        # ---- psi1 part (row a4)
        e = 1.0 / (S[n] + l2)
        p1 = s2 * np.exp(-0.5 * np.log1p(S[n] / l2).sum() - 0.5 * (a * a * e).sum(axis=1))
        L1 = dL1[n] * p1                                             # M
        v = a * a * e
        dvar += L1.sum() / s2
        dmu[n] += -(L1[:, None] * a * e).sum(axis=0)
        dS[n] += 0.5 * (L1[:, None] * (v - 1.0) * e).sum(axis=0)
        dZ += L1[:, None] * a * e
        dl += (L1[:, None] * (v + S[n] / l2) * e * ell).sum(axis=0)
        # ---- psi2 part (row a5)
        d = 1.0 / (2.0 * S[n] + l2)
        at = a * np.sqrt(d)
        r = (at * at).sum(axis=1)                                    # M
        G = at @ at.T
        cn = -0.5 * np.log1p(2.0 * S[n] / l2).sum()
        Pexp = np.exp(cn - 0.25 * (r[:, None] + r[None, :]) - 0.5 * G)
        psi2 += (s2 * s2) * np.exp(E1) * Pexp
        L = C * Pexp                                                 # M x M symmetric
        lam = L.sum(axis=1)
        Lam = lam.sum()
        T = L @ a                                                    # M x Q
        quad = (lam[:, None] * a * a).sum(axis=0) + (a * T).sum(axis=0)
        dvar += 2.0 * Lam / s2
        dmu[n] += -2.0 * d * (lam[:, None] * a).sum(axis=0)
        dS[n] += -d * Lam + d * d * quad
        dZ += d * (lam[:, None] * a + T)
        dl += Lam * 2.0 * S[n] / (ell * (2.0 * S[n] + l2)) + ell * d * d * quad
    # row-independent tails through E1 (uses the summed Psi2)
    LN = dLs * psi2
    rs = LN.sum(axis=1)
    LNZ = LN @ Z
    dZ += -(rs[:, None] * Z - LNZ) / l2
    dl += ((rs[:, None] * Z * Z).sum(axis=0) - (Z * LNZ).sum(axis=0)) / (l2 * ell)
    if not ard:
        dl = np.array([dl.sum()])
    return dvar, dl, dZ, dmu, dS


# ------------------------------------------------------------ independent known answers
def psi_quadrature(variance, lengthscale, Z, mu, S, nodes=80):
    """Psi1 and P_n by Gauss-Hermite quadrature of the *definition*
    Psi1[n,m] = E_{x~N(mu_n,S_n)} k(x, Z_m),  P_n[m,m'] = E k(x,Z_m) k(x,Z_m'),
    k(x,z) = s2 * prod_q exp(-(x_q-z_q)^2 / (2 l_q^2)).  The expectation factorises over
    q, so each factor is a 1-D integral.  Independent of every closed form above.
    Returns (Psi1[N,M], P[N,M,M]).  O(N M^2 Q nodes): tiny cases only.
    """
    Z, mu, S = _as2d(Z), _as2d(mu), _as2d(S)
    N, Q = mu.shape
    M = Z.shape[0]
    ell = _ell(lengthscale, Q)
    t, w = np.polynomial.hermite.hermgauss(nodes)
    w = w / np.sqrt(np.pi)
    psi1 = np.full((N, M), float(variance))
    psi2 = np.full((N, M, M), float(variance) ** 2)
    for q in range(Q):
        x = mu[:, q][:, None] + np.sqrt(2.0 * S[:, q])[:, None] * t[None, :]   # N x nodes
        k = np.exp(-np.square(x[:, None, :] - Z[None, :, q][:, :, None]) / (2.0 * ell[q] ** 2))  # N x M x nodes
        psi1 *= k @ w
        psi2 *= np.einsum("nak,nbk,k->nab", k, k, w)
    return psi1, psi2


def linear_functional(dL_dpsi0, dL_dpsi1, dL_dpsi2, variance, lengthscale, Z, mu, S):
    """F = <dL0,psi0> + <dL1,Psi1> + <dL2,Psi2>; its gradient is ``psi_backward``."""
    p0, p1, p2 = psi_forward(variance, lengthscale, Z, mu, S)
    N = p0.shape[0]
    dL0 = np.broadcast_to(np.asarray(dL_dpsi0, dtype=np.float64), (N,))
    return float((dL0 * p0).sum() + (np.asarray(dL_dpsi1) * p1).sum()
                 + (np.asarray(dL_dpsi2) * p2).sum())
