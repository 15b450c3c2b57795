"""CPU check of the ALGEBRA the tiled CUDA kernels implement (DESIGN.md "Factorised
exponent"): a numpy model of rgp_b200/csrc/fast_prep.cuh + psi2_kernels.cuh, step for
step (centring, w / H factorisation, lambda / W / ACC decomposition, the per-row and
final combiners), compared with the oracle.  It proves the formulas, not the CUDA code
(the GPU parity tests do that), and documents what each device buffer holds."""
import numpy as np
import pytest

from oracle.psi_oracle import psi_backward, psi_forward
from synth import make_inputs, make_upstream, relerr


def fast_model(variance, ell, Z, mu, S, dL0, dL1, dL2):
    N, Q = mu.shape
    l2 = ell ** 2
    o = Z.mean(axis=0)                                    # k_center
    Zc, mc = Z - o, mu - o
    d, e = 1.0 / (2 * S + l2), 1.0 / (S + l2)             # k_rowprep
    ws = S / (l2 * (2 * S + l2))                          # = -w >= 0
    A2 = np.hstack([d * mc, -0.25 * (d + 1 / l2)])
    A1 = np.hstack([e * mc, -0.5 * e])
    b2 = -0.25 * np.log1p(2 * S / l2).sum(1) - 0.5 * (d * mc * mc).sum(1)
    b1 = -0.5 * np.log1p(S / l2).sum(1) - 0.5 * (e * mc * mc).sum(1)
    ZB = np.hstack([Zc, Zc * Zc])                         # k_build_Z
    H = b2[:, None] + A2 @ ZB.T                           # hprime_gemm
    psi1 = variance * np.exp(b1[:, None] + A1 @ ZB.T)     # psi1_fwd
    C = variance ** 2 * 0.5 * (dL2 + dL2.T)               # k_build_C
    L1 = dL1 * psi1                                       # psi1_L1
    M = Z.shape[0]
    psi2 = np.zeros((M, M))
    lam = np.zeros((N, M)); Wq = np.zeros((N, Q)); ACC = np.zeros((M, Q))
    for n in range(N):                                    # k_psi2_fwd / k_psi2_bwd, one row at a time
        E = H[n][:, None] + H[n][None, :] + (Zc * ws[n]) @ Zc.T   # stage 1 (accumulator init + DMMA)
        P = np.exp(E)                                     # epilogue
        psi2 += P
        L = C * P
        lam[n] = L.sum(1)
        T = L @ Zc                                        # stage 2
        ACC += ws[n] * T
        Wq[n] = (Zc * T).sum(0)
    psi2 *= variance ** 2                                 # k_psi2_reduce
    R2, R1 = lam @ ZB, L1 @ ZB                            # rows_gemm
    U, V, LZ, LZ2 = R2[:, :Q], R2[:, Q:], R1[:, :Q], R1[:, Q:]
    Lam, Lam1 = lam.sum(1)[:, None], L1.sum(1)[:, None]   # k_rows_finalize
    quad = 2 * mc * mc * Lam - 4 * mc * U + V + Wq
    dmu = -2 * d * (mc * Lam - U) - e * (mc * Lam1 - LZ)
    B1 = mc * mc * Lam1 - 2 * mc * LZ + LZ2
    dS = -d * Lam + d * d * quad + 0.5 * e * (e * B1 - Lam1)
    dl = (Lam * 2 * S / (ell * (2 * S + l2)) + ell * d * d * quad + (V - Wq) / (l2 * ell)
          + ell * e * (e * B1 + (S / l2) * Lam1)).sum(0)
    dvar = ((2 * Lam + Lam1) / variance).sum() + dL0.sum()
    Gl, GL = lam.T @ A2, L1.T @ A1                        # dz_gemm
    dZ = 2 * Gl[:, :Q] + 4 * Zc * Gl[:, Q:] + 2 * ACC + GL[:, :Q] + 2 * Zc * GL[:, Q:]   # k_final_small
    return (np.full(N, variance), psi1, psi2), (dvar, dl, dZ, dmu, dS)


@pytest.mark.parametrize("N,M,Q,nc,shift", [(9, 6, 3, 0, 0.0), (17, 11, 5, 2, 0.0), (12, 8, 4, 1, 50.0)])
def test_factorised_algebra_matches_oracle(N, M, Q, nc, shift):
    var, ell, Z, mu, S = make_inputs(N, M, Q, seed=3 + N, n_control=nc)
    Z, mu = Z + shift, mu + shift                          # centring keeps this harmless
    dL0, dL1, dL2 = make_upstream(N, M, seed=N)
    dL2 = dL2 + np.random.default_rng(1).normal(size=(M, M)) / M ** 2    # not symmetric
    f, b = fast_model(var, ell, Z, mu, S, dL0, dL1, dL2)
    of = psi_forward(var, ell, Z, mu, S)
    ob = psi_backward(dL0, dL1, dL2, var, ell, Z, mu, S)
    for a, c in zip(f, of):
        assert relerr(a, c) < 1e-12
    for a, c in zip(b, ob):
        assert relerr(a, c) < 1e-11
