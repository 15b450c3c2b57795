"""CPU oracle for the sparse variational bounds either side of the psi path.
TEST INFRASTRUCTURE ONLY (see ``oracle/psi_oracle.py`` for who may import this).

Restates, line by line, with scipy replacing GPy's linalg helpers:
  * ``vardtc_inference``      autoreg/inference/vardtc.py:54-208   (collapsed bound)
  * ``svi_vardtc_inference``  autoreg/inference/svi_vardtc.py:43-195 (uncollapsed bound)
  * ``svi_kl_qu``             autoreg/inference/svi_vardtc.py:197-215
  * latent entropy / prior    autoreg/variational.py:4-24 as applied in
                              autoreg/layers.py:596-615
and the RBF ``K(Z,Z)`` pieces the layer adds on the M x M side
(autoreg/layers.py:105-107, :134) so a test can finite-difference the *whole* bound.

GPy helpers restated (GPy.util.linalg, not in /root/reference):
  jitchol(A)                     Cholesky, retrying with growing diagonal jitter
  dtrtrs(L, B, trans)            triangular solve, L lower
  backsub_both_sides(L, X, 'left')  = L^-T X L^-1 ;  'right' = L^-1 X L^-T
  tdot(A) = A A^T ; dpotri(L) = (L L^T)^-1 ; dtrtri(L) = L^-1

PARITY PINNED against the reference's own code: tests/golden/ref_bounds.npz and
ref_variational.npz hold outputs of autoreg/inference/vardtc.py, svi_vardtc.py and
variational.py EXECUTED in the build container (tests/golden/make_reference_golden.py; only the
GPy.util.linalg LAPACK wrappers and container classes they import are stubbed, GPy itself is
absent), and tests/test_reference_golden.py holds this file to them at <= 1e-12.  Also pinned
relationally the way the reference's tests pin themselves (``model.checkgrad`` -> finite
differences, row additivity, testing/minibatch_tests.py:98-100, :288-296).
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import cholesky, solve_triangular

LOG_2_PI = np.log(2.0 * np.pi)
CONST_JITTER = 1e-6          # vardtc.py:28, svi_vardtc.py:28


# ------------------------------------------------------------------ linalg restated
def jitchol(A, maxtries=5):
    A = np.ascontiguousarray(A, dtype=np.float64)
    try:
        return cholesky(A, lower=True)
    except np.linalg.LinAlgError:
        pass
    diagA = np.diag(A)
    if np.any(diagA <= 0.0):
        raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    for _ in range(maxtries):
        try:
            return cholesky(A + np.eye(A.shape[0]) * jitter, lower=True)
        except np.linalg.LinAlgError:
            jitter *= 10.0
    raise np.linalg.LinAlgError("not positive definite, even with jitter.")


def dtrtrs(L, B, trans=0):
    return solve_triangular(L, B, lower=True, trans=trans)


def backsub_both_sides(L, X, transpose="left"):
    if transpose == "left":                      # L^-T X L^-1
        tmp = dtrtrs(L, X, trans=1)
        return dtrtrs(L, tmp.T, trans=1).T
    tmp = dtrtrs(L, X, trans=0)                  # L^-1 X L^-T
    return dtrtrs(L, tmp.T, trans=0).T


def tdot(A):
    return A @ A.T


# ----------------------------------------------------------------- RBF on the M x M side
def rbf_K(variance, lengthscale, Z):
    Zs = Z / lengthscale
    r2 = np.square(Zs[:, None, :] - Zs[None, :, :]).sum(axis=2)
    return variance * np.exp(-0.5 * r2)


def rbf_K_grads(dL_dK, variance, lengthscale, Z):
    """(dvar, dl[Q], dZ[M,Q]) of sum(dL_dK * K(Z,Z)); GPy ``update_gradients_full`` and
    ``gradients_X(dL_dKmm, Z)`` (autoreg/layers.py:105, :134)."""
    K = rbf_K(variance, lengthscale, Z)
    W = dL_dK * K
    diff = Z[:, None, :] - Z[None, :, :]
    dvar = W.sum() / variance
    dl = np.einsum("ab,abq->q", W, np.square(diff)) / lengthscale ** 3
    Ws = W + W.T
    dZ = -np.einsum("ab,abq->aq", Ws, diff) / lengthscale ** 2
    return dvar, dl, dZ


# ------------------------------------------------------------------------- VarDTC
def vardtc_inference(psi0, psi1, psi2, Kmm, Y, noise_variance, Y_var=None):
    """autoreg/inference/vardtc.py:88-208 for uncertain inputs.

    psi0[N], psi1[N,M], psi2[M,M] are the *raw* kernel expectations (beta is applied
    here as vardtc.py:59-61,68 do).  ``Y`` is N x D; ``Y_var`` (N x D) switches on the
    uncertain-output branch (vardtc.py:70-77, :136-140, :201-206).
    Returns (logL, grads) with grads keys dL_dpsi0, dL_dpsi1, dL_dpsi2, dL_dKmm,
    dL_dthetaL [, dL_dYmean, dL_dYvar], woodbury_inv, woodbury_vector.
    """
    Y = np.asarray(Y, dtype=np.float64)
    N, D = Y.shape
    M = Kmm.shape[0]
    beta = 1.0 / max(float(noise_variance), 1e-6)                    # :102
    psi1b = psi1 * beta
    psi2b = psi2 * beta
    psi0b = psi0.sum() * beta                                        # :68
    psi1Y = Y.T @ psi1b                                              # :73/:79
    if Y_var is not None:
        Shalf = np.sqrt(Y_var.sum(axis=1))                           # :74
        psi1S = Shalf[:, None] * psi1b
        YRY = (np.square(Y).sum() + Y_var.sum()) * beta              # :77
    else:
        Shalf = psi1S = None
        YRY = np.square(Y).sum() * beta                              # :81-82

    Kmm = Kmm.copy()
    Kmm[np.diag_indices(M)] += CONST_JITTER                          # :114
    Lm = jitchol(Kmm)                                                # :116
    A = backsub_both_sides(Lm, psi2b, "right")                       # :120
    Lambda = np.eye(M) + A                                           # :124
    LL = jitchol(Lambda)
    LmLL = Lm @ LL
    logdet_L = 2.0 * np.log(np.diag(LL)).sum()                       # :130
    b = dtrtrs(LmLL, psi1Y.T).T                                      # :131  D x M
    bbt = np.square(b).sum()
    v = dtrtrs(LmLL, b.T, trans=1).T                                 # :133  D x M
    C = tdot(b.T)                                                    # :134  M x M
    if psi1S is not None:
        psi1SLLinv = dtrtrs(LmLL, psi1S.T).T                         # :137
        bbt += np.square(psi1SLLinv).sum()
        C = C + tdot(psi1SLLinv.T)
        psi1SP = dtrtrs(LmLL, psi1SLLinv.T, trans=1).T               # :140
    tmp = -backsub_both_sides(LL, C + D * np.eye(M))                 # :141
    dL_dpsi2R = backsub_both_sides(Lm, tmp + D * np.eye(M)) / 2.0    # :142

    logL_R = -N * np.log(beta)                                       # :149
    logL = -(D * (N * LOG_2_PI + logL_R + psi0b - np.trace(A)) + YRY - bbt) / 2.0 \
        - D * logdet_L / 2.0                                         # :150
    dL_dKmm = dL_dpsi2R - D * backsub_both_sides(Lm, A) / 2.0        # :156
    wd_inv = backsub_both_sides(
        Lm, np.eye(M) - backsub_both_sides(LL, np.eye(M), "left"), "left")  # :162
    dL_dthetaL = (YRY * beta + beta * D * psi0b - N * D * beta) / 2.0 \
        - beta * (dL_dpsi2R * psi2b).sum() - beta * np.trace(C)      # :169
    grads = {
        "dL_dpsi0": -D * (beta * np.ones(N)) / 2.0,                  # :175
        "dL_dpsi2": beta * dL_dpsi2R,                                # :184
        "dL_dKmm": dL_dKmm,
        "dL_dthetaL": dL_dthetaL,
        "woodbury_inv": wd_inv,
        "woodbury_vector": v.T,
    }
    if Y_var is not None:
        grads["dL_dpsi1"] = beta * (Y @ v + Shalf[:, None] * psi1SP)  # :179
        psi1LmiLLi = dtrtrs(LmLL, psi1b.T).T                         # :203
        grads["dL_dYmean"] = -Y * beta + psi1LmiLLi @ b.T            # :205
        grads["dL_dYvar"] = beta / -2.0 + np.square(psi1LmiLLi).sum(axis=1) / 2.0  # :206
    else:
        grads["dL_dpsi1"] = beta * (Y @ v)                           # :181
    return float(logL), grads


# ----------------------------------------------------------------------- SVI VarDTC
def svi_vardtc_inference(psi0, psi1, psi2, Kuu, Y, noise_variance, qU_mean, qU_var,
                         Y_var=None):
    """autoreg/inference/svi_vardtc.py:70-195 for uncertain inputs.  Returns
    (logL, grads, mid); ``mid`` feeds ``svi_kl_qu`` like ``self.mid`` (:101-105)."""
    Y = np.asarray(Y, dtype=np.float64)
    N, D = Y.shape
    M = Kuu.shape[0]
    beta = 1.0 / float(noise_variance)                               # :80
    psi0b = psi0.sum() * beta                                        # :48
    psi1b = psi1 * beta
    psi2b = psi2 * beta
    psi1Y = Y.T @ psi1b                                              # :58/:61
    if Y_var is not None:
        YRY = (np.square(Y).sum() + Y_var.sum()) * beta              # :59
    else:
        YRY = np.square(Y).sum() * beta                              # :63

    Kuu = Kuu.copy()
    Kuu[np.diag_indices(M)] += CONST_JITTER                          # :92
    Lm = jitchol(Kuu)
    mu, S = qU_mean, qU_var
    Ls = jitchol(S)                                                  # :96
    LinvLs = dtrtrs(Lm, Ls)
    Linvmu = dtrtrs(Lm, mu)
    psi1YLinvT = dtrtrs(Lm, psi1Y.T).T                               # :99
    mid = {"qU_L": Ls, "LinvLu": LinvLs, "L": Lm, "Linvmu": Linvmu}
    A = backsub_both_sides(Lm, psi2b, "right")                       # :108
    B = tdot(LinvLs) * D + tdot(Linvmu)                              # :112

    logL_R = -N * np.log(beta)
    core = -D * psi0b / 2.0 - YRY / 2.0 - (B * A).sum() / 2.0 \
        + np.trace(A) * D / 2.0 + (Linvmu * psi1YLinvT.T).sum()
    logL = -N * D * LOG_2_PI / 2.0 - D * logL_R / 2.0 + core         # :122-123

    tmp1 = backsub_both_sides(Lm, B @ A, "left")                     # :129
    tmp2 = Linvmu @ psi1YLinvT
    tmp3 = backsub_both_sides(Lm, -D * A - tmp2 - tmp2.T, "left") / 2.0
    dL_dKmm = (tmp1 + tmp1.T) / 2.0 + tmp3                           # :133
    dL_dthetaL = -D * N * beta / 2.0 - core * beta                   # :139
    t1 = backsub_both_sides(Lm, -A, "left")                          # :145
    KuuInvmu = dtrtrs(Lm, Linvmu, trans=1)                           # :153
    grads = {
        "dL_dpsi0": -D * (beta * np.ones(N)) / 2.0,                  # :162
        "dL_dpsi1": (Y @ KuuInvmu.T) * beta,                         # :165/:167
        "dL_dpsi2": beta * backsub_both_sides(Lm, D * np.eye(M) - B, "left") / 2.0,  # :169
        "dL_dKmm": dL_dKmm,
        "dL_dthetaL": dL_dthetaL,
        "dL_dqU_mean": t1 @ mu + dtrtrs(Lm, psi1YLinvT.T, trans=1),  # :146
        "dL_dqU_var": D / 2.0 * t1,                                  # :147
        "woodbury_inv": backsub_both_sides(Lm, np.eye(M) - tdot(LinvLs), "left"),
        "woodbury_vector": KuuInvmu,
    }
    if Y_var is not None:
        grads["dL_dYmean"] = -Y * beta + dtrtrs(Lm, psi1b.T).T @ dtrtrs(Lm, mu)   # :192
        grads["dL_dYvar"] = beta / -2.0 * np.ones((N, D))            # :193
    return float(logL), grads, mid


def svi_kl_qu(qU_mean, qU_var, mid):
    """autoreg/inference/svi_vardtc.py:197-215: KL(q(U)||p(U)) and its gradients."""
    M, D = qU_mean.shape
    Lu, L, Linvmu, LinvLu = mid["qU_L"], mid["L"], mid["Linvmu"], mid["LinvLu"]
    Linv = dtrtrs(L, np.eye(M))
    KuuInv = Linv.T @ Linv                                           # dpotri
    LuInv = dtrtrs(Lu, np.eye(M))                                    # dtrtri
    KL = D * M / -2.0 - np.log(np.diag(Lu)).sum() * D + np.log(np.diag(L)).sum() * D \
        + np.square(LinvLu).sum() / 2.0 * D + np.square(Linvmu).sum() / 2.0
    dKL_dqU_mean = dtrtrs(L, Linvmu, trans=1)
    dKL_dqU_var = (tdot(LuInv.T) / -2.0 + KuuInv / 2.0) * D
    dKL_dKuu = KuuInv * D / 2.0 - KuuInv @ (tdot(qU_mean) + qU_var * D) @ KuuInv / 2.0
    return float(KL), dKL_dqU_mean, dKL_dqU_var, dKL_dKuu


# -------------------------------------------------------- latent prior / entropy pieces
def normal_entropy_term(var):
    """-NormalEntropy.comp_value (autoreg/variational.py:7-9 as used at layers.py:611)
    = +entropy; returns (value added to the bound, d/dvar)."""
    return (1.0 + LOG_2_PI + np.log(var)).sum() / 2.0, 1.0 / (2.0 * var)


def normal_prior_term(mean, var):
    """-NormalPrior.comp_value (autoreg/variational.py:16-19 as used at layers.py:608);
    returns (value added to the bound, d/dmean, d/dvar)."""
    val = -(0.5 * (np.square(mean).sum() + (var - np.log(var)).sum()) - 0.5 * mean.size)
    return val, -mean, -(1.0 - 1.0 / var) * 0.5


# ------------------------------------------------------- one layer, end to end (tests)
def layer_bound_and_grads(variance, lengthscale, Z, mu, S, Y, noise_variance,
                          psi_fwd, psi_bwd, svi=None):
    """One sparse-GP layer with uncertain inputs: ELBO and its gradients with respect to
    (variance, lengthscale, Z, mu, S), the way autoreg/layers.py:66-134,574-580 assembles
    them.  ``psi_fwd(variance, l, Z, mu, S) -> (psi0, psi1, psi2)`` and
    ``psi_bwd(dL0, dL1, dL2, variance, l, Z, mu, S) -> (dvar, dl, dZ, dmu, dS)`` are
    injected, so the same harness runs the oracle or the CUDA path.
    ``svi`` = dict(qU_mean, qU_var, qU_ratio) selects the SVI bound.
    """
    lengthscale = np.asarray(lengthscale, dtype=np.float64)
    psi0, psi1, psi2 = psi_fwd(variance, lengthscale, Z, mu, S)
    Kmm = rbf_K(variance, lengthscale, Z)
    if svi is None:
        logL, g = vardtc_inference(psi0, psi1, psi2, Kmm, Y, noise_variance)
    else:
        logL, g, mid = svi_vardtc_inference(psi0, psi1, psi2, Kmm, Y, noise_variance,
                                            svi["qU_mean"], svi["qU_var"])
        KL, _, _, dKL_dKuu = svi_kl_qu(svi["qU_mean"], svi["qU_var"], mid)
        logL += -KL * svi.get("qU_ratio", 1.0)                       # layers.py:76
        g["dL_dKmm"] = g["dL_dKmm"] - dKL_dKuu * svi.get("qU_ratio", 1.0)  # layers.py:79
    dvar, dl, dZ, dmu, dS = psi_bwd(g["dL_dpsi0"], g["dL_dpsi1"], g["dL_dpsi2"],
                                    variance, lengthscale, Z, mu, S)
    kvar, kl, kZ = rbf_K_grads(g["dL_dKmm"], variance, lengthscale, Z)
    return logL, {"variance": dvar + kvar, "lengthscale": np.asarray(dl) + kl,
                  "Z": dZ + kZ, "mu": dmu, "S": dS, "inner": g}
