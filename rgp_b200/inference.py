"""Device-resident sparse variational bounds around the psi path.

The reference evaluates the bound on the host between the two psi phases
(``VarDTC.inference`` autoreg/inference/vardtc.py:88-208, ``SVI_VarDTC.inference``
autoreg/inference/svi_vardtc.py:70-195, ``comp_KL_qU`` :197-215), which forces Psi1 and
dL_dpsi1 (N x M, 17 GB each at the headline shape) across PCIe twice per evaluation.
Here the same algebra runs on the GPU: the M x M Cholesky / triangular solves go to
cuSOLVER / cuBLAS through ``torch.linalg`` (off the hot path, per the north star), the
N x M products (psi1^T Y, Y v) are cuBLAS GEMMs, and Psi1 / dL_dpsi1 never leave HBM.
The psi statistics themselves come from librgp_psi (``DevicePsi``).

GPy helpers restated on the device: jitchol (growing jitter), dtrtrs, backsub_both_sides
('left' = L^-T X L^-1, 'right' = L^-1 X L^-T), tdot.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from .device import DevicePsi

LOG_2_PI = math.log(2.0 * math.pi)
CONST_JITTER = 1e-6            # vardtc.py:28, svi_vardtc.py:28


def jitchol(A: torch.Tensor, maxtries: int = 5, pending: Optional[list] = None) -> torch.Tensor:
    """GPy ``jitchol``: plain Cholesky, and only if that fails growing diagonal jitter.
    With ``pending`` (a list) the success flag of the plain factorisation is appended to it
    instead of being read back - no host synchronisation; the caller checks all flags once
    (``DeviceBound.verify``) and re-runs in the careful mode if any is non-zero."""
    L, info = torch.linalg.cholesky_ex(A)
    if pending is not None:
        pending.append(info)
        return L
    if int(info) == 0:
        return L
    diag = torch.diagonal(A)
    if bool((diag <= 0).any()):
        raise RuntimeError("not pd: non-positive diagonal elements")
    jitter = float(diag.mean()) * 1e-6
    eye = torch.eye(A.shape[0], dtype=A.dtype, device=A.device)
    for _ in range(maxtries):
        L, info = torch.linalg.cholesky_ex(A + eye * jitter)
        if int(info) == 0:
            return L
        jitter *= 10.0
    raise RuntimeError("not positive definite, even with jitter.")


def dtrtrs(L: torch.Tensor, B: torch.Tensor, trans: int = 0) -> torch.Tensor:
    if trans:
        return torch.linalg.solve_triangular(L.mT, B, upper=True)
    return torch.linalg.solve_triangular(L, B, upper=False)


def backsub_both_sides(L: torch.Tensor, X: torch.Tensor, transpose: str = "left") -> torch.Tensor:
    t = 1 if transpose == "left" else 0
    tmp = dtrtrs(L, X, trans=t)
    return dtrtrs(L, tmp.mT, trans=t).mT


def rsolve_T(L: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """B L^-T  ( = dtrtrs(L, B.T).T of the reference) for tall B [N, M] without transposing it."""
    return torch.linalg.solve_triangular(L.mT, B, upper=True, left=False)


def rsolve(L: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """B L^-1  ( = dtrtrs(L, B.T, trans=1).T )."""
    return torch.linalg.solve_triangular(L, B, upper=False, left=False)


def tdot(A: torch.Tensor) -> torch.Tensor:
    return A @ A.mT


def rbf_K(variance: float, ell: torch.Tensor, Z: torch.Tensor) -> torch.Tensor:
    Zs = Z / ell
    r2 = (Zs[:, None, :] - Zs[None, :, :]).square().sum(-1)
    return variance * torch.exp(-0.5 * r2)


def rbf_K_grads(dL_dK: torch.Tensor, variance: float, ell: torch.Tensor, Z: torch.Tensor):
    """GPy ``update_gradients_full(dL_dKmm, Z)`` + ``gradients_X(dL_dKmm, Z)``
    (autoreg/layers.py:105, :134) on the device (M x M x Q, independent of N)."""
    W = dL_dK * rbf_K(variance, ell, Z)
    diff = Z[:, None, :] - Z[None, :, :]
    dvar = W.sum() / variance
    dl = torch.einsum("ab,abq->q", W, diff.square()) / ell ** 3
    dZ = -torch.einsum("ab,abq->aq", W + W.mT, diff) / ell ** 2
    return dvar, dl, dZ


class DeviceBound:
    """One sparse-GP layer with uncertain inputs, evaluated entirely on one GPU:
    psi forward -> bound algebra -> psi backward (+ the K(Z,Z) terms the layer adds,
    layers.py:98-134, 574-580)."""

    def __init__(self, device: Optional[int] = None, psi: Optional[DevicePsi] = None, group=None,
                 sharded: bool = False):
        """``sharded=True`` (or a ``group``): the rows handed to vardtc / svi are THIS rank's
        rows of a data set split across the ranks of ``group`` (default: the world).  The
        sums over rows - Psi2, Psi1^T Y, YRY, N, the uncertain-output M x M terms, and after the
        backward dvariance / dlengthscale / dZ - are all-reduced (exchanges C1 and
        C2; the reference's dead MPI path did the first, svi_vardtc.py:65-67); the M x M
        algebra then runs redundantly on every rank.  Returned bound and parameter gradients
        are global and identical on every rank; Psi1, dL_dpsi1 and the row gradients stay local."""
        self.psi = psi if psi is not None else DevicePsi(device)
        self.group, self.sharded = group, bool(sharded or group is not None)
        self._pending: Optional[list] = None     # deferred Cholesky success flags (see jitchol)

    # The M x M factorisations almost never need jitter.  A caller that evaluates several
    # layers (rgp_b200.layer) opens a deferred section so that no factorisation reads its
    # status back to the host; verify() reads all of them with one synchronisation.
    def defer_checks(self) -> None:
        self._pending = []

    def verify(self) -> bool:
        """True if every factorisation since defer_checks() succeeded; ends the section."""
        pending, self._pending = self._pending, None
        if not pending:
            return True
        return int(torch.stack([p.reshape(()) for p in pending]).abs().sum()) == 0

    def _chol(self, A: torch.Tensor) -> torch.Tensor:
        return jitchol(A, pending=self._pending)

    def allsum(self, tensors):
        """SUM over the ranks of the group of a list of tensors (one packed all-reduce);
        identity when not sharded."""
        if not self.sharded:
            return list(tensors)
        from .sharded import allreduce_packed
        return allreduce_packed(list(tensors), self.group)

    def _gather_stats(self, N, psi2, psi1Y, YRY):
        """Exchange C1: the row sums every rank needs before the M x M algebra."""
        if not self.sharded:
            return N, psi2, psi1Y, YRY
        n = torch.tensor([float(N)], dtype=psi2.dtype, device=psi2.device)
        n, psi2, psi1Y, YRY = self.allsum([n, psi2, psi1Y, YRY.reshape(1)])
        return int(round(float(n))), psi2, psi1Y, YRY.reshape(())

    # ------------------------------------------------------------------ VarDTC
    def vardtc(self, variance: float, ell, Z, mu, S, Y, noise_variance: float, Y_var=None
               ) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        """autoreg/inference/vardtc.py:88-208, uncertain inputs.  ``Y_var`` [N, D] switches on
        the uncertain-output branch (hidden layers: :70-77, :136-140, :179, :201-206), whose
        N x M triangular solves run as cuBLAS TRSMs on the rows resident in HBM; the two
        solves of the reference against the same factor (:137, :203) differ by a row scaling
        and are done once."""
        N, D = Y.shape
        M = Z.shape[0]
        beta = 1.0 / max(float(noise_variance), 1e-6)                         # :102
        _, psi1, psi2 = self.psi.forward(mu, S, Z, ell, variance)
        psi1Y = (Y.mT @ psi1) * beta                                          # :79  D x M
        YRY = Y.square().sum() * beta                                         # :81-82
        if Y_var is not None:
            YRY = YRY + Y_var.sum() * beta                                    # :77
        N, psi2, psi1Y, YRY = self._gather_stats(N, psi2, psi1Y, YRY)         # N is global from here on
        psi0b = variance * N * beta                                           # :68
        psi2b = psi2 * beta
        eye = torch.eye(M, dtype=Z.dtype, device=Z.device)
        Kmm = rbf_K(variance, ell, Z) + eye * CONST_JITTER                    # :110-114
        Lm = self._chol(Kmm)
        A = backsub_both_sides(Lm, psi2b, "right")                            # :120
        LL = self._chol(eye + A)                                                 # :124-125
        LmLL = Lm @ LL
        logdet_L = 2.0 * torch.log(torch.diagonal(LL)).sum()
        b = dtrtrs(LmLL, psi1Y.mT).mT                                         # :131
        bbt = b.square().sum()
        v = dtrtrs(LmLL, b.mT, trans=1).mT                                    # :133
        C = tdot(b.mT)
        if Y_var is not None:
            Shalf = Y_var.sum(dim=1).sqrt()                                   # :74
            psi1LmiLLi = rsolve_T(LmLL, psi1) * beta                          # :203  N x M
            psi1SLLinv = Shalf[:, None] * psi1LmiLLi                          # :137
            C_S, bbt_S = self.allsum([psi1SLLinv.mT @ psi1SLLinv, psi1SLLinv.square().sum().reshape(1)])
            bbt = bbt + bbt_S.reshape(())
            C = C + C_S                                                       # :139
            psi1SP = rsolve(LmLL, psi1SLLinv)                                 # :140
        tmp = -backsub_both_sides(LL, C + D * eye)                            # :141
        dL_dpsi2R = backsub_both_sides(Lm, tmp + D * eye) / 2.0               # :142
        logL_R = -N * math.log(beta)
        logL = -(D * (N * LOG_2_PI + logL_R + psi0b - torch.trace(A)) + YRY - bbt) / 2.0 \
            - D * logdet_L / 2.0                                              # :150
        dL_dKmm = dL_dpsi2R - D * backsub_both_sides(Lm, A) / 2.0             # :156
        dL_dthetaL = (YRY * beta + beta * D * psi0b - N * D * beta) / 2.0 \
            - beta * (dL_dpsi2R * psi2b).sum() - beta * torch.trace(C)        # :169
        dL_dpsi0 = -D * beta / 2.0                                            # :175 (constant over rows)
        dL_dpsi1 = (Y @ v) * beta                                             # :181
        extra = {"dL_dthetaL": dL_dthetaL, "woodbury_vector": v.mT}
        if Y_var is not None:
            dL_dpsi1 = dL_dpsi1 + (Shalf[:, None] * psi1SP) * beta            # :179
            extra["dL_dYmean"] = psi1LmiLLi @ b.mT - Y * beta                 # :205
            extra["dL_dYvar"] = psi1LmiLLi.square().sum(dim=1) / 2.0 - beta / 2.0   # :206  [N]
            del psi1SP, psi1SLLinv, psi1LmiLLi
        dL_dpsi2 = dL_dpsi2R * beta                                           # :184
        return logL, self._finish(variance, ell, Z, mu, S, dL_dpsi0, dL_dpsi1, dL_dpsi2, dL_dKmm, extra)

    # -------------------------------------------------------------------- SVI
    def svi(self, variance: float, ell, Z, mu, S, Y, noise_variance: float, qU_mean, qU_var,
            qU_ratio: float = 1.0, Y_var=None, fused: Optional[bool] = None
            ) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        """autoreg/inference/svi_vardtc.py:70-215 plus the KL scaling of layers.py:75-79.
        ``Y_var`` [N, D]: uncertain outputs (:56-59, :190-193).

        In this bound the upstream gradients dL_dpsi1 (:167) and dL_dpsi2 (:169) depend on q(U) and
        K(Z,Z) only, not on the statistics, so they are formed FIRST and the statistics and their
        gradients come out of ONE pass over the rows (``DevicePsi.fused``, ``rgp_psi_fused_dev``):
        the separate Psi2 forward pass - a quarter of a two-phase evaluation - is not run.
        ``fused=False`` forces the two-phase order of the reference."""
        N, D = Y.shape
        M = Z.shape[0]
        beta = 1.0 / float(noise_variance)                                    # :80
        if fused is None:
            fused = hasattr(self.psi, "fused")
        eye = torch.eye(M, dtype=Z.dtype, device=Z.device)
        Lm = self._chol(rbf_K(variance, ell, Z) + eye * CONST_JITTER)            # :88-93
        Ls = self._chol(qU_var)                                                  # :96
        LinvLs = dtrtrs(Lm, Ls)
        Linvmu = dtrtrs(Lm, qU_mean)
        B = tdot(LinvLs) * D + tdot(Linvmu)                                   # :112
        KuuInvmu = dtrtrs(Lm, Linvmu, trans=1)
        dL_dpsi0 = -D * beta / 2.0                                            # :162
        dL_dpsi1 = (Y @ KuuInvmu.mT) * beta                                   # :167
        dL_dpsi2 = beta * backsub_both_sides(Lm, D * eye - B, "left") / 2.0   # :169
        if fused:
            (psi1, psi2), (dvar, dl, dZ, dmu, dS) = self.psi.fused(mu, S, Z, ell, variance, dL_dpsi0, dL_dpsi1,
                                                                   dL_dpsi2)
            pgrads = (dvar, dl, dZ, dmu, dS)
        else:
            _, psi1, psi2 = self.psi.forward(mu, S, Z, ell, variance)
            pgrads = None
        psi1Y = (Y.mT @ psi1) * beta
        YRY = Y.square().sum() * beta
        if Y_var is not None:
            YRY = YRY + Y_var.sum() * beta                                    # :59
        N, psi2, psi1Y, YRY = self._gather_stats(N, psi2, psi1Y, YRY)         # :65-67 (allReduceArrays)
        psi0b = variance * N * beta
        psi2b = psi2 * beta
        psi1YLinvT = dtrtrs(Lm, psi1Y.mT).mT                                  # :99
        A = backsub_both_sides(Lm, psi2b, "right")                            # :108
        logL_R = -N * math.log(beta)
        core = -D * psi0b / 2.0 - YRY / 2.0 - (B * A).sum() / 2.0 + torch.trace(A) * D / 2.0 \
            + (Linvmu * psi1YLinvT.mT).sum()
        logL = -N * D * LOG_2_PI / 2.0 - D * logL_R / 2.0 + core              # :122-123
        tmp1 = backsub_both_sides(Lm, B @ A, "left")                          # :129
        tmp2 = Linvmu @ psi1YLinvT
        tmp3 = backsub_both_sides(Lm, -D * A - tmp2 - tmp2.mT, "left") / 2.0
        dL_dKmm = (tmp1 + tmp1.mT) / 2.0 + tmp3                               # :133
        dL_dthetaL = -D * N * beta / 2.0 - core * beta                        # :139
        t1 = backsub_both_sides(Lm, -A, "left")                               # :145
        dL_dqU_mean = t1 @ qU_mean + dtrtrs(Lm, psi1YLinvT.mT, trans=1)       # :146
        dL_dqU_var = D / 2.0 * t1                                             # :147
        # KL(q(U) || p(U)), svi_vardtc.py:197-215, scaled by qU_ratio (layers.py:76-79)
        Linv = dtrtrs(Lm, eye)
        KuuInv = Linv.mT @ Linv
        LuInv = dtrtrs(Ls, eye)
        KL = D * M / -2.0 - torch.log(torch.diagonal(Ls)).sum() * D + torch.log(torch.diagonal(Lm)).sum() * D \
            + LinvLs.square().sum() / 2.0 * D + Linvmu.square().sum() / 2.0
        dKL_dqU_mean = dtrtrs(Lm, Linvmu, trans=1)
        dKL_dqU_var = (tdot(LuInv.mT) / -2.0 + KuuInv / 2.0) * D
        dKL_dKuu = KuuInv * D / 2.0 - KuuInv @ (tdot(qU_mean) + qU_var * D) @ KuuInv / 2.0
        logL = logL - KL * qU_ratio
        extra = {"dL_dthetaL": dL_dthetaL,
                 "dL_dqU_mean": dL_dqU_mean - dKL_dqU_mean * qU_ratio,
                 "dL_dqU_var": dL_dqU_var - dKL_dqU_var * qU_ratio}
        if Y_var is not None:
            # :192  dtrtrs(Lm, psi1^T)^T . dtrtrs(Lm, mu) = psi1 . Kuu^-1 mu, without the N x M solve
            extra["dL_dYmean"] = (psi1 @ KuuInvmu) * beta - Y * beta
            extra["dL_dYvar"] = torch.full_like(Y, beta / -2.0)               # :193  [N, D]
        return logL, self._finish(variance, ell, Z, mu, S, dL_dpsi0, dL_dpsi1, dL_dpsi2,
                                  dL_dKmm - dKL_dKuu * qU_ratio, extra, pgrads)

    # ------------------------------------------------------------ shared tail
    def _finish(self, variance, ell, Z, mu, S, dL_dpsi0, dL_dpsi1, dL_dpsi2, dL_dKmm, extra, pgrads=None):
        dvar, dl, dZ, dmu, dS = pgrads if pgrads is not None else \
            self.psi.backward(mu, S, Z, ell, variance, dL_dpsi0, dL_dpsi1, dL_dpsi2)
        dvar, dl, dZ = self.allsum([dvar, dl, dZ])                            # exchange C2
        kvar, kl, kZ = rbf_K_grads(dL_dKmm, variance, ell, Z)
        out = {"variance": dvar.reshape(()) + kvar, "lengthscale": dl + kl, "Z": dZ + kZ, "mu": dmu, "S": dS,
               "dL_dKmm": dL_dKmm, "dL_dpsi1": dL_dpsi1, "dL_dpsi2": dL_dpsi2}
        out.update(extra)
        return out
