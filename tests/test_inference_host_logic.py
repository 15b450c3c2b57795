"""CPU test of the bound algebra in rgp_b200/inference.py (torch, fp64): the psi calls are
served by an oracle-backed stand-in (test code may use the oracle; the product wires in the
CUDA DevicePsi), so what is checked is the restated VarDTC / SVI algebra, its jitchol /
backsub helpers and the K(Z,Z) gradient pieces against oracle/bound_oracle.py."""
import numpy as np
import pytest
import torch

from oracle import bound_oracle as bo
from oracle.psi_oracle import psi_backward, psi_forward
from rgp_b200.inference import DeviceBound, backsub_both_sides, jitchol
from synth import make_inputs, relerr


class OraclePsi:
    """Same call signature as rgp_b200.device.DevicePsi, CPU tensors, oracle arithmetic."""

    def forward(self, mu, S, Z, ell, variance, **kw):
        p0, p1, p2 = psi_forward(variance, ell.numpy(), Z.numpy(), mu.numpy(), S.numpy())
        return torch.from_numpy(p0), torch.from_numpy(p1), torch.from_numpy(p2)

    def backward(self, mu, S, Z, ell, variance, dL0, dL1, dL2, **kw):
        N = mu.shape[0]
        d0 = np.full(N, dL0) if not isinstance(dL0, torch.Tensor) else dL0.numpy()
        out = psi_backward(d0, dL1.numpy(), dL2.numpy(), variance, ell.numpy(), Z.numpy(), mu.numpy(), S.numpy())
        return (torch.tensor([out[0]]),) + tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in out[1:])


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_linalg_helpers_match_gpy_semantics():
    rng = np.random.default_rng(0)
    A = rng.normal(size=(6, 6)); A = A @ A.T + 6 * np.eye(6)
    X = rng.normal(size=(6, 6)); X = X + X.T
    L = jitchol(_t(A))
    Ln = bo.jitchol(A)
    assert relerr(L.numpy(), Ln) < 1e-14
    for side in ("left", "right"):
        assert relerr(backsub_both_sides(L, _t(X), side).numpy(), bo.backsub_both_sides(Ln, X, side)) < 1e-12
    sing = np.ones((4, 4))                                     # needs jitter
    assert torch.isfinite(jitchol(_t(sing))).all()


@pytest.mark.parametrize("svi", [False, True])
def test_bound_algebra_matches_cpu_restatement(svi):
    N, M, Q, D = 60, 7, 3, 2
    var, ell, Z, mu, S = make_inputs(N, M, Q, seed=9)
    rng = np.random.default_rng(1)
    Y = rng.normal(size=(N, D))
    db = DeviceBound(psi=OraclePsi())
    if svi:
        W = rng.normal(size=(M, M)) * 0.1
        mode = dict(qU_mean=rng.normal(size=(M, D)), qU_var=W @ W.T + 0.5 * np.eye(M), qU_ratio=0.4)
        Lc, gc = db.svi(var, _t(ell), _t(Z), _t(mu), _t(S), _t(Y), 0.1, _t(mode["qU_mean"]), _t(mode["qU_var"]), 0.4)
    else:
        mode = None
        Lc, gc = db.vardtc(var, _t(ell), _t(Z), _t(mu), _t(S), _t(Y), 0.1)
    Lo, go = bo.layer_bound_and_grads(var, ell, Z, mu, S, Y, 0.1, psi_forward, psi_backward, svi=mode)
    assert abs(float(Lc) - Lo) < 1e-11 * abs(Lo)
    for k in ("variance", "lengthscale", "Z", "mu", "S"):
        assert relerr(gc[k].numpy(), go[k]) < 1e-10, k
