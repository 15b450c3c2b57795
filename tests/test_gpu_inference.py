"""GPU parity of the device-resident bounds against the CPU restatement
of autoreg/inference/vardtc.py and svi_vardtc.py (oracle/bound_oracle.py driven by the
oracle's psi functions).  Tolerance 1e-9 relative on the ELBO and on every gradient."""
import numpy as np
import pytest

from oracle import bound_oracle as bo
from oracle.psi_oracle import psi_backward, psi_forward
from synth import make_inputs, relerr

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("N,M,Q,D,nc", [(502, 100, 20, 1, 10), (490, 50, 20, 2, 10), (408, 200, 40, 59, 20),
                                        (3000, 64, 8, 3, 0)])
def test_vardtc_device_matches_cpu_restatement(N, M, Q, D, nc):
    from rgp_b200.inference import DeviceBound
    var, ell, Z, mu, S = make_inputs(N, M, Q, seed=N + D, n_control=nc)
    Y = np.random.default_rng(D).normal(size=(N, D))
    Lo, go = bo.layer_bound_and_grads(var, ell, Z, mu, S, Y, 0.05, psi_forward, psi_backward)
    Lc, gc = DeviceBound(0).vardtc(var, _t(ell), _t(Z), _t(mu), _t(S), _t(Y), 0.05)
    assert abs(float(Lc) - Lo) <= RTOL * abs(Lo)
    for k in ("variance", "lengthscale", "Z", "mu", "S"):
        assert relerr(gc[k].cpu().numpy(), go[k]) < RTOL, (k, relerr(gc[k].cpu().numpy(), go[k]))
    assert relerr(gc["dL_dKmm"].cpu().numpy(), go["inner"]["dL_dKmm"]) < RTOL
    assert abs(float(gc["dL_dthetaL"]) - go["inner"]["dL_dthetaL"]) <= RTOL * abs(go["inner"]["dL_dthetaL"])


@pytest.mark.parametrize("N,M,Q,D", [(502, 100, 20, 1), (1200, 40, 6, 4)])
def test_svi_device_matches_cpu_restatement(N, M, Q, D):
    from rgp_b200.inference import DeviceBound
    var, ell, Z, mu, S = make_inputs(N, M, Q, seed=N * 3 + D)
    rng = np.random.default_rng(N)
    Y = rng.normal(size=(N, D))
    W = rng.normal(size=(M, M)) * 0.05
    svi = dict(qU_mean=rng.normal(size=(M, D)), qU_var=W @ W.T + 0.5 * np.eye(M), qU_ratio=0.3)
    Lo, go = bo.layer_bound_and_grads(var, ell, Z, mu, S, Y, 0.1, psi_forward, psi_backward, svi=svi)
    Lc, gc = DeviceBound(0).svi(var, _t(ell), _t(Z), _t(mu), _t(S), _t(Y), 0.1, _t(svi["qU_mean"]),
                                _t(svi["qU_var"]), qU_ratio=0.3)
    assert abs(float(Lc) - Lo) <= RTOL * abs(Lo)
    for k in ("variance", "lengthscale", "Z", "mu", "S"):
        assert relerr(gc[k].cpu().numpy(), go[k]) < RTOL, (k, relerr(gc[k].cpu().numpy(), go[k]))
    _, _, mid = bo.svi_vardtc_inference(*psi_forward(var, ell, Z, mu, S), bo.rbf_K(var, ell, Z), Y, 0.1,
                                        svi["qU_mean"], svi["qU_var"])
    _, dKLm, dKLv, _ = bo.svi_kl_qu(svi["qU_mean"], svi["qU_var"], mid)
    assert relerr(gc["dL_dqU_mean"].cpu().numpy(), go["inner"]["dL_dqU_mean"] - 0.3 * dKLm) < RTOL
    assert relerr(gc["dL_dqU_var"].cpu().numpy(), go["inner"]["dL_dqU_var"] - 0.3 * dKLv) < RTOL


def test_svi_minibatch_additivity_on_device():
    """Two half minibatches sum to the full batch (testing/minibatch_tests.py:288-296):
    data-fit terms add over rows; the qU_ratio-scaled KL adds because the ratios do."""
    from rgp_b200.inference import DeviceBound
    N, M, Q, D = 4096, 48, 6, 2
    var, ell, Z, mu, S = make_inputs(N, M, Q, seed=5)
    rng = np.random.default_rng(6)
    Y = rng.normal(size=(N, D))
    qm, qv = rng.normal(size=(M, D)), 0.4 * np.eye(M)
    db = DeviceBound(0)
    args = lambda sl, r: (var, _t(ell), _t(Z), _t(mu[sl]), _t(S[sl]), _t(Y[sl]), 0.2, _t(qm), _t(qv), r)
    Lf, gf = db.svi(*args(slice(0, N), 1.0))
    La, ga = db.svi(*args(slice(0, N // 2), 0.5))
    Lb, gb = db.svi(*args(slice(N // 2, N), 0.5))
    # analytically exact; numerically limited by cond(Kuu) (jitter 1e-6) acting on the 1e-16
    # difference between Psi2 summed in one piece or two (the reference asserts 1e-14 on M = 3)
    np.testing.assert_allclose(float(La) + float(Lb), float(Lf), rtol=1e-9)
    for k in ("variance", "lengthscale", "Z"):
        assert relerr((ga[k] + gb[k]).cpu().numpy(), gf[k].cpu().numpy()) < 1e-8
