"""CPU test of the host logic in rgp_b200/layer.py (layer wiring, update order, uncertain-
output branches of rgp_b200/inference.py) with oracle-backed stand-ins for the CUDA pieces."""
import pytest

from model_standins import OracleLag, OraclePsi, OraclePsiFused, compare_with_oracle, stack_model
from rgp_b200.inference import DeviceBound
from rgp_b200.layer import DeviceDeepAutoreg
from synth import make_deep_model, relerr


@pytest.mark.parametrize("svi,control,wins,nDims", [
    (False, True, (0, 2, 3), (2, 1, 2)),
    (False, False, (0, 2, 3), (2, 1, 2)),
    (True, True, (0, 2, 3), (2, 1, 2)),
    (False, True, (0, 3), (1, 2)),
    (True, False, (0, 1, 1, 2), (3, 2, 1, 1)),
])
def test_deep_model_matches_oracle(svi, control, wins, nDims):
    m = make_deep_model(svi=svi, control=control, wins=wins, nDims=nDims)
    Y, latents, controls, params = stack_model(m)
    model = DeviceDeepAutoreg(m["wins"], nDims, [y.shape[0] for y in m["Ys"]], U_win=m["U_win"],
                              ctl_dim=1 if control else 0, svi=svi, bound=DeviceBound(psi=OraclePsi()),
                              lag_factory=OracleLag)
    out = model.evaluate(params, Y, latents, controls)
    compare_with_oracle(m, out, relerr, tol=1e-10)


def test_svi_fused_order_equals_two_phase_order():
    """DeviceBound.svi forms the upstream gradients before the statistics when the psi object offers a
    one-pass entry point; the result must not depend on that order."""
    m = make_deep_model(svi=True, control=True)
    Y, latents, controls, params = stack_model(m)
    outs = []
    for psi in (OraclePsi(), OraclePsiFused()):
        model = DeviceDeepAutoreg(m["wins"], (2, 1, 2), [y.shape[0] for y in m["Ys"]], U_win=m["U_win"], ctl_dim=1,
                                  svi=True, bound=DeviceBound(psi=psi), lag_factory=OracleLag)
        outs.append(model.evaluate(params, Y, latents, controls))
        compare_with_oracle(m, outs[-1], relerr, tol=1e-10)
    assert abs(float(outs[0][0]) - float(outs[1][0])) <= 1e-13 * abs(float(outs[0][0]))


def test_rejects_windowed_observed_layer():
    with pytest.raises(ValueError):
        DeviceDeepAutoreg((1, 2), (1, 1), [5], bound=DeviceBound(psi=OraclePsi()), lag_factory=OracleLag)


def test_deferred_cholesky_checks_fall_back_to_the_jitter_path():
    """qU_var = W W^T with rank-1 W: the plain factorisation fails, GPy's jitchol adds jitter.
    The deferred (no read-back) pass must notice and re-run in the careful mode."""
    import numpy as np
    import torch
    m = make_deep_model(svi=True, control=False)
    for p in m["params"]:
        p["qU_W"] = np.ones_like(p["qU_W"])
        p["qU_a"] = 0.0
    Y, latents, controls, params = stack_model(m)
    bound = DeviceBound(psi=OraclePsi())
    model = DeviceDeepAutoreg(m["wins"], (2, 1, 2), [y.shape[0] for y in m["Ys"]], U_win=m["U_win"], svi=True,
                              bound=bound, lag_factory=OracleLag)
    calls = []
    inner = model._evaluate
    model._evaluate = lambda *a: (calls.append(bound._pending is None), inner(*a))[1]
    logL, res, lat_grads, _ = model.evaluate(params, Y, latents, controls)
    assert calls == [False, True]                      # deferred pass, then careful pass
    assert np.isfinite(float(logL))
    assert all(torch.isfinite(g[0]).all() and torch.isfinite(g[1]).all() for g in lat_grads)
