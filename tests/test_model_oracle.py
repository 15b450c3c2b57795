"""Pins oracle/model_oracle.py (one whole objective evaluation of the deep autoregressive
model) the way the reference pins itself: ``model.checkgrad`` = finite differences of the
objective against the analytic gradient of every parameter block
(testing/minibatch_tests.py, testing/rnn_tests.py use checkgrad the same way)."""
import copy

import numpy as np
import pytest

from oracle.model_oracle import deep_autoreg_oracle
from synth import make_deep_model


def _objective(m):
    return deep_autoreg_oracle(m["wins"], m["Ys"], m["latents"], m["params"], Us=m["Us"],
                               U_win=m["U_win"], svi=m["svi"])


def _fd(m, getter, idx, h):
    vals = []
    for sgn in (+1, -1):
        mm = copy.deepcopy(m)
        arr = getter(mm)
        arr[idx] += sgn * h
        vals.append(_objective(mm)[0])
    return (vals[0] - vals[1]) / (2 * h)


@pytest.mark.parametrize("svi,control", [(False, True), (False, False), (True, True)])
def test_checkgrad_every_block(svi, control):
    m = make_deep_model(svi=svi, control=control)
    logL, res, lat_grads, ctl_grads = _objective(m)
    assert np.isfinite(logL)
    rng = np.random.default_rng(0)
    L = len(m["wins"])
    for i in range(L):
        p = m["params"][i]
        for key in ("Z", "lengthscale") + (("qU_mean", "qU_W") if svi else ()):
            a = p[key]
            for _ in range(3):
                idx = tuple(rng.integers(0, n) for n in a.shape)
                fd = _fd(m, lambda mm, i=i, key=key: mm["params"][i][key], idx, 1e-6)
                an = res[i][key][idx]
                assert abs(fd - an) <= 2e-6 * max(1.0, abs(an)), (i, key, idx, fd, an)
        for key in ("variance", "noise_variance") + (("qU_a",) if svi else ()):
            vals = []
            for sgn in (+1, -1):
                mm = copy.deepcopy(m)
                mm["params"][i][key] += sgn * 1e-6
                vals.append(_objective(mm)[0])
            fd = (vals[0] - vals[1]) / 2e-6
            an = float(res[i][key])
            assert abs(fd - an) <= 2e-6 * max(1.0, abs(an)), (i, key, fd, an)
    for lvl in range(L - 1):
        for s in range(len(m["Ys"])):
            for k in (0, 1):
                a = m["latents"][lvl][s][k]
                for _ in range(4):
                    idx = tuple(rng.integers(0, n) for n in a.shape)
                    fd = _fd(m, lambda mm, lvl=lvl, s=s, k=k: mm["latents"][lvl][s][k], idx, 1e-6)
                    an = lat_grads[lvl][s][k][idx]
                    assert abs(fd - an) <= 2e-6 * max(1.0, abs(an)), (lvl, s, k, idx, fd, an)


def test_control_gradients_are_scattered_from_zero():
    m = make_deep_model()
    _, res, _, ctl_grads = _objective(m)
    assert ctl_grads is not None and len(ctl_grads) == len(m["Ys"])
    # d bound / d control mean by finite differences (the reference computes it although the
    # controls are not parameters, layers.py:568-570)
    idx = (3, 0)
    fd = _fd(m, lambda mm: mm["Us"][1][0], idx, 1e-6)
    assert abs(fd - ctl_grads[1][0][idx]) <= 2e-6 * max(1.0, abs(fd))


def test_config1_fixture_is_what_the_oracle_computes():
    """tests/golden/actuator_config1.npz (real Actuator data, reference benchmark recipe) must
    stay what the oracle computes - a change of the oracle shows up here."""
    from synth import load_actuator_config1, relerr
    m, g = load_actuator_config1()
    logL, res, lat_grads, ctl_grads = _objective(m)
    # same code, same LAPACK: reproduces to rounding; conditioning (sens_*) bounds anything else
    assert abs(logL - float(g["logL"])) <= 1e-12 * abs(logL)
    for i in range(2):
        for k in ("variance", "lengthscale", "Z", "noise_variance"):
            assert relerr(res[i][k], g["g%d_%s" % (i, k)]) <= max(1e-10, float(g["sens_g%d_%s" % (i, k)])), (i, k)
    assert relerr(lat_grads[0][0][0], g["g_lat_mean"]) <= max(1e-10, float(g["sens_g_lat_mean"]))
    assert relerr(lat_grads[0][0][1], g["g_lat_var"]) <= max(1e-10, float(g["sens_g_lat_var"]))
