"""numpy models of the device-side elementary functions (rgp_b200/csrc/common.cuh exp_neg / exp_tab,
mlp_kernels.cuh tanh_fast): the same constants and the same operation order, evaluated in float64 without
FMA contraction, against numpy's exp / tanh.  They pin the accuracy claims written next to those functions
(the CUDA versions differ from these models only by FMA rounding, a few 1e-17 relative per step); the GPU
parity tests then check the kernels themselves against the oracle."""
import re
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG2E, LN2, MAGIC = 1.4426950408889634074, 6.93147180559945286227e-01, 6755399441055744.0


def _coeffs(fname, func):
    """The polynomial constants of `func` as written in the CUDA header (so the model cannot drift)."""
    src = open(os.path.join(ROOT, "rgp_b200", "csrc", fname)).read()
    body = src[src.index(func):]
    body = body[:body.index("\n}")]
    return [float(x) for x in re.findall(r"(?:p = |p = fma\(p, [ry], )(\d\.\d+e[+-]\d+)", body)]


def exp_neg_model(x):
    c = _coeffs("common.cuh", "RGP_DEVINL double exp_neg(double x)")
    assert len(c) == 10
    kd = x * LOG2E + MAGIC
    kf = kd - MAGIC
    r = kf * (-LN2) + x
    p = np.full_like(x, c[0])
    for ci in c[1:]:
        p = p * r + ci
    p = p * r + 1.0
    return np.where(x < -708.0, 0.0, np.ldexp(p, kf.astype(np.int64)))


def exp_tab_model(x):
    INV, STEP = 369.32993046757463, 2.7076061740622863e-03
    kd = x * INV + MAGIC
    nf = kd - MAGIC
    n = nf.astype(np.int64)
    r = nf * (-STEP) + x
    p = r * 4.16666666666666666667e-02 + 1.66666666666666666667e-01
    p = p * r + 0.5
    p = p * r + 1.0
    p = p * r + 1.0
    tab = np.exp2(np.arange(256) / 256.0)
    return np.where(x < -708.0, 0.0, np.ldexp(tab[n & 255] * p, n >> 8))


def tanh_fast_model(x):
    c = _coeffs("mlp_kernels.cuh", "RGP_DEVINL double tanh_fast(double x)")
    assert len(c) == 10
    y = -2.0 * np.abs(x)
    ys = np.maximum(y, -0.34)
    p = np.full_like(x, c[0])
    for ci in c[1:]:
        p = p * ys + ci
    em = np.where(y > -0.34, p * ys, exp_neg_model(np.maximum(y, -745.0)) - 1.0)
    return np.copysign(-em / (em + 2.0), x)


def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def test_exp_neg_accuracy():
    x = np.concatenate([np.linspace(-60.0, 2.0, 600001), -np.logspace(-12, 2.8, 20001)])
    # without FMA the range reduction k * ln2 rounds once more: relative error <= ~|x| * 1.5e-16 (the exponent's
    # absolute error); the device code fuses it.  Large |x| means a vanishing term of a sum over rows.
    assert np.all(_rel(exp_neg_model(x), np.exp(x)) <= 1e-15 + 1.5e-16 * np.abs(x))
    assert exp_neg_model(np.array([-709.0, -1e300]))[0] == 0.0 and exp_neg_model(np.array([-1e300]))[0] == 0.0


def test_exp_tab_accuracy():
    x = np.concatenate([np.linspace(-50.0, 0.5, 600001), -np.logspace(-12, 2.8, 20001)])
    assert np.all(_rel(exp_tab_model(x), np.exp(x)) <= 1.2e-15 + 1.5e-16 * np.abs(x))
    assert _rel(exp_tab_model(x[np.abs(x) < 30]), np.exp(x[np.abs(x) < 30])).max() < 5e-15
    assert exp_tab_model(np.array([-1e300]))[0] == 0.0              # padded inducing points (H = -1e300) give 0


def test_tanh_fast_accuracy():
    x = np.concatenate([np.linspace(-20, 20, 400001), np.logspace(-12, 1, 100001), -np.logspace(-12, 1, 1001)])
    got, ref = tanh_fast_model(x), np.tanh(x)
    assert _rel(got, ref).max() < 1e-14 and np.abs(got - ref).max() < 5e-16
    assert np.all(np.abs(got) <= 1.0) and tanh_fast_model(np.array([0.0]))[0] == 0.0
