"""The oracle against outputs of the REFERENCE'S OWN CODE (tests/golden/ref_*.npz, produced by
tests/golden/make_reference_golden.py executing autoreg/inference/vardtc.py, svi_vardtc.py,
variational.py, util.py and rnn_encoder.py from /root/reference in the build container, with
only the GPy / paramz helper imports stubbed).  This is what pins oracle/bound_oracle.py,
oracle/lag_oracle.py and the latent terms; the psi statistics (GPy's own arithmetic) stay
pinned by quadrature in test_oracle.py."""
import os

import numpy as np
import pytest

from oracle import bound_oracle as bo
from oracle.lag_oracle import get_conv_1D
from oracle.psi_oracle import psi_forward
from synth import relerr

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(G, "ref_bounds.npz"))


def _psi(ref):
    return psi_forward(float(ref["variance"]), ref["ell"], ref["Z"], ref["mu"], ref["S"])


@pytest.mark.parametrize("tag", ["c", "u"])
def test_vardtc_restatement_matches_reference_code(ref, tag):
    psi0, psi1, psi2 = _psi(ref)
    Kmm = bo.rbf_K(float(ref["variance"]), ref["ell"], ref["Z"])
    logL, g = bo.vardtc_inference(psi0, psi1, psi2, Kmm, ref["Y"], float(ref["noise"]),
                                  Y_var=ref["Y_var"] if tag == "u" else None)
    assert abs(logL - float(ref["vardtc_%s_logL" % tag])) <= 1e-13 * abs(logL)
    keys = ["dL_dpsi0", "dL_dpsi1", "dL_dpsi2", "dL_dKmm", "dL_dthetaL", "woodbury_vector", "woodbury_inv"]
    if tag == "u":
        keys += ["dL_dYmean", "dL_dYvar"]
    for k in keys:
        assert relerr(g[k], ref["vardtc_%s_%s" % (tag, k)]) <= 1e-12, k


@pytest.mark.parametrize("tag", ["c", "u"])
def test_svi_restatement_matches_reference_code(ref, tag):
    psi0, psi1, psi2 = _psi(ref)
    Kuu = bo.rbf_K(float(ref["variance"]), ref["ell"], ref["Z"])
    logL, g, mid = bo.svi_vardtc_inference(psi0, psi1, psi2, Kuu, ref["Y"], float(ref["noise"]), ref["qU_mean"],
                                           ref["qU_var"], Y_var=ref["Y_var"] if tag == "u" else None)
    assert abs(logL - float(ref["svi_%s_logL" % tag])) <= 1e-13 * abs(logL)
    keys = ["dL_dpsi0", "dL_dpsi1", "dL_dpsi2", "dL_dKmm", "dL_dthetaL", "dL_dqU_mean", "dL_dqU_var",
            "woodbury_vector"]
    if tag == "u":
        keys += ["dL_dYmean", "dL_dYvar"]
    for k in keys:
        assert relerr(g[k], ref["svi_%s_%s" % (tag, k)]) <= 1e-12, k
    KL, dm, dv, dK = bo.svi_kl_qu(ref["qU_mean"], ref["qU_var"], mid)
    assert abs(KL - float(ref["svi_%s_KL" % tag])) <= 1e-13 * abs(KL)
    assert relerr(dm, ref["svi_%s_dKL_dqU_mean" % tag]) <= 1e-12
    assert relerr(dv, ref["svi_%s_dKL_dqU_var" % tag]) <= 1e-12
    assert relerr(dK, ref["svi_%s_dKL_dKuu" % tag]) <= 1e-12


def test_latent_terms_match_reference_code():
    r = np.load(os.path.join(G, "ref_variational.npz"))
    val, dv = bo.normal_entropy_term(r["var"])               # the layer adds -comp_value (layers.py:611)
    assert abs(val + float(r["entropy_value"])) <= 1e-14 * abs(val)
    np.testing.assert_array_equal(dv, r["entropy_dvar"])
    val, dm, dv = bo.normal_prior_term(r["mean"], r["var"])  # -comp_value (layers.py:608)
    assert abs(val + float(r["prior_value"])) <= 1e-13 * abs(val)
    np.testing.assert_array_equal(dm, r["prior_dmean"])
    np.testing.assert_array_equal(dv, r["prior_dvar"])


def test_lag_window_matches_reference_code():
    r = np.load(os.path.join(G, "ref_conv.npz"))
    for w in (1, 2, 4):
        ours = get_conv_1D(r["arr"], w).reshape(r["arr"].shape[0] - w + 1, -1)
        np.testing.assert_array_equal(ours, r["win%d" % w])
