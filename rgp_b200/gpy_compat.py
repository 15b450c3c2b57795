"""Duck-typed stand-ins for the GPy objects either side of the psicomp boundary.

GPy (and paramz) cannot be installed in this image, so the parity tests cannot build a
real ``GPy.kern.RBF``.  These two classes reproduce exactly the slice of GPy's surface
that RGP touches on the hot path, so tests read like the reference's own calls:

  ``RBF``              GPy.kern.RBF(input_dim, variance, lengthscale, ARD, inv_l)
                       .psi0/.psi1/.psi2(Z, X)                      vardtc.py:59-61
                       .update_gradients_expectations(...)          layers.py:98-102
                       .gradients_Z_expectations(...)               layers.py:127-132
                       .gradients_qX_expectations(...)              layers.py:574-580
                       ``inv_l=True`` chain rule (every reference config builds its
                       kernels with it, e.g. autoreg/benchmark/methods.py:69):
                       l = 1/sqrt(inv_l + 1e-200), d/dinv_l = d/dl * (-l^3/2)
  ``NormalPosterior``  GPy.core.parameterization.variational.NormalPosterior
                       (.mean, .variance; built at autoreg/layers.py:484-489)

They hold plain numpy arrays, forward every expectation to ``self.psicomp`` and contain
no psi arithmetic of their own.  With real GPy present none of this is needed:
``kern.psicomp = PSICOMP_RBF_B200()`` is the whole integration.
"""
from __future__ import annotations

import numpy as np

from .psicomp import PSICOMP_RBF_B200


class NormalPosterior(object):
    def __init__(self, means, variances, name="latent space"):
        self.mean = np.array(means, dtype=np.float64, order="C")
        self.variance = np.array(variances, dtype=np.float64, order="C")
        assert self.mean.shape == self.variance.shape
        self.name = name

    @property
    def shape(self):
        return self.mean.shape

    @property
    def num_data(self):
        return self.mean.shape[0]

    @property
    def input_dim(self):
        return self.mean.shape[1]


class RBF(object):
    """Parameter holder with GPy's RBF expectation accessors; gradients land in
    ``.variance_gradient`` / ``.lengthscale_gradient`` (``inv_l_gradient`` when
    ``inv_l=True``) the way GPy writes ``.variance.gradient`` etc."""

    def __init__(self, input_dim, variance=1.0, lengthscale=None, ARD=False, inv_l=False,
                 psicomp=None, useGPU=True, device=0):
        self.input_dim = int(input_dim)
        self.ARD = bool(ARD)
        self.use_invLengthscale = bool(inv_l)
        n = self.input_dim if ARD else 1
        ls = np.ones(n) if lengthscale is None else np.asarray(lengthscale, dtype=np.float64).reshape(-1)
        assert ls.size == n, "lengthscale size must be %d" % n
        self.variance = np.array([float(variance)])
        if inv_l:
            self.inv_l = 1.0 / ls ** 2
        self._lengthscale = ls.copy()
        self.psicomp = psicomp if psicomp is not None else PSICOMP_RBF_B200(device=device)
        self.variance_gradient = np.zeros(1)
        self.lengthscale_gradient = np.zeros(n)
        self.inv_l_gradient = np.zeros(n)

    @property
    def lengthscale(self):
        if self.use_invLengthscale:
            return 1.0 / np.sqrt(self.inv_l + 1e-200)
        return self._lengthscale

    @lengthscale.setter
    def lengthscale(self, v):
        self._lengthscale = np.asarray(v, dtype=np.float64).reshape(-1).copy()

    # ---- expectations (forward)
    def psi0(self, Z, variational_posterior):
        return self.psicomp.psicomputations(self, Z, variational_posterior)[0]

    def psi1(self, Z, variational_posterior):
        return self.psicomp.psicomputations(self, Z, variational_posterior)[1]

    def psi2(self, Z, variational_posterior):
        return self.psicomp.psicomputations(self, Z, variational_posterior)[2]

    # ---- expectation gradients (backward)
    def update_gradients_expectations(self, dL_dpsi0, dL_dpsi1, dL_dpsi2, Z, variational_posterior):
        dvar, dl = self.psicomp.psiDerivativecomputations(
            self, dL_dpsi0, dL_dpsi1, dL_dpsi2, Z, variational_posterior)[:2]
        self.variance_gradient = np.array([dvar])
        if self.use_invLengthscale:
            self.inv_l_gradient = dl * (self.lengthscale ** 3 / -2.0)
        else:
            self.lengthscale_gradient = np.asarray(dl)

    def gradients_Z_expectations(self, dL_dpsi0, dL_dpsi1, dL_dpsi2, Z, variational_posterior):
        return self.psicomp.psiDerivativecomputations(
            self, dL_dpsi0, dL_dpsi1, dL_dpsi2, Z, variational_posterior)[2]

    def gradients_qX_expectations(self, dL_dpsi0, dL_dpsi1, dL_dpsi2, Z, variational_posterior):
        return self.psicomp.psiDerivativecomputations(
            self, dL_dpsi0, dL_dpsi1, dL_dpsi2, Z, variational_posterior)[3:]
