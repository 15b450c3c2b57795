"""The drop-in ``psicomp`` plugin: GPy's psi-statistics interface on B200.

GPy's RBF kernel delegates its expectations under a Gaussian q(X) to a ``psicomp``
object (GPy ``kern/src/psi_comp``: ``PSICOMP_RBF`` on CPU, ``PSICOMP_RBF_GPU`` with
pycuda).  RGP reaches it through the kernel only:

  forward   kern.psi0/psi1/psi2(Z, X)              autoreg/inference/vardtc.py:59-61,
                                                   autoreg/inference/svi_vardtc.py:48-50
  backward  kern.update_gradients_expectations     autoreg/layers.py:98-102
            kern.gradients_Z_expectations          autoreg/layers.py:127-132
            kern.gradients_qX_expectations         autoreg/layers.py:574-580

``PSICOMP_RBF_B200`` has the same two methods, argument meaning, return tuples and
error behaviour, and is installed with ``kern.psicomp = PSICOMP_RBF_B200()`` right
after kernel construction.  All arithmetic runs in librgp_psi.so
(hand-written sm_100a CUDA); there is no numpy fallback - without the library or a GPU
the first call raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from ._lib import Handle, IMPL_AUTO, IMPL_FAST, IMPL_REFERENCE

_IMPL = {"auto": IMPL_AUTO, "fast": IMPL_FAST, "reference": IMPL_REFERENCE}


def _f64(a) -> np.ndarray:
    """paramz Param / ObsAr are ndarray subclasses: take a plain C-order float64 view."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _fingerprint(*arrays) -> tuple:
    """Content fingerprint (shape + wrapping sum + xor of the raw 64-bit words).  The
    layer mutates X.mean / X.variance IN PLACE every evaluation (layers.py:537-543) and
    paramz mutates Z / lengthscale / variance in place, so array identity cannot key
    the cache ("Memoisation")."""
    out = []
    for a in arrays:
        w = a.reshape(-1).view(np.uint64)
        out.append((a.shape, int(np.add.reduce(w, dtype=np.uint64)) if w.size else 0,
                    int(np.bitwise_xor.reduce(w)) if w.size else 0))
    return tuple(out)


class PSICOMP_RBF_B200(object):
    """GPy ``PSICOMP_RBF`` interface backed by librgp_psi (sm_100a).

    Parameters
    ----------
    device : CUDA ordinal the handle binds to (one process per GPU).
    impl   : 'auto' | 'fast' | 'reference' - kernel family (see include/rgp_psi.h).
    cache  : keep the last forward / backward result, as GPy's ``Cache_this`` does; the
             three forward accessors and the three gradient accessors of one ELBO
             evaluation then cost one device evaluation each.
    """

    def __init__(self, device: int = 0, impl: str = "auto", cache: bool = True):
        if impl not in _IMPL:
            raise ValueError("impl must be one of %s" % sorted(_IMPL))
        self.device = int(device)
        self.impl = impl
        self.cache = bool(cache)
        self._handle = Handle(self.device)
        self._handle.set_option("impl", _IMPL[impl])
        self._fwd_key = self._fwd_val = None
        self._bwd_key = self._bwd_val = None

    # pickling / deepcopy: drop device state and cached arrays (ownership)
    def __getstate__(self):
        return {"device": self.device, "impl": self.impl, "cache": self.cache}

    def __setstate__(self, state):
        self.__init__(**state)

    def __deepcopy__(self, memo):
        return PSICOMP_RBF_B200(self.device, self.impl, self.cache)

    @property
    def handle(self) -> Handle:
        return self._handle

    # ------------------------------------------------------------------ argument plumbing
    @staticmethod
    def _unpack(args, n_lead):
        """Accept GPy >= 1.0 ``(kern, ...)`` and pre-1.0 ``(variance, lengthscale, ...)``."""
        first = args[0]
        if hasattr(first, "variance") and hasattr(first, "lengthscale"):
            return first.variance, first.lengthscale, args[1:]
        return args[0], args[1], args[2:]

    @staticmethod
    def _prepare(variance, lengthscale, Z, vp):
        if not (hasattr(vp, "mean") and hasattr(vp, "variance")) or hasattr(vp, "binary_prob"):
            # GPy raises for anything but a NormalPosterior (spelling kept from GPy)
            raise ValueError("unknown distriubtion received for psi-statistics")
        mu, S, Z = _f64(vp.mean), _f64(vp.variance), _f64(Z)
        if mu.ndim != 2 or S.shape != mu.shape or Z.ndim != 2 or Z.shape[1] != mu.shape[1]:
            raise ValueError("shape mismatch: mean %s variance %s Z %s" % (mu.shape, S.shape, Z.shape))
        N, Q = mu.shape
        ell_in = _f64(lengthscale).reshape(-1)
        ard = ell_in.size != 1
        if ard and ell_in.size != Q:
            raise ValueError("lengthscale has %d entries for input_dim %d" % (ell_in.size, Q))
        ell = ell_in if ard else np.full(Q, float(ell_in[0]))
        var = float(np.asarray(variance, dtype=np.float64).reshape(-1)[0])
        return var, ell, ard, Z, mu, S

    # ------------------------------------------------------------------------ forward
    def psicomputations(self, *args, **kwargs):
        """``psicomputations(kern, Z, variational_posterior, return_psi2_n=False)``
        -> ``(psi0[N], psi1[N,M], psi2[M,M])``.  Psi0[n] = variance."""
        return_psi2_n = kwargs.pop("return_psi2_n", False)
        variance, lengthscale, rest = self._unpack(args, 1)
        if len(rest) == 3:
            return_psi2_n = rest[2]
        Z, vp = rest[0], rest[1]
        if return_psi2_n:
            # N x M x M is never requested by RGP (4.4 TB at the headline shape)
            raise NotImplementedError("return_psi2_n=True is not supported by PSICOMP_RBF_B200")
        var, ell, _, Z, mu, S = self._prepare(variance, lengthscale, Z, vp)
        key = None
        if self.cache:
            key = (var,) + _fingerprint(ell, Z, mu, S)
            if key == self._fwd_key:
                return tuple(a.copy() for a in self._fwd_val)
        N, Q = mu.shape
        M = Z.shape[0]
        if N == 0:      # no rows: sums over the empty set (what GPy's numpy code returns); nothing to launch
            return np.empty(0), np.empty((0, M)), np.zeros((M, M))
        psi0 = np.empty(N)
        psi1 = np.empty((N, M))
        psi2 = np.empty((M, M))
        self._handle.forward_host(N, M, Q, _ptr(mu), _ptr(S), _ptr(Z), _ptr(ell), var,
                                  _ptr(psi0), _ptr(psi1), _ptr(psi2))
        if self.cache:
            self._fwd_key, self._fwd_val = key, (psi0.copy(), psi1.copy(), psi2.copy())
        return psi0, psi1, psi2

    # ----------------------------------------------------------------------- backward
    def psiDerivativecomputations(self, *args):
        """``psiDerivativecomputations(kern, dL_dpsi0, dL_dpsi1, dL_dpsi2, Z,
        variational_posterior)`` -> ``(dL_dvar, dL_dlengthscale, dL_dZ, dL_dmu, dL_dS)``.
        ``dL_dlengthscale`` is summed to one entry for a non-ARD kernel, as in GPy."""
        first = args[0]
        if hasattr(first, "variance") and hasattr(first, "lengthscale"):
            variance, lengthscale = first.variance, first.lengthscale
            dL0, dL1, dL2, Z, vp = args[1:6]
        else:                                     # pre-1.0: (dL0, dL1, dL2, variance, lengthscale, Z, vp)
            dL0, dL1, dL2, variance, lengthscale, Z, vp = args[:7]
        var, ell, ard, Z, mu, S = self._prepare(variance, lengthscale, Z, vp)
        N, Q = mu.shape
        M = Z.shape[0]
        dL0 = np.ascontiguousarray(np.broadcast_to(_f64(dL0).reshape(-1) if np.ndim(dL0) else
                                                   np.float64(dL0), (N,)))
        dL1 = _f64(dL1)
        dL2 = _f64(dL2)
        if dL1.shape != (N, M) or dL2.shape != (M, M):
            raise ValueError("dL_dpsi1 %s / dL_dpsi2 %s do not match N=%d M=%d"
                             % (dL1.shape, dL2.shape, N, M))
        if N == 0:
            return 0.0, np.zeros(Q if ard else 1), np.zeros((M, Q)), np.empty((0, Q)), np.empty((0, Q))
        key = None
        if self.cache:
            key = (var,) + _fingerprint(ell, Z, mu, S, dL0, dL1, dL2)
            if key == self._bwd_key:
                v = self._bwd_val
                return (v[0],) + tuple(a.copy() for a in v[1:])
        dmu = np.empty((N, Q))
        dS = np.empty((N, Q))
        dZ = np.empty((M, Q))
        dell = np.empty(Q)
        dvar = np.empty(1)
        self._handle.backward_host(N, M, Q, _ptr(mu), _ptr(S), _ptr(Z), _ptr(ell), var,
                                   _ptr(dL0), 0.0, _ptr(dL1), _ptr(dL2),
                                   _ptr(dmu), _ptr(dS), _ptr(dZ), _ptr(dell), _ptr(dvar))
        dl = dell if ard else np.array([dell.sum()])
        out = (float(dvar[0]), dl, dZ, dmu, dS)
        if self.cache:
            self._bwd_key = key
            self._bwd_val = (out[0],) + tuple(a.copy() for a in out[1:])
        return out
