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

from ._lib import Handle, IMPL_AUTO, IMPL_FAST, IMPL_REFERENCE, host_digest

_IMPL = {"auto": IMPL_AUTO, "fast": IMPL_FAST, "reference": IMPL_REFERENCE}


def _f64(a) -> np.ndarray:
    """paramz Param / ObsAr are ndarray subclasses: take a plain C-order float64 view."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------
# Cache key.  GPy wraps both psicomp methods in ``Cache_this(limit=10)`` and the three forward
# accessors / three gradient accessors of one ELBO evaluation rely on it.  GPy's cacher is valid
# because it subscribes to paramz change notifications; without paramz the only valid key is the
# CONTENT: the layer rewrites X.mean / X.variance in place every evaluation (layers.py:528-550)
# - with the same rows in a new order in the permuted-minibatch tests
# (testing/minibatch_tests.py:281-296) - and paramz mutates Z / lengthscale / variance in place.
# So the key is an ORDER-SENSITIVE 128-bit digest of the raw bytes, computed by librgp_psi's
# rgp_host_digest (xxh64-style lanes over 8 MiB slices on several threads, slice digests folded in
# order: memory-bandwidth bound, because at the headline shape one backward key covers 21 GB).  A
# wrapping sum / XOR of the words, which round 1 used, is permutation invariant and returned stale
# row-ordered results.
# ---------------------------------------------------------------------------------------
def _digest(a: np.ndarray) -> bytes:
    return host_digest(np.ascontiguousarray(a))


def _fingerprint(*arrays) -> tuple:
    """Order-sensitive content fingerprint: (shape, 128-bit digest of the bytes) per array."""
    return tuple((a.shape, _digest(a)) for a in arrays)


class PSICOMP_RBF_B200(object):
    """GPy ``PSICOMP_RBF`` interface backed by librgp_psi (sm_100a).

    Parameters
    ----------
    device : CUDA ordinal the handle binds to (one process per GPU).
    impl   : 'auto' | 'fast' | 'reference' - kernel family (see include/rgp_psi.h).
    cache  : keep the last forward / backward result, as GPy's ``Cache_this`` does; the
             three forward accessors and the three gradient accessors of one ELBO
             evaluation then cost one device evaluation each.  Keyed on an order-sensitive
             digest of the input bytes (see ``_fingerprint``).
    cache_copy_bytes : results up to this many bytes are handed out as fresh copies (a caller
             may then mutate what it got).  Above it the cached arrays THEMSELVES are returned
             on a hit - exactly what GPy's cacher does (it returns the stored tuple) - so an
             N x M Psi1 of 17 GB is neither duplicated in host memory nor copied per accessor.
    """

    def __init__(self, device: int = 0, impl: str = "auto", cache: bool = True,
                 cache_copy_bytes: int = 64 << 20):
        if impl not in _IMPL:
            raise ValueError("impl must be one of %s" % sorted(_IMPL))
        self.device = int(device)
        self.impl = impl
        self.cache = bool(cache)
        self.cache_copy_bytes = int(cache_copy_bytes)
        self._handle = Handle(self.device)
        self._handle.set_option("impl", _IMPL[impl])
        self._fwd_key = self._fwd_val = None
        self._bwd_key = self._bwd_val = None

    # pickling / deepcopy: drop device state and cached arrays (ownership)
    def __getstate__(self):
        return {"device": self.device, "impl": self.impl, "cache": self.cache,
                "cache_copy_bytes": self.cache_copy_bytes}

    def __setstate__(self, state):
        self.__init__(**state)

    def __deepcopy__(self, memo):
        return PSICOMP_RBF_B200(self.device, self.impl, self.cache, self.cache_copy_bytes)

    # cache entries: (values, copy?) - small results are stored and handed out as copies, large
    # ones by reference (GPy's Cache_this semantics)
    def _store(self, vals):
        nbytes = sum(a.nbytes for a in vals if isinstance(a, np.ndarray))
        if nbytes <= self.cache_copy_bytes:
            return tuple(a.copy() if isinstance(a, np.ndarray) else a for a in vals), True
        return tuple(vals), False

    @staticmethod
    def _hand_out(entry):
        vals, copy = entry
        return tuple(a.copy() if (copy and isinstance(a, np.ndarray)) else a for a in vals)

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
                return self._hand_out(self._fwd_val)
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
            self._fwd_key, self._fwd_val = key, self._store((psi0, psi1, psi2))
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
        dL0c = 0.0
        if np.size(dL0) == 1:                     # a constant dL_dpsi0 (vardtc.py:175: -D beta/2 for every row)
            dL0c, dL0 = float(np.asarray(dL0, dtype=np.float64).reshape(-1)[0]), None
        else:
            dL0 = _f64(dL0).reshape(-1)
            if dL0.shape != (N,):
                raise ValueError("dL_dpsi0 has %d entries for N=%d" % (dL0.size, N))
        dL1 = _f64(dL1)
        dL2 = _f64(dL2)
        if dL1.shape != (N, M) or dL2.shape != (M, M):
            raise ValueError("dL_dpsi1 %s / dL_dpsi2 %s do not match N=%d M=%d"
                             % (dL1.shape, dL2.shape, N, M))
        if N == 0:
            return 0.0, np.zeros(Q if ard else 1), np.zeros((M, Q)), np.empty((0, Q)), np.empty((0, Q))
        key = None
        if self.cache:
            key = (var, dL0c) + _fingerprint(ell, Z, mu, S, dL1, dL2) + (() if dL0 is None else _fingerprint(dL0))
            if key == self._bwd_key:
                return self._hand_out(self._bwd_val)
        dmu = np.empty((N, Q))
        dS = np.empty((N, Q))
        dZ = np.empty((M, Q))
        dell = np.empty(Q)
        dvar = np.empty(1)
        self._handle.backward_host(N, M, Q, _ptr(mu), _ptr(S), _ptr(Z), _ptr(ell), var,
                                   _ptr(dL0), dL0c, _ptr(dL1), _ptr(dL2),
                                   _ptr(dmu), _ptr(dS), _ptr(dZ), _ptr(dell), _ptr(dvar))
        dl = dell if ard else np.array([dell.sum()])
        out = (float(dvar[0]), dl, dZ, dmu, dS)
        if self.cache:
            self._bwd_key, self._bwd_val = key, self._store(out)
        return out
