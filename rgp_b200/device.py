"""Device-resident entry points: torch CUDA tensors in, torch CUDA tensors out.

The numpy plugin path (``psicomp.py``) pays host<->device copies of Psi1 and dL_dpsi1
(17 GB each at the headline shape).  Throughput callers keep q(X), Psi1, dL_dpsi1 and
the row gradients resident in HBM and call these functions instead; they enqueue on
torch's current stream and do not synchronise.  torch is plumbing only (allocation,
streams); all arithmetic is in librgp_psi.so.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._lib import Handle


def _check(t: torch.Tensor, name: str, shape=None) -> torch.Tensor:
    if not t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous float64 CUDA tensor")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t


class DevicePsi:
    """One handle per (process, GPU); methods mirror the C ABI."""

    def __init__(self, device: Optional[int] = None, impl: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("rgp_b200.device needs a CUDA device; there is no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.handle = Handle(self.device)
        self.handle.set_option("impl", impl)

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def forward(self, mu, S, Z, ell, variance: float, want_psi0=False, want_psi1=True,
                psi1_out: Optional[torch.Tensor] = None, psi2_out: Optional[torch.Tensor] = None
                ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], torch.Tensor]:
        N, Q = mu.shape
        M = Z.shape[0]
        _check(mu, "mu"); _check(S, "S", (N, Q)); _check(Z, "Z", (M, Q)); _check(ell, "ell", (Q,))
        dev = mu.device
        psi0 = torch.empty(N, dtype=torch.float64, device=dev) if want_psi0 else None
        psi1 = None
        if want_psi1:
            psi1 = psi1_out if psi1_out is not None else torch.empty((N, M), dtype=torch.float64, device=dev)
            _check(psi1, "psi1_out", (N, M))
        psi2 = psi2_out if psi2_out is not None else torch.empty((M, M), dtype=torch.float64, device=dev)
        _check(psi2, "psi2_out", (M, M))
        self.handle.forward_dev(self._stream(), N, M, Q, mu.data_ptr(), S.data_ptr(), Z.data_ptr(),
                                ell.data_ptr(), variance,
                                psi0.data_ptr() if psi0 is not None else None,
                                psi1.data_ptr() if psi1 is not None else None, psi2.data_ptr())
        return psi0, psi1, psi2

    def backward(self, mu, S, Z, ell, variance: float, dL_dpsi0, dL_dpsi1, dL_dpsi2,
                 dmu_out=None, dS_out=None):
        """dL_dpsi0 may be a python float (constant over rows, vardtc.py:175) or an N-vector."""
        N, Q = mu.shape
        M = Z.shape[0]
        _check(mu, "mu"); _check(S, "S", (N, Q)); _check(Z, "Z", (M, Q)); _check(ell, "ell", (Q,))
        _check(dL_dpsi2, "dL_dpsi2", (M, M))
        if dL_dpsi1 is not None:
            _check(dL_dpsi1, "dL_dpsi1", (N, M))
        dev = mu.device
        if isinstance(dL_dpsi0, torch.Tensor):
            _check(dL_dpsi0, "dL_dpsi0", (N,))
            p0, c0 = dL_dpsi0.data_ptr(), 0.0
        else:
            p0, c0 = None, float(dL_dpsi0)
        dmu = dmu_out if dmu_out is not None else torch.empty((N, Q), dtype=torch.float64, device=dev)
        dS = dS_out if dS_out is not None else torch.empty((N, Q), dtype=torch.float64, device=dev)
        _check(dmu, "dmu_out", (N, Q)); _check(dS, "dS_out", (N, Q))
        dZ = torch.empty((M, Q), dtype=torch.float64, device=dev)
        dell = torch.empty(Q, dtype=torch.float64, device=dev)
        dvar = torch.empty(1, dtype=torch.float64, device=dev)
        self.handle.backward_dev(self._stream(), N, M, Q, mu.data_ptr(), S.data_ptr(), Z.data_ptr(),
                                 ell.data_ptr(), variance, p0, c0,
                                 dL_dpsi1.data_ptr() if dL_dpsi1 is not None else None,
                                 dL_dpsi2.data_ptr(), dmu.data_ptr(), dS.data_ptr(), dZ.data_ptr(),
                                 dell.data_ptr(), dvar.data_ptr())
        return dvar, dell, dZ, dmu, dS

    def fused(self, mu, S, Z, ell, variance: float, dL_dpsi0, dL_dpsi1, dL_dpsi2, want_psi1: bool = True,
              psi1_out=None, dmu_out=None, dS_out=None):
        """Statistics and gradients from one pass (``rgp_psi_fused_dev``): for upstream gradients
        that do not depend on the statistics of this evaluation (the SVI bound).  Returns
        ((psi1 | None, psi2), (dvar, dell, dZ, dmu, dS))."""
        N, Q = mu.shape
        M = Z.shape[0]
        _check(mu, "mu"); _check(S, "S", (N, Q)); _check(Z, "Z", (M, Q)); _check(ell, "ell", (Q,))
        _check(dL_dpsi2, "dL_dpsi2", (M, M))
        if dL_dpsi1 is not None:
            _check(dL_dpsi1, "dL_dpsi1", (N, M))
        dev = mu.device
        if isinstance(dL_dpsi0, torch.Tensor):
            _check(dL_dpsi0, "dL_dpsi0", (N,))
            p0, c0 = dL_dpsi0.data_ptr(), 0.0
        else:
            p0, c0 = None, float(dL_dpsi0)
        f64 = dict(dtype=torch.float64, device=dev)
        psi1 = (psi1_out if psi1_out is not None else torch.empty((N, M), **f64)) if want_psi1 else None
        if psi1 is not None:
            _check(psi1, "psi1_out", (N, M))
        psi2 = torch.empty((M, M), **f64)
        dmu = dmu_out if dmu_out is not None else torch.empty((N, Q), **f64)
        dS = dS_out if dS_out is not None else torch.empty((N, Q), **f64)
        _check(dmu, "dmu_out", (N, Q)); _check(dS, "dS_out", (N, Q))
        dZ, dell, dvar = torch.empty((M, Q), **f64), torch.empty(Q, **f64), torch.empty(1, **f64)
        self.handle.fused_dev(self._stream(), N, M, Q, mu.data_ptr(), S.data_ptr(), Z.data_ptr(), ell.data_ptr(),
                              variance, p0, c0, dL_dpsi1.data_ptr() if dL_dpsi1 is not None else None,
                              dL_dpsi2.data_ptr(), psi1.data_ptr() if psi1 is not None else None, psi2.data_ptr(),
                              dmu.data_ptr(), dS.data_ptr(), dZ.data_ptr(), dell.data_ptr(), dvar.data_ptr())
        return (psi1, psi2), (dvar, dell, dZ, dmu, dS)
