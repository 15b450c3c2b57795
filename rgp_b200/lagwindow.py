"""Device-side lag-window build and latent-gradient scatter.

Mirrors what a layer does around the psi path: ``_update_conv`` / ``_init_XY``
(autoreg/layers.py:475-526, autoreg/util.py:6-12) builds the N x Q input rows from the
per-sequence latent (and control) series, and ``update_latent_gradients`` (:552-571) adds the
row gradients back - a Python double loop over sequences x time steps in the reference.
Here both are single kernel launches on the stacked series (librgp_psi, no CPU fallback).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from ._lib import Handle
from .device import _check


class LagWindow:
    """Geometry of one layer: sequence lengths, windows, dims.

    lat_lens[s] = T_s (latent steps of sequence s); rows of sequence s: N_s = T_s - X_win.
    ctl_lens[s] = control steps available (>= N_s + U_win - 1); the last N_s + U_win - 1
    are used, as the reference's ``[-N-U_win+1:]`` slice does.
    """

    def __init__(self, handle: Handle, lat_lens: Sequence[int], X_win: int, X_dim: int,
                 ctl_lens: Optional[Sequence[int]] = None, U_win: int = 0, U_dim: int = 0, device=None):
        if ctl_lens is None:
            U_win, U_dim = 0, 0
        if X_win * X_dim + U_win * U_dim <= 0:
            raise ValueError("empty window")
        self.handle, self.X_win, self.X_dim, self.U_win, self.U_dim = handle, X_win, X_dim, U_win, U_dim
        self.Q = X_win * X_dim + U_win * U_dim
        desc, row, lat, ctl = [], 0, 0, 0
        for s, T in enumerate(lat_lens):
            n = T - X_win
            if n <= 0:
                raise ValueError("sequence %d shorter than the window" % s)
            cl = int(ctl_lens[s]) if ctl_lens is not None else 0
            if ctl_lens is not None and cl < n + U_win - 1:
                raise ValueError("control series %d too short" % s)
            cstart = ctl + (cl - n - U_win + 1) if ctl_lens is not None else 0
            desc.append([row, n, lat, T, cstart, n + U_win - 1 if ctl_lens is not None else 0])
            row, lat, ctl = row + n, lat + T, ctl + cl
        self.N, self.lat_total, self.ctl_total, self.nseq = row, lat, ctl, len(desc)
        self.device = torch.device("cuda", handle.device) if device is None else device
        self.desc = torch.tensor(desc, dtype=torch.int64, device=self.device)

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def gather(self, lat: torch.Tensor, ctl: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """lat [lat_total, X_dim] (means or variances), ctl [ctl_total, U_dim] -> X [N, Q]."""
        if self.X_win:
            _check(lat, "lat", (self.lat_total, self.X_dim))
        if self.U_win:
            if ctl is None:
                raise ValueError("this window has a control part: ctl is required")
            _check(ctl, "ctl", (self.ctl_total, self.U_dim))
        if out is None:
            out = torch.empty((self.N, self.Q), dtype=torch.float64, device=self.device)
        _check(out, "out", (self.N, self.Q))
        self.handle.lag_gather(self._stream(), self.nseq, self.desc.data_ptr(), self.N, self.X_win, self.X_dim,
                               self.U_win, self.U_dim, lat.data_ptr() if self.X_win else None,
                               ctl.data_ptr() if (ctl is not None and self.U_win) else None, out.data_ptr())
        return out

    def scatter_add(self, dX: torch.Tensor, lat_grad: Optional[torch.Tensor] = None,
                    ctl_grad: Optional[torch.Tensor] = None, allocate: bool = True
                    ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """Adds dX [N, Q] onto lat_grad [lat_total, X_dim] / ctl_grad [ctl_total, U_dim], every
        element in the reference's order of ``+=`` (layers.py:552-571).  Omitted targets are
        allocated as zeros (``allocate=True``) or skipped."""
        _check(dX, "dX", (self.N, self.Q))
        if lat_grad is not None:
            _check(lat_grad, "lat_grad", (self.lat_total, self.X_dim))
        if ctl_grad is not None:
            _check(ctl_grad, "ctl_grad", (self.ctl_total, self.U_dim))
        if lat_grad is None and allocate and self.X_win:
            lat_grad = torch.zeros((self.lat_total, self.X_dim), dtype=torch.float64, device=self.device)
        if ctl_grad is None and allocate and self.U_win:
            ctl_grad = torch.zeros((self.ctl_total, self.U_dim), dtype=torch.float64, device=self.device)
        self.handle.lag_scatter(self._stream(), self.nseq, self.desc.data_ptr(), self.N, self.X_win, self.X_dim,
                                self.U_win, self.U_dim, dX.data_ptr(), self.lat_total,
                                lat_grad.data_ptr() if lat_grad is not None else None,
                                self.ctl_total, ctl_grad.data_ptr() if ctl_grad is not None else None)
        return lat_grad, ctl_grad

    def latent_terms(self, lat_mean: torch.Tensor, lat_var: torch.Tensor, dL_dYmean: torch.Tensor,
                     dL_dYvar: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """``_prepare_gradients`` of a hidden layer (layers.py:582-615): returns the freshly
        written latent gradients (gmean, gvar) [lat_total, X_dim] - output-side gradients on the
        steps t >= X_win plus the entropy / prior gradients (variational.py:4-24) - and the
        value the layer adds to its bound (0-d tensor)."""
        D = self.X_dim
        cols = 1 if dL_dYvar.dim() == 1 else int(dL_dYvar.shape[1])
        if tuple(dL_dYmean.shape) != (self.N, D) or dL_dYvar.shape[0] != self.N or cols not in (1, D):
            raise ValueError("dL_dYmean must be [N, D] and dL_dYvar [N] or [N, D]")
        _check(lat_mean, "lat_mean", (self.lat_total, D))
        _check(lat_var, "lat_var", (self.lat_total, D))
        gm, gv = torch.empty_like(lat_mean), torch.empty_like(lat_var)
        val = torch.empty((), dtype=torch.float64, device=self.device)
        dym, dyv = dL_dYmean.contiguous(), dL_dYvar.contiguous()
        _check(dym, "dL_dYmean")
        _check(dyv, "dL_dYvar")
        self.handle.latent_terms(self._stream(), self.nseq, self.desc.data_ptr(), self.X_win, D,
                                 lat_mean.data_ptr(), lat_var.data_ptr(), self.lat_total,
                                 dym.data_ptr(), dyv.data_ptr(), cols, gm.data_ptr(), gv.data_ptr(), val.data_ptr())
        return gm, gv, val
