"""MLP back-constraint of a hidden layer on the device (second half).

``DeepAutoreg_new(back_cstr=True)`` does not optimise the latent means directly: the first X_win
means of every sequence are parameters (``init_Xs``) and each later mean is produced by a small MLP
from the window before it and the aligned control / upper-layer window (``_encoder_freerun``,
autoreg/layers.py:623-666; network autoreg/mlp.py, theano); the gradient is back-propagated through
that recurrence step by step (``_encoder_update_gradient``, :668-715).  In the reference both are
Python loops over time steps around a theano call.  Here each direction is ONE kernel launch
(librgp_psi ``rgp_mlp_freerun_dev`` / ``rgp_mlp_freerun_bwd_dev``: one CTA per sequence, weights in
shared memory), wrapped in a ``torch.autograd.Function`` so it composes with
``rgp_b200.autograd.deep_autoreg_objective`` and with the encoders of the layers above.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from .lagwindow import LagWindow


def default_units(Q: int, X_dim: int) -> List[int]:
    return [Q, 2 * Q, Q + X_dim // 2, X_dim]                        # layers.py:441


class _FreeRun(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, init_means, ctl_mean, flat):
        lw: LagWindow = mod.lag
        lat = torch.empty((lw.lat_total, lw.X_dim), dtype=torch.float64, device=flat.device)
        lat.index_copy_(0, mod.init_index, init_means.reshape(-1, lw.X_dim))
        acts = torch.empty((lw.N, max(mod.nhid, 1)), dtype=torch.float64, device=flat.device)
        flat = flat.contiguous()
        ctl = ctl_mean.contiguous() if ctl_mean is not None else None
        lw.handle.mlp_freerun(lw._stream(), lw.nseq, lw.desc.data_ptr(), lw.X_win, lw.X_dim, lw.U_win, lw.U_dim,
                              mod.units, flat.data_ptr(), lat.data_ptr(), ctl.data_ptr() if ctl is not None else None,
                              acts.data_ptr())
        # save_for_backward (not a plain attribute): the output `lat` is among the saved tensors, and a
        # Python attribute would close an output -> grad_fn -> ctx -> output cycle the GC cannot see
        ctx.mod, ctx.has_ctl = mod, ctl is not None
        ctx.save_for_backward(flat, lat, acts, *((ctl,) if ctl is not None else ()))
        ctx.init_shape = init_means.shape
        return lat

    @staticmethod
    def backward(ctx, g_lat):
        mod = ctx.mod
        flat, lat, acts = ctx.saved_tensors[:3]
        ctl = ctx.saved_tensors[3] if ctx.has_ctl else None
        lw: LagWindow = mod.lag
        g = g_lat.contiguous().clone()                              # updated in place by the kernel
        g_ctl = torch.zeros_like(ctl) if ctl is not None else None
        pg = torch.empty((lw.nseq, flat.numel()), dtype=torch.float64, device=flat.device)
        lw.handle.mlp_freerun_bwd(lw._stream(), lw.nseq, lw.desc.data_ptr(), lw.X_win, lw.X_dim, lw.U_win, lw.U_dim,
                                  mod.units, flat.data_ptr(), lat.data_ptr(), ctl.data_ptr() if ctl is not None else None,
                                  acts.data_ptr(), g.data_ptr(), g_ctl.data_ptr() if g_ctl is not None else None,
                                  pg.data_ptr())
        g_init = g.index_select(0, mod.init_index).reshape(ctx.init_shape)
        return None, g_init, g_ctl, pg.sum(dim=0)


class MLPBackConstraint(nn.Module):
    """``lag``: the layer's LagWindow (sequence geometry, windows, dims).  ``MLP_dims``: hidden widths
    (None = the reference's default [2 Q, Q + X_dim / 2])."""

    def __init__(self, lag: LagWindow, MLP_dims: Optional[Sequence[int]] = None):
        super().__init__()
        if lag.X_win <= 0:
            raise ValueError("Neural Network constraints only applies autoregressive structure!")   # layers.py:438
        self.lag = lag
        Q = lag.Q
        self.units = default_units(Q, lag.X_dim) if MLP_dims is None else [Q] + list(MLP_dims) + [lag.X_dim]
        self.nhid = sum(self.units[1:-1])
        shapes = []
        for up, down in zip(self.units[:-1], self.units[1:]):
            shapes += [(down, up), (down,)]
        self.shapes = shapes
        flat = []
        for up, down in zip(self.units[:-1], self.units[1:]):       # mlp.py:26-30
            flat.append(((torch.rand(down, up, dtype=torch.float64) * 2 - 1) * math.sqrt(6.0 / (up + down))).reshape(-1))
            flat.append(torch.zeros(down, dtype=torch.float64))
        self.flat = nn.Parameter(torch.cat(flat).to(lag.device))    # packed [W0 | b0 | W1 | b1 ...]
        idx, off = [], 0
        for s in range(lag.nseq):
            T = int(lag.desc[s, 3])
            idx.append(torch.arange(off, off + lag.X_win))
            off += T
        self.init_index = torch.cat(idx).to(lag.device)

    def layer_params(self):
        """[(W [down, up], b [down]), ...] views of the packed parameter vector."""
        out, off = [], 0
        for up, down in zip(self.units[:-1], self.units[1:]):
            W = self.flat[off:off + down * up].reshape(down, up)
            off += down * up
            out.append((W, self.flat[off:off + down]))
            off += down
        return out

    def forward(self, init_means: torch.Tensor, ctl_mean: Optional[torch.Tensor] = None) -> torch.Tensor:
        """init_means [nseq, X_win, X_dim]; ctl_mean [ctl_total, U_dim] (stacked) -> latent means
        [lat_total, X_dim], stacked like every other per-level tensor."""
        if (ctl_mean is None) != (self.lag.U_win == 0):
            raise ValueError("control series do not match the layer geometry")
        return _FreeRun.apply(self, init_means, ctl_mean, self.flat)
