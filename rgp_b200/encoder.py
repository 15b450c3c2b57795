"""Recognition model on the device.

The reference's back-constrained model (``DeepAutoreg_rnn``, autoreg/model.py:284-553) does not
optimise q(X) directly: a stack of recurrent networks reads the observations (plus controls) and
emits, level by level, the mean and variance of every latent step
(``Mean_var_multilayer`` / ``Mean_var_rnn``, autoreg/rnn_encoder.py:20-152).  There the network
runs in torch on the CPU and its outputs / gradients are copied through numpy and paramz on every
evaluation (``forward_computation`` :216-254, ``backward_computation`` :256-281).  Here the same
architecture lives on the GPU in fp64, its outputs are handed to the objective as the stacked
latent tensors without leaving HBM, and the objective's latent gradients flow back through
autograd (``rgp_b200.autograd.deep_autoreg_objective``).

Architecture per level l (l = 0 reads the data): one single-layer RNN / GRU / LSTM over the
sequence, zero initial state, then ``mean = Linear(h)`` and ``var = softplus(Linear(h))``; level
l > 0 reads the concatenation [mean, var] of level l-1.  Parameter names match the reference's
(``layer_{l}.rnn.*``, ``layer_{l}.linear_mean.*``, ``layer_{l}.linear_var.*``) so state dicts are
interchangeable; tests/golden/ref_encoder.npz holds outputs and gradients of the reference's own
module for the parity test.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

_CELLS = {"rnn": nn.RNN, "gru": nn.GRU, "lstm": nn.LSTM}


class MeanVarLevel(nn.Module):
    """One level: recurrent cell -> (mean head, softplus variance head)."""

    def __init__(self, input_dim: int, output_dim: int, hidden_dim: int, rnn_type: str, bidirectional: bool):
        super().__init__()
        if rnn_type not in _CELLS:
            raise ValueError("Unknow rnn type")               # the reference's message, rnn_encoder.py:35
        self.rnn = _CELLS[rnn_type](input_size=input_dim, hidden_size=hidden_dim, num_layers=1,
                                    bidirectional=bidirectional)
        width = hidden_dim * (2 if bidirectional else 1)
        self.linear_mean = nn.Linear(width, output_dim)
        self.linear_var = nn.Linear(width, output_dim)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        h, _ = self.rnn(x)                                     # zero initial state (h_0_type='zero')
        return self.linear_mean(h), F.softplus(self.linear_var(h))


class RecognitionEncoder(nn.Module):
    """input_dims[l] / output_dims[l]: per level, lowest first (level 0 reads the observations,
    concatenated with the controls when the model has them, model.py:407-411)."""

    def __init__(self, input_dims: Sequence[int], output_dims: Sequence[int], hidden_dim: int,
                 rnn_type: str = "rnn", bidirectional: bool = False):
        super().__init__()
        if len(input_dims) != len(output_dims) or not input_dims:
            raise ValueError("Dim lengths must match")
        self.num_levels = len(input_dims)
        for l, (i, o) in enumerate(zip(input_dims, output_dims)):
            setattr(self, "layer_%d" % l, MeanVarLevel(i if l == 0 else 2 * i, o, hidden_dim, rnn_type, bidirectional))
        self.double()

    def forward(self, x: torch.Tensor) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
        """x [seq_len, batch, input_dims[0]] -> (means, vars), each a list over levels of
        [seq_len, batch, output_dims[l]] tensors."""
        means, variances = [], []
        for l in range(self.num_levels):
            m, v = getattr(self, "layer_%d" % l)(x)
            means.append(m)
            variances.append(v)
            x = torch.cat((m, v), dim=2)
        return means, variances

    def latents(self, x: torch.Tensor) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        """The encoder output in the layout of ``DeviceDeepAutoreg.evaluate``: per level a
        (mean, var) pair stacked sequence after sequence, [batch * seq_len, dim]."""
        means, variances = self.forward(x)
        stack = lambda t: t.permute(1, 0, 2).reshape(-1, t.shape[2]).contiguous()
        return [(stack(m), stack(v)) for m, v in zip(means, variances)]
