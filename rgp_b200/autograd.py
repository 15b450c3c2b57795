"""Autograd bridge: the device objective as one differentiable torch node.

``DeviceDeepAutoreg.evaluate`` returns the bound and its analytic gradients (computed by
librgp_psi and the bound algebra, not by autograd).  ``deep_autoreg_objective`` wraps that in a
``torch.autograd.Function`` so that whatever produced the latent tensors (the recognition model,
``rgp_b200.encoder``) or the kernel parameters (e.g. a positivity transform) receives those
gradients through ``.backward()`` - what ``DeepAutoreg_rnn.parameters_changed`` does by hand with
``encoder.backward_computation`` (autoreg/model.py:525-553, rnn_encoder.py:256-281).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

_KEYS = ("variance", "lengthscale", "Z", "noise_variance", "qU_mean", "qU_W", "qU_a")


class _Objective(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, layout, Y, controls, *tensors):
        params, latents = _unflatten(layout, tensors)
        logL, res, lat_grads, _ = model.evaluate(params, Y, latents, controls)
        grads: List[Optional[torch.Tensor]] = []
        for (kind, i, key) in layout:
            if kind == "p":
                g = res[i][key]
                grads.append(g if isinstance(g, torch.Tensor) else torch.as_tensor(g, dtype=torch.float64))
            else:
                grads.append(lat_grads[i][key])
        ctx.grads = grads
        ctx.shapes = [t.shape for t in tensors]
        return logL.reshape(()) if isinstance(logL, torch.Tensor) else torch.as_tensor(logL, dtype=torch.float64)

    @staticmethod
    def backward(ctx, grad_out):
        out = [None, None, None, None]
        for g, shp in zip(ctx.grads, ctx.shapes):
            out.append((grad_out * g.to(grad_out.device)).reshape(shp))
        return tuple(out)


def _unflatten(layout, tensors):
    n_layers = 1 + max(i for kind, i, _ in layout if kind == "p")
    n_levels = 1 + max(i for kind, i, _ in layout if kind == "l")
    params: List[Dict] = [dict() for _ in range(n_layers)]
    lat: List[List[Optional[torch.Tensor]]] = [[None, None] for _ in range(n_levels)]
    for (kind, i, key), t in zip(layout, tensors):
        if kind == "p":
            params[i][key] = t.detach()
        else:
            lat[i][key] = t.detach().contiguous()
    return params, [tuple(p) for p in lat]


def deep_autoreg_objective(model, params: Sequence[Dict], Y: torch.Tensor,
                           latents: Sequence[Tuple[torch.Tensor, torch.Tensor]],
                           controls: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> torch.Tensor:
    """Bound of the deep autoregressive model as a 0-d tensor wired into autograd.

    ``params[i]`` holds the level-i layer parameters as tensors (0-d for variance /
    noise_variance / qU_a); every tensor with ``requires_grad`` - and every latent tensor -
    receives its gradient on ``.backward()``.  Scalars given as Python floats are constants."""
    layout, tensors, const = [], [], [dict() for _ in params]
    for i, p in enumerate(params):
        for key in _KEYS:
            if key not in p:
                continue
            if isinstance(p[key], torch.Tensor):
                layout.append(("p", i, key))
                tensors.append(p[key])
            else:
                const[i][key] = p[key]
        for key in p:
            if key not in _KEYS:
                const[i][key] = p[key]
    for i, (m, v) in enumerate(latents):
        layout += [("l", i, 0), ("l", i, 1)]
        tensors += [m, v]
    return _Objective.apply(_WithConstants(model, const), tuple(layout), Y, controls, *tensors)


class _WithConstants:
    """Merges the non-tensor parameters back in before calling the model."""

    def __init__(self, model, const):
        self.model, self.const = model, const

    def evaluate(self, params, Y, latents, controls):
        merged = []
        for p, c in zip(params, self.const):
            q = dict(c)
            q.update(p)
            merged.append(q)
        return self.model.evaluate(merged, Y, latents, controls)
