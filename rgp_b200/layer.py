"""One objective evaluation of the deep autoregressive model, resident on one GPU
(SURVEY.md 8 f1 + f2 composed around the psi path).

Mirrors, for models without back-constraints:
  * ``Layer_new.update_layer``  autoreg/layers.py:617-621: ``_update_X`` (:528-550, lag-window
    rows), ``_inference_vardtc`` (:66-134, bound + parameter gradients), ``_update_qX_gradients``
    (:574-580), ``_prepare_gradients`` (:582-615, output-side gradients + latent prior /
    entropy, autoreg/variational.py:4-24);
  * ``update_latent_gradients``  (:552-572);
  * ``DeepAutoreg_new.parameters_changed``  autoreg/model.py:159-187 with the layer wiring of
    ``__init__`` (:95-110): layers updated top -> bottom, bound = sum, latent gradients
    scattered bottom -> top.

All sequences of a level live stacked in one [total_steps, dim] tensor pair (mean, variance);
the N x Q rows, Psi1, dL_dpsi1 and the row gradients never leave HBM.  Kernels: librgp_psi
(psi statistics, lag gather / scatter, latent terms); M x M algebra and the N x M GEMM / TRSM
through torch.linalg (cuBLAS / cuSOLVER).  No CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .inference import DeviceBound, tdot
from .lagwindow import LagWindow

Pair = Tuple[torch.Tensor, torch.Tensor]


class DeviceLayer:
    """Geometry + evaluation of one layer.

    lat_lens[s]  steps of this layer's own series (hidden: wins + T_s latent steps; observed:
                 T_s observations, ``X_win`` must be 0 as in the reference, whose
                 ``_update_conv`` cannot window plain arrays either)
    ctl_lens[s]  steps of the series feeding the layer from above (upper latent or controls)
    """

    def __init__(self, bound: DeviceBound, lat_lens: Sequence[int], X_win: int, X_dim: int,
                 ctl_lens: Optional[Sequence[int]], U_win: int, U_dim: int, observed: bool,
                 svi: bool = False, lag=None, device=None):
        if observed and X_win != 0:
            raise ValueError("observed layers are not windowed (X_win must be 0)")
        self.bound, self.observed, self.svi = bound, observed, svi
        self.X_win, self.X_dim, self.U_win, self.U_dim = X_win, X_dim, U_win, U_dim
        self.lat_lens = [int(t) for t in lat_lens]
        handle = bound.psi.handle if hasattr(bound.psi, "handle") else None
        self.lag = lag if lag is not None else LagWindow(handle, lat_lens, X_win, X_dim, ctl_lens, U_win, U_dim,
                                                         device=device)
        self.N, self.Q = self.lag.N, self.lag.Q
        # rows of sequence s are its steps X_win.. : index of the output steps in the stack
        idx, off = [], 0
        for T in self.lat_lens:
            idx.append(torch.arange(off + X_win, off + T))
            off += T
        self.out_index = torch.cat(idx).to(self.lag.device) if X_win > 0 else None

    # ------------------------------------------------------------------ update_layer
    def update(self, p: Dict, lat, ctl: Optional[Pair]) -> Dict:
        """``lat``: observed -> Y [T_total, D]; hidden -> (mean, var) [lat_total, X_dim].
        ``ctl``: (mean, var) [ctl_total, U_dim] or None.  Returns a dict with the layer bound
        ``logL``, parameter gradients, the row gradients ``dmu``/``dS`` [N, Q] and, for hidden
        layers, the prepared latent gradients ``gX`` = (gmean, gvar)."""
        if self.observed:
            Y, Y_var, lm, lv = lat, None, None, None
        else:
            lm, lv = lat
            Y = lm if self.out_index is None else lm.index_select(0, self.out_index)     # :494
            Y_var = lv if self.out_index is None else lv.index_select(0, self.out_index)
        cm, cv = ctl if ctl is not None else (None, None)
        mu = self.lag.gather(lm, cm)                                                    # :528-543
        S = self.lag.gather(lv, cv)
        var, ell, Z = float(p["variance"]), p["lengthscale"], p["Z"]
        out: Dict = {}
        if self.svi:
            M = Z.shape[0]
            qU_var = tdot(p["qU_W"]) + torch.eye(M, dtype=Z.dtype, device=Z.device) * float(p["qU_a"])   # :71
            logL, g = self.bound.svi(var, ell, Z, mu, S, Y, float(p["noise_variance"]), p["qU_mean"], qU_var,
                                     float(p.get("qU_ratio", 1.0)), Y_var=Y_var)
            out["qU_mean"] = g["dL_dqU_mean"]                                           # :174-176
            out["qU_W"] = (g["dL_dqU_var"] + g["dL_dqU_var"].mT) @ p["qU_W"]
            out["qU_a"] = torch.diagonal(g["dL_dqU_var"]).sum()
        else:
            logL, g = self.bound.vardtc(var, ell, Z, mu, S, Y, float(p["noise_variance"]), Y_var=Y_var)
        out.update(variance=g["variance"], lengthscale=g["lengthscale"], Z=g["Z"],
                   noise_variance=g["dL_dthetaL"], dmu=g["mu"], dS=g["S"])
        if not self.observed:                                                           # :582-615
            gm, gv, delta = self.lag.latent_terms(lm, lv, g["dL_dYmean"], g["dL_dYvar"])
            logL = logL + self.bound.allsum([delta.reshape(1)])[0].reshape(())   # local latents, global bound
            out["gX"] = (gm, gv)
        out["logL"] = logL
        return out

    # ------------------------------------------------------ update_latent_gradients
    def scatter(self, res: Dict, lat_grad: Optional[Pair], ctl_grad: Optional[Pair]) -> None:
        """Adds the row gradients of ``res`` onto this layer's own latent gradients and onto
        the gradients of the series above (both in place)."""
        for k, rows in ((0, res["dmu"]), (1, res["dS"])):
            self.lag.scatter_add(rows, lat_grad[k] if lat_grad is not None else None,
                                 ctl_grad[k] if ctl_grad is not None else None, allocate=False)


class DeviceDeepAutoreg:
    """``DeepAutoreg_new`` objective on the device.

    wins[i], nDims[i]   window / dimensionality per level, level 0 = observed layer
    seq_lens[s]         T_s, aligned observation steps per sequence (model.py:52-66)
    ctl_dim             dimensionality of the control series (0 = no controls); the control
                        series of sequence s has T_s + U_win - 1 steps (model.py:57-62)

    Multi-GPU: build ``bound=DeviceBound(device, sharded=True)`` and give every rank ITS
    sequences (``seq_lens``, Y, latents, controls of those sequences only).  Sequences are
    independent given the layer parameters, so rows shard on sequence boundaries with no halo;
    the bound and the parameter gradients come back global, latent gradients stay with the
    rank that owns the sequence.
    """

    def __init__(self, wins: Sequence[int], nDims: Sequence[int], seq_lens: Sequence[int], U_win: int = 1,
                 ctl_dim: int = 0, svi: bool = False, device: Optional[int] = None,
                 bound: Optional[DeviceBound] = None, lag_factory=None):
        L = len(wins)
        if L < 2 or len(nDims) != L:
            raise ValueError("need an observed layer and at least one hidden layer")
        if wins[0] != 0:
            raise ValueError("the observed layer is not windowed (wins[0] must be 0)")
        self.wins, self.nDims, self.seq_lens, self.U_win, self.ctl_dim = list(wins), list(nDims), list(seq_lens), U_win, ctl_dim
        self.bound = bound if bound is not None else DeviceBound(device)
        self.layers: List[DeviceLayer] = []
        for i in range(L):
            top = i == L - 1
            own = [(wins[i] + T) if i > 0 else T for T in seq_lens]
            if top:
                above = [T + U_win - 1 for T in seq_lens] if ctl_dim > 0 else None
                Uw, Ud = (U_win, ctl_dim) if ctl_dim > 0 else (0, 0)
            else:
                above, Uw, Ud = [wins[i + 1] + T for T in seq_lens], wins[i + 1], nDims[i + 1]
            lag = lag_factory(own, wins[i], nDims[i], above, Uw, Ud) if lag_factory is not None else None
            self.layers.append(DeviceLayer(self.bound, own, wins[i], nDims[i], above, Uw, Ud, observed=(i == 0),
                                           svi=svi, lag=lag, device=device))

    def evaluate(self, params: Sequence[Dict], Y: torch.Tensor, latents: Sequence[Pair],
                 controls: Optional[Pair] = None):
        """params[i]: parameters of the level-i layer; Y [sum T_s, nDims[0]]; latents[i-1]:
        (mean, var) of level i, stacked over sequences; controls: (mean, var) stacked.
        Returns (logL, layer_results, latent_grads, control_grads) like the model oracle."""
        self.bound.defer_checks()            # no per-factorisation read-back; checked once below
        out = self._evaluate(params, Y, latents, controls)
        if not self.bound.verify():          # some K(Z,Z) / Lambda needed jitter: careful re-run
            out = self._evaluate(params, Y, latents, controls)
        return out

    def _evaluate(self, params, Y, latents, controls):
        L = len(self.wins)
        res: List[Optional[Dict]] = [None] * L
        for i in range(L - 1, -1, -1):                                   # model.py:176, top first
            top = i == L - 1
            res[i] = self.layers[i].update(params[i], Y if i == 0 else latents[i - 1],
                                           controls if top else latents[i])
        logL = sum(r["logL"] for r in res)                               # :177
        lat_grads = [res[i]["gX"] for i in range(1, L)]
        ctl_grads = None
        if controls is not None:                                         # layers.py:589-592
            ctl_grads = (torch.zeros_like(controls[0]), torch.zeros_like(controls[1]))
        for i in range(L):                                               # :178, lowest first
            top = i == L - 1
            self.layers[i].scatter(res[i], lat_grads[i - 1] if i > 0 else None,
                                   ctl_grads if top else lat_grads[i])
        return logL, res, lat_grads, ctl_grads
