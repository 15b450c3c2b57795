"""CPU oracle for one full objective evaluation of the deep autoregressive model
(SURVEY.md 8 f1 + f2 composed).  TEST INFRASTRUCTURE ONLY (see oracle/psi_oracle.py).

Restates, with the same loops and the same order of in-place additions:
  * ``DeepAutoreg_new.parameters_changed``  autoreg/model.py:159-187 (layers updated top
    -> bottom, bound = sum of layer bounds, latent gradients scattered bottom -> top),
  * the layer wiring of ``DeepAutoreg_new.__init__``  autoreg/model.py:95-110,
  * ``Layer_new.update_layer``  autoreg/layers.py:617-621 =
    ``_update_X`` (:528-550)  ->  ``SparseGP_MPI._inference_vardtc`` (:66-134, :172-176)
    ->  ``_update_qX_gradients`` (:574-580)  ->  ``_prepare_gradients`` (:582-615),
  * ``update_latent_gradients``  (:552-572),
for models without back-constraints (no encoder).  The observed layer must have
``wins[0] == 0`` - the reference's ``_update_conv`` reads ``Xs_flat[i].mean`` and so cannot
window plain observed arrays either.

Pinning: the pieces this file composes (bounds, latent terms, lag windows) are held to outputs
of the reference's own code (tests/golden/ref_*.npz, tests/test_reference_golden.py); the
composition itself (layers.py / model.py need all of GPy to run and cannot be executed here) is
pinned relationally by finite differences of the whole objective with respect to every
parameter block (tests/test_model_oracle.py), the way the reference pins itself with
``model.checkgrad`` (testing/*_tests.py).
"""
from __future__ import annotations

import numpy as np

from . import bound_oracle as bo
from .lag_oracle import build_rows, scatter_rows_into
from .psi_oracle import psi_backward, psi_forward


def layer_oracle(p, Xs, Us, X_win, U_win, observed, top, svi=False,
                 psi_fwd=psi_forward, psi_bwd=psi_backward):
    """One ``Layer_new.update_layer``.

    p        dict(variance, lengthscale, Z, noise_variance [, qU_mean, qU_W, qU_a, qU_ratio])
    Xs       observed: list of T_s x D arrays; hidden: list of (mean, var) pairs
    Us       list of (mean, var) pairs or None
    Returns dict(logL, param grads, row grads dmu/dS, prepared latent grads gX [hidden]).
    """
    if observed:
        assert X_win == 0, "observed layers are not windowed (see module docstring)"
        Xm = [np.zeros((y.shape[0], 0)) for y in Xs]
        Xv = Xm
    else:
        Xm, Xv = [m for m, _ in Xs], [v for _, v in Xs]
    Um = [m for m, _ in Us] if Us is not None else None
    Uv = [v for _, v in Us] if Us is not None else None
    mu = build_rows(Xm, Um, X_win, U_win)                            # layers.py:528-543
    S = build_rows(Xv, Uv, X_win, U_win)
    if observed:
        Y, Y_var = np.vstack([y[X_win:] for y in Xs]), None          # :491-492
    else:
        Y = np.vstack([m[X_win:] for m in Xm])                       # :494
        Y_var = np.vstack([v[X_win:] for v in Xv])
    var, ell, Z = float(p["variance"]), np.asarray(p["lengthscale"], dtype=np.float64), p["Z"]
    M = Z.shape[0]
    psi0, psi1, psi2 = psi_fwd(var, ell, Z, mu, S)
    Kmm = bo.rbf_K(var, ell, Z)
    out = {}
    if svi:
        qU_var = bo.tdot(p["qU_W"]) + np.eye(M) * p["qU_a"]          # :71
        ratio = p.get("qU_ratio", 1.0)
        logL, g, mid = bo.svi_vardtc_inference(psi0, psi1, psi2, Kmm, Y, p["noise_variance"],
                                               p["qU_mean"], qU_var, Y_var=Y_var)
        KL, dKL_dm, dKL_dv, dKL_dK = bo.svi_kl_qu(p["qU_mean"], qU_var, mid)   # :75-79
        logL += -KL * ratio
        g["dL_dqU_mean"] = g["dL_dqU_mean"] - dKL_dm * ratio
        g["dL_dqU_var"] = g["dL_dqU_var"] - dKL_dv * ratio
        g["dL_dKmm"] = g["dL_dKmm"] - dKL_dK * ratio
        out["qU_mean"] = g["dL_dqU_mean"]                            # :174-176
        out["qU_W"] = (g["dL_dqU_var"] + g["dL_dqU_var"].T) @ p["qU_W"]
        out["qU_a"] = np.diag(g["dL_dqU_var"]).sum()
    else:
        logL, g = bo.vardtc_inference(psi0, psi1, psi2, Kmm, Y, p["noise_variance"], Y_var=Y_var)
    dvar, dl, dZ, dmu, dS = psi_bwd(g["dL_dpsi0"], g["dL_dpsi1"], g["dL_dpsi2"], var, ell, Z, mu, S)
    kvar, kl, kZ = bo.rbf_K_grads(g["dL_dKmm"], var, ell, Z)
    out.update(variance=dvar + kvar, lengthscale=np.asarray(dl) + kl, Z=dZ + kZ,
               noise_variance=g["dL_dthetaL"], dmu=dmu, dS=dS)
    if not observed:                                                 # _prepare_gradients :582-615
        gX, Y_off, delta = [], 0, 0.0
        for m, v in Xs:
            N = m.shape[0] - X_win
            gm, gv = np.zeros_like(m), np.zeros_like(v)
            gm[X_win:] += g["dL_dYmean"][Y_off:Y_off + N]
            dyv = g["dL_dYvar"][Y_off:Y_off + N]
            gv[X_win:] += dyv if dyv.ndim == 2 else dyv[:, None]     # :601-604
            if X_win > 0:
                val, dm_, dv_ = bo.normal_prior_term(m[:X_win], v[:X_win])
                delta += val
                gm[:X_win] += dm_
                gv[:X_win] += dv_
            val, dv_ = bo.normal_entropy_term(v[X_win:])
            delta += val
            gv[X_win:] += dv_
            gX.append((gm, gv))
            Y_off += N
        logL += delta
        out["gX"] = gX
    if Us is not None and top:                                       # :589-592
        out["gU"] = [(np.zeros_like(m), np.zeros_like(v)) for m, v in Us]
    out["logL"] = float(logL)
    return out


def deep_autoreg_oracle(wins, Ys, latents, params, Us=None, U_win=1, svi=False,
                        psi_fwd=psi_forward, psi_bwd=psi_backward):
    """One ``parameters_changed`` of a DeepAutoreg_new model (model.py:159-187).

    wins      window per level, level 0 = observed layer (must be 0)
    Ys        list over sequences of T_s x D observations (already aligned, model.py:52-66)
    latents   latents[i-1] = level i (1..L-1): list over sequences of (mean, var), each
              (wins[i] + T_s) x nDims[i]   (model.py:128-156)
    params    params[i] = parameter dict of the level-i layer
    Us        list over sequences of (mean, var) control series (var = 1e-10, model.py:65) or None
    Returns (logL, layer_results[level], latent_grads[i-1][seq] = (gmean, gvar), control grads).
    """
    L = len(wins)
    assert L >= 2 and len(latents) == L - 1
    res = [None] * L
    for i in range(L - 1, -1, -1):                                   # top layer first, model.py:176
        top = i == L - 1
        Xs = Ys if i == 0 else latents[i - 1]
        ctl = Us if top else latents[i]
        res[i] = layer_oracle(params[i], Xs, ctl, wins[i], U_win if top else wins[i + 1],
                              observed=(i == 0), top=top, svi=svi, psi_fwd=psi_fwd, psi_bwd=psi_bwd)
    logL = float(np.sum([r["logL"] for r in res]))                   # :177
    lat_grads = [res[i]["gX"] for i in range(1, L)]
    ctl_grads = res[L - 1].get("gU")
    for i in range(L):                                               # lowest layer first, :178
        top = i == L - 1
        Uw = U_win if top else wins[i + 1]
        ctl = Us if top else latents[i]
        gtarget_U = ctl_grads if top else lat_grads[i]
        U_dim = ctl[0][0].shape[1] if ctl is not None else 0
        for k, rows in ((0, res[i]["dmu"]), (1, res[i]["dS"])):
            if i == 0:          # observations carry no gradient: zero-width stand-ins give N_s
                gX = [np.zeros((y.shape[0], 0)) for y in Ys]
                X_dim = 0
            else:
                gX = [g[k] for g in lat_grads[i - 1]]
                X_dim = gX[0].shape[1]
            gU = [g[k] for g in gtarget_U] if gtarget_U is not None else None
            scatter_rows_into(rows, gX, gU, wins[i], Uw, X_dim, U_dim)
    return logL, res, lat_grads, ctl_grads
