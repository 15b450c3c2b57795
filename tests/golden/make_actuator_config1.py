"""Config 1 of BASELINE.json on the REAL Actuator data (build container only; reads
/root/reference/datasets/system_identification/actuator.mat):
    python tests/golden/make_actuator_config1.py

Builds the model the way the reference's benchmark does (autoreg/benchmark/tasks.py:141-158:
u -> p, windows 10/10, first 512 samples; autoreg/benchmark/methods.py:62-71:
DeepAutoreg([0, 10], Y, U=U, U_win=10, X_variance=0.05, RBF ARD kernels, lengthscale =
range(X)/sqrt(2), noise = 0.01 var(Y), kernel variance 1; model.py:52-66 alignment,
:128-141 latent init 'Y'; layers.py:250-255 k-means inducing inputs, M = 100, seeded here),
evaluates it once with the CPU oracle and stores inputs + bound + every gradient in
tests/golden/actuator_config1.npz.  The GPU box compares the device model with this file and
trains from it (scripts/train_actuator.py).

Conditioning.  At this initialisation K(Z,Z) + 1e-6 I has a condition number of 1e7-1e8 (a
smooth signal, lengthscales = range/sqrt(2)): the oracle's OWN outputs move by 1e-9 (bound) to
1e-6 (dZ) when Z is perturbed by one part in 1e15.  The fixture therefore also stores, per
output block, that measured sensitivity (``sens_*`` = max relative change over three random
1e-15 relative perturbations of Z); a comparison against this file is meaningful down to a
small multiple of it, not down to 1e-9.  (The synthetic shapes used for the 1e-9 parity bar
are well conditioned.)"""
import os
import sys

import numpy as np
import copy

import scipy.io
from sklearn.cluster import KMeans

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.lag_oracle import build_rows  # noqa: E402
from oracle.model_oracle import deep_autoreg_oracle  # noqa: E402

if __name__ == "__main__":
    d = scipy.io.loadmat("/root/reference/datasets/system_identification/actuator.mat")
    U_all, Y_all = np.asarray(d["u"], dtype=np.float64)[:512], np.asarray(d["p"], dtype=np.float64)[:512]
    win_in = win_out = 10
    M = 100
    wins, U_win = [0, win_out], win_in
    U = U_all[:-1].copy()                                  # model.py:57-62 (U_pre_step)
    Y = Y_all[U_win:].copy()
    T = Y.shape[0]
    Us = [(U, np.full(U.shape, 1e-10))]                     # model.py:65
    lat_mean = np.zeros((win_out + T, 1))                   # model.py:134-140 (init='Y')
    lat_mean[win_out:] = Y[:, :1]
    latents = [[(lat_mean, np.full(lat_mean.shape, 0.05))]]
    rng = np.random.default_rng(1)
    params = []
    for i in range(2):
        if i == 1:      # top layer: own window + controls
            X = build_rows([lat_mean], [U], win_out, U_win)
        else:           # observed layer: the hidden level's window
            X = build_rows([np.zeros((T, 0))], [lat_mean], 0, win_out)
        ell = (X.max(0) - X.min(0)) / np.sqrt(2.0)          # methods.py:68 (inv_l = 1/ell... sqrt form)
        Z = KMeans(n_clusters=M, n_init=10, max_iter=100, random_state=i).fit(X).cluster_centers_.copy()
        params.append(dict(variance=1.0, lengthscale=ell, Z=Z, noise_variance=0.01 * float(Y_all.var())))
    logL, res, lat_grads, ctl_grads = deep_autoreg_oracle(wins, [Y], latents, params, Us=Us, U_win=U_win)
    out = dict(Y=Y, U=U, lat_mean=lat_mean, lat_var=latents[0][0][1], logL=logL,
               g_lat_mean=lat_grads[0][0][0], g_lat_var=lat_grads[0][0][1],
               g_ctl_mean=ctl_grads[0][0], g_ctl_var=ctl_grads[0][1])
    sens = {}
    for trial in range(3):
        pp = copy.deepcopy(params)
        for p in pp:
            p["Z"] = p["Z"] * (1.0 + 1e-15 * rng.normal(size=p["Z"].shape))
        L2, r2, lg2, cg2 = deep_autoreg_oracle(wins, [Y], latents, pp, Us=Us, U_win=U_win)

        def upd(key, a, b):
            a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
            sens[key] = max(sens.get(key, 0.0), float(np.abs(a - b).max() / np.abs(b).max()))
        upd("sens_logL", L2, logL)
        for i in range(2):
            for k in ("variance", "lengthscale", "Z", "noise_variance"):
                upd("sens_g%d_%s" % (i, k), r2[i][k], res[i][k])
        upd("sens_g_lat_mean", lg2[0][0][0], lat_grads[0][0][0])
        upd("sens_g_lat_var", lg2[0][0][1], lat_grads[0][0][1])
        upd("sens_g_ctl_mean", cg2[0][0], ctl_grads[0][0])
    out.update(sens)
    print({k: "%.1e" % v for k, v in sens.items()})
    for i, (p, r) in enumerate(zip(params, res)):
        for k, v in p.items():
            out["p%d_%s" % (i, k)] = v
        for k in ("variance", "lengthscale", "Z", "noise_variance"):
            out["g%d_%s" % (i, k)] = r[k]
    np.savez_compressed(os.path.join(HERE, "actuator_config1.npz"), **out)
    print("actuator_config1.npz  logL =", logL, " T =", T)
