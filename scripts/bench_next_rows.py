"""Measurement for the SURVEY.md 8(f) rows built so far:
  f1  device-resident VarDTC bound: ms per ELBO+gradient evaluation of one layer
  f2  lag-window gather / scatter: achieved HBM GB/s against MEASURED_PEAKS.json (6546 GB/s)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgp_b200._lib import Handle  # noqa: E402
from rgp_b200.inference import DeviceBound  # noqa: E402
from rgp_b200.lagwindow import LagWindow  # noqa: E402


def timed(fn, reps):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda", 0)
    f64 = dict(dtype=torch.float64, device=dev)
    g = torch.Generator(device=dev).manual_seed(3)
    db = DeviceBound(0)
    for name, N, M, Q, D, reps in [("actuator_hidden", 502, 100, 20, 1, 20), ("mocap", 408, 200, 40, 59, 20),
                                   ("large", 1 << 20, 512, 64, 1, 2)]:
        mu = torch.randn((N, Q), generator=g, **f64)
        S = torch.rand((N, Q), generator=g, **f64) * 0.49 + 0.01
        Z = torch.randn((M, Q), generator=g, **f64)
        ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * Q ** 0.5
        Y = torch.randn((N, D), generator=g, **f64)
        h = db.psi.handle
        h.set_option("profile", 1); h.reset_counters()
        ms = timed(lambda: db.vardtc(1.3, ell, Z, mu, S, Y, 0.1), reps)
        kt = h.kernel_times(); h.set_option("profile", 0)
        ours = sum(v[0] for v in kt.values()) / (reps + 2)
        print(json.dumps({"row": "f1_vardtc_device", "config": name, "N": N, "M": M, "Q": Q, "D": D,
                          "ms_per_eval": ms, "ms_in_librgp_psi_kernels": ours,
                          "ms_in_bound_algebra_and_launch": ms - ours, "rows_per_s": N / (ms * 1e-3)}), flush=True)
    # f2: one long sequence set, X_win = U_win = 10 like config 1 but large
    T, nseq, Xw, Dx, Uw, Du = 1 << 20, 8, 10, 4, 10, 2
    lens = [T] * nseq
    lw = LagWindow(Handle(0), lens, Xw, Dx, [T] * nseq, Uw, Du)
    lat = torch.randn((lw.lat_total, Dx), generator=g, **f64)
    ctl = torch.randn((lw.ctl_total, Du), generator=g, **f64)
    X = torch.empty((lw.N, lw.Q), **f64)
    ms_g = timed(lambda: lw.gather(lat, ctl, out=X), 10)
    dlat = torch.zeros((lw.lat_total, Dx), **f64); dctl = torch.zeros((lw.ctl_total, Du), **f64)
    ms_s = timed(lambda: lw.scatter_add(X, dlat, dctl), 10)
    out_bytes = lw.N * lw.Q * 8
    src_bytes = (lw.lat_total * Dx + lw.ctl_total * Du) * 8
    print(json.dumps({"row": "f2_lag_gather", "N": lw.N, "Q": lw.Q, "ms": ms_g,
                      "algorithmic_GBps": (out_bytes + src_bytes) / ms_g / 1e6, "hbm_peak_GBps": 6546.2,
                      "frac": (out_bytes + src_bytes) / ms_g / 1e6 / 6546.2}), flush=True)
    print(json.dumps({"row": "f2_lag_scatter", "N": lw.N, "Q": lw.Q, "ms": ms_s,
                      "algorithmic_GBps": (out_bytes + 2 * src_bytes) / ms_s / 1e6, "hbm_peak_GBps": 6546.2,
                      "frac": (out_bytes + 2 * src_bytes) / ms_s / 1e6 / 6546.2}), flush=True)


if __name__ == "__main__":
    main()
