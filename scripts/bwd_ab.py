"""A/B timing of the Psi2 kernels on one GPU ("roles" = the warp-specialised experimental kernel: needs
RGP_PSI_LIB=rgp_b200/_lib/librgp_psi_debug.so) (CUDA events per launch through the handle's profile mode).
    python scripts/bwd_ab.py [rows] [M] [Q]
Prints one JSON line per variant: ms per launch and the fraction of the in-run DFMA peak."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rgp_b200.device import DevicePsi

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
M = int(sys.argv[2]) if len(sys.argv) > 2 else 512
Q = int(sys.argv[3]) if len(sys.argv) > 3 else 64
variants = sys.argv[4].split(",") if len(sys.argv) > 4 else ["default", "pipe", "rowloop"]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1); f64 = dict(dtype=torch.float64, device=dev)
mu = torch.randn((rows, Q), generator=g, **f64); S = torch.rand((rows, Q), generator=g, **f64) * 0.49 + 0.01
Z = torch.randn((M, Q), generator=g, **f64); ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * Q ** 0.5
dL1 = torch.randn((rows, M), generator=g, **f64) / M
dL2 = torch.randn((M, M), generator=g, **f64) / (M * M); dL2 = 0.5 * (dL2 + dL2.T)
P = M * (M + 1) // 2
fl = {"psi2_bwd": 2 * P * Q + 2 * M * M * Q + 8 * P, "psi2_fwd": 2 * P * Q + 8 * P,
      "psi2_bwd_fused": 2 * P * Q + 2 * M * M * Q + 8 * P}
ref = None
for name in variants:
    dp = DevicePsi(0)
    peak = dp.handle.fp64_peak(reps=3)
    dp.handle.set_option("bwd_pipe", {"pipe": 1, "rowloop": 0, "roles": 3}.get(name, 2))
    for _ in range(2):
        dp.forward(mu, S, Z, ell, 1.3)
        out = dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
    dp.handle.set_option("profile", 1); dp.handle.reset_counters()
    for _ in range(3):
        _, _, p2 = dp.forward(mu, S, Z, ell, 1.3)
        out = dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
    (q1, q2), fo = dp.fused(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
    (q1, q2), fo = dp.fused(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
    kt = dp.handle.kernel_times(); dp.handle.set_option("profile", 0)
    res = {"variant": name, "rows": rows, "M": M, "Q": Q, "peak_tflops": peak}
    for k in ("psi2_fwd", "psi2_bwd", "psi2_bwd_fused"):
        if k in kt:
            ms = kt[k][0] / kt[k][1]
            passes = 2 if (k == "psi2_bwd" and Q > 64) else 1
            res[k + "_ms"] = ms * passes
            res[k + "_frac"] = fl[k] * rows / (ms * passes * 1e-3) / 1e12 / peak
    cur = [t.double().clone() for t in out] + [p2.clone(), q2.clone()] + [t.clone() for t in fo]
    if ref is None:
        ref = cur
    else:
        res["max_rel_vs_first"] = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(cur, ref))
    res["fused_vs_two_phase"] = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip([q2] + list(fo), [p2] + list(out)))
    print(json.dumps(res), flush=True)
