"""One SVI evaluation (bound + every gradient) of one layer on the device: two-phase order of the
reference (Psi forward, bound algebra, Psi backward) against the fused order (upstream gradients
first, statistics + gradients from one pass, rgp_psi_fused_dev)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rgp_b200.inference import DeviceBound

dev = torch.device("cuda", 0); f64 = dict(dtype=torch.float64, device=dev)
g = torch.Generator(device=dev).manual_seed(2)
db = DeviceBound(0)
for N, M, Q, D in ((1 << 18, 512, 64, 1), (1 << 16, 1024, 64, 2), (1 << 18, 256, 32, 1), (502, 100, 20, 1)):
    mu = torch.randn((N, Q), generator=g, **f64); S = torch.rand((N, Q), generator=g, **f64) * 0.49 + 0.01
    Z = torch.randn((M, Q), generator=g, **f64); ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * Q ** 0.5
    Y = torch.randn((N, D), generator=g, **f64)
    qm = torch.randn((M, D), generator=g, **f64); W = torch.randn((M, M), generator=g, **f64) * 0.05
    qv = W @ W.mT + 0.5 * torch.eye(M, **f64)
    rec = {"row": "svi_layer_eval", "N": N, "M": M, "Q": Q, "D": D}
    outs = {}
    for mode in (False, True):
        fn = lambda: db.svi(1.3, ell, Z, mu, S, Y, 0.1, qm, qv, 1.0, fused=mode)
        outs[mode] = fn(); torch.cuda.synchronize()
        reps = 3 if N > 100000 else 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        rec["fused_ms" if mode else "two_phase_ms"] = e0.elapsed_time(e1) / reps
    a, b = outs[False], outs[True]
    rec["speedup"] = rec["two_phase_ms"] / rec["fused_ms"]
    rec["max_rel_diff"] = max([abs(float(a[0]) - float(b[0])) / abs(float(a[0]))] +
                              [float((a[1][k] - b[1][k]).abs().max() / a[1][k].abs().max())
                               for k in ("variance", "lengthscale", "Z", "mu", "S", "dL_dqU_mean", "dL_dqU_var")])
    print(json.dumps(rec), flush=True)
