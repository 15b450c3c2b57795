"""Times the Psi2 backward kernel variants (16-warp default, 8-warp, strip) at the headline tile
shape and checks that they agree with each other."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rgp_b200.device import DevicePsi
dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
M, Q = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (512, 64)
g = torch.Generator(device=dev).manual_seed(1); f64 = dict(dtype=torch.float64, device=dev)
mu = torch.randn((N, Q), generator=g, **f64); S = torch.rand((N, Q), generator=g, **f64) * 0.49 + 0.01
Z = torch.randn((M, Q), generator=g, **f64); ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * Q ** 0.5
dL1 = torch.randn((N, M), generator=g, **f64) / M; dL2 = torch.randn((M, M), generator=g, **f64) / M ** 2
ref = None
only = os.environ.get("BWD_VARIANTS", "bwd16,bwd8,mbar,strip")  # also: mbar_nolam, mbar_noatom (timing only).split(",")
for name, opts in (("bwd16", {"bwd_warps": 16}), ("bwd8", {"bwd_warps": 8}), ("mbar", {"bwd_mbar": 1}), ("mbar_nolam", {"bwd_mbar": 1, "debug_skip": 1}),
                   ("mbar_noatom", {"bwd_mbar": 1, "debug_skip": 3}),
                   ("strip", {"bwd_strip": 1})):
    if name not in only:
        continue
    dp = DevicePsi(0)
    for k, v in opts.items():
        dp.handle.set_option(k, v)
    out = dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
    torch.cuda.synchronize()
    dp.handle.set_option("profile", 1); dp.handle.reset_counters()
    for _ in range(3):
        out = dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
    kt = dp.handle.kernel_times(); dp.handle.set_option("profile", 0)
    rec = {"variant": name, "N": N, "M": M, "Q": Q, "psi2_bwd_ms": kt["psi2_bwd"][0] / kt["psi2_bwd"][1]}
    if ref is None:
        ref = out
    else:
        rec["max_rel_diff_vs_bwd16"] = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(out, ref))
    print(json.dumps(rec), flush=True)
