"""Tuning probe: forward kernel with 2 CTAs/SM (default) vs forced 1 CTA/SM (smem pad)."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rgp_b200.device import DevicePsi
dp = DevicePsi(0); dev = torch.device("cuda", 0)
N, M, Q = 1 << 19, 512, 64
g = torch.Generator(device=dev).manual_seed(1); f64 = dict(dtype=torch.float64, device=dev)
mu = torch.randn((N, Q), generator=g, **f64); S = torch.rand((N, Q), generator=g, **f64) * 0.49 + 0.01
Z = torch.randn((M, Q), generator=g, **f64); ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * 8.0
for pad in (0, 60000):
    dp.handle.set_option("fwd_smem_pad", pad)
    for _ in range(2): dp.forward(mu, S, Z, ell, 1.3, want_psi1=False)
    dp.handle.set_option("profile", 1); dp.handle.reset_counters()
    for _ in range(3): dp.forward(mu, S, Z, ell, 1.3, want_psi1=False)
    kt = dp.handle.kernel_times(); dp.handle.set_option("profile", 0)
    print(json.dumps({"fwd_smem_pad": pad, "psi2_fwd_ms": kt["psi2_fwd"][0] / kt["psi2_fwd"][1]}), flush=True)
