"""A/B of the small-inducing-set kernels (psi2_small.cuh) against the 64 x 64 block kernels: per-kernel times of one
forward + backward (+ fused pass) per shape.   python scripts/small_ab.py [N ["M,Q;M,Q" ["small_m,ks,warps;..."]]]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rgp_b200.device import DevicePsi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
dev = torch.device("cuda", 0)
f64 = dict(dtype=torch.float64, device=dev)
shapes = [(100, 20), (100, 40), (50, 20), (50, 40), (100, 30), (112, 46), (33, 20)]
if len(sys.argv) > 2:       # "M,Q;M,Q;..."
    shapes = [tuple(int(x) for x in p.split(",")) for p in sys.argv[2].split(";")]
variants = ((0, 0, 0), (1, 0, 0), (1, 4, 0), (1, 2, 0))
if len(sys.argv) > 3:       # "small_m,ks,warps;..."
    variants = [tuple(int(x) for x in p.split(",")) for p in sys.argv[3].split(";")]
for M, Q in shapes:
    g = torch.Generator(device=dev).manual_seed(1)
    mu = torch.randn((N, Q), generator=g, **f64); S = torch.rand((N, Q), generator=g, **f64) * 0.49 + 0.01
    Z = torch.randn((M, Q), generator=g, **f64); ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * Q ** 0.5
    dL1 = torch.randn((N, M), generator=g, **f64) / M
    dL2 = torch.randn((M, M), generator=g, **f64) / (M * M)
    for small_m, ks, warps in variants:
        dp = DevicePsi(0)
        dp.handle.set_option("small_m", small_m)
        dp.handle.set_option("small_ks", ks)
        dp.handle.set_option("small_warps", warps)
        for _ in range(2):
            dp.forward(mu, S, Z, ell, 1.3); dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2); dp.fused(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
        dp.handle.set_option("profile", 1); dp.handle.reset_counters()
        reps = 3
        for _ in range(reps):
            dp.forward(mu, S, Z, ell, 1.3); dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2); dp.fused(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
        kt = dp.handle.kernel_times()
        print(json.dumps({"N": N, "M": M, "Q": Q, "small_m": small_m, "ks": ks, "warps": warps,
                          "kernels_ms": {k: round(v[0] / reps, 3) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])[:6]}}),
              flush=True)
