"""One fused statistics + gradients pass at the headline tile shape (for ncu captures)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rgp_b200.device import DevicePsi
dev = torch.device("cuda", 0)
N, M, Q = (int(sys.argv[1]) if len(sys.argv) > 1 else 65536), 512, 64
g = torch.Generator(device=dev).manual_seed(1); f64 = dict(dtype=torch.float64, device=dev)
mu = torch.randn((N, Q), generator=g, **f64); S = torch.rand((N, Q), generator=g, **f64) * 0.49 + 0.01
Z = torch.randn((M, Q), generator=g, **f64); ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * Q ** 0.5
dL1 = torch.randn((N, M), generator=g, **f64) / M; dL2 = torch.randn((M, M), generator=g, **f64) / M ** 2
dp = DevicePsi(0)
for _ in range(2):
    out = dp.fused(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
torch.cuda.synchronize()
print("ok", float(out[0][1].sum()))
