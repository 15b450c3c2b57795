"""Per-kernel time split of one forward + backward at a given shape (CUDA events per launch, profile mode).
    python scripts/kernel_times.py N M Q [bwd_pipe]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rgp_b200.device import DevicePsi
N, M, Q = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1); f64 = dict(dtype=torch.float64, device=dev)
mu = torch.randn((N, Q), generator=g, **f64); S = torch.rand((N, Q), generator=g, **f64) * 0.49 + 0.01
Z = torch.randn((M, Q), generator=g, **f64); ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * Q ** 0.5
dL1 = torch.randn((N, M), generator=g, **f64) / M
dL2 = torch.randn((M, M), generator=g, **f64) / (M * M)
dp = DevicePsi(0)
if len(sys.argv) > 4:
    dp.handle.set_option("bwd_pipe", int(sys.argv[4]))
for _ in range(2):
    dp.forward(mu, S, Z, ell, 1.3); dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
dp.handle.set_option("profile", 1); dp.handle.reset_counters()
reps = 3
for _ in range(reps):
    dp.forward(mu, S, Z, ell, 1.3); dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
kt = dp.handle.kernel_times()
tot = sum(v[0] for v in kt.values())
print(json.dumps({"N": N, "M": M, "Q": Q, "bwd_pipe": sys.argv[4] if len(sys.argv) > 4 else "default", "total_ms": tot / reps,
                  "kernels": {k: round(v[0] / reps, 3) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])}}))
