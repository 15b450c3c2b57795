"""One ablation mask of scripts/bwd_ablate.py, for ncu.  usage: bwd_ablate_one.py MASK [rows]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rgp_b200.device import DevicePsi
mask = int(sys.argv[1]); rows = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
M, Q = 512, 64
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1); f64 = dict(dtype=torch.float64, device=dev)
mu = torch.randn((rows, Q), generator=g, **f64); S = torch.rand((rows, Q), generator=g, **f64) * 0.49 + 0.01
Z = torch.randn((M, Q), generator=g, **f64); ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * Q ** 0.5
dL1 = torch.randn((rows, M), generator=g, **f64) / M
dL2 = torch.randn((M, M), generator=g, **f64) / (M * M); dL2 = 0.5 * (dL2 + dL2.T)
dp = DevicePsi(0)
dp.handle.set_option("bwd_pipe", 0)
if mask:
    dp.handle.set_option("debug_skip", mask)
for _ in range(3):
    dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
torch.cuda.synchronize()
