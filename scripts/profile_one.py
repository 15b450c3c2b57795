"""Forward + backward three times at one shape, for ncu:  ncu --set full -k regex:k_psi2 -s 2 -c 2 python scripts/profile_one.py [rows M Q [small_m]]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rgp_b200.device import DevicePsi
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
M = int(sys.argv[2]) if len(sys.argv) > 2 else 512
Q = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1); f64 = dict(dtype=torch.float64, device=dev)
mu = torch.randn((rows, Q), generator=g, **f64); S = torch.rand((rows, Q), generator=g, **f64) * 0.49 + 0.01
Z = torch.randn((M, Q), generator=g, **f64); ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * Q ** 0.5
dL1 = torch.randn((rows, M), generator=g, **f64) / M
dL2 = torch.randn((M, M), generator=g, **f64) / (M * M)
dp = DevicePsi(0)
if len(sys.argv) > 4:
    dp.handle.set_option("small_m", int(sys.argv[4]))
for _ in range(3):
    dp.forward(mu, S, Z, ell, 1.3)
    dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
torch.cuda.synchronize()
