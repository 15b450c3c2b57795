import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rgp_b200.device import DevicePsi
dp = DevicePsi(0); dev = torch.device("cuda", 0)
dp.handle.set_option("bwd_warps", 16)   # these experiments instrument the 16-warp kernel
N, M, Q = 1 << 16, 512, 64
g = torch.Generator(device=dev).manual_seed(1); f64 = dict(dtype=torch.float64, device=dev)
mu = torch.randn((N, Q), generator=g, **f64); S = torch.rand((N, Q), generator=g, **f64) * 0.49 + 0.01
Z = torch.randn((M, Q), generator=g, **f64); ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * 8.0
dL1 = torch.randn((N, M), generator=g, **f64) / M; dL2 = torch.randn((M, M), generator=g, **f64) / M ** 2
dp.handle.set_option("debug_skip", int(sys.argv[1]) if len(sys.argv) > 1 else 247)
dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
torch.cuda.synchronize()
