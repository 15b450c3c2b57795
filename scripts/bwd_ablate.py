"""Where the Psi2 backward kernel's time goes: the row-at-a-time kernel with pieces removed (debug build,
results are wrong by construction).  RGP_PSI_LIB=rgp_b200/_lib/librgp_psi_debug.so python scripts/bwd_ablate.py"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rgp_b200.device import DevicePsi
rows, M, Q = 65536, 512, 64
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1); f64 = dict(dtype=torch.float64, device=dev)
mu = torch.randn((rows, Q), generator=g, **f64); S = torch.rand((rows, Q), generator=g, **f64) * 0.49 + 0.01
Z = torch.randn((M, Q), generator=g, **f64); ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * Q ** 0.5
dL1 = torch.randn((rows, M), generator=g, **f64) / M
dL2 = torch.randn((M, M), generator=g, **f64) / (M * M); dL2 = 0.5 * (dL2 + dL2.T)
NAMES = {0: "full kernel", 1: "no exp", 2: "no lambda sums/flush", 3: "no exp, no lambda", 4: "no stage 2-I folds / W",
         7: "no exp, lambda, folds", 8: "no epilogue (exp, C, L store)", 14: "skeleton: DMMA loops + LDS + barrier",
         30: "skeleton, no barrier", 46: "skeleton without stage 2-J", 78: "skeleton without stage 2-I",
         142: "skeleton without stage 1", 16: "full, no barrier", 110: "stage 1 only", 174: "stage 2-I only",
         206: "stage 2-J only"}
dp = DevicePsi(0)
dp.handle.set_option("bwd_pipe", 0)
for mask in [0, 1, 2, 3, 4, 7, 8, 14, 30, 16, 46, 78, 142, 110, 174, 206]:
    dp.handle.set_option("debug_skip", mask)
    dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
    dp.handle.set_option("profile", 1); dp.handle.reset_counters()
    for _ in range(2):
        dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
    kt = dp.handle.kernel_times(); dp.handle.set_option("profile", 0)
    print(json.dumps({"mask": mask, "what": NAMES[mask], "psi2_bwd_ms": kt["psi2_bwd"][0] / kt["psi2_bwd"][1]}), flush=True)
dp.handle.set_option("debug_skip", 0)
