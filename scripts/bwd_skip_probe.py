"""Timing experiment: cost of each non-MMA piece of the 16-warp backward kernel (results are
wrong when a piece is skipped; only psi2_bwd time is read)."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rgp_b200.device import DevicePsi
dp = DevicePsi(0); dev = torch.device("cuda", 0)
dp.handle.set_option("bwd_warps", 16)   # the knobs live in the 16-warp kernel
N, M, Q = 1 << 18, 512, 64
g = torch.Generator(device=dev).manual_seed(1); f64 = dict(dtype=torch.float64, device=dev)
mu = torch.randn((N, Q), generator=g, **f64); S = torch.rand((N, Q), generator=g, **f64) * 0.49 + 0.01
Z = torch.randn((M, Q), generator=g, **f64); ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * 8.0
dL1 = torch.randn((N, M), generator=g, **f64) / M; dL2 = torch.randn((M, M), generator=g, **f64) / M ** 2
names = {0: "baseline", 1: "no exp", 2: "no lambda sums", 4: "no Wq reduce", 16: "no L store", 32: "no flushes",
         64: "no ZW build", 128: "no barrier", 1 | 2 | 4 | 32 | 64: "no exp/sums/Wq/flush/build",
         1 | 2 | 4 | 16 | 32 | 64 | 128: "MMA loops only", 256: "no stage 2-I loop", 512: "no stage 2-J loop",
         1024: "no stage 1 loop (off-diag)", 256 | 512: "no stage 2 loops", 256 | 512 | 1024: "no MMA loops (off-diag)",
         247 | 256 | 512: "only stage 1 loop", 247 | 512 | 1024: "only stage 2-I loop", 247 | 256 | 1024: "only stage 2-J loop"}
for mask, nm in names.items():
    dp.handle.set_option("debug_skip", mask)
    dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
    dp.handle.set_option("profile", 1); dp.handle.reset_counters()
    for _ in range(2): dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
    kt = dp.handle.kernel_times(); dp.handle.set_option("profile", 0)
    print(json.dumps({"mask": mask, "what": nm, "psi2_bwd_ms": kt["psi2_bwd"][0] / kt["psi2_bwd"][1]}), flush=True)
dp.handle.set_option("debug_skip", 0)
