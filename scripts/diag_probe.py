"""Per-block cost of diagonal vs off-diagonal blocks: M=64 is one diagonal block, M=128 is two
diagonal + one off-diagonal.  Pure DMMA time per row: diag fwd 36 tiles, bwd 36+64; off-diag fwd 64, bwd 192
(x 16 DMMA x 16 cycles / 4 schedulers)."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rgp_b200.device import DevicePsi
dp = DevicePsi(0); dev = torch.device("cuda", 0)
dp.handle.set_option("bwd_warps", 16)   # these experiments instrument the 16-warp kernel
N, Q = 1 << 20, 64
res = {}
for M in (64, 128):
    g = torch.Generator(device=dev).manual_seed(1); f64 = dict(dtype=torch.float64, device=dev)
    mu = torch.randn((N, Q), generator=g, **f64); S = torch.rand((N, Q), generator=g, **f64) * 0.49 + 0.01
    Z = torch.randn((M, Q), generator=g, **f64); ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * 8.0
    dL1 = torch.randn((N, M), generator=g, **f64) / M; dL2 = torch.randn((M, M), generator=g, **f64) / M ** 2
    for _ in range(2):
        dp.forward(mu, S, Z, ell, 1.3, want_psi1=False); dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
    dp.handle.set_option("profile", 1); dp.handle.reset_counters()
    for _ in range(3):
        dp.forward(mu, S, Z, ell, 1.3, want_psi1=False); dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
    kt = dp.handle.kernel_times(); dp.handle.set_option("profile", 0)
    res[M] = {k: kt[k][0] / kt[k][1] for k in ("psi2_fwd", "psi2_bwd")}
cyc = lambda ms: ms * 1e-3 * 1.965e9 * 148 / N          # SM-cycles per row (all SMs busy)
out = {}
for k, pure_d, pure_o in (("psi2_fwd", 36 * 64, 64 * 64), ("psi2_bwd", 100 * 64, 192 * 64)):
    D = cyc(res[64][k]); O = cyc(res[128][k]) - 2 * D
    out[k] = {"diag_cycles_per_row": D, "diag_pure_dmma": pure_d, "diag_eff": pure_d / D,
              "offdiag_cycles_per_row": O, "offdiag_pure_dmma": pure_o, "offdiag_eff": pure_o / O}
print(json.dumps({"ms": res, "per_block": out}))
