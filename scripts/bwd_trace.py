"""Timeline trace of the 16-warp backward kernel: clock64 per warp at phase boundaries for 16
rows of one off-diagonal block in CTA 0.  Columns: t0 iteration start, t1 after first phase,
t2 after second phase, t3 after tile build, t4 after the barrier."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rgp_b200.device import DevicePsi
dp = DevicePsi(0); dev = torch.device("cuda", 0)
dp.handle.set_option("bwd_warps", 16)   # these experiments instrument the 16-warp kernel
N, M, Q = 1 << 16, 512, 64
g = torch.Generator(device=dev).manual_seed(1); f64 = dict(dtype=torch.float64, device=dev)
mu = torch.randn((N, Q), generator=g, **f64); S = torch.rand((N, Q), generator=g, **f64) * 0.49 + 0.01
Z = torch.randn((M, Q), generator=g, **f64); ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * 8.0
dL1 = torch.randn((N, M), generator=g, **f64) / M; dL2 = torch.randn((M, M), generator=g, **f64) / M ** 2
trace = torch.zeros(16 * 16 * 8, dtype=torch.int64, device=dev)
dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
dp.handle.set_option("trace_ptr", trace.data_ptr())
dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
torch.cuda.synchronize()
dp.handle.set_option("trace_ptr", 0)
t = trace.cpu().numpy().reshape(16, 16, 8)
base = t[0, :, 0].min()
for row in (4, 5):
    print("row", row)
    for w in range(16):
        grp = "B" if (w >> 2) & 1 else "A"
        e = t[row, w, :5] - base
        print(f"  warp {w:2d} grp {grp} start {e[0]:7d}  p1 +{e[1]-e[0]:6d}  p2 +{e[2]-e[1]:6d}  p3 +{e[3]-e[2]:5d}  bar-issue +{e[4]-e[3]:6d}  end {e[3]:7d}")
it = t[1:, 0, 0] - t[:-1, 0, 0]
print("iteration lengths (warp 0):", it.tolist())
