"""Tiny runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tests/tools/sanitize_small.py
Covers the default fast path (software-pipelined backward kernel with TMA row-vector staging), the
row-at-a-time backward kernel, the fused pass, the small-inducing-set kernels, the lag-window kernels and the latent-terms kernel; results are checked against the oracle."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from synth import make_inputs, make_upstream, relerr
from oracle.psi_oracle import psi_forward, psi_backward
from rgp_b200.device import DevicePsi
from rgp_b200.lagwindow import LagWindow

t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
worst = 0.0
variants = [{"small_m": 0}, {"small_m": 0, "bwd_pipe": 0}, {"small_m": 0, "bwd_pipe": 1}]
shapes = [(37, 70, 20), (21, 130, 64), (9, 64, 33)]
small_shapes = [(37, 100, 20), (21, 50, 20), (19, 100, 40), (5, 33, 7), (300, 112, 23)]   # 16-warp, 8-warp, wide-Q, tiny, > 1 row per CTA
if os.environ.get("SANITIZE_ONLY_SMALL"):     # only the small-inducing-set kernels (psi2_small.cuh)
    variants = []
for opts in variants + [{"small_m": 1}, {"small_m": 1, "small_ks": 4}]:
    dp = DevicePsi(0, impl=1)
    for k, v in opts.items():
        dp.handle.set_option(k, v)
    for (N, M, Q) in (small_shapes if opts.get("small_m") == 1 else shapes):
        var, ell, Z, mu, S = make_inputs(N, M, Q, seed=4)
        dL0, dL1, dL2 = make_upstream(N, M)
        of = psi_forward(var, ell, Z, mu, S); ob = psi_backward(dL0, dL1, dL2, var, ell, Z, mu, S)
        _, p1, p2 = dp.forward(t(mu), t(S), t(Z), t(ell), var)
        out = dp.backward(t(mu), t(S), t(Z), t(ell), var, t(dL0), t(dL1), t(dL2))
        (q1, q2), fo = dp.fused(t(mu), t(S), t(Z), t(ell), var, t(dL0), t(dL1), t(dL2))
        errs = [relerr(p1.cpu().numpy(), of[1]), relerr(p2.cpu().numpy(), of[2]), relerr(q2.cpu().numpy(), of[2])]
        errs += [relerr(a.cpu().numpy(), b) for a, b in zip(out, ob)] + [relerr(a.cpu().numpy(), b) for a, b in zip(fo, ob)]
        worst = max(worst, max(errs))
        print(opts, (N, M, Q), "max rel err %.2e" % max(errs), flush=True)
if os.environ.get("SANITIZE_ONLY_SMALL"):
    assert worst < 1e-10, worst
    print("sanitize_small (small kernels only): all results match the oracle, worst rel err %.2e" % worst)
    sys.exit(0)
lw = LagWindow(dp.handle, (9, 6), 3, 2, (9, 7), 2, 3)
lat, ctl = torch.randn((15, 2), dtype=torch.float64, device="cuda"), torch.randn((16, 3), dtype=torch.float64, device="cuda")
X = lw.gather(lat, ctl)
lw.scatter_add(X)
gm, gv, val = lw.latent_terms(lat, lat.abs() + 0.1, torch.randn((lw.N, 2), dtype=torch.float64, device="cuda"),
                              torch.randn(lw.N, dtype=torch.float64, device="cuda"))
torch.cuda.synchronize()
print("lag / latent kernels ok", float(val))
# MLP back-constraint kernels (shared-memory input window, pending-gradient ring)
from oracle import mlp_oracle as mo
from rgp_b200.backconstraint import MLPBackConstraint
from test_mlp_oracle import make_case
c = make_case(seed=2, X_win=3, X_dim=2, U_win=2, U_dim=1, n_steps=(12, 9), control=True)
lw2 = LagWindow(dp.handle, [3 + N for N in c["n_steps"]], 3, 2, [u.shape[0] for u in c["ctl"]], 2, 1)
enc = MLPBackConstraint(lw2)
with torch.no_grad():
    enc.flat.copy_(t(np.concatenate([np.concatenate([W.ravel(), b]) for W, b in c["params"]])))
init = t(np.stack(c["init"])).requires_grad_(True)
ctl2 = t(np.vstack(c["ctl"])).requires_grad_(True)
lat2 = enc(init, ctl2)
(lat2 * t(np.vstack(c["weights"]))).sum().backward()
Xo = mo.freerun(c["params"], c["init"], c["ctl"], c["n_steps"], 3, 2)
go = [w.copy() for w in c["weights"]]
pgo, cgo = mo.freerun_backward(c["params"], Xo, c["ctl"], go, 3, 2)
e_mlp = max(relerr(lat2.detach().cpu().numpy(), np.vstack(Xo)),
            relerr(enc.flat.grad.cpu().numpy(), np.concatenate([np.concatenate([dW.ravel(), db]) for dW, db in pgo])),
            relerr(init.grad.cpu().numpy(), np.stack([x[:3] for x in go])), relerr(ctl2.grad.cpu().numpy(), np.vstack(cgo)))
print("mlp back-constraint kernels: max rel err %.2e" % e_mlp)
worst = max(worst, e_mlp)
assert worst < 1e-10, worst
print("sanitize_small: all results match the oracle, worst rel err %.2e" % worst)
