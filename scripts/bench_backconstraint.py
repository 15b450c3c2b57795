"""Latency of the MLP back-constraint kernels (free-run recurrence + back-propagation through it):
one launch each, one CTA per sequence.  Shapes: config 1 (Actuator: 1 sequence, 502 steps, Q = 20,
default widths [20, 40, 20, 1]) and a batch of 64 sequences."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rgp_b200._lib import Handle
from rgp_b200.backconstraint import MLPBackConstraint
from rgp_b200.lagwindow import LagWindow

h = Handle(0)
for name, nseq, T, Xw, Dx, Uw, Du in (("config1_actuator", 1, 502, 10, 1, 10, 1), ("batch64", 64, 512, 10, 2, 10, 2),
                                      ("mocap_like", 8, 51, 20, 1, 20, 1)):
    lw = LagWindow(h, [Xw + T] * nseq, Xw, Dx, [T + Uw - 1] * nseq, Uw, Du)
    enc = MLPBackConstraint(lw)
    init = torch.randn((nseq, Xw, Dx), dtype=torch.float64, device="cuda", requires_grad=True)
    ctl = torch.randn((lw.ctl_total, Du), dtype=torch.float64, device="cuda")
    w = torch.randn((lw.lat_total, Dx), dtype=torch.float64, device="cuda")
    def step():
        enc.zero_grad(); init.grad = None
        lat = enc(init, ctl)
        (lat * w).sum().backward()
    step(); torch.cuda.synchronize()
    h.set_option("profile", 1); h.reset_counters()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record(); torch.cuda.synchronize()
    kt = h.kernel_times(); h.set_option("profile", 0)
    print(json.dumps({"row": "mlp_back_constraint", "config": name, "sequences": nseq, "steps": T, "units": enc.units,
                      "parameters": int(enc.flat.numel()),
                      "freerun_kernel_ms": kt["mlp_freerun"][0] / kt["mlp_freerun"][1],
                      "backprop_kernel_ms": kt["mlp_freerun_bwd"][0] / kt["mlp_freerun_bwd"][1],
                      "fwd_plus_bwd_wall_ms": e0.elapsed_time(e1) / reps,
                      "us_per_step_fwd": 1e3 * kt["mlp_freerun"][0] / kt["mlp_freerun"][1] / T}), flush=True)
