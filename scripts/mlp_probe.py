"""Timing experiment: free-run step latency with the GPU clocks kept up by a background GEMM stream, and
which part of a step costs the time (debug_skip masks: results wrong when a part is skipped)."""
import json, os, subprocess, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rgp_b200._lib import Handle
from rgp_b200.backconstraint import MLPBackConstraint
from rgp_b200.lagwindow import LagWindow
h = Handle(0)
lw = LagWindow(h, [10 + 502], 10, 1, [502 + 9], 10, 1)
enc = MLPBackConstraint(lw)
init = torch.randn((1, 10, 1), dtype=torch.float64, device="cuda")
ctl = torch.randn((lw.ctl_total, 1), dtype=torch.float64, device="cuda")
A = torch.randn((8192, 8192), device="cuda", dtype=torch.bfloat16)
side = torch.cuda.Stream()
def clocks():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-i", "0"],
                          capture_output=True, text=True).stdout.strip()
for busy in (False, True):
    for mask, what in ((0, "baseline"), (1, "no mat-vec"), (15, "barriers + window shift only")):
        h.set_option("debug_skip", mask)
        with torch.no_grad():
            if busy:
                for _ in range(200):
                    A @ A
                torch.cuda.synchronize()
            enc(init, ctl); torch.cuda.synchronize()
            h.set_option("profile", 1); h.reset_counters()
            for _ in range(5):
                if busy:
                    with torch.cuda.stream(side):
                        for _ in range(3):
                            A @ A
                enc(init, ctl)
            clk = clocks()
            kt = h.kernel_times(); h.set_option("profile", 0)
        print(json.dumps({"gpu_kept_busy": busy, "sm_mhz_during": clk, "mask": mask, "what": what,
                          "us_per_step": 1e3 * kt["mlp_freerun"][0] / kt["mlp_freerun"][1] / 502}), flush=True)
h.set_option("debug_skip", 0)
