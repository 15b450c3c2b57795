"""Shape sweep (configs[4] of BASELINE.json): forward+backward throughput and fraction of
the measured fp64 peak per (N, M, Q), device-resident inputs, CUDA-event timing.
Also the small-N system-ID shapes (configs 1-3), reported as latency per evaluation."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgp_b200.device import DevicePsi  # noqa: E402


def f_row(M, Q):
    P = M * (M + 1) // 2
    return 4 * P * Q + 2 * M * M * Q + 16 * M * Q + 16 * P


def run(dp, N, M, Q, reps=3):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    f64 = dict(dtype=torch.float64, device=dev)
    mu = torch.randn((N, Q), generator=g, **f64)
    S = torch.rand((N, Q), generator=g, **f64) * 0.49 + 0.01
    Z = torch.randn((M, Q), generator=g, **f64)
    ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * Q ** 0.5
    dL1 = torch.randn((N, M), generator=g, **f64) / M
    dL2 = torch.randn((M, M), generator=g, **f64) / (M * M)
    psi1 = torch.empty((N, M), **f64)

    def step():
        dp.forward(mu, S, Z, ell, 1.3, psi1_out=psi1)
        dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)

    step(); step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dp = DevicePsi(0)
    peak = dp.handle.fp64_peak(reps=5)
    print(json.dumps({"fp64_peak_tflops": peak}), flush=True)
    big = [(1 << 20, 128, 16), (1 << 20, 256, 32), (1 << 19, 512, 32), (1 << 19, 512, 64), (1 << 17, 1024, 64),
           (1 << 21, 64, 64), (1 << 20, 128, 64), (1 << 16, 512, 64), (1 << 20, 100, 20), (1 << 19, 200, 40),
           (1 << 20, 50, 20), (1 << 18, 500, 60)]
    for N, M, Q in big:
        ms = run(dp, N, M, Q)
        rps = N / (ms * 1e-3)
        print(json.dumps({"N": N, "M": M, "Q": Q, "ms": ms, "rows_per_s": rps, "tflops": rps * f_row(M, Q) / 1e12,
                          "frac_of_fp64_peak": rps * f_row(M, Q) / 1e12 / peak}), flush=True)
    # configs 1-3: one layer evaluation at the real shapes (latency-bound)
    for name, N, M, Q in [("actuator_hidden", 502, 100, 20), ("actuator_output", 502, 100, 10),
                          ("ballbeam", 490, 50, 20), ("mocap", 408, 200, 40)]:
        ms = run(dp, N, M, Q, reps=20)
        print(json.dumps({"config": name, "N": N, "M": M, "Q": Q, "ms_per_eval": ms, "rows_per_s": N / (ms * 1e-3)}),
              flush=True)


if __name__ == "__main__":
    main()
