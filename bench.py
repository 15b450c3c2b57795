#!/usr/bin/env python
"""bench.py - psi-statistics + gradients throughput (rows/s, fp64) on B200.

A "step" is one pass of the hot path over one batch of synthetic rows: forward
(Psi0, Psi1, Psi2) followed by backward (all five gradient blocks) with upstream
gradients supplied, at the headline shape of BASELINE.json: N = 4*2^20 rows per GPU,
M = 512, Q = 64, fp64.  Weak scaling: every rank owns N rows; after each phase the two
packed all-reduces of SURVEY.md 8(e) run inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA)
  python bench.py --impl reference ...                         reference arm: the
      reference's CPU implementation of the path (GPy's closed forms restated in numpy,
      oracle/psi_oracle.py - GPy itself is not installable here) on the host cores.

Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "psi-stat+grad rows/sec (fp64, N x M x Q)"
UNIT = "rows/s"


# ----------------------------------------------------------------- algorithmic work
def flops_row_fwd_psi2(M, Q):
    P = M * (M + 1) // 2
    return 2 * P * Q + 8 * P


def flops_row_bwd_psi2(M, Q):
    P = M * (M + 1) // 2
    return 2 * P * Q + 2 * M * M * Q + 8 * P


def flops_row_total(M, Q):
    """F_row of SURVEY.md 8(d) / BASELINE.md section 3."""
    P = M * (M + 1) // 2
    return 4 * P * Q + 2 * M * M * Q + 16 * M * Q + 16 * P


def bytes_row(M, Q):
    return 8 * (6 * Q + 2 * M + 1)


# ------------------------------------------------------------------------ CPU arm
def cpu_path_rows_per_s(M, Q, target_s, rows0=64, seed=20240607):
    """Time the oracle (GPy-structured numpy: chunk x M x M materialisation + GEMMs) on a
    bounded sample of the workload.  Returns (rows/s, rows, seconds, cores)."""
    import numpy as np
    from oracle.psi_oracle import psi_backward, psi_forward
    from synth import make_inputs, make_upstream
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()

    def run(rows):
        var, ell, Z, mu, S = make_inputs(rows, M, Q, seed=seed)
        dL0, dL1, dL2 = make_upstream(rows, M)
        t0 = time.perf_counter()
        psi_forward(var, ell, Z, mu, S, budget_bytes=1 << 30)
        psi_backward(dL0, dL1, dL2, var, ell, Z, mu, S, budget_bytes=1 << 30)
        return time.perf_counter() - t0

    run(min(rows0, 16))                       # warm BLAS threads
    rows = rows0
    t = run(rows)
    while t < target_s / 3 and rows < (1 << 20):
        rows = int(min(1 << 20, max(rows * 2, rows * target_s / max(t, 1e-3) * 0.8)))
        t = run(rows)
    return rows / t, rows, t, cores


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    M, Q = args.M, args.Q
    # size one step at ~3 s of CPU work, then time warmup + K steps of that sample
    rps, rows, t, cores = cpu_path_rows_per_s(M, Q, target_s=3.0)
    import numpy as np
    from oracle.psi_oracle import psi_backward, psi_forward
    from synth import make_inputs, make_upstream
    var, ell, Z, mu, S = make_inputs(rows, M, Q)
    dL0, dL1, dL2 = make_upstream(rows, M)

    def step():
        psi_forward(var, ell, Z, mu, S, budget_bytes=1 << 30)
        psi_backward(dL0, dL1, dL2, var, ell, Z, mu, S, budget_bytes=1 << 30)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = rows * args.steps / dt
    sample = "%d rows/step of the N=%d, M=%d, Q=%d workload" % (rows, args.rows, M, Q)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "psi0/1/2 + all gradients, N=%d M=%d Q=%d fp64" % (args.rows, M, Q),
                   "N_per_gpu": args.rows, "M": M, "Q": Q,
                   "note": "GPy closed forms restated in numpy (oracle/psi_oracle.py); GPy not installable"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------- clock sampling
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_id):
        self.gpu_id = gpu_id
        self.proc = None
        self.path = "/tmp/rgp_clocks_%d_%d.csv" % (os.getpid(), int(time.time()))

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_id), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); pw.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v == "Active":
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------ our arm
def ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from rgp_b200.sharded import ShardedPsi

    N, M, Q = args.rows, args.M, args.Q
    # synthetic inputs of SURVEY.md 8(d), generated on the device (17 GB of dL_dpsi1)
    g = torch.Generator(device=dev).manual_seed(20240607 + rank)
    f64 = dict(dtype=torch.float64, device=dev)
    mu = torch.randn((N, Q), generator=g, **f64)
    S = torch.rand((N, Q), generator=g, **f64) * 0.49 + 0.01
    gz = torch.Generator(device=dev).manual_seed(20240607)          # Z, ell replicated on every rank
    Z = torch.randn((M, Q), generator=gz, **f64)
    ell = (torch.rand(Q, generator=gz, **f64) * 0.7 + 0.7) * Q ** 0.5
    variance = 1.3
    dL1 = torch.randn((N, M), generator=g, **f64) / M
    dL2 = torch.randn((M, M), generator=gz, **f64) / (M * M)
    dL2 = 0.5 * (dL2 + dL2.T)
    psi1 = torch.empty((N, M), **f64)
    dmu = torch.empty((N, Q), **f64)
    dS = torch.empty((N, Q), **f64)

    sp = ShardedPsi(local, impl=args.kernels)
    h = sp.psi.handle

    def step():
        _, p1, p2 = sp.psi.forward(mu, S, Z, ell, variance, psi1_out=psi1)
        p0 = torch.full((1,), variance * N, **f64)
        if world > 1:
            from rgp_b200.sharded import reduce_forward, reduce_backward
            p0, p2, _ = reduce_forward(p0, p2)
        out = sp.psi.backward(mu, S, Z, ell, variance, -0.5, dL1, dL2, dmu_out=dmu, dS_out=dS)
        if world > 1:
            out = reduce_backward(out[0], out[1], out[2]) + out[3:]
        return p2, out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_tf = h.fp64_peak(reps=5)                   # fp64 roofline denominator, measured in-run
    for _ in range(args.warmup):
        step()
    barrier()
    h.set_option("profile", 1)
    h.reset_counters()
    sampler = ClockSampler(torch.cuda.get_device_properties(dev).uuid if False else local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        p2, out = step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = h.launch_count()
    ktimes = h.kernel_times()
    h.set_option("profile", 0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * N * args.steps / (ms * 1e-3)
    checksum = float(p2.sum().item()) + float(out[2].sum().item())

    # Informational, not the headline: the same rows through the FUSED entry point (statistics and
    # gradients from one pass; valid when the upstream gradients do not depend on the statistics, i.e.
    # the SVI bound - DESIGN.md section 9).  One warm pass, one timed pass, max over ranks.
    fused = None
    if not args.no_fused:
        def fstep():
            (q1, q2), fo = sp.psi.fused(mu, S, Z, ell, variance, -0.5, dL1, dL2, psi1_out=psi1, dmu_out=dmu, dS_out=dS)
            if world > 1:
                from rgp_b200.sharded import reduce_forward, reduce_backward
                _, q2, _ = reduce_forward(torch.full((1,), variance * N, **f64), q2)
                fo = reduce_backward(fo[0], fo[1], fo[2]) + fo[3:]
            return q2, fo
        q2, fo = fstep()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        q2, fo = fstep()
        f1.record()
        barrier()
        tf = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        fms = float(tf.item())
        fused = {"value": world * N / (fms * 1e-3), "unit": UNIT, "ms_per_step": fms,
                 "max_rel_diff_psi2_vs_two_phase": float((q2 - p2).abs().max() / p2.abs().max()),
                 "max_rel_diff_dZ_vs_two_phase": float((fo[2] - out[2]).abs().max() / out[2].abs().max()),
                 "note": "rgp_psi_fused_dev: one pass for statistics + gradients (SVI bound); not the headline metric"}

    # dominant kernel -> roofline (fp64 CUDA-core pipe; see DESIGN.md "Measurement")
    roof = None
    if ktimes:
        name, (tot_ms, cnt) = max(ktimes.items(), key=lambda kv: kv[1][0])
        per_launch_ms = tot_ms / max(cnt, 1)
        rows_per_launch = N * args.steps / max(cnt, 1)
        if "bwd" in name:
            fl = flops_row_bwd_psi2(M, Q)
        elif "psi2" in name:
            fl = flops_row_fwd_psi2(M, Q)
        else:
            fl = None
        if fl is not None:
            ach = fl * rows_per_launch / (per_launch_ms * 1e-3) / 1e12
            traffic = None
            try:                # dram bytes/row of this kernel from the committed ncu capture, scaled to this launch
                with open(os.path.join(ROOT, "profiles", "ncu_traffic_r01.json")) as f:
                    tr = json.load(f).get(name)
                if tr and (M, Q) == (512, 64):
                    traffic = tr["bytes_per_row"] * rows_per_launch
            except Exception:
                traffic = None
            roof = {"bound": "fp64", "kernel": name, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": ach / peak_tf if peak_tf else None, "traffic": traffic,
                    "traffic_note": "dram bytes/launch = ncu dram__bytes per row at a 65536-row capture "
                                    "(profiles/ncu_traffic_r01.json) x rows per launch",
                    "peak_source": "DFMA-chain microbenchmark run in this process (rgp_psi_fp64_peak); "
                                   "MEASURED_PEAKS.json has no fp64 entry",
                    "launch_ms": per_launch_ms, "launches": cnt,
                    "flops_per_row": fl, "rows_per_launch": rows_per_launch}
    hbm_peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak = json.load(f).get("hbm_gbs")
    except Exception:
        pass
    hbm_ach = value / world * bytes_row(M, Q) / 1e9
    whole = {"achieved_tflops": value / world * flops_row_total(M, Q) / 1e12, "peak_tflops": peak_tf,
             "frac": value / world * flops_row_total(M, Q) / 1e12 / peak_tf if peak_tf else None,
             "flops_per_row": flops_row_total(M, Q), "bytes_per_row": bytes_row(M, Q),
             "hbm_gbs_algorithmic": hbm_ach, "hbm_peak_gbs": hbm_peak if hbm_peak else 6650.0,
             "hbm_peak_source": "MEASURED_PEAKS.json" if hbm_peak else "fallback (B200_PROFILING.md)",
             "hbm_frac": hbm_ach / (hbm_peak if hbm_peak else 6650.0)}
    kshare = {k: {"ms": round(v[0], 3), "launches": v[1]} for k, v in
              sorted(ktimes.items(), key=lambda kv: -kv[1][0])}

    # end-to-end through the plugin's host-buffer C-ABI calls, pinned host memory
    Ne = min(N, args.e2e_rows)
    pin = lambda *shape: torch.empty(shape, dtype=torch.float64, pin_memory=True)
    h_mu, h_S = pin(Ne, Q), pin(Ne, Q)
    h_mu.copy_(mu[:Ne]); h_S.copy_(S[:Ne])
    h_Z, h_ell, h_dL2 = pin(M, Q), pin(Q), pin(M, M)
    h_Z.copy_(Z); h_ell.copy_(ell); h_dL2.copy_(dL2)
    h_dL1 = pin(Ne, M); h_dL1.copy_(dL1[:Ne])
    h_p1, h_p2 = pin(Ne, M), pin(M, M)
    h_dmu, h_dS, h_dZ, h_dl, h_dv = pin(Ne, Q), pin(Ne, Q), pin(M, Q), pin(Q), pin(1)
    torch.cuda.synchronize()
    P = lambda t_: t_.data_ptr()

    def e2e_step():
        h.forward_host(Ne, M, Q, P(h_mu), P(h_S), P(h_Z), P(h_ell), variance, None, P(h_p1), P(h_p2))
        h.backward_host(Ne, M, Q, P(h_mu), P(h_S), P(h_Z), P(h_ell), variance, None, -0.5, P(h_dL1),
                        P(h_dL2), P(h_dmu), P(h_dS), P(h_dZ), P(h_dl), P(h_dv))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    te = time.perf_counter() - t0
    t = torch.tensor([te], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    te = float(t.item())
    h2d = 8 * (2 * 2 * Ne * Q + 2 * M * Q + 2 * Q + M * M + Ne * M)
    d2h = 8 * (Ne * M + M * M + 2 * Ne * Q + M * Q + Q + 1)
    e2e = {"value": world * Ne * args.e2e_steps / te, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "rows_per_step": Ne, "steps": args.e2e_steps,
           "api": "rgp_psi_forward_host + rgp_psi_backward_host (pinned host buffers)"}

    cpu = None
    if rank == 0 and world == 1 and args.cpu_seconds > 0:
        rps, rows, secs, cores = cpu_path_rows_per_s(M, Q, target_s=args.cpu_seconds)
        cpu = {"value": rps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d rows of the same workload in %.1f s (numpy restatement of GPy's closed "
                         "forms, oracle/psi_oracle.py)" % (rows, secs)}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "psi0/1/2 forward + all gradients, N=%d rows/GPU, M=%d, Q=%d, fp64"
                                   % (N, M, Q), "N_per_gpu": N, "M": M, "Q": Q,
                       "parallelism": "rows sharded x%d, 2 packed all-reduces/step" % world,
                       "l2": "inputs (%.1f GB/GPU) exceed the 126 MB L2; no flush needed"
                             % ((2 * N * Q + 2 * N * M) * 8 / 1e9),
                       "kernels": {0: "auto", 1: "fast", 2: "reference"}[args.kernels]},
            "roofline": roof, "whole_step": whole, "kernel_ms": kshare,
            "cpu_baseline": cpu, "e2e": e2e, "fused_svi_pass": fused, "gpu_launches": launches, "clocks": clocks,
            "checksum": checksum,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=4 * 2 ** 20, help="rows per GPU (headline 4*2^20)")
    ap.add_argument("--M", type=int, default=512)
    ap.add_argument("--Q", type=int, default=64)
    ap.add_argument("--kernels", type=int, default=0, help="0 auto, 1 fast, 2 reference kernels")
    ap.add_argument("--e2e-rows", type=int, default=2 ** 20)
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-fused", action="store_true", help="skip the informational fused-pass measurement")
    args = ap.parse_args()
    if args.warmup < 3:
        print("warning: contract asks for >= 3 warm-up steps", file=sys.stderr)
    return reference_arm(args) if args.impl == "reference" else ours(args)


if __name__ == "__main__":
    sys.exit(main())
