#!/usr/bin/env python
"""bench.py - psi-statistics + gradients throughput (rows/s, fp64) on B200.

A "step" is one pass of the hot path over one batch of synthetic rows: forward
(Psi0, Psi1, Psi2) followed by backward (all five gradient blocks) with upstream
gradients supplied, at the headline shape of BASELINE.json: N = 4*2^20 rows per GPU,
M = 512, Q = 64, fp64.  Default scaling is weak (every rank owns N rows); after each phase the
two packed all-reduces of SURVEY.md 8(e) run inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA)
  python bench.py --scaling strong ...                         N rows in TOTAL, split over the ranks
  python bench.py --workload svi10m ...                        BASELINE.json config 4 (see svi_workload)
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


def workload_config(N, M, Q):
    """The same dict in both arms (the driver compares them)."""
    return {"workload": "psi0/1/2 forward + all gradients, N=%d rows/GPU, M=%d, Q=%d, fp64" % (N, M, Q),
            "N_per_gpu": N, "M": M, "Q": Q}


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


def bytes_row_bwd_psi2(M, Q):
    """Algorithmic HBM bytes per row of the Psi2 backward kernel: reads ws[Q] and H[M], writes lambda[M]
    and W[Q] once."""
    return 8 * (2 * Q + 2 * M)


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# ------------------------------------------------------------------------ CPU arm
def cpu_path_rows_per_s(M, Q, target_s, rows0=64, seed=20240607):
    """Time the oracle (GPy-structured numpy: chunk x M x M materialisation + GEMMs) on a
    bounded sample of the workload.  Returns (rows/s, rows, seconds, cores)."""
    from oracle.psi_oracle import psi_backward, psi_forward
    from synth import make_inputs, make_upstream
    cores = host_cores()

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
    # a FIXED sample per step (so the number does not depend on a sizing heuristic that reacts to box noise)
    rows = args.ref_rows
    from oracle.psi_oracle import psi_backward, psi_forward
    from synth import make_inputs, make_upstream
    cores = host_cores()
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
    cfg = workload_config(args.rows, M, Q)
    cfg["note"] = "GPy closed forms restated in numpy (oracle/psi_oracle.py) on the host cores; GPy not installable"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg,
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


# ----------------------------------------------------------------- multi-rank parity
def parity_check(dp, world, rank, dev):
    """Before anything is timed: (i) the row-sharded evaluation over all ranks against the same rows
    evaluated on ONE GPU (every rank evaluates the full small problem itself), (ii) the tiled kernels
    against the independent one-thread-per-output kernel family on a sub-range.  Headline tile shape
    (M = 512, Q = 64), ragged row counts.  Returns max relative difference over all ranks."""
    import torch
    import torch.distributed as dist
    from rgp_b200.device import DevicePsi
    from rgp_b200.sharded import reduce_backward, reduce_forward, row_partition
    M, Q = 512, 64
    n_total = 1500 * world + 37
    g = torch.Generator(device=dev).manual_seed(4242)            # identical on every rank
    f64 = dict(dtype=torch.float64, device=dev)
    mu = torch.randn((n_total, Q), generator=g, **f64)
    S = torch.rand((n_total, Q), generator=g, **f64) * 0.49 + 0.01
    Z = torch.randn((M, Q), generator=g, **f64)
    ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * Q ** 0.5
    dL1 = torch.randn((n_total, M), generator=g, **f64) / M
    dL2 = torch.randn((M, M), generator=g, **f64) / (M * M)
    var = 1.3
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    _, p1, p2 = dp.forward(mu, S, Z, ell, var)
    full = dp.backward(mu, S, Z, ell, var, -0.5, dL1, dL2)
    p1, p2, full = p1.clone(), p2.clone(), [t.clone() for t in full]
    worst = 0.0
    if world > 1:
        s, e = row_partition(n_total, world, rank)
        _, q1, q2 = dp.forward(mu[s:e], S[s:e], Z, ell, var)
        _, q2, _ = reduce_forward(torch.zeros(1, **f64), q2.clone())
        out = dp.backward(mu[s:e], S[s:e], Z, ell, var, -0.5, dL1[s:e].contiguous(), dL2)
        rv, rl, rz = reduce_backward(out[0].clone(), out[1].clone(), out[2].clone())
        # dvar carries N * dL_dpsi0_const: the shards' constants add up to the full one
        worst = max(rel(q2, p2), rel(q1, p1[s:e]), rel(rv, full[0]), rel(rl, full[1]), rel(rz, full[2]),
                    rel(out[3], full[3][s:e]), rel(out[4], full[4][s:e]))
    ref = DevicePsi(dev.index, impl=2)
    k = 1024
    _, r1, r2 = ref.forward(mu[:k], S[:k], Z, ell, var)
    rb = ref.backward(mu[:k], S[:k], Z, ell, var, -0.5, dL1[:k].contiguous(), dL2)
    _, f1, f2 = dp.forward(mu[:k], S[:k], Z, ell, var)
    fb = dp.backward(mu[:k], S[:k], Z, ell, var, -0.5, dL1[:k].contiguous(), dL2)
    kern = max([rel(f1, r1), rel(f2, r2)] + [rel(a, b) for a, b in zip(fb, rb)])
    t = torch.tensor([worst, kern], **f64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"ranks": world, "max_rel": max(float(t[0]), float(t[1])),
            "sharded_vs_single_gpu": float(t[0]) if world > 1 else None,
            "tiled_vs_reference_kernels": float(t[1]),
            "what": "N=%d rows (ragged shards), M=512, Q=64: all-reduced Psi2/dZ/dl/dvar and per-rank rows of "
                    "Psi1/dmu/dS vs the same rows on one GPU; tiled kernels vs the independent kernel family on "
                    "%d rows" % (n_total, k)}


# ------------------------------------------------------------------------ our arm
def model_shapes(local, peak_tf, rows=1 << 20):
    """rows/s of forward + backward at the shapes of the reference's models (one GPU, inputs resident, CUDA events,
    1 warm + 2 timed steps).  Never fails the bench: an error is reported in place of the numbers."""
    import torch
    from rgp_b200.device import DevicePsi
    out = []
    try:
        dev = torch.device("cuda", local)
        dp = DevicePsi(local)
        f64 = dict(dtype=torch.float64, device=dev)
        for M, Q in ((100, 20), (100, 40), (50, 20)):
            g = torch.Generator(device=dev).manual_seed(7)
            mu = torch.randn((rows, Q), generator=g, **f64)
            S = torch.rand((rows, Q), generator=g, **f64) * 0.49 + 0.01
            Z = torch.randn((M, Q), generator=g, **f64)
            ell = (torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * Q ** 0.5
            dL1 = torch.randn((rows, M), generator=g, **f64) / M
            dL2 = torch.randn((M, M), generator=g, **f64) / (M * M)
            psi1 = torch.empty((rows, M), **f64)
            dmu, dS = torch.empty((rows, Q), **f64), torch.empty((rows, Q), **f64)

            def step():
                dp.forward(mu, S, Z, ell, 1.3, psi1_out=psi1)
                dp.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2, dmu_out=dmu, dS_out=dS)
            step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 2
            rps = rows / (ms * 1e-3)
            out.append({"M": M, "Q": Q, "rows": rows, "ms_per_step": ms, "rows_per_s": rps,
                        "frac_of_fp64_peak": rps * flops_row_total(M, Q) / 1e12 / peak_tf if peak_tf else None})
            del mu, S, dL1, psi1, dmu, dS
        torch.cuda.empty_cache()
    except Exception as e:      # informational only
        out.append({"error": repr(e)})
    return out


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
    if args.workload == "svi10m":
        from svi_workload import run_svi10m
        return run_svi10m(args, world, rank, local, dev)
    from rgp_b200.sharded import ShardedPsi, reduce_backward, reduce_forward, row_partition

    M, Q = args.M, args.Q
    strong = args.scaling == "strong"
    if strong:
        s0, s1 = row_partition(args.rows, world, rank)
        N = s1 - s0
        N_total = args.rows
    else:
        N = args.rows
        N_total = N * world
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
    parity = parity_check(sp.psi, world, rank, dev)

    def make_step(n):
        """forward + backward over the first n local rows (views of the resident buffers)."""
        m_, s_, d1_, p1_, gm_, gs_ = mu[:n], S[:n], dL1[:n], psi1[:n], dmu[:n], dS[:n]

        def step():
            _, p1, p2 = sp.psi.forward(m_, s_, Z, ell, variance, psi1_out=p1_)
            p0 = torch.full((1,), variance * n, **f64)
            if world > 1:
                p0, p2, _ = reduce_forward(p0, p2)
            out = sp.psi.backward(m_, s_, Z, ell, variance, -0.5, d1_, dL2, dmu_out=gm_, dS_out=gs_)
            if world > 1:
                out = reduce_backward(out[0], out[1], out[2]) + out[3:]
            return p2, out
        return step

    step = make_step(N)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            res = fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), res

    peak_tf = h.fp64_peak(reps=5)                   # fp64 roofline denominator, measured in-run
    for _ in range(args.warmup):
        step()
    barrier()
    h.set_option("profile", 1)
    h.reset_counters()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, (p2, out) = timed(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = h.launch_count()
    ktimes = h.kernel_times()
    h.set_option("profile", 0)
    value = N_total * args.steps / (ms * 1e-3)
    checksum = float(p2.sum().item()) + float(out[2].sum().item())

    # The other scaling mode, informational (the headline `value` is the mode --scaling names): with
    # weak scaling as the headline, also time the headline N split over the ranks (north star: "N = 4M ...
    # >= 85 % scaling efficiency at 8 GPUs").  Efficiency is quoted against this run's own per-GPU rate.
    other = None
    if world > 1 and not strong and not args.no_strong:
        a, b = row_partition(args.rows, world, rank)
        sstep = make_step(b - a)
        for _ in range(2):
            sstep()
        sms, _ = timed(sstep, args.steps)
        sval = args.rows * args.steps / (sms * 1e-3)
        other = {"scaling": "strong", "N_total": args.rows, "rows_per_gpu": b - a, "value": sval, "unit": UNIT,
                 "ms_per_step": sms / args.steps, "steps": args.steps,
                 "efficiency_vs_this_runs_per_gpu_rate": sval / value,
                 "note": "same kernels, the headline N split over the ranks; 2 warm-up steps; max over ranks"}

    # Informational, not the headline: the same rows through the FUSED entry point (statistics and
    # gradients from one pass; valid when the upstream gradients do not depend on the statistics, i.e.
    # the SVI bound - DESIGN.md section 9).  One warm pass, one timed pass, max over ranks.
    fused = None
    if not args.no_fused:
        def fstep():
            (q1, q2), fo = sp.psi.fused(mu, S, Z, ell, variance, -0.5, dL1, dL2, psi1_out=psi1, dmu_out=dmu, dS_out=dS)
            if world > 1:
                _, q2, _ = reduce_forward(torch.full((1,), variance * N, **f64), q2)
                fo = reduce_backward(fo[0], fo[1], fo[2]) + fo[3:]
            return q2, fo
        fstep()
        fms, (q2, fo) = timed(fstep, 1)
        fused = {"value": N_total / (fms * 1e-3), "unit": UNIT, "ms_per_step": fms,
                 "max_rel_diff_psi2_vs_two_phase": float((q2 - p2).abs().max() / p2.abs().max()),
                 "max_rel_diff_dZ_vs_two_phase": float((fo[2] - out[2]).abs().max() / out[2].abs().max()),
                 "note": "rgp_psi_fused_dev: one pass for statistics + gradients (SVI bound); not the headline metric"}

    # Informational: the shapes of the reference's own models (configs 1 - 4: M = 50 ... 100, Q = 10 ... 40), which are
    # served by the small-inducing-set kernels, at 2^20 rows on this GPU.
    shapes = model_shapes(local, peak_tf) if (rank == 0 and not args.no_model_shapes) else None

    # dominant kernel -> roofline (fp64 CUDA-core pipe; see DESIGN.md "Measurement")
    roof = None
    if ktimes:
        name, (tot_ms, cnt) = max(ktimes.items(), key=lambda kv: kv[1][0])
        per_launch_ms = tot_ms / max(cnt, 1)
        rows_per_launch = N * args.steps / max(cnt, 1)
        fl = flops_row_bwd_psi2(M, Q) if "bwd" in name else (flops_row_fwd_psi2(M, Q) if "psi2" in name else None)
        if fl is not None:
            ach = fl * rows_per_launch / (per_launch_ms * 1e-3) / 1e12
            traffic = tnote = None
            try:                # dram bytes/row of this kernel from the committed ncu capture, scaled to this launch
                with open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")) as f:
                    tr = json.load(f).get(name)
                if tr and (M, Q) == (512, 64):
                    traffic = tr["bytes_per_row"] * rows_per_launch
                    tnote = ("dram__bytes_read+write per row from the ncu --set full capture of this kernel (%d-row "
                             "launch, profiles/ncu_traffic_r02.json) x rows per launch; algorithmic %d B/row"
                             % (tr.get("rows", 0), bytes_row_bwd_psi2(M, Q)))
            except Exception:
                traffic = None
            roof = {"bound": "fp64", "kernel": name, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": ach / peak_tf if peak_tf else None, "traffic": traffic, "traffic_note": tnote,
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
    per_gpu = value / world
    hbm_ach = per_gpu * bytes_row(M, Q) / 1e9
    whole = {"achieved_tflops": per_gpu * flops_row_total(M, Q) / 1e12, "peak_tflops": peak_tf,
             "frac": per_gpu * flops_row_total(M, Q) / 1e12 / peak_tf if peak_tf else None,
             "flops_per_row": flops_row_total(M, Q), "bytes_per_row": bytes_row(M, Q),
             "hbm_gbs_algorithmic": hbm_ach, "hbm_peak_gbs": hbm_peak if hbm_peak else 6650.0,
             "hbm_peak_source": "MEASURED_PEAKS.json" if hbm_peak else "fallback (B200_PROFILING.md)",
             "hbm_frac": hbm_ach / (hbm_peak if hbm_peak else 6650.0)}
    kshare = {k: {"ms": round(v[0], 3), "launches": v[1]} for k, v in
              sorted(ktimes.items(), key=lambda kv: -kv[1][0])}

    e2e = e2e_plugin(args, np, torch, dist, world, rank, local, dev, mu, S, Z, ell, variance, dL1, dL2, barrier)

    cpu = None
    if rank == 0 and world == 1 and args.cpu_seconds > 0:
        rps, rows, secs, cores = cpu_path_rows_per_s(M, Q, target_s=args.cpu_seconds)
        cpu = {"value": rps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d rows of the same workload in %.1f s (numpy restatement of GPy's closed "
                         "forms, oracle/psi_oracle.py)" % (rows, secs)}
    if rank == 0:
        cfg = workload_config(args.rows if not strong else N, M, Q)
        cfg.update({"parallelism": "rows sharded x%d, 2 packed all-reduces/step" % world,
                    "l2": "inputs (%.1f GB/GPU) exceed the 126 MB L2; no flush needed"
                          % ((2 * N * Q + 2 * N * M) * 8 / 1e9),
                    "kernels": {0: "auto", 1: "fast", 2: "reference"}[args.kernels]})
        if strong:
            cfg["N_total"] = N_total
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg, "parity": parity,
            "roofline": roof, "whole_step": whole, "kernel_ms": kshare,
            "cpu_baseline": cpu, "e2e": e2e, "strong_scaling": other, "fused_svi_pass": fused, "model_shapes": shapes,
            "gpu_launches": launches, "clocks": clocks, "checksum": checksum,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def e2e_plugin(args, np, torch, dist, world, rank, local, dev, mu, S, Z, ell, variance, dL1, dL2, barrier):
    """End to end the way RGP reaches the path: GPy-shaped kernel accessors (kern.psi0/psi1/psi2, then the
    three gradient accessors) on the drop-in plugin, with ordinary PAGEABLE numpy arrays in and fresh numpy
    arrays out, the plugin's content-keyed cache on (its default).  Every step mutates q(X) in place first (as
    the layer does), so each step computes once per phase and serves the other two accessors from the cache."""
    from rgp_b200.gpy_compat import RBF, NormalPosterior
    from rgp_b200.psicomp import PSICOMP_RBF_B200
    N, M, Q = mu.shape[0], Z.shape[0], Z.shape[1]
    per_row = 8 * (4 * Q + 2 * M + 2 * Q) + 8 * (2 * Q + M)       # caller arrays + results + cached references
    avail = None
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                avail = int(ln.split()[1]) * 1024
    except Exception:
        pass
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    Ne = min(N, args.e2e_rows) if args.e2e_rows > 0 else N
    if avail:
        cap = int(0.6 * avail / local_world / per_row)
        if cap < Ne:
            Ne = max(1 << 16, (cap >> 16) << 16)
    mu_h, S_h = mu[:Ne].cpu().numpy(), S[:Ne].cpu().numpy()       # plain (pageable) numpy
    dL1_h = dL1[:Ne].cpu().numpy()
    Z_h, ell_h, dL2_h = Z.cpu().numpy(), ell.cpu().numpy(), dL2.cpu().numpy()
    dL0_h = np.full(Ne, -0.5)
    pc = PSICOMP_RBF_B200(device=local)
    kern = RBF(Q, variance, ell_h, ARD=True, psicomp=pc)
    X = NormalPosterior(mu_h, S_h)
    X.mean, X.variance = mu_h, S_h                                # no copies: the caller's arrays

    def e2e_step(i):
        X.mean[i % Ne, 0] += 1e-9                                 # in-place update of q(X) (layers.py:537-543)
        p0, p1, p2 = kern.psi0(Z_h, X), kern.psi1(Z_h, X), kern.psi2(Z_h, X)
        kern.update_gradients_expectations(dL0_h, dL1_h, dL2_h, Z_h, X)
        dZ = kern.gradients_Z_expectations(dL0_h, dL1_h, dL2_h, Z_h, X)
        gmu, gS = kern.gradients_qX_expectations(dL0_h, dL1_h, dL2_h, Z_h, X)
        return float(p2[0, 0]) + float(dZ[0, 0]) + float(gmu[0, 0]) + float(p1[0, 0])

    n0 = pc.handle.launch_count()
    e2e_step(0)                                                   # warm: pinned ring, workspace
    per_step_launches = pc.handle.launch_count() - n0
    barrier()
    t0 = time.perf_counter()
    for i in range(args.e2e_steps):
        e2e_step(i + 1)
    torch.cuda.synchronize()
    te = time.perf_counter() - t0
    t = torch.tensor([te], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    te = float(t.item())
    h2d = 8 * (2 * 2 * Ne * Q + 2 * M * Q + 2 * Q + M * M + Ne * M + Ne)
    d2h = 8 * (Ne * M + M * M + 2 * Ne * Q + M * Q + Q + 1)
    return {"value": world * Ne * args.e2e_steps / te, "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "rows_per_step": Ne, "steps": args.e2e_steps,
            "launches_per_step": per_step_launches, "host_cores": host_cores(),
            "api": "kern.psi0/psi1/psi2 + update_gradients_expectations / gradients_Z_expectations / "
                   "gradients_qX_expectations -> PSICOMP_RBF_B200 (pageable numpy in, fresh numpy out, cache on: "
                   "6 accessor calls, 2 device evaluations, 6 content digests per step)",
            "rows_note": "N per GPU" if Ne == N else "largest row count the host memory allows (%.0f GB available / %d ranks)"
                         % ((avail or 0) / 1e9, local_world)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="psi", choices=["psi", "svi10m"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --rows per GPU; strong: --rows in total, split over the ranks")
    ap.add_argument("--rows", type=int, default=4 * 2 ** 20, help="rows per GPU (headline 4*2^20)")
    ap.add_argument("--M", type=int, default=512)
    ap.add_argument("--Q", type=int, default=64)
    ap.add_argument("--kernels", type=int, default=0, help="0 auto, 1 fast, 2 reference kernels")
    ap.add_argument("--e2e-rows", type=int, default=0, help="rows of the end-to-end measurement (0 = N, capped by host memory)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-rows", type=int, default=128, help="rows per step of the reference (CPU) arm")
    ap.add_argument("--no-fused", action="store_true", help="skip the informational fused-pass measurement")
    ap.add_argument("--no-model-shapes", action="store_true", help="skip the informational M = 100 / 50 measurements")
    ap.add_argument("--no-strong", action="store_true", help="skip the informational strong-scaling measurement")
    ap.add_argument("--svi-steps-total", type=int, default=10_000_000, help="svi10m: time steps of the synthetic sequence")
    ap.add_argument("--svi-minibatches", type=int, default=4, help="svi10m: minibatches evaluated (timed)")
    args = ap.parse_args()
    if args.warmup < 3:
        print("warning: contract asks for >= 3 warm-up steps", file=sys.stderr)
    return reference_arm(args) if args.impl == "reference" else ours(args)


if __name__ == "__main__":
    sys.exit(main())
