"""Shape sweep (configs[4] of BASELINE.json): psi statistics + all gradients at N = 64K ... 16M rows, M = 128 ... 1024,
Q = 16 ... 128, on 1 / 2 / 4 / 8 GPUs.

    python scripts/sweep.py [--quick | --small]                  1 GPU
    python -m torch.distributed.run --nproc-per-node G ... scripts/sweep.py

N is the TOTAL row count of a point; with G ranks every rank takes its row block (rgp_b200.sharded.row_partition)
and the two packed all-reduces run inside the timed region, as in bench.py.  One JSON line per point: rows/s (whole
job), fraction of the FP64 roofline per GPU (F_row accounting, measured DFMA peak of rank 0).  Scaling efficiency per
point = rows/s at G GPUs / (G x rows/s at 1 GPU), computed by scripts/sweep_merge.py from the per-G files.
Also (1 GPU only) the small-N system-ID shapes of configs 1-3, as latency per evaluation."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgp_b200.sharded import ShardedPsi, reduce_backward, reduce_forward, row_partition  # noqa: E402


def f_row(M, Q):
    P = M * (M + 1) // 2
    return 4 * P * Q + 2 * M * M * Q + 16 * M * Q + 16 * P


def run(sp, dev, world, rank, N_total, M, Q, warm=1, reps=1):
    s, e = row_partition(N_total, world, rank)
    N = e - s
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    gz = torch.Generator(device=dev).manual_seed(99)
    f64 = dict(dtype=torch.float64, device=dev)
    mu = torch.randn((N, Q), generator=g, **f64)
    S = torch.rand((N, Q), generator=g, **f64) * 0.49 + 0.01
    Z = torch.randn((M, Q), generator=gz, **f64)
    ell = (torch.rand(Q, generator=gz, **f64) * 0.7 + 0.7) * Q ** 0.5
    dL2 = torch.randn((M, M), generator=gz, **f64) / (M * M)
    if N * M * 8 > 30e9:      # 16M x 512: Psi1 and dL_dpsi1 (69 GB each) do not both fit next to the workspace;
        psi1 = torch.randn((N, M), generator=g, **f64) / M      # the backward pass reads the forward's Psi1 as
        dL1 = psi1                                              # its upstream gradient (same traffic, same flops)
    else:
        dL1 = torch.randn((N, M), generator=g, **f64) / M
        psi1 = torch.empty((N, M), **f64)
    dmu, dS = torch.empty((N, Q), **f64), torch.empty((N, Q), **f64)

    def step():
        _, _, p2 = sp.psi.forward(mu, S, Z, ell, 1.3, psi1_out=psi1)
        if world > 1:
            reduce_forward(torch.full((1,), 1.3 * N, **f64), p2)
        out = sp.psi.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2, dmu_out=dmu, dS_out=dS)
        if world > 1:
            reduce_backward(out[0], out[1], out[2])

    for _ in range(warm):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del mu, S, dL1, psi1, dmu, dS
    torch.cuda.empty_cache()
    return float(t)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sp = ShardedPsi(local)
    peak = sp.psi.handle.fp64_peak(reps=5)
    say = (lambda d: print(json.dumps(d), flush=True)) if rank == 0 else (lambda d: None)
    say({"fp64_peak_tflops": peak, "n_gpus": world})
    K, Mi = 1 << 10, 1 << 20
    points = [(64 * K, 128, 16), (Mi, 128, 16), (16 * Mi, 128, 16),
              (64 * K, 256, 32), (Mi, 256, 32), (16 * Mi, 256, 32),
              (64 * K, 512, 64), (Mi, 512, 64), (4 * Mi, 512, 64), (16 * Mi, 512, 64),
              (64 * K, 1024, 128), (Mi, 1024, 128),
              (256 * K, 1024, 64), (Mi, 512, 32), (2 * Mi, 64, 64), (Mi, 128, 64),
              (Mi, 100, 20), (Mi, 200, 40), (Mi, 50, 20), (256 * K, 500, 60)]
    if "--small" in sys.argv:    # the shapes of the reference's own models (small inducing sets) and their neighbours
        points = [(Mi, 100, 20), (Mi, 100, 10), (Mi, 100, 40), (Mi, 112, 23), (Mi, 50, 20), (Mi, 33, 20), (Mi, 128, 16),
                  (Mi, 200, 40)]
    if "--quick" in sys.argv:
        points = [p for p in points if p[0] * f_row(p[1], p[2]) < 3e13]
    for N, M, Q in points:
        big = N * f_row(M, Q) > 2e13 * world
        ms = run(sp, dev, world, rank, N, M, Q, warm=1, reps=1 if big else 3)
        rps = N / (ms * 1e-3)
        say({"N": N, "M": M, "Q": Q, "n_gpus": world, "ms": ms, "rows_per_s": rps,
             "tflops_per_gpu": rps * f_row(M, Q) / 1e12 / world,
             "frac_of_fp64_peak": rps * f_row(M, Q) / 1e12 / world / peak})
    if world == 1:   # configs 1-3: one layer evaluation at the real shapes (latency-bound)
        for name, N, M, Q in [("actuator_hidden", 502, 100, 20), ("actuator_output", 502, 100, 10),
                              ("ballbeam", 490, 50, 20), ("mocap", 408, 200, 40)]:
            ms = run(sp, dev, 1, 0, N, M, Q, warm=3, reps=50)
            say({"config": name, "N": N, "M": M, "Q": Q, "ms_per_eval": ms, "rows_per_s": N / (ms * 1e-3)})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
