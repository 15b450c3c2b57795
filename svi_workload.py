"""BASELINE.json config 4: SVI minibatch RGP on a synthetic 10M-step system-identification sequence.

    python bench.py --workload svi10m [--gpus N]        (torchrun for N > 1, like the headline bench)

What the reference does (svi_experiments/rgp_experiments.py:449-547, autoreg/model.py:215-267,
autoreg/data_streamers.py:203-267): the data set is a LIST of sequences, a minibatch is a slice of that
list, the model is rebuilt on the minibatch's sequences (``set_inputs_and_outputs``) and the KL term of
every layer is re-weighted by ``qU_ratio = |minibatch| / |data set|`` (autoreg/layers.py:231,328,76-79)
so that the minibatch bounds of one epoch add up to the full-data bound.  One optimiser step = one
ELBO + gradient evaluation of the uncollapsed SVI bound (autoreg/inference/svi_vardtc.py:70-215).

Here: the 10M-step record is cut into sequences; a minibatch is a set of whole sequences; its sequences
are dealt to the ranks (rows shard on sequence boundaries, no halo), every rank evaluates its sequences
with ONE fused psi pass per layer (``rgp_psi_fused_dev``) and the row sums are all-reduced
(``DeviceBound(sharded=True)``).  Model of the reference's experiments: one hidden layer, windows 20 / 20,
M = 100 inducing points -> kernel input dimensions 20 (observed layer) and 40 (hidden layer).

Reported: ms per minibatch evaluation (max over ranks, CUDA events), time steps / s, and three parity
figures measured before the timing: all ranks vs ONE GPU on the same minibatch, the bound and gradients
of a minibatch vs the sum over its two halves (testing/minibatch_tests.py:288-296), and a permuted
minibatch vs the same sequences in order (:281-286).
"""
from __future__ import annotations

import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WINS, NDIMS, U_WIN, CTL_DIM, M_IND = (0, 20), (1, 1), 20, 1, 100


def deal_sequences(seq_ids, world, rank):
    """Sequences of a minibatch owned by ``rank``: contiguous blocks, sizes differ by at most one
    (the same rule as rgp_b200.sharded.row_partition, on sequences instead of rows)."""
    n = len(seq_ids)
    if n < world:
        raise ValueError("a minibatch of %d sequences cannot be dealt to %d ranks" % (n, world))
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return list(seq_ids[start:start + base + (1 if rank < extra else 0)])


def gen_sequence(idx, T, dev):
    """Sequence ``idx`` of the synthetic system-identification record, a pure function of idx (so every
    sharding sees the same data): a band-limited excitation u, a static-nonlinear FIR response y with
    measurement noise, both roughly unit scale; plus the initial q(X) of the hidden layer."""
    import torch
    g = torch.Generator(device=dev).manual_seed(910_000 + int(idx))
    f64 = dict(dtype=torch.float64, device=dev)
    Tu = T + U_WIN - 1                                             # model.py:57-62: controls are U_win-1 steps longer
    t = torch.arange(Tu, **f64)
    amp = torch.rand(8, generator=g, **f64) * 0.5 + 0.2
    om = (torch.rand(8, generator=g, **f64) * 0.25 + 0.01)
    ph = torch.rand(8, generator=g, **f64) * 2 * math.pi
    u = (amp[:, None] * torch.sin(om[:, None] * t[None, :] + ph[:, None])).sum(0) / 1.2
    lag = lambda k: u[U_WIN - 1 - k:U_WIN - 1 - k + T]
    y = torch.tanh(0.8 * lag(3) + 0.5 * lag(10)) + 0.3 * lag(1) ** 2 - 0.2 * lag(7) * lag(15)
    y = (y - 0.15) * 1.6 + 0.05 * torch.randn(T, generator=g, **f64)
    lat_mean = torch.cat([y[:1].expand(WINS[1]), y]) + 0.05 * torch.randn(WINS[1] + T, generator=g, **f64)
    lat_var = 0.02 + 0.01 * torch.rand(WINS[1] + T, generator=g, **f64)
    return u[:, None], y[:, None], lat_mean[:, None], lat_var[:, None]


def make_params(dev, qU_ratio):
    """Layer parameters (level 0 = observed layer, Q = 20; level 1 = hidden layer, Q = 40), seeded."""
    import torch
    g = torch.Generator(device=dev).manual_seed(77)
    f64 = dict(dtype=torch.float64, device=dev)
    params = []
    for Q, D in ((WINS[1] * NDIMS[1], NDIMS[0]), (WINS[1] * NDIMS[1] + U_WIN * CTL_DIM, NDIMS[1])):
        params.append(dict(variance=1.2, lengthscale=(torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * Q ** 0.5,
                           Z=torch.randn((M_IND, Q), generator=g, **f64) * 0.8, noise_variance=0.1,
                           qU_mean=torch.randn((M_IND, D), generator=g, **f64) * 0.3,
                           qU_W=torch.randn((M_IND, M_IND), generator=g, **f64) * (0.3 / M_IND ** 0.5),
                           qU_a=0.4, qU_ratio=float(qU_ratio)))
    return params


def build_minibatch(seq_ids, T, dev):
    import torch
    us, ys, lm, lv = zip(*[gen_sequence(i, T, dev) for i in seq_ids])
    Y = torch.cat(ys)
    latents = [(torch.cat(lm), torch.cat(lv))]
    U = torch.cat(us)
    controls = (U, torch.full_like(U, 1e-10))                      # model.py:65
    return Y, latents, controls


def make_model(n_seqs, T, dev, sharded):
    from rgp_b200.inference import DeviceBound
    from rgp_b200.layer import DeviceDeepAutoreg
    return DeviceDeepAutoreg(WINS, NDIMS, [T] * n_seqs, U_win=U_WIN, ctl_dim=CTL_DIM, svi=True,
                             bound=DeviceBound(dev.index, sharded=sharded), device=dev.index)


PARAM_KEYS = ("variance", "lengthscale", "Z", "noise_variance", "qU_mean", "qU_W", "qU_a")


def flat_grads(res):
    import torch
    return torch.cat([torch.as_tensor(r[k]).reshape(-1) for r in res for k in PARAM_KEYS])


def parity(world, rank, dev, T):
    """Small minibatch (2 sequences per rank of T steps): (i) all ranks vs one GPU, (ii) a minibatch vs
    the sum over its two halves with qU_ratio halved, (iii) permuted vs ordered sequences."""
    import torch
    import torch.distributed as dist
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))
    ids = list(range(5000, 5000 + 2 * world))
    out = {"ranks": world, "sequences": len(ids), "steps_per_sequence": T}
    full = make_model(len(ids), T, dev, sharded=False)              # every rank: the whole minibatch on ONE GPU
    L1, r1, lat1, _ = full.evaluate(make_params(dev, 0.25), *build_minibatch(ids, T, dev))
    g1 = flat_grads(r1)
    worst = 0.0
    if world > 1:
        mine = deal_sequences(ids, world, rank)
        shard = make_model(len(mine), T, dev, sharded=True)
        Ls, rs, lats, _ = shard.evaluate(make_params(dev, 0.25), *build_minibatch(mine, T, dev))
        per = WINS[1] + T
        lo = ids.index(mine[0]) * per
        worst = max(abs(float(Ls) - float(L1)) / abs(float(L1)), rel(flat_grads(rs), g1),
                    rel(lats[0][0], lat1[0][0][lo:lo + per * len(mine)]),
                    rel(lats[0][1], lat1[0][1][lo:lo + per * len(mine)]))
        t = torch.tensor([worst], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst = float(t)
    out["all_ranks_vs_one_gpu"] = worst if world > 1 else None
    # (ii) additivity over two half-minibatches (each with half the KL weight), on one GPU
    half = len(ids) // 2
    La = Lb = None
    gsum = None
    for part in (ids[:half], ids[half:]):
        m = make_model(len(part), T, dev, sharded=False)
        Lp, rp, _, _ = m.evaluate(make_params(dev, 0.125), *build_minibatch(part, T, dev))
        gsum = flat_grads(rp) if gsum is None else gsum + flat_grads(rp)
        La, Lb = (Lp, Lb) if La is None else (La, Lp)
    out["two_halves_vs_whole_bound"] = abs(float(La) + float(Lb) - float(L1)) / abs(float(L1))
    out["two_halves_vs_whole_grads"] = rel(gsum, g1)
    # (iii) the same sequences in another order: same bound and parameter gradients
    perm = ids[1:] + ids[:1]
    mp = make_model(len(perm), T, dev, sharded=False)
    Lq, rq, _, _ = mp.evaluate(make_params(dev, 0.25), *build_minibatch(perm, T, dev))
    out["permuted_vs_ordered_bound"] = abs(float(Lq) - float(L1)) / abs(float(L1))
    out["permuted_vs_ordered_grads"] = rel(flat_grads(rq), g1)
    out["max_rel"] = max(v for k, v in out.items() if isinstance(v, float))
    return out


def run_svi10m(args, world, rank, local, dev):
    import torch
    import torch.distributed as dist
    total_steps = args.svi_steps_total
    T = 15625                                                       # 640 sequences of 15 625 steps = 10 M steps
    n_seq_total = total_steps // T
    mb_seqs = 64                                                    # sequences per minibatch: 1 M steps, qU_ratio 0.1
    if mb_seqs % world:
        mb_seqs = (mb_seqs // world) * world
    qU_ratio = mb_seqs / n_seq_total
    par = parity(world, rank, dev, T=2048)
    params = make_params(dev, qU_ratio)
    n_mb = max(1, args.svi_minibatches)
    batches = []
    for b in range(n_mb + 1):                                       # one extra for the warm-up evaluation
        ids = [(b * mb_seqs + s) % n_seq_total for s in range(mb_seqs)]
        mine = deal_sequences(ids, world, rank)
        batches.append(build_minibatch(mine, T, dev))
    model = make_model(mb_seqs // world, T, dev, sharded=world > 1)
    h = model.bound.psi.handle

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, args.warmup // 2)):
        out = model.evaluate(params, *batches[-1])
        float(out[0])
    barrier()
    h.set_option("profile", 1)
    h.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(1, args.steps)
    e0.record()
    n_eval = 0
    for _ in range(reps):
        for b in range(n_mb):
            out = model.evaluate(params, *batches[b])
            float(out[0])                                           # the optimiser reads the bound every step
            n_eval += 1
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t) / n_eval
    launches = h.launch_count()
    ktimes = h.kernel_times()
    h.set_option("profile", 0)
    steps_mb = mb_seqs * T
    rows_mb = steps_mb * len(WINS)                                  # one N x Q row per time step and layer
    if rank == 0:
        kshare = {k: {"ms": round(v[0], 3), "launches": v[1]} for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1][0])[:8]}
        line = {
            "metric": "SVI minibatch ELBO+gradient, time steps/sec (fp64)", "value": steps_mb / (ms * 1e-3),
            "unit": "steps/s", "n_gpus": world, "steps": n_eval, "warmup": max(1, args.warmup // 2),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "svi10m: SVI minibatch RGP, synthetic %d-step system-ID record = %d sequences x %d "
                                   "steps; minibatch = %d sequences (%d steps, qU_ratio %.4f); 1 hidden layer, windows "
                                   "20/20, M=%d, Q=20/40" % (n_seq_total * T, n_seq_total, T, mb_seqs, steps_mb, qU_ratio, M_IND),
                       "sequences_per_rank": mb_seqs // world, "layer_rows_per_minibatch": rows_mb,
                       "parallelism": "sequences of the minibatch dealt to %d ranks; fused psi pass per layer; "
                                      "packed all-reduces of the row sums" % world},
            "layer_rows_per_s": rows_mb / (ms * 1e-3), "parity": par, "kernel_ms_rank0": kshare,
            "gpu_launches": launches, "bound_last_minibatch": float(out[0]),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0
