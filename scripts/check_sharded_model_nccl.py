"""Run under torchrun on >= 2 GPUs: the deep-model objective with sequences sharded over ranks
(DeviceBound(sharded=True), NCCL all-reduces of the row sums) must reproduce the single-GPU
evaluation of all sequences computed on rank 0, and reports the weak-scaling time of one whole
evaluation at a large shape (sequences per rank fixed)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from rgp_b200.inference import DeviceBound  # noqa: E402
from rgp_b200.layer import DeviceDeepAutoreg  # noqa: E402
from synth import make_deep_model, stack_model  # noqa: E402


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def subset(m, seqs):
    sub = dict(m)
    sub["Ys"] = [m["Ys"][s] for s in seqs]
    sub["Us"] = [m["Us"][s] for s in seqs] if m["Us"] is not None else None
    sub["latents"] = [[lvl[s] for s in seqs] for lvl in m["latents"]]
    return sub


def build(m, nDims, dev, sharded):
    cuda = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    Y, latents, controls, params = stack_model(m, to=cuda)
    model = DeviceDeepAutoreg(m["wins"], nDims, [y.shape[0] for y in m["Ys"]], U_win=m["U_win"],
                              ctl_dim=m["Us"][0][0].shape[1] if m["Us"] is not None else 0, svi=m["svi"],
                              bound=DeviceBound(dev.index, sharded=sharded), device=dev.index)
    return model, (params, Y, latents, controls)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for svi in (False, True):
        wins, nDims, lens = (0, 4, 4), (3, 2, 2), tuple(300 + 17 * s for s in range(2 * world))
        m = make_deep_model(seed=31, svi=svi, wins=wins, nDims=nDims, seq_lens=lens, U_win=4, M=48)
        mine = [s for s in range(len(lens)) if s % world == rank]
        model, args = build(subset(m, mine), nDims, dev, sharded=True)
        logL, res, lat, _ = model.evaluate(*args)
        if rank == 0:
            one, a1 = build(m, nDims, dev, sharded=False)
            L1, r1, lat1, _ = one.evaluate(*a1)
            errs = {"logL": abs(float(logL) - float(L1)) / abs(float(L1))}
            for i in range(len(wins)):
                for k in ("variance", "lengthscale", "Z", "noise_variance") + (("qU_mean", "qU_W") if svi else ()):
                    errs["p%d_%s" % (i, k)] = rel(torch.as_tensor(res[i][k]), torch.as_tensor(r1[i][k]))
            offs = [np.cumsum([0] + [wins[i] + T for T in lens]) for i in (1, 2)]
            for lvl in range(2):
                for k in (0, 1):
                    want = torch.cat([lat1[lvl][k][offs[lvl][s]:offs[lvl][s + 1]] for s in mine])
                    errs["lat%d_%d" % (lvl, k)] = rel(lat[lvl][k], want)
            worst = max(errs.values())
            print(json.dumps({"check": "sharded_model_vs_single", "svi": svi, "ranks": world, "worst_rel_err": worst}),
                  flush=True)
            ok = ok and worst < 1e-9
    # weak scaling of one whole evaluation: 2 sequences of 2^16 steps per rank, M = 512, Q = 64 / 32
    wins, nDims, T = (0, 16), (1, 2), 1 << 16
    m = make_deep_model(seed=32, wins=wins, nDims=nDims, seq_lens=(T, T), U_win=16, U_dim=2, M=512)
    model, args = build(m, nDims, dev, sharded=True)
    model.evaluate(*args)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        out = model.evaluate(*args)
        float(out[0])
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        rows = 2 * T * world * len(wins)
        print(json.dumps({"row": "model_eval_sharded", "ranks": world, "rows_per_rank_per_layer": 2 * T, "M": 512,
                          "ms_per_eval_max_over_ranks": float(ms), "layer_rows_per_s": rows / (float(ms) * 1e-3)}),
              flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
