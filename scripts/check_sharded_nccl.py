"""Run under torchrun on >= 2 GPUs: rows sharded over ranks with ShardedPsi (NCCL) must
reproduce the single-GPU result computed on rank 0 (reference additivity property,
testing/minibatch_tests.py:288-296: rtol 1e-14 on sums, 1e-11 on gradients)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rgp_b200.device import DevicePsi  # noqa: E402
from rgp_b200.sharded import ShardedPsi, row_partition  # noqa: E402


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N, M, Q = 20000, 192, 24
    g = torch.Generator(device="cpu").manual_seed(5)          # identical on every rank
    f64 = dict(dtype=torch.float64)
    mu = torch.randn((N, Q), generator=g, **f64).to(dev)
    S = (torch.rand((N, Q), generator=g, **f64) * 0.49 + 0.01).to(dev)
    Z = torch.randn((M, Q), generator=g, **f64).to(dev)
    ell = ((torch.rand(Q, generator=g, **f64) * 0.7 + 0.7) * Q ** 0.5).to(dev)
    dL1 = (torch.randn((N, M), generator=g, **f64) / M).to(dev)
    dL2 = torch.randn((M, M), generator=g, **f64).to(dev) / M ** 2
    s, e = row_partition(N, world, rank)
    sp = ShardedPsi(local)
    p0, p1, p2 = sp.forward(mu[s:e].contiguous(), S[s:e].contiguous(), Z, ell, 1.3)
    dvar, dl, dZ, dmu, dS = sp.backward(mu[s:e].contiguous(), S[s:e].contiguous(), Z, ell, 1.3, -0.5,
                                        dL1[s:e].contiguous(), dL2)
    ok = True
    if rank == 0:
        one = DevicePsi(local)
        _, q1, q2 = one.forward(mu, S, Z, ell, 1.3)
        fvar, fl, fZ, fmu, fS = one.backward(mu, S, Z, ell, 1.3, -0.5, dL1, dL2)
        errs = {"psi0": abs(float(p0) - 1.3 * N) / (1.3 * N), "psi2": rel(p2, q2), "psi1": rel(p1, q1[s:e]),
                "dvar": rel(dvar, fvar), "dl": rel(dl, fl), "dZ": rel(dZ, fZ), "dmu": rel(dmu, fmu[s:e]),
                "dS": rel(dS, fS[s:e])}
        print("sharded-vs-single", world, "ranks:", errs, flush=True)
        ok = errs["psi2"] < 1e-13 and max(errs.values()) < 1e-11
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
