"""Row-sharded multi-GPU driver: one process per GPU, rows split across ranks.

Every psi quantity is either row-local (Psi1, dL/dmu, dL/dS) or a sum over rows
(Psi0, Psi2, dL/dZ, dL/dlengthscale, dL/dvariance) - SURVEY.md section 8(e).  The
reference planned exactly this split with MPI and left it dead:
  after the forward   allReduceArrays([psi0, psi2, YRY, psi1Y])  autoreg/inference/svi_vardtc.py:65-67
  after the backward  reduceArrays([kerngrad]) / ([Z.gradient])   autoreg/layers.py:110,137 (commented out)
Here the two exchanges are one packed ``all_reduce(SUM)`` each (NCCL over NVLink on
GPUs; gloo on CPU for the host-logic tests).  Payloads are tiny next to the compute
(M*M+1 and M*Q+Q+1 doubles), so they are latency-bound and not fused into kernels.

This module contains no psi arithmetic: shards are computed by ``DevicePsi`` (CUDA) and
only *combined* here, which is what the gloo tests exercise.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def row_partition(N: int, world_size: int, rank: int, align: int = 1,
                  boundaries: Optional[Sequence[int]] = None) -> Tuple[int, int]:
    """Contiguous row block [start, stop) of ``rank``.

    Blocks differ by at most ``align`` rows.  If ``boundaries`` (sorted row indices where
    sequences start, autoreg/layers.py:481-482 stacks sequences row-wise) is given, cuts
    snap to the nearest boundary so the later latent-gradient scatter
    (layers.py:552-571) needs no halo.  Snapping is strictly monotone: every rank keeps at least
    one whole sequence, and a world larger than the number of sequences raises on EVERY rank (the
    same arguments give the same answer everywhere), so no rank can enter a collective alone.
    """
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    units = (N + align - 1) // align
    base, extra = divmod(units, world_size)
    cuts = [0]
    for r in range(world_size):
        cuts.append(cuts[-1] + (base + (1 if r < extra else 0)) * align)
    cuts = [min(c, N) for c in cuts]
    if boundaries is not None and len(boundaries):
        b = sorted(set(int(x) for x in boundaries if 0 < int(x) < N) | {0, N})
        nseq = len(b) - 1
        if nseq < world_size:
            raise ValueError("cannot give each of %d ranks a whole sequence: only %d sequences" % (world_size, nseq))
        idx = [0]                                   # indices into b, strictly increasing
        for r, c in enumerate(cuts[1:-1], start=1):
            k = min(range(len(b)), key=lambda i: abs(b[i] - c))
            k = max(k, idx[-1] + 1)                  # at least one sequence for rank r-1
            k = min(k, nseq - (world_size - r))      # and one for every rank still to come
            idx.append(k)
        cuts = [b[i] for i in idx] + [N]
    elif N < world_size:
        raise ValueError("cannot shard %d rows over %d ranks" % (N, world_size))
    return cuts[rank], cuts[rank + 1]


def pack(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([t.reshape(-1) for t in tensors])


def unpack(buf: torch.Tensor, like: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    out, off = [], 0
    for t in like:
        n = t.numel()
        out.append(buf[off:off + n].reshape(t.shape))
        off += n
    return out


def allreduce_packed(tensors: Sequence[torch.Tensor], group=None) -> List[torch.Tensor]:
    """One SUM all-reduce over the concatenation of ``tensors`` (same dtype/device)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(tensors)
    buf = pack(tensors)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return unpack(buf, tensors)


def reduce_forward(psi0_sum: torch.Tensor, psi2: torch.Tensor, extra: Sequence[torch.Tensor] = (),
                   group=None):
    """Exchange C1: {Psi0 sum, Psi2[, Psi1^T Y, YRY ...]} summed over ranks."""
    out = allreduce_packed([psi0_sum.reshape(1), psi2, *extra], group)
    return out[0].reshape(()), out[1], out[2:]


def reduce_backward(dvar: torch.Tensor, dell: torch.Tensor, dZ: torch.Tensor, group=None):
    """Exchange C2: {dvariance, dlengthscale, dZ} summed over ranks."""
    out = allreduce_packed([dvar.reshape(1), dell, dZ], group)
    return out[0], out[1], out[2]


class ShardedPsi:
    """Psi statistics over rows sharded across the ranks of ``group``.

    Each rank passes only ITS rows of mu / S / dL_dpsi1 and gets back its rows of
    Psi1 / dmu / dS plus the globally reduced Psi0 sum, Psi2, dvar, dell, dZ.
    """

    def __init__(self, device: Optional[int] = None, group=None, impl: int = 0):
        from .device import DevicePsi
        self.psi = DevicePsi(device, impl=impl)
        self.group = group

    def forward(self, mu, S, Z, ell, variance: float, want_psi1: bool = True):
        _, psi1, psi2 = self.psi.forward(mu, S, Z, ell, variance, want_psi1=want_psi1)
        psi0_sum = torch.full((1,), variance * mu.shape[0], dtype=torch.float64, device=mu.device)
        psi0_sum, psi2, _ = reduce_forward(psi0_sum, psi2, group=self.group)
        return psi0_sum, psi1, psi2

    def backward(self, mu, S, Z, ell, variance: float, dL_dpsi0, dL_dpsi1, dL_dpsi2):
        dvar, dell, dZ, dmu, dS = self.psi.backward(mu, S, Z, ell, variance, dL_dpsi0, dL_dpsi1, dL_dpsi2)
        dvar, dell, dZ = reduce_backward(dvar, dell, dZ, group=self.group)
        return dvar, dell, dZ, dmu, dS
