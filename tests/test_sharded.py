"""CPU tests (gloo, world_size 2) of the multi-GPU host logic in rgp_b200/sharded.py:
row partitioning and the two packed all-reduces.  Shards are computed by the ORACLE here
(test code is allowed to; the product path computes them with CUDA), so what is tested
is that combining shards the way ShardedPsi does reproduces the full-batch result - the
reference's own additivity property (testing/minibatch_tests.py:288-296)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rgp_b200.sharded import (allreduce_packed, pack, reduce_backward, reduce_forward,
                              row_partition, unpack)
from oracle.psi_oracle import psi_backward, psi_forward
from synth import make_inputs, make_upstream, relerr


def test_row_partition_covers_rows_exactly_once():
    for N in (1, 7, 64, 1000, 4194304):
        for w in (1, 2, 3, 4, 8):
            if w > N:
                with pytest.raises(ValueError):           # every rank raises, none enters a collective alone
                    row_partition(N, w, 0)
                continue
            cuts = [row_partition(N, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == N
            for a, b in zip(cuts[:-1], cuts[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    assert row_partition(100, 4, 1, align=16) == (32, 64)
    with pytest.raises(ValueError):
        row_partition(10, 2, 2)


def test_row_partition_snaps_to_sequence_boundaries():
    # three sequences of 40, 25 and 35 rows stacked (layers.py:481-482)
    cuts = [row_partition(100, 2, r, boundaries=[0, 40, 65]) for r in range(2)]
    assert cuts == [(0, 40), (40, 100)]


def test_row_partition_never_hands_out_an_empty_block():
    """Snapping used to collapse cuts ([0, 0, 9, 9, 10] for N = 10, 4 ranks, boundaries [0, 9]): a rank
    with zero rows then raised inside DevicePsi while its peers sat in all_reduce.  Now every rank
    owns at least one sequence, or every rank raises."""
    for r in range(4):
        with pytest.raises(ValueError, match="only 2 sequences"):
            row_partition(10, 4, r, boundaries=[0, 9])
    rng = np.random.default_rng(0)
    for trial in range(200):
        nseq = int(rng.integers(1, 12))
        lens = rng.integers(1, 50, size=nseq)
        starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
        N = int(lens.sum())
        for w in range(1, nseq + 1):
            cuts = [row_partition(N, w, r, boundaries=starts) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == N
            for (a0, a1), (b0, b1) in zip(cuts[:-1], cuts[1:]):
                assert a1 == b0
            for a0, a1 in cuts:
                assert a1 > a0 and a0 in set(starts.tolist()) | {N}


def test_pack_unpack_roundtrip():
    ts = [torch.arange(6.0).reshape(2, 3), torch.ones(1), torch.arange(4.0)]
    out = unpack(pack(ts), ts)
    for a, b in zip(ts, out):
        assert torch.equal(a, b)
    assert [torch.equal(a, b) for a, b in zip(allreduce_packed(ts), ts)] == [True] * 3   # no group: identity


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, M, Q, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        var, ell, Z, mu, S = make_inputs(N, M, Q, seed=77)
        dL0, dL1, dL2 = make_upstream(N, M, seed=78)
        s, e = row_partition(N, world, rank)
        p0, p1, p2 = psi_forward(var, ell, Z, mu[s:e], S[s:e])          # this rank's shard
        psi0_sum, psi2, _ = reduce_forward(torch.tensor([p0.sum()]), torch.from_numpy(p2))
        dvar, dl, dZ, dmu, dS = psi_backward(dL0[s:e], dL1[s:e], dL2, var, ell, Z, mu[s:e], S[s:e])
        gvar, gl, gZ = reduce_backward(torch.tensor([dvar]), torch.from_numpy(dl), torch.from_numpy(dZ))
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), psi0=psi0_sum.numpy(), psi2=psi2.numpy(),
                 psi1=p1, dvar=gvar.numpy(), dl=gl.numpy(), dZ=gZ.numpy(), dmu=dmu, dS=dS, s=s, e=e)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_reduction_reproduces_full_batch(tmp_path):
    N, M, Q, world = 96, 9, 4, 2
    mp.spawn(_worker, args=(world, _free_port(), N, M, Q, str(tmp_path)), nprocs=world, join=True)
    var, ell, Z, mu, S = make_inputs(N, M, Q, seed=77)
    dL0, dL1, dL2 = make_upstream(N, M, seed=78)
    f = psi_forward(var, ell, Z, mu, S)
    b = psi_backward(dL0, dL1, dL2, var, ell, Z, mu, S)
    rows1, rowsmu, rowsS = [], [], []
    for r in range(world):
        g = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        np.testing.assert_allclose(g["psi0"], f[0].sum(), rtol=1e-14)      # ELBO-level: rtol 1e-14
        assert relerr(g["psi2"], f[2]) < 1e-13
        assert abs(g["dvar"][0] - b[0]) < 1e-11 * abs(b[0])                # gradients: rtol 1e-11
        assert relerr(g["dl"], b[1]) < 1e-11 and relerr(g["dZ"], b[2]) < 1e-11
        rows1.append(g["psi1"]); rowsmu.append(g["dmu"]); rowsS.append(g["dS"])
    assert relerr(np.vstack(rows1), f[1]) < 1e-14
    assert relerr(np.vstack(rowsmu), b[3]) < 1e-13 and relerr(np.vstack(rowsS), b[4]) < 1e-13


def test_minibatch_sequences_are_dealt_in_contiguous_blocks():
    """svi_workload.deal_sequences (config 4): every sequence of a minibatch goes to exactly one rank, blocks are
    contiguous (so a rank's latent gradients are one slice of the single-GPU result) and differ by at most one."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from svi_workload import deal_sequences
    ids = list(range(100, 164))
    for world in (1, 2, 3, 8):
        parts = [deal_sequences(ids, world, r) for r in range(world)]
        assert sum(parts, []) == ids
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    with pytest.raises(ValueError):
        deal_sequences(ids[:3], 4, 0)
