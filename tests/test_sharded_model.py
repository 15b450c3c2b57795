"""World-size-2 gloo test of the sequence-sharded deep model (SURVEY.md 8e applied to f1+f2):
each rank evaluates ITS sequences; the bound and the parameter gradients must equal the
single-process evaluation of all sequences, latent gradients must equal the owner's slices.
Psi / lag arithmetic comes from the oracle-backed stand-ins (host logic under test)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from model_standins import OracleLag, OraclePsi, stack_model
from rgp_b200.inference import DeviceBound
from rgp_b200.layer import DeviceDeepAutoreg
from synth import make_deep_model, relerr

WINS, NDIMS, LENS = (0, 2, 3), (2, 1, 2), (9, 7, 8, 6)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _subset(m, seqs):
    sub = dict(m)
    sub["Ys"] = [m["Ys"][s] for s in seqs]
    sub["Us"] = [m["Us"][s] for s in seqs] if m["Us"] is not None else None
    sub["latents"] = [[lvl[s] for s in seqs] for lvl in m["latents"]]
    return sub


def _evaluate(m, bound):
    Y, latents, controls, params = stack_model(m)
    model = DeviceDeepAutoreg(m["wins"], NDIMS, [y.shape[0] for y in m["Ys"]], U_win=m["U_win"],
                              ctl_dim=1 if m["Us"] is not None else 0, svi=m["svi"], bound=bound, lag_factory=OracleLag)
    return model.evaluate(params, Y, latents, controls)


def _worker(rank, world, port, svi, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = make_deep_model(svi=svi, wins=WINS, nDims=NDIMS, seq_lens=LENS)
        mine = [s for s in range(len(LENS)) if s % world == rank]
        logL, res, lat_grads, _ = _evaluate(_subset(m, mine), DeviceBound(psi=OraclePsi(), sharded=True))
        out = {"logL": float(logL)}
        for i, r in enumerate(res):
            for k in ("variance", "lengthscale", "Z", "noise_variance") + (("qU_mean", "qU_W", "qU_a") if svi else ()):
                out["p%d_%s" % (i, k)] = np.asarray(r[k])
        for lvl, g in enumerate(lat_grads):
            out["g%d_m" % lvl], out["g%d_v" % lvl] = g[0].numpy(), g[1].numpy()
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), **out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("svi", [False, True])
def test_sequence_sharded_model_equals_single_process(tmp_path, svi):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), svi, str(tmp_path)), nprocs=world, join=True)
    m = make_deep_model(svi=svi, wins=WINS, nDims=NDIMS, seq_lens=LENS)
    logL, res, lat_grads, _ = _evaluate(m, DeviceBound(psi=OraclePsi()))
    offs = [np.cumsum([0] + [WINS[i] + T for T in LENS]) for i in (1, 2)]
    for rank in range(world):
        g = np.load(os.path.join(str(tmp_path), "r%d.npz" % rank))
        assert abs(g["logL"] - float(logL)) <= 1e-12 * abs(float(logL))
        for i, r in enumerate(res):
            for k in ("variance", "lengthscale", "Z", "noise_variance") + (("qU_mean", "qU_W", "qU_a") if svi else ()):
                assert relerr(g["p%d_%s" % (i, k)], np.asarray(r[k])) <= 1e-10, (i, k)
        mine = [s for s in range(len(LENS)) if s % world == rank]
        for lvl in range(2):
            for k, name in ((0, "m"), (1, "v")):
                full = lat_grads[lvl][k].numpy()
                want = np.vstack([full[offs[lvl][s]:offs[lvl][s + 1]] for s in mine])
                assert relerr(g["g%d_%s" % (lvl, name)], want) <= 1e-10, (lvl, name)
