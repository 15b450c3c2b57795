"""Recognition model: parity with the reference's own torch module
(tests/golden/ref_encoder.npz, produced by executing autoreg/rnn_encoder.py) and the autograd
bridge that feeds the objective's latent gradients back into the encoder."""
import os

import numpy as np
import pytest
import torch

from model_standins import OracleLag, OraclePsi, stack_model
from rgp_b200.autograd import deep_autoreg_objective
from rgp_b200.encoder import RecognitionEncoder
from rgp_b200.inference import DeviceBound
from rgp_b200.layer import DeviceDeepAutoreg
from synth import make_deep_model

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_encoder.npz")
VARIANTS = {"rnn": ("rnn", False), "gru": ("gru", False), "lstm_bi": ("lstm", True)}


def load_reference_case(tag, device="cpu"):
    r = np.load(G)
    rt, bi = VARIANTS[tag]
    net = RecognitionEncoder([3, 2], [2, 1], 4, rnn_type=rt, bidirectional=bi)
    sd = {k.split("__param__")[1]: torch.from_numpy(r[k]) for k in r.files if k.startswith(tag + "__param__")}
    assert sorted(sd) == sorted(net.state_dict())          # the reference's parameter names
    net.load_state_dict(sd)
    return r, net.to(device)


@pytest.mark.parametrize("tag", sorted(VARIANTS))
def test_encoder_matches_reference_module(tag):
    r, net = load_reference_case(tag)
    means, variances = net(torch.from_numpy(r[tag + "__input"]))
    for i in range(2):
        np.testing.assert_allclose(means[i].detach().numpy(), r["%s__mean%d" % (tag, i)], rtol=0, atol=1e-15)
        np.testing.assert_allclose(variances[i].detach().numpy(), r["%s__var%d" % (tag, i)], rtol=0, atol=1e-15)
    torch.autograd.backward(means + variances,
                            [torch.from_numpy(r["%s__gmean%d" % (tag, i)]) for i in range(2)] +
                            [torch.from_numpy(r["%s__gvar%d" % (tag, i)]) for i in range(2)])
    for name, p in net.named_parameters():
        np.testing.assert_allclose(p.grad.numpy(), r["%s__grad__%s" % (tag, name)], rtol=1e-13, atol=1e-15)


def test_unknown_cell_type_raises_like_the_reference():
    with pytest.raises(ValueError, match="Unknow rnn type"):
        RecognitionEncoder([3], [2], 4, rnn_type="transformer")


def _encoder_model(seed=4):
    """Equal-length sequences (the encoder batches them, model.py:427-431); wins = U_win so that
    the encoder's T_full outputs are the wins[i] + T latent steps of each level."""
    T, w, B = 9, 2, 3
    m = make_deep_model(seed=seed, wins=(0, w, w), nDims=(2, 1, 2), seq_lens=(T,) * B, U_win=w, control=False)
    torch.manual_seed(seed)
    enc = RecognitionEncoder([2, 1], [1, 2], 5, rnn_type="gru")
    x = torch.from_numpy(np.random.default_rng(seed).normal(size=(T + w, B, 2)))
    model = DeviceDeepAutoreg(m["wins"], (2, 1, 2), [T] * B, U_win=w, bound=DeviceBound(psi=OraclePsi()),
                              lag_factory=OracleLag)
    Y, _, _, params = stack_model(m)
    return m, enc, x, model, Y, params


def test_autograd_bridge_feeds_latent_gradients_into_the_encoder():
    m, enc, x, model, Y, params = _encoder_model()
    latents = enc.latents(x)
    for p in params:
        p["Z"] = p["Z"].clone().requires_grad_(True)
    L = deep_autoreg_objective(model, params, Y, latents)
    L.backward()
    # the same thing by hand: analytic latent gradients pushed through the encoder graph
    lat2 = enc.latents(x)
    _, res, lat_grads, _ = model.evaluate([{k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in p.items()}
                                           for p in params], Y, [(a.detach(), b.detach()) for a, b in lat2], None)
    got = {n: p.grad.clone() for n, p in enc.named_parameters()}
    enc.zero_grad()
    torch.autograd.backward([t for pair in lat2 for t in pair], [g for pair in lat_grads for g in pair])
    for n, p in enc.named_parameters():
        torch.testing.assert_close(got[n], p.grad, rtol=1e-12, atol=1e-14)
    for i, p in enumerate(params):
        torch.testing.assert_close(p["Z"].grad, res[i]["Z"], rtol=1e-12, atol=1e-14)


def test_encoder_parameter_gradient_by_finite_differences():
    m, enc, x, model, Y, params = _encoder_model(seed=6)
    L = deep_autoreg_objective(model, params, Y, enc.latents(x))
    L.backward()
    name, p = next((n, q) for n, q in enc.named_parameters() if n.endswith("linear_var.weight"))
    an = p.grad[0, 1].item()
    vals = []
    for sgn in (+1, -1):
        with torch.no_grad():
            p[0, 1] += sgn * 1e-6
            vals.append(float(deep_autoreg_objective(model, params, Y, enc.latents(x))))
            p[0, 1] -= sgn * 1e-6
    fd = (vals[0] - vals[1]) / 2e-6
    assert abs(fd - an) <= 1e-6 * max(1.0, abs(an)), (fd, an)
