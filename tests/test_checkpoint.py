"""HDF5 checkpoint reader on the authors' own trained models shipped with the reference
(examples/*.h5).  Runs where /root/reference is mounted (the build container); skipped elsewhere.
The parse is validated by the files' own redundancy: ``param_array`` (the flat optimiser vector) must
have exactly as many entries as all named parameter datasets together, and every value of the named
datasets must occur in it."""
import os

import numpy as np
import pytest

from rgp_b200.checkpoint import HDF5FormatError, layer_parameters, load_checkpoint

REF = "/root/reference/examples"
FILES = ["alex_walk_run_m1_sf1.0.h5", "alex_walk_run_m2_sf1.0.h5", "walk_run_2.h5"]


@pytest.mark.parametrize("fname", FILES)
def test_shipped_checkpoints_parse_and_are_self_consistent(fname):
    path = os.path.join(REF, fname)
    if not os.path.exists(path):
        pytest.skip("reference checkpoints are not mounted here")
    ck = load_checkpoint(path)
    flat = ck["param_array"]
    named = {k: v for k, v in ck.items() if k != "param_array"}
    assert flat.ndim == 1 and flat.size == sum(v.size for v in named.values())
    assert np.isfinite(flat).all()
    pool = np.sort(flat)
    for k, v in named.items():                      # every named value is somewhere in the flat vector
        idx = np.clip(np.searchsorted(pool, np.ravel(v)), 0, pool.size - 1)
        assert np.array_equal(pool[idx], np.ravel(v)), k


def test_trained_mocap_model_has_the_shapes_of_config3():
    path = os.path.join(REF, "alex_walk_run_m1_sf1.0.h5")
    if not os.path.exists(path):
        pytest.skip("reference checkpoints are not mounted here")
    layers = layer_parameters(load_checkpoint(path))
    assert [L["Z"].shape for L in layers] == [(100, 20), (100, 40), (100, 40)]      # M = 100, Q = 20 / 40 / 40
    assert all(L["lengthscale"].shape == (L["Z"].shape[1],) and (L["lengthscale"] > 0).all() for L in layers)
    assert all(L["variance"] > 0 and L["noise_variance"] > 0 for L in layers)
    seq = sorted(k for k in layers[1] if k.startswith("qX_") and k.endswith("_mean"))
    assert len(seq) == 4 and all(layers[1][k].shape[1] == 1 for k in seq)            # 4 sequences, 1-d latents


def test_back_constrained_checkpoint_carries_the_mlp():
    path = os.path.join(REF, "walk_run_2.h5")
    if not os.path.exists(path):
        pytest.skip("reference checkpoints are not mounted here")
    layers = layer_parameters(load_checkpoint(path))
    mlp = layers[1]["mlp"]
    assert [W.shape for W, _ in mlp][-1][0] == 1 and all(W.shape[0] == b.shape[0] for W, b in mlp)
    for (W0, _), (W1, _) in zip(mlp[:-1], mlp[1:]):
        assert W1.shape[1] == W0.shape[0]                                            # widths chain


def test_rejects_files_that_are_not_hdf5(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file at all")
    with pytest.raises(HDF5FormatError):
        load_checkpoint(str(p))
