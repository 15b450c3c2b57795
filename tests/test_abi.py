"""CPU tests of the boundary: the C-ABI library loads and exports every symbol that
include/rgp_psi.h declares; the host-side plugin logic (argument handling, errors,
pickling) behaves like GPy's psicomp.  No compute calls without a GPU."""
import copy
import os
import pickle
import re

import numpy as np
import pytest

import rgp_b200
from rgp_b200 import _lib
from rgp_b200.gpy_compat import RBF, NormalPosterior
from rgp_b200.psicomp import PSICOMP_RBF_B200, _fingerprint
from conftest import HAS_GPU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rgp_psi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rgp_(?:psi|lag|latent|mlp|host)_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    lib = rgp_b200.load_library()
    names = _declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_lib.SIGNATURES) == names           # the ctypes table covers the header exactly
    assert lib.rgp_psi_abi_version() == 4


def test_shared_object_contains_sm100a_code_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", rgp_b200.library_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    assert not re.search(r"sm_(?!100a)\d+", out.stdout)


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_create_fails_loudly_without_a_device():
    with pytest.raises(rgp_b200.PsiError) as ei:
        rgp_b200.Handle(0)._ensure()
    assert "no CPU fallback" in str(ei.value)
    kern = RBF(3, ARD=True)
    X = NormalPosterior(np.zeros((4, 3)), np.ones((4, 3)))
    with pytest.raises(rgp_b200.PsiError):
        kern.psi2(np.zeros((2, 3)), X)


def test_invalid_arguments_are_reported_not_aborted():
    lib = rgp_b200.load_library()
    assert lib.rgp_psi_create(0, None) == -1
    assert b"null" in lib.rgp_psi_last_error()
    assert lib.rgp_psi_forward_dev(None, None, 1, 1, 1, None, None, None, None, 1.0, None, None, None) == -1
    assert lib.rgp_psi_set_option(None, b"impl", 0) == -1
    assert lib.rgp_psi_launch_count(None) == -1


def test_psicomp_rejects_non_normal_posteriors_like_gpy():
    pc = PSICOMP_RBF_B200()
    kern = RBF(2, ARD=True, psicomp=pc)
    with pytest.raises(ValueError, match="unknown distriubtion"):
        pc.psicomputations(kern, np.zeros((3, 2)), np.zeros((5, 2)))

    class SpikeAndSlab(NormalPosterior):
        binary_prob = 0.5

    with pytest.raises(ValueError):
        pc.psicomputations(kern, np.zeros((3, 2)), SpikeAndSlab(np.zeros((5, 2)), np.ones((5, 2))))
    with pytest.raises(ValueError, match="shape mismatch"):
        pc.psicomputations(kern, np.zeros((3, 4)), NormalPosterior(np.zeros((5, 2)), np.ones((5, 2))))
    with pytest.raises(NotImplementedError):
        pc.psicomputations(kern, np.zeros((3, 2)), NormalPosterior(np.zeros((5, 2)), np.ones((5, 2))),
                           return_psi2_n=True)


def test_psicomp_survives_deepcopy_and_pickle_without_device_state():
    pc = PSICOMP_RBF_B200(device=0, impl="reference")
    kern = RBF(3, ARD=True, inv_l=True, psicomp=pc)
    k2 = copy.deepcopy(kern)                           # minibatch_tests.py:91 deep-copies models
    assert k2.psicomp is not pc and k2.psicomp.impl == "reference"
    k3 = pickle.loads(pickle.dumps(kern))
    assert k3.psicomp.impl == "reference" and k3.psicomp._handle._h is None


def test_fingerprint_sees_in_place_mutation():
    a = np.arange(12.0).reshape(3, 4)
    f1 = _fingerprint(a)
    a[1, 2] += 1e-9                                    # layers.py:537-543 mutates X in place
    assert _fingerprint(a) != f1
    assert _fingerprint(a.copy()) == _fingerprint(a)


def test_fingerprint_is_order_sensitive():
    """The layer rewrites X in place with the same rows in a NEW ORDER in the reference's permuted
    minibatch tests (testing/minibatch_tests.py:281-296, autoreg/layers.py:528-550): a row permutation,
    a swap of two rows and a swap of two elements must all change the key (round 1's sum/xor key did not)."""
    rng = np.random.default_rng(3)
    mu, S = rng.normal(size=(50, 7)), rng.uniform(0.1, 1, size=(50, 7))
    f0 = _fingerprint(mu, S)
    perm = rng.permutation(50)
    assert _fingerprint(mu[perm], S[perm]) != f0
    assert _fingerprint(mu[::-1], S) != f0
    mu2 = mu.copy()
    mu2[[3, 17]] = mu2[[17, 3]]                       # two rows of mu alone
    assert _fingerprint(mu2, S) != f0
    mu3 = mu.copy()
    mu3[4, 1], mu3[4, 2] = mu3[4, 2], mu3[4, 1]       # two elements
    assert _fingerprint(mu3, S) != f0
    mu.flat[:] = mu[perm].ravel()                     # in place, same object
    assert _fingerprint(mu, S) != f0
    assert _fingerprint(mu.reshape(7, 50)) != _fingerprint(mu)          # shape is part of the key


def test_fingerprint_large_arrays_are_hashed_in_ordered_slices():
    """Arrays above one 8 MiB slice are hashed slice by slice on several threads; exchanging two whole
    slices, or two words inside one slice, must change the key; equal content gives equal keys whatever
    the thread count."""
    import ctypes as C
    from rgp_b200._lib import load
    a = np.arange(3 * (1 << 20) + 5, dtype=np.float64)            # 24 MiB + a ragged tail: 4 slices
    f0 = _fingerprint(a)
    b = a.copy()
    n = 1 << 20                                                    # doubles per slice
    b[:n], b[n:2 * n] = a[n:2 * n], a[:n]
    assert _fingerprint(b) != f0
    c = a.copy()
    c[5], c[6] = a[6], a[5]
    assert _fingerprint(c) != f0
    assert _fingerprint(a.copy()) == f0
    lib = load()
    outs = []
    for threads in (1, 2, 7):
        out = (C.c_uint64 * 2)()
        assert lib.rgp_host_digest(C.c_void_p(a.ctypes.data), a.nbytes, threads, out) == 0
        outs.append(bytes(out))
    assert outs[0] == outs[1] == outs[2]
    out = (C.c_uint64 * 2)()
    assert lib.rgp_host_digest(None, 0, 0, out) == 0                # empty buffer is fine
    assert lib.rgp_host_digest(None, 8, 0, out) == -1 and b"digest" in lib.rgp_psi_last_error()
    # every single-bit flip of a short buffer changes both halves of the digest
    base = np.arange(40, dtype=np.uint8)
    d0 = _fingerprint(base)[0][1]
    for byte in range(40):
        x = base.copy()
        x[byte] ^= 1
        d = _fingerprint(x)[0][1]
        assert d[:8] != d0[:8] and d[8:] != d0[8:]


def test_cache_hands_out_copies_below_the_threshold_and_the_stored_arrays_above():
    pc = PSICOMP_RBF_B200(cache=True, cache_copy_bytes=100)
    small = (np.zeros(3), 1.5)
    entry = pc._store(small)
    out = pc._hand_out(entry)
    assert out[0] is not small[0] and entry[0][0] is not small[0] and out[1] == 1.5
    big = (np.zeros(100),)
    entry = pc._store(big)
    assert pc._hand_out(entry)[0] is big[0]             # GPy's Cache_this returns the stored object
    k = pickle.loads(pickle.dumps(pc))
    assert k.cache_copy_bytes == 100


def test_inv_lengthscale_chain_rule_matches_reference_kernels():
    kern = RBF(2, ARD=True, inv_l=True, lengthscale=[2.0, 0.5])
    np.testing.assert_allclose(kern.lengthscale, [2.0, 0.5])
    np.testing.assert_allclose(kern.inv_l, [0.25, 4.0])


def test_plugin_handles_zero_rows_without_a_device():
    """Empty q(X) (N = 0): the sums over rows are empty - zeros of the right shapes, as GPy's numpy code
    returns - and no kernel is launched (so this runs without a GPU)."""
    import numpy as np
    from oracle.psi_oracle import psi_backward, psi_forward
    from rgp_b200.gpy_compat import RBF, NormalPosterior
    from rgp_b200.psicomp import PSICOMP_RBF_B200
    M, Q = 5, 3
    Z = np.random.default_rng(0).normal(size=(M, Q))
    ell = np.array([1.0, 2.0, 0.5])
    pc = PSICOMP_RBF_B200(cache=False)
    kern = RBF(Q, 1.3, ell, ARD=True, psicomp=pc)
    X = NormalPosterior(np.empty((0, Q)), np.empty((0, Q)))
    p0, p1, p2 = pc.psicomputations(kern, Z, X)
    o0, o1, o2 = psi_forward(1.3, ell, Z, X.mean, X.variance)
    assert p0.shape == o0.shape == (0,) and p1.shape == o1.shape == (0, M)
    np.testing.assert_array_equal(p2, o2)
    out = pc.psiDerivativecomputations(kern, np.empty(0), np.empty((0, M)), np.ones((M, M)), Z, X)
    ref = psi_backward(np.empty(0), np.empty((0, M)), np.ones((M, M)), 1.3, ell, Z, X.mean, X.variance)
    assert out[0] == 0.0 and float(ref[0]) == 0.0
    for a, b in zip(out[1:], ref[1:]):
        assert np.shape(a) == np.shape(b)
        np.testing.assert_allclose(a, b, atol=0)
