"""Golden vectors produced by EXECUTING the reference's own Python in the build container:
    python tests/golden/make_reference_golden.py        (needs /root/reference; not run on the GPU box)

What runs for real (loaded by path from /root/reference, unmodified):
  autoreg/inference/vardtc.py      VarDTC.inference            (certain and uncertain outputs)
  autoreg/inference/svi_vardtc.py  SVI_VarDTC.inference, comp_KL_qU
  autoreg/variational.py           NormalEntropy, NormalPrior  (values and gradients)
  autoreg/util.py                  get_conv_1D
  autoreg/rnn_encoder.py           Mean_var_multilayer         (torch recognition model)

What is stubbed, because GPy / paramz are not installable here (no network): only the thin
third-party helpers those files import, restated from their documented LAPACK semantics -
GPy.util.linalg {jitchol, dtrtrs, dtrtri, dpotri, pdinv, tdot, backsub_both_sides},
GPy.util.diag.add, the VariationalPosterior / NormalPosterior containers, the Posterior
container, and empty base classes.  The kernel object handed to the reference is a stub whose
psi0 / psi1 / psi2 / K return oracle/psi_oracle.py values: the psi statistics themselves live
in GPy and stay pinned by quadrature (tests/test_oracle.py), not by these files.

The fixtures pin oracle/bound_oracle.py, oracle/lag_oracle.py and the latent terms to outputs
of the reference's code (tests/test_reference_golden.py), and the CUDA / device paths are
compared with the same files on the GPU box.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import scipy.linalg as sl

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/autoreg"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))


# --------------------------------------------------------------------------- GPy stubs
def _install_stubs():
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    def jitchol(A, maxtries=5):
        A = np.ascontiguousarray(A)
        try:
            return sl.cholesky(A, lower=True)
        except sl.LinAlgError:
            pass
        d = np.diag(A)
        if np.any(d <= 0.0):
            raise sl.LinAlgError("not pd: non-positive diagonal elements")
        jitter = d.mean() * 1e-6
        for _ in range(maxtries):
            try:
                return sl.cholesky(A + np.eye(A.shape[0]) * jitter, lower=True)
            except sl.LinAlgError:
                jitter *= 10
        raise sl.LinAlgError("not positive definite, even with jitter.")

    def dtrtrs(A, B, lower=1, trans=0, unitdiag=0):
        return sl.lapack.dtrtrs(np.asfortranarray(A), np.asfortranarray(B), lower=lower, trans=int(trans),
                                unitdiag=unitdiag)

    def dtrtri(L):
        return sl.lapack.dtrtri(np.asfortranarray(L), lower=1)[0]

    def dpotri(A, lower=1):
        R, info = sl.lapack.dpotri(np.asfortranarray(A), lower=lower)
        R = np.tril(R) + np.tril(R, -1).T if lower else np.triu(R) + np.triu(R, 1).T   # symmetrify
        return R, info

    def tdot(mat):
        return mat.dot(mat.T)

    def backsub_both_sides(L, X, transpose="left"):
        if transpose == "left":
            tmp, _ = dtrtrs(L, X, lower=1, trans=1)
            return dtrtrs(L, tmp.T, lower=1, trans=1)[0].T
        tmp, _ = dtrtrs(L, X, lower=1, trans=0)
        return dtrtrs(L, tmp.T, lower=1, trans=0)[0].T

    def pdinv(A):
        L = jitchol(A)
        logdet = 2.0 * np.sum(np.log(np.diag(L)))
        Li = dtrtri(L)
        Ai, _ = dpotri(L, lower=1)
        return Ai, L, Li, logdet

    mod("GPy")
    mod("GPy.util")
    lin = mod("GPy.util.linalg")
    for f in (jitchol, dtrtrs, dtrtri, dpotri, tdot, backsub_both_sides, pdinv):
        setattr(lin, f.__name__, f)
    dg = mod("GPy.util.diag")

    def add(A, b):
        A[np.diag_indices(A.shape[0])] += b
    dg.add = add
    sys.modules["GPy.util"].diag = dg
    sys.modules["GPy.util"].linalg = lin

    class VariationalPosterior(object):
        def __init__(self, means, variances, name=None):
            self.mean, self.variance = np.asanyarray(means), np.asanyarray(variances)
            self.num_data, self.input_dim = self.mean.shape
            self.shape = self.mean.shape

    class NormalPosterior(VariationalPosterior):
        pass

    core = mod("GPy.core")
    core.Model = core.Parameterized = type("Parameterized", (object,), {})
    core.Param = lambda name, value, *a: value
    mod("GPy.core.parameterization")
    var = mod("GPy.core.parameterization.variational")
    var.VariationalPosterior, var.NormalPosterior = VariationalPosterior, NormalPosterior
    mod("GPy.inference")
    lfi = mod("GPy.inference.latent_function_inference")
    lfi.LatentFunctionInference = type("LatentFunctionInference", (object,), {})
    post = mod("GPy.inference.latent_function_inference.posterior")

    class Posterior(object):
        def __init__(self, woodbury_inv=None, woodbury_vector=None, K=None, mean=None, cov=None, K_chol=None):
            self.woodbury_inv, self.woodbury_vector = woodbury_inv, woodbury_vector
    post.Posterior = Posterior
    mod("matplotlib")
    mod("matplotlib.pyplot")
    mod("paramz")
    tr = mod("paramz.transformations")
    tr.Transformation, tr._lim_val, tr.epsilon = object, 36.0, np.finfo(np.float64).resolution
    return NormalPosterior


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class StubKern:
    """Serves the psi statistics / K(Z,Z) the reference asks its kernel for."""

    def __init__(self, variance, ell):
        from oracle.psi_oracle import psi_forward
        from oracle.bound_oracle import rbf_K
        self._fwd, self._K, self.variance, self.ell = psi_forward, rbf_K, variance, ell

    def _psi(self, Z, X):
        return self._fwd(self.variance, self.ell, np.asarray(Z), X.mean, X.variance)

    def psi0(self, Z, X): return self._psi(Z, X)[0]
    def psi1(self, Z, X): return self._psi(Z, X)[1]
    def psi2(self, Z, X): return self._psi(Z, X)[2]
    def K(self, Z): return self._K(self.variance, self.ell, np.asarray(Z))


class Lik:
    def __init__(self, v):
        self.variance = np.array([v])


class Grad(np.ndarray):
    """ndarray with a .gradient slot, like a paramz Param."""
    def __new__(cls, a):
        o = np.asarray(a, dtype=np.float64).copy().view(cls)
        o.gradient = np.zeros_like(a)
        return o


def main():
    NormalPosterior = _install_stubs()
    from synth import make_inputs
    vardtc = _load("ref_vardtc", "inference/vardtc.py")
    svi = _load("ref_svi_vardtc", "inference/svi_vardtc.py")
    variational = _load("ref_variational", "variational.py")
    util = _load("ref_util", "util.py")

    # ---- bounds ---------------------------------------------------------------------
    N, M, Q, D = 40, 6, 3, 2
    variance, ell, Z, mu, S = make_inputs(N, M, Q, seed=21, n_control=1)
    rng = np.random.default_rng(22)
    Ym, Yv = rng.normal(size=(N, D)), rng.uniform(0.02, 0.4, size=(N, D))
    noise = 0.13
    W = rng.normal(size=(M, M)) * 0.2
    qU_mean, qU_var = rng.normal(size=(M, D)), W @ W.T + 0.4 * np.eye(M)
    X = NormalPosterior(mu, S)
    kern = StubKern(variance, ell)
    out = dict(variance=variance, ell=ell, Z=Z, mu=mu, S=S, Y=Ym, Y_var=Yv, noise=noise, qU_mean=qU_mean,
               qU_var=qU_var)
    for tag, Y in (("c", Ym), ("u", NormalPosterior(Ym, Yv))):
        post, logL, g = vardtc.VarDTC().inference(kern, X, Z, Lik(noise), Y)
        out["vardtc_%s_logL" % tag] = float(np.squeeze(logL))
        out["vardtc_%s_woodbury_vector" % tag] = post.woodbury_vector
        out["vardtc_%s_woodbury_inv" % tag] = post.woodbury_inv
        for k, v in g.items():
            out["vardtc_%s_%s" % (tag, k)] = np.asarray(v, dtype=np.float64)
        inf = svi.SVI_VarDTC()
        post, logL, g = inf.inference(kern, X, Z, Lik(noise), Y, qU_mean, qU_var)
        KL, dKL_dm, dKL_dv, dKL_dK = inf.comp_KL_qU(qU_mean, qU_var)
        out["svi_%s_logL" % tag] = float(np.squeeze(logL))
        out["svi_%s_woodbury_vector" % tag] = post.woodbury_vector
        for k, v in g.items():
            out["svi_%s_%s" % (tag, k)] = np.asarray(v, dtype=np.float64)
        out["svi_%s_KL" % tag], out["svi_%s_dKL_dqU_mean" % tag] = KL, dKL_dm
        out["svi_%s_dKL_dqU_var" % tag], out["svi_%s_dKL_dKuu" % tag] = dKL_dv, dKL_dK
    np.savez_compressed(os.path.join(HERE, "ref_bounds.npz"), **out)
    print("ref_bounds.npz:", len(out), "arrays")

    # ---- latent prior / entropy -------------------------------------------------------
    T, Dx = 17, 3
    m, v = rng.normal(size=(T, Dx)), rng.uniform(0.01, 2.0, size=(T, Dx))

    def vp(mean, var):
        p = NormalPosterior(Grad(mean), Grad(var))
        return p
    pe, pp = vp(m, v), vp(m, v)
    ent, pri = variational.NormalEntropy(), variational.NormalPrior()
    ev, pv = ent.comp_value(pe), pri.comp_value(pp)
    ent.update_gradients(pe)
    pri.update_gradients(pp)
    np.savez_compressed(os.path.join(HERE, "ref_variational.npz"), mean=m, var=v, entropy_value=ev,
                        entropy_dvar=np.asarray(pe.variance.gradient), prior_value=pv,
                        prior_dmean=np.asarray(pp.mean.gradient), prior_dvar=np.asarray(pp.variance.gradient))
    print("ref_variational.npz")

    # ---- lag windows --------------------------------------------------------------------
    a = rng.normal(size=(11, 2))
    np.savez_compressed(os.path.join(HERE, "ref_conv.npz"), arr=a,
                        **{"win%d" % w: np.asarray(util.get_conv_1D(a, w)).reshape(a.shape[0] - w + 1, -1)
                           for w in (1, 2, 4)})
    print("ref_conv.npz")

    # ---- recognition model --------------------------------------------------------------
    import torch
    enc_mod = _load("ref_rnn_encoder", "rnn_encoder.py")
    enc_out = {}
    for rnn_type, bidir in (("rnn", False), ("gru", False), ("lstm", True)):
        torch.manual_seed(5)
        net = enc_mod.Mean_var_multilayer(2, [3, 2], [2, 1], 4, rnn_type=rnn_type, bidirectional=bidir).double()
        x = torch.from_numpy(rng.normal(size=(9, 3, 3)))                # (seq_len, batch, input_dim)
        means, vars_ = net.forward(x)
        gm = [torch.from_numpy(rng.normal(size=tuple(t.shape))) for t in means]
        gv = [torch.from_numpy(rng.normal(size=tuple(t.shape))) for t in vars_]
        torch.autograd.backward(means + vars_, gm + gv)
        tag = "%s%s" % (rnn_type, "_bi" if bidir else "")
        enc_out[tag + "__input"] = x.numpy()
        for i, (a_, b_, c_, d_) in enumerate(zip(means, vars_, gm, gv)):
            enc_out["%s__mean%d" % (tag, i)] = a_.detach().numpy()
            enc_out["%s__var%d" % (tag, i)] = b_.detach().numpy()
            enc_out["%s__gmean%d" % (tag, i)] = c_.numpy()
            enc_out["%s__gvar%d" % (tag, i)] = d_.numpy()
        for name, p in net.named_parameters():
            enc_out["%s__param__%s" % (tag, name)] = p.detach().numpy()
            enc_out["%s__grad__%s" % (tag, name)] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "ref_encoder.npz"), **enc_out)
    print("ref_encoder.npz:", len(enc_out), "arrays")


if __name__ == "__main__":
    main()
