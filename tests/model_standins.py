"""Test-only stand-ins that let the host logic of rgp_b200.layer / rgp_b200.inference run on
CPU tensors: same call signatures as DevicePsi / LagWindow, oracle arithmetic.  (Tests may
use the oracle; the product wires in the CUDA implementations and has no CPU path.)"""
import numpy as np
import torch

from oracle import bound_oracle as bo
from oracle.lag_oracle import build_rows, scatter_rows_into
from oracle.psi_oracle import psi_backward, psi_forward
from synth import stack_model  # noqa: F401  (re-exported)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


class OraclePsi:
    def forward(self, mu, S, Z, ell, variance, **kw):
        p0, p1, p2 = psi_forward(variance, ell.numpy(), Z.numpy(), mu.numpy(), S.numpy())
        return t(p0), t(p1), t(p2)

    def backward(self, mu, S, Z, ell, variance, dL0, dL1, dL2, **kw):
        N = mu.shape[0]
        d0 = np.full(N, dL0) if not isinstance(dL0, torch.Tensor) else dL0.numpy()
        out = psi_backward(d0, dL1.numpy(), dL2.numpy(), variance, ell.numpy(), Z.numpy(), mu.numpy(), S.numpy())
        return (torch.tensor([out[0]]),) + tuple(t(a) for a in out[1:])


class OraclePsiFused(OraclePsi):
    """Adds the one-pass entry point of DevicePsi (oracle arithmetic: simply both phases)."""

    def fused(self, mu, S, Z, ell, variance, dL0, dL1, dL2, want_psi1=True):
        _, p1, p2 = self.forward(mu, S, Z, ell, variance)
        return (p1, p2), self.backward(mu, S, Z, ell, variance, dL0, dL1, dL2)


class OracleLag:
    """LagWindow stand-in on stacked CPU tensors."""

    def __init__(self, lat_lens, X_win, X_dim, ctl_lens, U_win, U_dim):
        self.lat_lens, self.ctl_lens = list(lat_lens), (list(ctl_lens) if ctl_lens is not None else None)
        self.X_win, self.X_dim, self.U_win, self.U_dim = X_win, X_dim, U_win, U_dim
        self.N = sum(T - X_win for T in lat_lens)
        self.Q = X_win * X_dim + (U_win * U_dim if ctl_lens is not None else 0)
        self.device = torch.device("cpu")

    def _split(self, a, lens, dim):
        if a is None:
            return [np.zeros((T, dim)) for T in lens]
        out, off = [], 0
        for T in lens:
            out.append(a.numpy()[off:off + T])       # views: scatter writes through
            off += T
        return out

    def gather(self, lat, ctl=None, out=None):
        Xs = self._split(lat, self.lat_lens, 0 if lat is None else self.X_dim)
        Us = self._split(ctl, self.ctl_lens, self.U_dim) if self.ctl_lens is not None else None
        return t(build_rows(Xs, Us, self.X_win, self.U_win))

    def scatter_add(self, dX, lat_grad=None, ctl_grad=None, allocate=True):
        assert not allocate
        gX = self._split(lat_grad, self.lat_lens, 0 if lat_grad is None else self.X_dim)
        gU = self._split(ctl_grad, self.ctl_lens, self.U_dim) if (self.ctl_lens is not None and ctl_grad is not None) else None
        scatter_rows_into(dX.numpy(), gX, gU, self.X_win if lat_grad is not None else 0, self.U_win,
                          self.X_dim if lat_grad is not None else 0, self.U_dim)
        return lat_grad, ctl_grad

    def latent_terms(self, lm, lv, dYm, dYv):
        gm, gv = np.zeros_like(lm.numpy()), np.zeros_like(lv.numpy())
        off, yoff, delta = 0, 0, 0.0
        for T in self.lat_lens:
            N = T - self.X_win
            m, v = lm.numpy()[off:off + T], lv.numpy()[off:off + T]
            gm[off + self.X_win:off + T] += dYm.numpy()[yoff:yoff + N]
            dyv = dYv.numpy()[yoff:yoff + N]
            gv[off + self.X_win:off + T] += dyv if dyv.ndim == 2 else dyv[:, None]
            if self.X_win > 0:
                val, a, b = bo.normal_prior_term(m[:self.X_win], v[:self.X_win])
                delta += val
                gm[off:off + self.X_win] += a
                gv[off:off + self.X_win] += b
            val, b = bo.normal_entropy_term(v[self.X_win:])
            delta += val
            gv[off + self.X_win:off + T] += b
            off, yoff = off + T, yoff + N
        return t(gm), t(gv), torch.tensor(delta, dtype=torch.float64)


def compare_with_oracle(m, out, relerr, tol=1e-10, to_np=lambda a: a.numpy()):
    """Checks a DeviceDeepAutoreg.evaluate result against oracle/model_oracle.py."""
    from oracle.model_oracle import deep_autoreg_oracle
    logL, res, lat_grads, ctl_grads = out
    oL, ores, olat, octl = deep_autoreg_oracle(m["wins"], m["Ys"], m["latents"], m["params"], Us=m["Us"],
                                               U_win=m["U_win"], svi=m["svi"])
    assert abs(float(logL) - oL) <= tol * abs(oL), (float(logL), oL)
    keys = ["variance", "lengthscale", "Z", "noise_variance"] + (["qU_mean", "qU_W", "qU_a"] if m["svi"] else [])
    worst = 0.0
    for i, (r, o) in enumerate(zip(res, ores)):
        for k in keys:
            e = relerr(to_np(r[k]) if isinstance(r[k], torch.Tensor) else r[k], o[k])
            worst = max(worst, e)
            assert e <= tol, (i, k, e)
    for lvl, (g, og) in enumerate(zip(lat_grads, olat)):
        for k in (0, 1):
            e = relerr(to_np(g[k]), np.vstack([s[k] for s in og]))
            worst = max(worst, e)
            assert e <= tol, (lvl, k, e)
    if octl is not None:
        for k in (0, 1):
            e = relerr(to_np(ctl_grads[k]), np.vstack([s[k] for s in octl]))
            worst = max(worst, e)
            assert e <= tol, ("ctl", k, e)
    return worst
