"""Tiny fast-path run for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from synth import make_inputs, make_upstream, relerr
from oracle.psi_oracle import psi_forward, psi_backward
from rgp_b200.psicomp import PSICOMP_RBF_B200
from rgp_b200.gpy_compat import RBF, NormalPosterior
pc = PSICOMP_RBF_B200(impl="fast", cache=False)
for (N, M, Q) in [(37, 70, 20), (21, 130, 64)]:
    var, ell, Z, mu, S = make_inputs(N, M, Q, seed=4)
    dL0, dL1, dL2 = make_upstream(N, M)
    k = RBF(Q, var, ell, ARD=True, psicomp=pc); X = NormalPosterior(mu, S)
    f = pc.psicomputations(k, Z, X); b = pc.psiDerivativecomputations(k, dL0, dL1, dL2, Z, X)
    of = psi_forward(var, ell, Z, mu, S); ob = psi_backward(dL0, dL1, dL2, var, ell, Z, mu, S)
    print((N, M, Q), [relerr(a, c) for a, c in zip(f, of)], [relerr(a, c) for a, c in zip(b, ob)])
