"""Regenerate tests/golden/*.npz from the CPU oracle (run in the build container):
    python tests/golden/make_golden.py
GPy is not installable here, so these are oracle outputs (oracle/psi_oracle.py, itself
pinned by quadrature and finite differences), stored so the GPU box can compare the
CUDA path against fixed numbers without recomputing them, and so a later change to the
oracle shows up as a diff.  Shapes: a tiny ragged case and the two layers of config 1
(Actuator, N=502, M=100, Q=20 / Q=10)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.psi_oracle import psi_backward, psi_forward  # noqa: E402
from synth import make_inputs, make_upstream  # noqa: E402

CASES = {
    "tiny_ragged": dict(N=13, M=7, Q=3, n_control=1, seed=11),
    "actuator_hidden": dict(N=502, M=100, Q=20, n_control=10, seed=12),
    "actuator_output": dict(N=502, M=100, Q=10, n_control=0, seed=13),
}

if __name__ == "__main__":
    for name, c in CASES.items():
        var, ell, Z, mu, S = make_inputs(c["N"], c["M"], c["Q"], seed=c["seed"], n_control=c["n_control"])
        dL0, dL1, dL2 = make_upstream(c["N"], c["M"], seed=c["seed"] + 100)
        p0, p1, p2 = psi_forward(var, ell, Z, mu, S)
        dvar, dl, dZ, dmu, dS = psi_backward(dL0, dL1, dL2, var, ell, Z, mu, S)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), variance=var, ell=ell, Z=Z, mu=mu, S=S,
                            dL0=dL0, dL1=dL1, dL2=dL2, psi0=p0, psi1=p1, psi2=p2,
                            dvar=dvar, dl=dl, dZ=dZ, dmu=dmu, dS=dS)
        print(name, "written")
