"""Time-boxed probe: could stage 2 of the Psi2 backward kernel (T = L . Z', 64 x 64 x 64 per
row and block, IEEE fp64 on DMMA today) run on the tcgen05 INT8 tensor path with an Ozaki-style split?

This script measures the NUMERICAL side on the CPU (numpy, exact integer arithmetic = what INT8 MMAs with INT32
accumulators compute) on operands of the headline shape, and prints the arithmetic of the PROJECTED cost; it does not
run on the GPU.  Scheme: row-scale L and column-scale Z' by powers of two, cut each into s signed 7-bit slices,
form the slice products with k + l < s exactly in integers, recombine in fp64.  Reference: numpy longdouble.

    python scripts/ozaki_probe.py            -> one JSON line per slice count
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from synth import make_inputs, make_upstream  # noqa: E402


def slices(A, axis, s, bits=7):
    """A = 2^e (per row/col) * sum_k S_k 2^(-bits (k+1)),  S_k integer in [-2^bits, 2^bits]."""
    amax = np.abs(A).max(axis=axis, keepdims=True)
    amax[amax == 0] = 1.0
    e = np.ceil(np.log2(amax))                       # |A| / 2^e <= 1
    R = A / np.exp2(e)
    out = []
    for k in range(s):
        R = R * (1 << bits)
        Sk = np.rint(R)
        R = R - Sk
        out.append(Sk.astype(np.int64))
    return out, e


def main():
    M, Q, N = 512, 64, 6
    var, ell, Z, mu, S = make_inputs(N, M, Q, seed=3)
    _, _, dL2 = make_upstream(N, M)
    o = Z.mean(0)
    Zc, muc = Z - o, mu - o
    l2 = ell ** 2
    rows = []
    for n in range(N):
        d = 1.0 / (2 * S[n] + l2)
        ws = S[n] / (l2 * (2 * S[n] + l2))
        b2 = -0.25 * np.log1p(2 * S[n] / l2).sum() - 0.5 * (d * muc[n] ** 2).sum()
        H = b2 + Zc @ (d * muc[n]) - 0.25 * (Zc ** 2) @ (d + 1 / l2)
        I, J = slice(0, 64), slice(64, 128)           # one off-diagonal block
        E = H[I, None] + H[None, J] + (Zc[I] * ws) @ Zc[J].T
        L = var ** 2 * 0.5 * (dL2[I, J] + dL2[J, I].T) * np.exp(E)
        rows.append((L, Zc[J]))
    for s in range(3, 9):
        worst = 0.0
        for L, ZJ in rows:
            ref = (L.astype(np.longdouble) @ ZJ.astype(np.longdouble))
            La, ea = slices(L, 1, s)
            Zb, eb = slices(ZJ, 0, s)
            T = np.zeros((64, Q))
            for k in range(s):
                for l in range(s - k):
                    prod = La[k] @ Zb[l]                 # exact: |sum| <= 64 * 128 * 128 < 2^31
                    assert np.abs(prod).max() < 2 ** 31
                    T += prod.astype(np.float64) * np.exp2(-7.0 * (k + l + 2))
            T = T * np.exp2(ea) * np.exp2(eb)
            worst = max(worst, float(np.abs(T - ref).max() / np.abs(ref).max()))
        fp64 = max(float(np.abs((L @ ZJ) - (L.astype(np.longdouble) @ ZJ.astype(np.longdouble))).max() /
                         np.abs(L @ ZJ).max()) for L, ZJ in rows)
        nprod = s * (s + 1) // 2
        # projected cycles per row and block on one SM (B200: 64 fp64 FMA / clk; INT8 dense 4.5 POP/s = 7.7 K MAC / clk)
        mma = nprod * 64 ** 3 / 7700.0
        split = 64 * 64 * s * 3 / 64.0                  # ~3 FP64-pipe ops per slice and element of L (scale, round, subtract)
        recomb = 64 * Q * s * 2 / 64.0                  # per output: s int64 partial sums -> fp64 (convert + fma)
        print(json.dumps({"probe": "ozaki_int8_stage2", "slices": s, "int8_products": nprod,
                          "max_rel_err_vs_longdouble": worst, "plain_fp64_rel_err": fp64,
                          "projected_cycles": {"int8_mma": round(mma), "slicing_L_fp64_pipe": round(split),
                                               "recombination_fp64_pipe": round(recomb),
                                               "total": round(mma + split + recomb), "dmma_today": 64 ** 3 // 64}}))


if __name__ == "__main__":
    main()
