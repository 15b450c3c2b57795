"""(Lives under tests/: it times the CPU oracle beside the device path.)
Latency of ONE whole objective evaluation (bound + every gradient) of the deep
autoregressive model at the shapes of BASELINE.json configs 1-3, on the device
(rgp_b200.layer.DeviceDeepAutoreg) with the CPU oracle timed beside it on the host cores.
Synthetic data of the configs' shapes (the datasets are not needed for timing)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from model_standins import compare_with_oracle, stack_model  # noqa: E402
from oracle.model_oracle import deep_autoreg_oracle  # noqa: E402
from rgp_b200.layer import DeviceDeepAutoreg  # noqa: E402
from synth import make_deep_model, relerr  # noqa: E402

CONFIGS = [
    # name, wins, nDims, seq_lens, U_win, ctl_dim, M, cpu reps
    ("config1_actuator", (0, 10), (1, 1), (502,), 10, 1, 100, 2),
    ("config2_ballbeam_2hidden", (0, 10, 10), (1, 1, 1), (490,), 10, 1, 50, 2),
    ("config3_mocap", (0, 20, 20), (59, 1, 1), (102, 102, 102, 102), 20, 1, 200, 1),
    ("synthetic_large", (0, 16), (1, 2), (1 << 17, 1 << 17), 16, 2, 512, 0),
    ("synthetic_large_svi", (0, 16), (1, 2), (1 << 17, 1 << 17), 16, 2, 512, 0),
]


def main():
    cuda = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for name, wins, nDims, seq_lens, U_win, ctl_dim, M, cpu_reps in CONFIGS:
        svi = name.endswith("_svi")        # SVI bound: statistics + gradients from one fused pass per layer
        m = make_deep_model(wins=wins, nDims=nDims, seq_lens=seq_lens, U_win=U_win, U_dim=ctl_dim, M=M,
                            control=ctl_dim > 0, svi=svi)
        Y, latents, controls, params = stack_model(m, to=cuda)
        model = DeviceDeepAutoreg(list(wins), nDims, list(seq_lens), U_win=U_win, ctl_dim=ctl_dim, svi=svi, device=0)
        h = model.bound.psi.handle
        out = model.evaluate(params, Y, latents, controls)
        torch.cuda.synchronize()
        reps = 20 if cpu_reps else 2
        h.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            out = model.evaluate(params, Y, latents, controls)
            float(out[0])                                      # the optimiser reads the bound every step
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / reps * 1e3
        rec = {"row": "model_eval", "config": name, "layers": len(wins), "rows_per_layer": int(sum(seq_lens)),
               "M": M, "Q_per_layer": [l.Q for l in model.layers], "gpu_ms_per_eval": e0.elapsed_time(e1) / reps,
               "wall_ms_per_eval": wall, "librgp_launches_per_eval": h.launch_count() / reps}
        if cpu_reps:
            t0 = time.perf_counter()
            for _ in range(cpu_reps):
                deep_autoreg_oracle(m["wins"], m["Ys"], m["latents"], m["params"], Us=m["Us"], U_win=U_win)
            rec["cpu_oracle_ms_per_eval"] = (time.perf_counter() - t0) / cpu_reps * 1e3
            rec["cpu_threads"] = torch.get_num_threads()
            rec["worst_rel_err_vs_oracle"] = compare_with_oracle(
                m, out, relerr, tol=1e-8, to_np=lambda a: a.detach().cpu().numpy())
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
