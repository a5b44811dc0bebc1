"""Profiling target: ONE newton_solve launch of BASELINE config B (batch 1024, seed 1234) after one warm-up launch, so
that `ncu -k regex:agb_newton_solve -s 1 -c 1` captures exactly one solve.  Writes gpurun_out/profile_solve.json with the
Newton steps the captured launch executed (the per-step normalisation of the ncu counters).  Not a bench."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import algames_b200 as ab

name = sys.argv[1] if len(sys.argv) > 1 else "B"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
model, N, dt, obj, con, opts, x0, xf = ab.workloads.CONFIGS[name](batch=B)
gb = ab.GameBatch(model, N, dt, obj, con, B, device=0)
gb.set_instance_params(x0=x0, xf=xf)
rng = np.random.default_rng(opts.seed)
Z0 = opts.amplitude_init * rng.random((B, N, model.n + model.m)); L0 = opts.amplitude_init * rng.random((B, model.p, N - 1, model.n))
gb.set_initial(Z0, L0)
for _ in range(2):
    out = gb.newton_solve(opts, want=("stats", "status"))
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"config": name, "batch": B, "newton_steps_per_launch": float(out["stats"][:, 6].sum()), "residual_evals_per_launch": float(out["stats"][:, 8].sum()),
           "converged": int((out["status"] == 0).sum()), "kernel_ms_last": gb.last_solve_ms()}, open(f"gpurun_out/profile_solve_{name}.json", "w"))
gb.close()
