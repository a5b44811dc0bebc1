"""Times one solve of every BASELINE configuration on one GPU (device time of the solve kernel, CUDA events inside the library)
and writes gpurun_out/configs.json.  Not a bench line: bench.py measures config B only."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import algames_b200 as ab
rows = []
for cfg, B in [("A", 1024), ("A'", 1024), ("B", 1024), ("B", 8192), ("C", 8192), ("D", 4096), ("E", 8192)]:
    if cfg in ("A", "A'"):
        model, N, dt, obj, con, opts, x0, xf = ab.workloads.CONFIGS[cfg]()
        x0 = np.tile(x0, (B, 1))                # the reference's own single-instance cases, replicated
    else:
        model, N, dt, obj, con, opts, x0, xf = ab.workloads.CONFIGS[cfg](batch=B)
    gb = ab.GameBatch(model, N, dt, obj, con, B, device=0)
    gb.set_instance_params(x0=x0, xf=xf)
    gb.random_initial(opts.amplitude_init, opts.seed)
    ms = []
    for r in range(3):
        out = gb.newton_solve(opts, want=("stats", "status"))
        ms.append(gb.last_solve_ms())
    st = out["stats"]; conv = float((out["status"] == 0).mean())
    rows.append({"config": cfg, "batch": B, "players": model.p, "N": N, "model": type(model).__name__, "kernel_ms": min(ms),
                 "converged_fraction": conv, "converged_instances_per_s": conv * B / (min(ms) / 1e3),
                 "newton_steps_per_instance": float(st[:, 6].mean()), "residual_evals_per_instance": float(st[:, 8].mean())})
    print(rows[-1], flush=True)
    gb.close()
# MPC loop (config D): 1024 streams x 200 warm-started re-solves, disturbance on x0 every step
for B, steps in [(1024, 200)]:
    model, N, dt, obj, con, opts, x0, xf = ab.workloads.config_d(batch=B)
    gb = ab.GameBatch(model, N, dt, obj, con, B, device=0)
    for collect in (True, False):
        t0 = time.perf_counter()
        r = ab.mpc.mpc_run(gb, opts, x0, steps, xf=xf, disturbance_std=1e-3, seed=3, collect=collect)
        torch.cuda.synchronize()
        dtm = time.perf_counter() - t0
        row = {"config": "D-mpc", "streams": B, "resolves": steps, "mode": "host-collected" if collect else "device loop (disturbance H2D only)",
               "wall_ms": dtm * 1e3, "resolves_per_s": B * steps / dtm, "last_solve_kernel_ms": gb.last_solve_ms()}
        if collect:
            stats, status, xs = r
            row["converged_fraction"] = float((status == 0).mean()); row["newton_steps_per_resolve"] = float(stats[:, :, 6].mean())
        rows.append(row); print(row, flush=True)
    gb.close()
json.dump(rows, open("gpurun_out/configs.json", "w"), indent=1)
