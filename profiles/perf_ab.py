"""Kernel-only timing of the BASELINE shapes for A/B experiments (device time of the solve kernel from the library's own
CUDA events, best of `reps`).  usage: python profiles/perf_ab.py <tag> [config:batch ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import algames_b200 as ab

tag = sys.argv[1]
todo = sys.argv[2:] or ["B:1024", "B:8192", "D:4096", "E:8192", "C:2048"]
rows = []
for item in todo:
    cfg, B = item.split(":"); B = int(B)
    model, N, dt, obj, con, opts, x0, xf = ab.workloads.CONFIGS[cfg](batch=B)
    gb = ab.GameBatch(model, N, dt, obj, con, B, device=0, lib_path=os.environ.get("AGB_LIB"))
    gb.set_instance_params(x0=x0, xf=xf)
    gb.random_initial(opts.amplitude_init, opts.seed)
    o = ab.Options(**{**opts.to_dict(), "dual_reset": True})
    ms = []
    for r in range(5 if B * N < 400000 else 3):
        out = gb.newton_solve(o, want=("stats", "status"))
        ms.append(gb.last_solve_ms())
    conv = float((out["status"] == 0).mean())
    rows.append({"tag": tag, "config": cfg, "batch": B, "kernel_ms": min(ms), "kernel_ms_all": ms, "converged_fraction": conv,
                 "converged_per_s": conv * B / (min(ms) / 1e3), "newton_per_instance": float(out["stats"][:, 6].mean()),
                 "checksum": float(np.abs(out["stats"][:, :6]).sum())})
    print(json.dumps(rows[-1]), flush=True)
    gb.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open(f"gpurun_out/perf_{tag}.json", "w"), indent=1)
