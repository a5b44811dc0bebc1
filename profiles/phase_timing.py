"""Critical-path profile of the solve kernel: needs a library built with -DAGB_PHASE_TIMING (AGB_EXTRA_NVCC_FLAGS, see
__graft_entry__.build) at AGB_LIB.  Thread 0 of every CTA accumulates clock64() deltas between block barriers; this script
prints the per-phase share of the CTA's lifetime, averaged over the batch.  usage: python profiles/phase_timing.py [config] [batch]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import algames_b200 as ab

name = sys.argv[1] if len(sys.argv) > 1 else "B"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
lib = os.environ.get("AGB_LIB")
model, N, dt, obj, con, opts, x0, xf = ab.workloads.CONFIGS[name](batch=B)
gb = ab.GameBatch(model, N, dt, obj, con, B, device=0, lib_path=lib)
gb.set_instance_params(x0=x0, xf=xf)
gb.random_initial(opts.amplitude_init, opts.seed)
gb.set_history(4)
for _ in range(2):
    out = gb.newton_solve(opts, want=("stats", "status"))
hist, _ = gb.get_history()
prof = hist.reshape(B, -1)[:, :16]
names = ["residual", "line-search evals", "kkt terminal setup", "phase 1 (Aug)", "phase 2 (GJ | rows | H)", "phase 3 (P update)", "forward sweep",
         "costate pre-pass", "costate recursion", "update_traj", "load kept residual", "load / rollout / store / AL update", "(GJ alone, inside phase 2)"]
tot = prof[:, :12].sum(axis=1)
newton = out["stats"][:, 6]
rep = {"config": name, "batch": B, "kernel_ms": gb.last_solve_ms(), "cycles_per_cta_mean": float(tot.mean()), "newton_steps_mean": float(newton.mean()),
       "cycles_per_newton_step": float(tot.sum() / newton.sum()), "share": {}, "cycles_per_newton_step_by_phase": {}}
for k, nm in enumerate(names):
    rep["share"][nm] = float(prof[:, k].sum() / tot.sum())
    rep["cycles_per_newton_step_by_phase"][nm] = float(prof[:, k].sum() / newton.sum())
K = N - 1
rep["per_stage_cycles"] = {nm: rep["cycles_per_newton_step_by_phase"][nm] / K for nm in names[3:9] + names[12:]}
print(json.dumps(rep, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rep, open(f"gpurun_out/phase_timing_{name}_{B}.json", "w"), indent=1)
