"""Critical-path profile of the fused receding-horizon loop (agb_mpc_run) on config D: needs a library built with
-DAGB_PHASE_TIMING at AGB_LIB (profiles/build_variant.sh timing "-DAGB_PHASE_TIMING" agb_kernels_p3m).  Thread 0 of every CTA
accumulates clock64() deltas per phase over ALL re-solves of its stream.  usage: python profiles/phase_timing_mpc.py [streams] [resolves]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import algames_b200 as ab

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 50
lib = os.environ.get("AGB_LIB")
model, N, dt, obj, con, opts, x0, xf = ab.workloads.config_d(batch=B)
gb = ab.GameBatch(model, N, dt, obj, con, B, device=0, lib_path=lib)
gb.set_history(4)
dist = 1e-3 * np.random.default_rng(3456).standard_normal((R, B, model.n))
first = ab.Options(**{**opts.to_dict(), "dual_reset": True})
for _ in range(2):
    gb.set_instance_params(x0=x0, xf=xf)
    gb.random_initial(opts.amplitude_init, opts.seed)
    stats, status, xs = gb.mpc_run(first, R, 1, dist)
    ms = gb.last_solve_ms()
hist, _ = gb.get_history()
prof = hist.reshape(B, -1)[:, :16]
names = ["residual", "line-search evals", "kkt terminal setup", "phase 1 (Aug)", "phase 2 (GJ | rows | H)", "phase 3 (P update)", "forward sweep",
         "costate pre-pass", "costate recursion", "update_traj", "load kept residual", "load / store / records / AL update", "(GJ alone, inside phase 2)", "final record", "shift + advance", "rollout"]
tot = prof[:, :12].sum(axis=1) + prof[:, 13:16].sum(axis=1)
newton = stats[:, :, 6].sum()
evals = stats[:, :, 8].sum()
rep = {"config": "D (fused MPC loop)", "streams": B, "resolves": R, "loop_ms": ms, "cycles_per_resolve": float(tot.sum() / (B * R)),
       "newton_steps_per_resolve": float(newton / (B * R)), "residual_evals_per_resolve": float(evals / (B * R)),
       "converged_fraction": float((status == 0).mean()), "share": {}, "cycles_per_resolve_by_phase": {}}
for k, nm in enumerate(names):
    rep["share"][nm] = float(prof[:, k].sum() / tot.sum())
    rep["cycles_per_resolve_by_phase"][nm] = float(prof[:, k].sum() / (B * R))
print(json.dumps(rep, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rep, open(f"gpurun_out/phase_timing_mpc_D_{B}.json", "w"), indent=1)
