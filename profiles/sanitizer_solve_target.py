"""compute-sanitizer target: small structured solves (block-pivot Gauss-Jordan, every layout family) and one fused MPC run."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, algames_b200 as ab
for name, B, N in (("B", 3, 8), ("E", 2, 8), ("C", 1, 6), ("A", 1, None)):
    kw = {"batch": B}
    if N: kw["N"] = N
    model, N_, dt, obj, con, opts, x0, xf = (ab.workloads.CONFIGS[name](**kw) if name != "A" else ab.workloads.config_a())
    B_ = x0.shape[0]
    gb = ab.GameBatch(model, N_, dt, obj, con, B_, device=0)
    gb.set_instance_params(x0=x0, xf=xf); gb.random_initial(opts.amplitude_init, opts.seed)
    o = ab.Options(**{**opts.to_dict(), "outer_iter": 2, "inner_iter": 3})
    out = gb.newton_solve(o)
    print(name, out["status"], out["stats"][:, 6])
    gb.close()
