import sys, os
sys.path.insert(0, '.')
import numpy as np, algames_b200 as ab
for name, B, N, kw in (("Q", 2, 5, {"p": 2}), ("B", 2, 8, {})):
    model, N, dt, obj, con, opts, x0, xf = ab.workloads.CONFIGS[name](batch=B, N=N, **kw)
    gb = ab.GameBatch(model, N, dt, obj, con, B, device=0, solver=ab._capi.SOLVER_BAND)
    print(name, gb.band_info())
    gb.set_instance_params(x0=x0, xf=xf); gb.random_initial(opts.amplitude_init, opts.seed)
    o = ab.Options(**{**opts.to_dict(), "outer_iter": 2, "inner_iter": 3})
    out = gb.newton_solve(o)
    print(out["status"], out["stats"][:, 6])
    gb.close()
