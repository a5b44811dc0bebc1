"""Receding-horizon (MPC) driver on top of GameBatch: the caller-side loop that Options.shift = 1 / dual_reset = false
exist for (src/struct/options.jl:16-17, :114-115).  The reference ships no such loop (it lives in AlgamesDriving.jl,
README.md:6), so this is the synthetic BASELINE config D: solve, apply the first control (x0 <- x_2 + disturbance),
shift the previous solution by one knot as the warm start, re-solve keeping duals and penalties."""
from __future__ import annotations

import numpy as np

from .problem import GameBatch, Options


def mpc_run(batch: GameBatch, opts: Options, x0, n_resolves: int, xf=None, disturbance_std: float = 0.0, seed: int = 0,
            stream: int = 0, collect: bool = True, disturbances=None):
    """Run `n_resolves` warm-started re-solves for every stream of the batch.  Returns per-step stats
    [n_resolves, B, 10], status [n_resolves, B] and the executed closed-loop states [n_resolves + 1, B, n]."""
    rng = np.random.default_rng(seed)
    B, n = batch.batch, batch.n
    batch.set_instance_params(x0=x0, xf=xf)
    batch.random_initial(opts.amplitude_init, opts.seed)
    first = Options(**{**opts.to_dict(), "dual_reset": True})
    warm = Options(**{**opts.to_dict(), "dual_reset": False, "shift": 1})
    stats, status, xs = [], [], [np.asarray(x0, float).copy()]
    for t in range(n_resolves):
        if collect:
            out = batch.newton_solve(first if t == 0 else warm, want=("Z", "stats", "status"))
            stats.append(out["stats"]); status.append(out["status"])
        else:
            batch.newton_solve(first if t == 0 else warm, want=())      # same stream as mpc_advance; no D2H copies
        if disturbances is not None:
            d = np.ascontiguousarray(disturbances[t])              # caller-supplied [n_resolves, B, n]
        else:
            d = disturbance_std * rng.standard_normal((B, n)) if disturbance_std > 0 else None
        if collect:
            xs.append(out["Z"][:, 1, :n] + (0.0 if d is None else d))
        batch.mpc_advance(1, d)
    if not collect:
        return None
    return np.array(stats), np.array(status), np.array(xs)
