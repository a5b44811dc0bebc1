"""Generates the golden fixtures in this directory with the NumPy oracle (the Julia reference cannot run in the build
image — SURVEY F2 — so these are outputs of the pinned restatement, not of Algames.jl itself).
    python tests/golden/make_golden.py            # all cases
    python tests/golden/make_golden.py V          # only the named ones
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import oracle.algames_oracle as O  # noqa: E402
import parity  # noqa: E402

CASES = [("A", 1, None, 100), ("A'", 1, None, 101), ("B", 3, 40, 102), ("C", 2, 12, 103), ("E", 2, 14, 104), ("V", 2, 14, 105)]

for name, B, N, seed in CASES:
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    model, N, dt, obj, con, opts, x0, xf = parity.small_config(name, B, N)
    rng = np.random.default_rng(seed)
    Z0 = opts.amplitude_init * rng.random((B, N, model.n + model.m))
    L0 = opts.amplitude_init * rng.random((B, model.p, N - 1, model.n))
    Z, L, stats, conv, nn = [], [], [], [], []
    for b in range(B):
        op = parity.oracle_problem(model, N, dt, obj, con, opts, x0[b], None if xf is None else xf[b])
        O.newton_solve(op, Z0=Z0[b], L0=L0[b])
        last = op.stats[-1]
        Z.append(np.concatenate([op.pdtraj.X, op.pdtraj.U], axis=1)); L.append(op.pdtraj.du.copy())
        stats.append([last.res, last.dyn, last.con, last.sta, last.opt]); conv.append(op.converged); nn.append(op.n_newton)
    d = dict(config=name, N=N, x0=x0[:B], Z0=Z0, L0=L0, Z=np.array(Z), L=np.array(L), stats=np.array(stats),
             converged=np.array(conv), n_newton=np.array(nn))
    if xf is not None:
        d["xf"] = xf[:B]
    fn = os.path.join(HERE, "solve_%s.npz" % name.replace("'", "p"))
    np.savez_compressed(fn, **d)
    print(fn, "converged", conv, "newton", nn)
