"""CPU tier: the plain-C restatement (oracle/algames_oracle.c — the timed CPU baseline) against the NumPy oracle."""
import numpy as np
import pytest

import algames_b200 as ab
import oracle.algames_oracle as O
import parity
from oracle import c_oracle


@pytest.mark.parametrize("name,B,N", [("A", 1, None), ("A'", 1, None), ("B", 3, 40), ("C", 2, 12), ("D", 2, 12), ("V", 2, 12)])
def test_c_oracle_matches_numpy_oracle(name, B, N):
    model, N, dt, obj, con, opts, x0, xf = parity.small_config(name, B, N)
    if x0.shape[0] < B:
        x0 = np.tile(x0[:1], (B, 1))
    p, J = model.p, ab.problem._joint
    desc = ab.problem._make_desc(model, N, dt, obj, con)
    tile = lambda v: np.tile(v, (B, 1))
    xfj = tile(J(obj.xf, p, 4)) if xf is None else xf
    rng = np.random.default_rng(21)
    Z0 = 1e-8 * rng.random((B, N, model.n + model.m)); L0 = 1e-8 * rng.random((B, p, N - 1, model.n))
    out, used = c_oracle.newton_solve(desc, opts.to_c(), x0, xfj, tile(J(obj.Q, p, 4)), tile(J(obj.R, p, 2)),
                                      tile(J(obj.uf, p, 2)), Z0, L0, nthreads=2)
    assert used >= 1
    for b in range(B):
        op = parity.oracle_problem(model, N, dt, obj, con, opts, x0[b], None if xf is None else xf[b])
        O.newton_solve(op, Z0=Z0[b], L0=L0[b])
        Zo = np.concatenate([op.pdtraj.X, op.pdtraj.U], axis=1)
        assert np.abs(out["Z"][b] - Zo).max() < parity.TOL_SOLVE
        assert np.abs(out["L"][b] - op.pdtraj.du).max() < parity.TOL_SOLVE * max(1.0, np.abs(op.pdtraj.du).max())
        assert int(out["stats"][b, 6]) == op.n_newton and (out["status"][b] == 0) == op.converged
        last = op.stats[-1]
        assert np.allclose(out["stats"][b, :5], [last.res, last.dyn, last.con, last.sta, last.opt], atol=parity.TOL_SOLVE)
