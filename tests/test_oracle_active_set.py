"""Active-set analysis of the reference (src/active_set/*.jl) restated in the oracle only — it is a host-side
analysis off the solve path.  Restates test/active_set/active_set_stamp.jl, active_set_core.jl and active_set_methods.jl."""
import numpy as np

import oracle.algames_oracle as O


def test_cstamp_validity():
    # test/active_set/active_set_stamp.jl:4-36
    N, p = 10, 4
    table = [("v", 1, 2, 3, True), ("v", 1, 1, 3, False), ("v", 1, 3, 3, True), ("v", 1, 5, 3, False), ("v", 0, 3, 3, False),
             ("v", 1, 3, 1, False), ("v", 3, 1, 3, False), ("v", 1, 2, 11, False),
             ("h", 1, 3, 3, True), ("h", 2, 2, 3, False), ("h", 3, 1, 3, True), ("h", 3, 1, 1, False), ("h", 3, 1, 11, False),
             ("h", 5, 1, 10, False), ("h", 4, 1, 10, True)]
    for dim, i, j, k, want in table:
        assert O.valid_c(dim, "col", i, j, k, N, p) == want, (dim, i, j, k)


def _problem(radius, seed=0):
    rng = np.random.default_rng(seed)
    N, dt, p = 10, 0.1, 3
    model = O.make_model("unicycle", p)
    ps = O.ProblemSize(N, model)
    obj = O.GameObjective([rng.random(4) for _ in range(p)], [rng.random(2) for _ in range(p)],
                          [(i + 1) * np.ones(4) for i in range(p)], [2 * (i + 1) * np.ones(2) for i in range(p)], N, model)
    gc = O.GameConstraintValues(ps)
    gc.add_collision_avoidance(radius)
    return O.GameProblem(N, dt, rng.random(model.n), model, O.Options(), obj, gc)


def test_active_set_core_sizes():
    # test/active_set/active_set_core.jl + active_set_core.jl:81-82
    ps = O.ProblemSize(10, O.make_model("unicycle", 3))
    a = O.ActiveSetCore(ps)
    assert a.Sv == ps.S + 9 * 3 and a.Sh == ps.S + 9 * 6
    assert a.res.shape == (a.Sv,) and a.jac.shape == (a.Sv, a.Sh)
    assert sorted(a.vert.values()) == list(range(ps.S, a.Sv)) and sorted(a.horiz.values()) == list(range(ps.S, a.Sh))


def test_active_masks_all_and_none():
    # test/active_set/active_set_methods.jl:36-82: tiny radius — at the zero trajectory every pair coincides (c = r² ≥ 0: all
    # active); at a spread-out random trajectory none is
    prob = _problem(1e-8)
    ps, gc = prob.probsize, prob.game_con
    a = O.ActiveSetCore(ps)
    pd = prob.pdtraj
    pd.X[:] = 0.0
    gc.evaluate(pd.X, pd.U)
    O.active_vertical_mask(a, gc); O.active_horizontal_mask(a, gc)
    assert np.array_equal(a.vmask, np.arange(ps.S + 9 * 3)) and np.array_equal(a.hmask, np.arange(ps.S + 9 * 6))
    pd.X[:] = 1e3 * np.random.default_rng(100).random(pd.X.shape)
    gc.evaluate(pd.X, pd.U)
    O.active_vertical_mask(a, gc); O.active_horizontal_mask(a, gc)
    assert np.array_equal(a.vmask, np.arange(ps.S)) and np.array_equal(a.hmask, np.arange(ps.S))


def test_nullspace_dimensions():
    # test/active_set/active_set_methods.jl:85-117: radius 1, x0 in [0,1]^n ⇒ every pair active at the rolled-out iterate
    prob = _problem(1.0, seed=1)
    ps = prob.probsize
    N, p = ps.N, ps.p
    O.rollout_rk3(prob.model, prob.pdtraj)
    a = O.ActiveSetCore(ps)
    r = O.as_residual(a, prob)
    J = O.as_residual_jacobian(a, prob)
    assert np.array_equal(r[:ps.S], prob.core.res) and np.abs(r[ps.S:]).max() > 0
    # the border blocks are each other's transposes pair by pair: column (i,j) on opt_i x_k rows == row (min,max) on x_k columns
    k, i, j = 5, 1, 2
    col = J[prob.core.vert[("opt", i, "x", k)], a.horiz[("h", "col", i, j, k)]]
    row = J[a.vert[("v", "col", i, j, k)], prob.core.horiz[("x", k)]]
    assert np.allclose(col, row) and np.abs(col).max() > 0
    mat = O.update_nullspace(a, prob)
    assert mat.shape == (ps.S + (N - 1) * p * (p - 1), (N - 1) * p)                       # :114
    assert len(a.null_vec) == (N - 1) * p and len(a.null_vec[0]) == ps.S + (N - 1) * p * (p - 1)   # :115-116
    djac = a.jac[np.ix_(a.vmask, a.hmask)]
    assert np.abs(djac @ mat).max() < 1e-9 * np.abs(djac).max()                           # they are null vectors
    assert all(abs(np.mean(np.abs(v)) - 1.0) < 1e-12 for v in a.null_vec)                 # add_matrix! scaling
