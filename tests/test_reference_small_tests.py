"""The reference's small unit tests for types on the path that tests/test_oracle_golden.py does not already restate:
test/constraints/game_constraints.jl, test/struct/{regularizer,problem_size,options,statistics}.jl,
test/problem/{problem,global_quantities}.jl — on the oracle, and on the host package where it mirrors the type."""
import numpy as np

import algames_b200 as ab
import oracle.algames_oracle as O


def _lq_problem(seed=0, constraints=False):
    # test/problem/problem.jl:12-29 and test/problem/global_quantities.jl:4-21: 3-player unicycle, N = 10, random weights
    rng = np.random.default_rng(seed)
    N, dt, p = 10, 0.1, 3
    model = O.make_model("unicycle", p)
    ps = O.ProblemSize(N, model)
    obj = O.GameObjective([rng.random(4) for _ in range(p)], [rng.random(2) for _ in range(p)],
                          [(i + 1) * np.ones(4) for i in range(p)], [2 * (i + 1) * np.ones(2) for i in range(p)], N, model)
    gc = O.GameConstraintValues(ps)
    if constraints:
        gc.add_control_bound(0.1 * np.ones(model.m), -0.1 * np.ones(model.m))
        gc.add_wall_constraint([O.Wall(np.array([0.0, 1.0]), np.array([1.0, 0.0]), np.array([1.0, 1.0]) / np.sqrt(2))])
    return O.GameProblem(N, dt, rng.random(model.n), model, O.Options(), obj, gc)


def test_game_constraints_set_constraint_params():
    # test/constraints/game_constraints.jl:3-39
    N, p = 10, 3
    model = O.make_model("unicycle", p)
    gc = O.GameConstraintValues(O.ProblemSize(N, model))
    assert gc.probsize.p == model.p and len(gc.state_conval) == model.p
    rng = np.random.default_rng(1)
    gc.add_control_bound(rng.random(model.m), -rng.random(model.m))
    gc.add_collision_avoidance(1.0)
    opts = O.Options()
    opts.rho_increase, opts.rho_0, opts.rho_max, opts.lambda_max = 2.0, 3.0, 4.0, 5.0
    gc.set_constraint_params(opts)
    assert gc.alpha_dual == opts.alpha_dual and gc.alphax_dual == opts.alphax_dual[:p]
    assert gc.active_set_tolerance == opts.active_set_tolerance
    for cv in (gc.state_conval[p - 1][0], gc.control_conval[0]):
        assert (cv.phi, cv.mu0, cv.mu_max, cv.lam_max) == (2.0, 3.0, 4.0, 5.0)


def test_regularizer_set_and_mult():
    # test/struct/regularizer.jl:3-16
    reg = O.Regularizer()
    reg.set(1e-1)
    assert reg.x == reg.u == reg.lam == 1e-1
    reg.mult(1e-1)
    assert reg.x == reg.u == reg.lam == 1e-1 ** 2


def test_problem_size_equality():
    # test/struct/problem_size.jl:3-9: the 3-player unicycle and the 3-player planar double integrator share every size
    for mk, PS in ((lambda name: O.make_model(name, 3), O.ProblemSize),):
        assert PS(10, mk("unicycle")) == PS(10, mk("double_integrator"))
        assert not (PS(10, mk("unicycle")) == PS(11, mk("unicycle")))
        assert not (PS(10, O.make_model("unicycle", 2)) == PS(10, mk("unicycle")))
    a, b = ab.ProblemSize(10, ab.UnicycleGame(p=3)), ab.ProblemSize(10, ab.DoubleIntegratorGame(p=3, d=2))
    assert vars(a) == vars(b) and a.S == O.ProblemSize(10, O.make_model("unicycle", 3)).S == 12 * 3 * 9 + 6 * 9 + 12 * 9


def test_options_defaults_agree_between_oracle_and_host():
    # test/struct/options.jl + the defaults of src/struct/options.jl:5-116
    o, h = O.Options(), ab.Options()
    expect = dict(reg_0=1e-3, ls_iter=25, beta=0.01, alpha_decrease=0.5, delta_min=1e-9, rho_0=1.0, rho_increase=10.0,
                  rho_max=1e7, lambda_max=1e7, alpha_dual=1.0, active_set_tolerance=1e-4, eps_dyn=1e-3, eps_sta=1e-3,
                  eps_con=1e-3, eps_opt=1e-3, outer_iter=7, inner_iter=20, amplitude_init=1e-8, shift=2 ** 10, seed=100,
                  dual_reset=True, regularize=True)
    for k, v in expect.items():
        assert getattr(o, k) == v and getattr(h, k) == v, k
    assert o.alphax_dual == h.alphax_dual == [1.0] * 10


def test_problem_constructor_pushes_options_into_constraints():
    # test/problem/problem.jl:12-29 + problem.jl:48 (set_constraint_params! in the constructor)
    prob = _lq_problem(constraints=True)
    assert isinstance(prob, O.GameProblem) and prob.probsize.S == len(prob.core.res)
    cv = prob.game_con.control_conval[0]
    assert (cv.phi, cv.mu0, cv.mu_max, cv.lam_max) == (10.0, 1.0, 1e7, 1e7)


def test_global_quantities_run_and_are_consistent():
    # test/problem/global_quantities.jl:4-28: residual!, ibr_residual!, residual_jacobian!, ibr_residual_jacobian! on a random
    # unconstrained LQ unicycle game; beyond "runs": the best-response quantities are the masked full-game ones
    prob = _lq_problem(seed=2)
    i = 2
    res = O.residual(prob, prob.pdtraj).copy()
    J = O.residual_jacobian(prob, prob.pdtraj).copy()
    assert res.shape == (prob.probsize.S,) and J.shape == (prob.probsize.S,) * 2 and np.isfinite(J).all()
    vm, hm = O.vertical_mask(prob.core, i), O.horizontal_mask(prob.core, i)
    ires = O.ibr_residual(prob, prob.pdtraj, i).copy()
    iJ = O.ibr_residual_jacobian(prob, prob.pdtraj, i)
    assert np.allclose(ires[vm], res[vm], rtol=0, atol=1e-14)
    assert np.allclose(iJ[np.ix_(vm, hm)], J[np.ix_(vm, hm)], rtol=0, atol=1e-14)


def test_statistics_record_and_reset():
    # test/struct/statistics.jl:3-66: record!(stats, prob, …) appends one entry (outer index, Δ_traj, violations); the
    # best-response form (:54-60) does the same for player i; reset! empties the history
    prob = _lq_problem(seed=3, constraints=True)
    O.record(prob, prob.pdtraj, 0.1, 2)
    assert len(prob.stats) == 1 and prob.stats[0].outer == 2 and prob.stats[0].delta == 0.1
    O.record_ibr(prob, prob.pdtraj, 0.1, 3, 2)
    assert [r.outer for r in prob.stats] == [2, 3] and [r.delta for r in prob.stats] == [0.1, 0.1]
    assert all(np.isfinite([r.res, r.dyn, r.con, r.sta, r.opt]).all() for r in prob.stats)
    # host mirror (fields of Statistics; t_elap is not kept — see the class docstring)
    st = ab.problem.Statistics()
    st.record([0.3, 0.0, 0.0, 0.0, 0.0, 0.1, 1, 1, 1, 0])
    st.record([0.3, 0.0, 0.0, 0.0, 0.0, 0.1, 2, 2, 2, 0])
    assert st.iter == 2 and st.outer_iter == [1, 2] and st.delta == [0.1, 0.1] and st.res == [0.3, 0.3]
    st.reset()
    assert st.iter == 0 and st.outer_iter == [] and st.dyn_vio == []
