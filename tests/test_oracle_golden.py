"""Pins the CPU oracle against every known-answer vector the reference's own tests hold
for the hot path.  Each test cites the reference test file:line it restates.
CPU only."""
import math

import numpy as np
import pytest

from oracle import algames_oracle as O

ones = lambda *s: np.ones(s)


# ---------------------------------------------------------------- models (test/dynamics/*.jl)
def test_model_index_sets():
    # test/dynamics/double_integrator.jl, unicycle.jl, bicycle.jl: pu/px/pz are component-major
    for mdl in (O.DoubleIntegratorGame(p=3, d=2), O.UnicycleGame(p=3), O.BicycleGame(p=3)):
        assert (mdl.n, mdl.m, mdl.p) == (12, 6, 3)
        assert [list(v + 1) for v in mdl.pu] == [[1, 4], [2, 5], [3, 6]]
        assert [list(v + 1) for v in mdl.px] == [[1, 4], [2, 5], [3, 6]]
        assert [list(v + 1) for v in mdl.pz] == [[1, 4, 7, 10], [2, 5, 8, 11], [3, 6, 9, 12]]
    di3 = O.DoubleIntegratorGame(p=2, d=3)
    assert (di3.n, di3.m) == (12, 6) and list(di3.pz[0] + 1) == [1, 3, 5, 7, 9, 11]


def test_dynamics_values():
    # src/dynamics/double_integrator.jl:27-31, unicycle.jl:27-32, bicycle.jl:28-41
    x = np.arange(1.0, 13.0) / 10
    u = np.arange(1.0, 7.0) / 7
    di = O.DoubleIntegratorGame(p=3)
    assert np.allclose(di.f(x, u), np.concatenate([x[6:], u]))
    un = O.UnicycleGame(p=3)
    f = un.f(x, u)
    assert np.isclose(f[1], math.cos(x[7]) * x[10]) and np.isclose(f[4], math.sin(x[7]) * x[10])
    assert np.allclose(f[6:], u)
    bi = O.BicycleGame(p=3)
    f = bi.f(x, u)
    beta = math.atan2(0.05 * math.tan(u[4]), 0.1)
    assert np.isclose(f[1], x[7] * math.cos(beta + x[10])) and np.isclose(f[4], x[7] * math.sin(beta + x[10]))
    assert np.isclose(f[7], u[1]) and np.isclose(f[10], x[7] * math.sin(beta) / 0.05)


@pytest.mark.parametrize("mdl", [O.DoubleIntegratorGame(p=3), O.UnicycleGame(p=3), O.BicycleGame(p=2)])
def test_rk2_jacobian_matches_finite_differences(mdl):
    # test/problem/local_quantities.jl:24-57: discrete_jacobian!(RK2) == d/dx, d/du of discrete_dynamics(RK2)
    rng = np.random.default_rng(3)
    x, u, dt = rng.random(mdl.n), rng.random(mdl.m), 0.2
    A, B = O.rk2_jacobian(mdl, x, u, dt)
    h = 1e-6
    for j in range(mdl.n):
        e = np.zeros(mdl.n); e[j] = h
        assert np.allclose((O.rk2(mdl, x + e, u, dt) - O.rk2(mdl, x - e, u, dt)) / (2 * h), A[:, j], atol=1e-8)
    for j in range(mdl.m):
        e = np.zeros(mdl.m); e[j] = h
        assert np.allclose((O.rk2(mdl, x, u + e, dt) - O.rk2(mdl, x, u - e, dt)) / (2 * h), B[:, j], atol=1e-8)


def test_rk2_close_to_euler():
    # test/problem/local_quantities.jl:4-14
    mdl = O.DoubleIntegratorGame(p=3, d=2)
    rng = np.random.default_rng(0)
    x, u, dt = rng.random(mdl.n), rng.random(mdl.m), 0.01
    assert np.abs(O.rk2(mdl, x, u, dt) - (x + dt * mdl.f(x, u))).sum() < 1e-3


# ---------------------------------------------------------------- index maps (test/core/newton_core.jl)
def test_vertical_horizontal_indices():
    mdl = O.UnicycleGame(p=2)
    ps = O.ProblemSize(3, mdl)
    core = O.NewtonCore(ps)
    n, mi, m = ps.n, ps.mi, ps.m
    r = lambda a, b: list(range(a, b))
    # test/core/newton_core.jl:12-16 (1-based there)
    assert list(core.vert[("opt", 1, "x", 2)]) == r(0, n)
    assert list(core.vert[("opt", 1, "u", 1)]) == r(n, n + mi[0])
    assert list(core.vert[("opt", 1, "x", 3)]) == r(n + mi[0], 2 * n + mi[0])
    assert list(core.vert[("opt", 1, "u", 2)]) == r(2 * n + mi[0], 2 * n + 2 * mi[0])
    assert list(core.vert[("opt", 2, "x", 2)]) == r(2 * n + 2 * mi[0], 3 * n + 2 * mi[0])
    # :54-59
    assert list(core.horiz[("x", 2)]) == r(0, n)
    assert list(core.horiz[("u", 1, 1)]) == r(n, n + mi[0])
    assert list(core.horiz[("u", 2, 1)]) == r(n + mi[0], n + m)
    assert list(core.horiz[("λ", 1, 1)]) == r(n + m, 2 * n + m)
    assert list(core.horiz[("λ", 2, 1)]) == r(2 * n + m, 3 * n + m)
    assert list(core.horiz[("x", 3)]) == r(3 * n + m, 4 * n + m)
    # :18-41, 61-84: both maps are permutations of 1..S
    assert sorted(np.concatenate(list(core.vert.values()))) == r(0, ps.S)
    assert sorted(np.concatenate(list(core.horiz.values()))) == r(0, ps.S)
    assert ps.S == n * 2 * 2 + m * 2 + n * 2      # problem_size.jl:22


def test_stamp_validity_table():
    # test/core/stamp.jl:8-104 (representative rows of the validity table)
    N, p = 10, 3
    assert O.valid_v("opt", 1, "x", 1, 2, N, p) and not O.valid_v("opt", 1, "x", 1, 1, N, p)
    assert O.valid_v("opt", 2, "u", 2, 1, N, p) and not O.valid_v("opt", 2, "u", 1, 1, N, p)
    assert not O.valid_v("opt", 2, "u", 2, N, N, p)
    assert O.valid_v("dyn", 1, "x", 1, N - 1, N, p) and not O.valid_v("dyn", 1, "x", 1, N, N, p)
    assert not O.valid_v("dyn", 2, "x", 1, 1, N, p)
    assert O.valid_h("x", 1, N, N, p) and not O.valid_h("x", 1, 1, N, p) and not O.valid_h("x", 2, 2, N, p)
    assert O.valid_h("λ", 3, N - 1, N, p) and not O.valid_h("λ", 3, N, N, p)
    assert O.valid("opt", 1, "x", 1, 2, "λ", 1, 1, N, p) and not O.valid("opt", 1, "x", 1, 2, "λ", 2, 1, N, p)
    assert O.valid("dyn", 1, "x", 1, 1, "u", 3, 1, N, p) and not O.valid("dyn", 1, "x", 1, 1, "λ", 1, 1, N, p)


# ---------------------------------------------------------------- constraints
def test_control_bound_constraint_values():
    # test/constraints/control_bound_constraint.jl:8-15
    U = np.array([13.0, 1.0, -12.0, 1.0, 2.0, 30.0])
    u_max = [np.inf, np.inf, -11.0, 15.0, 2.0, 30.0]
    u_min = [-np.inf, -10.0, -np.inf, 1.0, -2.0, -30.0]
    con = O.ControlBoundConstraint(6, u_max, u_min)
    assert list(con.evaluate(None, U)) == [-1.0, -14.0, 0.0, 0.0, -11.0, 0.0, -4.0, -60.0]
    assert list(con.inds + 1) == [3, 4, 5, 6, 8, 10, 11, 12]
    J = con.jacobian(None, U)
    assert J.shape == (8, 6) and J[0, 2] == 1 and J[4, 1] == -1 and np.abs(J).sum() == 8


def test_state_bound_constraint_values():
    # test/constraints/state_bound_constraint.jl:8-15 (same numbers on the state)
    X = np.array([13.0, 1.0, -12.0, 1.0, 2.0, 30.0])
    con = O.StateBoundConstraint(6, [np.inf, np.inf, -11.0, 15.0, 2.0, 30.0], [-np.inf, -10.0, -np.inf, 1.0, -2.0, -30.0])
    assert list(con.evaluate(X, None)) == [-1.0, -14.0, 0.0, 0.0, -11.0, 0.0, -4.0, -60.0]


def test_wall_constraint_values():
    # test/constraints/wall_constraint.jl:4-30
    s2 = math.sqrt(2)
    X = np.array([13.0, 1.0, -12.0, 1.0])
    x1 = [0.0, 0.0, 1.0, 3.0, -2.0]; y1 = [1.0, -1.0, 2.0, 2.0, 0.0]
    x2 = [1.0, 1.0, 2.0, 2.0, 0.0]; y2 = [0.0, 0.0, 1.0, 1.0, 0.0]
    xv = np.array([1.0, 1.0, 1.0, 1.0, 0.0]) / s2; yv = np.array([1.0, -1.0, 1.0, -1.0, s2]) / s2
    con = O.WallConstraint(4, x1, y1, x2, y2, xv, yv, x=3, y=1)      # x=4, y=2 one-based
    assert np.abs(con.evaluate(X, None) - np.array([s2 / 2, 0.0, -s2 / 2, 0.0, 0.0])).sum() < 1e-10
    J = con.jacobian(X, None)
    h = 1e-7
    for j in range(4):
        e = np.zeros(4); e[j] = h
        fd = (con.evaluate(X + e, None) - con.evaluate(X - e, None)) / (2 * h)
        assert np.allclose(fd, J[:, j], atol=1e-6)


def test_collision_avoidance_pairs_and_circle_indices():
    # test/constraints/constraints_methods.jl:4-28 and :76-98
    mdl = O.DoubleIntegratorGame(p=3)
    gc = O.GameConstraintValues(O.ProblemSize(20, mdl))
    gc.add_collision_avoidance(1.0)
    pu = mdl.pu
    for i, others in enumerate([[1, 2], [0, 2], [0, 1]]):
        assert [list(cv.con.x1) for cv in gc.state_conval[i]] == [list(pu[i])] * 2
        assert [list(cv.con.x2) for cv in gc.state_conval[i]] == [list(pu[j]) for j in others]
        assert all(cv.con.radius == 2.0 for cv in gc.state_conval[i])      # r_i + r_j, :27-29
        assert all(cv.inds == list(range(2, 21)) for cv in gc.state_conval[i])
    gc = O.GameConstraintValues(O.ProblemSize(20, mdl))
    gc.add_circle_constraint([1.0, 2, 3, 4, 5], [-1.0, -2, -3, -4, -5], [0.1, 0.2, 0.3, 0.4, 0.5])
    for i in range(3):
        assert gc.state_conval[i][0].con.xi == mdl.px[i][0] and gc.state_conval[i][0].con.yi == mdl.px[i][1]


def test_al_expansion():
    # test/constraints/constraint_derivatives.jl:4-34
    N, dt = 10, 0.1
    mdl = O.UnicycleGame(p=3)
    ps = O.ProblemSize(N, mdl)
    pd = O.PrimalDualTraj(ps, dt, f=ones, amplitude=0.1)
    gc = O.GameConstraintValues(ps)
    gc.add_control_bound(np.ones(mdl.m), -np.ones(mdl.m))
    cv = gc.control_conval[0]
    for k in range(N - 2):
        cv.lam[k] = (k + 1) * np.ones(2 * mdl.m)
    cv.evaluate(pd.X, pd.U)
    assert np.array_equal(cv.vals[0], np.concatenate([-0.9 * np.ones(mdl.m), -1.1 * np.ones(mdl.m)]))
    assert np.array_equal(cv.vals[-1], cv.vals[0])
    cv.jacobian(pd.X, pd.U)
    assert np.array_equal(cv.jac[0], np.vstack([np.eye(mdl.m), -np.eye(mdl.m)]))
    cv.cost_expansion()
    for j in (0, -1):
        Irho = np.diag(((cv.vals[j] >= 0) | (cv.lam[j] > 0)) * cv.mu[j][0])
        assert np.array_equal(cv.grad[j], cv.jac[j].T @ cv.lam[j] + cv.jac[j].T @ Irho @ cv.vals[j])
        assert np.array_equal(cv.hess[j], cv.jac[j].T @ Irho @ cv.jac[j])
    # last knot has λ = 0 and c < 0 ⇒ inactive ⇒ zero expansion; first knot has λ > 0 ⇒ active
    assert not cv.hess[-1].any() and cv.hess[0].any()


def test_penalty_and_dual_updates():
    # test/constraints/constraints_methods.jl:135-229
    mdl = O.DoubleIntegratorGame(p=3)
    ps = O.ProblemSize(20, mdl)
    gc = O.GameConstraintValues(ps)
    gc.add_control_bound(10 * np.ones(mdl.m), -10 * np.ones(mdl.m))
    gc.add_circle_constraint([1.0, 2, 3, 4, 5], [-1.0, -2, -3, -4, -5], [0.1, 0.2, 0.3, 0.4, 0.5])
    opts = O.Options(rho_0=1e-3, rho_increase=1e1, rho_max=1e-1, lambda_max=1e1)
    gc.set_constraint_params(opts)
    cc, sc = gc.control_conval[0], gc.state_conval[0][0]
    assert (cc.mu0, sc.mu0, cc.phi, sc.phi, cc.mu_max, sc.mu_max) == (1e-3, 1e-3, 10, 10, 0.1, 0.1)
    gc.reset_penalties()
    assert np.array_equal(cc.mu[0], 1e-3 * np.ones(12)) and np.array_equal(sc.mu[0], 1e-3 * np.ones(5))
    gc.penalty_update()
    assert np.allclose(cc.mu[0], 1e-2, rtol=1e-15) and np.allclose(sc.mu[0], 1e-2, rtol=1e-15)
    gc.penalty_update()
    assert np.allclose(cc.mu[0], 1e-1, rtol=1e-15)
    for _ in range(4):
        gc.penalty_update()
    assert np.array_equal(cc.mu[0], 1e-1 * np.ones(12)) and np.array_equal(sc.mu[0], 1e-1 * np.ones(5))   # clamped
    gc.reset_penalties()
    assert np.array_equal(cc.mu[0], 1e-3 * np.ones(12))
    # duals
    gc.reset_duals()
    gc.dual_update()
    assert not cc.lam.any() and not sc.lam.any()
    pd = O.PrimalDualTraj(ps, 0.1)
    O.init_traj(pd, np.zeros(mdl.n), lambda k: np.ones(k), 1e2)
    gc.evaluate(pd.X, pd.U)
    assert np.array_equal(cc.vals[0], np.concatenate([90 * np.ones(6), -110 * np.ones(6)]))
    gc.dual_update()
    assert np.allclose(cc.lam[0], 1e-3 * np.concatenate([90 * np.ones(6), np.zeros(6)]), rtol=1e-15)
    assert np.allclose(sc.lam[0], 1e-3 * np.maximum(0, sc.vals[0]), rtol=1e-15)
    O.init_traj(pd, np.zeros(mdl.n), lambda k: np.ones(k), 1e5)
    gc.evaluate(pd.X, pd.U)
    assert np.array_equal(cc.vals[0], np.concatenate([(1e5 - 10) * np.ones(6), -(1e5 + 10) * np.ones(6)]))
    gc.dual_update()
    assert np.array_equal(cc.lam[0], np.concatenate([opts.lambda_max * np.ones(6), np.zeros(6)]))   # saturation
    gc.reset_duals()
    assert not cc.lam.any() and not sc.lam.any()


def test_active_set_predicate():
    # test/active_set/active_set_methods.jl:16-34: active = (c >= -tol) | (λ > 0)
    mdl = O.DoubleIntegratorGame(p=2)
    ps = O.ProblemSize(5, mdl)
    gc = O.GameConstraintValues(ps)
    gc.add_control_bound(np.ones(mdl.m), -np.ones(mdl.m))
    cv = gc.control_conval[0]
    cv.vals[:] = -1e-3
    cv.vals[0, 0] = -5e-5
    cv.lam[1, 1] = 0.2
    a = gc.active_set(tol=1e-4)[0]
    assert a[0, 0] and a[1, 1] and a.sum() == 2


# ---------------------------------------------------------------- objective
def test_lqr_expansion_dt_scaling():
    # test/objective/objective.jl:10-64
    N, dt = 10, 0.1
    mdl = O.UnicycleGame(p=3)
    rng = np.random.default_rng(1)
    Q = [rng.random(4) for _ in range(3)]; R = [rng.random(2) for _ in range(3)]
    xf = [(i + 1) * np.ones(4) for i in range(3)]; uf = [2 * (i + 1) * np.ones(2) for i in range(3)]
    obj = O.GameObjective(Q, R, xf, uf, N, mdl)
    # zero cost at the targets (:23-32): gradient vanishes there
    X = np.array([1.0, 2, 3] * 4); U = np.array([2.0, 4, 6] * 2)
    for i in range(3):
        q, r = obj.gradient(i, 1, X, U, dt)
        assert np.abs(q).max() < 1e-14 and np.abs(r).max() < 1e-14
    x, u = 10 * rng.random(12), 10 * rng.random(6)
    Qj, Rj = np.zeros(12), np.zeros(6); Qj[mdl.pz[0]] = Q[0]; Rj[mdl.pu[0]] = R[0]
    xfj, ufj = np.zeros(12), np.zeros(6); xfj[mdl.pz[0]] = 1; ufj[mdl.pu[0]] = 2
    q, r = obj.gradient(0, 1, x, u, dt)
    assert np.abs(q - Qj * (x - xfj) * dt).sum() < 1e-10 and np.abs(r - Rj * (u - ufj) * dt).sum() < 1e-10
    q, r = obj.gradient(0, N, x, u, dt)
    assert np.abs(q - Qj * (x - xfj)).sum() < 1e-10 and np.abs(r).sum() < 1e-10
    Qh, Rh = obj.hessian(0, 1, x, u, dt)
    assert np.abs(Qh - np.diag(Qj) * dt).sum() < 1e-10 and np.abs(Rh - np.diag(Rj) * dt).sum() < 1e-10
    Qh, Rh = obj.hessian(0, N, x, u, dt)
    assert np.abs(Qh - np.diag(Qj)).sum() < 1e-10 and np.abs(Rh).sum() < 1e-10


def test_collision_cost_values_and_derivatives():
    # test/objective/objective.jl:113-201
    mdl = O.DoubleIntegratorGame(p=2)
    mu, r = 10.0, 0.2
    pxi, pxj = mdl.px[0], mdl.px[1]
    x = np.array([1.0, 1.1, 2.0, 2.0, 0, 0, 0, 0])
    assert abs(O.collision_stage_cost(mu, r, pxi, pxj, x) - 0.05) < 1e-10        # :141
    assert abs(O.collision_stage_cost(mu, r, pxi, pxj, x) * 0.1 - 0.005) < 1e-10  # :147 (dt-scaled)
    rng = np.random.default_rng(5)
    z = rng.random(8)
    cost = lambda xx, rr: O.collision_stage_cost(mu, rr, pxi, pxj, xx)
    for rr in (1e3, 1e-3):          # active / inactive (:166-167)
        g = O._collision_cost_grad(mu, rr, pxi, pxj, z)
        H = O._collision_cost_hess(mu, rr, pxi, pxj, z)
        h = 1e-5
        gfd = np.array([(cost(z + h * e, rr) - cost(z - h * e, rr)) / (2 * h) for e in np.eye(8)])
        assert np.abs(g - gfd).sum() / max(np.linalg.norm(g), 1e-30) < 1e-6 or np.abs(g - gfd).sum() < 1e-7
        Hfd = np.array([(O._collision_cost_grad(mu, rr, pxi, pxj, z + h * e) -
                         O._collision_cost_grad(mu, rr, pxi, pxj, z - h * e)) / (2 * h) for e in np.eye(8)])
        assert np.abs(H - Hfd).sum() < 1e-2                                        # :188, :193
    obj = O.GameObjective([np.ones(4)] * 2, [np.ones(2)] * 2, [np.zeros(4)] * 2, [np.zeros(2)] * 2, 10, mdl)
    obj.add_collision_cost([1.0, 2.0], [10.0, 20.0])
    assert [(c[0], c[1]) for c in obj.collision[0]] == [(10.0, 1.0)] and [(c[0], c[1]) for c in obj.collision[1]] == [(20.0, 2.0)]


# ---------------------------------------------------------------- trajectories (test/struct/primal_dual_traj.jl)
def test_traj_ops():
    N, dt = 10, 0.2
    mdl = O.UnicycleGame(p=3)
    ps = O.ProblemSize(N, mdl)
    n, m, p = mdl.n, mdl.m, mdl.p
    rng = np.random.default_rng(0)
    pd = O.PrimalDualTraj(ps, dt)
    assert pd.X.shape == (N, n) and pd.U.shape == (N, m) and pd.du.shape == (p, N - 1, n)
    x0 = rng.random(n)
    O.init_traj(pd, x0, lambda k: np.ones(k), 10.0)                    # :19-26
    assert np.array_equal(pd.X[0], x0) and np.array_equal(pd.X[1], 10 * np.ones(n))
    assert np.array_equal(pd.U[0], 10 * np.ones(m)) and np.array_equal(pd.du[0, 0], 10 * np.ones(n))
    core = O.NewtonCore(ps)
    d = O.PrimalDualTraj(ps, dt)
    dtraj = rng.random(ps.S)
    O.init_traj(d, x0, lambda k: np.ones(k), 10.0)
    O.set_traj(core, d, dtraj)                                         # :45-63
    assert np.array_equal(d.X[0], x0)
    assert np.array_equal(d.X[1], dtraj[core.horiz[("x", 2)]]) and np.array_equal(d.X[-1], dtraj[core.horiz[("x", N)]])
    assert np.array_equal(d.U[0][mdl.pu[1]], dtraj[core.horiz[("u", 2, 1)]])
    assert np.array_equal(d.U[1][mdl.pu[2]], dtraj[core.horiz[("u", 3, 2)]])
    assert np.array_equal(d.du[2, -1], dtraj[core.horiz[("λ", 3, N - 1)]])
    assert np.array_equal(O.get_traj(core, d), dtraj)                  # :66-83 round trip
    tgt, src, dl = (O.PrimalDualTraj(ps, dt) for _ in range(3))
    O.init_traj(tgt, x0, lambda k: np.ones(k), 0.0)
    O.init_traj(src, x0, lambda k: np.ones(k), 10.0)
    O.init_traj(dl, x0, lambda k: np.ones(k), 100.0)
    O.update_traj(tgt, src, 0.5, dl)                                   # :94-107
    assert np.array_equal(tgt.X[0], x0) and np.array_equal(tgt.X[1:], 60 * np.ones((N - 1, n)))
    assert np.array_equal(tgt.U[: N - 1], 60 * np.ones((N - 1, m))) and np.array_equal(tgt.du, 60 * np.ones((p, N - 1, n)))
    O.init_traj(pd, 1e3 * np.ones(n), lambda k: np.ones(k), 10.0)
    assert O.delta_step(pd, 0.5) == 10.0 * 0.5                         # :119-123


def test_init_traj_shift():
    # src/struct/primal_dual_traj.jl:34-41 with s=1 (MPC warm start)
    mdl = O.DoubleIntegratorGame(p=2)
    ps = O.ProblemSize(5, mdl)
    pd = O.PrimalDualTraj(ps, 0.1)
    pd.X[:] = np.arange(5)[:, None]; pd.U[:] = 10 + np.arange(5)[:, None]; pd.du[:] = 20 + np.arange(4)[None, :, None]
    O.init_traj(pd, -np.ones(8), lambda k: np.ones(k), 7.0, s=1)
    assert list(pd.X[:, 0]) == [-1, 2, 3, 4, 7] and list(pd.U[:, 0]) == [11, 12, 13, 14, 7]
    assert list(pd.du[0, :, 0]) == [21, 22, 23, 7]


# ---------------------------------------------------------------- violations (test/struct/violations.jl)
def _empty_problem(mdl, N, dt, gc=None):
    ps = O.ProblemSize(N, mdl)
    p = mdl.p
    obj = O.GameObjective([np.ones(4)] * p, [np.ones(2)] * p, [np.zeros(4)] * p, [np.zeros(2)] * p, N, mdl)
    return O.GameProblem(N, dt, np.zeros(mdl.n), mdl, O.Options(), obj, gc or O.GameConstraintValues(ps))


def test_violations():
    N, dt = 10, 0.1
    mdl = O.UnicycleGame(p=3)
    ps = O.ProblemSize(N, mdl)
    prob = _empty_problem(mdl, N, dt)
    pd = O.PrimalDualTraj(ps, dt)
    O.init_traj(pd, np.zeros(mdl.n), lambda k: np.zeros(k), 1.0)
    assert O.dynamics_violation(prob, pd) == 0.0                                     # :13-17
    O.init_traj(pd, np.ones(mdl.n), lambda k: np.ones(k), 1.0)
    assert abs(O.dynamics_violation(prob, pd) - np.abs(O.dynamics_residual(mdl, pd, 1)).max()) < 1e-10
    gc = O.GameConstraintValues(ps)
    gc.add_control_bound(0.1 * np.ones(mdl.m), -0.1 * np.ones(mdl.m))
    prob = _empty_problem(mdl, N, dt, gc)
    O.init_traj(pd, np.zeros(mdl.n), lambda k: np.ones(k), 1.0)
    assert O.control_violation(prob, pd) == 0.9                                      # :24-31
    gc = O.GameConstraintValues(ps)
    gc.add_wall_constraint([O.Wall(np.array([0.0, 1]), np.array([1.0, 0]), np.array([1.0, 1]) / math.sqrt(2))])
    prob = _empty_problem(mdl, N, dt, gc)
    assert abs(O.state_violation(prob, pd) - math.sqrt(2) / 2) < 1e-10               # :38-45
    core = O.NewtonCore(ps)
    assert O.optimality_violation(core) == 0.0                                        # :53-56
    core.res[core.vert[("opt", 2, "u", 5)]] += 1e2
    assert O.optimality_violation(core) == 1e2                                        # :62-68


# ---------------------------------------------------------------- solver (test/problem/solver_methods.jl)
def _solver_case(model, p, constrained, outer, inner):
    N, dt = 20, 0.1
    mdl = O.make_model(model, p)
    ps = O.ProblemSize(N, mdl)
    x0 = {1: [1.0, 1.0, 0.0, 0.9], 2: [1.0, 2.0, 1.0, 2.0, 0, 0, 0.9, 0.9]}[p]
    if constrained:
        x0 = [1.0, 2.0, 1.1, 2.0, 0, 0, 0.9, 0.9]
    opts = O.Options(outer_iter=outer, inner_iter=inner, ls_iter=25, reg_0=1e-7, eps_dyn=1e-10, eps_opt=1e-10)
    obj = O.GameObjective([np.ones(4)] * p, [0.5 * np.ones(2)] * p, [np.zeros(4)] * p, [-np.ones(2)] * p, N, mdl)
    gc = O.GameConstraintValues(ps)
    if constrained:                                           # :147-160
        gc.add_collision_avoidance(0.05)
        gc.add_control_bound(np.ones(mdl.m), -np.ones(mdl.m))
        gc.add_circle_constraint([1.5, 0.2, 0.3], [1.25, 0.2, 0.3], [0.2, 0.2, 0.3])
    return O.GameProblem(N, dt, x0, mdl, opts, obj, gc)


@pytest.mark.parametrize("model,p,outer,inner", [
    ("double_integrator", 1, 1, 1),     # :6-34   LQ, one Newton step
    ("unicycle", 1, 7, 20),             # :36-65
    ("double_integrator", 2, 1, 1),     # :68-97  2-player LQ game, one Newton step
    ("unicycle", 2, 7, 20),             # :100-129
])
def test_solver_unconstrained(model, p, outer, inner):
    prob = O.newton_solve(_solver_case(model, p, False, outer, inner))
    assert np.abs(prob.core.res).sum() / prob.probsize.S < 1e-6
    assert O.dynamics_violation(prob, prob.pdtraj) < 1e-6
    if outer == 1:
        assert prob.n_newton == 1


def test_solver_constrained_two_player_unicycle():
    # :132-182 — runs with the *previous* block's opts (the rebinding at :164 never reaches prob)
    prob = O.newton_solve(_solver_case("unicycle", 2, True, 7, 20))
    last = prob.stats[-1]
    assert np.abs(prob.core.res).sum() / prob.probsize.S < 1e-3
    assert last.dyn < 1e-3 and last.sta < 1e-3 and last.con < 1e-3 and last.opt < 1e-3


def test_dense_jacobian_matches_finite_difference_of_residual_for_lq_game():
    # For an unconstrained LQ game the Gauss-Newton Jacobian is the exact derivative of residual!
    # (global_quantities.jl:109-174 vs :9-65) — checks the block placement of every stamp.
    prob = _solver_case("double_integrator", 2, False, 1, 1)
    prob.N = 6
    mdl = prob.model
    prob = O.GameProblem(6, 0.1, prob.x0, mdl, O.Options(), O.GameObjective([np.ones(4)] * 2, [0.5 * np.ones(2)] * 2, [np.zeros(4)] * 2, [-np.ones(2)] * 2, 6, mdl), O.GameConstraintValues(O.ProblemSize(6, mdl)))
    rng = np.random.default_rng(2)
    pd = prob.pdtraj
    pd.X[:] = rng.random(pd.X.shape); pd.U[:] = rng.random(pd.U.shape); pd.du[:] = rng.random(pd.du.shape)
    J = O.residual_jacobian(prob, pd, regularize=False)
    v0 = O.get_traj(prob.core, pd)
    r0 = O.residual(prob, pd).copy()
    for c in rng.choice(prob.probsize.S, 12, replace=False):
        d = pd.copy()
        v = v0.copy(); v[c] += 1e-6
        O.set_traj(prob.core, d, v)
        assert np.allclose((O.residual(prob, d) - r0) / 1e-6, J[:, c], atol=1e-6)


# ---------------------------------------------------------------- iterative best response
def test_player_masks():
    # test/core/newton_core.jl:118-160
    N, p = 6, 5
    mdl = O.UnicycleGame(p=p)
    ps = O.ProblemSize(N, mdl)
    core = O.NewtonCore(ps)
    n, ni, mi = ps.n, 4, 2
    for fn in (O.vertical_mask, O.horizontal_mask):
        split = []
        for i in range(1, p + 1):
            assert len(fn(core, i)) == (N - 1) * (2 * n + mi)
            ms = fn(core, i, splitted_state=True)
            assert len(ms) == (N - 1) * (2 * ni + mi)
            split.append(set(ms.tolist()))
        assert not set.intersection(*split)


@pytest.mark.parametrize("model,p,outer,inner,tol", [
    ("double_integrator", 1, 1, 1, 1e-6),      # test/problem/solver_methods.jl:187-218
    ("unicycle", 1, 7, 20, 1e-6),              # :220-250
    ("double_integrator", 2, 1, 1, 5e-2),      # :252-282
    ("unicycle", 2, 7, 20, 5e-2),              # :284-314
])
def test_ibr_solver(model, p, outer, inner, tol):
    prob = _solver_case(model, p, False, outer, inner)
    if p == 1:
        rng = np.random.default_rng(100)
        O.init_traj(prob.pdtraj, prob.x0, lambda k: rng.random(k), prob.opts.amplitude_init)
        prob.pdtraj_trial = prob.pdtraj.copy()
        O.rollout_rk3(prob.model, prob.pdtraj)
        prob.stats, prob.n_newton = [], 0
        O.ibr_newton_solve_player(prob, 1)
        O.residual(prob, prob.pdtraj)
    else:
        O.ibr_newton_solve(prob)
    assert np.abs(prob.core.res).sum() / prob.probsize.S < tol
    assert O.dynamics_violation(prob, prob.pdtraj) < 1e-6


def test_ibr_masked_system_is_the_single_player_problem():
    # the masked Jacobian of player i is exactly the KKT Jacobian of the 1-player problem obtained by freezing the others
    prob = _solver_case("unicycle", 2, True, 7, 20)
    rng = np.random.default_rng(3)
    pd = prob.pdtraj
    pd.X[:] = rng.normal(size=pd.X.shape); pd.U[:] = rng.normal(size=pd.U.shape); pd.du[:] = rng.normal(size=pd.du.shape)
    prob.opts.reg.set(1e-3)
    Jf = O.residual_jacobian(prob, pd)
    for i in (1, 2):
        vm, hm = O.vertical_mask(prob.core, i), O.horizontal_mask(prob.core, i)
        Ji = O.ibr_residual_jacobian(prob, pd, i)
        assert np.array_equal(Ji[np.ix_(vm, hm)], Jf[np.ix_(vm, hm)])
