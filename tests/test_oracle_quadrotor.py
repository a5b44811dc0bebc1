"""QuadrotorGame in the oracle (src/dynamics/quadrotor.jl; SURVEY §8 f3 — oracle only, no device model yet).
The reference pins the index sets and sizes (test/dynamics/quadrotor.jl:4-23); the rigid-body arithmetic that lives in
Rotations.jl is unpinned, so beyond the index sets these tests check internal consistency: hover equilibrium, the MRP
rotation/kinematics pair (Ṙ = R [ω]×), Jacobians against central differences, and the KKT Jacobian of a small
2-player game against differences of the residual."""
import numpy as np
import pytest

import oracle.algames_oracle as O


def test_quadrotor_index_sets():
    # test/dynamics/quadrotor.jl:4-23 (1-based there)
    p = 3
    model = O.make_model("quadrotor", p)
    assert (model.n, model.m, model.p) == (12 * p, 4 * p, p)
    assert model.ni == [12] * p and model.mi == [4] * p
    assert [list(v + 1) for v in model.pu] == [[1, 4, 7, 10], [2, 5, 8, 11], [3, 6, 9, 12]]
    assert [list(v + 1) for v in model.px] == [[1, 4], [2, 5], [3, 6]]
    assert list(model.pz[0] + 1) == [1, 4, 7, 10, 13, 16, 19, 22, 25, 28, 31, 34]
    assert list(model.pz[2] + 1) == [3, 6, 9, 12, 15, 18, 21, 24, 27, 30, 33, 36]
    with pytest.raises(AssertionError):                                   # quadrotor.jl:21
        O.QuadrotorGame(p=5)
    for q in range(1, 5):                                                 # :38-43: dynamics defined for p = 1..4
        mq = O.make_model("quadrotor", q)
        assert mq.f(np.random.default_rng(q).random(mq.n), np.random.default_rng(q).random(mq.m)).shape == (12 * q,)


def test_quadrotor_hover_and_wrenches():
    model = O.make_model("quadrotor", 2)
    hover = model.mass * 9.81 / 4 / model.kf
    x = np.zeros(model.n); x[0:2] = [1.0, -1.0]
    u = np.full(model.m, hover)
    assert np.abs(model.f(x, u)).max() < 1e-14                            # thrust balances gravity, no torque
    assert np.allclose(model.forces(x, u, 0), 0.0, atol=1e-14) and np.allclose(model.moments(x, u, 1), 0.0)
    u2 = u.copy(); u2[[0, 2, 4, 6]] = -1.0                                # player 1's rotors commanded negative: F = max(0, kf w) = 0
    assert np.allclose(model.forces(x, u2, 0), model.mass * model.gravity)
    assert np.allclose(model.moments(x, u2, 0), [0.0, 0.0, 0.0])          # yaw moment km (w1 − w2 + w3 − w4) cancels for equal w


def test_mrp_rotation_and_kinematics_are_consistent():
    rng = np.random.default_rng(0)
    for _ in range(5):
        q, w = rng.uniform(-0.4, 0.4, 3), rng.uniform(-1, 1, 3)
        R = O.QuadrotorGame._mrp_rotation(q)
        assert np.abs(R @ R.T - np.eye(3)).max() < 1e-14 and abs(np.linalg.det(R) - 1) < 1e-14
        qd, h = O.QuadrotorGame._mrp_kinematics(q, w), 1e-6
        Rd = (O.QuadrotorGame._mrp_rotation(q + h * qd) - O.QuadrotorGame._mrp_rotation(q - h * qd)) / (2 * h)
        S = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        assert np.abs(Rd - R @ S).max() < 1e-8                            # body-rate convention: Ṙ = R [ω]×


def test_quadrotor_jacobians_match_central_differences():
    model = O.make_model("quadrotor", 3)
    rng = np.random.default_rng(1)
    x, u = 0.3 * rng.random(model.n), rng.random(model.m)
    eps = 1e-6
    for J, k, f in ((model.fx(x, u), model.n, lambda d: model.f(x + d, u)), (model.fu(x, u), model.m, lambda d: model.f(x, u + d))):
        fd = np.array([(f(eps * np.eye(k)[a]) - f(-eps * np.eye(k)[a])) / (2 * eps) for a in range(k)]).T
        assert np.abs(J - fd).max() < 1e-6
    # players are decoupled: player i's rows depend on player i's columns only
    Fx = model.fx(x, u)
    for i in range(3):
        others = [a for a in range(model.n) if a % 3 != i]
        assert np.abs(Fx[np.ix_(model.pz[i], others)]).max() == 0.0


def test_quadrotor_game_kkt_jacobian_matches_residual_differences():
    # the generic residual! / residual_jacobian! (global_quantities.jl:9-193) with 12-state players, b = 2·24 + 8 + 24 = 80
    p, N, dt = 2, 3, 0.05
    model = O.make_model("quadrotor", p)
    hover = model.mass * 9.81 / 4 / model.kf
    obj = O.GameObjective([np.ones(12)] * p, [0.1 * np.ones(4)] * p,
                          [np.r_[1.0, 0.5, 1.0, np.zeros(9)], np.r_[-1.0, 0.5, 1.0, np.zeros(9)]], [hover * np.ones(4)] * p, N, model)
    ps = O.ProblemSize(N, model)
    assert ps.S == (N - 1) * (p * model.n + model.m + model.n) == 160
    prob = O.GameProblem(N, dt, 0.05 * np.arange(model.n), model, O.Options(), obj, O.GameConstraintValues(ps))
    rng = np.random.default_rng(2)
    pd = prob.pdtraj
    # multipliers at zero: the Gauss-Newton Jacobian drops only λᵀ∂²f (global_quantities.jl:140-160), which vanishes
    # there, so J·d must equal the directional derivative of the residual on every row — including the Aᵀ Δλ columns
    pd.X[1:] = 0.1 * rng.normal(size=pd.X[1:].shape); pd.U[:] = hover + 0.1 * rng.normal(size=pd.U.shape); pd.du[:] = 0.0
    prob.opts.reg.set(0.0)                                                 # no proximal terms in J
    O.residual(prob, pd)
    J = O.residual_jacobian(prob, pd).copy()
    d = rng.normal(size=ps.S)                                              # columns follow set_traj!'s order
    O.set_traj(prob.core, prob.dpdtraj, d)
    h = 1e-6
    plus, minus = pd.copy(), pd.copy()
    O.update_traj(plus, pd, h, prob.dpdtraj); O.update_traj(minus, pd, -h, prob.dpdtraj)
    rp = O.residual(prob, plus).copy(); rm = O.residual(prob, minus).copy()
    fd = (rp - rm) / (2 * h)
    assert np.abs(J @ d - fd).max() < 1e-6 * max(1.0, np.abs(fd).max())


# ---------------------------------------------------------------- 3-D constraints (oracle only)
S2 = 1.0 / np.sqrt(2.0)


def _wall3d():
    # test/constraints/wall_constraint.jl:33-62 (x = 4, y = 2, z = 1 there; 0-based here)
    return O.Wall3DConstraint(6, [0] * 5, [0] * 5, [0] * 5, [1] * 5, [0] * 5, [0, 0, 1, 0, 0], [1, 1, 1, 0, 0], [1] * 5, [0, 1, 1, 0, 1],
                              [0, 0, -S2, 0, 0], [0, -S2, 0, 0, -S2], [1, S2, S2, 1, S2], 3, 1, 0)


def test_wall3d_constraint_values():
    con = _wall3d()
    cases = [([0.0, 0.10, -12.0, 0.10, 12.0, 11.0], [0.0, -0.1 * S2, -0.1 * S2, 0.0, -0.1 * S2]),       # :63
             ([1.0, 0.55, -12.0, 0.55, 12.0, 11.0], [1.0, 0.45 * S2, 0.45 * S2, 1.0, 0.45 * S2]),       # :64
             ([1.0, 1.25, -12.0, 0.75, 12.0, 11.0], [0.0, 0.0, 0.0, 1.0, -0.25 * S2])]                  # :65
    for X, want in cases:
        assert np.abs(con.evaluate(np.array(X), None) - np.array(want)).sum() < 1e-10
    # :68-71: jacobian! equals the derivative of evaluate (away from the mask edges)
    X = np.array([0.4, 0.55, -12.0, 0.55, 12.0, 11.0]); h = 1e-7
    fd = np.array([(con.evaluate(X + h * np.eye(6)[a], None) - con.evaluate(X - h * np.eye(6)[a], None)) / (2 * h) for a in range(6)]).T
    assert np.abs(con.jacobian(X, None) - fd).sum() < 1e-7


def test_cylinder_constraint_values():
    # test/constraints/cylinder_constraint.jl:4-35 (x = 2, y = 3, z = 4 there)
    con = O.CylinderConstraint(4, [1, 1, 1, 0, 1], [0, 0, 0, 1, 0], [1, 1, 3, 1, 2], ["z", "z", "z", "x", "y"],
                               [5, 2, 0.5, 2, 10], [3, 2, 7, 3, 1], 1, 2, 3)
    X0 = np.array([13.0, 1.0, 1.0, 2.0])
    assert np.abs(con.evaluate(X0, None) - np.array([8.0, 3.0, 0.0, 8.0, 1.0])).sum() < 1e-10               # :21
    h = 1e-7
    fd = np.array([(con.evaluate(X0 + h * np.eye(4)[a], None) - con.evaluate(X0 - h * np.eye(4)[a], None)) / (2 * h) for a in range(4)]).T
    assert np.abs(con.jacobian(X0, None) - fd).sum() < 1e-6                                                 # :30-34


def test_spherical_collision_avoidance_indices():
    # test/constraints/constraints_methods.jl:30-57: 3-player, 3-D double integrator; the collision convals of player i
    # pair its first three state components with each opponent's (pu == pz[:3] for this model)
    p = 3
    model = O.make_model("double_integrator", p, d=3)
    gc = O.GameConstraintValues(O.ProblemSize(20, model))
    gc.add_spherical_collision_avoidance(1.0)
    for i in range(p):
        others = [j for j in range(p) if j != i]
        assert len(gc.state_conval[i]) == 2
        for cv, j in zip(gc.state_conval[i], others):
            assert list(cv.con.x1) == list(model.pu[i]) and list(cv.con.x2) == list(model.pu[j])
            assert cv.con.radius == 2.0 and list(cv.inds) == list(range(2, 21))


def test_3d_adders_build_one_conval_per_player_on_position_components():
    # add_wall_constraint!(…, Vector{Wall3D}) / (…, Vector{CylinderWall}) (constraints_methods.jl:208-285) for a quadrotor game
    model = O.make_model("quadrotor", 2)
    gc = O.GameConstraintValues(O.ProblemSize(5, model))
    gc.add_wall3d_constraint([O.Wall3D(np.zeros(3), np.array([1.0, 0, 0]), np.array([1.0, 1, 0]), np.array([0, 0, 1.0]))])
    gc.add_cylinder_constraint([O.CylinderWall(np.array([0.5, 0.5, 0.0]), "z", 2.0, 0.3)], i=1)
    assert [len(c) for c in gc.state_conval] == [1, 2]
    w = gc.state_conval[1][0].con
    assert (w.x, w.y, w.z) == tuple(model.pz[1][:3]) and w.length() == 1
    cyl = gc.state_conval[1][1].con
    x = np.zeros(model.n); x[model.pz[1][:3]] = [0.6, 0.5, 1.0]            # inside the cylinder, 0.1 from its axis
    assert np.isclose(cyl.evaluate(x, None)[0], 0.3 ** 2 - 0.1 ** 2)
    x[model.pz[1][2]] = 2.5                                                 # above the top cap: inactive
    assert cyl.evaluate(x, None)[0] == 0.0


@pytest.mark.parametrize("p,N", [(1, 6), (2, 5)])
def test_quadrotor_game_solves_on_the_oracle(p, N):
    # the generic newton_solve! loop (solver_methods.jl:5-65) with 12-state / 4-control players: hover-to-waypoint LQ game,
    # for p = 2 with spherical collision avoidance (inactive along the way); same convergence test as every other model
    model = O.make_model("quadrotor", p)
    hover = model.mass * 9.81 / 4 / model.kf
    xf = [np.r_[0.3 * (1 - 2 * i), 0.2, 1.1, np.zeros(9)] for i in range(p)]
    obj = O.GameObjective([np.ones(12)] * p, [0.1 * np.ones(4)] * p, xf, [hover * np.ones(4)] * p, N, model)
    gc = O.GameConstraintValues(O.ProblemSize(N, model))
    if p > 1:
        gc.add_spherical_collision_avoidance(0.1)
    x0 = np.zeros(model.n); x0[2 * p:3 * p] = 1.0; x0[0:p] = np.linspace(-0.5, 0.5, p) if p > 1 else 0.0
    prob = O.GameProblem(N, 0.1, x0, model, O.Options(), obj, gc)
    O.newton_solve(prob)
    last = prob.stats[-1]
    assert prob.converged and max(last.dyn, last.con, last.sta, last.opt) < 1e-3
    assert prob.pdtraj.X[-1][2 * p] > 1.0                                  # climbing towards z = 1.1
