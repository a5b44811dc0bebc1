"""Synthetic inputs of the BASELINE.json configurations.  Each builder returns
(model, N, dt, game_obj, game_con, opts, x0[B,n], xf[B,n] or None)."""
from __future__ import annotations

import math

import numpy as np

from .problem import (BicycleGame, CylinderWall, DoubleIntegratorGame, GameConstraintValues, GameObjective, Options, ProblemSize,
                      QuadrotorGame, UnicycleGame, Wall, Wall3D, add_circle_constraint, add_collision_avoidance, add_collision_cost,
                      add_control_bound, add_spherical_collision_avoidance, add_state_bound, add_velocity_bound, add_wall_constraint)


def config_a():
    """2-player UnicycleGame, N=20: test/problem/solver_methods.jl:132-182 verbatim (runs with the previous
    block's opts, :108-126 — see SURVEY §8d)."""
    p, N, dt = 2, 20, 0.1
    model = UnicycleGame(p=p)
    ps = ProblemSize(N, model)
    obj = GameObjective([np.ones(4)] * p, [0.5 * np.ones(2)] * p, [np.zeros(4)] * p, [-np.ones(2)] * p, N, model)
    con = GameConstraintValues(ps)
    add_collision_avoidance(con, 0.05)
    add_control_bound(con, np.ones(model.m), -np.ones(model.m))
    add_circle_constraint(con, [1.5, 0.2, 0.3], [1.25, 0.2, 0.3], [0.2, 0.2, 0.3])
    opts = Options(outer_iter=7, inner_iter=20, ls_iter=25, reg_0=1e-7, eps_dyn=1e-10, eps_opt=1e-10)
    x0 = np.array([[1.0, 2.0, 1.1, 2.0, 0.0, 0.0, 0.9, 0.9]])
    return model, N, dt, obj, con, opts, x0, None


def config_a_prime():
    """examples/intro_example.jl:11-72 verbatim: 3-player BicycleGame, N=20."""
    p, N, dt = 3, 20, 0.1
    model = BicycleGame(p=p)
    ps = ProblemSize(N, model)
    xf = [np.array([2, 0.4, 0, 0.0]), np.array([2, 0.0, 0, 0]), np.array([3, -0.4, 0, 0])]
    obj = GameObjective([10 * np.ones(4)] * p, [0.1 * np.ones(2)] * p, xf, [np.zeros(2)] * p, N, model)
    add_collision_cost(obj, np.ones(p), 5.0 * np.ones(p))
    con = GameConstraintValues(ps)
    add_collision_avoidance(con, 0.08)
    add_control_bound(con, 5 * np.ones(model.m), -5 * np.ones(model.m))
    add_state_bound(con, 0, 5 * np.ones(model.n), -5 * np.ones(model.n))
    add_wall_constraint(con, [Wall([0.0, -0.4], [1.0, -0.4], [0.0, -1.0])])
    add_circle_constraint(con, [1.0, 2.0, 3.0], [1.0, 2.0, 3.0], [0.1, 0.2, 0.3])
    x0 = np.array([[0.1, 0.0, 0.5, -0.4, 0.0, 0.7, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]])
    return model, N, dt, obj, con, Options(), x0, None


def config_b(batch=1024, seed=1234, N=40):
    """3-player DoubleIntegratorGame(d=2), N=40 — shape of examples/ibr_example.jl:12-74, jittered x0."""
    p, dt = 3, 0.1
    model = DoubleIntegratorGame(p=p, d=2)
    ps = ProblemSize(N, model)
    obj = GameObjective([50 * np.ones(4)] * p, [0.01 * np.ones(2)] * p, [np.zeros(4)] * p, [np.zeros(2)] * p, N, model)
    add_collision_cost(obj, 3.0 * np.ones(p), 2.0 * np.ones(p))
    con = GameConstraintValues(ps)
    add_collision_avoidance(con, 0.25)
    th = np.array([2 * math.pi * (1 - 1 / p) * i / (p - 1) for i in range(p)])
    base = np.zeros(model.n)
    base[0:p] = (0.5 + th / 10) * np.cos(th)
    base[p:2 * p] = (0.5 + th / 10) * np.sin(th)
    rng = np.random.default_rng(seed)
    x0 = np.tile(base, (batch, 1))
    x0[:, :2 * p] += rng.uniform(-0.1, 0.1, (batch, 2 * p))
    x0[:, 2 * p:] += rng.uniform(-0.05, 0.05, (batch, 2 * p))
    return model, N, dt, obj, con, Options(), x0, None


def config_c(batch=8192, seed=2345, N=50):
    """4-player UnicycleGame with collision constraints + control bounds, N=50."""
    p, dt = 4, 0.1
    model = UnicycleGame(p=p)
    ps = ProblemSize(N, model)
    obj = GameObjective([np.ones(4)] * p, [0.5 * np.ones(2)] * p, [np.zeros(4)] * p, [np.zeros(2)] * p, N, model)
    con = GameConstraintValues(ps)
    add_collision_avoidance(con, 0.15)
    add_control_bound(con, 2 * np.ones(model.m), -2 * np.ones(model.m))
    ang = np.array([2 * math.pi * i / p for i in range(p)])
    rng = np.random.default_rng(seed)
    x0 = np.zeros((batch, model.n))
    x0[:, 0:p] = 2 * np.cos(ang) + rng.uniform(-0.1, 0.1, (batch, p))
    x0[:, p:2 * p] = 2 * np.sin(ang) + rng.uniform(-0.1, 0.1, (batch, p))
    x0[:, 2 * p:3 * p] = ang + math.pi                   # heading to the centre
    x0[:, 3 * p:4 * p] = 0.5
    xf = np.zeros((batch, model.n))
    xf[:, 0:p] = -2 * np.cos(ang)                        # antipodal point
    xf[:, p:2 * p] = -2 * np.sin(ang)
    xf[:, 2 * p:3 * p] = ang + math.pi
    return model, N, dt, obj, con, Options(), x0, xf


def _lane_game(batch, seed, N, ramp):
    """Synthetic 3-player Unicycle lane scenario used by configs D (ramp merge) and E (highway)."""
    p, dt = 3, 0.1
    model = UnicycleGame(p=p)
    ps = ProblemSize(N, model)
    obj = GameObjective([np.array([0.0, 1.0, 1.0, 1.0])] * p, [0.5 * np.ones(2)] * p, [np.zeros(4)] * p, [np.zeros(2)] * p, N, model)
    con = GameConstraintValues(ps)
    add_collision_avoidance(con, 0.08)
    add_control_bound(con, 5 * np.ones(model.m), -5 * np.ones(model.m))
    half = 0.4 if ramp else 0.5
    lane = [Wall([-50.0, half], [50.0, half], [0.0, 1.0]), Wall([-50.0, -half], [50.0, -half], [0.0, -1.0])]
    rng = np.random.default_rng(seed)
    x0 = np.zeros((batch, model.n)); xf = np.zeros((batch, model.n))
    if ramp:
        add_wall_constraint(con, lane, 0); add_wall_constraint(con, lane, 1)
        add_wall_constraint(con, [lane[0], Wall([-50.0, -0.9], [50.0, -0.9], [0.0, -1.0])], 2)
        x0[:, 0:p] = np.array([0.0, 0.6, 0.3]) + rng.uniform(-0.05, 0.05, (batch, p))
        x0[:, p:2 * p] = np.array([0.15, -0.15, -0.6])
        x0[:, 3 * p:4 * p] = 1.0
        xf[:, p:2 * p] = np.array([0.15, -0.15, -0.15])
        xf[:, 3 * p:4 * p] = 1.0
    else:
        add_wall_constraint(con, lane)
        x0[:, 0:p] = np.sort(rng.uniform(0, 3, (batch, p)), axis=1)
        lanes = rng.choice([-0.25, 0.25], (batch, p))
        x0[:, p:2 * p] = lanes
        x0[:, 3 * p:4 * p] = rng.uniform(0.5, 1.5, (batch, p))
        xf[:, p:2 * p] = lanes
        xf[:, 3 * p:4 * p] = rng.uniform(0.8, 1.2, (batch, p))
    return model, N, dt, obj, con, Options(), x0, xf


def config_d(batch=4096, seed=3456, N=40):
    """MPC ramp merge (synthetic; AlgamesDriving.jl is not in the reference tree): shift=1, dual_reset=False."""
    model, N, dt, obj, con, opts, x0, xf = _lane_game(batch, seed, N, ramp=True)
    opts.shift, opts.dual_reset = 1, False
    return model, N, dt, obj, con, opts, x0, xf


def config_e(batch=65536, seed=4567, N=60):
    """Monte-Carlo highway sweep (synthetic)."""
    return _lane_game(batch, seed, N, ramp=False)


def config_v(batch=64, seed=5678, N=20):
    """Highway lane game of config E with speed limits: add_velocity_bound! (velocity_constraint.jl:1-28) gives every
    player one StateBound conval per limited player — both sides for player 1, an upper limit for player 2, a lower
    limit for player 3 — next to the collision, wall and control-bound rows.  Desired speeds above the limit keep
    some of the rows active at the solution."""
    model, N, dt, obj, con, opts, x0, xf = _lane_game(batch, seed, N, ramp=False)
    add_velocity_bound(model, con, [1.1, 1.0, np.inf], [0.6, -np.inf, 0.7])
    return model, N, dt, obj, con, opts, x0, xf


def config_s(batch=8, seed=6789, N=12):
    """Row-order stress for several StateBound convals per player (add_state_bound! called repeatedly,
    constraints_methods.jl:87-98): 2-player unicycle, player 1 with three convals whose finite entries interleave
    (components out of order across convals, both sides, one conval without rows), player 2 with two; plus collision
    avoidance and control bounds so the bound rows sit between other row kinds."""
    p, dt = 2, 0.1
    model = UnicycleGame(p=p)
    ps = ProblemSize(N, model)
    n, inf = model.n, np.inf
    obj = GameObjective([np.ones(4)] * p, [0.5 * np.ones(2)] * p, [np.array([2.0, 0.3, 0.0, 1.0]), np.array([2.0, -0.3, 0.0, 1.0])],
                        [np.zeros(2)] * p, N, model)
    con = GameConstraintValues(ps)
    add_collision_avoidance(con, 0.1)
    add_control_bound(con, 3 * np.ones(model.m), -3 * np.ones(model.m))

    def bound(hi=(), lo=()):
        x_max, x_min = np.full(n, inf), np.full(n, -inf)
        for a, v in hi:
            x_max[a] = v
        for a, v in lo:
            x_min[a] = v
        return x_max, x_min

    # joint state (component-major): 0,1 = x; 2,3 = y; 4,5 = heading; 6,7 = speed
    add_state_bound(con, 0, *bound(hi=[(6, 1.2), (2, 0.5)], lo=[(2, -0.1)]))
    add_state_bound(con, 0, *bound())                                   # a conval without rows
    add_state_bound(con, 0, *bound(hi=[(0, 1.5), (7, 1.1)], lo=[(6, 0.7), (3, -0.5)]))
    add_state_bound(con, 1, *bound(lo=[(7, 0.6), (0, -1.0)]))
    add_state_bound(con, 1, *bound(hi=[(7, 1.1), (3, 0.1)], lo=[(4, -0.4)]))
    rng = np.random.default_rng(seed)
    x0 = np.zeros((batch, n))
    x0[:, 0:2] = rng.uniform(0.0, 0.3, (batch, 2))
    x0[:, 2:4] = np.array([0.25, -0.25]) + rng.uniform(-0.05, 0.05, (batch, 2))
    x0[:, 6:8] = rng.uniform(0.8, 1.4, (batch, 2))
    return model, N, dt, obj, con, Options(), x0, None


def config_s_random(layout_seed, batch=4, seed=6789, N=8):
    """Config S with a random StateBound layout: every (component, side) of the joint state is given, with probability
    1/2, to one of up to four convals of each player (bounds wide enough to stay feasible, tight enough that some are
    active on a random iterate).  Used by the randomized row-order tests."""
    model, N, dt, obj, con, opts, x0, xf = config_s(batch=batch, seed=seed, N=N)
    con.state_bound = [[] for _ in range(model.p)]
    rng = np.random.default_rng(layout_seed)
    n = model.n
    for i in range(model.p):
        ncon = int(rng.integers(1, 5))
        his = [np.full(n, np.inf) for _ in range(ncon)]
        los = [np.full(n, -np.inf) for _ in range(ncon)]
        for a in range(n):
            if rng.random() < 0.5:
                his[int(rng.integers(ncon))][a] = float(rng.uniform(0.5, 3.0))
            if rng.random() < 0.5:
                los[int(rng.integers(ncon))][a] = float(rng.uniform(-3.0, -0.5))
        for hi, lo in zip(his, los):
            add_state_bound(con, i, hi, lo)
    return model, N, dt, obj, con, opts, x0, xf


def config_q(batch=8, seed=7890, N=12, p=2, constraints=True):
    """QuadrotorGame (dynamics/quadrotor.jl): p quadrotors fly from a hover at z = 1 to waypoints on the other side, crossing
    each other; spherical collision avoidance (constraints_methods.jl:45-81), rotor-command bounds, a 3-D wall under the flight
    corridor (wall_constraint.jl:141-249) and a vertical cylinder next to it (cylinder_constraint.jl:33-137).  The reference
    ships no quadrotor example; shapes follow test/dynamics/quadrotor.jl and the oracle's test_oracle_quadrotor.py."""
    dt = 0.1
    model = QuadrotorGame(p=p)
    ps = ProblemSize(N, model)
    hover = model.mass * 9.81 / 4 / 1.245
    xf = [np.r_[0.4 * (1 - 2 * (i % 2)), 0.2 * (i // 2) + 0.2 * (i % 2), 1.1, np.zeros(9)] for i in range(p)]      # lanes 0.2 apart
    obj = GameObjective([np.ones(12)] * p, [1.0 * np.ones(4)] * p, xf, [hover * np.ones(4)] * p, N, model)
    con = GameConstraintValues(ps)
    if constraints:
        if p > 1:
            add_spherical_collision_avoidance(con, 0.15)
        add_control_bound(con, 2.0 * np.ones(model.m), np.zeros(model.m))
        add_wall_constraint(con, [Wall3D([-1.0, -1.0, 0.9], [1.0, -1.0, 0.9], [1.0, 1.0, 0.9], [0.0, 0.0, -1.0])])
        add_wall_constraint(con, [CylinderWall([0.0, 0.6, 0.0], "z", 3.0, 0.2)], 0)
    rng = np.random.default_rng(seed)
    x0 = np.zeros((batch, model.n))
    x0[:, 0:p] = np.array([-0.4 * (1 - 2 * (i % 2)) for i in range(p)]) + rng.uniform(-0.05, 0.05, (batch, p))
    x0[:, p:2 * p] = np.array([0.2 * (i // 2) + 0.2 * (i % 2) for i in range(p)]) + rng.uniform(-0.05, 0.05, (batch, p))
    x0[:, 2 * p:3 * p] = 1.0 + rng.uniform(-0.02, 0.02, (batch, p))
    return model, N, dt, obj, con, Options(), x0, None


def config_b3(batch=8, seed=8901, N=10):
    """3-player DoubleIntegratorGame(d = 3) with spherical collision avoidance: the game of the reference's
    test/constraints/constraints_methods.jl:30-57 given an objective (band solver)."""
    p, dt = 3, 0.1
    model = DoubleIntegratorGame(p=p, d=3)
    obj = GameObjective([10 * np.ones(6)] * p, [0.1 * np.ones(3)] * p, [np.zeros(6)] * p, [np.zeros(3)] * p, N, model)
    add_collision_cost(obj, 1.0 * np.ones(p), 2.0 * np.ones(p))
    con = GameConstraintValues(ProblemSize(N, model))
    add_spherical_collision_avoidance(con, 0.2)
    add_control_bound(con, 3 * np.ones(model.m), -3 * np.ones(model.m))
    rng = np.random.default_rng(seed)
    th = 2 * math.pi * np.arange(p) / p
    x0 = np.zeros((batch, model.n))
    x0[:, 0:p] = 0.6 * np.cos(th) + rng.uniform(-0.05, 0.05, (batch, p))
    x0[:, p:2 * p] = 0.6 * np.sin(th) + rng.uniform(-0.05, 0.05, (batch, p))
    x0[:, 2 * p:3 * p] = rng.uniform(-0.2, 0.2, (batch, p))
    return model, N, dt, obj, con, Options(), x0, None


CONFIGS = {"B3": config_b3, "Q": config_q, "S": config_s, "V": config_v, "A": config_a, "A'": config_a_prime, "B": config_b, "C": config_c, "D": config_d, "E": config_e}
for _ls in range(1, 5):                                   # "S1" … "S4": random StateBound layouts
    CONFIGS["S%d" % _ls] = (lambda ls: (lambda batch=4, N=8: config_s_random(ls, batch=batch, N=N)))(_ls)
