"""The parity tests build the oracle's problem from the package's own descriptor (`spec_of`), so a host-side mis-translation of a
constraint schema would be common to both sides.  Here the oracle's problem is built a second time NATIVELY — with the oracle's own
constructors and adders, following the reference's scripts line by line — and must agree with the spec_of route entry by entry:
residual and KKT Jacobian at a random iterate with random multipliers and penalties."""
import os, sys
import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import algames_b200 as ab
import oracle.algames_oracle as O
import parity


def native_a_prime():
    """examples/intro_example.jl:11-72 with the oracle's own API."""
    p, N, dt = 3, 20, 0.1
    model = O.BicycleGame(p)
    ps = O.ProblemSize(N, model)
    xf = [np.array([2, 0.4, 0, 0.0]), np.array([2, 0.0, 0, 0]), np.array([3, -0.4, 0, 0])]
    obj = O.GameObjective([10 * np.ones(4)] * p, [0.1 * np.ones(2)] * p, xf, [np.zeros(2)] * p, N, model)
    obj.add_collision_cost(np.ones(p), 5.0 * np.ones(p))
    con = O.GameConstraintValues(ps)
    con.add_collision_avoidance(0.08)
    con.add_control_bound(5 * np.ones(model.m), -5 * np.ones(model.m))
    con.add_state_bound(0, 5 * np.ones(model.n), -5 * np.ones(model.n))
    con.add_wall_constraint([O.Wall([0.0, -0.4], [1.0, -0.4], [0.0, -1.0])])
    con.add_circle_constraint([1.0, 2.0, 3.0], [1.0, 2.0, 3.0], [0.1, 0.2, 0.3])
    x0 = np.array([0.1, 0.0, 0.5, -0.4, 0.0, 0.7, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
    return O.GameProblem(N, dt, x0, model, O.Options(), obj, con)


def native_a():
    """test/problem/solver_methods.jl:132-182 with the oracle's own API."""
    p, N, dt = 2, 20, 0.1
    model = O.UnicycleGame(p)
    ps = O.ProblemSize(N, model)
    obj = O.GameObjective([np.ones(4)] * p, [0.5 * np.ones(2)] * p, [np.zeros(4)] * p, [-np.ones(2)] * p, N, model)
    con = O.GameConstraintValues(ps)
    con.add_collision_avoidance(0.05)
    con.add_control_bound(np.ones(model.m), -np.ones(model.m))
    con.add_circle_constraint([1.5, 0.2, 0.3], [1.25, 0.2, 0.3], [0.2, 0.2, 0.3])
    x0 = np.array([1.0, 2.0, 1.1, 2.0, 0.0, 0.0, 0.9, 0.9])
    return O.GameProblem(N, dt, x0, model, O.Options(), obj, con)


def _randomise(prob, seed):
    rng = np.random.default_rng(seed)
    pd = prob.pdtraj
    pd.X[:] = rng.normal(size=pd.X.shape); pd.U[:] = rng.normal(size=pd.U.shape); pd.du[:] = rng.normal(size=pd.du.shape)
    pd.X[0] = prob.x0
    for _, _, cv in prob.game_con.all_convals():
        cv.mu[:] = 10 ** rng.uniform(0, 3, cv.mu.shape)
        cv.lam[:] = rng.random(cv.lam.shape) * (rng.random(cv.lam.shape) > 0.5)


@pytest.mark.parametrize("name,native", [("A'", native_a_prime), ("A", native_a)])
def test_spec_route_equals_native_oracle_construction(name, native):
    model, N, dt, obj, con, opts, x0, xf = parity.small_config(name)
    via_spec = parity.oracle_problem(model, N, dt, obj, con, opts, np.asarray(x0).reshape(-1, model.n)[0])
    nat = native()
    assert [len(c) for c in nat.game_con.state_conval] == [len(c) for c in via_spec.game_con.state_conval]
    assert len(nat.game_con.control_conval) == len(via_spec.game_con.control_conval)
    _randomise(via_spec, 7); _randomise(nat, 7)
    lam_a, mu_a = O.pack_multipliers(via_spec); lam_b, mu_b = O.pack_multipliers(nat)
    assert np.array_equal(lam_a, lam_b) and np.array_equal(mu_a, mu_b)          # same rows in the same order
    ra = O.residual(via_spec).copy(); rb = O.residual(nat).copy()
    assert np.abs(ra - rb).max() <= 1e-13 * max(1.0, np.abs(rb).max())
    Ja = np.array(O.residual_jacobian(via_spec, regularize=False)); Jb = np.array(O.residual_jacobian(nat, regularize=False))
    assert np.abs(Ja - Jb).max() <= 1e-13 * max(1.0, np.abs(Jb).max())
