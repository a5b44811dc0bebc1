"""add_velocity_bound! / velocity_index (src/constraints/velocity_constraint.jl) and several StateBound convals per
player: the reference's own test (test/constraints/velocity_constraint.jl) restated for the oracle and for the host
package, the descriptor row schema (agb_sizes_of — host arithmetic only), and device-logic parity on the CTA
emulator.  The GPU parity of the same schema is in tests/test_gpu_parity.py (config "V")."""
import ctypes as C

import numpy as np
import pytest

import parity
from test_emulated_kernels import emu_lib  # noqa: F401  (session fixture: builds tests/emu/libagb_emu.so)

import algames_b200 as ab
import oracle.algames_oracle as O

INF = np.inf
# (model name, v_max, v_min, StateBound convals per player) — test/constraints/velocity_constraint.jl:3-45
REFERENCE_CASES = [
    ("unicycle", [1.0, 1.0, 1.0], [-1.0, -1.0, -1.0], 3),        # :13-19
    ("bicycle", [1.0, INF, INF], [1.0, -1.0, -INF], 2),          # :26-32
    ("bicycle", [INF, INF, INF], [-INF, -INF, -INF], 0),         # :37-43
]


@pytest.mark.parametrize("name,v_max,v_min,count", REFERENCE_CASES)
def test_reference_velocity_constraint_test_on_the_oracle(name, v_max, v_min, count):
    model = O.make_model(name, 3)
    gc = O.GameConstraintValues(O.ProblemSize(10, model))
    gc.add_velocity_bound(model, np.array(v_max), np.array(v_min))
    assert len(gc.state_conval) == model.p
    assert all(len(cvs) == count for cvs in gc.state_conval)
    # every conval bounds exactly the limited player's speed component, on knots 2:N
    limited = [a for a in range(3) if v_max[a] != INF or v_min[a] != -INF]
    for cvs in gc.state_conval:
        for a, cv in zip(limited, cvs):
            vi = O.velocity_index(model, a)
            assert np.isfinite(cv.con.x_max).sum() == (v_max[a] != INF) and np.isfinite(cv.con.x_min).sum() == (v_min[a] != -INF)
            assert cv.con.x_max[vi] == v_max[a] and cv.con.x_min[vi] == v_min[a]
            assert list(cv.inds) == list(range(2, 11))


def test_velocity_index():
    # velocity_constraint.jl:30-43 (1-based pz[i][4] / pz[i][3]); 0-based here
    for p in (2, 3, 4):
        uni, bic = ab.UnicycleGame(p=p), ab.BicycleGame(p=p)
        for i in range(p):
            assert ab.velocity_index(uni, i) == O.velocity_index(O.make_model("unicycle", p), i) == 3 * p + i
            assert ab.velocity_index(bic, i) == O.velocity_index(O.make_model("bicycle", p), i) == 2 * p + i
    with pytest.raises(NotImplementedError):
        ab.velocity_index(ab.DoubleIntegratorGame(p=2), 0)
    with pytest.raises(NotImplementedError):
        O.velocity_index(O.make_model("double_integrator", 2), 0)
    with pytest.raises(AssertionError):
        ab.velocity_index(ab.UnicycleGame(p=2), 2)


@pytest.mark.parametrize("name,v_max,v_min,count", REFERENCE_CASES)
def test_reference_velocity_constraint_test_on_the_host_package(name, v_max, v_min, count):
    model = {"unicycle": ab.UnicycleGame, "bicycle": ab.BicycleGame}[name](p=3)
    N = 10
    con = ab.GameConstraintValues(ab.ProblemSize(N, model))
    ab.add_velocity_bound(model, con, v_max, v_min)
    assert len(con.state_bound) == model.p and all(len(b) == count for b in con.state_bound)
    # descriptor: has_state_bound counts the convals, rows = finite entries, one block per player
    obj = ab.GameObjective([np.ones(4)] * 3, [np.ones(2)] * 3, [np.zeros(4)] * 3, [np.zeros(2)] * 3, N, model)
    d = ab.problem._make_desc(model, N, 0.1, obj, con)
    assert list(d.has_state_bound)[:3] == [count] * 3
    rows = int(np.isfinite(v_max).sum() + np.isfinite(v_min).sum())
    sz = ab._capi.Sizes()
    assert ab._capi.load().agb_sizes_of(C.byref(d), C.byref(sz)) == 0
    assert (sz.nrow_state, sz.nrow_control, sz.nrow) == (3 * rows, 0, 3 * rows)
    # the spec the tests hand to the oracle rebuilds the same convals in the same order
    prob = ab.GameProblem(N, 0.1, np.zeros(model.n), model, ab.Options(), obj, con, lib_path="unused")
    op = O.problem_from_spec(ab.spec_of(prob))
    ref = O.GameConstraintValues(O.ProblemSize(N, O.make_model(name, 3)))
    ref.add_velocity_bound(O.make_model(name, 3), np.array(v_max), np.array(v_min))
    for i in range(3):
        assert len(op.game_con.state_conval[i]) == len(ref.state_conval[i]) == count
        for a, b in zip(op.game_con.state_conval[i], ref.state_conval[i]):
            assert np.array_equal(a.con.x_max, b.con.x_max) and np.array_equal(a.con.x_min, b.con.x_min)


def test_state_bound_schema_edges():
    model = ab.UnicycleGame(p=2)
    n = model.n
    con = ab.GameConstraintValues(ab.ProblemSize(5, model))
    hi = np.full(n, INF); hi[0] = 1.0
    ab.add_state_bound(con, 0, hi, np.full(n, -INF))
    ab.add_state_bound(con, 0, np.full(n, INF), np.full(n, -INF))            # a conval without rows is fine
    lo = np.full(n, -INF); lo[0] = 2.0                                        # other side of the same component: separate
    ab.add_state_bound(con, 0, np.full(n, INF), lo)                           # convals ⇒ no max >= min check between them
    with pytest.raises(NotImplementedError):                                  # same component, same side: no device form
        ab.add_state_bound(con, 0, hi, np.full(n, -INF))
    assert len(con.state_bound[0]) == 3
    with pytest.raises(ValueError):                                           # state_bound_constraint.jl:62-68
        ab.add_state_bound(con, 1, np.zeros(n), np.ones(n))
    with pytest.raises(AssertionError):                                       # velocity_constraint.jl:14
        ab.add_velocity_bound(model, con, INF, -INF, i=0)
    obj = ab.GameObjective([np.ones(4)] * 2, [np.ones(2)] * 2, [np.zeros(4)] * 2, [np.zeros(2)] * 2, 5, model)
    d = ab.problem._make_desc(model, 5, 0.1, obj, con)
    assert d.has_state_bound[0] == 3 and d.x_max_con[0][0] == 0 and d.x_min_con[0][0] == 2
    sz = ab._capi.Sizes()
    lib = ab._capi.load()
    assert lib.agb_sizes_of(C.byref(d), C.byref(sz)) == 0 and sz.nrow == 2
    d.x_min_con[0][0] = 3                                                     # conval index out of range
    assert lib.agb_sizes_of(C.byref(d), C.byref(sz)) == -1   # AGB_EINVAL
    d.x_min_con[0][0] = 0                                                     # now one conval holds max 1 < min 2
    assert lib.agb_sizes_of(C.byref(d), C.byref(sz)) == -1   # AGB_EINVAL


def test_velocity_bound_per_function_parity_emulated(emu_lib):  # noqa: F811
    parity.check_per_function(emu_lib, "V", seed=1, N=10)


def test_velocity_bound_solve_emulated(emu_lib):  # noqa: F811
    out = parity.check_solve_vs_oracle(emu_lib, "V", B=2, N=12)
    assert out["status"][0] == 0
    # speed limits are active at the solution: multipliers of the limit rows are positive somewhere
    nrow_player = 2 + 4 + 2                                                   # collision | 4 speed-limit rows | walls
    sb = np.concatenate([out["conlam"][0][:, i * nrow_player + 2:i * nrow_player + 6] for i in range(3)], axis=1)
    assert (sb > 0).any()


def test_interleaved_state_bound_convals_emulated(emu_lib):  # noqa: F811
    # config S: three / two StateBound convals per player with interleaved components and sides, one without rows;
    # the AL rows must follow the reference's conval order (oracle: one ALConVal per add_state_bound!)
    parity.check_per_function(emu_lib, "S", seed=5)
    out = parity.check_solve_vs_oracle(emu_lib, "S", B=2)
    assert (out["status"] == 0).all() and (out["conlam"] > 0).any(axis=(0, 1)).sum() >= 3


@pytest.mark.parametrize("name", ["S1", "S2", "S3", "S4"])
def test_random_state_bound_layouts_emulated(emu_lib, name):  # noqa: F811
    # random assignment of (component, side) pairs to up to four convals per player: values, AL rows, multiplier updates
    # and the Newton step must follow the oracle's conval-by-conval row order
    parity.check_per_function(emu_lib, name, seed=7)
