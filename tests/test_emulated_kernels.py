"""CPU tier: the CUDA library's sources compiled by g++ against the test-only CTA emulator (tests/emu) and driven
through the same C ABI, compared with the oracle.  Checks kernel *logic* without a GPU; the real parity tests are
tests/test_gpu_parity.py."""
import os
import subprocess

import numpy as np
import pytest

import parity

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu", "libagb_emu.so")


@pytest.fixture(scope="session")
def emu_lib():
    src = [os.path.join(HERE, "emu", f) for f in ("emu.cpp", "cuda_runtime.h")]
    csrc = os.path.join(HERE, "..", "algames.jl_b200", "csrc")
    src += [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if not os.path.exists(EMU) or any(os.path.getmtime(s) > os.path.getmtime(EMU) for s in src):
        subprocess.run([os.path.join(HERE, "emu", "build.sh")], check=True)
    return EMU


@pytest.mark.parametrize("name,N", [("A", None), ("A'", None), ("B", 12), ("C", 10), ("D", 10), ("E", 12)])
def test_per_function_parity_emulated(emu_lib, name, N):
    parity.check_per_function(emu_lib, name, seed=1, N=N)


@pytest.mark.parametrize("name", ["A'", "B"])
def test_rollout_emulated(emu_lib, name):
    parity.check_rollout(emu_lib, name)


def test_solve_config_a_emulated(emu_lib):
    # test/problem/solver_methods.jl:132-182 through the device control flow
    out = parity.check_solve_vs_oracle(emu_lib, "A", B=1)
    assert out["status"][0] == 0 and (out["stats"][0, 1:5] < 1e-3).all()


def test_solve_config_b_short_horizon_emulated(emu_lib):
    parity.check_solve_vs_oracle(emu_lib, "B", B=2, N=12)


def test_unconstrained_lq_game_one_newton_step_emulated(emu_lib):
    # test/problem/solver_methods.jl:68-97: 2-player LQ game solved by exactly one Newton step
    import algames_b200 as ab
    p, N, dt = 2, 20, 0.1
    model = ab.DoubleIntegratorGame(p=p)
    obj = ab.GameObjective([np.ones(4)] * p, [0.5 * np.ones(2)] * p, [np.zeros(4)] * p, [-np.ones(2)] * p, N, model)
    con = ab.GameConstraintValues(ab.ProblemSize(N, model))
    opts = ab.Options(outer_iter=1, inner_iter=1, ls_iter=25, reg_0=1e-7, eps_dyn=1e-10, eps_opt=1e-10)
    prob = ab.GameProblem(N, dt, [1.0, 2.0, 1.0, 2.0, 0, 0, 0.9, 0.9], model, opts, obj, con, lib_path=emu_lib)
    ab.newton_solve(prob)
    assert np.abs(prob.core.res).sum() / prob.probsize.S < 1e-6
    assert prob.stats.dyn_vio[-1].max < 1e-6 and prob.stats.newton_steps == 1


def test_statistics_history_single_problem_emulated(emu_lib):
    # prob.stats keeps one entry per record!(stats, …) like the reference's Statistics (struct/statistics.jl:44-57)
    import algames_b200 as ab
    import oracle.algames_oracle as O
    model, N, dt, obj, con, opts, x0, _ = ab.workloads.config_a()
    prob = ab.GameProblem(N, dt, x0[0], model, opts, obj, con, lib_path=emu_lib)
    ab.init_traj(prob)
    Z0 = np.concatenate([prob.pdtraj.X, prob.pdtraj.U], axis=1).copy(); L0 = prob.pdtraj.du.copy()
    ab.newton_solve(prob, init=False)
    op = O.newton_solve(O.problem_from_spec(ab.spec_of(prob)), Z0=Z0, L0=L0)
    st = prob.stats
    assert st.iter == len(op.stats) == len(st.res) == len(st.opt_vio) and st.iter > opts.outer_iter
    assert st.outer_iter == [r.outer for r in op.stats]
    assert np.allclose(st.res, [r.res for r in op.stats], rtol=1e-6, atol=1e-9)
    assert np.allclose([v.max for v in st.dyn_vio], [r.dyn for r in op.stats], rtol=1e-6, atol=1e-9)
    assert np.allclose(st.delta, [r.delta for r in op.stats], rtol=1e-6, atol=1e-9)
    assert st.sta_vio[-1].max < 1e-3 and st.con_vio[-1].max < 1e-3          # test/problem/solver_methods.jl:178-182


# ---- big layout (duals, AL multipliers and pair/self Hessian blocks in global memory; agb_internal.h): 4-player games always
# use it (config C above); the test hook forces it on small 3-player instances
@pytest.mark.parametrize("lay", ["1", "2", "3"])
@pytest.mark.parametrize("name,N", [("A'", None), ("B", 12), ("E", 12)])
def test_big_layout_per_function_parity_emulated(emu_lib, monkeypatch, name, N, lay):
    monkeypatch.setenv("AGB_FORCE_BIG_LAYOUT", lay)
    parity.check_per_function(emu_lib, name, seed=2, N=N)


@pytest.mark.parametrize("lay", ["1", "2", "3"])
def test_big_layout_solve_emulated(emu_lib, monkeypatch, lay):
    monkeypatch.setenv("AGB_FORCE_BIG_LAYOUT", lay)
    parity.check_solve_vs_oracle(emu_lib, "B", B=2, N=12)
    parity.check_ibr_solve(emu_lib, "B", B=1, N=8, ibr_iter=1)


def test_error_behaviour_emulated(emu_lib):
    import algames_b200 as ab
    model = ab.UnicycleGame(p=2)
    con = ab.GameConstraintValues(ab.ProblemSize(10, model))
    with pytest.raises(ValueError):                      # control_bound_constraint.jl:62-68
        ab.add_control_bound(con, -np.ones(4), np.ones(4))
    obj = ab.GameObjective([np.ones(4)] * 2, [np.ones(2)] * 2, [np.zeros(4)] * 2, [np.zeros(2)] * 2, 10, model)
    with pytest.raises(ab.AlgamesError):
        ab.GameBatch(model, 10, 0.1, obj, con, 0, lib_path=emu_lib)       # empty batch
    with pytest.raises(ab.AlgamesError):
        ab.GameBatch(model, 1, 0.1, obj, con, 1, lib_path=emu_lib)        # N < 2
    gb = ab.GameBatch(model, 10, 0.1, obj, con, 1, lib_path=emu_lib)
    with pytest.raises(ValueError):
        gb.set_instance_params(x0=np.zeros((2, 8)))                       # ragged batch
    gb.close()


@pytest.mark.parametrize("f", ["solve_C.npz", "solve_E.npz"])
def test_golden_emulated(emu_lib, f):
    parity.check_golden(emu_lib, os.path.join(HERE, "golden", f))


def test_mpc_loop_emulated(emu_lib):
    """Receding-horizon loop (config D shape, short horizon): closed-loop state follows x_2 of each solve, warm starts
    keep duals, and every re-solve converges."""
    import algames_b200 as ab
    model, N, dt, obj, con, opts, x0, xf = ab.workloads.config_d(batch=2, N=10)
    gb = ab.GameBatch(model, N, dt, obj, con, 2, lib_path=emu_lib)
    stats, status, xs = ab.mpc.mpc_run(gb, opts, x0, 3, xf=xf, disturbance_std=1e-3, seed=3)
    assert stats.shape == (3, 2, 10) and (status == 0).all()
    assert np.isfinite(xs).all() and (xs[-1][:, 0] > xs[0][:, 0]).all()        # the cars move forward
    assert (stats[1:, :, 6] <= stats[0, :, 6] + 2).all()                        # warm starts are not harder than the cold start
    gb.close()


def test_gauss_jordan_pivot_fallback_emulated(emu_lib):
    parity.check_pivot_fallback(emu_lib)


# ---- iterative best response -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,N", [("A", None), ("A'", 8), ("B", 8), ("C", 6)])
def test_ibr_per_function_parity_emulated(emu_lib, name, N):
    parity.check_ibr_per_function(emu_lib, name, N=N)


def test_ibr_solve_emulated(emu_lib):
    parity.check_ibr_solve(emu_lib, "B", B=1, N=10, ibr_iter=2)


def test_ibr_reference_scenarios_emulated(emu_lib):
    """test/problem/solver_methods.jl:187-315 through the device path: 1-player best response = the optimal-control solve
    (res < 1e-6), 2-player IBR leaves the loose 5e-2 residual the reference asserts."""
    import algames_b200 as ab
    N, dt = 20, 0.1
    for mdl, p, outer, inner, ibr_iter, tol in [(ab.DoubleIntegratorGame, 1, 1, 1, 1, 1e-6), (ab.UnicycleGame, 1, 7, 20, 1, 1e-6),
                                                 (ab.DoubleIntegratorGame, 2, 1, 1, 100, 5e-2)]:
        model = mdl(p=p)
        x0 = {1: [1.0, 1.0, 0.0, 0.9], 2: [1.0, 2.0, 1.0, 2.0, 0, 0, 0.9, 0.9]}[p]
        obj = ab.GameObjective([np.ones(4)] * p, [0.5 * np.ones(2)] * p, [np.zeros(4)] * p, [-np.ones(2)] * p, N, model)
        con = ab.GameConstraintValues(ab.ProblemSize(N, model))
        opts = ab.Options(outer_iter=outer, inner_iter=inner, ls_iter=25, reg_0=1e-7, eps_dyn=1e-10, eps_opt=1e-10)
        prob = ab.GameProblem(N, dt, x0, model, opts, obj, con, lib_path=emu_lib)
        ab.ibr_newton_solve(prob, ab.IBROptions(ibr_iter=ibr_iter))
        assert np.abs(prob.core.res).sum() / prob.probsize.S < tol
        assert prob.stats.dyn_vio[-1].max < 1e-6


def test_solve_from_host_matches_staged_calls_emulated(emu_lib):
    """agb_solve_from_host (chunked pipeline) == set_instance_params + set_initial + newton_solve_batch, bit for bit."""
    import algames_b200 as ab
    model, N, dt, obj, con, opts, x0, xf = ab.workloads.config_b(batch=5, N=8)
    gb = ab.GameBatch(model, N, dt, obj, con, 5, lib_path=emu_lib)
    gb.set_instance_params(x0=x0)
    Z0, L0 = gb.random_initial()
    ref = gb.newton_solve(opts)
    got = gb.solve_from_host(opts, x0, Z0, L0)
    for k in ("Z", "L", "conlam", "conmu", "stats", "status"):
        assert np.array_equal(ref[k], got[k]), k
    gb.close()


@pytest.mark.parametrize("model_name,p,N", [("double_integrator", 1, 2), ("unicycle", 1, 3), ("bicycle", 4, 5), ("unicycle", 2, 2)])
def test_edge_shapes_emulated(emu_lib, model_name, p, N):
    parity.check_edge_shape(emu_lib, model_name, p, N)


@pytest.mark.parametrize("name,N", [("B", 10), ("D", 8)])
def test_per_function_parity_high_penalties_emulated(emu_lib, name, N):
    parity.check_per_function(emu_lib, name, seed=5, N=N, mu_exp=(4, 7))


def test_bulk_parity_vs_c_oracle_emulated(emu_lib):
    """The bulk comparison of tests/test_gpu_parity.py::test_bulk_parity_vs_c_oracle on a handful of instances."""
    parity.check_bulk_vs_c_oracle(emu_lib, "B", 6, 20, report=lambda s: None)
    parity.check_bulk_vs_c_oracle(emu_lib, "D", 2, 12, resolves=2, report=lambda s: None)


def test_working_set_too_large_is_rejected_emulated(emu_lib):
    import algames_b200 as ab
    model = ab.UnicycleGame(p=4)
    N = 200          # (N = 120 fits since the big layout keeps duals and multipliers in global memory)
    obj = ab.GameObjective([np.ones(4)] * 4, [np.ones(2)] * 4, [np.zeros(4)] * 4, [np.zeros(2)] * 4, N, model)
    con = ab.GameConstraintValues(ab.ProblemSize(N, model))
    ab.add_collision_avoidance(con, 0.1)
    with pytest.raises(ab.AlgamesError, match="shared memory"):
        ab.GameBatch(model, N, 0.1, obj, con, 1, lib_path=emu_lib)


# ---- round-2 boundary features on the emulator -------------------------------------------------------------------------
@pytest.mark.parametrize("name,N", [("A", None), ("A'", 10), ("B", 8), ("E", 8)])
def test_violation_vectors_emulated(emu_lib, name, N):
    parity.check_violation_vectors(emu_lib, name, N)


def test_ibr_history_emulated(emu_lib):
    parity.check_ibr_history(emu_lib, "B", N=8, ibr_iter=2)
    parity.check_ibr_history(emu_lib, "A", N=None, ibr_iter=1)


def test_status_codes_emulated(emu_lib):
    parity.check_status_codes(emu_lib)


def test_local_gather_emulated(emu_lib, monkeypatch):
    monkeypatch.setenv("AGB_EMU_DEVICES", "3")
    parity.check_local_gather(emu_lib, ndev=3, total=7)


def test_batch_requires_equal_options_emulated(emu_lib):
    import algames_b200 as ab
    model, N, dt, obj, con, opts, x0, _ = ab.workloads.config_b(batch=2, N=6)
    a = ab.GameProblem(N, dt, x0[0], model, ab.Options(), obj, con, lib_path=emu_lib)
    b = ab.GameProblem(N, dt, x0[1], model, ab.Options(inner_iter=3), obj, con, lib_path=emu_lib)
    with pytest.raises(ValueError, match="Options"):
        ab.newton_solve([a, b])
    b.opts = ab.Options()
    ab.newton_solve([a, b])
    assert a.status == b.status == "converged"
    assert a.stats.dyn_vio[-1].vio.shape == (N - 1,) and a.stats.opt_vio[-1].vio.shape == (N,)
    assert a.stats.dyn_vio[-1].vio.max() == a.stats.dyn_vio[-1].max
    # the cached single-problem handle re-pushes the objective (xf edited after the first solve)
    ab.newton_solve(a)
    z_before = a.pdtraj.X[-1].copy()
    a.game_obj.xf[0][:] = [0.3, 0.3, 0.0, 0.0]
    ab.newton_solve(a)
    assert np.abs(a.pdtraj.X[-1] - z_before).max() > 1e-3


# ---- band solver: QuadrotorGame, 3-D constraints, AGB_SOLVER_BAND, singular fallback -------------------------------------
@pytest.mark.parametrize("name,N,kw", [("Q", 5, {"p": 1}), ("Q", 4, {"p": 2}), ("B", 6, {}), ("A", None, {}), ("E", 6, {})])
def test_band_per_function_parity_emulated(emu_lib, name, N, kw):
    if name == "A":
        import algames_b200 as ab
        ab.workloads.CONFIGS["A_"] = lambda batch=2: tuple(np.tile(v, (batch, 1)) if i == 6 else v for i, v in enumerate(ab.workloads.config_a()))
        parity.check_band_per_function(emu_lib, "A_", seed=3)
    else:
        parity.check_band_per_function(emu_lib, name, seed=3, N=N, **kw)


def test_band_double_integrator_3d_emulated(emu_lib):
    # DoubleIntegratorGame(d = 3) + add_spherical_collision_avoidance! (test/constraints/constraints_methods.jl:30-57)
    parity.check_band_per_function(emu_lib, "B3", seed=2, N=5)
    parity.check_band_solve_vs_oracle(emu_lib, "B3", B=1, N=6)


def test_band_quadrotor_solve_emulated(emu_lib):
    parity.check_band_solve_vs_oracle(emu_lib, "Q", B=1, N=5, p=1)
    parity.check_band_solve_vs_oracle(emu_lib, "Q", B=1, N=4, p=2)


def test_band_equals_structured_emulated(emu_lib):
    parity.check_band_equals_structured(emu_lib, "B", B=3, N=8)
    parity.check_band_equals_structured(emu_lib, "D", B=2, N=8)


def test_singular_fallback_emulated(emu_lib, monkeypatch):
    parity.check_singular_fallback(emu_lib, monkeypatch)


def test_mpc_fused_loop_equals_stepwise_emulated(emu_lib):
    """agb_mpc_run (the receding-horizon loop inside the solve kernel) == the step-wise loop, bit for bit, on the CTA emulator."""
    parity.check_mpc_fused_equals_stepwise(emu_lib, "D", B=3, N=10, resolves=4)
    parity.check_mpc_fused_equals_stepwise(emu_lib, "D", B=2, N=8, resolves=3, force_layout=2)


@pytest.mark.parametrize("name,N,partial", [("B", 6, True), ("C", 5, False)])
def test_active_set_analysis_emulated(emu_lib, name, N, partial):
    """src/active_set/*.jl on the device kernels (CTA emulator) vs the oracle: bordered residual / Jacobian, masks, null space."""
    parity.check_active_set_analysis(emu_lib, name, N=N, partial=partial)


def test_active_set_reference_nullspace_test_emulated(emu_lib):
    """test/active_set/active_set_methods.jl:96-127 through the host mirror on the CTA emulator: at the zero initial iterate every
    collision gradient vanishes (exactly zero border rows) and null.mat still has the reference's (N−1)p columns."""
    import algames_b200 as ab
    rng = np.random.default_rng(0)
    N, dt, p = 6, 0.1, 3
    model = ab.UnicycleGame(p=p)
    ps = ab.ProblemSize(N, model)
    obj = ab.GameObjective([rng.random(4) for _ in range(p)], [rng.random(2) for _ in range(p)],
                           [(i + 1) * np.ones(4) for i in range(p)], [2 * (i + 1) * np.ones(2) for i in range(p)], N, model)
    con = ab.GameConstraintValues(ps)
    ab.add_collision_avoidance(con, 1.0)
    prob = ab.GameProblem(N, dt, rng.random(model.n), model, ab.Options(), obj, con, lib_path=emu_lib)
    asc = ab.ActiveSetCore(ps)
    null = ab.update_nullspace(asc, prob)
    assert null.mat.shape == (ps.S + (N - 1) * p * (p - 1), (N - 1) * p)
    assert len(null.vec) == (N - 1) * p and len(null.vec[0]) == ps.S + (N - 1) * p * (p - 1)
    assert all(abs(np.mean(np.abs(v)) - 1.0) < 1e-12 for v in null.vec)


@pytest.mark.parametrize("name,B,N,kw", [("Q", 1, 4, {"p": 2}), ("Q", 1, 5, {"p": 1}), ("B", 2, 8, {}), ("C", 1, 6, {}), ("B3", 1, 5, {})])
def test_band_window_equals_global_emulated(emu_lib, monkeypatch, name, B, N, kw):
    """The shared-memory elimination window of the band solver == the global-memory elimination, bit for bit (CTA emulator)."""
    parity.check_band_window_equals_global(emu_lib, monkeypatch, name, B=B, N=N, **kw)
