"""Parity checks shared by the CPU tier (CTA emulator, tests/emu) and the GPU tier (real libalgames_b200.so).
Every check drives the library through the C ABI (ctypes) and compares with the NumPy oracle on the same inputs."""
import numpy as np

import algames_b200 as ab
import oracle.algames_oracle as O

# tolerances (floating point, FP64): per-function results on identical inputs, and converged solutions
TOL_FUNC = 1e-10       # relative, residual / Jacobian / Newton step vs the oracle
TOL_SOLVE = 1e-6       # absolute, primal/dual trajectories and violations of a full newton_solve (north_star bound)


def small_config(name, batch=1, N=None):
    W = ab.workloads
    if name == "A":
        return W.config_a()
    if name == "A'":
        return W.config_a_prime()
    kw = {"batch": batch}
    if N is not None:
        kw["N"] = N
    return W.CONFIGS[name](**kw)


def oracle_problem(model, N, dt, obj, con, opts, x0, xf_joint=None):
    prob = ab.GameProblem(N, dt, x0, model, opts, obj, con, lib_path="unused")
    xfb = None
    if xf_joint is not None:
        p = model.p
        xfb = np.stack([np.asarray(xf_joint)[[i + c * p for c in range(model.ni[0])]] for i in range(p)])
    return O.problem_from_spec(ab.spec_of(prob), xf=xfb)


def random_state(oprob, x0, rng, mu_exp=(0, 3)):
    pd = oprob.pdtraj
    pd.X[:] = rng.normal(size=pd.X.shape); pd.U[:] = rng.normal(size=pd.U.shape); pd.du[:] = rng.normal(size=pd.du.shape)
    pd.X[0] = x0
    for _, _, cv in oprob.game_con.all_convals():
        cv.mu[:] = 10 ** rng.uniform(mu_exp[0], mu_exp[1], cv.mu.shape)
        cv.lam[:] = rng.random(cv.lam.shape) * (rng.random(cv.lam.shape) > 0.5)
    lam, mu = O.pack_multipliers(oprob)
    Z = np.concatenate([pd.X, pd.U], axis=1)
    return Z, pd.du.copy(), lam, mu


def check_per_function(lib_path, name, seed=0, reg=1e-3, N=None, mu_exp=(0, 3)):
    """residual!, residual_jacobian!, Δtraj solve, line search, update_traj!, Δ_step, evaluate!, dual/penalty update,
    active set — one random (far from converged) iterate with random multipliers and penalties 10^mu_exp (the reference
    runs its penalties up to ρ_max = 1e7, options.jl:59)."""
    model, N, dt, obj, con, opts, x0, xf = small_config(name, 2, N)
    B = 2
    x0 = np.tile(x0[:1], (B, 1)) if x0.shape[0] < B else x0[:B]
    xf = None if xf is None else xf[:B]
    rng = np.random.default_rng(seed)
    gb = ab.GameBatch(model, N, dt, obj, con, B, lib_path=lib_path)
    gb.set_instance_params(x0=x0, xf=xf)
    oprobs, Zs, Ls, lams, mus = [], [], [], [], []
    for b in range(B):
        op = oracle_problem(model, N, dt, obj, con, opts, x0[b], None if xf is None else xf[b])
        Z, L, lam, mu = random_state(op, x0[b], rng, mu_exp)
        oprobs.append(op); Zs.append(Z); Ls.append(L); lams.append(lam); mus.append(mu)
    has_con = lams[0].size > 0
    gb.set_initial(np.stack(Zs), np.stack(Ls), np.stack(lams) if has_con else None, np.stack(mus) if has_con else None)
    res, norms = gb.residual()
    J = gb.residual_jacobian_dense(reg, reg)
    d = gb.kkt_solve(reg, reg)
    alpha, j = gb.line_search(opts, reg)
    c = gb.evaluate()
    act = gb.active_set(opts.active_set_tolerance)
    for b, op in enumerate(oprobs):
        pd = op.pdtraj
        ores = O.residual(op, pd).copy()
        scale = np.abs(ores).max()
        assert np.abs(res[b] - ores).max() <= TOL_FUNC * scale
        rec = O.record(op, pd, 0.0, 1)
        assert np.allclose(norms[b], [rec.res, rec.dyn, rec.con, rec.sta, rec.opt], rtol=1e-12, atol=1e-12)
        op.opts.reg.set(reg)
        O.residual(op, pd)
        Jo = O.residual_jacobian(op, pd)
        assert np.abs(J[b] - Jo).max() <= TOL_FUNC * np.abs(Jo).max()
        ref = -np.linalg.solve(Jo, ores)
        # forward error: 1e-9 relative; beyond cond(J) ~ 1e10 (penalties >= 1e4) dense LU itself is only good to
        # cond·ε, so the bound becomes 1e-4·cond(J)·ε — four orders below what backward stability guarantees
        tol_fwd = 1e-9 if mu_exp[1] <= 3 else max(1e-9, 1e-4 * np.linalg.cond(Jo) * np.finfo(float).eps)
        assert np.abs(d[b] - ref).max() <= tol_fwd * np.abs(ref).max()
        # the step solves the reference's linear system at least as well as dense LU does
        assert np.abs(Jo @ d[b] + ores).max() <= 10 * max(np.abs(Jo @ ref + ores).max(), 1e-12 * scale)
        # line search from the same step
        O.set_traj(op.core, op.dpdtraj, d[b])
        op.pdtraj_trial = op.pdtraj.copy()          # x_1 = x0 on the trial trajectory (solver_methods.jl:14)
        res_norm = np.abs(ores).sum() / len(ores)
        op.core.res[:] = ores
        a_o, j_o = O.line_search(op, res_norm)
        assert j[b] == j_o and alpha[b] == a_o
        # constraint values / active set
        if has_con:
            op.game_con.evaluate(pd.X, pd.U)
            vals, _ = pack_vals(op)
            assert np.allclose(c[b], vals, rtol=1e-13, atol=1e-13)
            lam_o, _ = O.pack_multipliers(op)
            assert np.array_equal(act[b], (vals >= -opts.active_set_tolerance) | (lam_o > 0))
    # trial residual (regularize_residual!) at alpha = 0.25
    rt, _ = gb.residual(reg, reg, 0.25)
    for b, op in enumerate(oprobs):
        O.update_traj(op.pdtraj_trial, op.pdtraj, 0.25, op.dpdtraj)
        O.residual(op, op.pdtraj_trial)
        O.regularize_residual(op, op.pdtraj_trial, op.pdtraj)
        assert np.abs(rt[b] - op.core.res).max() <= TOL_FUNC * np.abs(op.core.res).max()
    # update_traj! + Δ_step
    dstep = gb.update_traj(alpha)
    Z, L, _, _ = gb.get_state()
    n = model.n
    for b, op in enumerate(oprobs):
        O.update_traj(op.pdtraj, op.pdtraj, alpha[b], op.dpdtraj)
        assert abs(dstep[b] - O.delta_step(op.dpdtraj, alpha[b])) <= 1e-12 * max(1.0, abs(dstep[b]))
        assert np.abs(Z[b, :, :n] - op.pdtraj.X).max() <= 1e-12 * max(1.0, np.abs(op.pdtraj.X).max())
        assert np.abs(Z[b, :-1, n:] - op.pdtraj.U[:-1]).max() <= 1e-12 * max(1.0, np.abs(op.pdtraj.U).max())
        assert np.abs(L[b] - op.pdtraj.du).max() <= 1e-12 * max(1.0, np.abs(op.pdtraj.du).max())
    # dual ascent + penalty schedule at the new iterate
    if has_con:
        gb.dual_update(opts); gb.penalty_update(opts)
        _, _, cl, cm = gb.get_state()
        for b, op in enumerate(oprobs):
            op.game_con.evaluate(op.pdtraj.X, op.pdtraj.U)
            op.game_con.dual_update(); op.game_con.penalty_update()
            lam_o, mu_o = O.pack_multipliers(op)
            assert np.allclose(cl[b], lam_o, rtol=1e-13, atol=1e-13) and np.allclose(cm[b], mu_o, rtol=1e-13)
        gb.reset(opts)
        _, _, cl, cm = gb.get_state()
        assert (cl == 0).all() and (cm == opts.rho_0).all()
    gb.close()


def pack_vals(op):
    ps, gc = op.probsize, op.game_con
    out = []
    for k in range(1, ps.N):
        row = []
        for i in range(ps.p):
            for cv in gc.state_conval[i]:
                row.append(cv.vals[k - 1])
        for cv in gc.control_conval:
            row.append(cv.vals[k - 1])
        out.append(np.concatenate(row) if row else np.zeros(0))
    return np.array(out), None


def check_rollout(lib_path, name):
    model, N, dt, obj, con, opts, x0, xf = small_config(name, 1)
    gb = ab.GameBatch(model, N, dt, obj, con, 1, lib_path=lib_path)
    gb.set_instance_params(x0=x0[:1])
    rng = np.random.default_rng(3)
    Z0 = rng.normal(size=(1, N, model.n + model.m)); L0 = np.zeros((1, model.p, N - 1, model.n))
    gb.set_initial(Z0, L0)
    gb.rollout()
    Z, _, _, _ = gb.get_state()
    op = oracle_problem(model, N, dt, obj, con, opts, x0[0])
    op.pdtraj.X[:] = Z0[0, :, :model.n]; op.pdtraj.U[:] = Z0[0, :, model.n:]; op.pdtraj.X[0] = x0[0]
    O.rollout_rk3(op.model, op.pdtraj)
    assert np.abs(Z[0, :, :model.n] - op.pdtraj.X).max() <= 1e-12 * max(1.0, np.abs(op.pdtraj.X).max())
    gb.close()


def solve_batch(lib_path, name, B, N=None, opts_override=None):
    model, N, dt, obj, con, opts, x0, xf = small_config(name, B, N)
    if x0.shape[0] < B:
        x0 = np.tile(x0[:1], (B, 1)); x0[:, :2 * model.p] += 0.01 * np.arange(B)[:, None]
    if opts_override:
        for k, v in opts_override.items():
            setattr(opts, k, v)
    gb = ab.GameBatch(model, N, dt, obj, con, B, lib_path=lib_path)
    gb.set_instance_params(x0=x0, xf=xf)
    Z0, L0 = gb.random_initial(opts.amplitude_init, opts.seed)
    gb.set_history(opts.outer_iter * opts.inner_iter + 1)
    out = gb.newton_solve(opts)
    out["hist"], out["hist_count"] = gb.get_history()
    return (model, N, dt, obj, con, opts, x0, xf), gb, Z0, L0, out


def check_solve_vs_oracle(lib_path, name, B=2, N=None, which=None):
    """Full newton_solve! on the same initial iterate: converged trajectories, duals, AL multipliers, final record."""
    cfg, gb, Z0, L0, out = solve_batch(lib_path, name, B, N)
    model, N, dt, obj, con, opts, x0, xf = cfg
    for b in (range(B) if which is None else which):
        op = oracle_problem(model, N, dt, obj, con, opts, x0[b], None if xf is None else xf[b])
        O.newton_solve(op, Z0=Z0[b], L0=L0[b])
        last, st = op.stats[-1], out["stats"][b]
        Zo = np.concatenate([op.pdtraj.X, op.pdtraj.U], axis=1)
        assert np.abs(out["Z"][b] - Zo).max() < TOL_SOLVE
        assert np.abs(out["L"][b] - op.pdtraj.du).max() < TOL_SOLVE * max(1.0, np.abs(op.pdtraj.du).max())
        assert np.allclose(st[:5], [last.res, last.dyn, last.con, last.sta, last.opt], atol=TOL_SOLVE)
        assert (out["status"][b] == 0) == op.converged
        # the whole Statistics history (statistics.jl:44-57): one record per inner iteration + the final one
        cnt = int(out["hist_count"][b])
        assert cnt == len(op.stats)
        ho = np.array([[r.outer, r.res, r.dyn, r.con, r.sta, r.opt, r.delta] for r in op.stats])
        hd = out["hist"][b, :cnt, :7]
        assert np.array_equal(hd[:, 0], ho[:, 0])
        assert np.allclose(hd[:, 1:], ho[:, 1:], rtol=1e-6, atol=TOL_SOLVE), np.abs(hd[:, 1:] - ho[:, 1:]).max(axis=0)
        lam_o, mu_o = O.pack_multipliers(op)
        if lam_o.size:
            assert np.allclose(out["conlam"][b], lam_o, atol=TOL_SOLVE * max(1.0, np.abs(lam_o).max()))
            assert np.allclose(out["conmu"][b], mu_o, rtol=1e-12)
    gb.close()
    return out


def check_solution_properties(cfg, out, gb_factory):
    """Size-independent properties of a solved batch: the reported record is reproduced by re-evaluating the residual
    at the returned iterate, converged instances meet every tolerance, and a solve restarted from the solution with the
    returned multipliers stops immediately."""
    model, N, dt, obj, con, opts, x0, xf = cfg
    gb = gb_factory()
    gb.set_instance_params(x0=x0, xf=xf)
    gb.set_initial(out["Z"], out["L"], out["conlam"] if out["conlam"].size else None, out["conmu"] if out["conmu"].size else None)
    _, norms = gb.residual(want_res=False)
    assert np.allclose(norms, out["stats"][:, :5], rtol=1e-9, atol=1e-12)
    conv = out["status"] == 0
    eps = np.array([opts.eps_dyn, opts.eps_con, opts.eps_sta, opts.eps_opt])
    assert (out["stats"][conv][:, 1:5] < eps).all()
    assert np.isfinite(out["Z"]).all() and np.isfinite(out["L"]).all()
    return gb, conv


def check_golden(lib_path, path):
    """Committed golden vector (tests/golden/make_golden.py): same inputs -> same converged solution and Newton count."""
    g = np.load(path)
    name, B = str(g["config"]), int(g["x0"].shape[0])
    model, N, dt, obj, con, opts, x0, xf = small_config(name, B, int(g["N"]))
    gb = ab.GameBatch(model, N, dt, obj, con, B, lib_path=lib_path)
    gb.set_instance_params(x0=g["x0"], xf=(g["xf"] if "xf" in g.files else None))
    gb.set_initial(g["Z0"], g["L0"])
    out = gb.newton_solve(opts)
    assert np.abs(out["Z"] - g["Z"]).max() < TOL_SOLVE, path
    assert np.abs(out["L"] - g["L"]).max() < TOL_SOLVE * max(1.0, np.abs(g["L"]).max()), path
    assert np.allclose(out["stats"][:, :5], g["stats"], atol=TOL_SOLVE), path
    assert np.array_equal(out["status"] == 0, g["converged"]), path
    assert np.array_equal(out["stats"][:, 6].astype(int), g["n_newton"]), path
    gb.close()


def check_pivot_fallback(lib_path):
    """The kernel's gain-system solver on systems [S | RHS]: diagonally dominant S (threshold test passes, no row
    exchange), S with zero / tiny diagonal entries but well conditioned (must fall back to full partial pivoting),
    and a singular S (must report failure)."""
    for p in (2, 3, 4):
        model = ab.DoubleIntegratorGame(p=p)
        N = 4
        obj = ab.GameObjective([np.ones(4)] * p, [np.ones(2)] * p, [np.zeros(4)] * p, [np.zeros(2)] * p, N, model)
        con = ab.GameConstraintValues(ab.ProblemSize(N, model))
        m, n = model.m, model.n
        B = 6
        rng = np.random.default_rng(p)
        aug = rng.normal(size=(B, m, m + n + 1))
        aug[0, :, :m] += 10 * np.eye(m)                       # dominant diagonal
        aug[1, :, :m] += 10 * np.eye(m)
        for b in (2, 3):                                      # zero diagonal, strong off-diagonal (a cyclic shift)
            aug[b, :, :m] = 0.1 * rng.normal(size=(m, m)) + 5 * np.roll(np.eye(m), 1, axis=1)
            aug[b, np.arange(m), np.arange(m)] = 0.0 if b == 2 else 1e-14
        aug[4, :, :m] = np.diag(np.r_[1e-3, np.ones(m - 1)]) + 0.5 * np.roll(np.eye(m), 1, axis=0)   # weak leading pivot
        aug[5, :, :m] = np.ones((m, m))                       # singular
        gb = ab.GameBatch(model, N, 0.1, obj, con, B, lib_path=lib_path)
        out, ok = gb.debug_gain_solve(aug)
        gb.close()
        for b in range(5):
            ref = np.linalg.solve(aug[b, :, :m], aug[b, :, m:])
            assert ok[b], (p, b)
            assert np.abs(out[b, :, m:] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()) * np.linalg.cond(aug[b, :, :m]), (p, b)
        assert not ok[5] or not np.isfinite(out[5]).all()


# ---------------------------------------------------------------------------------------------------------------
# Iterative best response
# ---------------------------------------------------------------------------------------------------------------
def check_ibr_per_function(lib_path, name, seed=4, reg=1e-3, N=None):
    """ibr_residual! (+ proximal term) and the masked Newton step of every player on a random iterate."""
    model, N, dt, obj, con, opts, x0, xf = small_config(name, 1, N)
    rng = np.random.default_rng(seed)
    gb = ab.GameBatch(model, N, dt, obj, con, 1, lib_path=lib_path)
    gb.set_instance_params(x0=x0[:1], xf=None if xf is None else xf[:1])
    op = oracle_problem(model, N, dt, obj, con, opts, x0[0], None if xf is None else xf[0])
    Z, L, lam, mu = random_state(op, x0[0], rng)
    has_con = lam.size > 0
    gb.set_initial(Z[None], L[None], lam[None] if has_con else None, mu[None] if has_con else None)
    op.opts.reg.set(reg)
    for i in range(1, model.p + 1):
        vm, hm = O.vertical_mask(op.core, i), O.horizontal_mask(op.core, i)
        res, norms = gb.ibr_residual(i - 1)
        ores = O.ibr_residual(op, op.pdtraj, i).copy()
        masked = np.zeros_like(ores); masked[vm] = ores[vm]
        assert np.abs(res[0] - masked).max() <= TOL_FUNC * np.abs(masked).max()
        stats0 = list(op.stats)
        rec = O.record_ibr(op, op.pdtraj, 0.0, 1, i)
        op.stats = stats0
        assert np.allclose(norms[0], [np.abs(ores[vm]).sum() / len(vm), rec.dyn, rec.con, rec.sta, rec.opt], rtol=1e-12, atol=1e-12)
        d = gb.ibr_kkt_solve(i - 1, reg, reg)[0]
        O.ibr_residual(op, op.pdtraj, i)
        J = O.ibr_residual_jacobian(op, op.pdtraj, i)
        ref = np.zeros(op.probsize.S)
        ref[hm] = -np.linalg.solve(J[np.ix_(vm, hm)], ores[vm])
        assert np.abs(d - ref).max() <= 1e-9 * np.abs(ref).max(), (i, np.abs(d - ref).max() / np.abs(ref).max())
        # trial residual with the proximal term
        rt, _ = gb.ibr_residual(i - 1, reg, reg, 0.5)
        O.set_traj(op.core, op.dpdtraj, d)
        op.pdtraj_trial = op.pdtraj.copy()
        O.update_traj(op.pdtraj_trial, op.pdtraj, 0.5, op.dpdtraj)
        O.ibr_residual(op, op.pdtraj_trial, i)
        O.regularize_ibr_residual(op, op.pdtraj_trial, op.pdtraj, i)
        mt = np.zeros_like(ores); mt[vm] = op.core.res[vm]
        assert np.abs(rt[0] - mt).max() <= TOL_FUNC * np.abs(mt).max()
    gb.close()


def check_ibr_solve(lib_path, name, B=1, N=None, ibr_iter=3, opts_override=None):
    """ibr_newton_solve! on the same initial iterate: trajectories, Newton counts, final full-game record."""
    model, N, dt, obj, con, opts, x0, xf = small_config(name, B, N)
    if x0.shape[0] < B:
        x0 = np.tile(x0[:1], (B, 1))
    for k, v in (opts_override or {}).items():
        setattr(opts, k, v)
    gb = ab.GameBatch(model, N, dt, obj, con, B, lib_path=lib_path)
    gb.set_instance_params(x0=x0, xf=xf)
    Z0, L0 = gb.random_initial(opts.amplitude_init, opts.seed)
    out = gb.ibr_newton_solve(opts, ab.IBROptions(ibr_iter=ibr_iter))
    for b in range(B):
        op = oracle_problem(model, N, dt, obj, con, opts, x0[b], None if xf is None else xf[b])
        O.ibr_newton_solve(op, Z0=Z0[b], L0=L0[b], ibr_iter=ibr_iter)
        Zo = np.concatenate([op.pdtraj.X, op.pdtraj.U], axis=1)
        assert np.abs(out["Z"][b] - Zo).max() < TOL_SOLVE
        assert np.abs(out["L"][b] - op.pdtraj.du).max() < TOL_SOLVE * max(1.0, np.abs(op.pdtraj.du).max())
        assert int(out["stats"][b, 6]) == op.n_newton and int(out["stats"][b, 7]) == op.ibr_sweeps
        O.residual(op, op.pdtraj)
        assert abs(out["stats"][b, 0] - np.abs(op.core.res).sum() / op.probsize.S) < TOL_SOLVE
    gb.close()
    return out


# ---------------------------------------------------------------------------------------------------------------
# Bulk parity at BASELINE sizes: every instance of a batch against the C restatement of the reference algorithm
# (oracle/algames_oracle.c: explicit KKT Jacobian + pivoted band LU per Newton step, cross-checked against the NumPy
# oracle in tests/test_c_oracle.py)
# ---------------------------------------------------------------------------------------------------------------
def c_oracle_solve(cfg, x0, xf, Z0, L0, opts, conlam=None, conmu=None, hist_max=0, nthreads=0):
    from oracle import c_oracle
    model, N, dt, obj, con = cfg[0], cfg[1], cfg[2], cfg[3], cfg[4]
    p, J, B = model.p, ab.problem._joint, x0.shape[0]
    desc = ab.problem._make_desc(model, N, dt, obj, con)
    tile = lambda v: np.tile(v, (B, 1))
    xfj = tile(J(obj.xf, p, 4)) if xf is None else xf
    out, _ = c_oracle.newton_solve(desc, opts.to_c(), x0, xfj, tile(J(obj.Q, p, 4)), tile(J(obj.R, p, 2)), tile(J(obj.uf, p, 2)),
                                   Z0, L0, nthreads=nthreads, conlam=conlam, conmu=conmu, hist_max=hist_max)
    return out


def compare_bulk(dev, ref, tol=TOL_SOLVE):
    """Instance-by-instance comparison of a device solve with the C oracle's on the same inputs.  An instance FORKS when
    its iteration trace differs (status, Newton steps, outer iterations or number of records): round-off differences
    between the two linear solvers flipped a discrete decision (line-search step, active set, exit test), after which the
    two runs are different — equally valid — executions of the same algorithm.  Non-forked instances must agree to `tol`:
    converged ones on trajectories, duals and multipliers; the others on their final record (relative)."""
    B = dev["status"].shape[0]
    st_d, st_r = dev["stats"], ref["stats"]
    fork = (dev["status"] != ref["status"]) | (st_d[:, 6] != st_r[:, 6]) | (st_d[:, 7] != st_r[:, 7])
    if "hist_count" in dev and "hist_count" in ref:
        fork |= dev["hist_count"] != ref["hist_count"]
    same = ~fork
    conv = (ref["status"] == 0) & same
    nonc = (ref["status"] != 0) & same
    rep = {"n": int(B), "forked": int(fork.sum()), "converged": int((dev["status"] == 0).sum()), "ref_converged": int((ref["status"] == 0).sum()),
           "nonconverged_compared": int(nonc.sum()), "fork_idx": np.nonzero(fork)[0][:16].tolist()}
    ax = tuple(range(1, dev["Z"].ndim))
    zerr = np.abs(dev["Z"] - ref["Z"]).max(axis=ax)
    lscale = np.maximum(1.0, np.abs(ref["L"]).max(axis=(1, 2, 3)))
    lerr = np.abs(dev["L"] - ref["L"]).max(axis=(1, 2, 3)) / lscale
    rec_rel = (np.abs(st_d[:, :5] - st_r[:, :5]) / np.maximum(np.abs(st_r[:, :5]), 1e-9)).max(axis=1)
    rec_abs = np.abs(st_d[:, :5] - st_r[:, :5]).max(axis=1)
    if dev["conlam"].size:
        cscale = np.maximum(1.0, np.abs(ref["conlam"]).max(axis=(1, 2)))
        cerr = np.abs(dev["conlam"] - ref["conlam"]).max(axis=(1, 2)) / cscale
        mu_same = np.all(np.isclose(dev["conmu"], ref["conmu"], rtol=1e-12, atol=0), axis=(1, 2))
    else:
        cerr, mu_same = np.zeros(B), np.ones(B, bool)
    pick = lambda v, msk: float(v[msk].max()) if msk.any() else 0.0
    rep.update(worst_Z_converged=pick(zerr, conv), worst_L_converged=pick(lerr, conv), worst_conlam_converged=pick(cerr, conv),
               worst_record_abs_converged=pick(rec_abs, conv), worst_record_rel_nonconverged=pick(rec_rel, nonc),
               worst_Z_nonconverged=pick(zerr, nonc), worst_Z_forked=pick(zerr, fork))
    if "hist" in dev and "hist" in ref:               # whole convergence trace of the non-forked instances
        worst = 0.0
        for b in np.nonzero(same)[0]:
            c = int(ref["hist_count"][b]); c = min(c, dev["hist"].shape[1], ref["hist"].shape[1])
            hd, hr = dev["hist"][b, :c], ref["hist"][b, :c]
            assert np.array_equal(hd[:, [0, 7]], hr[:, [0, 7]]), b
            worst = max(worst, float((np.abs(hd[:, 1:7] - hr[:, 1:7]) / np.maximum(np.abs(hr[:, 1:7]), 1e-6)).max()))
        rep["worst_history_rel"] = worst
    rep["ok_converged"] = bool((zerr[conv] < tol).all() and (lerr[conv] < tol).all() and (cerr[conv] < tol).all() and mu_same[conv].all()
                               and (rec_abs[conv] < tol).all())
    rep["ok_nonconverged"] = bool(((rec_rel[nonc] < tol) | (rec_abs[nonc] < tol)).all() and mu_same[nonc].all())
    return rep


def check_bulk_vs_c_oracle(lib_path, name, B, N=None, resolves=0, max_fork_frac=0.0, disturbance_std=1e-3, report=print):
    """Cold solve of every instance of config `name` on the device and with the C oracle from the same initial iterate;
    then `resolves` MPC re-solves (shift = 1, multipliers carried, dual_reset = false), each compared on the device's own
    shifted iterate so that re-solve r is checked on identical inputs whatever happened before."""
    cfg = small_config(name, B, N)
    model, N, dt, obj, con, opts, x0, xf = cfg
    n = model.n
    gb = ab.GameBatch(model, N, dt, obj, con, B, lib_path=lib_path)
    gb.set_instance_params(x0=x0, xf=xf)
    Z0, L0 = gb.random_initial(opts.amplitude_init, opts.seed)
    hmax = opts.outer_iter * opts.inner_iter + 1
    gb.set_history(hmax)
    cold = ab.Options(**{**opts.to_dict(), "dual_reset": True})
    dev = gb.newton_solve(cold)
    dev["hist"], dev["hist_count"] = gb.get_history()
    ref = c_oracle_solve(cfg, x0, xf, Z0, L0, cold, hist_max=hmax)
    reps = [compare_bulk(dev, ref)]
    reps[0]["what"] = f"{name} cold B={B} N={N}"
    rng = np.random.default_rng(11)
    warm = ab.Options(**{**opts.to_dict(), "dual_reset": False, "shift": 1})
    for r in range(resolves):
        d = disturbance_std * rng.standard_normal((B, n))
        x0 = dev["Z"][:, 1, :n] + d
        gb.mpc_advance(1, d)
        Zs, Ls, cl, cm = gb.get_state()
        dev = gb.newton_solve(warm)
        dev["hist"], dev["hist_count"] = gb.get_history()
        ref = c_oracle_solve(cfg, x0, xf, Zs, Ls, warm, conlam=cl, conmu=cm, hist_max=hmax)
        reps.append(compare_bulk(dev, ref))
        reps[-1]["what"] = f"{name} warm re-solve {r + 1} B={B} N={N}"
    gb.close()
    for rep in reps:
        report("bulk parity: " + ", ".join(f"{k}={v if not isinstance(v, float) else format(v, '.2e')}" for k, v in rep.items()))
        assert rep["ok_converged"], rep
        assert rep["ok_nonconverged"], rep
        assert rep["forked"] <= max_fork_frac * rep["n"], rep
    return reps


def check_edge_shape(lib_path, model_name, p, N):
    """Smallest horizons (N = 2: one stage, no backward recursion), single player, and a 4-player bicycle game carrying every
    constraint type: full solve vs the NumPy oracle."""
    model = {"double_integrator": ab.DoubleIntegratorGame, "unicycle": ab.UnicycleGame, "bicycle": ab.BicycleGame}[model_name](p=p)
    rng = np.random.default_rng(p * 10 + N)
    obj = ab.GameObjective([1 + rng.random(4) for _ in range(p)], [0.1 + rng.random(2) for _ in range(p)],
                           [rng.normal(size=4) for _ in range(p)], [0.1 * rng.normal(size=2) for _ in range(p)], N, model)
    con = ab.GameConstraintValues(ab.ProblemSize(N, model))
    if p > 1:
        ab.add_collision_cost(obj, 0.5 * np.ones(p), 2.0 * np.ones(p))
        ab.add_collision_avoidance(con, 0.05)
    if model_name == "bicycle":
        ab.add_control_bound(con, np.r_[2 * np.ones(p), 0.5 * np.ones(p)], np.r_[-2 * np.ones(p), -np.inf * np.ones(p)])
        ab.add_state_bound(con, min(1, p - 1), 5 * np.ones(model.n), np.r_[-5 * np.ones(model.n - 2), -np.inf, -np.inf])
        ab.add_wall_constraint(con, [ab.Wall([0.0, -0.4], [1.0, -0.4], [0.0, -1.0])], min(2, p - 1))
        ab.add_circle_constraint(con, [1.0], [1.0], [0.2])
    x0 = rng.normal(size=model.n)
    opts = ab.Options()
    gb = ab.GameBatch(model, N, 0.1, obj, con, 1, lib_path=lib_path)
    gb.set_instance_params(x0=x0[None])
    Z0, L0 = gb.random_initial()
    out = gb.newton_solve(opts)
    prob = ab.GameProblem(N, 0.1, x0, model, opts, obj, con, lib_path=lib_path)
    op = O.problem_from_spec(ab.spec_of(prob))
    O.newton_solve(op, Z0=Z0[0], L0=L0[0])
    Zo = np.concatenate([op.pdtraj.X, op.pdtraj.U], axis=1)
    assert np.abs(out["Z"][0] - Zo).max() < TOL_SOLVE
    assert int(out["stats"][0, 6]) == op.n_newton and (out["status"][0] == 0) == op.converged
    gb.close()


# ---------------------------------------------------------------------------------------------------------------
# Round-2 boundary features
# ---------------------------------------------------------------------------------------------------------------
def check_violation_vectors(lib_path, name, N=None):
    """agb_violations vs the oracle's restatement of struct/violations.jl on a random iterate."""
    model, N, dt, obj, con, opts, x0, xf = small_config(name, 1, N)
    rng = np.random.default_rng(9)
    gb = ab.GameBatch(model, N, dt, obj, con, 1, lib_path=lib_path)
    gb.set_instance_params(x0=x0[:1], xf=None if xf is None else xf[:1])
    op = oracle_problem(model, N, dt, obj, con, opts, x0[0], None if xf is None else xf[0])
    Z, L, lam, mu = random_state(op, x0[0], rng)
    gb.set_initial(Z[None], L[None], lam[None] if lam.size else None, mu[None] if lam.size else None)
    got = gb.violations()
    want = O.violation_vectors(op, op.pdtraj)
    for g, w, nm in zip(got, want, ("dyn", "con", "sta", "opt")):
        assert g.shape == (1,) + w.shape, nm
        assert np.allclose(g[0], w, rtol=1e-12, atol=1e-12), (nm, np.abs(g[0] - w).max())
    _, norms = gb.residual(want_res=False)
    assert np.allclose([g.max() for g in got], norms[0, 1:5], rtol=1e-13, atol=0)      # .max of every vector = the record's maxima
    gb.close()


def check_ibr_history(lib_path, name="B", N=10, ibr_iter=2):
    """IBR history (statistics.jl:59-72): every record!(stats, …, k, i) of the sweeps, against the oracle's."""
    model, N, dt, obj, con, opts, x0, xf = small_config(name, 1, N)
    gb = ab.GameBatch(model, N, dt, obj, con, 1, lib_path=lib_path)
    gb.set_instance_params(x0=x0[:1], xf=None if xf is None else xf[:1])
    Z0, L0 = gb.random_initial(opts.amplitude_init, opts.seed)
    gb.set_history(2048)
    out = gb.ibr_newton_solve(opts, ab.IBROptions(ibr_iter=ibr_iter))
    hist, count = gb.get_history()
    op = oracle_problem(model, N, dt, obj, con, opts, x0[0], None if xf is None else xf[0])
    O.ibr_newton_solve(op, Z0=Z0[0], L0=L0[0], ibr_iter=ibr_iter)
    cnt = int(count[0])
    assert cnt == len(op.stats), (cnt, len(op.stats))
    ho = np.array([[r.outer, r.res, r.dyn, r.con, r.sta, r.opt, r.delta] for r in op.stats])
    hd = hist[0, :cnt]
    assert np.array_equal(hd[:, 0], ho[:, 0])
    assert np.allclose(hd[:, 1:7], ho[:, 1:], rtol=1e-6, atol=TOL_SOLVE), np.abs(hd[:, 1:7] - ho[:, 1:]).max(axis=0)
    assert set(hd[:, 8].astype(int)) == set(range(model.p)) and hd[:, 9].max() <= ibr_iter
    # switching the log off changes nothing in the result
    gb.set_history(0)
    gb.set_initial(Z0, L0)
    out2 = gb.ibr_newton_solve(opts, ab.IBROptions(ibr_iter=ibr_iter))
    assert np.array_equal(out["Z"], out2["Z"]) and np.array_equal(out["stats"][:, :8], out2["stats"][:, :8])
    gb.close()


def check_status_codes(lib_path):
    """The six per-instance outcomes of include/algames_b200.h, on the device and in the C oracle."""
    S = ab._capi
    model, N, dt, obj, con, opts, x0, xf = small_config("B", 4, 12)
    cfg = (model, N, dt, obj, con, opts, x0, xf)

    def run(o, x):
        gb = ab.GameBatch(model, N, dt, obj, con, x.shape[0], lib_path=lib_path)
        gb.set_instance_params(x0=x)
        Z0, L0 = gb.random_initial()
        out = gb.newton_solve(o)
        gb.close()
        ref = c_oracle_solve(cfg, x, None, Z0, L0, o)
        assert np.array_equal(out["status"], ref["status"]), (out["status"], ref["status"])
        return out["status"]

    assert (run(opts, x0) == S.CONVERGED).all()
    tight = ab.Options(**{**opts.to_dict(), "eps_opt": 1e-300, "eps_dyn": 1e-300, "outer_iter": 2, "inner_iter": 3, "delta_min": 0.0, "ls_iter": 40})
    assert (run(tight, x0) == S.MAX_OUTER).all()                       # inner_iter exhausted, tolerances unreachable
    stall = ab.Options(**{**opts.to_dict(), "eps_opt": 1e-300, "outer_iter": 1, "delta_min": 1e30})
    assert (run(stall, x0) == S.STALLED).all()                         # first step is "too small"
    nols = ab.Options(**{**opts.to_dict(), "eps_opt": 1e-300, "outer_iter": 1, "ls_iter": 1, "delta_min": 0.0})
    assert (run(nols, x0) == S.LINE_SEARCH_FAILED).all()               # no trial allowed: j == ls_iter at once
    xn = x0.copy(); xn[2, 1] = np.inf
    st = run(opts, xn)
    assert st[2] == S.NONFINITE and (np.delete(st, 2) == S.CONVERGED).all()


def check_local_gather(lib_path, ndev=2, total=5):
    """Single-process multi-GPU path of the C ABI (agb_create_sharded-style layout built from GameBatch handles): uneven
    contiguous shards, one push all-gather, every rank's buffer holds every rank's results."""
    model, N, dt, obj, con, opts, x0, xf = small_config("B", total, 10)
    from algames_b200 import distributed as D
    bounds = [D.shard_bounds(total, ndev, r) for r in range(ndev)]
    batches = [hi - lo for lo, hi in bounds]
    gbs, outs = [], []
    rng = np.random.default_rng(opts.seed)
    Z0 = opts.amplitude_init * rng.random((total, N, model.n + model.m)); L0 = opts.amplitude_init * rng.random((total, model.p, N - 1, model.n))
    for r, (lo, hi) in enumerate(bounds):
        gb = ab.GameBatch(model, N, dt, obj, con, hi - lo, device=r, lib_path=lib_path)
        gb.set_instance_params(x0=x0[lo:hi]); gb.set_initial(Z0[lo:hi], L0[lo:hi])
        gb.peer_init(ndev, r, batches)
        gbs.append(gb)
    ab.GameBatch.peer_connect_local(gbs)
    for gb in gbs:
        gb.newton_solve_async(opts, 0)
        gb.allgather()
    for gb in gbs:
        gb.allgather_wait()
    for gb in gbs:
        outs.append(gb.newton_solve(opts, want=("Z", "L", "stats", "status")))      # same solve again, host copies: the reference result
    for gb in gbs:
        ptr, offs = gb.gathered_view()
        assert ptr and len(offs) == ndev + 1 and offs[0] == 0
        for r in range(ndev):
            got = gb.unpack_gathered(r)
            for k in ("Z", "L", "stats", "status"):
                assert np.array_equal(got[k], outs[r][k]), (r, k)
    for gb in gbs:
        gb.close()


# ---------------------------------------------------------------------------------------------------------------
# Band solver (agb_band.cuh): QuadrotorGame / 3-D constraints, AGB_SOLVER_BAND on the planar configs, singular fallback
# ---------------------------------------------------------------------------------------------------------------
def check_band_per_function(lib_path, name, seed=0, reg=1e-3, N=None, **kw):
    """residual!, residual_jacobian!, Δtraj and evaluate! of the band solver on a random iterate vs the NumPy oracle."""
    W = ab.workloads
    model, N, dt, obj, con, opts, x0, xf = W.CONFIGS[name](batch=2, **({"N": N} if N else {}), **kw)
    B = 2
    rng = np.random.default_rng(seed)
    gb = ab.GameBatch(model, N, dt, obj, con, B, lib_path=lib_path, solver=ab._capi.SOLVER_BAND)
    gb.set_instance_params(x0=x0[:B], xf=None if xf is None else xf[:B])
    oprobs, Zs, Ls, lams, mus = [], [], [], [], []
    for b in range(B):
        op = oracle_problem(model, N, dt, obj, con, opts, x0[b], None if xf is None else xf[b])
        pd = op.pdtraj
        Z, L, lam, mu = random_state(op, x0[b], rng)
        if model.name == "quadrotor":                       # keep the random iterate in the model's sensible range
            pd.X[:] *= 0.3; pd.U[:] = 1.0 + 0.3 * pd.U; pd.X[0] = x0[b]
            Z = np.concatenate([pd.X, pd.U], axis=1)
        oprobs.append(op); Zs.append(Z); Ls.append(L); lams.append(lam); mus.append(mu)
    has_con = lams[0].size > 0
    gb.set_initial(np.stack(Zs), np.stack(Ls), np.stack(lams) if has_con else None, np.stack(mus) if has_con else None)
    res, norms = gb.residual()
    J = gb.residual_jacobian_dense(reg, reg)
    d = gb.kkt_solve(reg, reg)
    c = gb.evaluate() if has_con else None
    for b, op in enumerate(oprobs):
        pd = op.pdtraj
        ores = O.residual(op, pd).copy()
        assert np.abs(res[b] - ores).max() <= TOL_FUNC * np.abs(ores).max(), np.abs(res[b] - ores).max()
        rec = O.record(op, pd, 0.0, 1)
        assert np.allclose(norms[b], [rec.res, rec.dyn, rec.con, rec.sta, rec.opt], rtol=1e-12, atol=1e-12)
        op.opts.reg.set(reg)
        O.residual(op, pd)
        Jo = O.residual_jacobian(op, pd)
        assert np.abs(J[b] - Jo).max() <= 1e-9 * np.abs(Jo).max(), np.abs(J[b] - Jo).max() / np.abs(Jo).max()
        ref = -np.linalg.solve(Jo, ores)
        assert np.abs(d[b] - ref).max() <= 1e-8 * np.abs(ref).max(), np.abs(d[b] - ref).max() / np.abs(ref).max()
        if has_con:
            op.game_con.evaluate(pd.X, pd.U)
            vals, _ = pack_vals(op)
            assert np.allclose(c[b], vals, rtol=1e-13, atol=1e-13)
    # trial residual with the proximal term
    rt, _ = gb.residual(reg, reg, 0.25)
    for b, op in enumerate(oprobs):
        O.set_traj(op.core, op.dpdtraj, d[b])
        op.pdtraj_trial = op.pdtraj.copy()
        O.update_traj(op.pdtraj_trial, op.pdtraj, 0.25, op.dpdtraj)
        O.residual(op, op.pdtraj_trial)
        O.regularize_residual(op, op.pdtraj_trial, op.pdtraj)
        assert np.abs(rt[b] - op.core.res).max() <= TOL_FUNC * np.abs(op.core.res).max()
    gb.close()


def check_band_solve_vs_oracle(lib_path, name, B=2, N=None, which=None, **kw):
    """Full newton_solve! through the band solver vs the NumPy oracle (trajectories, duals, multipliers, history)."""
    W = ab.workloads
    model, N, dt, obj, con, opts, x0, xf = W.CONFIGS[name](batch=B, **({"N": N} if N else {}), **kw)
    gb = ab.GameBatch(model, N, dt, obj, con, B, lib_path=lib_path, solver=ab._capi.SOLVER_BAND)
    gb.set_instance_params(x0=x0, xf=xf)
    Z0, L0 = gb.random_initial(opts.amplitude_init, opts.seed)
    gb.set_history(opts.outer_iter * opts.inner_iter + 1)
    out = gb.newton_solve(opts)
    hist, count = gb.get_history()
    for b in (range(B) if which is None else which):
        op = oracle_problem(model, N, dt, obj, con, opts, x0[b], None if xf is None else xf[b])
        O.newton_solve(op, Z0=Z0[b], L0=L0[b])
        Zo = np.concatenate([op.pdtraj.X, op.pdtraj.U], axis=1)
        assert np.abs(out["Z"][b] - Zo).max() < TOL_SOLVE, np.abs(out["Z"][b] - Zo).max()
        assert np.abs(out["L"][b] - op.pdtraj.du).max() < TOL_SOLVE * max(1.0, np.abs(op.pdtraj.du).max())
        assert (out["status"][b] == 0) == op.converged and int(out["stats"][b, 6]) == op.n_newton
        cnt = int(count[b])
        assert cnt == len(op.stats)
        ho = np.array([[r.outer, r.res, r.dyn, r.con, r.sta, r.opt, r.delta] for r in op.stats])
        assert np.array_equal(hist[b, :cnt, 0], ho[:, 0])
        assert np.allclose(hist[b, :cnt, 1:7], ho[:, 1:], rtol=1e-6, atol=TOL_SOLVE)
        lam_o, mu_o = O.pack_multipliers(op)
        if lam_o.size:
            assert np.allclose(out["conlam"][b], lam_o, atol=TOL_SOLVE * max(1.0, np.abs(lam_o).max()))
            assert np.allclose(out["conmu"][b], mu_o, rtol=1e-12)
    gb.close()
    return out


def check_band_equals_structured(lib_path, name, B=4, N=None):
    """The planar configs through both solvers of the library: same statuses and Newton counts, solutions within 1e-6."""
    cfg = small_config(name, B, N)
    model, N, dt, obj, con, opts, x0, xf = cfg
    if x0.shape[0] < B:
        x0 = np.tile(x0[:1], (B, 1)); x0[:, :2 * model.p] += 0.01 * np.arange(B)[:, None]
    outs = []
    for solver in (ab._capi.SOLVER_AUTO, ab._capi.SOLVER_BAND):
        gb = ab.GameBatch(model, N, dt, obj, con, B, lib_path=lib_path, solver=solver)
        gb.set_instance_params(x0=x0, xf=xf)
        gb.random_initial(opts.amplitude_init, opts.seed)
        outs.append(gb.newton_solve(ab.Options(**{**opts.to_dict(), "dual_reset": True})))
        gb.close()
    a, b = outs
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["stats"][:, 6], b["stats"][:, 6])
    conv = a["status"] == 0
    assert np.abs(a["Z"][conv] - b["Z"][conv]).max() < TOL_SOLVE and np.allclose(a["stats"][:, :5], b["stats"][:, :5], rtol=1e-5, atol=TOL_SOLVE)
    return a


def check_band_window_equals_global(lib_path, monkeypatch, name, B=2, N=None, **kw):
    """band_solve_window (elimination window in shared memory, one barrier per column, slot map instead of row swaps) against
    band_solve_global (the first form, selected with AGB_BAND_WINDOW=0): the same operations in the same order — every result,
    the whole Statistics history and the Newton step of a random iterate must agree BIT FOR BIT."""
    W = ab.workloads
    model, N, dt, obj, con, opts, x0, xf = W.CONFIGS[name](batch=B, **({"N": N} if N else {}), **kw)
    res = {}
    for win in ("0", "1"):
        monkeypatch.setenv("AGB_BAND_WINDOW", win)
        gb = ab.GameBatch(model, N, dt, obj, con, B, lib_path=lib_path, solver=ab._capi.SOLVER_BAND)
        assert (gb.band_info()[0] > 0) == (win == "1")               # the window path really is the one that runs
        gb.set_instance_params(x0=x0, xf=xf)
        gb.random_initial(opts.amplitude_init, opts.seed)
        gb.set_history(opts.outer_iter * opts.inner_iter + 1)
        out = gb.newton_solve(opts)
        out["hist"], out["count"] = gb.get_history()
        rng = np.random.default_rng(11)
        Z = rng.normal(size=out["Z"].shape); L = rng.normal(size=out["L"].shape)
        gb.set_initial(Z, L)
        out["dtraj"] = gb.kkt_solve(1e-3, 1e-3)
        res[win] = out
        gb.close()
    monkeypatch.delenv("AGB_BAND_WINDOW")
    a, b = res["0"], res["1"]
    assert a["stats"][:, 6].min() >= 1                                   # the solves did factorise
    for k in ("status", "Z", "L", "conlam", "conmu", "stats", "count"):
        assert np.array_equal(a[k], b[k]), k
    for i in range(B):
        assert np.array_equal(a["hist"][i, :a["count"][i]], b["hist"][i, :b["count"][i]])
    assert np.array_equal(a["dtraj"], b["dtraj"])
    return a


def check_singular_fallback(lib_path, monkeypatch):
    """Plumbing of the fallback: with the test hook AGB_TEST_FORCE_SINGULAR=k the structured kernel reports every k-th
    instance as AGB_SINGULAR after its first factorisation; the band solver must re-solve exactly those, from the same
    initial iterate, and the batch must come out as if nothing had happened (stats[9] = 2 marks the re-solved ones)."""
    model, N, dt, obj, con, opts, x0, xf = small_config("B", 6, 10)
    res = {}
    for hook in ("0", "3"):
        monkeypatch.setenv("AGB_TEST_FORCE_SINGULAR", hook)
        gb = ab.GameBatch(model, N, dt, obj, con, 6, lib_path=lib_path)
        gb.set_instance_params(x0=x0)
        gb.random_initial(opts.amplitude_init, opts.seed)
        res[hook] = gb.newton_solve(opts)
        gb.close()
    monkeypatch.delenv("AGB_TEST_FORCE_SINGULAR")
    a, b = res["0"], res["3"]
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    assert np.array_equal(b["stats"][:, 9] >= 2, np.arange(6) % 3 == 0), b["stats"][:, 9]
    assert np.abs(a["Z"] - b["Z"]).max() < TOL_SOLVE and np.array_equal(a["stats"][:, 6], b["stats"][:, 6])


def check_mpc_fused_equals_stepwise(lib_path, name="D", B=8, N=None, resolves=6, disturbance_std=1e-3, seed=5, force_layout=None):
    """agb_mpc_run (every stream's whole receding-horizon loop inside ONE launch of the solve kernel) against the step-wise loop
    agb_newton_solve_batch + agb_mpc_advance: per-re-solve stats and status, executed states, and the handle's final state
    (Z, L, multipliers, penalties) must be IDENTICAL bit for bit — the streams are independent and the arithmetic is the same."""
    import os
    cfg = small_config(name, B, N)
    model, N_, dt, obj, con, opts, x0, xf = cfg
    rng = np.random.default_rng(seed)
    dist = disturbance_std * rng.standard_normal((resolves, B, model.n))
    if force_layout is not None:
        os.environ["AGB_FORCE_BIG_LAYOUT"] = str(force_layout)
    try:
        g1 = ab.GameBatch(model, N_, dt, obj, con, B, lib_path=lib_path)
        g2 = ab.GameBatch(model, N_, dt, obj, con, B, lib_path=lib_path)
    finally:
        os.environ.pop("AGB_FORCE_BIG_LAYOUT", None)
    stats1, status1, xs1 = ab.mpc.mpc_run(g1, opts, x0, resolves, xf=xf, disturbances=dist)
    g2.set_instance_params(x0=x0, xf=xf)
    g2.random_initial(opts.amplitude_init, opts.seed)
    first = ab.Options(**{**opts.to_dict(), "dual_reset": True})
    stats2, status2, xs2 = g2.mpc_run(first, resolves, 1, dist)
    assert np.array_equal(status1, status2), (status1, status2)
    assert np.array_equal(stats1, stats2), np.abs(stats1 - stats2).max()
    assert np.array_equal(xs1[1:], xs2)
    for a, b in zip(g1.get_state(), g2.get_state()):
        assert np.array_equal(a, b)
    v1, v2 = g1.newton_solve(ab.Options(**{**opts.to_dict(), "dual_reset": False})), g2.newton_solve(ab.Options(**{**opts.to_dict(), "dual_reset": False}))
    assert np.array_equal(v1["Z"], v2["Z"]) and np.array_equal(v1["stats"], v2["stats"])      # x0 and the warm start carried over too
    g1.close(); g2.close()
    return stats2, status2


def check_active_set_analysis(lib_path, name="B", N=8, seed=3, partial=True):
    """src/active_set/*.jl on the device (agb_active_set_* / agb_update_nullspace) against the oracle's restatement at a random
    iterate: bordered residual and Jacobian entry by entry, the active masks, and the null space of the masked Jacobian as a
    SUBSPACE (dimension, J·v = 0, and the two orthonormal bases span each other — a basis is unique only up to rotation).
    With `partial` the multipliers and positions are arranged so that only some collision pairs are active."""
    cfg = small_config(name, 1, N)
    model, N_, dt, obj, con, opts, x0, xf = cfg
    x0 = np.asarray(x0).reshape(-1, model.n)[0]
    xfj = None if xf is None else np.asarray(xf).reshape(-1, model.n)[0]
    op = oracle_problem(model, N_, dt, obj, con, opts, x0, xfj)
    rng = np.random.default_rng(seed)
    Z, L, lam, mu = random_state(op, x0, rng)
    if partial:                                   # players on top of each other at the early knots (active pairs), far apart later
        p = model.p
        for k in range(1, N_):
            for i in range(p):
                if k < N_ // 2:
                    Z[k, i], Z[k, p + i] = 0.01 * rng.normal(), 0.01 * rng.normal()
                else:
                    Z[k, i], Z[k, p + i] = 10.0 * (i + 1) + rng.normal(), -7.0 * (i + 1) + rng.normal()
        lam = np.zeros_like(lam)
    op.pdtraj.X[:], op.pdtraj.U[:], op.pdtraj.du[:] = Z[:, :model.n], Z[:, model.n:], L
    O.unpack_multipliers(op, lam, mu)
    gb = ab.GameBatch(model, N_, dt, obj, con, 1, lib_path=lib_path)
    gb.set_instance_params(x0=x0[None], xf=None if xfj is None else xfj[None])
    gb.set_initial(Z[None], L[None], lam[None], mu[None])
    asc = O.ActiveSetCore(op.probsize)
    assert gb.active_set_sizes() == (asc.Sv, asc.Sh)
    r_ref = O.as_residual(asc, op).copy()
    r_dev = gb.active_set_residual()[0]
    scale = max(1.0, np.abs(r_ref).max())
    assert np.abs(r_dev - r_ref).max() <= TOL_FUNC * scale, np.abs(r_dev - r_ref).max()
    J_ref = O.as_residual_jacobian(asc, op).copy()
    J_dev = gb.active_set_jacobian()[0]
    assert np.abs(J_dev - J_ref).max() <= TOL_FUNC * max(1.0, np.abs(J_ref).max()), np.abs(J_dev - J_ref).max()
    tol = opts.active_set_tolerance
    op.game_con.active_set_tolerance = tol
    N_ref = O.update_nullspace(asc, op)
    vm, hm = gb.active_set_masks(tol)
    assert np.array_equal(np.flatnonzero(vm[0]), asc.vmask) and np.array_equal(np.flatnonzero(hm[0]), asc.hmask)
    if partial:
        assert op.probsize.S < len(asc.vmask) < asc.Sv          # the case really is a partial active set
    basis = gb.update_nullspace(tol)[0]                          # [dim, Sh]
    assert basis.shape == (N_ref.shape[1], asc.Sh), (basis.shape, N_ref.shape)
    inactive = np.setdiff1d(np.arange(asc.Sh), asc.hmask)
    assert not basis[:, inactive].any()
    Bm = basis[:, asc.hmask]                                     # rows: orthonormal vectors in the masked column space
    assert np.abs(Bm @ Bm.T - np.eye(Bm.shape[0])).max() < 1e-10
    dj = J_ref[np.ix_(asc.vmask, asc.hmask)]
    assert np.abs(dj @ Bm.T).max() <= 1e-9 * max(1.0, np.abs(dj).max())
    sv = np.linalg.svd(dj, compute_uv=False)
    if sv[-1] > 1e-10 * sv[0]:
        # full numerical row rank — the null space is well defined: projecting the oracle's basis onto the device's loses nothing
        assert np.abs(N_ref - Bm.T @ (Bm @ N_ref)).max() < 1e-8
    else:
        # a singular value at round-off level (> atol = 1e-20, so nullspace() still counts it as rank): the reference's basis is then
        # ANY n - m of the n - m + 1 numerically null directions; the device's must lie inside that numerical null space
        Vt = np.linalg.svd(dj, full_matrices=True)[2]
        NN = Vt[int((sv > 1e-10 * sv[0]).sum()):].T
        gap = sv[sv > 1e-10 * sv[0]][-1]
        assert np.abs(Bm.T - NN @ (NN.T @ Bm.T)).max() <= 1e3 * np.abs(dj @ Bm.T).max() / gap + 1e-12
    gb.close()
    return basis.shape[0]
