# AlgamesB200.jl — drop-in batched `newton_solve!` for Algames.jl backed by libalgames_b200.so (NVIDIA B200, sm_100a).
#
#     using Algames, AlgamesB200
#     probs = [GameProblem(N, dt, x0s[b], model, opts, game_obj, game_con) for b in 1:B]
#     newton_solve!(probs)                      # all instances on the GPU; each `prob` is mutated like Algames.newton_solve!
#
# NOTE: no Julia toolchain exists in the build image (SURVEY.md F2), so this file has never been executed; it is kept
# deliberately thin — descriptor packing + `ccall`s that mirror, one for one, the ctypes binding in ../_capi.py that the
# test-suite exercises.  Struct layouts below must match include/algames_b200.h field for field.
module AlgamesB200

using Algames
using StaticArrays
using LinearAlgebra
import Algames: newton_solve!

const LIB = get(ENV, "ALGAMES_B200_LIB", joinpath(@__DIR__, "..", "libalgames_b200.so"))
const MAX_P, MAX_N, MAX_M, MAX_WALLS, MAX_CIRCLES, NSTATS = 4, 16, 8, 8, 8, 10

# ---- include/algames_b200.h : agb_problem_desc (isbits, C layout) --------------------------------------------------
struct AgbProblemDesc
    model::Cint; p::Cint; d::Cint; N::Cint
    dt::Cdouble; lf::Cdouble; lr::Cdouble
    Q::NTuple{MAX_N,Cdouble}; R::NTuple{MAX_M,Cdouble}; xf::NTuple{MAX_N,Cdouble}; uf::NTuple{MAX_M,Cdouble}
    has_collision_cost::Cint
    cc_radius::NTuple{MAX_P,Cdouble}; cc_mu::NTuple{MAX_P,Cdouble}
    col_radius::NTuple{MAX_P * MAX_P,Cdouble}                 # row-major [i][j]
    has_control_bound::Cint
    u_max::NTuple{MAX_M,Cdouble}; u_min::NTuple{MAX_M,Cdouble}
    has_state_bound::NTuple{MAX_P,Cint}
    x_max::NTuple{MAX_P * MAX_N,Cdouble}; x_min::NTuple{MAX_P * MAX_N,Cdouble}
    n_walls::NTuple{MAX_P,Cint}
    walls::NTuple{MAX_P * MAX_WALLS * 6,Cdouble}
    n_circles::NTuple{MAX_P,Cint}
    circles::NTuple{MAX_P * MAX_CIRCLES * 3,Cdouble}
    x_max_con::NTuple{MAX_P * MAX_N,Cint}; x_min_con::NTuple{MAX_P * MAX_N,Cint}   # owning conval of each finite bound
end

# ---- agb_options: live fields of Algames.Options (src/struct/options.jl) ------------------------------------------
struct AgbOptions
    reg_0::Cdouble; regularize::Cint; alpha_decrease::Cdouble; beta::Cdouble; ls_iter::Cint; delta_min::Cdouble
    rho_0::Cdouble; rho_increase::Cdouble; rho_max::Cdouble; lambda_max::Cdouble; alpha_dual::Cdouble
    alphax_dual::NTuple{MAX_P,Cdouble}; active_set_tolerance::Cdouble
    eps_dyn::Cdouble; eps_sta::Cdouble; eps_con::Cdouble; eps_opt::Cdouble
    outer_iter::Cint; inner_iter::Cint; dual_reset::Cint
end

AgbOptions(o::Options) = AgbOptions(o.reg_0, o.regularize, o.α_decrease, o.β, o.ls_iter, o.Δ_min, o.ρ_0[1], o.ρ_increase,
    o.ρ_max[1], o.λ_max, o.α_dual, ntuple(i -> Cdouble(o.αx_dual[i]), MAX_P), o.active_set_tolerance,
    o.ϵ_dyn, o.ϵ_sta, o.ϵ_con, o.ϵ_opt, o.outer_iter, o.inner_iter, o.dual_reset)

pad(v, n) = ntuple(i -> i <= length(v) ? Cdouble(v[i]) : 0.0, n)
padi(v, n) = ntuple(i -> i <= length(v) ? Cint(v[i]) : Cint(0), n)

model_id(::DoubleIntegratorGame) = 0
model_id(::UnicycleGame) = 1
model_id(::BicycleGame) = 2

# joint diagonal weights / targets out of the per-player LQR costs built by GameObjective (objective.jl:24-32):
# Q_i is non-zero on pz[i] only, so the joint vector is the sum over players; xf = -Q \ q on the weighted entries.
function joint_lqr(game_obj::GameObjective, model)
    n, m, p = model.n, model.m, model.p
    Q = zeros(n); R = zeros(m); xf = zeros(n); uf = zeros(m)
    for i in 1:p
        c = game_obj.obj[i][1].cost[1]
        q, r = diag(c.Q), diag(c.R)
        for a in model.pz[i]
            Q[a] = q[a]; xf[a] = q[a] != 0 ? -c.q[a] / q[a] : 0.0
        end
        for a in model.pu[i]
            R[a] = r[a]; uf[a] = r[a] != 0 ? -c.r[a] / r[a] : 0.0
        end
    end
    return Q, R, xf, uf
end

function make_desc(prob::GameProblem)
    model, ps, gc, go = prob.model, prob.probsize, prob.game_con, prob.game_obj
    n, m, p, N = ps.n, ps.m, ps.p, ps.N
    Q, R, xf, uf = joint_lqr(go, model)
    has_cc = 0; ccr = zeros(MAX_P); ccm = zeros(MAX_P)
    for i in 1:p, j in 2:length(go.obj[i])
        c = go.obj[i][j].cost[1]
        if c isa Algames.CollisionCost
            has_cc = 1; ccr[i] = c.r; ccm[i] = c.μ
        end
    end
    col = zeros(MAX_P, MAX_P); hsb = zeros(Cint, MAX_P)
    xmax = fill(Inf, MAX_N, MAX_P); xmin = fill(-Inf, MAX_N, MAX_P)
    xmaxc = zeros(Cint, MAX_N, MAX_P); xminc = zeros(Cint, MAX_N, MAX_P)
    nw = zeros(Cint, MAX_P); walls = zeros(6, MAX_WALLS, MAX_P)
    nc = zeros(Cint, MAX_P); circ = zeros(3, MAX_CIRCLES, MAX_P)
    for i in 1:p, con in gc.state_conlist[i].constraints
        if con isa TrajectoryOptimization.CollisionConstraint
            j = findfirst(q -> ps.px[q] == con.x2, 1:p)
            col[j, i] = con.radius                      # stored transposed: tuple below is row-major [i][j]
        elseif con isa Algames.StateBoundConstraint
            # several StateBound convals per player (add_velocity_bound!, velocity_constraint.jl:13-28) are merged
            # component-wise; xmaxc / xminc keep the 0-based conval that owns each finite entry
            g = hsb[i]; hsb[i] += 1
            for a in 1:n
                if isfinite(con.x_max[a])
                    isfinite(xmax[a, i]) && error("AlgamesB200: two state bounds of player $i on component $a (upper)")
                    xmax[a, i] = con.x_max[a]; xmaxc[a, i] = g
                end
                if isfinite(con.x_min[a])
                    isfinite(xmin[a, i]) && error("AlgamesB200: two state bounds of player $i on component $a (lower)")
                    xmin[a, i] = con.x_min[a]; xminc[a, i] = g
                end
            end
        elseif con isa Algames.WallConstraint
            for q in 1:length(con.x1)
                nw[i] += 1
                walls[:, nw[i], i] .= (con.x1[q], con.y1[q], con.x2[q], con.y2[q], con.xv[q], con.yv[q])
            end
        elseif con isa TrajectoryOptimization.CircleConstraint
            for q in 1:length(con.x)
                nc[i] += 1
                circ[:, nc[i], i] .= (con.x[q], con.y[q], con.radius[q])
            end
        else
            error("AlgamesB200: unsupported state constraint $(typeof(con))")
        end
    end
    hcb = 0; umax = fill(Inf, MAX_M); umin = fill(-Inf, MAX_M)
    for con in gc.control_conlist.constraints
        con isa Algames.ControlBoundConstraint || error("AlgamesB200: unsupported control constraint $(typeof(con))")
        hcb = 1; umax[1:m] .= con.u_max; umin[1:m] .= con.u_min
    end
    lf = model isa BicycleGame ? model.lf : 0.05
    lr = model isa BicycleGame ? model.lr : 0.05
    return AgbProblemDesc(model_id(model), p, 2, N, prob.pdtraj.pr[1].dt, lf, lr,
        pad(Q, MAX_N), pad(R, MAX_M), pad(xf, MAX_N), pad(uf, MAX_M), has_cc, pad(ccr, MAX_P), pad(ccm, MAX_P),
        pad(vec(col), MAX_P * MAX_P), hcb, pad(umax, MAX_M), pad(umin, MAX_M), padi(hsb, MAX_P),
        pad(vec(xmax), MAX_P * MAX_N), pad(vec(xmin), MAX_P * MAX_N), padi(nw, MAX_P),
        pad(vec(walls), MAX_P * MAX_WALLS * 6), padi(nc, MAX_P), pad(vec(circ), MAX_P * MAX_CIRCLES * 3),
        padi(vec(xmaxc), MAX_P * MAX_N), padi(vec(xminc), MAX_P * MAX_N))
end

# ---- agb_sizes ---------------------------------------------------------------------------------------------------
struct AgbSizes
    n::Cint; m::Cint; p::Cint; N::Cint; S::Cint; nrow::Cint; nrow_state::Cint; nrow_control::Cint
end

"""
    canonical_convals(prob) -> Vector of (conval, is_control)

The convals of `prob.game_con` in the row order of the ABI's `conlam` / `conmu` arrays (include/algames_b200.h): per
player the collision convals by ascending opponent, then the state bounds, walls and circles each in the order they were
added; then the control bound.  Rows inside a conval keep the conval's own order (finite upper bounds then finite lower
bounds for the bound constraints, control_bound_constraint.jl:33-35; one row per wall / circle).
"""
function canonical_convals(prob::GameProblem)
    ps, gc = prob.probsize, prob.game_con
    out = Tuple{Any,Bool}[]
    for i in 1:ps.p
        cvs = gc.state_conval[i]
        col = [cv for cv in cvs if cv.con isa TrajectoryOptimization.CollisionConstraint]
        sort!(col, by = cv -> findfirst(q -> ps.px[q] == cv.con.x2, 1:ps.p))
        append!(out, [(cv, false) for cv in col])
        for T in (Algames.StateBoundConstraint, Algames.WallConstraint, TrajectoryOptimization.CircleConstraint)
            append!(out, [(cv, false) for cv in cvs if cv.con isa T])
        end
    end
    append!(out, [(cv, true) for cv in gc.control_conval])
    return out
end

"""
    scatter_multipliers!(prob, lam, mu)

Writes the solver's AL multipliers and penalties (`lam`, `mu`: nrow × (N-1), stage s = state rows of knot s+1 followed by
the control-bound rows of knot s) back into `conval.λ[l]`, `conval.μ[l]` as `newton_solve!` leaves them in the reference
(dual_update! / penalty_update!, constraints_methods.jl:329-365, :421-430).
"""
function scatter_multipliers!(prob::GameProblem, lam::AbstractMatrix, mu::AbstractMatrix)
    row = 0
    for (cv, is_control) in canonical_convals(prob)
        len = length(cv.λ[1])
        for (l, k) in enumerate(cv.inds)                 # state convals: knots 2:N; control conval: knots 1:N-1
            s = is_control ? k : k - 1
            cv.λ[l] = typeof(cv.λ[l])(lam[row+1:row+len, s])
            cv.μ[l] = typeof(cv.μ[l])(mu[row+1:row+len, s])
        end
        row += len
    end
    return nothing
end

"Inverse of `scatter_multipliers!`: the convals' λ, μ in the ABI's row order (warm starts with `dual_reset = false`)."
function gather_multipliers!(lam::AbstractMatrix, mu::AbstractMatrix, prob::GameProblem)
    row = 0
    for (cv, is_control) in canonical_convals(prob)
        len = length(cv.λ[1])
        for (l, k) in enumerate(cv.inds)
            s = is_control ? k : k - 1
            lam[row+1:row+len, s] .= cv.λ[l]; mu[row+1:row+len, s] .= cv.μ[l]
        end
        row += len
    end
    return nothing
end

check(rc, h) = rc == 0 || error("libalgames_b200 ($rc): " *
    unsafe_string(ccall((:agb_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))

"""
    newton_solve!(probs::AbstractVector{<:GameProblem}; device=0)

Batched replacement of `Algames.newton_solve!` (src/problem/solver_methods.jl:5-65).  All problems must share the
model, sizes and constraint schema of `probs[1]`; x0 and the LQR weights/targets may differ per instance.
C arrays are row-major: a Julia `Array{Float64}` with reversed dims has exactly the layout the ABI expects.
"""
function newton_solve!(probs::AbstractVector{<:GameProblem}; device::Integer=0)
    B = length(probs); prob = probs[1]; opts = prob.opts
    ps = prob.probsize; n, m, p, N = ps.n, ps.m, ps.p, ps.N
    desc = Ref(make_desc(prob)); h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:agb_create, LIB), Cint, (Ref{AgbProblemDesc}, Cint, Cint, Ref{Ptr{Cvoid}}), desc, B, device, h)
    rc == 0 || error("agb_create ($rc): " * unsafe_string(ccall((:agb_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    try
        x0 = Array{Float64}(undef, n, B); xf = similar(x0); Q = similar(x0)
        R = Array{Float64}(undef, m, B); uf = similar(R)
        Z0 = Array{Float64}(undef, n + m, N, B); L0 = Array{Float64}(undef, n, N - 1, p, B)
        for (b, q) in enumerate(probs)
            # host keeps Julia's RNG semantics for the initial iterate (solver_methods.jl:12-15)
            Algames.Random.seed!(q.opts.seed)
            init_traj!(q.pdtraj; x0=q.x0, f=q.opts.f_init, amplitude=q.opts.amplitude_init, s=q.opts.shift)
            x0[:, b] .= q.x0
            Q[:, b], R[:, b], xf[:, b], uf[:, b] = joint_lqr(q.game_obj, q.model)
            for k in 1:N; Z0[:, k, b] .= q.pdtraj.pr[k].z; end
            for i in 1:p, k in 1:N-1; L0[:, k, i, b] .= q.pdtraj.du[i][k]; end
        end
        check(ccall((:agb_set_instance_params, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h[], x0, xf, Q, R, uf), h[])
        sz = Ref(AgbSizes(0, 0, 0, 0, 0, 0, 0, 0))
        check(ccall((:agb_get_sizes, LIB), Cint, (Ptr{Cvoid}, Ref{AgbSizes}), h[], sz), h[])
        nrow = Int(sz[].nrow)
        conλ = Array{Float64}(undef, nrow, N - 1, B); conμ = similar(conλ)     # [B][N-1][nrow] in C order
        warm = nrow > 0 && !opts.dual_reset            # dual_reset = false keeps the convals' λ, μ (solver_methods.jl:25)
        if warm
            for (b, q) in enumerate(probs); gather_multipliers!(view(conλ, :, :, b), view(conμ, :, :, b), q); end
        end
        GC.@preserve conλ conμ check(ccall((:agb_set_initial, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h[], Z0, L0, warm ? pointer(conλ) : C_NULL, warm ? pointer(conμ) : C_NULL), h[])
        Z = similar(Z0); L = similar(L0); stats = Array{Float64}(undef, NSTATS, B); status = Vector{Cint}(undef, B)
        # one log entry per record!(stats, …) of the reference loop (statistics.jl:44-57)
        maxrec = opts.outer_iter * opts.inner_iter + 1
        check(ccall((:agb_set_history, LIB), Cint, (Ptr{Cvoid}, Cint), h[], maxrec), h[])
        hist = Array{Float64}(undef, 8, maxrec, B); nrec = Vector{Cint}(undef, B)
        o = Ref(AgbOptions(opts))
        GC.@preserve Z L conλ conμ stats status hist nrec begin
            check(ccall((:agb_newton_solve_batch, LIB), Cint,
                (Ptr{Cvoid}, Ref{AgbOptions}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cint}),
                h[], o, Z, L, nrow > 0 ? pointer(conλ) : C_NULL, nrow > 0 ? pointer(conμ) : C_NULL, stats, status), h[])
            check(ccall((:agb_get_history, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Cint}), h[], hist, nrec), h[])
        end
        for (b, q) in enumerate(probs)
            for k in 1:N
                Algames.RobotDynamics.set_state!(q.pdtraj.pr[k], SVector{n}(Z[1:n, k, b]))
                k < N && Algames.RobotDynamics.set_control!(q.pdtraj.pr[k], SVector{m}(Z[n+1:n+m, k, b]))
            end
            for i in 1:p, k in 1:N-1; q.pdtraj.du[i][k] = SVector{n}(L[:, k, i, b]); end
            scatter_multipliers!(q, view(conλ, :, :, b), view(conμ, :, :, b))
            evaluate!(q.game_con, q.pdtraj.pr)             # conval.vals at the returned iterate
            residual!(q)                                   # prob.core.res, as the reference leaves it
            reset!(q.stats)                                # the device log replays every record! of the solve
            for r in 1:min(nrec[b], maxrec)
                k, res, dyn, con, sta, opt, Δ = hist[1:7, r, b]
                dv = DynamicsViolation(N); dv.max = dyn        # violations.jl:11-16 (per-knot vectors stay zero)
                cv = ControlViolation(N); cv.max = con
                sv = StateViolation(N); sv.max = sta
                ov = OptimalityViolation(N); ov.max = opt
                record!(q.stats, 0.0, res, Δ, dv, cv, sv, ov, Int(k))      # statistics.jl:30-42
            end
        end
        return status
    finally
        ccall((:agb_destroy, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
end

end # module
