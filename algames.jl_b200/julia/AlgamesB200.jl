# AlgamesB200.jl — drop-in batched `newton_solve!` for Algames.jl backed by libalgames_b200.so (NVIDIA B200, sm_100a).
#
#     using Algames, AlgamesB200
#     probs = [GameProblem(N, dt, x0s[b], model, opts, game_obj, game_con) for b in 1:B]
#     newton_solve!(probs)                      # all instances on the GPU; each `prob` is mutated like Algames.newton_solve!
#
# NOTE: no Julia toolchain exists in the build image (SURVEY.md F2), so this file has never been executed; it is kept
# deliberately thin (its struct layouts are checked against the library by agb_abi_check in __init__) — descriptor packing + `ccall`s that mirror, one for one, the ctypes binding in ../_capi.py that the
# test-suite exercises.  Struct layouts below must match include/algames_b200.h field for field.
module AlgamesB200

using Algames
using StaticArrays
using LinearAlgebra
import Algames: newton_solve!

const LIB = get(ENV, "ALGAMES_B200_LIB", joinpath(@__DIR__, "..", "libalgames_b200.so"))
const MAX_P, MAX_N, MAX_M, MAX_WALLS, MAX_CIRCLES, NSTATS, NHIST, IPC_BYTES = 4, 16, 8, 8, 8, 10, 10, 64

# per-instance status codes (include/algames_b200.h)
const CONVERGED, MAX_OUTER, LINE_SEARCH_FAILED, STALLED, SINGULAR, NONFINITE = 0, 1, 2, 3, 4, 5

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

# agb_device_view (only its size is needed here: 10 pointers + one UInt64)
struct AgbDeviceView
    ptrs::NTuple{10,Ptr{Cvoid}}; results_bytes::Culonglong
end
struct AgbIBROptions
    ibr_iter::Cint; ordering::NTuple{MAX_P,Cint}; delta_min::Cdouble
end

"The AGB_ABI_WORDS values agb_abi_check compares with the header the library was built from (include/algames_b200.h)."
abi_layout() = Cint[sizeof(AgbProblemDesc), fieldoffset(AgbProblemDesc, 5), fieldoffset(AgbProblemDesc, 8),
    fieldoffset(AgbProblemDesc, 15), fieldoffset(AgbProblemDesc, 19), fieldoffset(AgbProblemDesc, 23), fieldoffset(AgbProblemDesc, 25),
    fieldoffset(AgbProblemDesc, 26), sizeof(AgbOptions), fieldoffset(AgbOptions, 12), fieldoffset(AgbOptions, 14), fieldoffset(AgbOptions, 20),
    sizeof(AgbIBROptions), fieldoffset(AgbIBROptions, 3), sizeof(AgbSizes), sizeof(AgbDeviceView),
    MAX_P, MAX_N, MAX_M, MAX_WALLS, MAX_CIRCLES, NSTATS, NHIST, IPC_BYTES]

function __init__()
    # the structs above mirror the C header by hand: refuse to run against a library built from another layout
    w = abi_layout()
    rc = ccall((:agb_abi_check, LIB), Cint, (Ptr{Cint}, Cint), w, length(w))
    rc == 0 || error("AlgamesB200: " * unsafe_string(ccall((:agb_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
end

"""
    Batch(probs; device=0)

Persistent device state for a vector of GameProblems sharing one schema: ONE agb_handle (device buffers, streams, the loaded
kernels) and page-locked host staging arrays, created once and reused by every `newton_solve!(batch)` / `mpc_step!(batch)`.
`close(batch)` (or the finalizer) releases it.  `batch.status[b]` holds the per-instance status code of the last solve.
"""
mutable struct Batch
    h::Ptr{Cvoid}
    probs::Vector
    n::Int; m::Int; p::Int; N::Int; nrow::Int
    x0::Matrix{Float64}; Z0::Array{Float64,3}; L0::Array{Float64,4}
    Z::Array{Float64,3}; L::Array{Float64,4}; conλ::Array{Float64,3}; conμ::Array{Float64,3}
    stats::Matrix{Float64}; status::Vector{Cint}
end

function Batch(probs::AbstractVector{<:GameProblem}; device::Integer=0)
    prob = probs[1]; ps = prob.probsize; n, m, p, N = ps.n, ps.m, ps.p, ps.N; B = length(probs)
    desc = Ref(make_desc(prob)); h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:agb_create, LIB), Cint, (Ref{AgbProblemDesc}, Cint, Cint, Ref{Ptr{Cvoid}}), desc, B, device, h)
    rc == 0 || error("agb_create ($rc): " * unsafe_string(ccall((:agb_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    sz = Ref(AgbSizes(0, 0, 0, 0, 0, 0, 0, 0))
    check(ccall((:agb_get_sizes, LIB), Cint, (Ptr{Cvoid}, Ref{AgbSizes}), h[], sz), h[])
    nrow = Int(sz[].nrow)
    b = Batch(h[], collect(probs), n, m, p, N, nrow, zeros(n, B), zeros(n + m, N, B), zeros(n, N - 1, p, B),
              zeros(n + m, N, B), zeros(n, N - 1, p, B), zeros(nrow, N - 1, B), zeros(nrow, N - 1, B), zeros(NSTATS, B), zeros(Cint, B))
    # page-lock the staging arrays so that agb_solve_from_host's chunked copy/solve pipeline overlaps (CUDA.jl: Mem.register;
    # without CUDA.jl the copies still work, just without overlap)
    finalizer(close, b)
    return b
end

function Base.close(b::Batch)
    b.h == C_NULL && return
    ccall((:agb_destroy, LIB), Cvoid, (Ptr{Cvoid},), b.h); b.h = C_NULL
    return nothing
end

"""
    newton_solve!(batch::Batch)  ->  nothing

One call of the hot path (`agb_solve_from_host`: x0 + initial iterate in, trajectories / duals / multipliers / stats out,
copies pipelined with the solve in chunks) for every problem of the batch; mutates each `prob` like Algames.newton_solve!.
Returns `nothing` like the reference; per-instance outcomes are in `batch.status`.
"""
function newton_solve!(b::Batch)
    probs = b.probs; opts = probs[1].opts; n, m, p, N = b.n, b.m, b.p, b.N
    for (k, q) in enumerate(probs)
        Algames.Random.seed!(q.opts.seed)                  # Julia's RNG semantics for the initial iterate (solver_methods.jl:12-15)
        init_traj!(q.pdtraj; x0=q.x0, f=q.opts.f_init, amplitude=q.opts.amplitude_init, s=q.opts.shift)
        b.x0[:, k] .= q.x0
        for t in 1:N; b.Z0[:, t, k] .= q.pdtraj.pr[t].z; end
        for i in 1:p, t in 1:N-1; b.L0[:, t, i, k] .= q.pdtraj.du[i][t]; end
    end
    if b.nrow > 0 && !opts.dual_reset                       # warm start: the convals' λ, μ go in through agb_set_initial
        for (k, q) in enumerate(probs); gather_multipliers!(view(b.conλ, :, :, k), view(b.conμ, :, :, k), q); end
        check(ccall((:agb_set_initial, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            b.h, C_NULL, C_NULL, b.conλ, b.conμ), b.h)
    end
    o = Ref(AgbOptions(opts))
    check(ccall((:agb_solve_from_host, LIB), Cint,
        (Ptr{Cvoid}, Ref{AgbOptions}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cint}),
        b.h, o, b.x0, b.Z0, b.L0, b.Z, b.L, b.nrow > 0 ? pointer(b.conλ) : C_NULL, b.nrow > 0 ? pointer(b.conμ) : C_NULL, b.stats, b.status), b.h)
    scatter_results!(b)
    return nothing
end

"Per-knot violation vectors of the resident iterate (agb_violations; struct/violations.jl `.vio`)."
function violations(b::Batch)
    B = length(b.probs); N = b.N
    dyn = zeros(N - 1, B); con = zeros(N - 1, B); sta = zeros(N, B); opt = zeros(N, B)
    check(ccall((:agb_violations, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), b.h, dyn, con, sta, opt), b.h)
    return dyn, con, sta, opt
end

"""
    update_nullspace!(ascore::ActiveSetCore, batch::Batch, k::Integer=1; atol=1e-20)

`Algames.update_nullspace!` (src/active_set/active_set_methods.jl:173-184) for instance `k` of a resident batch, on the device:
active masks (`agb_active_set_masks`), null space of the masked bordered Jacobian (`agb_update_nullspace`), then the reference's own
`add_matrix!` (active_set_core.jl:29-52) fills `ascore.null`.  The basis is orthonormal; it is unique only up to a rotation.
"""
function update_nullspace!(ascore, b::Batch, k::Integer=1; atol::Float64=1e-20)
    B = length(b.probs)
    sv = Ref{Cint}(0); sh = Ref{Cint}(0)
    check(ccall((:agb_active_set_sizes, LIB), Cint, (Ptr{Cvoid}, Ref{Cint}, Ref{Cint}), b.h, sv, sh), b.h)
    Sv, Sh = Int(sv[]), Int(sh[])
    tol = b.probs[1].opts.active_set_tolerance
    vmask = zeros(UInt8, Sv, B); hmask = zeros(UInt8, Sh, B)
    check(ccall((:agb_active_set_masks, LIB), Cint, (Ptr{Cvoid}, Cdouble, Ptr{UInt8}, Ptr{UInt8}), b.h, tol, vmask, hmask), b.h)
    maxdim = Sh - Sv + (b.N - 1) * b.p
    null = zeros(Sh, maxdim, B); dim = zeros(Cint, B)
    check(ccall((:agb_update_nullspace, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cint, Ptr{Float64}, Ptr{Cint}),
        b.h, tol, atol, maxdim, null, dim), b.h)
    ascore.vmask = findall(!iszero, view(vmask, :, k)); ascore.hmask = findall(!iszero, view(hmask, :, k))
    Algames.reset!(ascore.null)
    Algames.add_matrix!(ascore.null, null[ascore.hmask, 1:dim[k], k], ascore.hmask)
    return nothing
end

function scatter_results!(b::Batch)
    n, m, p, N = b.n, b.m, b.p, b.N
    dyn, con, sta, opt = violations(b)
    for (k, q) in enumerate(b.probs)
        for t in 1:N
            Algames.RobotDynamics.set_state!(q.pdtraj.pr[t], SVector{n}(b.Z[1:n, t, k]))
            t < N && Algames.RobotDynamics.set_control!(q.pdtraj.pr[t], SVector{m}(b.Z[n+1:n+m, t, k]))
        end
        for i in 1:p, t in 1:N-1; q.pdtraj.du[i][t] = SVector{n}(b.L[:, t, i, k]); end
        b.nrow > 0 && scatter_multipliers!(q, view(b.conλ, :, :, k), view(b.conμ, :, :, k))
        evaluate!(q.game_con, q.pdtraj.pr); residual!(q)
        reset!(q.stats)
        dv = DynamicsViolation(N); dv.vio .= dyn[:, k]; dv.max = b.stats[2, k]
        cv = ControlViolation(N); cv.vio .= con[:, k]; cv.max = b.stats[3, k]
        sv = StateViolation(N); sv.vio .= sta[:, k]; sv.max = b.stats[4, k]
        ov = OptimalityViolation(N); ov.vio .= opt[:, k]; ov.max = b.stats[5, k]
        record!(q.stats, 0.0, b.stats[1, k], b.stats[6, k], dv, cv, sv, ov, Int(b.stats[8, k]))   # the final record (:63)
    end
end

"""
    mpc_step!(batch::Batch; disturbance=nothing)

One receding-horizon step on the device (`agb_mpc_advance`): x0 ← x₂ of the resident solution (+ disturbance n×B), the
solution is shifted by one knot as the warm start (Options.shift = 1, primal_dual_traj.jl:34-41), then re-solved with the
multipliers carried (dual_reset = false).  Only x₂…, stats and status come back to the host.
"""
function mpc_step!(b::Batch; disturbance::Union{Nothing,Matrix{Float64}}=nothing)
    check(ccall((:agb_mpc_advance, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        b.h, 1, disturbance === nothing ? C_NULL : pointer(disturbance), C_NULL, C_NULL), b.h)
    o = AgbOptions(b.probs[1].opts)
    warm = Ref(AgbOptions(o.reg_0, o.regularize, o.alpha_decrease, o.beta, o.ls_iter, o.delta_min, o.rho_0, o.rho_increase, o.rho_max,
        o.lambda_max, o.alpha_dual, o.alphax_dual, o.active_set_tolerance, o.eps_dyn, o.eps_sta, o.eps_con, o.eps_opt, o.outer_iter, o.inner_iter, Cint(0)))
    check(ccall((:agb_newton_solve_batch, LIB), Cint,
        (Ptr{Cvoid}, Ref{AgbOptions}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cint}),
        b.h, warm, b.Z, C_NULL, C_NULL, C_NULL, b.stats, b.status), b.h)
    return nothing
end

"""
    mpc_run!(batch::Batch, resolves; shift=1, disturbances=nothing) -> (stats, status, xs)

The whole receding-horizon loop of every stream in ONE kernel launch (`agb_mpc_run`): stream b's CTA runs `resolves` ×
(newton_solve!, x0 ← x_{1+shift} + disturbances[:, b, t], shift the iterate, carry multipliers and penalties) back to back on
the device; bit for bit the results of `resolves` × (`newton_solve!`, `mpc_step!`).  `disturbances` is n × B × resolves.
Returns per-re-solve `stats` (NSTATS × B × resolves), `status` (B × resolves) and the executed states `xs` (n × B × resolves).
`opts.dual_reset` applies to the first re-solve only.
"""
function mpc_run!(b::Batch, resolves::Integer; shift::Integer=1, disturbances::Union{Nothing,Array{Float64,3}}=nothing)
    B = length(b.probs)
    stats = zeros(NSTATS, B, resolves); status = zeros(Cint, B, resolves); xs = zeros(b.n, B, resolves)
    o = Ref(AgbOptions(b.probs[1].opts))
    check(ccall((:agb_mpc_run, LIB), Cint,
        (Ptr{Cvoid}, Ref{AgbOptions}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Cint}, Ptr{Float64}),
        b.h, o, resolves, shift, disturbances === nothing ? C_NULL : pointer(disturbances), stats, status, xs), b.h)
    return stats, status, xs
end

"""
    ShardedBatch(probs; devices=nothing)

Single-process multi-GPU form (agb_create_sharded): the problems are split contiguously over the visible GPUs, every shard
is solved by its own handle with no communication, and `allgather!(sb)` pushes every shard's result slab into every GPU's
gather buffer over NVLink peer memory (the path's single collective, SURVEY §8e).
"""
mutable struct ShardedBatch
    handles::Vector{Ptr{Cvoid}}
    bounds::Vector{UnitRange{Int}}          # problems of shard r (contiguous split, ranks < B % ndev get one more)
    probs::Vector
end

function ShardedBatch(probs::AbstractVector{<:GameProblem}; devices::Union{Nothing,Vector{Cint}}=nothing)
    B = length(probs); desc = Ref(make_desc(probs[1]))
    ndev = Ref{Cint}(0); hs = Vector{Ptr{Cvoid}}(undef, 64)
    rc = ccall((:agb_create_sharded, LIB), Cint, (Ref{AgbProblemDesc}, Cint, Cint, Ptr{Cint}, Ptr{Ptr{Cvoid}}, Ref{Cint}),
               desc, B, devices === nothing ? 0 : length(devices), devices === nothing ? C_NULL : pointer(devices), hs, ndev)
    rc == 0 || error("agb_create_sharded ($rc): " * unsafe_string(ccall((:agb_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    nd = Int(ndev[]); resize!(hs, nd)
    bounds = UnitRange{Int}[]; lo = 1
    for r in 0:nd-1
        cnt = div(B, nd) + (r < B % nd ? 1 : 0); push!(bounds, lo:lo+cnt-1); lo += cnt
    end
    return ShardedBatch(hs, bounds, collect(probs))
end

"""
    newton_solve!(sb::ShardedBatch)

Every shard solved on its own GPU (asynchronously, no communication), then ONE all-gather: afterwards every GPU's gather
buffer holds every shard's trajectories, duals, stats and status (`agb_gathered_view` / `agb_unpack_gathered`).
"""
function newton_solve!(sb::ShardedBatch)
    o = Ref(AgbOptions(sb.probs[1].opts))
    for (h, rng) in zip(sb.handles, sb.bounds)          # (initial iterates are pushed per shard with agb_set_instance_params /
        check(ccall((:agb_newton_solve_async, LIB), Cint, (Ptr{Cvoid}, Ref{AgbOptions}, Ptr{Cvoid}), h, o, ccall((:agb_get_stream, LIB), Ptr{Cvoid}, (Ptr{Cvoid},), h)), h)
    end                                                 #  agb_set_initial exactly as Batch does for one device)
    allgather!(sb.handles)
    return nothing
end

function allgather!(hs::Vector{Ptr{Cvoid}})
    for h in hs; check(ccall((:agb_allgather, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h, C_NULL), h); end
    for h in hs; check(ccall((:agb_allgather_wait, LIB), Cint, (Ptr{Cvoid},), h), h); end
    return nothing
end

"""
    newton_solve!(probs::AbstractVector{<:GameProblem}; device=0)  ->  nothing

Batched replacement of `Algames.newton_solve!` (src/problem/solver_methods.jl:5-65) for a one-off batch: builds a `Batch`,
solves, releases it.  All problems must share the model, sizes, constraint schema and Options of `probs[1]`; x0 and the LQR
weights/targets may differ per instance.  In a loop (MPC, Monte-Carlo sweeps) create the `Batch` once instead.
C arrays are row-major: a Julia `Array{Float64}` with reversed dims has exactly the layout the ABI expects.
"""
function newton_solve!(probs::AbstractVector{<:GameProblem}; device::Integer=0)
    b = Batch(probs; device=device)
    try
        xf = zeros(b.n, length(probs)); Q = similar(xf); R = zeros(b.m, length(probs)); uf = similar(R)
        for (k, q) in enumerate(probs)
            Q[:, k], R[:, k], xf[:, k], uf[:, k] = joint_lqr(q.game_obj, q.model)
        end
        check(ccall((:agb_set_instance_params, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            b.h, C_NULL, xf, Q, R, uf), b.h)
        newton_solve!(b)
    finally
        close(b)
    end
    return nothing
end

end # module
