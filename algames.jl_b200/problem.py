"""Host-side mirror of the reference's exported API for the hot path (src/Algames.jl:124-143).

Same names, argument meaning and error behaviour as the Julia package — `DoubleIntegratorGame/UnicycleGame/BicycleGame`,
`ProblemSize`, `GameObjective` + `add_collision_cost`, `GameConstraintValues` + `add_*`, `Options`, `GameProblem`,
`newton_solve` (Julia: `newton_solve!`) — with the work done by libalgames_b200.so on the GPU.  Julia's `!` suffix is
dropped and Unicode option names are spelled out (ρ_0 → rho_0, ϵ_dyn → eps_dyn, Δ_min → delta_min, ...).

Host arithmetic here is limited to packing descriptors and drawing the tiny random initial iterate
(primal_dual_traj.jl:29-44); everything numerical runs in the CUDA library.  Nothing in this package imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field, fields
from typing import List, Optional, Sequence

import numpy as np

from . import _capi

__all__ = [
    "DoubleIntegratorGame", "UnicycleGame", "BicycleGame", "QuadrotorGame", "Wall3D", "CylinderWall", "add_spherical_collision_avoidance", "ProblemSize", "Options", "GameObjective",
    "add_collision_cost", "GameConstraintValues", "add_collision_avoidance", "add_control_bound", "add_state_bound", "add_velocity_bound", "velocity_index",
    "add_circle_constraint", "add_wall_constraint", "Wall", "GameProblem", "GameBatch", "newton_solve",
    "residual", "residual_jacobian", "kkt_solve", "line_search", "update_traj", "rollout", "dual_update",
    "penalty_update", "reset", "evaluate", "active_set", "Statistics", "spec_of", "IBROptions", "ibr_newton_solve",
    "ActiveSetCore", "NullSpace", "active_set_residual", "active_set_residual_jacobian", "active_masks", "update_nullspace",
]


# ------------------------------------------------------------------------------------------------------------
# Models (src/dynamics/*.jl): component-major joint layout, pu[i] = [i, i+p], px[i] = [i, i+p], pz[i] = [i, i+p, ...]
# ------------------------------------------------------------------------------------------------------------
class _GameModel:
    name = "abstract"

    def __init__(self, p: int, ni: int = 4, mi: int = 2):
        if not 1 <= p <= _capi.MAX_P:
            raise ValueError(f"p must be in 1..{_capi.MAX_P}")
        self.p, self.n, self.m = p, ni * p, mi * p
        self.ni, self.mi = [ni] * p, [mi] * p
        self.pu = [[i + j * p for j in range(mi)] for i in range(p)]      # 0-based
        self.px = [[i + j * p for j in range(2)] for i in range(p)]
        self.pz = [[i + j * p for j in range(ni)] for i in range(p)]


class DoubleIntegratorGame(_GameModel):
    """dynamics/double_integrator.jl:2-33: d = 2 (structured kernels) or d = 3 (band solver)."""
    name = "double_integrator"

    def __init__(self, p: int = 2, d: int = 2):
        if d not in (2, 3):
            raise NotImplementedError("DoubleIntegratorGame: d must be 2 or 3")
        super().__init__(p, 2 * d, d)
        self.d = d


class UnicycleGame(_GameModel):
    """dynamics/unicycle.jl:2-34."""
    name = "unicycle"

    def __init__(self, p: int = 2):
        super().__init__(p)


class BicycleGame(_GameModel):
    """dynamics/bicycle.jl:2-43."""
    name = "bicycle"

    def __init__(self, p: int = 2, lf: float = 0.05, lr: float = 0.05):
        super().__init__(p)
        self.lf, self.lr = lf, lr


class QuadrotorGame(_GameModel):
    """dynamics/quadrotor.jl:2-208: 12 states [r, q (MRP attitude), v, ω] and 4 rotor commands per player, component-major
    joint layout.  Solved by the library's band solver (the structured kernels are specialised for planar players)."""
    name = "quadrotor"

    def __init__(self, p: int = 2, mass: float = 0.5):
        super().__init__(p, 12, 4)
        self.mass = mass


class ProblemSize:
    """struct/problem_size.jl:5-35."""

    def __init__(self, N: int, model: _GameModel):
        self.N, self.n, self.m, self.p = N, model.n, model.m, model.p
        self.ni, self.mi, self.pu, self.px, self.pz = model.ni, model.mi, model.pu, model.px, model.pz
        self.S = self.n * self.p * (N - 1) + self.m * (N - 1) + self.n * (N - 1)
        self.b = self.n * self.p + self.m + self.n


# ------------------------------------------------------------------------------------------------------------
# Options (struct/options.jl:5-116) — live fields only; dead ones (θ, α_0, γ, ...) are accepted and ignored
# ------------------------------------------------------------------------------------------------------------
@dataclass
class Options:
    amplitude_init: float = 1e-8
    shift: int = 2 ** 10
    regularize: bool = True
    reg_0: float = 1e-3
    alpha_decrease: float = 0.5
    beta: float = 0.01
    ls_iter: int = 25
    delta_min: float = 1e-9
    rho_0: float = 1.0
    rho_increase: float = 10.0
    rho_max: float = 1e7
    lambda_max: float = 1e7
    alpha_dual: float = 1.0
    alphax_dual: List[float] = field(default_factory=lambda: [1.0] * 10)
    active_set_tolerance: float = 1e-4
    eps_dyn: float = 1e-3
    eps_sta: float = 1e-3
    eps_con: float = 1e-3
    eps_opt: float = 1e-3
    outer_iter: int = 7
    inner_iter: int = 20
    seed: int = 100
    dual_reset: bool = True
    inner_print: bool = False
    outer_print: bool = False

    def to_c(self) -> _capi.OptionsC:
        o = _capi.OptionsC()
        for name in ("reg_0", "alpha_decrease", "beta", "delta_min", "rho_0", "rho_increase", "rho_max", "lambda_max",
                     "alpha_dual", "active_set_tolerance", "eps_dyn", "eps_sta", "eps_con", "eps_opt"):
            setattr(o, name, float(getattr(self, name)))
        for name in ("ls_iter", "outer_iter", "inner_iter"):
            setattr(o, name, int(getattr(self, name)))
        o.regularize, o.dual_reset = int(self.regularize), int(self.dual_reset)
        ax = list(self.alphax_dual) + [1.0] * _capi.MAX_P
        for i in range(_capi.MAX_P):
            o.alphax_dual[i] = float(ax[i])
        return o

    def to_dict(self) -> dict:
        return {f.name: (list(getattr(self, f.name)) if f.name == "alphax_dual" else getattr(self, f.name))
                for f in fields(self)}


@dataclass
class IBROptions:
    """struct/options.jl:123-136.  `ordering` lists players 0-based (the reference's default is 1:100)."""
    ibr_iter: int = 100
    ordering: Optional[List[int]] = None
    delta_min: float = 1e-9

    def to_c(self, p: int) -> _capi.IBROptionsC:
        o = _capi.IBROptionsC()
        o.ibr_iter, o.delta_min = int(self.ibr_iter), float(self.delta_min)
        order = list(range(p)) if self.ordering is None else list(self.ordering)[:p]
        if sorted(order) != list(range(p)):
            raise ValueError("ordering must be a permutation of the players")
        for i in range(_capi.MAX_P):
            o.ordering[i] = order[i] if i < p else 0
        return o


# ------------------------------------------------------------------------------------------------------------
# Objective (objective/objective.jl:6-35, :84-100)
# ------------------------------------------------------------------------------------------------------------
def _diag(v, k):
    v = np.asarray(v, float)
    if v.ndim == 2:
        if not np.allclose(v, np.diag(np.diag(v))):
            raise NotImplementedError("only diagonal Q / R are supported (the reference builds Diagonal LQR costs)")
        v = np.diag(v)
    if v.shape != (k,):
        raise ValueError(f"expected {k} weights, got shape {v.shape}")
    return v.copy()


class GameObjective:
    """Per-player LQR cost (+ optional soft collision cost).  Q[i] (4,), R[i] (2,), xf[i] (4,), uf[i] (2,)."""

    def __init__(self, Q, R, xf, uf, N: int, model: _GameModel):
        p = model.p
        if not (len(Q) == len(R) == len(xf) == len(uf) == p):
            raise ValueError("Q, R, xf, uf must have one entry per player")
        self.p, self.N, self.model = p, N, model
        ni, mi = model.ni[0], model.mi[0]
        self.Q = [_diag(Q[i], ni) for i in range(p)]
        self.R = [_diag(R[i], mi) for i in range(p)]
        self.xf = [np.asarray(xf[i], float).reshape(ni).copy() for i in range(p)]
        self.uf = [np.asarray(uf[i], float).reshape(mi).copy() for i in range(p)]
        self.collision_cost = None     # (radius[p], mu[p])


def add_collision_cost(game_obj: GameObjective, radius, mu):
    """add_collision_cost!(game_obj, radius, μ) (objective.jl:84-100): every ordered pair (i, j≠i), all knots."""
    radius = np.broadcast_to(np.asarray(radius, float), (game_obj.p,)).copy()
    mu = np.broadcast_to(np.asarray(mu, float), (game_obj.p,)).copy()
    game_obj.collision_cost = (radius, mu)


# ------------------------------------------------------------------------------------------------------------
# Constraints (constraints/game_constraints.jl, constraints_methods.jl:5-195)
# ------------------------------------------------------------------------------------------------------------
@dataclass
class Wall:
    """constraints_methods.jl:155-159."""
    p1: Sequence[float]
    p2: Sequence[float]
    v: Sequence[float]


@dataclass
class Wall3D:
    """constraints_methods.jl:201-206: corner points p1, p2, p3 of the rectangle and its outward normal v."""
    p1: Sequence[float]
    p2: Sequence[float]
    p3: Sequence[float]
    v: Sequence[float]


@dataclass
class CylinderWall:
    """constraints_methods.jl:249-254: base point p, axis v in {"x", "y", "z"}, length l, radius r."""
    p: Sequence[float]
    v: str
    l: float
    r: float


class GameConstraintValues:
    """Constraint schema of a game.  Row order of the AL multipliers is canonical (see include/algames_b200.h)."""

    def __init__(self, probsize: ProblemSize):
        self.probsize = probsize
        p = probsize.p
        self.col_radius = np.zeros((p, p))
        self.control_bound = None                  # (u_max[m], u_min[m])
        self.state_bound: List[List[tuple]] = [[] for _ in range(p)]   # StateBound convals (x_max[n], x_min[n]) of player i
        self.walls: List[List[Wall]] = [[] for _ in range(p)]
        self.circles: List[List[tuple]] = [[] for _ in range(p)]
        self.spherical = False                     # add_spherical_collision_avoidance: col_radius acts on three components
        self.walls3d: List[List[Wall3D]] = [[] for _ in range(p)]
        self.cylinders: List[List[CylinderWall]] = [[] for _ in range(p)]


def add_collision_avoidance(game_con: GameConstraintValues, radius, i: Optional[int] = None, j: Optional[int] = None):
    """add_collision_avoidance!(game_con, radius) / (game_con, i, j, radius) — players 0-based here.
    A scalar or per-player radius gives the pair radius r_i + r_j (constraints_methods.jl:21-39)."""
    p = game_con.probsize.p
    if i is not None:
        game_con.col_radius[i, j] = float(radius)
        return
    r = np.broadcast_to(np.asarray(radius, float), (p,))
    for a in range(p):
        for b in range(p):
            if a != b:
                game_con.col_radius[a, b] = r[a] + r[b]


def add_spherical_collision_avoidance(game_con: GameConstraintValues, radius, i: Optional[int] = None, j: Optional[int] = None):
    """add_spherical_collision_avoidance!(game_con, radius) / (game_con, i, j, radius) (constraints_methods.jl:45-81): the same
    CollisionConstraint on the first three state components.  One game uses either the planar or the spherical form."""
    if game_con.col_radius.any() and not game_con.spherical:
        raise NotImplementedError("planar and spherical collision avoidance in one game")
    game_con.spherical = True
    add_collision_avoidance(game_con, radius, i, j)


def _check_bounds(hi, lo, k):
    hi = np.broadcast_to(np.asarray(hi, float), (k,)).copy()
    lo = np.broadcast_to(np.asarray(lo, float), (k,)).copy()
    if not np.all(hi >= lo):
        raise ValueError("Upper bounds must be greater than or equal to lower bounds")   # control_bound_constraint.jl:62-68
    return hi, lo


def add_control_bound(game_con: GameConstraintValues, u_max, u_min):
    if game_con.control_bound is not None:
        raise NotImplementedError("one control bound per game")
    game_con.control_bound = _check_bounds(u_max, u_min, game_con.probsize.m)


def _merge_state_bounds(bounds, n):
    """Component-wise union of one player's StateBound convals: (x_max, x_min, owner_of_max, owner_of_min).
    The device keeps one max and one min row per (player, component); a second conval bounding the same
    component from the same side has no slot there."""
    hi, lo = np.full(n, np.inf), np.full(n, -np.inf)
    hi_con, lo_con = np.zeros(n, int), np.zeros(n, int)
    for g, (h, l) in enumerate(bounds):
        for a in range(n):
            if np.isfinite(h[a]):
                if np.isfinite(hi[a]):
                    raise NotImplementedError("two state bounds of one player on the same component and side")
                hi[a], hi_con[a] = h[a], g
            if np.isfinite(l[a]):
                if np.isfinite(lo[a]):
                    raise NotImplementedError("two state bounds of one player on the same component and side")
                lo[a], lo_con[a] = l[a], g
    return hi, lo, hi_con, lo_con


def add_state_bound(game_con: GameConstraintValues, i: int, x_max, x_min):
    """add_state_bound!(game_con, i, x_max, x_min) (constraints_methods.jl:87-98): one more StateBoundConstraint
    conval on knots 2:N for player i (0-based here)."""
    n = game_con.probsize.n
    bound = _check_bounds(x_max, x_min, n)
    _merge_state_bounds(game_con.state_bound[i] + [bound], n)       # raises if the schema has no device form
    game_con.state_bound[i].append(bound)


def velocity_index(model: _GameModel, i: int) -> int:
    """velocity_index(model, i) (velocity_constraint.jl:30-43), 0-based joint state index of player i's speed."""
    if not 0 <= i < model.p:
        raise AssertionError("player out of range")
    if isinstance(model, UnicycleGame):
        return model.pz[i][3]
    if isinstance(model, BicycleGame):
        return model.pz[i][2]
    raise NotImplementedError(f"Velocity Index is not implemented for {type(model).__name__}.")


def add_velocity_bound(model: _GameModel, game_con: GameConstraintValues, v_max, v_min, i: Optional[int] = None):
    """add_velocity_bound!(model, game_con, v_max, v_min) / (model, game_con, i, v_max, v_min)
    (velocity_constraint.jl:1-28).  Bounding player i's speed adds one StateBoundConstraint on that joint
    component to EVERY player's state constraint list; the vector form skips players whose bounds are both
    infinite."""
    p, n = model.p, model.n
    if i is None:
        v_max, v_min = np.asarray(v_max, float), np.asarray(v_min, float)
        if not (len(v_max) == len(v_min) == p):
            raise AssertionError("v_max and v_min need one entry per player")
        for a in range(p):
            if v_max[a] != np.inf or v_min[a] != -np.inf:
                add_velocity_bound(model, game_con, float(v_max[a]), float(v_min[a]), i=a)
        return
    if not (v_max != np.inf or v_min != -np.inf):
        raise AssertionError("at least one of v_max, v_min must be finite")
    x_max, x_min = np.full(n, np.inf), np.full(n, -np.inf)
    vi = velocity_index(model, i)
    x_max[vi], x_min[vi] = v_max, v_min
    for j in range(p):
        add_state_bound(game_con, j, x_max, x_min)


def add_circle_constraint(game_con: GameConstraintValues, xc, yc, radius, i: Optional[int] = None):
    for a in (range(game_con.probsize.p) if i is None else [i]):
        for t in zip(xc, yc, radius):
            game_con.circles[a].append(tuple(float(v) for v in t))
        if len(game_con.circles[a]) > _capi.MAX_CIRCLES:
            raise ValueError(f"at most {_capi.MAX_CIRCLES} circles per player")


def add_wall_constraint(game_con: GameConstraintValues, walls: Sequence, i: Optional[int] = None):
    """add_wall_constraint!(game_con, [i,] walls) for Vector{Wall} (constraints_methods.jl:161-195), Vector{Wall3D} (:208-247)
    or Vector{CylinderWall} (:256-285)."""
    kind = {Wall: "walls", Wall3D: "walls3d", CylinderWall: "cylinders"}[type(walls[0])]
    for a in (range(game_con.probsize.p) if i is None else [i]):
        getattr(game_con, kind)[a].extend(walls)
        if len(getattr(game_con, kind)[a]) > _capi.MAX_WALLS:
            raise ValueError(f"at most {_capi.MAX_WALLS} walls per player")


# ------------------------------------------------------------------------------------------------------------
# Descriptor packing
# ------------------------------------------------------------------------------------------------------------
def _joint(per_player, p, k):
    """per-player (k,) vectors -> component-major joint vector (objective.jl:24-28 expand_vector)."""
    out = np.zeros(k * p)
    for i in range(p):
        out[[i + c * p for c in range(k)]] = per_player[i]
    return out


def _make_desc(model, N, dt, game_obj: GameObjective, game_con: GameConstraintValues, solver: int = 0) -> _capi.ProblemDesc:
    d = _capi.ProblemDesc()
    p, n, m = model.p, model.n, model.m
    ni, mi = model.ni[0], model.mi[0]
    d.model, d.p, d.d, d.N, d.dt = _capi.MODEL_IDS[model.name], p, getattr(model, "d", 2), N, dt
    d.lf, d.lr = getattr(model, "lf", 0.05), getattr(model, "lr", 0.05)
    d.quad_mass, d.solver, d.spherical_collision = getattr(model, "mass", 0.5), int(solver), int(game_con.spherical)
    for i in range(p):
        d.n_walls3d[i] = len(game_con.walls3d[i])
        for q, w in enumerate(game_con.walls3d[i]):
            for e, v in enumerate([*w.p1, *w.p2, *w.p3, *w.v]):
                d.walls3d[i][q][e] = float(v)
        d.n_cylinders[i] = len(game_con.cylinders[i])
        for q, w in enumerate(game_con.cylinders[i]):
            for e, v in enumerate([*w.p, "xyz".index(w.v), w.l, w.r]):
                d.cylinders[i][q][e] = float(v)
    for name, vec in (("Q", _joint(game_obj.Q, p, ni)), ("xf", _joint(game_obj.xf, p, ni)),
                      ("R", _joint(game_obj.R, p, mi)), ("uf", _joint(game_obj.uf, p, mi))):
        arr = getattr(d, name)
        for a, v in enumerate(vec):
            arr[a] = v
    if game_obj.collision_cost is not None:
        d.has_collision_cost = 1
        for i in range(p):
            d.cc_radius[i], d.cc_mu[i] = game_obj.collision_cost[0][i], game_obj.collision_cost[1][i]
    for i in range(p):
        for j in range(p):
            d.col_radius[i][j] = game_con.col_radius[i, j]
        if game_con.state_bound[i]:
            d.has_state_bound[i] = len(game_con.state_bound[i])
            hi, lo, hi_con, lo_con = _merge_state_bounds(game_con.state_bound[i], n)
            for a in range(n):
                d.x_max[i][a], d.x_min[i][a] = hi[a], lo[a]
                d.x_max_con[i][a], d.x_min_con[i][a] = int(hi_con[a]), int(lo_con[a])
        d.n_walls[i] = len(game_con.walls[i])
        for q, w in enumerate(game_con.walls[i]):
            for e, v in enumerate([w.p1[0], w.p1[1], w.p2[0], w.p2[1], w.v[0], w.v[1]]):
                d.walls[i][q][e] = float(v)
        d.n_circles[i] = len(game_con.circles[i])
        for q, c in enumerate(game_con.circles[i]):
            for e in range(3):
                d.circles[i][q][e] = c[e]
    if game_con.control_bound is not None:
        d.has_control_bound = 1
        for a in range(m):
            d.u_max[a], d.u_min[a] = game_con.control_bound[0][a], game_con.control_bound[1][a]
    return d


def _sb_spec(bounds):
    items = [{"x_max": hi.tolist(), "x_min": lo.tolist()} for hi, lo in bounds]
    return None if not items else (items[0] if len(items) == 1 else items)


def spec_of(prob: "GameProblem") -> dict:
    """Neutral plain-dict description of a problem (what the tests hand to the oracle's problem_from_spec)."""
    model, obj, con = prob.model, prob.game_obj, prob.game_con
    p = model.p
    return {
        "model": model.name, "p": p, "d": getattr(model, "d", 2), "lf": getattr(model, "lf", 0.05), "mass": getattr(model, "mass", 0.5),
        "spherical": bool(con.spherical),
        "walls3d": [[[*map(float, w.p1), *map(float, w.p2), *map(float, w.p3), *map(float, w.v)] for w in con.walls3d[i]] for i in range(p)],
        "cylinders": [[[*map(float, w.p), w.v, float(w.l), float(w.r)] for w in con.cylinders[i]] for i in range(p)],
        "lr": getattr(model, "lr", 0.05), "N": prob.probsize.N, "dt": prob.dt, "x0": np.asarray(prob.x0, float).tolist(),
        "Q": [q.tolist() for q in obj.Q], "R": [r.tolist() for r in obj.R],
        "xf": [x.tolist() for x in obj.xf], "uf": [u.tolist() for u in obj.uf],
        "collision_cost": None if obj.collision_cost is None else
        {"radius": obj.collision_cost[0].tolist(), "mu": obj.collision_cost[1].tolist()},
        "collision_radius": con.col_radius.tolist() if con.col_radius.any() else None,
        # per player: None, one {"x_max", "x_min"} dict, or the list of them in the order they were added
        "state_bounds": [_sb_spec(sbs) for sbs in con.state_bound],
        "walls": [[[w.p1[0], w.p1[1], w.p2[0], w.p2[1], w.v[0], w.v[1]] for w in con.walls[i]] for i in range(p)],
        "circles": [[list(c) for c in con.circles[i]] for i in range(p)],
        "control_bounds": None if con.control_bound is None else
        {"u_max": con.control_bound[0].tolist(), "u_min": con.control_bound[1].tolist()},
        "opts": prob.opts.to_dict(),
    }


# ------------------------------------------------------------------------------------------------------------
# Batch of games sharing one schema: thin object wrapper over an agb_handle
# ------------------------------------------------------------------------------------------------------------
class AlgamesError(RuntimeError):
    pass


class GameBatch:
    """`batch` independent GameProblems with one schema (model, players, horizon, constraint lists) resident on one GPU.

    Arrays use the ABI layouts: x0 [B,n], Z [B,N,n+m], L [B,p,N-1,n], conlam/conmu [B,N-1,nrow], res/dtraj [B,S].
    """

    def __init__(self, model, N, dt, game_obj, game_con, batch: int, device: int = 0, lib_path: Optional[str] = None,
                 solver: int = _capi.SOLVER_AUTO):
        """`solver`: _capi.SOLVER_AUTO (structured kernels where the schema allows, band solver otherwise and as the fallback
        for singular stage systems) or _capi.SOLVER_BAND (explicit KKT band + pivoted LU for every instance)."""
        self.lib = _capi.load(lib_path)
        self.model, self.N, self.dt, self.batch, self.device = model, N, dt, batch, device
        self.probsize = ProblemSize(N, model)
        self.desc = _make_desc(model, N, dt, game_obj, game_con, solver)
        h = C.c_void_p()
        rc = self.lib.agb_create(C.byref(self.desc), batch, device, C.byref(h))
        if rc != 0:
            raise AlgamesError(f"agb_create failed ({rc}): {self.lib.agb_last_error(None).decode()}")
        self.h = h
        sz = _capi.Sizes()
        self._ck(self.lib.agb_get_sizes(self.h, C.byref(sz)))
        self.sizes = sz
        self.n, self.m, self.p, self.S, self.nrow = sz.n, sz.m, sz.p, sz.S, sz.nrow

    def _ck(self, rc):
        if rc != 0:
            raise AlgamesError(f"libalgames_b200 error {rc}: {self.lib.agb_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.agb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- shapes
    def _z(self):
        return (self.batch, self.N, self.n + self.m)

    def _l(self):
        return (self.batch, self.p, self.N - 1, self.n)

    def _c(self):
        return (self.batch, self.N - 1, self.nrow)

    @staticmethod
    def _arr(a, shape):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
        return a

    # ---- state
    def set_instance_params(self, x0=None, xf=None, Q=None, R=None, uf=None):
        B, n, m = self.batch, self.n, self.m
        a = [self._arr(x0, (B, n)), self._arr(xf, (B, n)), self._arr(Q, (B, n)), self._arr(R, (B, m)), self._arr(uf, (B, m))]
        self._ck(self.lib.agb_set_instance_params(self.h, *[_capi.dptr(v) for v in a]))

    def set_initial(self, Z0, L0, conlam=None, conmu=None):
        a = [self._arr(Z0, self._z()), self._arr(L0, self._l()), self._arr(conlam, self._c()), self._arr(conmu, self._c())]
        self._ck(self.lib.agb_set_initial(self.h, *[_capi.dptr(v) for v in a]))

    def random_initial(self, amplitude=1e-8, seed=100):
        """init_traj!(…; f=rand, amplitude) for every instance (primal_dual_traj.jl:29-44).  The reference draws from
        Julia's MersenneTwister(seed); that stream is not reproducible here, numpy's default_rng(seed) is used instead."""
        rng = np.random.default_rng(seed)
        Z0 = amplitude * rng.random(self._z())
        L0 = amplitude * rng.random(self._l())
        self.set_initial(Z0, L0)
        return Z0, L0

    def shift_initial(self, s: int, Zfresh=None, Lfresh=None):
        a = [self._arr(Zfresh, self._z()), self._arr(Lfresh, self._l())]
        self._ck(self.lib.agb_shift_initial(self.h, int(s), *[_capi.dptr(v) for v in a]))

    def mpc_advance(self, s: int = 1, disturbance=None, Zfresh=None, Lfresh=None):
        """x0 <- x_{1+s} of the resident solution (+ disturbance [B,n]) and init_traj! shift by s, all on device."""
        a = [self._arr(disturbance, (self.batch, self.n)), self._arr(Zfresh, self._z()), self._arr(Lfresh, self._l())]
        self._ck(self.lib.agb_mpc_advance(self.h, int(s), *[_capi.dptr(v) for v in a]))

    def mpc_advance_async(self, s: int = 1, disturbance_dev: int = 0):
        """agb_mpc_advance_async: the same step with a DEVICE pointer for the disturbance (0 = none) and no host sync."""
        self._ck(self.lib.agb_mpc_advance_async(self.h, int(s), C.c_void_p(disturbance_dev) if disturbance_dev else None))

    def mpc_run(self, opts: "Options", resolves: int, s: int = 1, disturbances=None, want_xs: bool = True):
        """agb_mpc_run: the whole receding-horizon loop of every stream in one launch.  disturbances [resolves, B, n] or None.
        Returns per-re-solve stats [resolves, B, 10], status [resolves, B] and the executed states [resolves, B, n]."""
        R, B = int(resolves), self.batch
        d = self._arr(disturbances, (R, B, self.n))
        stats = np.empty((R, B, _capi.NSTATS)); status = np.empty((R, B), dtype=np.int32)
        xs = np.empty((R, B, self.n)) if want_xs else None
        o = opts.to_c()
        self._ck(self.lib.agb_mpc_run(self.h, C.byref(o), R, int(s), _capi.dptr(d), _capi.dptr(stats),
                                      status.ctypes.data_as(C.POINTER(C.c_int)), _capi.dptr(xs)))
        return stats, status, xs

    def mpc_run_async(self, opts: "Options", resolves: int, s: int, disturbance_dev: int, stats_dev: int, status_dev: int,
                      xs_dev: int = 0, stream: int = 0):
        """agb_mpc_run_async: device pointers (integers) in and out, nothing synchronises."""
        o = opts.to_c()
        vp = lambda a: C.c_void_p(a) if a else None
        self._ck(self.lib.agb_mpc_run_async(self.h, C.byref(o), int(resolves), int(s), vp(disturbance_dev), vp(stats_dev),
                                            vp(status_dev), vp(xs_dev), vp(stream)))

    def stream(self) -> int:
        """agb_get_stream: the handle's own cudaStream_t (as an integer)."""
        return int(self.lib.agb_get_stream(self.h) or 0)

    def join_stream(self, stream: int):
        """agb_join_stream: `stream` waits (on the device) for everything enqueued on the handle's own stream."""
        self._ck(self.lib.agb_join_stream(self.h, C.c_void_p(stream) if stream else None))

    def get_state(self):
        Z, L = np.empty(self._z()), np.empty(self._l())
        cl, cm = np.empty(self._c()), np.empty(self._c())
        self._ck(self.lib.agb_get_state(self.h, _capi.dptr(Z), _capi.dptr(L), _capi.dptr(cl), _capi.dptr(cm)))
        return Z, L, cl, cm

    # ---- per-function entry points
    def rollout(self):
        self._ck(self.lib.agb_rollout(self.h))

    def residual(self, reg_x=0.0, reg_u=0.0, alpha=0.0, want_res=True):
        res = np.empty((self.batch, self.S)) if want_res else None
        norms = np.empty((self.batch, 5))
        self._ck(self.lib.agb_residual(self.h, reg_x, reg_u, alpha, _capi.dptr(res), _capi.dptr(norms)))
        return res, norms

    def residual_jacobian_dense(self, reg_x=0.0, reg_u=0.0):
        J = np.empty((self.batch, self.S, self.S))
        self._ck(self.lib.agb_residual_jacobian_dense(self.h, reg_x, reg_u, _capi.dptr(J)))
        return J

    def kkt_solve(self, reg_x=0.0, reg_u=0.0):
        d = np.empty((self.batch, self.S))
        self._ck(self.lib.agb_kkt_solve(self.h, reg_x, reg_u, _capi.dptr(d)))
        return d

    def line_search(self, opts: Options, reg: float):
        alpha = np.empty(self.batch)
        j = np.empty(self.batch, dtype=np.int32)
        oc = opts.to_c()
        self._ck(self.lib.agb_line_search(self.h, C.byref(oc), reg, reg, _capi.dptr(alpha), _capi.iptr(j)))
        return alpha, j

    def update_traj(self, alpha):
        alpha = self._arr(np.broadcast_to(np.asarray(alpha, float), (self.batch,)), (self.batch,))
        d = np.empty(self.batch)
        self._ck(self.lib.agb_update_traj(self.h, _capi.dptr(alpha), _capi.dptr(d)))
        return d

    def dual_update(self, opts: Options):
        oc = opts.to_c()
        self._ck(self.lib.agb_dual_update(self.h, C.byref(oc)))

    def penalty_update(self, opts: Options):
        oc = opts.to_c()
        self._ck(self.lib.agb_penalty_update(self.h, C.byref(oc)))

    def reset(self, opts: Options):
        oc = opts.to_c()
        self._ck(self.lib.agb_reset_duals_penalties(self.h, C.byref(oc)))

    def evaluate(self):
        c = np.empty(self._c())
        self._ck(self.lib.agb_evaluate_constraints(self.h, _capi.dptr(c)))
        return c

    def active_set(self, tol):
        a = np.empty(self._c(), dtype=np.uint8)
        self._ck(self.lib.agb_active_set(self.h, float(tol), a.ctypes.data_as(C.POINTER(C.c_ubyte))))
        return a.astype(bool)

    # ---- active-set analysis (src/active_set/*.jl) --------------------------------------------------------------
    def active_set_sizes(self):
        """(Sv, Sh) of ActiveSetCore (active_set_core.jl:81-82)."""
        sv, sh = C.c_int(0), C.c_int(0)
        self._ck(self.lib.agb_active_set_sizes(self.h, C.byref(sv), C.byref(sh)))
        return sv.value, sh.value

    def active_set_residual(self):
        sv, _ = self.active_set_sizes()
        r = np.empty((self.batch, sv))
        self._ck(self.lib.agb_active_set_residual(self.h, _capi.dptr(r)))
        return r

    def active_set_jacobian(self):
        sv, sh = self.active_set_sizes()
        J = np.empty((self.batch, sv, sh))
        self._ck(self.lib.agb_active_set_jacobian_dense(self.h, _capi.dptr(J)))
        return J

    def active_set_masks(self, tol):
        sv, sh = self.active_set_sizes()
        vm, hm = np.empty((self.batch, sv), dtype=np.uint8), np.empty((self.batch, sh), dtype=np.uint8)
        ub = C.POINTER(C.c_ubyte)
        self._ck(self.lib.agb_active_set_masks(self.h, float(tol), vm.ctypes.data_as(ub), hm.ctypes.data_as(ub)))
        return vm.astype(bool), hm.astype(bool)

    def update_nullspace(self, tol, atol=1e-20, max_dim=None):
        """agb_update_nullspace: per instance an orthonormal basis [dim, Sh] of nullspace(jac[vmask, hmask]) (rows of the inactive
        columns are zero).  Returns a list of arrays, one per instance."""
        sv, sh = self.active_set_sizes()
        if max_dim is None:
            max_dim = sh - self.S + (self.sizes.N - 1) * self.p          # every pair active gives Sh − Sv; leave room for rank deficiency
        out = np.zeros((self.batch, max_dim, sh)); dim = np.zeros(self.batch, dtype=np.int32)
        self._ck(self.lib.agb_update_nullspace(self.h, float(tol), float(atol), int(max_dim), _capi.dptr(out), _capi.iptr(dim)))
        if (dim > max_dim).any():
            raise AlgamesError(f"null space of dimension {int(dim.max())} exceeds max_dim = {max_dim}")
        return [out[b, :dim[b]].copy() for b in range(self.batch)]

    def band_info(self):
        """agb_band_info: (bytes of the band solver's shared-memory elimination window per CTA, band scratch slots)."""
        w = C.c_int(0); sl = C.c_int(0)
        self._ck(self.lib.agb_band_info(self.h, C.byref(w), C.byref(sl)))
        return w.value, sl.value

    def debug_gain_solve(self, aug):
        """agb_debug_gain_solve: the kernel's Gauss-Jordan on aug [B, m, m+n+1]; returns (reduced systems, ok flags)."""
        aug = self._arr(aug, (self.batch, self.m, self.m + self.n + 1))
        out = np.empty_like(aug)
        ok = np.empty(self.batch, dtype=np.int32)
        self._ck(self.lib.agb_debug_gain_solve(self.h, _capi.dptr(aug), _capi.dptr(out), _capi.iptr(ok)))
        return out, ok.astype(bool)

    # ---- the hot path
    def newton_solve(self, opts: Options, want=("Z", "L", "conlam", "conmu", "stats", "status"), out=None):
        """agb_newton_solve_batch: solve every instance and copy the requested results to host arrays (`out` may hold
        preallocated, e.g. pinned, arrays of the right shapes)."""
        if out is not None:
            oc = opts.to_c()
            self._ck(self.lib.agb_newton_solve_batch(
                self.h, C.byref(oc), _capi.dptr(out.get("Z")), _capi.dptr(out.get("L")), _capi.dptr(out.get("conlam")),
                _capi.dptr(out.get("conmu")), _capi.dptr(out.get("stats")), _capi.iptr(out.get("status"))))
            return out
        out = {}
        if "Z" in want: out["Z"] = np.empty(self._z())
        if "L" in want: out["L"] = np.empty(self._l())
        if "conlam" in want: out["conlam"] = np.empty(self._c())
        if "conmu" in want: out["conmu"] = np.empty(self._c())
        if "stats" in want: out["stats"] = np.empty((self.batch, _capi.NSTATS))
        if "status" in want: out["status"] = np.empty(self.batch, dtype=np.int32)
        oc = opts.to_c()
        self._ck(self.lib.agb_newton_solve_batch(
            self.h, C.byref(oc), _capi.dptr(out.get("Z")), _capi.dptr(out.get("L")), _capi.dptr(out.get("conlam")),
            _capi.dptr(out.get("conmu")), _capi.dptr(out.get("stats")), _capi.iptr(out.get("status"))))
        return out

    def set_history(self, max_records: int):
        """agb_set_history: keep every record!(stats, …) of later newton_solve calls (statistics.jl:44-57); 0 switches it off."""
        self._ck(self.lib.agb_set_history(self.h, int(max_records)))
        self._hist_max = int(max_records)

    def get_history(self):
        """agb_get_history → (hist [B, max_records, 8], count [B]); columns: outer k, res, dyn, con, sta, opt, Δ_traj, inner l."""
        hist = np.empty((self.batch, self._hist_max, _capi.NHIST))
        count = np.empty(self.batch, dtype=np.int32)
        self._ck(self.lib.agb_get_history(self.h, _capi.dptr(hist), _capi.iptr(count)))
        return hist, count

    def ibr_newton_solve(self, opts: Options, ibr_opts: Optional[IBROptions] = None):
        """agb_ibr_newton_solve_batch: ibr_newton_solve!(prob; ibr_opts) for every instance (solver_methods.jl:133-166)."""
        out = {"Z": np.empty(self._z()), "L": np.empty(self._l()), "conlam": np.empty(self._c()), "conmu": np.empty(self._c()),
               "stats": np.empty((self.batch, _capi.NSTATS)), "status": np.empty(self.batch, dtype=np.int32)}
        oc, ic = opts.to_c(), (ibr_opts or IBROptions()).to_c(self.p)
        self._ck(self.lib.agb_ibr_newton_solve_batch(
            self.h, C.byref(oc), C.byref(ic), _capi.dptr(out["Z"]), _capi.dptr(out["L"]), _capi.dptr(out["conlam"]),
            _capi.dptr(out["conmu"]), _capi.dptr(out["stats"]), _capi.iptr(out["status"])))
        return out

    def ibr_residual(self, player: int, reg_x=0.0, reg_u=0.0, alpha=0.0):
        res, norms = np.empty((self.batch, self.S)), np.empty((self.batch, 5))
        self._ck(self.lib.agb_ibr_residual(self.h, int(player), reg_x, reg_u, alpha, _capi.dptr(res), _capi.dptr(norms)))
        return res, norms

    def ibr_kkt_solve(self, player: int, reg_x=0.0, reg_u=0.0):
        d = np.empty((self.batch, self.S))
        self._ck(self.lib.agb_ibr_kkt_solve(self.h, int(player), reg_x, reg_u, _capi.dptr(d)))
        return d

    def solve_from_host(self, opts: Options, x0, Z0, L0, out=None):
        """agb_solve_from_host: one call from host buffers to host results, copies pipelined with the solve in chunks."""
        x0, Z0, L0 = self._arr(x0, (self.batch, self.n)), self._arr(Z0, self._z()), self._arr(L0, self._l())
        if out is None:
            out = {"Z": np.empty(self._z()), "L": np.empty(self._l()), "conlam": np.empty(self._c()), "conmu": np.empty(self._c()),
                   "stats": np.empty((self.batch, _capi.NSTATS)), "status": np.empty(self.batch, dtype=np.int32)}
        oc = opts.to_c()
        self._ck(self.lib.agb_solve_from_host(
            self.h, C.byref(oc), _capi.dptr(x0), _capi.dptr(Z0), _capi.dptr(L0), _capi.dptr(out.get("Z")), _capi.dptr(out.get("L")),
            _capi.dptr(out.get("conlam")), _capi.dptr(out.get("conmu")), _capi.dptr(out.get("stats")), _capi.iptr(out.get("status"))))
        return out

    def newton_solve_async(self, opts: Options, stream: int = 0):
        oc = opts.to_c()
        self._ck(self.lib.agb_newton_solve_async(self.h, C.byref(oc), C.c_void_p(stream)))

    def device_view(self) -> _capi.DeviceView:
        v = _capi.DeviceView()
        self._ck(self.lib.agb_get_device_view(self.h, C.byref(v)))
        return v

    def violations(self):
        """agb_violations: per-knot vectors of dynamics_violation / control_violation / state_violation / optimality_violation
        at the resident iterate (struct/violations.jl): dyn [B,N-1], con [B,N-1], sta [B,N], opt [B,N]."""
        B, N = self.batch, self.N
        dyn, con, sta, opt = np.empty((B, N - 1)), np.empty((B, N - 1)), np.empty((B, N)), np.empty((B, N))
        self._ck(self.lib.agb_violations(self.h, _capi.dptr(dyn), _capi.dptr(con), _capi.dptr(sta), _capi.dptr(opt)))
        return dyn, con, sta, opt

    # ---- multi-GPU: the path's single all-gather, pushed over peer memory by the copy engines (agb_peer_* / agb_allgather)
    def peer_init(self, nranks: int, rank: int, batches):
        b = np.ascontiguousarray(batches, dtype=np.int32)
        assert b.shape == (nranks,)
        self._ck(self.lib.agb_peer_init(self.h, int(nranks), int(rank), _capi.iptr(b)))
        self._peer = (int(nranks), int(rank), b.copy())

    def peer_export(self) -> bytes:
        buf = (C.c_ubyte * _capi.IPC_BYTES)()
        self._ck(self.lib.agb_peer_export(self.h, buf))
        return bytes(buf)

    def peer_connect(self, handles: Sequence[bytes]):
        blob = b"".join(handles)
        assert len(blob) == self._peer[0] * _capi.IPC_BYTES
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._ck(self.lib.agb_peer_connect(self.h, buf))

    @staticmethod
    def peer_connect_local(batches: Sequence["GameBatch"]):
        lib = batches[0].lib
        arr = (C.c_void_p * len(batches))(*[b.h for b in batches])
        rc = lib.agb_peer_connect_local(arr, len(batches))
        if rc != 0:
            raise AlgamesError(f"agb_peer_connect_local failed ({rc}): {lib.agb_last_error(None).decode()}")

    def allgather(self, stream: Optional[int] = None):
        self._ck(self.lib.agb_allgather(self.h, C.c_void_p(stream) if stream else None))

    def allgather_wait(self):
        self._ck(self.lib.agb_allgather_wait(self.h))

    def gathered_view(self):
        """(device pointer of this rank's gather buffer, byte offsets [nranks+1] of the ranks' slabs)."""
        ptr = C.c_void_p()
        offs = (C.c_ulonglong * (self._peer[0] + 1))()
        self._ck(self.lib.agb_gathered_view(self.h, C.byref(ptr), offs))
        return ptr.value, list(offs)

    def unpack_gathered(self, src_rank: int):
        """Rank `src_rank`'s results as they arrived in this rank's gather buffer."""
        Bs = int(self._peer[2][src_rank])
        out = {"Z": np.empty((Bs, self.N, self.n + self.m)), "L": np.empty((Bs, self.p, self.N - 1, self.n)),
               "stats": np.empty((Bs, _capi.NSTATS)), "status": np.empty(Bs, dtype=np.int32)}
        self._ck(self.lib.agb_unpack_gathered(self.h, int(src_rank), _capi.dptr(out["Z"]), _capi.dptr(out["L"]),
                                              _capi.dptr(out["stats"]), _capi.iptr(out["status"])))
        return out

    def launch_count(self) -> int:
        return int(self.lib.agb_launch_count(self.h))

    def last_solve_ms(self) -> float:
        return float(self.lib.agb_last_solve_ms(self.h))


# ------------------------------------------------------------------------------------------------------------
# GameProblem: the reference's single-problem object (problem/problem.jl:19-53)
# ------------------------------------------------------------------------------------------------------------
@dataclass
class Violation:
    """DynamicsViolation / ControlViolation / StateViolation / OptimalityViolation (struct/violations.jl:5-16, …): `.max`
    for every record; `.vio` (the per-knot vector) for the final record of a solve (agb_violations)."""
    max: float
    vio: Optional[np.ndarray] = None


class Statistics:
    """struct/statistics.jl:5-57: one entry per record!(stats, …) of the solve (newton_solve fills the whole history from
    the device log; the best-response solver reports its final record only).  `t_elap` is not kept: instances of a batch
    share one kernel launch, its duration is GameBatch.last_solve_ms()."""

    def __init__(self):
        self.iter = 0
        self.outer_iter: List[int] = []
        self.res: List[float] = []
        self.delta: List[float] = []
        self.dyn_vio: List[Violation] = []
        self.con_vio: List[Violation] = []
        self.sta_vio: List[Violation] = []
        self.opt_vio: List[Violation] = []
        self.newton_steps = 0
        self.residual_evals = 0

    def reset(self):
        """reset!(stats) (struct/statistics.jl:74-80)."""
        self.__init__()

    def record_history(self, hist):
        """hist [count, 8] rows of agb_get_history."""
        for h in hist:
            self.iter += 1
            self.outer_iter.append(int(h[0])); self.res.append(float(h[1]))
            self.dyn_vio.append(Violation(float(h[2]))); self.con_vio.append(Violation(float(h[3])))
            self.sta_vio.append(Violation(float(h[4]))); self.opt_vio.append(Violation(float(h[5])))
            self.delta.append(float(h[6]))

    def record(self, st):
        self.iter += 1
        self.res.append(float(st[0])); self.dyn_vio.append(Violation(float(st[1]))); self.con_vio.append(Violation(float(st[2])))
        self.sta_vio.append(Violation(float(st[3]))); self.opt_vio.append(Violation(float(st[4])))
        self.delta.append(float(st[5])); self.newton_steps = int(st[6]); self.outer_iter.append(int(st[7]))
        self.residual_evals = int(st[8])


class _Core:
    def __init__(self, S):
        self.res = np.zeros(S)
        self.jac = None


class _PrimalDualTraj:
    """X [N,n], U [N,m] (last row = the unused terminal control), du [p,N-1,n] (struct/primal_dual_traj.jl:5-20)."""

    def __init__(self, ps: ProblemSize):
        self.X, self.U, self.du = np.zeros((ps.N, ps.n)), np.zeros((ps.N, ps.m)), np.zeros((ps.p, ps.N - 1, ps.n))


class GameProblem:
    """GameProblem(N, dt, x0, model, opts, game_obj, game_con)."""

    def __init__(self, N, dt, x0, model, opts: Options, game_obj: GameObjective, game_con: GameConstraintValues,
                 device: int = 0, lib_path: Optional[str] = None):
        self.probsize = ProblemSize(N, model)
        self.N, self.dt, self.model, self.opts = N, dt, model, opts
        self.x0 = np.asarray(x0, float).reshape(model.n).copy()
        self.game_obj, self.game_con = game_obj, game_con
        self.core = _Core(self.probsize.S)
        self.pdtraj = _PrimalDualTraj(self.probsize)
        self.stats = Statistics()
        self.status = None
        self.conlam = self.conmu = None
        self._device, self._lib_path, self._batch = device, lib_path, None

    def batch(self) -> GameBatch:
        if self._batch is None:
            self._batch = GameBatch(self.model, self.N, self.dt, self.game_obj, self.game_con, 1, self._device, self._lib_path)
        o, p = self.game_obj, self.probsize.p      # the objective may have been edited since the handle was created
        ni, mi = self.probsize.ni[0], self.probsize.mi[0]
        self._batch.set_instance_params(x0=self.x0[None, :], xf=_joint(o.xf, p, ni)[None], Q=_joint(o.Q, p, ni)[None],
                                        R=_joint(o.R, p, mi)[None], uf=_joint(o.uf, p, mi)[None])
        return self._batch

    def _push(self):
        b = self.batch()
        Z = np.concatenate([self.pdtraj.X, self.pdtraj.U], axis=1)[None]
        b.set_initial(Z, self.pdtraj.du[None], None if self.conlam is None else self.conlam[None],
                      None if self.conmu is None else self.conmu[None])
        return b

    def _pull(self, b: GameBatch):
        Z, L, cl, cm = b.get_state()
        n = self.probsize.n
        self.pdtraj.X[:], self.pdtraj.U[:], self.pdtraj.du[:] = Z[0, :, :n], Z[0, :, n:], L[0]
        self.conlam, self.conmu = cl[0], cm[0]


def _same_schema(a: GameProblem, b: GameProblem) -> bool:
    sa, sb = spec_of(a), spec_of(b)
    for k in ("x0", "xf", "Q", "R", "uf", "opts"):
        sa.pop(k); sb.pop(k)
    return sa == sb


def _check_batch(plist):
    """One schema and ONE set of solver options per batch: the kernel takes a single agb_options for all instances."""
    p0 = plist[0]
    for q in plist[1:]:
        if not _same_schema(p0, q):
            raise ValueError("all problems of a batch must share model, sizes and constraint schema")
        if q.opts.to_dict() != p0.opts.to_dict():
            raise ValueError("all problems of a batch must share the same Options (one agb_options per launch)")


def init_traj(prob: GameProblem, rng=None):
    """init_traj!(pdtraj; x0, f=rand, amplitude, s=shift) (primal_dual_traj.jl:29-44), on the host arrays."""
    o, ps, pd = prob.opts, prob.probsize, prob.pdtraj
    rng = rng or np.random.default_rng(o.seed)
    s, N = o.shift, ps.N
    for k in range(1, N + 1):
        if k + s <= N:
            pd.X[k - 1], pd.U[k - 1] = pd.X[k + s - 1], pd.U[k + s - 1]
        else:
            z = o.amplitude_init * rng.random(ps.n + ps.m)
            pd.X[k - 1], pd.U[k - 1] = z[:ps.n], z[ps.n:]
    for i in range(ps.p):
        for k in range(1, N):
            pd.du[i, k - 1] = pd.du[i, k + s - 1] if k + s <= N - 1 else o.amplitude_init * rng.random(ps.n)
    pd.X[0] = prob.x0


def newton_solve(probs, init: bool = True):
    """newton_solve!(prob) for one GameProblem or a list sharing one schema (solver_methods.jl:5-65).
    Mutates prob.pdtraj, prob.stats, prob.core.res, prob.conlam/conmu, prob.status like the reference mutates prob."""
    single = isinstance(probs, GameProblem)
    plist = [probs] if single else list(probs)
    if not plist:
        return None
    p0 = plist[0]
    _check_batch(plist)
    B, ps = len(plist), p0.probsize
    batch = p0.batch() if B == 1 else GameBatch(p0.model, p0.N, p0.dt, p0.game_obj, p0.game_con, B, p0._device, p0._lib_path)
    try:
        if init:
            for q in plist:
                init_traj(q)
        obj = [q.game_obj for q in plist]
        batch.set_instance_params(
            x0=np.stack([q.x0 for q in plist]),
            xf=np.stack([_joint(o.xf, ps.p, ps.ni[0]) for o in obj]), Q=np.stack([_joint(o.Q, ps.p, ps.ni[0]) for o in obj]),
            R=np.stack([_joint(o.R, ps.p, ps.mi[0]) for o in obj]), uf=np.stack([_joint(o.uf, ps.p, ps.mi[0]) for o in obj]))
        Z0 = np.stack([np.concatenate([q.pdtraj.X, q.pdtraj.U], axis=1) for q in plist])
        L0 = np.stack([q.pdtraj.du for q in plist])
        have_duals = all(q.conlam is not None for q in plist)
        batch.set_initial(Z0, L0, np.stack([q.conlam for q in plist]) if have_duals else None,
                          np.stack([q.conmu for q in plist]) if have_duals else None)
        max_rec = p0.opts.outer_iter * p0.opts.inner_iter + 1          # every record!(stats, …) the loop can make
        batch.set_history(max_rec)
        out = batch.newton_solve(p0.opts)
        hist, count = batch.get_history()
        res, _ = batch.residual()
        try:
            vio = batch.violations()
        except AlgamesError:                      # band-solver schemas (QuadrotorGame, 3-D constraints): maxima only
            vio = ()
        for b, q in enumerate(plist):
            q.pdtraj.X[:], q.pdtraj.U[:], q.pdtraj.du[:] = out["Z"][b, :, :ps.n], out["Z"][b, :, ps.n:], out["L"][b]
            q.conlam, q.conmu = out["conlam"][b], out["conmu"][b]
            q.stats = Statistics(); q.stats.record_history(hist[b, :min(int(count[b]), max_rec)])
            for lst, v in zip((q.stats.dyn_vio, q.stats.con_vio, q.stats.sta_vio, q.stats.opt_vio), vio):
                if lst:
                    lst[-1].vio = v[b].copy()             # per-knot vectors of the final record (struct/violations.jl)
            q.stats.newton_steps, q.stats.residual_evals = int(out["stats"][b, 6]), int(out["stats"][b, 8])
            q.status = _capi.STATUS_NAMES[int(out["status"][b])]
            q.core.res[:] = res[b]
    finally:
        if B > 1:
            batch.close()
    return None


def ibr_newton_solve(probs, ibr_opts: Optional[IBROptions] = None, init: bool = True):
    """ibr_newton_solve!(prob; ibr_opts) for one GameProblem or a list sharing one schema (solver_methods.jl:133-166)."""
    single = isinstance(probs, GameProblem)
    plist = [probs] if single else list(probs)
    if not plist:
        return None
    p0 = plist[0]
    _check_batch(plist)
    B, ps = len(plist), p0.probsize
    batch = p0.batch() if B == 1 else GameBatch(p0.model, p0.N, p0.dt, p0.game_obj, p0.game_con, B, p0._device, p0._lib_path)
    try:
        if init:
            for q in plist:
                init_traj(q)
        obj = [q.game_obj for q in plist]
        batch.set_instance_params(
            x0=np.stack([q.x0 for q in plist]),
            xf=np.stack([_joint(o.xf, ps.p, ps.ni[0]) for o in obj]), Q=np.stack([_joint(o.Q, ps.p, ps.ni[0]) for o in obj]),
            R=np.stack([_joint(o.R, ps.p, ps.mi[0]) for o in obj]), uf=np.stack([_joint(o.uf, ps.p, ps.mi[0]) for o in obj]))
        batch.set_initial(np.stack([np.concatenate([q.pdtraj.X, q.pdtraj.U], axis=1) for q in plist]),
                          np.stack([q.pdtraj.du for q in plist]))
        io = ibr_opts or IBROptions()
        # every record!(stats, …, k, i) of the sweeps (statistics.jl:59-72); capped so that ibr_iter = 100 stays cheap
        max_rec = min(io.ibr_iter * ps.p * (p0.opts.outer_iter * p0.opts.inner_iter + 1), 4096)
        batch.set_history(max_rec)
        out = batch.ibr_newton_solve(p0.opts, ibr_opts)
        hist, count = batch.get_history()
        batch.set_history(0)
        res, _ = batch.residual()
        for b, q in enumerate(plist):
            q.pdtraj.X[:], q.pdtraj.U[:], q.pdtraj.du[:] = out["Z"][b, :, :ps.n], out["Z"][b, :, ps.n:], out["L"][b]
            q.conlam, q.conmu = out["conlam"][b], out["conmu"][b]
            q.stats = Statistics(); q.stats.record_history(hist[b, :min(int(count[b]), max_rec)])
            q.stats.record(out["stats"][b])              # + the full-game record at the returned iterate
            q.status = _capi.STATUS_NAMES[int(out["status"][b])]
            q.core.res[:] = res[b]
    finally:
        if B > 1:
            batch.close()
    return None


# ---- stand-alone exported functions of the reference, single-problem form ----------------------------------------
def residual(prob: GameProblem):
    """residual!(prob): fills prob.core.res (reference row order)."""
    b = prob._push()
    res, _ = b.residual()
    prob.core.res[:] = res[0]
    return prob.core.res


def residual_jacobian(prob: GameProblem, reg: float = 0.0):
    """residual_jacobian!(prob) + regularize_residual_jacobian!(prob): dense S×S prob.core.jac."""
    b = prob._push()
    prob.core.jac = b.residual_jacobian_dense(reg, reg)[0]
    return prob.core.jac


def kkt_solve(prob: GameProblem, reg: float = 0.0):
    """Δtraj = −(lu(jac) \\ res) (solver_methods.jl:87), reference column order."""
    return prob._push().kkt_solve(reg, reg)[0]


def line_search(prob: GameProblem, reg: float = 0.0):
    b = prob._push()
    b.kkt_solve(reg, reg)
    a, j = b.line_search(prob.opts, reg)
    return float(a[0]), int(j[0])


def update_traj(prob: GameProblem, alpha: float, reg: float = 0.0):
    b = prob._push()
    b.kkt_solve(reg, reg)
    d = b.update_traj(alpha)
    prob._pull(b)
    return float(d[0])


def rollout(prob: GameProblem):
    b = prob._push(); b.rollout(); prob._pull(b)


def dual_update(prob: GameProblem):
    b = prob._push(); b.dual_update(prob.opts); prob._pull(b)


def penalty_update(prob: GameProblem):
    b = prob._push(); b.penalty_update(prob.opts); prob._pull(b)


def reset(prob: GameProblem):
    b = prob._push(); b.reset(prob.opts); prob._pull(b)


def evaluate(prob: GameProblem):
    return prob._push().evaluate()[0]


class NullSpace:
    """NullSpace (active_set_core.jl:5-27): `mat` len(hmask) x dim, `vec` its columns scattered to Sh entries and scaled to unit
    mean absolute value, `dtraj` / `dlam` their first S / remaining entries (the reference's Δtraj, Δλ)."""

    def __init__(self, sh):
        self.mat = np.zeros((sh, 0)); self.vec = []; self.dtraj = []; self.dlam = []


class ActiveSetCore:
    """ActiveSetCore(probsize) (active_set_core.jl:57-160): the Newton system bordered by the collision rows (pairs i < j) and
    columns (ordered pairs); `vmask` / `hmask` are 0-based index arrays here."""

    def __init__(self, probsize: "ProblemSize"):
        self.probsize = probsize
        N, p, S = probsize.N, probsize.p, probsize.S
        self.Sv, self.Sh = S + p * (p - 1) * (N - 1) // 2, S + p * (p - 1) * (N - 1)
        self.res = np.zeros(self.Sv); self.jac = np.zeros((self.Sv, self.Sh))
        self.vmask = np.arange(self.Sv); self.hmask = np.arange(self.Sh)
        self.null = NullSpace(self.Sh)


def active_set_residual(ascore: ActiveSetCore, prob: "GameProblem"):
    """residual!(ascore, prob, pdtraj) (active_set_methods.jl:96-124)."""
    ascore.res[:] = prob._push().active_set_residual()[0]
    return ascore.res


def active_set_residual_jacobian(ascore: ActiveSetCore, prob: "GameProblem"):
    """residual_jacobian!(ascore, prob, pdtraj) (active_set_methods.jl:131-170)."""
    ascore.jac[:] = prob._push().active_set_jacobian()[0]
    return ascore.jac


def active_masks(ascore: ActiveSetCore, prob: "GameProblem", tol: Optional[float] = None):
    """active_vertical_mask! + active_horizontal_mask! (active_set_methods.jl:28-74) at the problem's iterate and multipliers."""
    vm, hm = prob._push().active_set_masks(prob.opts.active_set_tolerance if tol is None else tol)
    ascore.vmask, ascore.hmask = np.flatnonzero(vm[0]), np.flatnonzero(hm[0])
    return ascore.vmask, ascore.hmask


def update_nullspace(ascore: ActiveSetCore, prob: "GameProblem", atol: float = 1e-20, tol: Optional[float] = None):
    """update_nullspace!(ascore, prob, pdtraj; atol) (active_set_methods.jl:173-184)."""
    tol = prob.opts.active_set_tolerance if tol is None else tol
    b = prob._push()
    vm, hm = b.active_set_masks(tol)
    ascore.vmask, ascore.hmask = np.flatnonzero(vm[0]), np.flatnonzero(hm[0])
    basis = b.update_nullspace(tol, atol)[0]
    # add_matrix! (active_set_core.jl:29-52): mat keeps the masked rows, vec the scattered vectors scaled to unit mean |.|
    ascore.null.mat = basis[:, ascore.hmask].T.copy()
    ascore.null.vec = [basis[d] / np.mean(np.abs(basis[d])) for d in range(basis.shape[0])]
    S = prob.probsize.S
    ascore.null.dtraj = [v[:S] for v in ascore.null.vec]; ascore.null.dlam = [v[S:] for v in ascore.null.vec]
    return ascore.null


def active_set(prob: GameProblem, tol: Optional[float] = None):
    return prob._push().active_set(prob.opts.active_set_tolerance if tol is None else tol)[0]
