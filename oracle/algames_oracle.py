"""CPU oracle: NumPy float64 restatement of Algames.jl's newton_solve! hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package imports this module; only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may execute it, and there only as the checker / the timed CPU baseline.

Parity status: the reference is Julia and neither Julia nor its third-party packages
(Altro 0.3.0, TrajectoryOptimization 0.4.1, RobotDynamics 0.3.1, UMFPACK) exist in the
build image, so the reference itself cannot be executed here.  This restatement is pinned
against every known-answer vector the reference's own test-suite holds for the path
(tests/test_oracle_golden.py cites each one).  Arithmetic that lives in the third-party
packages and is pinned by NO reference test is marked "parity unpinned" below:
  * RK2 (explicit midpoint) / RK3 tableaux of RobotDynamics 0.3.1,
  * CollisionConstraint / CircleConstraint formulas of TrajectoryOptimization 0.4.1,
  * MRP rotation matrix / kinematics of Rotations.jl used by QuadrotorGame (an oracle-only model this round).

Every function cites the reference file:line it follows (paths relative to the
reference root).  Indices are 0-based here, 1-based in the reference.
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

__all__ = [
    "DoubleIntegratorGame", "UnicycleGame", "BicycleGame", "ProblemSize", "Options",
    "GameObjective", "GameConstraintValues", "GameProblem", "Wall", "newton_solve",
    "residual", "residual_jacobian", "inner_iteration", "line_search", "problem_from_spec",
]


# --------------------------------------------------------------------------------------
# Models  (src/dynamics/*.jl)
# --------------------------------------------------------------------------------------
class _GameModel:
    """Joint multi-player model, component-major state layout.

    Index sets follow src/dynamics/double_integrator.jl:18-20 (same in unicycle.jl:18-20,
    bicycle.jl:19-21): pu[i] = [i, i+p, ...], px[i] = [i, i+p], pz[i] = [i, i+p, ...].
    """

    name = "abstract"

    def __init__(self, p: int, ni: int, mi: int):
        self.p = p
        self.n = ni * p
        self.m = mi * p
        self.ni = [ni] * p
        self.mi = [mi] * p
        self.pu = [np.array([i + j * p for j in range(mi)]) for i in range(p)]
        self.px = [np.array([i + j * p for j in range(2)]) for i in range(p)]
        self.pz = [np.array([i + j * p for j in range(ni)]) for i in range(p)]

    def f(self, x, u):
        raise NotImplementedError

    def fx(self, x, u):
        raise NotImplementedError

    def fu(self, x, u):
        raise NotImplementedError


class DoubleIntegratorGame(_GameModel):
    """src/dynamics/double_integrator.jl:13-31: xdot = [x[m+1:n]; u]."""

    name = "double_integrator"

    def __init__(self, p: int = 2, d: int = 2):
        super().__init__(p, 2 * d, d)
        self.d = d

    def f(self, x, u):
        return np.concatenate([x[self.m:], u])

    def fx(self, x, u):
        F = np.zeros((self.n, self.n))
        F[: self.m, self.m:] = np.eye(self.m)
        return F

    def fu(self, x, u):
        F = np.zeros((self.n, self.m))
        F[self.m:, :] = np.eye(self.m)
        return F


class UnicycleGame(_GameModel):
    """src/dynamics/unicycle.jl:14-32: x=[x..,y..,θ..,v..], u=[ω..,a..]."""

    name = "unicycle"

    def __init__(self, p: int = 2):
        super().__init__(p, 4, 2)

    def f(self, x, u):
        p = self.p
        th, v = x[2 * p:3 * p], x[3 * p:4 * p]
        return np.concatenate([np.cos(th) * v, np.sin(th) * v, u])

    def fx(self, x, u):
        p = self.p
        F = np.zeros((self.n, self.n))
        for i in range(p):
            th, v = x[2 * p + i], x[3 * p + i]
            F[i, 2 * p + i] = -math.sin(th) * v
            F[i, 3 * p + i] = math.cos(th)
            F[p + i, 2 * p + i] = math.cos(th) * v
            F[p + i, 3 * p + i] = math.sin(th)
        return F

    def fu(self, x, u):
        F = np.zeros((self.n, self.m))
        F[2 * self.p:, :] = np.eye(self.m)
        return F


class BicycleGame(_GameModel):
    """src/dynamics/bicycle.jl:15-41: x=[x..,y..,v..,ψ..], u=[a..,δ..] (code uses sin for ẏ, :37)."""

    name = "bicycle"

    def __init__(self, p: int = 2, lf: float = 0.05, lr: float = 0.05):
        super().__init__(p, 4, 2)
        self.lf, self.lr = lf, lr

    def f(self, x, u):
        p, lr, L = self.p, self.lr, self.lr + self.lf
        v, psi = x[2 * p:3 * p], x[3 * p:4 * p]
        a, dl = u[:p], u[p:2 * p]
        beta = np.arctan2(lr * np.tan(dl), L)
        return np.concatenate([v * np.cos(beta + psi), v * np.sin(beta + psi), a,
                               v * np.sin(beta) / lr])

    def _beta(self, dl):
        lr, L = self.lr, self.lr + self.lf
        t = math.tan(dl)
        beta = math.atan2(lr * t, L)
        dbeta = lr * L * (1.0 + t * t) / (L * L + lr * lr * t * t)
        return beta, dbeta

    def fx(self, x, u):
        p, lr = self.p, self.lr
        F = np.zeros((self.n, self.n))
        for i in range(p):
            v, psi = x[2 * p + i], x[3 * p + i]
            beta, _ = self._beta(u[p + i])
            F[i, 2 * p + i] = math.cos(beta + psi)
            F[i, 3 * p + i] = -v * math.sin(beta + psi)
            F[p + i, 2 * p + i] = math.sin(beta + psi)
            F[p + i, 3 * p + i] = v * math.cos(beta + psi)
            F[3 * p + i, 2 * p + i] = math.sin(beta) / lr
        return F

    def fu(self, x, u):
        p, lr = self.p, self.lr
        F = np.zeros((self.n, self.m))
        for i in range(p):
            v, psi = x[2 * p + i], x[3 * p + i]
            beta, db = self._beta(u[p + i])
            F[i, p + i] = -v * math.sin(beta + psi) * db
            F[p + i, p + i] = v * math.cos(beta + psi) * db
            F[2 * p + i, i] = 1.0
            F[3 * p + i, p + i] = v * math.cos(beta) * db / lr
        return F


class QuadrotorGame(_GameModel):
    """src/dynamics/quadrotor.jl:2-208 — ORACLE ONLY this round (the device kernels are specialised for
    4-state / 2-control players and `agb_create` has no model id for it).

    Per player 12 states [r(3), q(3) = MRP attitude, v(3), ω(3)] and 4 rotor commands, joint layout component-major
    (:31-33, pinned by test/dynamics/quadrotor.jl:4-23).  forces / moments / dynamics follow :48-108.  The MRP rotation
    matrix and kinematics come from Rotations.jl (third party, absent from /root/reference, exercised by no reference
    test): restated from the package's published formulas — PARITY UNPINNED.
    Jacobians: the reference differentiates with ForwardDiff (exact); here complex-step differentiation (exact to
    round-off for this polynomial/rational right-hand side), with max(0, ·) following the real part.
    """

    name = "quadrotor"

    def __init__(self, p: int = 2, mass: float = 0.5):
        assert p <= 4                                                   # :21
        super().__init__(p, 12, 4)
        self.mass = mass
        self.J = np.array([0.0023, 0.0023, 0.004])                      # :22
        self.gravity = np.array([0.0, 0.0, -9.81])                      # :24
        self.motor_dist, self.kf, self.km = 0.1750, 1.245, 1.0          # :25-29

    @staticmethod
    def _mrp_rotation(q):
        """RotMatrix(MRP(q)) [3P Rotations.jl]: I + (8 S² + 4 (1 − |q|²) S) / (1 + |q|²)², S = skew(q)."""
        S = np.array([[0, -q[2], q[1]], [q[2], 0, -q[0]], [-q[1], q[0], 0]], dtype=q.dtype)
        n2 = q @ q
        return np.eye(3, dtype=q.dtype) + (8.0 * (S @ S) + 4.0 * (1.0 - n2) * S) / (1.0 + n2) ** 2

    @staticmethod
    def _mrp_kinematics(q, w):
        """Rotations.kinematics(MRP(q), ω) [3P Rotations.jl]: q̇ = ¼ ((1 − |q|²) I + 2 S + 2 q qᵀ) ω (body rates)."""
        S = np.array([[0, -q[2], q[1]], [q[2], 0, -q[0]], [-q[1], q[0], 0]], dtype=q.dtype)
        A = (1.0 - q @ q) * np.eye(3, dtype=q.dtype) + 2.0 * S + 2.0 * np.outer(q, q)
        return 0.25 * (A @ w)

    def _rotor_forces(self, u, i):
        w = np.array([u[j * self.p + i] for j in range(4)])
        F = np.where((self.kf * w).real > 0, self.kf * w, 0.0 * w)      # max(0, kf·w), :58-61
        return w, F

    def forces(self, x, u, i):                                          # :48-69
        P = self.p
        q = np.array([x[3 * P + i], x[4 * P + i], x[5 * P + i]])
        _, F = self._rotor_forces(u, i)
        body = np.array([0.0 * F[0], 0.0 * F[0], F.sum()])
        return self.mass * self.gravity + self._mrp_rotation(q) @ body

    def moments(self, x, u, i):                                         # :71-92
        w, F = self._rotor_forces(u, i)
        L = self.motor_dist
        return np.array([L * (F[1] - F[3]), L * (F[2] - F[0]), self.km * (w[0] - w[1] + w[2] - w[3])])

    def player_dynamics(self, x, u, i):                                 # :100-119
        P = self.p
        q = np.array([x[3 * P + i], x[4 * P + i], x[5 * P + i]])
        v = np.array([x[6 * P + i], x[7 * P + i], x[8 * P + i]])
        om = np.array([x[9 * P + i], x[10 * P + i], x[11 * P + i]])
        F, tau = self.forces(x, u, i), self.moments(x, u, i)
        return v, self._mrp_kinematics(q, om), F / self.mass, (tau - np.cross(om, self.J * om)) / self.J

    def f(self, x, u):                                                  # :121-205 (component-major interleave)
        x, u = np.asarray(x), np.asarray(u)
        out = np.zeros(self.n, dtype=np.result_type(x.dtype, u.dtype, float))
        for i in range(self.p):
            parts = np.concatenate(self.player_dynamics(x, u, i))
            out[[i + c * self.p for c in range(12)]] = parts
        return out

    def _cstep(self, x, u, wrt):
        h = 1e-30
        base = np.asarray(x if wrt == "x" else u, float)
        cols = []
        for a in range(len(base)):
            z = base.astype(complex); z[a] += 1j * h
            cols.append(self.f(z, np.asarray(u, float)).imag / h if wrt == "x" else self.f(np.asarray(x, float), z).imag / h)
        return np.array(cols).T

    def fx(self, x, u):
        return self._cstep(x, u, "x")

    def fu(self, x, u):
        return self._cstep(x, u, "u")


def make_model(name: str, p: int, d: int = 2, lf: float = 0.05, lr: float = 0.05):
    if name == "quadrotor":
        return QuadrotorGame(p=p)
    if name == "double_integrator":
        return DoubleIntegratorGame(p=p, d=d)
    if name == "unicycle":
        return UnicycleGame(p=p)
    if name == "bicycle":
        return BicycleGame(p=p, lf=lf, lr=lr)
    raise ValueError(name)


# --------------------------------------------------------------------------------------
# Integrators  [3P RobotDynamics 0.3.1 — tableaux parity unpinned; SURVEY §8 a8/a16]
# --------------------------------------------------------------------------------------
def rk2(model, x, u, dt):
    """discrete_dynamics(RK2,…) as called at src/problem/local_quantities.jl:13 (explicit midpoint)."""
    k1 = model.f(x, u) * dt
    k2 = model.f(x + k1 / 2, u) * dt
    return x + k2


def rk2_jacobian(model, x, u, dt):
    """[A|B] of rk2 — analytic chain rule replacing ForwardDiff (local_quantities.jl:26)."""
    xm = x + (dt / 2) * model.f(x, u)
    Fxm = model.fx(xm, u)
    A = np.eye(model.n) + dt * Fxm @ (np.eye(model.n) + (dt / 2) * model.fx(x, u))
    B = dt * (Fxm @ ((dt / 2) * model.fu(x, u)) + model.fu(xm, u))
    return A, B


def rk3(model, x, u, dt):
    """discrete_dynamics(RK3,…) used by rollout! at src/problem/solver_methods.jl:17-18."""
    k1 = model.f(x, u) * dt
    k2 = model.f(x + k1 / 2, u) * dt
    k3 = model.f(x - k1 + 2 * k2, u) * dt
    return x + (k1 + 4 * k2 + k3) / 6


# --------------------------------------------------------------------------------------
# ProblemSize / index maps  (src/struct/problem_size.jl, src/core/newton_core.jl)
# --------------------------------------------------------------------------------------
class ProblemSize:
    """src/struct/problem_size.jl:5-35."""

    def __init__(self, N: int, model):
        self.N, self.n, self.m, self.p = N, model.n, model.m, model.p
        self.ni, self.mi = model.ni, model.mi
        self.pu, self.px, self.pz = model.pu, model.px, model.pz
        self.S = self.n * self.p * (N - 1) + self.m * (N - 1) + self.n * (N - 1)  # :22

    def __eq__(self, other):   # :37-46, field by field
        if not isinstance(other, ProblemSize):
            return NotImplemented
        f = lambda v: [list(e) if hasattr(e, "__len__") else e for e in v] if hasattr(v, "__len__") else v
        return all(f(getattr(self, k)) == f(getattr(other, k)) for k in ("N", "n", "m", "p", "ni", "mi", "pu", "px", "pz", "S"))

    __hash__ = None


def valid_v(prob, i0, n1, i1, v1, N, p):
    """src/core/stamp.jl:199-214 (VStamp validity); players/knots 1-based like the reference."""
    if prob == "opt" and 1 <= i0 <= p:
        if n1 == "u" and i1 == i0 and 1 <= v1 <= N - 1:
            return True
        if n1 == "x" and i1 == 1 and 2 <= v1 <= N:
            return True
    if prob == "dyn" and i0 == 1:
        if n1 == "x" and i1 == 1 and 1 <= v1 <= N - 1:
            return True
    return False


def valid_h(n2, i2, v2, N, p):
    """src/core/stamp.jl:220-229 (HStamp validity)."""
    if n2 == "u" and 1 <= i2 <= p and 1 <= v2 <= N - 1:
        return True
    if n2 == "λ" and 1 <= i2 <= p and 1 <= v2 <= N - 1:
        return True
    if n2 == "x" and i2 == 1 and 2 <= v2 <= N:
        return True
    return False


def valid(prob, i0, n1, i1, v1, n2, i2, v2, N, p):
    """src/core/stamp.jl:171-192 (full Stamp validity)."""
    b1 = valid_v(prob, i0, n1, i1, v1, N, p)
    b2 = False
    if prob == "opt" and 1 <= i0 <= p:
        if n2 == "u" and 1 <= i2 <= p and 1 <= v2 <= N - 1:
            b2 = True
        elif n2 == "λ" and i2 == i0 and 1 <= v2 <= N - 1:
            b2 = True
        elif n2 == "x" and i2 == 1 and 2 <= v2 <= N:
            b2 = True
    if prob == "dyn" and i0 == 1:
        if n2 == "u" and 1 <= i2 <= p and 1 <= v2 <= N - 1:
            b2 = True
        elif n2 == "x" and i2 == 1 and 2 <= v2 <= N:
            b2 = True
    return b1 and b2


class NewtonCore:
    """Index algebra of src/core/newton_core.jl:40-89 with closed-form 0-based offsets.

    vert[("opt",i,"x",k)]  rows of player i's stationarity w.r.t. x_k   (k = 2..N, 1-based knots)
    vert[("opt",i,"u",k)]  rows of player i's stationarity w.r.t. u_{i,k} (k = 1..N-1)
    vert[("dyn",k)]        dynamics rows of stage k                       (k = 1..N-1)
    horiz[("x",k)], horiz[("u",i,k)], horiz[("λ",i,k)]  column blocks.
    """

    def __init__(self, ps: ProblemSize):
        self.ps = ps
        N, n, p, mi = ps.N, ps.n, ps.p, ps.mi
        self.vert, self.horiz = {}, {}
        off = 0
        for i in range(1, p + 1):                      # newton_core.jl:46-55
            for k in range(1, N):
                self.vert[("opt", i, "x", k + 1)] = np.arange(off, off + n); off += n
                self.vert[("opt", i, "u", k)] = np.arange(off, off + mi[i - 1]); off += mi[i - 1]
        for k in range(1, N):                          # :56-60
            self.vert[("dyn", k)] = np.arange(off, off + n); off += n
        assert off == ps.S
        off = 0
        for k in range(1, N):                          # :73-87
            self.horiz[("x", k + 1)] = np.arange(off, off + n); off += n
            for i in range(1, p + 1):
                self.horiz[("u", i, k)] = np.arange(off, off + mi[i - 1]); off += mi[i - 1]
            for i in range(1, p + 1):
                self.horiz[("λ", i, k)] = np.arange(off, off + n); off += n
        assert off == ps.S
        self.res = np.zeros(ps.S)
        self.jac = None


# --------------------------------------------------------------------------------------
# Options / Regularizer  (src/struct/options.jl:5-116, src/struct/regularizer.jl)
# --------------------------------------------------------------------------------------
@dataclass
class Regularizer:
    x: float = 1e-3
    u: float = 1e-3
    lam: float = 1e-3

    def set(self, v):          # regularizer.jl:25-29
        self.x = self.u = self.lam = v

    def mult(self, v):         # regularizer.jl:31-35
        self.x, self.u, self.lam = self.x * v, self.u * v, self.lam * v


@dataclass
class Options:
    amplitude_init: float = 1e-8
    shift: int = 2 ** 10
    regularize: bool = True
    reg: Regularizer = field(default_factory=Regularizer)
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


# --------------------------------------------------------------------------------------
# Constraints  (src/constraints/*.jl + [3P] TrajectoryOptimization / Altro)
# --------------------------------------------------------------------------------------
class _Con:
    """A vector-valued inequality constraint c(z) <= 0 with dense Jacobian w.r.t. x or u."""
    kind = "state"

    def length(self):
        raise NotImplementedError

    def evaluate(self, x, u):
        raise NotImplementedError

    def jacobian(self, x, u):
        raise NotImplementedError


class CollisionConstraint(_Con):
    """[3P TrajectoryOptimization 0.4.1, parity unpinned] c = r² − ‖x[x1] − x[x2]‖².
    Built by add_collision_avoidance! (src/constraints/constraints_methods.jl:5-19)."""

    def __init__(self, n, x1, x2, radius):
        self.n, self.x1, self.x2, self.radius = n, np.asarray(x1), np.asarray(x2), float(radius)

    def length(self):
        return 1

    def evaluate(self, x, u):
        d = x[self.x1] - x[self.x2]
        return np.array([self.radius ** 2 - d @ d])

    def jacobian(self, x, u):
        d = x[self.x1] - x[self.x2]
        J = np.zeros((1, self.n))
        J[0, self.x1] = -2 * d
        J[0, self.x2] = 2 * d
        return J


class CircleConstraint(_Con):
    """[3P TrajectoryOptimization 0.4.1, parity unpinned] c_l = r_l² − (x−xc_l)² − (y−yc_l)².
    Built by add_circle_constraint! (constraints_methods.jl:121-139)."""

    def __init__(self, n, xc, yc, radius, xi, yi):
        self.n = n
        self.xc, self.yc, self.r = (np.asarray(v, float) for v in (xc, yc, radius))
        self.xi, self.yi = xi, yi

    def length(self):
        return len(self.xc)

    def evaluate(self, x, u):
        return self.r ** 2 - (x[self.xi] - self.xc) ** 2 - (x[self.yi] - self.yc) ** 2

    def jacobian(self, x, u):
        J = np.zeros((len(self.xc), self.n))
        J[:, self.xi] = -2 * (x[self.xi] - self.xc)
        J[:, self.yi] = -2 * (x[self.yi] - self.yc)
        return J


class ControlBoundConstraint(_Con):
    """src/constraints/control_bound_constraint.jl:94-106: [u−u_max; u_min−u][finite]."""
    kind = "control"

    def __init__(self, m, u_max, u_min):
        self.m = m
        self.u_max = np.broadcast_to(np.asarray(u_max, float), (m,)).copy()
        self.u_min = np.broadcast_to(np.asarray(u_min, float), (m,)).copy()
        if not np.all(self.u_max >= self.u_min):
            raise ValueError("Upper bounds must be greater than or equal to lower bounds")  # :62-68
        b = np.concatenate([-self.u_max, self.u_min])
        self.inds = np.flatnonzero(np.isfinite(b))                                      # :33-35

    def length(self):
        return len(self.inds)

    def evaluate(self, x, u):
        return np.concatenate([u - self.u_max, self.u_min - u])[self.inds]

    def jacobian(self, x, u):
        J = np.vstack([np.eye(self.m), -np.eye(self.m)])
        return J[self.inds]


class StateBoundConstraint(_Con):
    """src/constraints/state_bound_constraint.jl:80-92 (same shape as the control bound, on x)."""

    def __init__(self, n, x_max, x_min):
        self.n = n
        self.x_max = np.broadcast_to(np.asarray(x_max, float), (n,)).copy()
        self.x_min = np.broadcast_to(np.asarray(x_min, float), (n,)).copy()
        if not np.all(self.x_max >= self.x_min):
            raise ValueError("Upper bounds must be greater than or equal to lower bounds")
        b = np.concatenate([-self.x_max, self.x_min])
        self.inds = np.flatnonzero(np.isfinite(b))

    def length(self):
        return len(self.inds)

    def evaluate(self, x, u):
        return np.concatenate([x - self.x_max, self.x_min - x])[self.inds]

    def jacobian(self, x, u):
        J = np.vstack([np.eye(self.n), -np.eye(self.n)])
        return J[self.inds]


@dataclass
class Wall:
    """src/constraints/constraints_methods.jl:155-159."""
    p1: np.ndarray
    p2: np.ndarray
    v: np.ndarray


class WallConstraint(_Con):
    """src/constraints/wall_constraint.jl:56-89: ((x−p1)·v)·1[left]·1[right]."""

    def __init__(self, n, x1, y1, x2, y2, xv, yv, x=0, y=1):
        self.n = n
        self.x1, self.y1, self.x2, self.y2, self.xv, self.yv = (
            np.asarray(v, float) for v in (x1, y1, x2, y2, xv, yv))
        self.x, self.y = x, y

    def length(self):
        return len(self.x1)

    def _mask(self, X):
        x, y = X[self.x], X[self.y]
        left = (x - self.x1) * (self.x2 - self.x1) + (y - self.y1) * (self.y2 - self.y1) > 0
        right = (x - self.x2) * (self.x1 - self.x2) + (y - self.y2) * (self.y1 - self.y2) > 0
        return x, y, (left & right).astype(float)

    def evaluate(self, X, u):
        x, y, msk = self._mask(X)
        return ((x - self.x1) * self.xv + (y - self.y1) * self.yv) * msk

    def jacobian(self, X, u):
        _, _, msk = self._mask(X)
        J = np.zeros((len(self.x1), self.n))
        J[:, self.x] = msk * self.xv
        J[:, self.y] = msk * self.yv
        return J


@dataclass
class Wall3D:
    """src/constraints/constraints_methods.jl:201-206: corner points p1, p2, p3 of the rectangle, outward normal v."""
    p1: np.ndarray
    p2: np.ndarray
    p3: np.ndarray
    v: np.ndarray


class Wall3DConstraint(_Con):
    """src/constraints/wall_constraint.jl:141-233 (ORACLE ONLY): ((X−p1)·v) · 1[left]·1[right]·1[bottom]·1[top];
    pinned by test/constraints/wall_constraint.jl:33-70."""

    def __init__(self, n, x1, y1, z1, x2, y2, z2, x3, y3, z3, xv, yv, zv, x=0, y=1, z=2):
        self.n = n
        (self.x1, self.y1, self.z1, self.x2, self.y2, self.z2, self.x3, self.y3, self.z3, self.xv, self.yv, self.zv) = (
            np.asarray(v, float) for v in (x1, y1, z1, x2, y2, z2, x3, y3, z3, xv, yv, zv))
        self.x, self.y, self.z = x, y, z

    def length(self):
        return len(self.x1)

    def _mask(self, X):
        x, y, z = X[self.x], X[self.y], X[self.z]
        dot = lambda ax, ay, az, bx, by, bz, cx, cy, cz: (x - ax) * (bx - cx) + (y - ay) * (by - cy) + (z - az) * (bz - cz)
        left = dot(self.x1, self.y1, self.z1, self.x2, self.y2, self.z2, self.x1, self.y1, self.z1) > 0       # :203
        right = dot(self.x2, self.y2, self.z2, self.x1, self.y1, self.z1, self.x2, self.y2, self.z2) > 0      # :204
        bottom = dot(self.x3, self.y3, self.z3, self.x2, self.y2, self.z2, self.x3, self.y3, self.z3) > 0     # :205
        top = dot(self.x2, self.y2, self.z2, self.x3, self.y3, self.z3, self.x2, self.y2, self.z2) > 0        # :206
        return x, y, z, (left & right & bottom & top).astype(float)

    def evaluate(self, X, u):
        x, y, z, msk = self._mask(X)
        return ((x - self.x1) * self.xv + (y - self.y1) * self.yv + (z - self.z1) * self.zv) * msk           # :207-208

    def jacobian(self, X, u):
        _, _, _, msk = self._mask(X)
        J = np.zeros((len(self.x1), self.n))
        J[:, self.x], J[:, self.y], J[:, self.z] = msk * self.xv, msk * self.yv, msk * self.zv                # :231-235
        return J


@dataclass
class CylinderWall:
    """src/constraints/constraints_methods.jl:249-254: base point p, axis v in {"x","y","z"}, length l, radius r."""
    p: np.ndarray
    v: str
    l: float
    r: float


class CylinderConstraint(_Con):
    """src/constraints/cylinder_constraint.jl:33-129 (ORACLE ONLY): r² − (squared distance to the axis), active
    only between the end caps; pinned by test/constraints/cylinder_constraint.jl:4-22."""

    def __init__(self, n, p1, p2, p3, v, l, r, x=0, y=1, z=2):
        self.n = n
        self.p1, self.p2, self.p3, self.l, self.r = (np.asarray(a, float) for a in (p1, p2, p3, l, r))
        self.v = [str(a) for a in v]
        assert all(a in ("x", "y", "z") for a in self.v)
        self.axis = np.array([[a == "x", a == "y", a == "z"] for a in self.v], dtype=float)     # one-hot axis per cylinder
        self.x, self.y, self.z = x, y, z

    def length(self):
        return len(self.p1)

    def _terms(self, X):
        t = np.stack([X[self.x] - self.p1, X[self.y] - self.p2, X[self.z] - self.p3], axis=1)  # t0 = X − p, :76-79
        along = (t * self.axis).sum(axis=1)
        valid = ((along > 0.0) & (along < self.l)).astype(float)                              # :81-84
        return t, valid

    def evaluate(self, X, u):
        t, valid = self._terms(X)
        return (self.r ** 2 - (t ** 2 * (1.0 - self.axis)).sum(axis=1)) * valid               # :86-91

    def jacobian(self, X, u):
        t, valid = self._terms(X)
        J = np.zeros((len(self.p1), self.n))
        g = -2.0 * t * (1.0 - self.axis) * valid[:, None]                                     # :122-126
        J[:, self.x], J[:, self.y], J[:, self.z] = g[:, 0], g[:, 1], g[:, 2]
        return J


class ALConVal:
    """[3P Altro 0.3.0 ALConVal] per-knot vals/jac/λ/μ/grad/hess; formulas pinned by
    test/constraints/constraint_derivatives.jl:22-34 and test/constraints/constraints_methods.jl:176-229."""

    def __init__(self, n, m, con, inds):
        self.con, self.inds = con, list(inds)      # inds: 1-based knot indices
        P, K = con.length(), len(self.inds)
        w = n if con.kind == "state" else m
        self.vals = np.zeros((K, P))
        self.jac = np.zeros((K, P, w))
        self.lam = np.zeros((K, P))
        self.mu = np.ones((K, P))
        self.grad = np.zeros((K, w))
        self.hess = np.zeros((K, w, w))
        self.c_max = np.zeros(K)
        self.phi, self.mu0, self.mu_max, self.lam_max = 10.0, 1.0, 1e8, 1e8

    def evaluate(self, X, U):
        for j, k in enumerate(self.inds):
            u = U[k - 1] if k - 1 < len(U) else None
            self.vals[j] = self.con.evaluate(X[k - 1], u)

    def jacobian(self, X, U):
        for j, k in enumerate(self.inds):
            u = U[k - 1] if k - 1 < len(U) else None
            self.jac[j] = self.con.jacobian(X[k - 1], u)

    def cost_expansion(self):
        """Altro cost_expansion!(::Inequality): Iρ = diag((c>=0)|(λ>0))·μ (constraint_derivatives test :28-34)."""
        for j in range(len(self.inds)):
            c, lam, mu, J = self.vals[j], self.lam[j], self.mu[j], self.jac[j]
            a = (c >= 0) | (lam > 0)
            Irho = a * mu
            self.grad[j] = J.T @ lam + J.T @ (Irho * c)
            self.hess[j] = J.T @ (Irho[:, None] * J)

    def max_violation(self):
        """[3P TrajOpt max_violation!] c_max = max(0, max c); pinned test/struct/violations.jl:28-35."""
        for j in range(len(self.inds)):
            self.c_max[j] = max(0.0, float(np.max(self.vals[j]))) if self.vals.shape[1] else 0.0


def velocity_index(model, i):
    """src/constraints/velocity_constraint.jl:30-43 (0-based player and index here)."""
    assert 0 <= i < model.p
    if model.name == "unicycle":
        return model.pz[i][3]
    if model.name == "bicycle":
        return model.pz[i][2]
    raise NotImplementedError("Velocity Index is not implemented for DoubleIntegratorGame.")


class GameConstraintValues:
    """src/constraints/game_constraints.jl:5-53 + adders of constraints_methods.jl."""

    def __init__(self, probsize: ProblemSize):
        self.probsize = probsize
        self.alpha_dual = 1.0
        self.alphax_dual = [1.0] * probsize.p
        self.active_set_tolerance = 0.0
        self.state_conval: List[List[ALConVal]] = [[] for _ in range(probsize.p)]
        self.control_conval: List[ALConVal] = []

    # -- adders -----------------------------------------------------------------------
    def _add_state(self, i, con):
        ps = self.probsize
        self.state_conval[i].append(ALConVal(ps.n, ps.m, con, range(2, ps.N + 1)))   # knots 2:N

    def add_collision_avoidance(self, radius, i=None, j=None):
        """constraints_methods.jl:5-39; scalar radius r ⇒ pair radius 2r, vector ⇒ r_i+r_j."""
        ps = self.probsize
        if i is not None:
            self._add_state(i, CollisionConstraint(ps.n, ps.px[i], ps.px[j], radius))
            return
        rad = np.broadcast_to(np.asarray(radius, float), (ps.p,))
        for a in range(ps.p):
            for b in range(ps.p):
                if b != a:
                    self._add_state(a, CollisionConstraint(ps.n, ps.px[a], ps.px[b], rad[a] + rad[b]))

    def add_spherical_collision_avoidance(self, radius, i=None, j=None):
        """constraints_methods.jl:45-81 (ORACLE ONLY): CollisionConstraint on the first three state components."""
        ps = self.probsize
        if i is not None:
            self._add_state(i, CollisionConstraint(ps.n, ps.pz[i][:3], ps.pz[j][:3], radius))       # :52-55
            return
        rad = np.broadcast_to(np.asarray(radius, float), (ps.p,))
        for a in range(ps.p):
            for b in range(ps.p):
                if b != a:
                    self._add_state(a, CollisionConstraint(ps.n, ps.pz[a][:3], ps.pz[b][:3], rad[a] + rad[b]))

    def add_wall3d_constraint(self, walls, i=None):
        """add_wall_constraint!(game_con, i, walls::Vector{Wall3D}) (constraints_methods.jl:208-247, ORACLE ONLY)."""
        ps = self.probsize
        for a in (range(ps.p) if i is None else [i]):
            cols = [[getattr(w, f)[c] for w in walls] for f in ("p1", "p2", "p3", "v") for c in range(3)]
            self._add_state(a, Wall3DConstraint(ps.n, *cols, ps.pz[a][0], ps.pz[a][1], ps.pz[a][2]))

    def add_cylinder_constraint(self, walls, i=None):
        """add_wall_constraint!(game_con, i, walls::Vector{CylinderWall}) (constraints_methods.jl:256-285, ORACLE ONLY)."""
        ps = self.probsize
        for a in (range(ps.p) if i is None else [i]):
            self._add_state(a, CylinderConstraint(ps.n, [w.p[0] for w in walls], [w.p[1] for w in walls], [w.p[2] for w in walls],
                                                  [w.v for w in walls], [w.l for w in walls], [w.r for w in walls],
                                                  ps.pz[a][0], ps.pz[a][1], ps.pz[a][2]))

    def add_state_bound(self, i, x_max, x_min):
        self._add_state(i, StateBoundConstraint(self.probsize.n, x_max, x_min))       # :87-98

    def add_velocity_bound(self, model, v_max, v_min, i=None):
        """src/constraints/velocity_constraint.jl:1-28: a bound on player i's speed becomes one StateBoundConstraint
        on that joint component in EVERY player's list; the vector form skips players with both bounds infinite."""
        p, n = model.p, model.n
        if i is None:
            assert len(v_max) == len(v_min) == p                                       # :4
            for a in range(p):
                if v_max[a] != np.inf or v_min[a] != -np.inf:                          # :6
                    self.add_velocity_bound(model, v_max[a], v_min[a], i=a)
            return
        assert v_max != np.inf or v_min != -np.inf                                     # :14
        x_max, x_min = np.full(n, np.inf), np.full(n, -np.inf)
        vi = velocity_index(model, i)
        x_max[vi], x_min[vi] = v_max, v_min                                            # :18-22
        for j in range(p):                                                             # :24-26
            self.add_state_bound(j, x_max, x_min)

    def add_control_bound(self, u_max, u_min):
        ps = self.probsize
        con = ControlBoundConstraint(ps.m, u_max, u_min)
        self.control_conval.append(ALConVal(ps.n, ps.m, con, range(1, ps.N)))          # knots 1:N-1, :109

    def add_circle_constraint(self, xc, yc, radius, i=None):
        ps = self.probsize
        for a in (range(ps.p) if i is None else [i]):                                  # :121-148
            self._add_state(a, CircleConstraint(ps.n, xc, yc, radius, ps.px[a][0], ps.px[a][1]))

    def add_wall_constraint(self, walls, i=None):
        ps = self.probsize
        for a in (range(ps.p) if i is None else [i]):                                  # :161-195
            con = WallConstraint(ps.n, [w.p1[0] for w in walls], [w.p1[1] for w in walls],
                                 [w.p2[0] for w in walls], [w.p2[1] for w in walls],
                                 [w.v[0] for w in walls], [w.v[1] for w in walls],
                                 ps.px[a][0], ps.px[a][1])
            self._add_state(a, con)

    # -- iteration helpers ------------------------------------------------------------
    def all_convals(self):
        for i in range(self.probsize.p):
            for cv in self.state_conval[i]:
                yield ("state", i, cv)
        for cv in self.control_conval:
            yield ("control", -1, cv)

    def set_constraint_params(self, opts: Options):
        """game_constraints.jl:33-53."""
        self.alpha_dual = opts.alpha_dual
        self.alphax_dual = list(opts.alphax_dual[: self.probsize.p])
        self.active_set_tolerance = opts.active_set_tolerance
        for _, _, cv in self.all_convals():
            cv.phi, cv.mu0, cv.mu_max, cv.lam_max = (opts.rho_increase, opts.rho_0,
                                                     opts.rho_max, opts.lambda_max)

    def reset_duals(self):                    # constraints_methods.jl:301-313
        for _, _, cv in self.all_convals():
            cv.lam[:] = 0.0

    def reset_penalties(self):                # :315-327
        for _, _, cv in self.all_convals():
            cv.mu[:] = cv.mu0

    def reset(self):                          # :295-299
        self.reset_duals()
        self.reset_penalties()

    def penalty_update(self):
        """:329-341 → Altro.penalty_update!: μ ← clamp(ϕ·μ, 0, μ_max) (pinned by constraints_methods test :176-193)."""
        for _, _, cv in self.all_convals():
            cv.mu[:] = np.clip(cv.phi * cv.mu, 0.0, cv.mu_max)

    def evaluate(self, X, U):                 # :367-379
        for _, _, cv in self.all_convals():
            cv.evaluate(X, U)

    def jacobian(self, X, U):                 # :382-394
        for _, _, cv in self.all_convals():
            cv.jacobian(X, U)

    def dual_update(self):
        """:349-365 and :421-440 (all in-scope constraints are Inequality ⇒ clamp to [0, λ_max])."""
        for kind, i, cv in self.all_convals():
            a = self.alphax_dual[i] if kind == "state" else self.alpha_dual
            cv.lam[:] = np.clip(cv.lam + a * cv.mu * cv.vals, 0.0, cv.lam_max)

    def active_set(self, tol=None):
        """Altro.update_active_set!: active = (c >= −tol) | (λ > 0); pinned test/active_set/active_set_methods.jl:16-34."""
        tol = self.active_set_tolerance if tol is None else tol
        return [((cv.vals >= -tol) | (cv.lam > 0)) for _, _, cv in self.all_convals()]


# --------------------------------------------------------------------------------------
# Objective  (src/objective/objective.jl + [3P] TrajOpt LQRCost expansion with dt scaling)
# --------------------------------------------------------------------------------------
class GameObjective:
    """src/objective/objective.jl:6-35: per-player diagonal LQR expanded to joint dims,
    plus optional soft collision costs (:84-100, :109-173)."""

    def __init__(self, Q, R, xf, uf, N, model):
        self.model, self.N = model, N
        n, m, p = model.n, model.m, model.p
        self.Q = np.zeros((p, n)); self.R = np.zeros((p, m))
        self.xf = np.zeros((p, n)); self.uf = np.zeros((p, m))
        for i in range(p):                                               # :23-28 expand_vector
            self.Q[i, model.pz[i]] = np.diag(Q[i]) if np.ndim(Q[i]) == 2 else Q[i]
            self.R[i, model.pu[i]] = np.diag(R[i]) if np.ndim(R[i]) == 2 else R[i]
            self.xf[i, model.pz[i]] = xf[i]
            self.uf[i, model.pu[i]] = uf[i]
        self.collision = [[] for _ in range(p)]     # per player: list of (mu, r, pxi, pxj)

    def add_collision_cost(self, radius, mu):
        """:84-100: for every ordered pair (i,j≠i) a CollisionCost(μ_i, r_i, px_i, px_j) on all N knots."""
        model = self.model
        for i in range(model.p):
            for j in range(model.p):
                if j != i:
                    self.collision[i].append((float(mu[i]), float(radius[i]), model.px[i], model.px[j]))

    # dt scaling [3P cost_gradient!/cost_hessian!, pinned test/objective/objective.jl:52-64]:
    # stage knots ×dt on q,r,Q,R; terminal knot q,Q ×1 and r,R ×0.
    def gradient(self, i, k, x, u, dt):
        """(q, r) of player i at 1-based knot k."""
        N = self.N
        dtx = dt if k < N else 1.0
        dtu = dt if k < N else 0.0
        q = self.Q[i] * (x - self.xf[i])
        r = self.R[i] * (u - self.uf[i]) if k < N else np.zeros(self.model.m)
        for (mu, rad, pxi, pxj) in self.collision[i]:
            q = q + _collision_cost_grad(mu, rad, pxi, pxj, x)
        return q * dtx, r * dtu

    def hessian(self, i, k, x, u, dt):
        N = self.N
        dtx = dt if k < N else 1.0
        dtu = dt if k < N else 0.0
        Qh = np.diag(self.Q[i]).copy()
        Rh = np.diag(self.R[i]).copy() if k < N else np.zeros((self.model.m, self.model.m))
        for (mu, rad, pxi, pxj) in self.collision[i]:
            Qh += _collision_cost_hess(mu, rad, pxi, pxj, x)
        return Qh * dtx, Rh * dtu


def collision_stage_cost(mu, r, pxi, pxj, x):
    """objective.jl:127-131: ½ μ max(0, r − ‖xi−xj‖)² (pinned 0.05 at test/objective/objective.jl:141)."""
    return 0.5 * mu * max(0.0, r - np.linalg.norm(x[pxi] - x[pxj])) ** 2


def _collision_cost_grad(mu, r, pxi, pxj, x):
    """objective.jl:134-149 (note the ε-regularised direction, kept verbatim)."""
    n = len(x)
    eps = 1e-10
    eps_norm = eps * math.sqrt(n)
    d = x[pxi] - x[pxj]
    dn = np.linalg.norm(d)
    q = np.zeros(n)
    if max(0.0, r - dn) > 0.0:
        g = mu * (r * (eps + d) / (eps_norm + dn) - d)
        q[pxi] = -g
        q[pxj] = g
    return q


def _collision_cost_hess(mu, r, pxi, pxj, x):
    """objective.jl:157-173."""
    n = len(x)
    d = x[pxi] - x[pxj]
    dn = np.linalg.norm(d)
    Qh = np.zeros((n, n))
    if max(0.0, r - dn) > 0.0:
        blk = mu * (np.eye(len(d)) - r * np.eye(len(d)) / dn + r * np.outer(d, d) / dn ** 3)
        Qh[np.ix_(pxi, pxi)] = blk
        Qh[np.ix_(pxi, pxj)] = -blk
        Qh[np.ix_(pxj, pxi)] = -blk
        Qh[np.ix_(pxj, pxj)] = blk
    return Qh


# --------------------------------------------------------------------------------------
# PrimalDualTraj  (src/struct/primal_dual_traj.jl)
# --------------------------------------------------------------------------------------
class PrimalDualTraj:
    """X[N,n], U[N,m] (row N-1 of U is the unused terminal control), du[p,N-1,n]."""

    def __init__(self, ps: ProblemSize, dt, f=None, amplitude=1e-8, rng=None):
        self.ps, self.dt = ps, dt
        rng = rng or np.random.default_rng(0)
        f = f or (lambda *s: rng.random(s))
        self.du = amplitude * f(ps.p, ps.N - 1, ps.n)               # :16-17
        z = amplitude * f(ps.N, ps.n + ps.m)                        # :18-20
        self.X = z[:, : ps.n].copy()
        self.U = z[:, ps.n:].copy()

    def copy(self):
        return copy.deepcopy(self)


def init_traj(pd: PrimalDualTraj, x0, f, amplitude, s=2 ** 10):
    """primal_dual_traj.jl:29-44 (in place, ascending k, so shifted reads see old entries)."""
    ps = pd.ps
    N = ps.N
    for k in range(1, N + 1):
        if k + s <= N:
            pd.X[k - 1] = pd.X[k + s - 1]
            pd.U[k - 1] = pd.U[k + s - 1]
        else:
            z = amplitude * f(ps.n + ps.m)
            pd.X[k - 1], pd.U[k - 1] = z[: ps.n], z[ps.n:]
    for i in range(ps.p):
        for k in range(1, N):
            pd.du[i, k - 1] = pd.du[i, k + s - 1] if k + s <= N - 1 else amplitude * f(ps.n)
    pd.X[0] = x0


def update_traj(target, source, alpha, delta):
    """primal_dual_traj.jl:109-128: x_{2..N}, u_{1..N-1}, all λ."""
    N = target.ps.N
    target.X[1:] = source.X[1:] + alpha * delta.X[1:]
    target.U[: N - 1] = source.U[: N - 1] + alpha * delta.U[: N - 1]
    target.du[:] = source.du + alpha * delta.du


def delta_step(dpd, alpha):
    """primal_dual_traj.jl:130-147 (duals excluded)."""
    ps = dpd.ps
    s = np.abs(dpd.X[1:]).sum() + np.abs(dpd.U[: ps.N - 1]).sum()
    return s * alpha / ((ps.N - 1) * (ps.n + ps.m))


def set_traj(core: NewtonCore, dpd, dtraj):
    """primal_dual_traj.jl:46-75: scatter the solution vector (column order) into knots."""
    ps = core.ps
    for k in range(1, ps.N):
        dpd.X[k] = dtraj[core.horiz[("x", k + 1)]]
        for i in range(1, ps.p + 1):
            dpd.U[k - 1, ps.pu[i - 1]] = dtraj[core.horiz[("u", i, k)]]
        for i in range(1, ps.p + 1):
            dpd.du[i - 1, k - 1] = dtraj[core.horiz[("λ", i, k)]]


def get_traj(core: NewtonCore, dpd):
    """primal_dual_traj.jl:78-107."""
    ps = core.ps
    out = np.zeros(ps.S)
    for k in range(1, ps.N):
        out[core.horiz[("x", k + 1)]] = dpd.X[k]
        for i in range(1, ps.p + 1):
            out[core.horiz[("u", i, k)]] = dpd.U[k - 1, ps.pu[i - 1]]
            out[core.horiz[("λ", i, k)]] = dpd.du[i - 1, k - 1]
    return out


def rollout_rk3(model, pd):
    """rollout!(RK3, model, traj) (solver_methods.jl:17-18) [3P, parity unpinned]."""
    for k in range(pd.ps.N - 1):
        pd.X[k + 1] = rk3(model, pd.X[k], pd.U[k], pd.dt)


# --------------------------------------------------------------------------------------
# GameProblem  (src/problem/problem.jl:19-53)
# --------------------------------------------------------------------------------------
@dataclass
class Record:
    """One Statistics entry (src/struct/statistics.jl:44-57); only the .max values."""
    outer: int
    res: float
    dyn: float
    con: float
    sta: float
    opt: float
    delta: float


class GameProblem:
    def __init__(self, N, dt, x0, model, opts: Options, game_obj, game_con):
        self.probsize = ProblemSize(N, model)
        self.N, self.dt, self.model, self.opts = N, dt, model, opts
        self.x0 = np.asarray(x0, float)
        self.game_obj, self.game_con = game_obj, game_con
        self.core = NewtonCore(self.probsize)
        self.pdtraj = PrimalDualTraj(self.probsize, dt)
        self.pdtraj_trial = PrimalDualTraj(self.probsize, dt)
        self.dpdtraj = PrimalDualTraj(self.probsize, dt)
        self.stats: List[Record] = []
        self.n_newton = 0
        game_con.set_constraint_params(opts)                   # problem.jl:48


# --------------------------------------------------------------------------------------
# Residual / Jacobian  (src/problem/global_quantities.jl, src/constraints/constraint_derivatives.jl)
# --------------------------------------------------------------------------------------
def dynamics_residual(model, pd, k):
    """local_quantities.jl:5-14, 1-based stage k."""
    return rk2(model, pd.X[k - 1], pd.U[k - 1], pd.dt) - pd.X[k]


def residual(prob: GameProblem, pd: Optional[PrimalDualTraj] = None):
    """global_quantities.jl:9-65.  Writes and returns prob.core.res (reference row order)."""
    pd = pd or prob.pdtraj
    ps, core, model, obj, gc = prob.probsize, prob.core, prob.model, prob.game_obj, prob.game_con
    N, p, pu = ps.N, ps.p, ps.pu
    res = core.res
    res[:] = 0.0
    for i in range(1, p + 1):                                         # :24-41 cost
        for k in range(1, N + 1):
            q, r = obj.gradient(i - 1, k, pd.X[k - 1], pd.U[k - 1], pd.dt)
            if k >= 2:
                res[core.vert[("opt", i, "x", k)]] += q
            if k <= N - 1:
                res[core.vert[("opt", i, "u", k)]] += r[pu[i - 1]]
    for k in range(1, N):                                             # :43-54 dynamics penalty
        A, B = rk2_jacobian(model, pd.X[k - 1], pd.U[k - 1], pd.dt)
        for i in range(1, p + 1):
            lam = pd.du[i - 1, k - 1]
            if k >= 2:
                res[core.vert[("opt", i, "x", k)]] += A.T @ lam
            res[core.vert[("opt", i, "u", k)]] += B[:, pu[i - 1]].T @ lam
            res[core.vert[("opt", i, "x", k + 1)]] += -lam
    _constraint_residual(prob, pd)                                    # :57
    for k in range(1, N):                                             # :60-63
        res[core.vert[("dyn", k)]] += dynamics_residual(model, pd, k)
    return res


def _expand_convals(prob, pd):
    for _, _, cv in prob.game_con.all_convals():
        cv.evaluate(pd.X, pd.U)
        cv.jacobian(pd.X, pd.U)
        cv.cost_expansion()


def _constraint_residual(prob, pd):
    """constraint_derivatives.jl:39-74."""
    ps, core, gc = prob.probsize, prob.core, prob.game_con
    _expand_convals(prob, pd)
    for i in range(1, ps.p + 1):
        for cv in gc.state_conval[i - 1]:
            for j, k in enumerate(cv.inds):
                if 2 <= k <= ps.N:
                    core.res[core.vert[("opt", i, "x", k)]] += cv.grad[j]
    for cv in gc.control_conval:
        for j, k in enumerate(cv.inds):
            for i in range(1, ps.p + 1):
                core.res[core.vert[("opt", i, "u", k)]] += cv.grad[j][ps.pu[i - 1]]


def regularize_residual(prob, pd, pd_ref):
    """global_quantities.jl:67-86."""
    ps, core, reg = prob.probsize, prob.core, prob.opts.reg
    for k in range(1, ps.N):
        dx = pd.X[k] - pd_ref.X[k]
        du_ = pd.U[k - 1] - pd_ref.U[k - 1]
        for i in range(1, ps.p + 1):
            core.res[core.vert[("opt", i, "x", k + 1)]] += reg.x * dx
            core.res[core.vert[("opt", i, "u", k)]] += reg.u * du_[ps.pu[i - 1]]


def residual_jacobian(prob: GameProblem, pd: Optional[PrimalDualTraj] = None, regularize=True):
    """global_quantities.jl:109-193 → dense S×S array in the reference's (row, col) order."""
    pd = pd or prob.pdtraj
    ps, core, model, obj, gc = prob.probsize, prob.core, prob.model, prob.game_obj, prob.game_con
    N, n, p, pu, S = ps.N, ps.n, ps.p, ps.pu, ps.S
    J = np.zeros((S, S))
    V, H = core.vert, core.horiz

    def add(vkey, hkey, M):
        J[np.ix_(V[vkey], H[hkey])] += M

    for i in range(1, p + 1):                                         # :128-145 cost
        for k in range(1, N + 1):
            Qh, Rh = obj.hessian(i - 1, k, pd.X[k - 1], pd.U[k - 1], pd.dt)
            if k >= 2:
                add(("opt", i, "x", k), ("x", k), Qh)
            if k <= N - 1:
                add(("opt", i, "u", k), ("u", i, k), Rh[np.ix_(pu[i - 1], pu[i - 1])])
    _expand_convals(prob, pd)                                         # :148 → constraint_derivatives.jl:1-36
    for i in range(1, p + 1):
        for cv in gc.state_conval[i - 1]:
            for j, k in enumerate(cv.inds):
                if 2 <= k <= N:
                    add(("opt", i, "x", k), ("x", k), cv.hess[j])
    for cv in gc.control_conval:
        for j, k in enumerate(cv.inds):
            for i in range(1, p + 1):
                add(("opt", i, "u", k), ("u", i, k), cv.hess[j][np.ix_(pu[i - 1], pu[i - 1])])
    for k in range(1, N):                                             # :151-172 dynamics
        A, B = rk2_jacobian(model, pd.X[k - 1], pd.U[k - 1], pd.dt)
        if k >= 2:
            add(("dyn", k), ("x", k), A)
        for i in range(1, p + 1):
            add(("dyn", k), ("u", i, k), B[:, pu[i - 1]])
        add(("dyn", k), ("x", k + 1), -np.eye(n))
        for i in range(1, p + 1):
            if k >= 2:
                add(("opt", i, "x", k), ("λ", i, k), A.T)
            add(("opt", i, "u", k), ("λ", i, k), B[:, pu[i - 1]].T)
            add(("opt", i, "x", k + 1), ("λ", i, k), -np.eye(n))
    if regularize:                                                    # :176-193
        reg = prob.opts.reg
        for k in range(1, N):
            for i in range(1, p + 1):
                add(("opt", i, "x", k + 1), ("x", k + 1), reg.x * np.eye(n))
                add(("opt", i, "u", k), ("u", i, k), reg.u * np.eye(ps.mi[i - 1]))
    core.jac = J
    return J


# --------------------------------------------------------------------------------------
# Violations  (src/struct/violations.jl)
# --------------------------------------------------------------------------------------
def dynamics_violation(prob, pd):
    return max(np.max(np.abs(dynamics_residual(prob.model, pd, k))) for k in range(1, prob.N))   # :18-26


def control_violation(prob, pd):
    vio = np.zeros(prob.N - 1)                                        # :57-67
    for cv in prob.game_con.control_conval:
        cv.evaluate(pd.X, pd.U)
        cv.max_violation()
        idx = np.array(cv.inds) - 1
        vio[idx] = np.maximum(vio[idx], cv.c_max)
    return float(vio.max())


def state_violation(prob, pd):
    vio = np.zeros(prob.N)                                            # :101-114
    for i in range(prob.probsize.p):
        for cv in prob.game_con.state_conval[i]:
            cv.evaluate(pd.X, pd.U)
            cv.max_violation()
            idx = np.array(cv.inds) - 1
            vio[idx] = np.maximum(vio[idx], cv.c_max)
    return float(vio.max())


def optimality_violation(core: NewtonCore):
    ps = core.ps                                                      # :153-168
    best = 0.0
    for i in range(1, ps.p + 1):
        for k in range(1, ps.N + 1):
            if k >= 2:
                best = max(best, float(np.max(np.abs(core.res[core.vert[("opt", i, "x", k)]]))))
            if k <= ps.N - 1:
                best = max(best, float(np.max(np.abs(core.res[core.vert[("opt", i, "u", k)]]))))
    return best


def violation_vectors(prob, pd):
    """The `.vio` vectors of dynamics_violation / control_violation / state_violation / optimality_violation
    (struct/violations.jl:18-26, 57-67, 101-114, 153-168): dyn [N-1], con [N-1], sta [N], opt [N].  Leaves residual!(prob, pd)
    in core.res like record! does."""
    ps, core = prob.probsize, prob.core
    residual(prob, pd)
    dyn = np.array([np.max(np.abs(dynamics_residual(prob.model, pd, k))) for k in range(1, prob.N)])
    con = np.zeros(prob.N - 1)
    for cv in prob.game_con.control_conval:
        cv.evaluate(pd.X, pd.U)
        cv.max_violation()
        idx = np.array(cv.inds) - 1
        con[idx] = np.maximum(con[idx], cv.c_max)
    sta = np.zeros(prob.N)
    for i in range(ps.p):
        for cv in prob.game_con.state_conval[i]:
            cv.evaluate(pd.X, pd.U)
            cv.max_violation()
            idx = np.array(cv.inds) - 1
            sta[idx] = np.maximum(sta[idx], cv.c_max)
    opt = np.zeros(prob.N)
    for i in range(1, ps.p + 1):
        for k in range(1, ps.N + 1):
            if k >= 2:
                opt[k - 1] = max(opt[k - 1], float(np.max(np.abs(core.res[core.vert[("opt", i, "x", k)]]))))
            if k <= ps.N - 1:
                opt[k - 1] = max(opt[k - 1], float(np.max(np.abs(core.res[core.vert[("opt", i, "u", k)]]))))
    return dyn, con, sta, opt


def record(prob, pd, delta, k_out):
    """statistics.jl:44-57: residual! is re-run WITHOUT regularisation before the norms are taken."""
    residual(prob, pd)
    rec = Record(outer=k_out, res=float(np.abs(prob.core.res).sum() / prob.probsize.S),
                 dyn=dynamics_violation(prob, pd), con=control_violation(prob, pd),
                 sta=state_violation(prob, pd), opt=optimality_violation(prob.core), delta=delta)
    prob.stats.append(rec)
    return rec


# --------------------------------------------------------------------------------------
# Solver  (src/problem/solver_methods.jl:5-125)
# --------------------------------------------------------------------------------------
def kkt_solve(J, res):
    """Δtraj = −(lu(jac) \\ res) (solver_methods.jl:87).  UMFPACK stand-in: LAPACK dense LU with
    partial pivoting for small S, SuperLU (scipy splu) above that."""
    S = len(res)
    if S <= 1200:
        return -np.linalg.solve(J, res)
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    return -spla.splu(sp.csc_matrix(J)).solve(res)


def line_search(prob: GameProblem, res_norm: float):
    """solver_methods.jl:105-125."""
    opts, core = prob.opts, prob.core
    j, alpha = 1, 1.0
    while j < opts.ls_iter:
        update_traj(prob.pdtraj_trial, prob.pdtraj, alpha, prob.dpdtraj)
        residual(prob, prob.pdtraj_trial)
        if opts.regularize:
            regularize_residual(prob, prob.pdtraj_trial, prob.pdtraj)
        res_norm_trial = np.abs(core.res).sum() / len(core.res)
        if res_norm_trial <= (1.0 - alpha * opts.beta) * res_norm:
            break
        alpha *= opts.alpha_decrease
        j += 1
    return alpha, j


def inner_iteration(prob: GameProblem, LS_count: int, delta: float, k: int, l: int):
    """solver_methods.jl:67-103.  Returns (LS_count, control_flow, Δ)."""
    core, opts = prob.core, prob.opts
    residual(prob, prob.pdtraj)
    if opts.regularize:
        regularize_residual(prob, prob.pdtraj, prob.pdtraj)
    rec = record(prob, prob.pdtraj, delta, k)
    res_norm = np.abs(core.res).sum() / len(core.res)
    delta = 0.0
    if rec.opt < opts.eps_opt:
        return LS_count, "break", delta
    J = residual_jacobian(prob, prob.pdtraj, regularize=True)
    dtraj = kkt_solve(J, core.res)
    prob.last_dtraj = dtraj
    prob.n_newton += 1
    set_traj(core, prob.dpdtraj, dtraj)
    alpha, j = line_search(prob, res_norm)
    LS_count = LS_count + 1 if j == opts.ls_iter else 0
    update_traj(prob.pdtraj, prob.pdtraj, alpha, prob.dpdtraj)
    delta = delta_step(prob.dpdtraj, alpha)
    if delta < opts.delta_min:
        return LS_count, "break", delta
    return LS_count, "continue", delta


def newton_solve(prob: GameProblem, Z0=None, L0=None, rng=None):
    """solver_methods.jl:5-65.

    Z0 [N, n+m] / L0 [p, N-1, n]: the initial iterate BEFORE x_1←x0 and the RK3 rollout (what
    init_traj! produces).  The reference draws it from Julia's MersenneTwister (seed 100); that stream
    cannot be reproduced here, so callers pass the iterate explicitly (or a numpy Generator).
    With opts.shift < N the previous trajectory is shifted (MPC warm start) exactly as :13 does.
    """
    opts, model, gc, ps = prob.opts, prob.model, prob.game_con, prob.probsize
    rng = rng or np.random.default_rng(opts.seed)
    if Z0 is not None:
        prob.pdtraj.X[:] = np.asarray(Z0)[:, : ps.n]
        prob.pdtraj.U[:] = np.asarray(Z0)[:, ps.n:]
        prob.pdtraj.du[:] = np.asarray(L0)
        prob.pdtraj.X[0] = prob.x0
    else:
        init_traj(prob.pdtraj, prob.x0, lambda k: rng.random(k), opts.amplitude_init, opts.shift)
    prob.pdtraj_trial = prob.pdtraj.copy()
    prob.dpdtraj.X[:] = 0; prob.dpdtraj.U[:] = 0; prob.dpdtraj.du[:] = 0
    rollout_rk3(model, prob.pdtraj)                                        # :17
    prob.stats = []                                                        # :24
    prob.n_newton = 0
    if opts.dual_reset:
        gc.reset()                                                         # :25
    delta = 0.0
    out = 0
    for k in range(1, opts.outer_iter + 1):                                # :30
        out = k
        opts.reg.set(opts.reg_0)
        LS_count = 0
        for l in range(1, opts.inner_iter + 1):                            # :38
            opts.reg.set(opts.reg_0 * l ** 4)                              # :39
            LS_count, flow, delta = inner_iteration(prob, LS_count, delta, k, l)
            if LS_count >= 1 or flow == "break":                           # :43
                break
        last = prob.stats[-1]
        if k == opts.outer_iter or (last.dyn < opts.eps_dyn and last.con < opts.eps_con and
                                    last.sta < opts.eps_sta and last.opt < opts.eps_opt):   # :49-55
            break
        gc.evaluate(prob.pdtraj.X, prob.pdtraj.U)                          # :57
        gc.dual_update()                                                   # :58
        gc.penalty_update()                                                # :61
    rec = record(prob, prob.pdtraj, delta, out)                            # :63
    prob.converged = bool(rec.dyn < opts.eps_dyn and rec.con < opts.eps_con and
                          rec.sta < opts.eps_sta and rec.opt < opts.eps_opt)
    return prob


# --------------------------------------------------------------------------------------
# Neutral problem spec → oracle objects (the tests build both sides from one dict)
# --------------------------------------------------------------------------------------
def problem_from_spec(spec: dict, x0=None, xf=None, opts: Optional[Options] = None) -> GameProblem:
    """Build a GameProblem from the plain-dict spec the host package emits (`GameProblem.to_spec()`).

    Constraint order is canonical: per player [collision pairs j≠i ascending | state bound | walls |
    circles]; one control-bound conval.  That is also the row order of the C-ABI's conλ/conμ arrays.
    """
    if spec["model"] == "quadrotor":
        model = QuadrotorGame(p=spec["p"], mass=spec.get("mass", 0.5))
    else:
        model = make_model(spec["model"], spec["p"], spec.get("d", 2), spec.get("lf", 0.05), spec.get("lr", 0.05))
    N, dt, p = spec["N"], spec["dt"], spec["p"]
    ps = ProblemSize(N, model)
    xf_ = np.asarray(spec["xf"] if xf is None else xf, float).reshape(p, -1)
    obj = GameObjective([np.asarray(q, float) for q in spec["Q"]], [np.asarray(r, float) for r in spec["R"]],
                        list(xf_), [np.asarray(u, float) for u in spec["uf"]], N, model)
    if spec.get("collision_cost"):
        obj.add_collision_cost(spec["collision_cost"]["radius"], spec["collision_cost"]["mu"])
    gc = GameConstraintValues(ps)
    rad = spec.get("collision_radius")          # p×p matrix of pair radii, 0 ⇒ no constraint
    sb = spec.get("state_bounds") or [None] * p
    walls = spec.get("walls") or [[] for _ in range(p)]
    circles = spec.get("circles") or [[] for _ in range(p)]
    for i in range(p):
        if rad is not None:
            for j in range(p):
                if j != i and rad[i][j] > 0:
                    (gc.add_spherical_collision_avoidance if spec.get("spherical") else gc.add_collision_avoidance)(rad[i][j], i, j)
        for bound in ([] if sb[i] is None else [sb[i]] if isinstance(sb[i], dict) else sb[i]):
            gc.add_state_bound(i, bound["x_max"], bound["x_min"])
        if len(walls[i]):
            gc.add_wall_constraint([Wall(np.array(w[0:2]), np.array(w[2:4]), np.array(w[4:6])) for w in walls[i]], i)
        if len(circles[i]):
            c = np.asarray(circles[i], float)
            gc.add_circle_constraint(c[:, 0], c[:, 1], c[:, 2], i)
        w3 = (spec.get("walls3d") or [[] for _ in range(p)])[i]
        if len(w3):
            gc.add_wall3d_constraint([Wall3D(np.array(w[0:3]), np.array(w[3:6]), np.array(w[6:9]), np.array(w[9:12])) for w in w3], i)
        cy = (spec.get("cylinders") or [[] for _ in range(p)])[i]
        if len(cy):
            gc.add_cylinder_constraint([CylinderWall(np.array(w[0:3]), w[3], w[4], w[5]) for w in cy], i)
    if spec.get("control_bounds"):
        gc.add_control_bound(spec["control_bounds"]["u_max"], spec["control_bounds"]["u_min"])
    opts = opts or options_from_dict(spec.get("opts", {}))
    return GameProblem(N, dt, spec["x0"] if x0 is None else x0, model, opts, obj, gc)


def options_from_dict(d: dict) -> Options:
    o = Options()
    for k, v in d.items():
        if k == "alphax_dual":
            v = list(v)
        if hasattr(o, k):
            setattr(o, k, v)
    return o


def pack_multipliers(prob: GameProblem):
    """Constraint λ and μ in the C-ABI layout [N-1][nrow]: stage k (1-based) holds the state rows of
    knot k+1 (player-major, canonical order) followed by the control rows of knot k."""
    ps, gc = prob.probsize, prob.game_con
    lam, mu = [], []
    for k in range(1, ps.N):
        rl, rm = [], []
        for i in range(ps.p):
            for cv in gc.state_conval[i]:
                rl.append(cv.lam[k - 1]); rm.append(cv.mu[k - 1])     # inds = 2..N ⇒ j = k-1 ↔ knot k+1
        for cv in gc.control_conval:
            rl.append(cv.lam[k - 1]); rm.append(cv.mu[k - 1])
        lam.append(np.concatenate(rl) if rl else np.zeros(0))
        mu.append(np.concatenate(rm) if rm else np.zeros(0))
    return np.array(lam), np.array(mu)


def unpack_multipliers(prob: GameProblem, lam, mu):
    ps, gc = prob.probsize, prob.game_con
    for k in range(1, ps.N):
        o = 0
        for i in range(ps.p):
            for cv in gc.state_conval[i]:
                P = cv.con.length()
                cv.lam[k - 1] = lam[k - 1, o:o + P]; cv.mu[k - 1] = mu[k - 1, o:o + P]; o += P
        for cv in gc.control_conval:
            P = cv.con.length()
            cv.lam[k - 1] = lam[k - 1, o:o + P]; cv.mu[k - 1] = mu[k - 1, o:o + P]; o += P


# --------------------------------------------------------------------------------------
# Iterative best response (src/problem/solver_methods.jl:133-289, global_quantities.jl:199-365,
# core/newton_core.jl:205-294, struct/violations.jl:28-37,69-82,116-134,170-183)
# Players are 1-based here, like the reference.
# --------------------------------------------------------------------------------------
def vertical_mask(core: NewtonCore, i: int, splitted_state: bool = False):
    """newton_core.jl:205-245: rows of player i's best-response problem: opt_i x, opt_i u_i, dyn."""
    ps = core.ps
    pi = np.asarray(ps.pz[i - 1]) if splitted_state else np.arange(ps.n)
    msk = []
    for k in range(2, ps.N + 1):
        msk.extend(core.vert[("opt", i, "x", k)][pi])
    for k in range(1, ps.N):
        msk.extend(core.vert[("opt", i, "u", k)])
    for k in range(1, ps.N):
        msk.extend(core.vert[("dyn", k)][pi])
    return np.array(msk)


def horizontal_mask(core: NewtonCore, i: int, splitted_state: bool = False):
    """newton_core.jl:248-294: columns x, u_i, λ_i."""
    ps = core.ps
    pi = np.asarray(ps.pz[i - 1]) if splitted_state else np.arange(ps.n)
    msk = []
    for k in range(2, ps.N + 1):
        msk.extend(core.horiz[("x", k)][pi])
    for k in range(1, ps.N):
        msk.extend(core.horiz[("u", i, k)])
    for k in range(1, ps.N):
        msk.extend(core.horiz[("λ", i, k)][pi])
    return np.array(msk)


def ibr_residual(prob: GameProblem, pd: PrimalDualTraj, i: int):
    """global_quantities.jl:204-256: cost and dynamics-penalty terms of player i only; constraint terms of every
    player (harmless: the caller masks the rows); dynamics rows."""
    ps, core, model, obj = prob.probsize, prob.core, prob.model, prob.game_obj
    N, pu = ps.N, ps.pu
    res = core.res
    res[:] = 0.0
    for k in range(1, N + 1):
        q, r = obj.gradient(i - 1, k, pd.X[k - 1], pd.U[k - 1], pd.dt)
        if k >= 2:
            res[core.vert[("opt", i, "x", k)]] += q
        if k <= N - 1:
            res[core.vert[("opt", i, "u", k)]] += r[pu[i - 1]]
    for k in range(1, N):
        A, B = rk2_jacobian(model, pd.X[k - 1], pd.U[k - 1], pd.dt)
        lam = pd.du[i - 1, k - 1]
        if k >= 2:
            res[core.vert[("opt", i, "x", k)]] += A.T @ lam
        res[core.vert[("opt", i, "u", k)]] += B[:, pu[i - 1]].T @ lam
        res[core.vert[("opt", i, "x", k + 1)]] += -lam
    _constraint_residual(prob, pd)
    for k in range(1, N):
        res[core.vert[("dyn", k)]] += dynamics_residual(model, pd, k)
    return res


def regularize_ibr_residual(prob, pd, pd_ref, i):
    """global_quantities.jl:258-276."""
    ps, core, reg = prob.probsize, prob.core, prob.opts.reg
    for k in range(1, ps.N):
        core.res[core.vert[("opt", i, "x", k + 1)]] += reg.x * (pd.X[k] - pd_ref.X[k])
        core.res[core.vert[("opt", i, "u", k)]] += reg.u * (pd.U[k - 1] - pd_ref.U[k - 1])[ps.pu[i - 1]]


def ibr_residual_jacobian(prob: GameProblem, pd: PrimalDualTraj, i: int, regularize=True):
    """global_quantities.jl:288-365 (dense S×S; only player i's cost/dynamics blocks, every player's constraint blocks)."""
    ps, core, model, obj, gc = prob.probsize, prob.core, prob.model, prob.game_obj, prob.game_con
    N, n, pu, S = ps.N, ps.n, ps.pu, ps.S
    J = np.zeros((S, S))
    V, H = core.vert, core.horiz

    def add(vkey, hkey, M):
        J[np.ix_(V[vkey], H[hkey])] += M

    for k in range(1, N + 1):
        Qh, Rh = obj.hessian(i - 1, k, pd.X[k - 1], pd.U[k - 1], pd.dt)
        if k >= 2:
            add(("opt", i, "x", k), ("x", k), Qh)
        if k <= N - 1:
            add(("opt", i, "u", k), ("u", i, k), Rh[np.ix_(pu[i - 1], pu[i - 1])])
    _expand_convals(prob, pd)
    for pl in range(1, ps.p + 1):
        for cv in gc.state_conval[pl - 1]:
            for j, k in enumerate(cv.inds):
                if 2 <= k <= N:
                    add(("opt", pl, "x", k), ("x", k), cv.hess[j])
    for cv in gc.control_conval:
        for j, k in enumerate(cv.inds):
            for pl in range(1, ps.p + 1):
                add(("opt", pl, "u", k), ("u", pl, k), cv.hess[j][np.ix_(pu[pl - 1], pu[pl - 1])])
    for k in range(1, N):
        A, B = rk2_jacobian(model, pd.X[k - 1], pd.U[k - 1], pd.dt)
        if k >= 2:
            add(("dyn", k), ("x", k), A)
        add(("dyn", k), ("u", i, k), B[:, pu[i - 1]])
        add(("dyn", k), ("x", k + 1), -np.eye(n))
        if k >= 2:
            add(("opt", i, "x", k), ("λ", i, k), A.T)
        add(("opt", i, "u", k), ("λ", i, k), B[:, pu[i - 1]].T)
        add(("opt", i, "x", k + 1), ("λ", i, k), -np.eye(n))
    if regularize:
        reg = prob.opts.reg
        for k in range(1, N):
            add(("opt", i, "x", k + 1), ("x", k + 1), reg.x * np.eye(n))
            add(("opt", i, "u", k), ("u", i, k), reg.u * np.eye(ps.mi[i - 1]))
    core.jac = J
    return J


def record_ibr(prob, pd, delta, k_out, i):
    """statistics.jl:59-72: player-i violations (violations.jl:28-37, 69-82, 116-134, 170-183).
    res is the full-game residual norm (logging only)."""
    ps, core = prob.probsize, prob.core
    residual(prob, pd)
    res_full = float(np.abs(core.res).sum() / ps.S)
    pz, pui = np.asarray(ps.pz[i - 1]), np.asarray(ps.pu[i - 1])
    dyn = max(np.max(np.abs(dynamics_residual(prob.model, pd, k)[pz])) for k in range(1, prob.N))
    con = np.zeros(prob.N - 1)
    for cv in prob.game_con.control_conval:
        cv.evaluate(pd.X, pd.U)
        # the reference indexes the stacked [u-u_max; u_min-u] rows with pu[i]; with all bounds finite those are the
        # upper-bound rows of player i (violations.jl:69-82)
        cmax = np.array([max(0.0, float(np.max(v[pui]))) for v in cv.vals])
        idx = np.array(cv.inds) - 1
        con[idx] = np.maximum(con[idx], cmax)
    sta = np.zeros(prob.N)
    for cv in prob.game_con.state_conval[i - 1]:
        cv.evaluate(pd.X, pd.U)
        cv.max_violation()
        idx = np.array(cv.inds) - 1
        sta[idx] = np.maximum(sta[idx], cv.c_max)
    opt = 0.0
    for k in range(1, ps.N + 1):
        if k >= 2:
            opt = max(opt, float(np.max(np.abs(core.res[core.vert[("opt", i, "x", k)]]))))
        if k <= ps.N - 1:
            opt = max(opt, float(np.max(np.abs(core.res[core.vert[("opt", i, "u", k)]]))))
    rec = Record(outer=k_out, res=res_full, dyn=float(dyn), con=float(con.max()) if len(con) else 0.0,
                 sta=float(sta.max()), opt=opt, delta=delta)
    prob.stats.append(rec)
    return rec


def ibr_line_search(prob, res_norm, i):
    """solver_methods.jl:267-289."""
    opts, core = prob.opts, prob.core
    vm = vertical_mask(core, i)
    j, alpha = 1, 1.0
    while j < opts.ls_iter:
        update_traj(prob.pdtraj_trial, prob.pdtraj, alpha, prob.dpdtraj)
        ibr_residual(prob, prob.pdtraj_trial, i)
        if opts.regularize:
            regularize_ibr_residual(prob, prob.pdtraj_trial, prob.pdtraj, i)
        if np.abs(core.res[vm]).sum() / len(vm) <= (1.0 - alpha * opts.beta) * res_norm:
            break
        alpha *= opts.alpha_decrease
        j += 1
    return alpha, j


def ibr_inner_iteration(prob, LS_count, delta, k, l, i):
    """solver_methods.jl:226-265."""
    core, opts = prob.core, prob.opts
    vm, hm = vertical_mask(core, i), horizontal_mask(core, i)
    ibr_residual(prob, prob.pdtraj, i)
    masked = core.res.copy()
    rec = record_ibr(prob, prob.pdtraj, delta, k, i)          # (re-runs the full residual!, like the reference's record!)
    core.res[:] = masked
    # NB the reference's optimality_violation(core, i) is taken on core.res as left by residual_norm → residual!, whose
    # player-i rows equal the ibr_residual! ones; record_ibr above already used the full residual for them.
    res_norm = np.abs(core.res[vm]).sum() / len(vm)
    delta = 0.0
    if rec.opt < opts.eps_opt:
        return LS_count, "break", delta
    J = ibr_residual_jacobian(prob, prob.pdtraj, i, regularize=True)
    dtraj = np.zeros(prob.probsize.S)
    dtraj[hm] = -np.linalg.solve(J[np.ix_(vm, hm)], core.res[vm])
    prob.last_dtraj = dtraj
    prob.n_newton += 1
    set_traj(core, prob.dpdtraj, dtraj)
    alpha, j = ibr_line_search(prob, res_norm, i)
    LS_count = LS_count + 1 if j == opts.ls_iter else 0
    update_traj(prob.pdtraj, prob.pdtraj, alpha, prob.dpdtraj)
    delta = delta_step(prob.dpdtraj, alpha)
    if delta < opts.delta_min:
        return LS_count, "break", delta
    return LS_count, "continue", delta


def ibr_newton_solve_player(prob: GameProblem, i: int):
    """solver_methods.jl:168-224: best response of player i (1-based)."""
    opts, gc = prob.opts, prob.game_con
    if opts.dual_reset:
        gc.reset()
        prob.pdtraj.du[:] = 0.0                               # reset_duals!(pdtraj) (primal_dual_traj.jl:149-158)
        prob.pdtraj_trial.du[:] = 0.0
    delta, out = 0.0, 0
    for k in range(1, opts.outer_iter + 1):
        out = k
        LS_count = 0
        for l in range(1, opts.inner_iter + 1):
            opts.reg.set(opts.reg_0 * l ** 4)
            LS_count, flow, delta = ibr_inner_iteration(prob, LS_count, delta, k, l, i)
            if LS_count >= 1 or flow == "break":
                break
        last = prob.stats[-1]
        if k == opts.outer_iter or (last.dyn < opts.eps_dyn and last.con < opts.eps_con and
                                    last.sta < opts.eps_sta and last.opt < opts.eps_opt):
            break
        gc.evaluate(prob.pdtraj.X, prob.pdtraj.U)
        gc.dual_update()
        gc.penalty_update()
    record_ibr(prob, prob.pdtraj, delta, out, i)
    return prob


def ibr_newton_solve(prob: GameProblem, Z0=None, L0=None, ibr_iter=100, ordering=None, delta_min=1e-9, rng=None):
    """solver_methods.jl:133-166.  Z0/L0 as in newton_solve."""
    opts, ps = prob.opts, prob.probsize
    ordering = list(ordering) if ordering is not None else list(range(1, ps.p + 1))
    prob.stats = []
    prob.n_newton = 0
    rng = rng or np.random.default_rng(opts.seed)
    if Z0 is not None:
        prob.pdtraj.X[:] = np.asarray(Z0)[:, : ps.n]
        prob.pdtraj.U[:] = np.asarray(Z0)[:, ps.n:]
        prob.pdtraj.du[:] = np.asarray(L0)
        prob.pdtraj.X[0] = prob.x0
    else:
        init_traj(prob.pdtraj, prob.x0, lambda k: rng.random(k), opts.amplitude_init, opts.shift)
    prob.pdtraj_trial = prob.pdtraj.copy()
    prob.dpdtraj.X[:] = 0; prob.dpdtraj.U[:] = 0; prob.dpdtraj.du[:] = 0
    rollout_rk3(prob.model, prob.pdtraj)
    change = [True] * ps.p
    prob.ibr_sweeps = 0
    for q in range(ibr_iter):
        prob.ibr_sweeps = q + 1
        for idx in range(ps.p):
            i = ordering[idx]
            ibr_newton_solve_player(prob, i)
            # maximum over the WHOLE recorded history (stats is never reset inside the loop, :156)
            change[i - 1] = not (delta_min > max(r.delta for r in prob.stats))
        residual(prob, prob.pdtraj)
        if all(not c for c in change):
            break
    return prob


# --------------------------------------------------------------------------------------
# Active-set analysis  (src/active_set/*.jl) — ORACLE ONLY (host-side analysis, off the solve path)
# --------------------------------------------------------------------------------------
def valid_c(dim, con, i, j, k, N, p):
    """valid(::CStamp, N, p) (active_set_stamp.jl:64-81; pinned test/active_set/active_set_stamp.jl:6-36), 1-based."""
    if dim == "v":
        return i < j and 1 <= i <= p and 1 <= j <= p and 2 <= k <= N
    if dim == "h":
        return 1 <= i <= p and 1 <= j <= p and 2 <= k <= N and i != j
    return False


class ActiveSetCore:
    """active_set_core.jl:57-160: the Newton system augmented with one row per unordered collision pair and knot
    (`("v","col",i,j,k)`, i < j) and one column per ordered pair (`("h","col",i,j,k)`): Sv = S + p(p−1)(N−1)/2 rows,
    Sh = S + p(p−1)(N−1) columns.  The first S rows / columns are NewtonCore's."""

    def __init__(self, ps: ProblemSize):
        self.ps = ps
        N, p, S = ps.N, ps.p, ps.S
        self.Sv, self.Sh = S + p * (p - 1) * (N - 1) // 2, S + p * (p - 1) * (N - 1)     # :81-82
        self.vert, self.horiz = {}, {}
        off = S
        for k in range(2, N + 1):                                                        # :118-126
            for i in range(1, p + 1):
                for j in range(i + 1, p + 1):
                    self.vert[("v", "col", i, j, k)] = off; off += 1
        assert off == self.Sv
        off = S
        for k in range(2, N + 1):                                                        # :151-159
            for i in range(1, p + 1):
                for j in range(1, p + 1):
                    if j != i:
                        self.horiz[("h", "col", i, j, k)] = off; off += 1
        assert off == self.Sh
        self.res = np.zeros(self.Sv)
        self.jac = np.zeros((self.Sv, self.Sh))
        self.vmask = np.arange(self.Sv)                                                  # :91-92
        self.hmask = np.arange(self.Sh)
        self.null_mat = np.zeros((0, 0))
        self.null_vec, self.null_dtraj, self.null_dlam = [], [], []                      # NullSpace, :5-27


def _collision_conval(game_con, i, j):
    """get_collision_conval (active_set_methods.jl:80-94), players 1-based."""
    px = game_con.probsize.px
    for cv in game_con.state_conval[i - 1]:
        if isinstance(cv.con, CollisionConstraint) and np.array_equal(cv.con.x2, px[j - 1]):
            return cv
    return None


def as_active(game_con, dim, i, j, k, tol=None):
    """active(game_con, stamp) (active_set_methods.jl:5-26): the `active` flag Altro keeps for the (i,j) collision row of knot k."""
    ps = game_con.probsize
    if not valid_c(dim, "col", i, j, k, ps.N, ps.p):
        return False
    cv = _collision_conval(game_con, i, j)
    l = cv.inds.index(k)
    tol = game_con.active_set_tolerance if tol is None else tol
    return bool((cv.vals[l, 0] >= -tol) | (cv.lam[l, 0] > 0))


def active_vertical_mask(ascore, game_con):
    """active_set_methods.jl:28-50: rows of the Newton system + the rows of the active collision pairs."""
    ps = ascore.ps
    keep = list(range(ps.S))
    for k in range(2, ps.N + 1):
        for i in range(1, ps.p + 1):
            for j in range(i + 1, ps.p + 1):
                if as_active(game_con, "v", i, j, k):
                    keep.append(ascore.vert[("v", "col", i, j, k)])
    ascore.vmask = np.array(keep)


def active_horizontal_mask(ascore, game_con):
    """active_set_methods.jl:52-74."""
    ps = ascore.ps
    keep = list(range(ps.S))
    for k in range(2, ps.N + 1):
        for i in range(1, ps.p + 1):
            for j in range(1, ps.p + 1):
                if j != i and as_active(game_con, "h", i, j, k):
                    keep.append(ascore.horiz[("h", "col", i, j, k)])
    ascore.hmask = np.array(keep)


def as_residual(ascore, prob, pd=None):
    """residual!(ascore, prob, pdtraj) (active_set_methods.jl:96-124): [KKT residual; collision values of the pairs i < j]."""
    pd = pd or prob.pdtraj
    ps = ascore.ps
    ascore.res[:] = 0.0
    ascore.res[:ps.S] = residual(prob, pd)
    prob.game_con.evaluate(pd.X, pd.U)
    for i in range(1, ps.p + 1):
        for j in range(i + 1, ps.p + 1):
            cv = _collision_conval(prob.game_con, i, j)
            if cv is not None:
                for l, k in enumerate(cv.inds):
                    ascore.res[ascore.vert[("v", "col", i, j, k)]] += cv.vals[l, 0]
    return ascore.res


def as_residual_jacobian(ascore, prob, pd=None):
    """residual_jacobian!(ascore, prob, pdtraj) (active_set_methods.jl:131-170): the KKT Jacobian bordered by ∇cᵀ in the
    column of the ordered pair (i,j) on player i's opt-x rows, and by ∇c in the row of the unordered pair on the x columns."""
    pd = pd or prob.pdtraj
    ps, core = ascore.ps, prob.core
    ascore.jac[:] = 0.0
    residual(prob, pd)
    ascore.jac[:ps.S, :ps.S] = residual_jacobian(prob, pd, regularize=False)     # :141 residual_jacobian! only — no regularize_residual_jacobian!
    prob.game_con.jacobian(pd.X, pd.U)
    for i in range(1, ps.p + 1):
        for j in range(1, ps.p + 1):
            if j == i:
                continue
            cv = _collision_conval(prob.game_con, i, j)
            if cv is None:
                continue
            for l, k in enumerate(cv.inds):
                ascore.jac[core.vert[("opt", i, "x", k)], ascore.horiz[("h", "col", i, j, k)]] += cv.jac[l, 0]       # :154-156
                if valid_c("v", "col", i, j, k, ps.N, ps.p):                                                        # :157-161
                    ascore.jac[ascore.vert[("v", "col", i, j, k)], core.horiz[("x", k)]] += cv.jac[l, 0]
    return ascore.jac


def update_nullspace(ascore, prob, pd=None, atol=1e-20):
    """update_nullspace! (active_set_methods.jl:173-184) + add_matrix! (active_set_core.jl:29-52): null space of the
    active-set Jacobian (LinearAlgebra.nullspace = right singular vectors of the singular values ≤ atol), each vector
    scattered to the full column set and scaled to unit mean absolute value."""
    pd = pd or prob.pdtraj
    ps = ascore.ps
    prob.game_con.evaluate(pd.X, pd.U)                       # update_active_set!(game_con, traj)
    active_vertical_mask(ascore, prob.game_con)
    active_horizontal_mask(ascore, prob.game_con)
    as_residual_jacobian(ascore, prob, pd)
    djac = ascore.jac[np.ix_(ascore.vmask, ascore.hmask)]
    _, sv, Vt = np.linalg.svd(djac, full_matrices=True)
    rank = int((sv > atol).sum())
    mat = Vt[rank:].T                                       # columns span the null space
    assert ps.S <= mat.shape[0] <= ascore.Sh                # active_set_core.jl:37
    ascore.null_mat = mat
    ascore.null_vec, ascore.null_dtraj, ascore.null_dlam = [], [], []
    for c in range(mat.shape[1]):
        vec = np.zeros(ascore.Sh)
        vec[ascore.hmask] = mat[:, c]
        vec /= np.mean(np.abs(vec))
        ascore.null_vec.append(vec); ascore.null_dtraj.append(vec[:ps.S]); ascore.null_dlam.append(vec[ps.S:])
    return ascore.null_mat
