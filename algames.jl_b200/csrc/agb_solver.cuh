// agb_solver.cuh — device code of the batched ALGAMES Newton/KKT + augmented-Lagrangian solve (sm_100a, FP64).
//
// One CTA (kThreads = 128) owns one game instance; the whole iterate, the KKT right-hand side / Newton step, the
// feedback gains of the stage-wise factorisation and the AL multipliers stay in shared memory for the entire
// newton_solve! loop (reference: src/problem/solver_methods.jl:5-125).  The KKT system is never
// formed: in time-major order it is block tridiagonal, and its block-LU with the pivot order
// (λ_k via the -I blocks, then u_k through S_k = Hu + BᵀPB with partial pivoting, then x_k) is the game Riccati
// recursion implemented in kkt_solve() below.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "agb_internal.h"

namespace agb {

#define AGB_FULL 0xffffffffu

struct Acc {            // norms of one residual evaluation (statistics.jl:44-57, violations.jl:18-168)
  double sum, opt, dyn, con, sta;
};

// ------------------------------------------------------------------------------------------------------------
// Per-player continuous dynamics and their Jacobians (all in-scope models are separable per player)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dyn_f(int model, double lf, double lr, const double s[4], const double u[2], double f[4]) {
  if (model == AGB_MODEL_DOUBLE_INTEGRATOR) {          // dynamics/double_integrator.jl:27-31
    f[0] = s[2]; f[1] = s[3]; f[2] = u[0]; f[3] = u[1];
  } else if (model == AGB_MODEL_UNICYCLE) {            // dynamics/unicycle.jl:27-32   s=[x,y,θ,v], u=[ω,a]
    double sn, cs; sincos(s[2], &sn, &cs);
    f[0] = cs * s[3]; f[1] = sn * s[3]; f[2] = u[0]; f[3] = u[1];
  } else {                                             // dynamics/bicycle.jl:28-41    s=[x,y,v,ψ], u=[a,δ]
    double beta = atan2(lr * tan(u[1]), lr + lf);
    double sn, cs; sincos(beta + s[3], &sn, &cs);
    f[0] = s[2] * cs; f[1] = s[2] * sn; f[2] = u[0]; f[3] = s[2] * sin(beta) / lr;
  }
}

// Fx[c][c'] (row-major 4x4), Fu[c][j] (4x2)
__device__ __forceinline__ void dyn_jac(int model, double lf, double lr, const double s[4], const double u[2],
                                        double Fx[16], double Fu[8]) {
#pragma unroll
  for (int q = 0; q < 16; q++) Fx[q] = 0.0;
#pragma unroll
  for (int q = 0; q < 8; q++) Fu[q] = 0.0;
  if (model == AGB_MODEL_DOUBLE_INTEGRATOR) {
    Fx[0 * 4 + 2] = 1.0; Fx[1 * 4 + 3] = 1.0; Fu[2 * 2 + 0] = 1.0; Fu[3 * 2 + 1] = 1.0;
  } else if (model == AGB_MODEL_UNICYCLE) {
    double sn, cs; sincos(s[2], &sn, &cs);
    Fx[0 * 4 + 2] = -sn * s[3]; Fx[0 * 4 + 3] = cs;
    Fx[1 * 4 + 2] = cs * s[3];  Fx[1 * 4 + 3] = sn;
    Fu[2 * 2 + 0] = 1.0; Fu[3 * 2 + 1] = 1.0;
  } else {
    double L = lr + lf, t = tan(u[1]);
    double beta = atan2(lr * t, L);
    double dbeta = lr * L * (1.0 + t * t) / (L * L + lr * lr * t * t);
    double sn, cs; sincos(beta + s[3], &sn, &cs);
    double sb, cb; sincos(beta, &sb, &cb);
    double v = s[2];
    Fx[0 * 4 + 2] = cs;      Fx[0 * 4 + 3] = -v * sn;
    Fx[1 * 4 + 2] = sn;      Fx[1 * 4 + 3] = v * cs;
    Fx[3 * 4 + 2] = sb / lr;
    Fu[0 * 2 + 1] = -v * sn * dbeta;
    Fu[1 * 2 + 1] = v * cs * dbeta;
    Fu[2 * 2 + 0] = 1.0;
    Fu[3 * 2 + 1] = v * cb * dbeta / lr;
  }
}

// discrete_dynamics(RK2,…) = explicit midpoint (problem/local_quantities.jl:13) and its Jacobian [A|B] (:26),
// by the chain rule: A = I + dt·Fx(xm)(I + dt/2·Fx(x)),  B = dt·(Fx(xm)·dt/2·Fu(x) + Fu(xm)).
__device__ __forceinline__ void rk2_jac(int model, double dt, double lf, double lr, const double s[4], const double u[2],
                                        double xn[4], double A[16], double B[8]) {
  double f0[4], sm[4], f1[4];
  dyn_f(model, lf, lr, s, u, f0);
#pragma unroll
  for (int c = 0; c < 4; c++) sm[c] = s[c] + (f0[c] * dt) / 2;
  dyn_f(model, lf, lr, sm, u, f1);
#pragma unroll
  for (int c = 0; c < 4; c++) xn[c] = s[c] + f1[c] * dt;
  double Fx0[16], Fu0[8], Fxm[16], Fum[8];
  dyn_jac(model, lf, lr, s, u, Fx0, Fu0);
  dyn_jac(model, lf, lr, sm, u, Fxm, Fum);
  const double h = dt / 2;
#pragma unroll
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < 4; q++) acc += Fxm[r * 4 + q] * ((q == c ? 1.0 : 0.0) + h * Fx0[q * 4 + c]);
      A[r * 4 + c] = (r == c ? 1.0 : 0.0) + dt * acc;
    }
#pragma unroll
    for (int j = 0; j < 2; j++) {
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < 4; q++) acc += Fxm[r * 4 + q] * (h * Fu0[q * 2 + j]);
      B[r * 2 + j] = dt * (acc + Fum[r * 2 + j]);
    }
  }
}

__device__ __forceinline__ void rk2_only(int model, double dt, double lf, double lr, const double s[4], const double u[2],
                                         double xn[4]) {
  double f0[4], sm[4], f1[4];
  dyn_f(model, lf, lr, s, u, f0);
#pragma unroll
  for (int c = 0; c < 4; c++) sm[c] = s[c] + (f0[c] * dt) / 2;
  dyn_f(model, lf, lr, sm, u, f1);
#pragma unroll
  for (int c = 0; c < 4; c++) xn[c] = s[c] + f1[c] * dt;
}

// discrete_dynamics(RK3,…) used by rollout! (solver_methods.jl:17-18)
__device__ __forceinline__ void rk3_step(int model, double dt, double lf, double lr, const double s[4], const double u[2],
                                         double xn[4]) {
  double k1[4], k2[4], k3[4], t[4];
  dyn_f(model, lf, lr, s, u, k1);
#pragma unroll
  for (int c = 0; c < 4; c++) { k1[c] *= dt; t[c] = s[c] + k1[c] / 2; }
  dyn_f(model, lf, lr, t, u, k2);
#pragma unroll
  for (int c = 0; c < 4; c++) { k2[c] *= dt; t[c] = s[c] - k1[c] + 2 * k2[c]; }
  dyn_f(model, lf, lr, t, u, k3);
#pragma unroll
  for (int c = 0; c < 4; c++) { k3[c] *= dt; xn[c] = s[c] + (k1[c] + 4 * k2[c] + k3[c]) / 6; }
}

// ------------------------------------------------------------------------------------------------------------
// Instance context: shared-memory views of one game instance
// ------------------------------------------------------------------------------------------------------------
template <int P>
struct Inst {
  static constexpr int n = 4 * P, m = 2 * P, b = P * n + m + n, W = m + n + 1, KUS = m * (n + 1);
  static constexpr int OX = 0, OU = P * n, OD = P * n + m;     // offsets inside one stage of R: [rx(p·n) | ru(m) | rd(n)]

  const DevDesc* __restrict__ d;
  int N, K, nrow, model, npairs, has_cc;
  double dt;
  double *X, *U, *L, *R, *KU, *AB, *CL, *CM, *CW, *CC, *Pm, *Sv, *Aug, *Acl, *Hpos, *Hd, *xf, *Q, *Rw, *uf, *red;
  int tid, lane, warp;

  __device__ void bind(const DevDesc* dd, double* sm) {
    d = dd; N = dd->N; K = dd->K; nrow = dd->nrow; model = dd->model; npairs = dd->npairs; has_cc = dd->has_cc; dt = dd->dt;
    X = sm + dd->o_X; U = sm + dd->o_U; L = sm + dd->o_L; R = sm + dd->o_R; KU = sm + dd->o_KU; AB = sm + dd->o_AB;
    CL = sm + dd->o_CL; CM = sm + dd->o_CM; CW = sm + dd->o_CW; CC = sm + dd->o_CC; Pm = sm + dd->o_P; Sv = sm + dd->o_Sv;
    Aug = sm + dd->o_Aug; Acl = sm + dd->o_Acl; Hpos = sm + dd->o_Hpos; Hd = sm + dd->o_Hd;
    xf = sm + dd->o_par; Q = xf + n; Rw = Q + n; uf = Rw + m; red = sm + dd->o_red;
    tid = threadIdx.x; lane = tid & 31; warp = tid >> 5;
  }

  // ---- iterate accessors; TRIAL reads Z + alpha·Δ with Δ held in R (update_traj!, primal_dual_traj.jl:109-128)
  template <bool TRIAL> __device__ __forceinline__ double xg(int k, int a, double alpha) const {
    double v = X[k * n + a];
    if (TRIAL) { if (k > 0) v += alpha * R[(k - 1) * b + OD + a]; }
    return v;
  }
  template <bool TRIAL> __device__ __forceinline__ double ug(int s, int idx, double alpha) const {
    double v = U[s * m + idx];
    if (TRIAL) v += alpha * R[s * b + OU + idx];
    return v;
  }
  template <bool TRIAL> __device__ __forceinline__ double lg(int i, int s, int a, double alpha) const {
    double v = L[(i * K + s) * n + a];
    if (TRIAL) v += alpha * R[s * b + OX + i * n + a];
    return v;
  }

  // [A|B] of stage s, player i at the last expansion point (DI: constant, never stored)
  __device__ __forceinline__ void loadAB(int s, int i, double A[16], double B[8]) const {
    if (model == AGB_MODEL_DOUBLE_INTEGRATOR) {
#pragma unroll
      for (int q = 0; q < 16; q++) A[q] = 0.0;
#pragma unroll
      for (int q = 0; q < 8; q++) B[q] = 0.0;
      A[0] = A[5] = A[10] = A[15] = 1.0; A[0 * 4 + 2] = dt; A[1 * 4 + 3] = dt;
      B[0 * 2 + 0] = dt * (dt / 2); B[1 * 2 + 1] = dt * (dt / 2); B[2 * 2 + 0] = dt; B[3 * 2 + 1] = dt;
    } else {
      const double* src = AB + (s * P + i) * 24;
#pragma unroll
      for (int q = 0; q < 16; q++) A[q] = src[q];
#pragma unroll
      for (int q = 0; q < 8; q++) B[q] = src[16 + q];
    }
  }
  __device__ __forceinline__ double Ael(int s, int i, int r, int c) const {   // A_i[r][c] of stage s
    if (model == AGB_MODEL_DOUBLE_INTEGRATOR) return (r == c) ? 1.0 : ((c == r + 2) ? dt : 0.0);
    return AB[(s * P + i) * 24 + r * 4 + c];
  }
  __device__ __forceinline__ double Bel(int s, int i, int r, int j) const {   // B_i[r][j] of stage s
    if (model == AGB_MODEL_DOUBLE_INTEGRATOR) return (r == j) ? dt * (dt / 2) : ((r == j + 2) ? dt : 0.0);
    return AB[(s * P + i) * 24 + 16 + r * 2 + j];
  }

  __device__ __forceinline__ int pair_index(int i, int j) const { return i * (P - 1) + (j < i ? j : j - 1); }

  // AL expansion of one inequality row (Altro cost_expansion!, pinned by test/constraints/constraint_derivatives.jl:28-34):
  // a = (c >= 0) | (λ > 0), w = a·μ, returns λ + w·c.
  __device__ __forceinline__ double al_row(int s, int row, double c, double& w) const {
    double lam = CL[s * nrow + row], mu = CM[s * nrow + row];
    w = ((c >= 0.0) || (lam > 0.0)) ? mu : 0.0;
    return lam + w * c;
  }

  // ---------------------------------------------------------------------------------------------------------
  // residual!  (problem/global_quantities.jl:9-65, constraints/constraint_derivatives.jl:39-74)
  // ---------------------------------------------------------------------------------------------------------
  // pass 1: per (stage, player): RK2 step and its Jacobian blocks; dynamics rows.
  template <bool TRIAL> __device__ void pass1(double alpha, double* Rout, Acc& acc) {
    for (int item = tid; item < K * P; item += kThreads) {
      int s = item / P, i = item - s * P;
      double st[4], u[2], xn[4], A[16], B[8];
#pragma unroll
      for (int c = 0; c < 4; c++) st[c] = xg<TRIAL>(s, c * P + i, alpha);
#pragma unroll
      for (int j = 0; j < 2; j++) u[j] = ug<TRIAL>(s, j * P + i, alpha);
      if (model == AGB_MODEL_DOUBLE_INTEGRATOR) {
        rk2_only(model, dt, d->lf, d->lr, st, u, xn);
      } else {
        rk2_jac(model, dt, d->lf, d->lr, st, u, xn, A, B);
        double* dst = AB + (s * P + i) * 24;
#pragma unroll
        for (int q = 0; q < 16; q++) dst[q] = A[q];
#pragma unroll
        for (int q = 0; q < 8; q++) dst[16 + q] = B[q];
      }
#pragma unroll
      for (int c = 0; c < 4; c++) {
        double r = xn[c] - xg<TRIAL>(s + 1, c * P + i, alpha);      // local_quantities.jl:13
        acc.sum += fabs(r); acc.dyn = fmax(acc.dyn, fabs(r));
        if (Rout) Rout[s * b + OD + c * P + i] = r;
      }
    }
  }

  // one element of player i's stationarity row block w.r.t. x at knot k (1..K), joint comp a
  template <bool TRIAL> __device__ double xrow_elem(int i, int k, int a, double alpha, double reg_x, Acc& acc) {
    const int c = a / P, ia = a - c * P, s = k - 1;
    const double dtx = (k < K) ? dt : 1.0;                      // terminal knot is not dt-scaled (objective test :52-64)
    const double xa = xg<TRIAL>(k, a, alpha);
    double v = 0.0;
    if (ia == i) v = dtx * Q[a] * (xa - xf[a]);                 // LQR gradient (objective.jl:24-32)
    if (c < 2) {
      for (int j = 0; j < P; j++) {
        if (j == i) continue;
        if (ia != i && ia != j) continue;
        const double dx = xg<TRIAL>(k, i, alpha) - xg<TRIAL>(k, j, alpha);
        const double dy = xg<TRIAL>(k, P + i, alpha) - xg<TRIAL>(k, P + j, alpha);
        const double dc = (c == 0) ? dx : dy;
        const double sgn = (ia == i) ? -1.0 : 1.0;
        const bool owner = (ia == i) && (c == 0);
        if (has_cc) {                                            // CollisionCost (objective.jl:134-173)
          const double dn = sqrt(dx * dx + dy * dy);
          const double rr = d->cc_radius[i], mu = d->cc_mu[i];
          double b00 = 0.0, b01 = 0.0, b11 = 0.0;
          if (fmax(0.0, rr - dn) > 0.0) {
            const double eps = 1e-10, eps_norm = eps * sqrt((double)n);
            const double g = mu * (rr * (eps + dc) / (eps_norm + dn) - dc);
            v += sgn * g * dtx;
            if (owner) {
              const double dn3 = dn * dn * dn;
              b00 = dtx * mu * (1.0 - rr / dn + rr * dx * dx / dn3);
              b01 = dtx * mu * (rr * dx * dy / dn3);
              b11 = dtx * mu * (1.0 - rr / dn + rr * dy * dy / dn3);
            }
          }
          if (owner) {
            double* cc = CC + (k * npairs + pair_index(i, j)) * 3;
            cc[0] = b00; cc[1] = b01; cc[2] = b11;
          }
        }
        const int row = d->col_row[i][j];
        if (row >= 0) {                                          // CollisionConstraint: c = r² − ‖xi−xj‖²
          const double rad = d->col_radius[i][j];
          const double cv = rad * rad - (dx * dx + dy * dy);
          double w; const double g = al_row(s, row, cv, w);
          v += sgn * 2.0 * dc * g;
          acc.sta = fmax(acc.sta, cv);
          if (owner) CW[s * nrow + row] = w;
        }
      }
      if (ia == i) {
        const double px = xg<TRIAL>(k, i, alpha), py = xg<TRIAL>(k, P + i, alpha);
        const int nw = d->n_walls[i];
        for (int q = 0; q < nw; q++) {                           // WallConstraint (constraints/wall_constraint.jl:56-89)
          const double* wl = d->walls[i][q];
          const double x1 = wl[0], y1 = wl[1], x2 = wl[2], y2 = wl[3], xv = wl[4], yv = wl[5];
          const bool left = (px - x1) * (x2 - x1) + (py - y1) * (y2 - y1) > 0.0;
          const bool right = (px - x2) * (x1 - x2) + (py - y2) * (y1 - y2) > 0.0;
          const double msk = (left && right) ? 1.0 : 0.0;
          const double cv = ((px - x1) * xv + (py - y1) * yv) * msk;
          const int row = d->wall_row[i] + q;
          double w; const double g = al_row(s, row, cv, w);
          v += msk * ((c == 0) ? xv : yv) * g;
          acc.sta = fmax(acc.sta, cv);
          if (c == 0) CW[s * nrow + row] = w * msk;              // effective weight: the wall Jacobian carries the mask
        }
        const int nc = d->n_circles[i];
        for (int q = 0; q < nc; q++) {                           // CircleConstraint: c = r² − (x−xc)² − (y−yc)²
          const double* cl = d->circles[i][q];
          const double ex = px - cl[0], ey = py - cl[1];
          const double cv = cl[2] * cl[2] - ex * ex - ey * ey;
          const int row = d->circle_row[i] + q;
          double w; const double g = al_row(s, row, cv, w);
          v += -2.0 * ((c == 0) ? ex : ey) * g;
          acc.sta = fmax(acc.sta, cv);
          if (c == 0) CW[s * nrow + row] = w;
        }
      }
    }
    {                                                            // StateBoundConstraint (state_bound_constraint.jl:80-92)
      int row = d->sbmax_row[i][a];
      if (row >= 0) {
        const double cv = xa - d->x_max[i][a];
        double w; const double g = al_row(s, row, cv, w);
        v += g; acc.sta = fmax(acc.sta, cv); CW[s * nrow + row] = w;
      }
      row = d->sbmin_row[i][a];
      if (row >= 0) {
        const double cv = d->x_min[i][a] - xa;
        double w; const double g = al_row(s, row, cv, w);
        v -= g; acc.sta = fmax(acc.sta, cv); CW[s * nrow + row] = w;
      }
    }
    if (k < K) {                                                 // + A_kᵀ λ_{i,k}   (global_quantities.jl:45-53)
#pragma unroll
      for (int q = 0; q < 4; q++) v += Ael(k, ia, q, c) * lg<TRIAL>(i, k, q * P + ia, alpha);
    }
    v -= lg<TRIAL>(i, k - 1, a, alpha);                          // − λ_{i,k−1}
    if (TRIAL) v += reg_x * (alpha * R[s * b + OD + a]);         // regularize_residual! (:67-86)
    return v;
  }

  // one element of player i's stationarity row block w.r.t. u_{i,s}: own control comp j
  template <bool TRIAL> __device__ double urow_elem(int i, int s, int j, double alpha, double reg_u, Acc& acc) {
    const int idx = j * P + i;
    const double ua = ug<TRIAL>(s, idx, alpha);
    double v = dt * Rw[idx] * (ua - uf[idx]);
#pragma unroll
    for (int q = 0; q < 4; q++) v += Bel(s, i, q, j) * lg<TRIAL>(i, s, q * P + i, alpha);
    int row = d->ub_row[idx];                                    // ControlBoundConstraint (control_bound_constraint.jl:94-106)
    if (row >= 0) {
      const double cv = ua - d->u_max[idx];
      double w; const double g = al_row(s, row, cv, w);
      v += g; acc.con = fmax(acc.con, cv); CW[s * nrow + row] = w;
    }
    row = d->lb_row[idx];
    if (row >= 0) {
      const double cv = d->u_min[idx] - ua;
      double w; const double g = al_row(s, row, cv, w);
      v -= g; acc.con = fmax(acc.con, cv); CW[s * nrow + row] = w;
    }
    if (TRIAL) v += reg_u * (alpha * R[s * b + OU + idx]);
    return v;
  }

  __device__ Acc block_reduce(Acc a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a.sum += __shfl_xor_sync(AGB_FULL, a.sum, o);
      a.opt = fmax(a.opt, __shfl_xor_sync(AGB_FULL, a.opt, o));
      a.dyn = fmax(a.dyn, __shfl_xor_sync(AGB_FULL, a.dyn, o));
      a.con = fmax(a.con, __shfl_xor_sync(AGB_FULL, a.con, o));
      a.sta = fmax(a.sta, __shfl_xor_sync(AGB_FULL, a.sta, o));
    }
    __syncthreads();            // red[] may still be read from the previous reduction
    if (lane == 0) { double* r = red + warp * 5; r[0] = a.sum; r[1] = a.opt; r[2] = a.dyn; r[3] = a.con; r[4] = a.sta; }
    __syncthreads();
    Acc t = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) {
      const double* r = red + w * 5;
      t.sum += r[0]; t.opt = fmax(t.opt, r[1]); t.dyn = fmax(t.dyn, r[2]); t.con = fmax(t.con, r[3]); t.sta = fmax(t.sta, r[4]);
    }
    return t;
  }

  // Full residual evaluation.  !TRIAL: at Z, rows stored to Rout (= R).  TRIAL: at Z + alpha·Δ (Δ in R) with the proximal
  // terms of regularize_residual!; rows go to Rout if non-null (may be global memory), norms are always returned.
  // NaN-safe maxima: fmax drops NaNs, so a non-finite residual is caught through `sum`.
  template <bool TRIAL> __device__ Acc residual(double alpha, double reg_x, double reg_u, double* Rout) {
    Acc acc = {0.0, 0.0, 0.0, 0.0, 0.0};
    pass1<TRIAL>(alpha, Rout, acc);
    __syncthreads();
    const int nx = P * K * n;
    for (int item = tid; item < nx; item += kThreads) {
      int a = item % n, t = item / n;
      int s = t % K, i = t / K;
      double v = xrow_elem<TRIAL>(i, s + 1, a, alpha, reg_x, acc);
      acc.sum += fabs(v); acc.opt = fmax(acc.opt, fabs(v));
      if (Rout) Rout[s * b + OX + i * n + a] = v;
    }
    const int nu = K * m;
    for (int item = tid; item < nu; item += kThreads) {
      int idx = item % m, s = item / m;
      int j = idx / P, i = idx - j * P;
      double v = urow_elem<TRIAL>(i, s, j, alpha, reg_u, acc);
      acc.sum += fabs(v); acc.opt = fmax(acc.opt, fabs(v));
      if (Rout) Rout[s * b + OU + idx] = v;
    }
    Acc t = block_reduce(acc);
    return t;
  }

  // ---------------------------------------------------------------------------------------------------------
  // residual_jacobian! blocks (global_quantities.jl:109-193) at the last expansion point (uses CW, CC, AB, X)
  // ---------------------------------------------------------------------------------------------------------
  // H^x_{i,k}[a][b] on the position sub-block a,b < 2p (collision cost + Gauss-Newton terms of position constraints)
  __device__ double hpos_entry(int i, int k, int a, int bq) const {
    const int ca = a / P, ia = a - ca * P, cb = bq / P, ib = bq - cb * P, s = k - 1;
    double v = 0.0;
    for (int j = 0; j < P; j++) {
      if (j == i) continue;
      if ((ia != i && ia != j) || (ib != i && ib != j)) continue;
      const double sg = (ia == ib) ? 1.0 : -1.0;
      if (has_cc) v += sg * CC[(k * npairs + pair_index(i, j)) * 3 + ca + cb];
      const int row = d->col_row[i][j];
      if (row >= 0) {
        const double w = CW[s * nrow + row];
        if (w != 0.0) {
          const double dx = X[k * n + i] - X[k * n + j], dy = X[k * n + P + i] - X[k * n + P + j];
          v += sg * 4.0 * (ca ? dy : dx) * (cb ? dy : dx) * w;
        }
      }
    }
    if (ia == i && ib == i) {
      const int nw = d->n_walls[i];
      for (int q = 0; q < nw; q++) {
        const double* wl = d->walls[i][q];
        v += CW[s * nrow + d->wall_row[i] + q] * (ca ? wl[5] : wl[4]) * (cb ? wl[5] : wl[4]);
      }
      const int nc = d->n_circles[i];
      if (nc > 0) {
        const double px = X[k * n + i], py = X[k * n + P + i];
        for (int q = 0; q < nc; q++) {
          const double* cl = d->circles[i][q];
          const double ex = px - cl[0], ey = py - cl[1];
          v += CW[s * nrow + d->circle_row[i] + q] * 4.0 * (ca ? ey : ex) * (cb ? ey : ex);
        }
      }
    }
    return v;
  }
  // diagonal part of H^x_{i,k}: dt·Q_i + state-bound terms + reg.x
  __device__ double hd_entry(int i, int k, int a, double reg_x) const {
    const int ia = a % P, s = k - 1;
    double v = reg_x;
    if (ia == i) v += ((k < K) ? dt : 1.0) * Q[a];
    int row = d->sbmax_row[i][a]; if (row >= 0) v += CW[s * nrow + row];
    row = d->sbmin_row[i][a];     if (row >= 0) v += CW[s * nrow + row];
    return v;
  }
  // H^u diagonal of stage s, joint control comp idx: dt·R + control-bound terms + reg.u
  __device__ double hu_entry(int s, int idx, double reg_u) const {
    double v = dt * Rw[idx] + reg_u;
    int row = d->ub_row[idx]; if (row >= 0) v += CW[s * nrow + row];
    row = d->lb_row[idx];     if (row >= 0) v += CW[s * nrow + row];
    return v;
  }

  // assemble compact H^x_{·,k} into Hpos/Hd with the threads [t0, t0+nt) of the CTA
  __device__ void assemble_H(int k, double reg_x, int t0, int nt) {
    const int t = tid - t0;
    if (t < 0 || t >= nt) return;
    constexpr int q2 = 2 * P;
    for (int item = t; item < P * q2 * q2; item += nt) {
      int bq = item % q2, r = item / q2;
      int a = r % q2, i = r / q2;
      Hpos[item] = hpos_entry(i, k, a, bq);
    }
    for (int item = t; item < P * n; item += nt) {
      int a = item % n, i = item / n;
      Hd[item] = hd_entry(i, k, a, reg_x);
    }
  }
  __device__ __forceinline__ double Hc(int i, int a, int bq) const {
    double v = (a == bq) ? Hd[i * n + a] : 0.0;
    if (a < 2 * P && bq < 2 * P) v += Hpos[(i * 2 * P + a) * 2 * P + bq];
    return v;
  }

  // Gauss-Jordan with partial pivoting on the m x W augmented system, one column per lane (warp 0).
  __device__ bool gj_warp(double* aug) {
    double a[m];
    const bool act = lane < W;
#pragma unroll
    for (int r = 0; r < m; r++) a[r] = act ? aug[r * W + lane] : 0.0;
    bool ok = true;
#pragma unroll
    for (int t = 0; t < m; t++) {
      int pr = t; double best = fabs(a[t]);
#pragma unroll
      for (int r = t + 1; r < m; r++) { double v = fabs(a[r]); if (v > best) { best = v; pr = r; } }
      pr = __shfl_sync(AGB_FULL, pr, t);
#pragma unroll
      for (int r = t + 1; r < m; r++) if (r == pr) { double tmp = a[r]; a[r] = a[t]; a[t] = tmp; }
      const double pv = __shfl_sync(AGB_FULL, a[t], t);
      if (!(fabs(pv) > 0.0) || isinf(pv)) ok = false;
      const double at = a[t] * (1.0 / pv);
#pragma unroll
      for (int r = 0; r < m; r++) {
        if (r != t) { const double f = __shfl_sync(AGB_FULL, a[r], t); a[r] = fma(-f, at, a[r]); }
      }
      a[t] = at;
    }
    if (act) {
#pragma unroll
      for (int r = 0; r < m; r++) aug[r * W + lane] = a[r];
    }
    return ok;
  }

  // ---------------------------------------------------------------------------------------------------------
  // Δtraj = −(lu(jac) \ res)  (solver_methods.jl:87-88): R holds res on entry, Δ on exit (same stage-major slots:
  // rx(i,s) → Δλ_{i,s}, ru(s) → Δu_s, rd(s) → Δx_{s+1}).  Returns false on a singular / non-finite pivot.
  // ---------------------------------------------------------------------------------------------------------
  __device__ bool kkt_solve(double reg_x, double reg_u) {
    int ok = 1;
    double* Pc = Pm;                 // current P_i (knot s+1), [P][n][n]
    double* Pn = Pm + P * n * n;     // next
    double* sc = Sv;                 // [P][n]
    double* sn = Sv + P * n;
    // terminal knot: P_i = H_{i,N}, s_i = r^x_{i,N}
    assemble_H(K, reg_x, 0, kThreads);
    __syncthreads();
    for (int item = tid; item < P * n * n; item += kThreads) {
      int bq = item % n, r = item / n;
      int a = r % n, i = r / n;
      Pc[item] = Hc(i, a, bq);
    }
    for (int item = tid; item < P * n; item += kThreads) {
      int a = item % n, i = item / n;
      sc[item] = R[(K - 1) * b + OX + i * n + a];
    }
    __syncthreads();
    for (int s = K - 1; s >= 0; s--) {
      const double* Rs = R + s * b;
      // ---- phase B: augmented system  [Hu + Y B | Y A | Y rd + Bᵀ s + ru],  Y_r = B_iᵀ P_i (row r = (j,i))
      for (int item = tid; item < m * W; item += kThreads) {
        const int col = item % W, r = item / W;
        const int j = r / P, i = r - j * P;
        const double* Pi = Pc + i * n * n;
        double Bi[4];
#pragma unroll
        for (int q = 0; q < 4; q++) Bi[q] = Bel(s, i, q, j);
        double v;
        if (col < m + n) {
          int i2, c2; bool isB = col < m;
          if (isB) { c2 = col / P; i2 = col - c2 * P; } else { int a2 = col - m; c2 = a2 / P; i2 = a2 - c2 * P; }
          v = 0.0;
#pragma unroll
          for (int q = 0; q < 4; q++) {                 // Y[r][(q,i2)]
            double y = 0.0;
#pragma unroll
            for (int q1 = 0; q1 < 4; q1++) y += Bi[q1] * Pi[(q1 * P + i) * n + q * P + i2];
            v += y * (isB ? Bel(s, i2, q, c2) : Ael(s, i2, q, c2));
          }
          if (isB && col == r) v += hu_entry(s, r, reg_u);
        } else {
          v = Rs[OU + r];
#pragma unroll
          for (int q1 = 0; q1 < 4; q1++) {
            const double* prow = Pi + (q1 * P + i) * n;
            double y = sc[i * n + q1 * P + i];
            for (int a2 = 0; a2 < n; a2++) y += prow[a2] * Rs[OD + a2];
            v += Bi[q1] * y;
          }
        }
        Aug[r * W + col] = v;
      }
      __syncthreads();
      // ---- phase C: warp 0 solves for the gains; the other warps assemble H^x of knot s meanwhile
      if (warp == 0) {
        if (!gj_warp(Aug)) ok = 0;
        if (lane >= m && lane < W) {
          double* ku = KU + s * KUS;
#pragma unroll
          for (int r = 0; r < m; r++) ku[r * (n + 1) + (lane - m)] = Aug[r * W + lane];
        }
      } else if (s > 0) {
        assemble_H(s, reg_x, 32, kThreads - 32);
      }
      __syncthreads();
      if (s == 0) break;
      const double* ku = KU + s * KUS;
      // ---- phase D: closed loop  Acl = A − B Ku,  ccl = rd − B ku   (n x (n+1))
      for (int item = tid; item < n * (n + 1); item += kThreads) {
        const int col = item % (n + 1), a = item / (n + 1);
        const int c = a / P, i = a - c * P;
        double v;
        if (col < n) { const int c2 = col / P, i2 = col - c2 * P; v = (i2 == i) ? Ael(s, i, c, c2) : 0.0; }
        else v = Rs[OD + a];
#pragma unroll
        for (int j = 0; j < 2; j++) v -= Bel(s, i, c, j) * ku[(j * P + i) * (n + 1) + col];
        Acl[item] = v;
      }
      __syncthreads();
      // ---- phase E: P_i ← H_{i,s} + Aᵀ P_i Acl,  s_i ← r^x_{i,s} + Aᵀ (P_i ccl + s_i)
      for (int item = tid; item < P * P * (n + 1); item += kThreads) {
        const int col = item % (n + 1), t = item / (n + 1);
        const int i2 = t % P, i = t / P;
        const double* Pi = Pc + i * n * n;
        double T[4];
#pragma unroll
        for (int q = 0; q < 4; q++) T[q] = (col == n) ? sc[i * n + q * P + i2] : 0.0;
        for (int a2 = 0; a2 < n; a2++) {
          const double e = Acl[a2 * (n + 1) + col];
#pragma unroll
          for (int q = 0; q < 4; q++) T[q] += Pi[(q * P + i2) * n + a2] * e;
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
          double o = 0.0;
#pragma unroll
          for (int q = 0; q < 4; q++) o += Ael(s, i2, q, c) * T[q];
          const int a = c * P + i2;
          if (col < n) Pn[(i * n + a) * n + col] = o + Hc(i, a, col);
          else sn[i * n + a] = o + R[(s - 1) * b + OX + i * n + a];
        }
      }
      __syncthreads();
      { double* t = Pc; Pc = Pn; Pn = t; t = sc; sc = sn; sn = t; }
    }
    // ---- forward sweep (warp 0): Δu_s = −Ku Δx_s − ku,  Δx_{s+1} = A Δx_s + B Δu_s + rd
    if (warp == 0) {
      for (int s = 0; s < K; s++) {
        const double* ku = KU + s * KUS;
        double* Rs = R + s * b;
        const double* dxp = R + (s - 1) * b + OD;      // Δx_s (only read when s > 0)
        double du = 0.0;
        if (lane < m) {
          double acc = ku[lane * (n + 1) + n];
          if (s > 0) for (int a = 0; a < n; a++) acc += ku[lane * (n + 1) + a] * dxp[a];
          du = -acc;
          Rs[OU + lane] = du;
        }
        __syncwarp();
        double v = 0.0;
        if (lane < n) {
          const int c = lane / P, i = lane - c * P;
          v = Rs[OD + lane];
          if (s > 0) {
#pragma unroll
            for (int q = 0; q < 4; q++) v += Ael(s, i, c, q) * dxp[q * P + i];
          }
#pragma unroll
          for (int j = 0; j < 2; j++) v += Bel(s, i, c, j) * Rs[OU + j * P + i];
        }
        __syncwarp();
        if (lane < n) Rs[OD + lane] = v;
        __syncwarp();
      }
    }
    __syncthreads();
    // ---- costate: g_{i,k} = H_{i,k} Δx_k + r^x_{i,k} (parallel), then Δλ_{i,k−1} = g_{i,k} + A_kᵀ Δλ_{i,k} (per-player warp)
    for (int item = tid; item < P * K * n; item += kThreads) {
      const int a = item % n, t = item / n;
      const int s = t % K, i = t / K, k = s + 1;
      const double* dx = R + s * b + OD;
      double v = R[s * b + OX + i * n + a] + hd_entry(i, k, a, reg_x) * dx[a];
      if (a < 2 * P) {
        for (int bq = 0; bq < 2 * P; bq++) v += hpos_entry(i, k, a, bq) * dx[bq];
      }
      R[s * b + OX + i * n + a] = v;
    }
    __syncthreads();
    if (warp < P) {
      const int i = warp;
      for (int s = K - 2; s >= 0; s--) {
        double v = 0.0;
        if (lane < n) {
          const int c = lane / P, ia = lane - c * P;
          v = R[s * b + OX + i * n + lane];
#pragma unroll
          for (int q = 0; q < 4; q++) v += Ael(s + 1, ia, q, c) * R[(s + 1) * b + OX + i * n + q * P + ia];
        }
        __syncwarp();
        if (lane < n) R[s * b + OX + i * n + lane] = v;
        __syncwarp();
      }
    }
    __syncthreads();
    ok = __syncthreads_and(ok);
    return ok != 0;
  }

  // line_search (solver_methods.jl:105-125): returns alpha, j through references; n_eval counts residual evaluations
  __device__ void line_search(const agb_options& o, double reg, double res_norm, double& alpha, int& j, int& n_eval) {
    const double S = (double)(K * b);
    const double rr = o.regularize ? reg : 0.0;
    alpha = 1.0; j = 1;
    while (j < o.ls_iter) {
      Acc t = residual<true>(alpha, rr, rr, nullptr);
      n_eval++;
      const double trial = t.sum / S;
      if (trial <= (1.0 - alpha * o.beta) * res_norm) break;
      alpha *= o.alpha_decrease;
      j++;
    }
  }

  // update_traj!(pdtraj, pdtraj, α, Δpdtraj) + Δ_step (primal_dual_traj.jl:109-147)
  __device__ double update_traj(double alpha) {
    double loc = 0.0;
    for (int item = tid; item < K * b; item += kThreads) {
      const int q = item % b, s = item / b;
      const double dv = R[item];
      if (q < OU) { const int i = q / n, a = q - i * n; L[(i * K + s) * n + a] += alpha * dv; }
      else if (q < OD) { U[s * m + (q - OU)] += alpha * dv; loc += fabs(dv); }
      else { X[(s + 1) * n + (q - OD)] += alpha * dv; loc += fabs(dv); }
    }
    Acc a = {loc, 0.0, 0.0, 0.0, 0.0};
    a = block_reduce(a);
    return a.sum * alpha / (double)(K * (n + m));
  }

  // rollout!(RK3, model, traj): independent per player (separable dynamics)
  __device__ void rollout() {
    if (tid < P) {
      const int i = tid;
      double st[4], u[2], xn[4];
#pragma unroll
      for (int c = 0; c < 4; c++) st[c] = X[c * P + i];
      for (int s = 0; s < K; s++) {
        u[0] = U[s * m + i]; u[1] = U[s * m + P + i];
        rk3_step(model, dt, d->lf, d->lr, st, u, xn);
#pragma unroll
        for (int c = 0; c < 4; c++) { st[c] = xn[c]; X[(s + 1) * n + c * P + i] = xn[c]; }
      }
    }
    __syncthreads();
  }

  // value of AL constraint row `row` of stage s at the resident iterate (evaluate!, constraints_methods.jl:367-379)
  __device__ double con_value(int s, int row) const {
    const int k = s + 1;
    if (row >= d->nrow_state) {
      for (int idx = 0; idx < m; idx++) {
        if (d->ub_row[idx] == row) return U[s * m + idx] - d->u_max[idx];
        if (d->lb_row[idx] == row) return d->u_min[idx] - U[s * m + idx];
      }
      return 0.0;
    }
    int i = 0;
    while (i + 1 < P && row >= d->srow_off[i + 1]) i++;
    for (int j = 0; j < P; j++) {
      if (j != i && d->col_row[i][j] == row) {
        const double dx = X[k * n + i] - X[k * n + j], dy = X[k * n + P + i] - X[k * n + P + j];
        const double rad = d->col_radius[i][j];
        return rad * rad - (dx * dx + dy * dy);
      }
    }
    for (int a = 0; a < n; a++) {
      if (d->sbmax_row[i][a] == row) return X[k * n + a] - d->x_max[i][a];
      if (d->sbmin_row[i][a] == row) return d->x_min[i][a] - X[k * n + a];
    }
    const double px = X[k * n + i], py = X[k * n + P + i];
    if (row >= d->wall_row[i] && row < d->wall_row[i] + d->n_walls[i]) {
      const double* wl = d->walls[i][row - d->wall_row[i]];
      const bool left = (px - wl[0]) * (wl[2] - wl[0]) + (py - wl[1]) * (wl[3] - wl[1]) > 0.0;
      const bool right = (px - wl[2]) * (wl[0] - wl[2]) + (py - wl[3]) * (wl[1] - wl[3]) > 0.0;
      return ((px - wl[0]) * wl[4] + (py - wl[1]) * wl[5]) * ((left && right) ? 1.0 : 0.0);
    }
    if (row >= d->circle_row[i] && row < d->circle_row[i] + d->n_circles[i]) {
      const double* cl = d->circles[i][row - d->circle_row[i]];
      const double ex = px - cl[0], ey = py - cl[1];
      return cl[2] * cl[2] - ex * ex - ey * ey;
    }
    return 0.0;
  }
  __device__ __forceinline__ int row_player(int row) const {     // owner of a state row, -1 for control rows
    if (row >= d->nrow_state) return -1;
    int i = 0;
    while (i + 1 < P && row >= d->srow_off[i + 1]) i++;
    return i;
  }

  // evaluate! + dual_update! (constraints_methods.jl:349-365, 421-440): λ ← clamp(λ + α·μ∘c, 0, λ_max)
  __device__ void dual_update(const agb_options& o) {
    for (int item = tid; item < K * nrow; item += kThreads) {
      const int row = item % nrow, s = item / nrow;
      const int i = row_player(row);
      const double a = (i >= 0) ? o.alphax_dual[i] : o.alpha_dual;
      const double c = con_value(s, row);
      CL[item] = fmin(fmax(CL[item] + a * CM[item] * c, 0.0), o.lambda_max);
    }
    __syncthreads();
  }
  // penalty_update! → Altro.penalty_update!: μ ← clamp(ϕ·μ, 0, μ_max) (pinned by test/constraints/constraints_methods.jl:176-193)
  __device__ void penalty_update(const agb_options& o) {
    for (int item = tid; item < K * nrow; item += kThreads) CM[item] = fmin(fmax(o.rho_increase * CM[item], 0.0), o.rho_max);
    __syncthreads();
  }
  __device__ void reset_duals_penalties(const agb_options& o) {  // reset!(game_con) (constraints_methods.jl:295-327)
    for (int item = tid; item < K * nrow; item += kThreads) { CL[item] = 0.0; CM[item] = o.rho_0; }
    __syncthreads();
  }

  // ---- global <-> shared staging ---------------------------------------------------------------------------
  __device__ void load_params(const Buffers& g, int inst) {
    for (int a = tid; a < n; a += kThreads) { xf[a] = g.xf[(size_t)inst * n + a]; Q[a] = g.Q[(size_t)inst * n + a]; }
    for (int a = tid; a < m; a += kThreads) { Rw[a] = g.R[(size_t)inst * m + a]; uf[a] = g.uf[(size_t)inst * m + a]; }
  }
  __device__ void load_iterate(const double* Zg, const double* Lg, int inst) {
    const double* z = Zg + (size_t)inst * N * (n + m);
    for (int item = tid; item < N * (n + m); item += kThreads) {
      const int q = item % (n + m), k = item / (n + m);
      if (q < n) X[k * n + q] = z[item]; else U[k * m + (q - n)] = z[item];
    }
    const double* l = Lg + (size_t)inst * P * K * n;
    for (int item = tid; item < P * K * n; item += kThreads) L[item] = l[item];
  }
  __device__ void store_iterate(double* Zg, double* Lg, int inst) const {
    double* z = Zg + (size_t)inst * N * (n + m);
    for (int item = tid; item < N * (n + m); item += kThreads) {
      const int q = item % (n + m), k = item / (n + m);
      z[item] = (q < n) ? X[k * n + q] : U[k * m + (q - n)];
    }
    double* l = Lg + (size_t)inst * P * K * n;
    for (int item = tid; item < P * K * n; item += kThreads) l[item] = L[item];
  }
  __device__ void load_duals(const Buffers& g, int inst) {
    const size_t o = (size_t)inst * K * nrow;
    for (int item = tid; item < K * nrow; item += kThreads) { CL[item] = g.conlam[o + item]; CM[item] = g.conmu[o + item]; }
  }
  __device__ void store_duals(const Buffers& g, int inst) const {
    const size_t o = (size_t)inst * K * nrow;
    for (int item = tid; item < K * nrow; item += kThreads) { g.conlam[o + item] = CL[item]; g.conmu[o + item] = CM[item]; }
  }
};

}  // namespace agb
