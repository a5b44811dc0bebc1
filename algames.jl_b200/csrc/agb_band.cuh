// agb_band.cuh — the BAND solver: newton_solve! (src/problem/solver_methods.jl:5-125) on the explicit KKT band.
//
// The structured kernels (agb_solver.cuh) eliminate the KKT system stage by stage and are specialised for planar players
// with 4 states / 2 controls.  This file is the general form of the same path, written for every schema the reference
// accepts: per-player dynamics of any size (QuadrotorGame: 12 states, 4 rotor commands, dynamics/quadrotor.jl), planar and
// 3-D constraints (Wall3D, cylinders, spherical collision avoidance; constraints/wall_constraint.jl:141-249,
// cylinder_constraint.jl:33-137, constraints_methods.jl:45-81), and — like the reference's lu(jac) \ res
// (solver_methods.jl:87) — a factorisation with row pivoting over the WHOLE matrix, so it also serves as the fallback for
// instances on which the structured elimination meets a singular stage system although the KKT matrix is regular.
//
// One CTA per instance.  The KKT Jacobian lives in global memory as a band in time-major order — rows
// [dyn_s | opt u_s | opt x_{s+1}], columns [λ_s | u_s | x_{s+1}] per stage, bandwidths kl = max(2n−1, n+m),
// ku = max(p·n+n−1, p·n+m) — and is factorised in place by an LU with partial pivoting (row exchanges inside the band, fill
// kept in the extra kl columns of every stored row); every row of the band is assembled by exactly one work item, so no
// atomics are needed.  Discrete Jacobians [A | B] of the RK2 map come from forward-mode differentiation (dual numbers),
// which is what the reference does with ForwardDiff (problem/local_quantities.jl:26).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <type_traits>
#include "agb_internal.h"

namespace agb {
namespace band {

// ------------------------------------------------------------------------------------------------------------
// forward-mode scalar: value + one directional derivative
// ------------------------------------------------------------------------------------------------------------
struct dual { double v, d; };
__device__ __forceinline__ dual mk(double v, double d = 0.0) { dual r; r.v = v; r.d = d; return r; }
__device__ __forceinline__ dual operator+(dual a, dual b) { return mk(a.v + b.v, a.d + b.d); }
__device__ __forceinline__ dual operator-(dual a, dual b) { return mk(a.v - b.v, a.d - b.d); }
__device__ __forceinline__ dual operator*(dual a, dual b) { return mk(a.v * b.v, a.d * b.v + a.v * b.d); }
__device__ __forceinline__ dual operator/(dual a, dual b) { const double q = a.v / b.v; return mk(q, (a.d - q * b.d) / b.v); }
__device__ __forceinline__ dual operator+(dual a, double b) { return mk(a.v + b, a.d); }
__device__ __forceinline__ dual operator+(double a, dual b) { return mk(a + b.v, b.d); }
__device__ __forceinline__ dual operator-(dual a, double b) { return mk(a.v - b, a.d); }
__device__ __forceinline__ dual operator-(double a, dual b) { return mk(a - b.v, -b.d); }
__device__ __forceinline__ dual operator-(dual a) { return mk(-a.v, -a.d); }
__device__ __forceinline__ dual operator*(dual a, double b) { return mk(a.v * b, a.d * b); }
__device__ __forceinline__ dual operator*(double a, dual b) { return mk(a * b.v, a * b.d); }
__device__ __forceinline__ dual operator/(dual a, double b) { return mk(a.v / b, a.d / b); }
__device__ __forceinline__ dual sin_(dual a) { double s, c; sincos(a.v, &s, &c); return mk(s, c * a.d); }
__device__ __forceinline__ dual cos_(dual a) { double s, c; sincos(a.v, &s, &c); return mk(c, -s * a.d); }
__device__ __forceinline__ dual tan_(dual a) { const double t = tan(a.v); return mk(t, (1.0 + t * t) * a.d); }
__device__ __forceinline__ dual atan2c_(dual y, double x) { return mk(atan2(y.v, x), x / (x * x + y.v * y.v) * y.d); }   // constant 2nd argument
__device__ __forceinline__ dual max0_(dual a) { return a.v > 0.0 ? a : mk(0.0, 0.0); }
__device__ __forceinline__ double sin_(double a) { return sin(a); }
__device__ __forceinline__ double cos_(double a) { return cos(a); }
__device__ __forceinline__ double tan_(double a) { return tan(a); }
__device__ __forceinline__ double atan2c_(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ double max0_(double a) { return a > 0.0 ? a : 0.0; }

constexpr int kMaxNi = 12, kMaxMi = 4;

// per-player continuous dynamics ẋ_i = f(x_i, u_i) (all models are separable per player)
template <class T> __device__ void dyn_f(const DevDesc* d, const T* s, const T* u, T* f) {
  switch (d->model) {
    case AGB_MODEL_DOUBLE_INTEGRATOR: {                    // dynamics/double_integrator.jl:27-31, any dimension d = mi
      const int dd = d->mi;
      for (int c = 0; c < dd; c++) { f[c] = s[dd + c]; f[dd + c] = u[c]; }
    } break;
    case AGB_MODEL_UNICYCLE:                               // dynamics/unicycle.jl:27-32
      f[0] = cos_(s[2]) * s[3]; f[1] = sin_(s[2]) * s[3]; f[2] = u[0]; f[3] = u[1];
      break;
    case AGB_MODEL_BICYCLE: {                              // dynamics/bicycle.jl:28-41
      const T beta = atan2c_(d->lr * tan_(u[1]), d->lr + d->lf);
      f[0] = s[2] * cos_(beta + s[3]); f[1] = s[2] * sin_(beta + s[3]); f[2] = u[0]; f[3] = s[2] * sin_(beta) / d->lr;
    } break;
    default: {                                             // dynamics/quadrotor.jl:48-119: [r, q (MRP), v, ω]
      const double mass = d->quad_mass, J0 = 0.0023, J1 = 0.0023, J2 = 0.004, Lm = 0.1750, kf = 1.245, km = 1.0, grav = -9.81;
      const T q0 = s[3], q1 = s[4], q2 = s[5], w0 = s[9], w1 = s[10], w2 = s[11];
      const T F0 = max0_(kf * u[0]), F1 = max0_(kf * u[1]), F2 = max0_(kf * u[2]), F3 = max0_(kf * u[3]);
      const T Ft = ((F0 + F1) + F2) + F3;
      // third column of RotMatrix(MRP(q)) = I + (8 S² + 4 (1 − |q|²) S) / (1 + |q|²)²  [Rotations.jl]
      const T n2 = (q0 * q0 + q1 * q1) + q2 * q2;
      const T den = (1.0 + n2) * (1.0 + n2), c4 = 4.0 * (1.0 - n2);
      const T r0 = (8.0 * (q0 * q2) + c4 * q1) / den;
      const T r1 = (8.0 * (q1 * q2) - c4 * q0) / den;
      const T r2 = 1.0 - (8.0 * (q0 * q0 + q1 * q1)) / den;
      f[0] = s[6]; f[1] = s[7]; f[2] = s[8];
      // Rotations.kinematics(MRP(q), ω) = ¼ ((1 − |q|²) ω + 2 q × ω + 2 q (q·ω))
      const T qw = (q0 * w0 + q1 * w1) + q2 * w2, om = 1.0 - n2;
      f[3] = 0.25 * ((om * w0 + 2.0 * (q1 * w2 - q2 * w1)) + 2.0 * (q0 * qw));
      f[4] = 0.25 * ((om * w1 + 2.0 * (q2 * w0 - q0 * w2)) + 2.0 * (q1 * qw));
      f[5] = 0.25 * ((om * w2 + 2.0 * (q0 * w1 - q1 * w0)) + 2.0 * (q2 * qw));
      f[6] = (r0 * Ft) / mass; f[7] = (r1 * Ft) / mass; f[8] = (mass * grav + r2 * Ft) / mass;
      const T t0 = Lm * (F1 - F3), t1 = Lm * (F2 - F0), t2 = km * (((u[0] - u[1]) + u[2]) - u[3]);
      f[9] = (t0 - (w1 * (J2 * w2) - w2 * (J1 * w1))) / J0;
      f[10] = (t1 - (w2 * (J0 * w0) - w0 * (J2 * w2))) / J1;
      f[11] = (t2 - (w0 * (J1 * w1) - w1 * (J0 * w0))) / J2;
    } break;
  }
}

// discrete_dynamics(RK2,…): explicit midpoint (problem/local_quantities.jl:13)
template <class T> __device__ void rk2(const DevDesc* d, const T* s, const T* u, T* xn) {
  const int ni = d->ni;
  T f0[kMaxNi], sm[kMaxNi], f1[kMaxNi];
  dyn_f(d, s, u, f0);
  for (int c = 0; c < ni; c++) sm[c] = s[c] + (f0[c] * d->dt) / 2.0;
  dyn_f(d, sm, u, f1);
  for (int c = 0; c < ni; c++) xn[c] = s[c] + f1[c] * d->dt;
}
// discrete_dynamics(RK3,…) of rollout! (solver_methods.jl:17-18)
__device__ inline void rk3(const DevDesc* d, const double* s, const double* u, double* xn) {
  const int ni = d->ni;
  const double dt = d->dt;
  double k1[kMaxNi], k2[kMaxNi], k3[kMaxNi], t[kMaxNi];
  dyn_f(d, s, u, k1);
  for (int c = 0; c < ni; c++) { k1[c] *= dt; t[c] = s[c] + k1[c] / 2; }
  dyn_f(d, t, u, k2);
  for (int c = 0; c < ni; c++) { k2[c] *= dt; t[c] = s[c] - k1[c] + 2 * k2[c]; }
  dyn_f(d, t, u, k3);
  for (int c = 0; c < ni; c++) { k3[c] *= dt; xn[c] = s[c] + (k1[c] + 4 * k2[c] + k3[c]) / 6; }
}

struct Norms { double sum, opt, dyn, con, sta; };

// -DAGB_BAND_TIMING (profiling builds of agb_band.cu only): thread 0 of CTA 0 accumulates clock64() per phase, printed at kernel end
#ifdef AGB_BAND_TIMING
__device__ long long g_bt[16];
#define BT_MARK(slot) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long t_ = clock64(); g_bt[slot] += t_ - bt_t0; bt_t0 = t_; } } while (0)
#define BT_START() long long bt_t0 = clock64()
#else
#define BT_MARK(slot) do { } while (0)
#define BT_START() do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------------------
// one instance
// ------------------------------------------------------------------------------------------------------------
struct Ctx {
  const DevDesc* d;
  int p, n, m, ni, mi, N, K, b, S, nrow, kl, ku, wd, nab;
  int tid, nt;
  double dt;
  const double *xf, *Q, *R, *uf;                     // this instance's objective (joint, component-major)
  double *X, *U, *L, *Xt, *Ut, *Lt, *dX, *dU, *dL, *res, *rhs, *AB, *XN, *bandm;
  double *lam, *mu;                                   // [K][nrow], this instance's slice of the result buffers
  double* red;                                        // shared: 8 doubles per warp + broadcast slots
  double* win = nullptr;                              // shared-memory elimination window (band_solve_window), nullptr: global path

  __device__ void bind(const DevDesc* dd, double* scratch, double* red_) {
    d = dd; p = dd->p; n = dd->n; m = dd->m; ni = dd->ni; mi = dd->mi; N = dd->N; K = dd->K; b = dd->b; S = dd->S; nrow = dd->nrow;
    kl = dd->kl; ku = dd->ku; wd = dd->wd; nab = ni * (ni + mi); dt = dd->dt;
    tid = threadIdx.x; nt = blockDim.x; red = red_;
    double* q = scratch;
    X = q; q += N * n; U = q; q += N * m; L = q; q += p * K * n;
    Xt = q; q += N * n; Ut = q; q += N * m; Lt = q; q += p * K * n;
    dX = q; q += N * n; dU = q; q += N * m; dL = q; q += p * K * n;
    res = q; q += S; rhs = q; q += S; AB = q; q += K * p * nab; XN = q; q += K * n;
    q += ((q - scratch) & 1);                         // 16-byte alignment of the band rows is not required, keep it even anyway
    bandm = q;
  }
  // time-major numbering
  __device__ __forceinline__ int row_dyn(int s, int a) const { return s * b + a; }
  __device__ __forceinline__ int row_u(int s, int idx) const { return s * b + n + idx; }
  __device__ __forceinline__ int row_x(int s, int i, int a) const { return s * b + n + m + i * n + a; }
  __device__ __forceinline__ int col_l(int s, int i, int a) const { return s * b + i * n + a; }
  __device__ __forceinline__ int col_u(int s, int idx) const { return s * b + p * n + idx; }
  __device__ __forceinline__ int col_x(int s, int a) const { return s * b + p * n + m + a; }
  __device__ __forceinline__ double& be(int r, int c) const { return bandm[(size_t)r * wd + (c - r + kl)]; }

  // ---- block reduction of (sum, max, max, max, max) ----------------------------------------------------------
  __device__ Norms reduce(Norms a) {
    for (int o = 16; o > 0; o >>= 1) {
      a.sum += __shfl_xor_sync(0xffffffffu, a.sum, o);
      a.opt = fmax(a.opt, __shfl_xor_sync(0xffffffffu, a.opt, o)); a.dyn = fmax(a.dyn, __shfl_xor_sync(0xffffffffu, a.dyn, o));
      a.con = fmax(a.con, __shfl_xor_sync(0xffffffffu, a.con, o)); a.sta = fmax(a.sta, __shfl_xor_sync(0xffffffffu, a.sta, o));
    }
    __syncthreads();
    if ((tid & 31) == 0) { double* r = red + (tid >> 5) * 5; r[0] = a.sum; r[1] = a.opt; r[2] = a.dyn; r[3] = a.con; r[4] = a.sta; }
    __syncthreads();
    Norms t = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int w = 0; w < nt / 32; w++) {
      const double* r = red + w * 5;
      t.sum += r[0]; t.opt = fmax(t.opt, r[1]); t.dyn = fmax(t.dyn, r[2]); t.con = fmax(t.con, r[3]); t.sta = fmax(t.sta, r[4]);
    }
    return t;
  }

  // ---- constraint rows of player i at knot s+1 / control rows of stage s: F(row, c, nnz, idx, g) ---------------
  // canonical order: [collision j≠i ascending | state bounds conval by conval (max rows, min rows) | walls | circles |
  // 3-D walls | cylinders], then the control-bound rows [u_max rows | u_min rows]
  template <class F> __device__ void state_rows(int i, const double* x, F&& fn) const {
    int r = d->srow_off[i];
    int idx[6]; double g[6];
    for (int j = 0; j < p; j++) {
      if (j == i || !(d->col_radius[i][j] > 0.0)) continue;
      const int nd = d->spherical ? 3 : 2;             // CollisionConstraint on px (planar) or on the first three components
      double d2 = 0.0;
      for (int c = 0; c < nd; c++) {
        const double dd = x[c * p + i] - x[c * p + j];
        d2 += dd * dd; idx[c] = c * p + i; g[c] = -2.0 * dd; idx[nd + c] = c * p + j; g[nd + c] = 2.0 * dd;
      }
      const double rad = d->col_radius[i][j];
      fn(r++, rad * rad - d2, 2 * nd, idx, g);
    }
    for (int q = 0; q < d->sb_ncon[i]; q++) {          // StateBoundConstraint convals in the order they were added
      for (int a = 0; a < n; a++) if (d->has_x_max[i][a] && d->x_max_con[i][a] == q) { idx[0] = a; g[0] = 1.0; fn(r++, x[a] - d->x_max[i][a], 1, idx, g); }
      for (int a = 0; a < n; a++) if (d->has_x_min[i][a] && d->x_min_con[i][a] == q) { idx[0] = a; g[0] = -1.0; fn(r++, d->x_min[i][a] - x[a], 1, idx, g); }
    }
    const double px = x[i], py = x[p + i];
    for (int t = 0; t < d->n_walls[i]; t++) {           // WallConstraint (constraints/wall_constraint.jl:56-89)
      const double* wl = d->walls[i][t];
      const bool left = (px - wl[0]) * (wl[2] - wl[0]) + (py - wl[1]) * (wl[3] - wl[1]) > 0.0;
      const bool right = (px - wl[2]) * (wl[0] - wl[2]) + (py - wl[3]) * (wl[1] - wl[3]) > 0.0;
      const double msk = (left && right) ? 1.0 : 0.0;
      idx[0] = i; g[0] = msk * wl[4]; idx[1] = p + i; g[1] = msk * wl[5];
      fn(r++, ((px - wl[0]) * wl[4] + (py - wl[1]) * wl[5]) * msk, 2, idx, g);
    }
    for (int t = 0; t < d->n_circles[i]; t++) {         // CircleConstraint
      const double* cl = d->circles[i][t];
      const double ex = px - cl[0], ey = py - cl[1];
      idx[0] = i; g[0] = -2.0 * ex; idx[1] = p + i; g[1] = -2.0 * ey;
      fn(r++, cl[2] * cl[2] - ex * ex - ey * ey, 2, idx, g);
    }
    if (d->n_walls3d[i] > 0 || d->n_cyl[i] > 0) {
      const double pz = x[2 * p + i];
      for (int t = 0; t < d->n_walls3d[i]; t++) {       // Wall3DConstraint (constraints/wall_constraint.jl:141-233)
        const double* w = d->walls3d[i][t];            // p1 = w[0..2], p2 = w[3..5], p3 = w[6..8], v = w[9..11]
        auto dot = [&](int a, int bq, int c) { return (px - w[a]) * (w[bq] - w[c]) + (py - w[a + 1]) * (w[bq + 1] - w[c + 1]) + (pz - w[a + 2]) * (w[bq + 2] - w[c + 2]); };
        const bool left = dot(0, 3, 0) > 0.0, right = dot(3, 0, 3) > 0.0, bottom = dot(6, 3, 6) > 0.0, top = dot(3, 6, 3) > 0.0;
        const double msk = (left && right && bottom && top) ? 1.0 : 0.0;
        idx[0] = i; idx[1] = p + i; idx[2] = 2 * p + i; g[0] = msk * w[9]; g[1] = msk * w[10]; g[2] = msk * w[11];
        fn(r++, ((px - w[0]) * w[9] + (py - w[1]) * w[10] + (pz - w[2]) * w[11]) * msk, 3, idx, g);
      }
      for (int t = 0; t < d->n_cyl[i]; t++) {           // CylinderConstraint (constraints/cylinder_constraint.jl:33-129)
        const double* cy = d->cyl[i][t];               // base point, axis, length, radius
        const int ax = (int)cy[3];
        const double t3[3] = {px - cy[0], py - cy[1], pz - cy[2]};
        const double along = t3[ax];
        const double valid = (along > 0.0 && along < cy[4]) ? 1.0 : 0.0;
        double dist2 = 0.0;
        for (int c = 0; c < 3; c++) { const double off = (c == ax) ? 0.0 : 1.0; dist2 += t3[c] * t3[c] * off; idx[c] = c * p + i; g[c] = -2.0 * t3[c] * off * valid; }
        fn(r++, (cy[5] * cy[5] - dist2) * valid, 3, idx, g);
      }
    }
  }
  template <class F> __device__ void control_rows(const double* u, F&& fn) const {
    if (!d->has_cb) return;
    int idx[1]; double g[1];
    for (int a = 0; a < m; a++) if (d->ub_row[a] >= 0) { idx[0] = a; g[0] = 1.0; fn(d->ub_row[a], u[a] - d->u_max[a], 1, idx, g); }
    for (int a = 0; a < m; a++) if (d->lb_row[a] >= 0) { idx[0] = a; g[0] = -1.0; fn(d->lb_row[a], d->u_min[a] - u[a], 1, idx, g); }
  }

  // ---- residual! (+ regularize_residual!) and, with jac, residual_jacobian! (+ its regularisation) -----------------
  // (problem/global_quantities.jl:9-193, constraints/constraint_derivatives.jl:1-74)
  __device__ Norms assemble(const double* Xa, const double* Ua, const double* La, const double* Xref, const double* Uref,
                            double reg_res, bool jac, double reg_jac) {
    // pass 1: RK2 values and forward-mode Jacobians [A | B] per (stage, player); zero the band
    BT_START();
    const int ncol = ni + mi;
    for (int item = tid; item < K * p * (ncol + 1); item += nt) {
      const int c = item % (ncol + 1), sp = item / (ncol + 1), i = sp % p, s = sp / p;
      if (c == ncol) {
        double st[kMaxNi], u[kMaxMi], xn[kMaxNi];
        for (int q = 0; q < ni; q++) st[q] = Xa[s * n + q * p + i];
        for (int q = 0; q < mi; q++) u[q] = Ua[s * m + q * p + i];
        rk2(d, st, u, xn);
        for (int q = 0; q < ni; q++) XN[s * n + q * p + i] = xn[q];
      } else {
        dual st[kMaxNi], u[kMaxMi], xn[kMaxNi];
        for (int q = 0; q < ni; q++) st[q] = mk(Xa[s * n + q * p + i], q == c ? 1.0 : 0.0);
        for (int q = 0; q < mi; q++) u[q] = mk(Ua[s * m + q * p + i], ni + q == c ? 1.0 : 0.0);
        rk2(d, st, u, xn);
        double* ab = AB + (size_t)(s * p + i) * nab;
        for (int q = 0; q < ni; q++) ab[q * ncol + c] = xn[q].d;
      }
    }
    BT_MARK(0);
    if (jac) for (size_t q = tid; q < (size_t)S * wd; q += nt) bandm[q] = 0.0;
    __syncthreads();
    BT_MARK(1);
    // pass 2a: one work item per band row — the smooth part of the row (dynamics defect, LQR gradient, costate terms) and its
    // Jacobian entries; rows are disjoint, so the order of the additions into every entry is the order of the serial form
    Norms nm = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int r = tid; r < S; r += nt) {
      const int s = r / b, e = r - s * b, k = s + 1;
      const double* x = Xa + k * n;
      const double* u = Ua + s * m;
      if (e < n) {                                                              // dyn_s: f_RK2(x_s, u_s) − x_{s+1}
        const int a = e, ia = a % p, ca = a / p;
        res[r] = XN[s * n + a] - x[a];
        if (jac) {
          const double* ab = AB + (size_t)(s * p + ia) * nab + ca * ncol;
          if (s >= 1) for (int q = 0; q < ni; q++) be(r, col_x(s - 1, q * p + ia)) += ab[q];
          for (int q = 0; q < mi; q++) be(r, col_u(s, q * p + ia)) += ab[ni + q];
          be(r, col_x(s, a)) += -1.0;
        }
      } else if (e < n + m) {                                                   // opt_i u_{i,s}: own control components
        const int idx = e - n, i = idx % p, j = idx / p;
        const double* ab = AB + (size_t)(s * p + i) * nab;
        const double* lamd = La + (size_t)(i * K + s) * n;
        double v = dt * R[idx] * (u[idx] - uf[idx]);
        for (int q = 0; q < ni; q++) v += ab[q * ncol + ni + j] * lamd[q * p + i];
        if (reg_res != 0.0) v += reg_res * (u[idx] - Uref[s * m + idx]);
        res[r] = v;
        if (jac) {
          be(r, col_u(s, idx)) += dt * R[idx] + reg_jac;
          for (int q = 0; q < ni; q++) be(r, col_l(s, i, q * p + i)) += ab[q * ncol + ni + j];
        }
      } else {                                                                  // opt_i x_{k}: row a of player i
        const int t = e - n - m, i = t / n, a = t - i * n;
        const double dtx = (k < K) ? dt : 1.0;                                  // the terminal knot is not dt-scaled
        const int ia = a % p, ca = a / p;
        double v = -La[(size_t)(i * K + s) * n + a];                            // −λ_{i,s}
        if (ia == i) v += dtx * Q[a] * (x[a] - xf[a]);                          // LQR gradient (objective.jl:24-32)
        if (s + 1 < K) {                                                        // + A_{s+1}ᵀ λ_{i,s+1}
          const double* ab = AB + (size_t)((s + 1) * p + ia) * nab;
          const double* lamn = La + (size_t)(i * K + s + 1) * n;
          for (int q = 0; q < ni; q++) v += ab[q * ncol + ca] * lamn[q * p + ia];
        }
        if (reg_res != 0.0) v += reg_res * (x[a] - Xref[k * n + a]);
        res[r] = v;
        if (jac) {
          be(r, col_l(s, i, a)) += -1.0;
          if (s + 1 < K) {
            const double* ab = AB + (size_t)((s + 1) * p + ia) * nab;
            for (int q = 0; q < ni; q++) be(r, col_l(s + 1, i, q * p + ia)) += ab[q * ncol + ca];
          }
          be(r, col_x(s, a)) += (ia == i ? dtx * Q[a] : 0.0) + reg_jac;
        }
      }
    }
    __syncthreads();
    // pass 2b: one work item per (stage, row group) — collision cost and augmented-Lagrangian terms, which add into several rows of
    // their own group: group 0 = control rows of stage s, group 1+i = state rows of player i at knot s+1
    for (int item = tid; item < K * (1 + p); item += nt) {
      const int grp = item % (1 + p), s = item / (1 + p), k = s + 1;
      const double* x = Xa + k * n;
      const double* u = Ua + s * m;
      if (grp == 0) {
        control_rows(u, [&](int row, double c, int nnz, const int* ix, const double* g) {
          const double lm = lam[s * nrow + row], mm = mu[s * nrow + row];
          const double w = ((c >= 0.0) || (lm > 0.0)) ? mm : 0.0, gl = lm + w * c;
          nm.con = fmax(nm.con, c);
          res[row_u(s, ix[0])] += g[0] * gl;
          if (jac) be(row_u(s, ix[0]), col_u(s, ix[0])) += w * g[0] * g[0];
          (void)nnz;
        });
      } else {
        const int i = grp - 1;
        const double dtx = (k < K) ? dt : 1.0;                                  // the terminal knot is not dt-scaled
        if (d->has_cc) {                                                        // soft collision cost on px (objective.jl:134-173)
          for (int j = 0; j < p; j++) {
            if (j == i) continue;
            const double dx = x[i] - x[j], dy = x[p + i] - x[p + j], dn = sqrt(dx * dx + dy * dy), rr = d->cc_radius[i], mc = d->cc_mu[i];
            if (fmax(0.0, rr - dn) > 0.0) {
              const double eps = 1e-10, en = eps * sqrt((double)n);
              const double gx = mc * (rr * (eps + dx) / (en + dn) - dx), gy = mc * (rr * (eps + dy) / (en + dn) - dy);
              res[row_x(s, i, i)] -= gx * dtx; res[row_x(s, i, p + i)] -= gy * dtx;
              res[row_x(s, i, j)] += gx * dtx; res[row_x(s, i, p + j)] += gy * dtx;
              if (jac) {
                const double dn3 = dn * dn * dn;
                const double h[2][2] = {{mc * (1 - rr / dn + rr * dx * dx / dn3), mc * (rr * dx * dy / dn3)}, {mc * (rr * dx * dy / dn3), mc * (1 - rr / dn + rr * dy * dy / dn3)}};
                const int pi[2] = {i, p + i}, pj[2] = {j, p + j};
                for (int a = 0; a < 2; a++) for (int c = 0; c < 2; c++) {
                  be(row_x(s, i, pi[a]), col_x(s, pi[c])) += dtx * h[a][c]; be(row_x(s, i, pi[a]), col_x(s, pj[c])) -= dtx * h[a][c];
                  be(row_x(s, i, pj[a]), col_x(s, pi[c])) -= dtx * h[a][c]; be(row_x(s, i, pj[a]), col_x(s, pj[c])) += dtx * h[a][c];
                }
              }
            }
          }
        }
        state_rows(i, x, [&](int row, double c, int nnz, const int* ix, const double* g) {
          const double lm = lam[s * nrow + row], mm = mu[s * nrow + row];
          const double w = ((c >= 0.0) || (lm > 0.0)) ? mm : 0.0, gl = lm + w * c;
          nm.sta = fmax(nm.sta, c);
          for (int e = 0; e < nnz; e++) {
            res[row_x(s, i, ix[e])] += g[e] * gl;
            if (jac) for (int f = 0; f < nnz; f++) be(row_x(s, i, ix[e]), col_x(s, ix[f])) += w * g[e] * g[f];
          }
        });
      }
    }
    __syncthreads();
    BT_MARK(2);
    for (int q = tid; q < S; q += nt) {
      const double v = fabs(res[q]);
      nm.sum += v;
      if (q % b < n) nm.dyn = fmax(nm.dyn, v); else nm.opt = fmax(nm.opt, v);
    }
    nm = reduce(nm);
    BT_MARK(3);
    return nm;
  }

  // ---- Δtraj = −(lu(jac) \ res): band LU with partial pivoting, in place; rhs ← solution.  false: singular ------------
  __device__ bool band_solve() {
    if (win == nullptr) return band_solve_global();
    switch ((kl + ku + 2 + 31) / 32) {                                          // 32-column chunks of a window row (+ right-hand side) per lane
      case 1: return band_solve_window<1>();
      case 2: return band_solve_window<2>();
      case 3: return band_solve_window<3>();
      case 4: return band_solve_window<4>();
      case 5: return band_solve_window<5>();
      case 6: return band_solve_window<6>();
      case 7: return band_solve_window<7>();
      default: return band_solve_window<8>();
    }
  }

  // The same elimination, operation for operation (identical pivots, identical roundings), with the active rows held in shared
  // memory.  At column j the live part of the matrix is rows j..j+kl, columns j..j+kl+ku: a window of WR = kl+2 physical row slots
  // (kl+1 live rows + the slot the next row streams into) of CW = kl+ku+1 entries addressed by ABSOLUTE column modulo CW, plus the
  // right-hand side in position CW.  Absolute columns make a row interchange a swap of two entries of the logical-row -> slot map
  // (double buffered by the parity of j) instead of a data move.  A column costs ONE block barrier:
  //   [rank-1 update, two rows per warp pass with all loads issued before the FMAs, the pivot row cached in registers; the lane that
  //    owns column j+1 keeps the arg-max of its warp's updated rows | pivot row retired to the global band (U, for the back
  //    substitution) | row j+kl+1 streamed into the spare slot | next map written]  barrier  [pivot of column j+1 = arg-max of the
  //    per-warp candidates and of the streamed row, first row on ties — the global path's choice].
  // The position of column j is zeroed as each row is updated: it is column j+CW of the next step.
  // Back substitution: warps 1.. stream the retired U rows back from the global band into two shared buffers of RB rows each while
  // warp 0 substitutes out of the other buffer (the last CW solution entries in a shared ring), reproducing the 128 partial sums of
  // the global path's reduction four per lane.
  template <int NCH> __device__ bool band_solve_window() {
    BT_START();
    const int CW = kl + ku + 1, WR = kl + 2, WS = NCH * 32 + 1;                  // row: CW columns | right-hand side | padding to NCH chunks of 32 (+1: odd stride)
    const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    int* lp = reinterpret_cast<int*>(win + (size_t)WR * WS);                    // lp[2][WR]
    int RW = 32; while (RW < CW + 1) RW <<= 1;
    double* xr = win + (size_t)WR * WS + WR;                                    // x ring [RW] (behind the maps: 2 WR ints = WR doubles)
    double* cand0 = xr + RW;                                                    // pivot candidates [2][value [nw+1] | row (as double) [nw+1]], by parity of j
    auto load_row = [&](int r, int c0, int q) -> double {                       // entry q of row r for the window whose first column is c0
      if (q == CW) return -res[r];
      const int c = c0 + q;
      return (c >= r - kl && c <= r + ku && c < S) ? bandm[(size_t)r * wd + (c - r + kl)] : 0.0;
    };
    const int r0 = (kl + 1 < S) ? kl + 1 : S;                                   // rows 0..r0-1 start in the window, slot = row
    for (int r = warp; r < r0; r += nw) {
      double* dst = win + (size_t)r * WS;
      for (int q = lane; q <= CW; q += 32) dst[q] = load_row(r, 0, q);          // columns 0..CW-1 sit at positions 0..CW-1
    }
    for (int q = tid; q < WR; q += nt) lp[q] = q * WS;                          // logical row -> element offset of its slot
    for (int q = tid; q < WR * (WS - CW - 1); q += nt) win[(size_t)(q / (WS - CW - 1)) * WS + CW + 1 + q % (WS - CW - 1)] = 0.0;   // row padding
    __syncthreads();
    // pivot of column 0: first row of maximal |a_r0| (every warp, same data, same answer)
    double best = -1.0; int pr = 0;
    for (int r = lane; r < r0; r += 32) { const double v = fabs(win[(size_t)r * WS]); if (v > best) { best = v; pr = r; } }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o); const int orr = __shfl_xor_sync(0xffffffffu, pr, o);
      if (ob > best || (ob == best && orr < pr)) { best = ob; pr = orr; }
    }
    int spare = (WR - 1) * WS;                                                  // element offset of the free physical slot (uniform)
    int posj = 0, lrj = 0;                                                      // j % CW, j % WR
    for (int j = 0; j < S; j++) {
      if (!(best > 0.0)) { __syncthreads(); return false; }                     // (uniform)
      const int rmax = (j + kl < S - 1) ? j + kl : S - 1, cmax = (j + ku + kl < S - 1) ? j + ku + kl : S - 1;
      const int* M = lp + (j & 1) * WR;
      int* Mn = lp + ((j & 1) ^ 1) * WR;
      int lrp = lrj + (pr - j); if (lrp >= WR) lrp -= WR;
      const int Pj = M[lrj], P = M[lrp];                                        // slots of logical row j and of the pivot row
      const double* prow = win + P;
      const double inv = 1.0 / prow[posj];
      const int nr = rmax - j, nc = cmax - j;
      int pos1 = posj + 1; if (pos1 >= CW) pos1 -= CW;                          // position of column j+1
      // every lane owns the same POSITIONS lane + 32 t of a window row for the whole elimination (position = column mod CW, position
      // CW = right-hand side), so the update addresses a row with compile-time offsets and no per-column index arithmetic.  At
      // column j the positions other than posj hold the live columns j+1 .. j+CW-1; a pivot row is zero beyond its fill, so
      // updating all of them performs the global path's operations plus subtractions of f * 0
      double pv[NCH];
#pragma unroll
      for (int t = 0; t < NCH; t++) pv[t] = prow[lane + 32 * t];
      const int lj = posj & 31;                                                 // the lane that owns the pivot column's position
      // retire the pivot row: U row j and its right-hand side go to the global band
      for (int cc = tid; cc <= nc + 1; cc += nt) {
        if (cc == nc + 1) rhs[j] = prow[CW];
        else { int q = posj + cc; if (q >= CW) q -= CW; bandm[(size_t)j * wd + kl + cc] = prow[q]; }
      }
      // the row that enters the window at step j+1 (values held in registers until the update is done)
      const int rn = j + kl + 1;
      double nv0 = 0.0, nv1 = 0.0;
      if (rn < S) {
        if (tid <= CW) nv0 = load_row(rn, j + 1, tid);
        if (tid + nt <= CW) nv1 = load_row(rn, j + 1, tid + nt);
      }
      BT_MARK(7);
      // rank-1 update, rows j+1+ri, two rows per pass with the loads issued before the FMAs; the lane that owns the position of
      // column j+1 tracks this warp's pivot candidate (rows ascend: strict > keeps the first maximum)
      double cbest = -1.0; int crow = 0;
      const int l1 = pos1 & 31;
      const bool track = (lane == l1) && (nc >= 1);
      auto pass = [&](auto rc, int ri0) {
        constexpr int R = decltype(rc)::value;
        double* row[R]; double f[R]; double v[R][NCH];
#pragma unroll
        for (int q = 0; q < R; q++) {
          const int ri = ri0 + q * nw, r = j + 1 + ri;
          int lr = lrj + 1 + ri; if (lr >= WR) lr -= WR;
          row[q] = win + ((r == pr) ? Pj : M[lr]);
        }
#pragma unroll
        for (int q = 0; q < R; q++) f[q] = row[q][posj] * inv;
#pragma unroll
        for (int q = 0; q < R; q++)
#pragma unroll
          for (int t = 0; t < NCH; t++) v[q][t] = row[q][lane + 32 * t];
#pragma unroll
        for (int q = 0; q < R; q++)
#pragma unroll
          for (int t = 0; t < NCH; t++) v[q][t] -= f[q] * pv[t];
        __syncwarp();                                                           // every lane has read its factors from position posj
#pragma unroll
        for (int q = 0; q < R; q++)
#pragma unroll
          for (int t = 0; t < NCH; t++) row[q][lane + 32 * t] = v[q][t];
        if (lane == lj) {
#pragma unroll
          for (int q = 0; q < R; q++) row[q][posj] = 0.0;                           // eliminated; the position becomes column j + CW
        }
        if (track) {                                                            // (this lane stored position pos1 itself, just above)
#pragma unroll
          for (int q = 0; q < R; q++) { const double av = fabs(row[q][pos1]); if (av > cbest) { cbest = av; crow = j + 1 + ri0 + q * nw; } }
        }
      };
      {
        int ri = warp;
        for (; ri + nw < nr; ri += 2 * nw) pass(std::integral_constant<int, 2>(), ri);
        for (; ri < nr; ri += nw) pass(std::integral_constant<int, 1>(), ri);
      }
      BT_MARK(8);
      // next step's map, the candidates and the new row
      int lrn = lrj + kl + 1; if (lrn >= WR) lrn -= WR;                           // logical position of row j+kl+1 (that of row j-1)
      for (int q = tid; q < WR; q += nt) {
        int v = M[q];
        if (q == lrp) v = Pj;
        if (q == lrn) v = spare;
        Mn[q] = v;
      }
      double* cand = cand0 + (j & 1) * 2 * (nw + 1);                             // (a fast warp may write step j+1's while a slow one still reads step j's)
      if (lane == l1) { cand[warp] = cbest; cand[nw + 1 + warp] = (double)crow; }
      if (tid == 0) { cand[nw] = (rn < S) ? fabs(nv0) : -1.0; cand[2 * nw + 1] = (double)rn; }   // entry 0 of the new row is column j+1
      if (rn < S) {
        double* dst = win + spare;
        // columns j+1 .. j+CW: position of column j+1+q is (pos1 + q) mod CW
        if (tid <= CW) { int q = pos1 + tid; if (q >= CW) q -= CW; dst[tid == CW ? CW : q] = nv0; }
        if (tid + nt <= CW) { int q = pos1 + tid + nt; if (q >= CW) q -= CW; dst[tid + nt == CW ? CW : q] = nv1; }
      }
      spare = P;
      posj = pos1; lrj = (lrj + 1 == WR) ? 0 : lrj + 1;
      __syncthreads();
      // pivot of column j+1: first row of maximal |a| among the candidates (rows of one warp ascend, strict > keeps the first)
      best = -1.0; pr = j + 1;
      if (j + 1 < S) {
        for (int w = 0; w <= nw; w++) {
          const double v = cand[w]; const int rr = (int)cand[nw + 1 + w];
          if (v > best || (v == best && v >= 0.0 && rr < pr)) { best = v; pr = rr; }
        }
        if (nc < 1) best = -1.0;                                                // (cannot happen for j + 1 < S: column j+1 exists)
      } else best = 1.0;
      BT_MARK(9);
    }
    BT_MARK(4);
    // ---- back substitution ----
    {
      const int RB = (int)(((size_t)WR * WS / 2) / (size_t)(CW + 1));            // U rows per buffer (>= 1)
      double* buf[2] = {win, win + (size_t)RB * (CW + 1)};
      // chunk g holds rows hi_g-1 down to lo_g (hi_0 = S); row j of a chunk at buf + (hi_g-1-j)*(CW+1): entries c = j..j+CW-1, then rhs
      auto load_chunk = [&](int hi, double* dst, int first, int step) {
        const int lo = hi - RB > 0 ? hi - RB : 0;
        for (int item = first; item < (hi - lo) * (CW + 1); item += step) {
          const int jr = item / (CW + 1), q = item - jr * (CW + 1), j = hi - 1 - jr;
          double v;
          if (q == CW) v = rhs[j];
          else { const int c = j + q; v = (c < S && q <= kl + ku) ? bandm[(size_t)j * wd + kl + q] : 0.0; }
          dst[item] = v;
        }
      };
      load_chunk(S, buf[0], tid, nt);
      __syncthreads();
      int cur = 0;
      for (int hi = S; hi > 0; hi -= RB, cur ^= 1) {
        const int lo = hi - RB > 0 ? hi - RB : 0;
        if (warp == 0) {
          const double* bsrc = buf[cur];
          for (int j = hi - 1; j >= lo; j--) {
            const double* urow = bsrc + (size_t)(hi - 1 - j) * (CW + 1);
            const int cmax = (j + ku + kl < S - 1) ? j + ku + kl : S - 1;
            double a[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int w = 0; w < 4; w++)
              for (int c = j + 1 + lane + 32 * w; c <= cmax; c += 128) a[w] += urow[c - j] * xr[c & (RW - 1)];
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
              for (int w = 0; w < 4; w++) a[w] += __shfl_xor_sync(0xffffffffu, a[w], o);
            }
            if (lane == 0) {
              double t = 0.0;
#pragma unroll
              for (int w = 0; w < 4; w++) t += a[w];
              const double x = (urow[CW] - t) / urow[0];
              xr[j & (RW - 1)] = x; rhs[j] = x;
            }
            __syncwarp();
          }
        } else if (lo > 0) {
          load_chunk(lo, buf[cur ^ 1], tid - 32, nt - 32);
        }
        __syncthreads();
      }
    }
    BT_MARK(5);
    return true;
  }

  __device__ bool band_solve_global() {
    for (int q = tid; q < S; q += nt) rhs[q] = -res[q];
    __syncthreads();
    int* ired = reinterpret_cast<int*>(red + 48);
    for (int j = 0; j < S; j++) {
      const int rmax = (j + kl < S - 1) ? j + kl : S - 1, cmax = (j + ku + kl < S - 1) ? j + ku + kl : S - 1;
      // pivot: first row of maximal |a_rj| in j..rmax
      if (tid < 32) {
        double best = -1.0; int br = j;
        for (int r = j + tid; r <= rmax; r += 32) { const double v = fabs(be(r, j)); if (v > best) { best = v; br = r; } }
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o); const int orr = __shfl_xor_sync(0xffffffffu, br, o);
          if (ob > best || (ob == best && orr < br)) { best = ob; br = orr; }
        }
        if (tid == 0) { ired[0] = br; red[47] = best; }
      }
      __syncthreads();
      const int pr = ired[0];
      if (!(red[47] > 0.0)) return false;                                       // (uniform)
      if (pr != j) {
        for (int c = j + tid; c <= cmax; c += nt) { const double t = be(j, c); be(j, c) = be(pr, c); be(pr, c) = t; }
        if (tid == 0) { const double t = rhs[j]; rhs[j] = rhs[pr]; rhs[pr] = t; }
        __syncthreads();
      }
      const double inv = 1.0 / be(j, j);
      const int nr = rmax - j, nc = cmax - j;
      for (int item = tid; item < nr * (nc + 1); item += nt) {
        const int r = j + 1 + item / (nc + 1), cc = item % (nc + 1);
        const double f = be(r, j) * inv;
        if (cc == nc) rhs[r] -= f * rhs[j];
        else be(r, j + 1 + cc) -= f * be(j, j + 1 + cc);
      }
      __syncthreads();
    }
    for (int j = S - 1; j >= 0; j--) {                                          // back substitution
      const int cmax = (j + ku + kl < S - 1) ? j + ku + kl : S - 1;
      double acc = 0.0;
      for (int c = j + 1 + tid; c <= cmax; c += nt) acc += be(j, c) * rhs[c];
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if ((tid & 31) == 0) red[40 + (tid >> 5)] = acc;
      __syncthreads();
      if (tid == 0) { double t = 0.0; for (int w = 0; w < nt / 32; w++) t += red[40 + w]; rhs[j] = (rhs[j] - t) / be(j, j); }
      __syncthreads();
    }
    return true;
  }

  __device__ void scatter_step() {                    // set_traj!: solution vector → ΔX, ΔU, ΔΛ (primal_dual_traj.jl:46-75)
    for (int q = tid; q < K * b; q += nt) {
      const int s = q / b, e = q % b;
      if (e < p * n) { const int i = e / n, a = e % n; dL[(size_t)(i * K + s) * n + a] = rhs[q]; }
      else if (e < p * n + m) dU[s * m + (e - p * n)] = rhs[q];
      else dX[(s + 1) * n + (e - p * n - m)] = rhs[q];
    }
    __syncthreads();
  }
  __device__ void axpy_traj(double alpha, double* Xo, double* Uo, double* Lo) {   // update_traj! (:109-128)
    for (int q = tid; q < N * n; q += nt) Xo[q] = (q < n) ? X[q] : X[q] + alpha * dX[q];
    for (int q = tid; q < K * m; q += nt) Uo[q] = U[q] + alpha * dU[q];
    for (int q = tid; q < p * K * n; q += nt) Lo[q] = L[q] + alpha * dL[q];
    __syncthreads();
  }
  __device__ double delta_sum() {
    Norms a = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int q = n + tid; q < N * n; q += nt) a.sum += fabs(dX[q]);
    for (int q = tid; q < K * m; q += nt) a.sum += fabs(dU[q]);
    return reduce(a).sum;
  }
  __device__ void rollout() {
    if (tid < p) {
      const int i = tid;
      double st[kMaxNi], u[kMaxMi], xn[kMaxNi];
      for (int c = 0; c < ni; c++) st[c] = X[c * p + i];
      for (int s = 0; s < K; s++) {
        for (int c = 0; c < mi; c++) u[c] = U[s * m + c * p + i];
        rk3(d, st, u, xn);
        for (int c = 0; c < ni; c++) { st[c] = xn[c]; X[(s + 1) * n + c * p + i] = xn[c]; }
      }
    }
    __syncthreads();
  }
  // evaluate! + dual_update! + penalty_update! (constraints_methods.jl:329-365, :421-440)
  __device__ void dual_penalty_update(const agb_options& o) {
    for (int item = tid; item < K * (1 + p); item += nt) {
      const int grp = item % (1 + p), s = item / (1 + p);
      auto upd = [&](int row, double c, double a) {
        double* lm = &lam[s * nrow + row]; double* mm = &mu[s * nrow + row];
        *lm = fmin(fmax(*lm + a * (*mm) * c, 0.0), o.lambda_max);
        *mm = fmin(fmax(o.rho_increase * (*mm), 0.0), o.rho_max);
      };
      if (grp == 0) control_rows(U + s * m, [&](int row, double c, int, const int*, const double*) { upd(row, c, o.alpha_dual); });
      else state_rows(grp - 1, X + (s + 1) * n, [&](int row, double c, int, const int*, const double*) { upd(row, c, o.alphax_dual[grp - 1]); });
    }
    __syncthreads();
  }
};

}  // namespace band
}  // namespace agb
