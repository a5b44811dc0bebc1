/*
 * algames_oracle.c — plain-C restatement of Algames.jl's newton_solve! hot path, used ONLY as the timed CPU
 * baseline (bench.py: cpu_baseline / --impl reference) and cross-checked against the NumPy oracle in tests/.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY: nothing in the shipped package links or loads this file.
 *
 * It follows the reference function by function (paths relative to the reference root):
 *   newton_solve!, inner_iteration, line_search      src/problem/solver_methods.jl:5-125
 *   residual!, regularize_residual!                  src/problem/global_quantities.jl:9-86
 *   residual_jacobian!, regularize_residual_jacobian! src/problem/global_quantities.jl:109-193
 *   constraint_residual!/constraint_jacobian_residual! src/constraints/constraint_derivatives.jl:1-74
 *   dual_update!, penalty_update!, reset!            src/constraints/constraints_methods.jl:295-445
 *   violations / record!                             src/struct/violations.jl:18-168, src/struct/statistics.jl:44-57
 * and, like the reference, forms the S x S KKT Jacobian explicitly and factorises it from scratch every Newton step
 * (reference: UMFPACK sparse LU, solver_methods.jl:87).  UMFPACK is not available here; the stand-in is a band LU with
 * partial pivoting (LAPACK dgbsv algorithm) on the time-major permutation of the same matrix — rows
 * [dyn_k | opt u_k | opt x_{k+1}], columns [λ_k | u_k | x_{k+1}] per stage — which has lower/upper bandwidths
 * 2n-1 and p·n+n-1.  Instances are independent and are spread over host cores with a pthread work queue
 * (libgomp is not installed in the image, so no OpenMP).
 *
 * Parity status: same as oracle/algames_oracle.py (pinned against the reference's known-answer tests through that
 * file; RK2/RK3 tableaux and CollisionConstraint/CircleConstraint formulas are "parity unpinned" third-party code).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>
#include "../include/algames_b200.h"

#define MAXROW 256   /* >= p(p-1) + p(2n + walls + circles) + 2m at the ABI maxima (220) */

typedef struct {
  const agb_problem_desc* d;
  const agb_options* o;
  int p, n, m, N, K, b, S, nrow, kl, ku, wd;
  double dt;
  const double *xf, *Q, *R, *uf;            /* joint, this instance */
  /* row schema */
  int nrow_state;
  int row_owner[MAXROW];
  /* state */
  double *X, *U, *L, *Xt, *Ut, *Lt, *dX, *dU, *dL, *lam, *mu, *res, *band, *rhs;
  int n_newton, n_eval;
  double* hist; int hist_max, n_rec;   /* optional record!(stats, …) log of this instance (statistics.jl:44-57) */
} Work;

/* ---- models (src/dynamics/*.jl), per player: s[4], u[2] ---------------------------------------------------- */
static void dyn_f(const Work* w, const double* s, const double* u, double* f) {
  const agb_problem_desc* d = w->d;
  if (d->model == AGB_MODEL_DOUBLE_INTEGRATOR) { f[0] = s[2]; f[1] = s[3]; f[2] = u[0]; f[3] = u[1]; }
  else if (d->model == AGB_MODEL_UNICYCLE) { f[0] = cos(s[2]) * s[3]; f[1] = sin(s[2]) * s[3]; f[2] = u[0]; f[3] = u[1]; }
  else {
    double beta = atan2(d->lr * tan(u[1]), d->lr + d->lf);
    f[0] = s[2] * cos(beta + s[3]); f[1] = s[2] * sin(beta + s[3]); f[2] = u[0]; f[3] = s[2] * sin(beta) / d->lr;
  }
}
static void dyn_jac(const Work* w, const double* s, const double* u, double* Fx, double* Fu) {
  const agb_problem_desc* d = w->d;
  memset(Fx, 0, 16 * sizeof(double)); memset(Fu, 0, 8 * sizeof(double));
  if (d->model == AGB_MODEL_DOUBLE_INTEGRATOR) { Fx[2] = 1; Fx[7] = 1; Fu[4] = 1; Fu[7] = 1; }
  else if (d->model == AGB_MODEL_UNICYCLE) {
    double sn = sin(s[2]), cs = cos(s[2]);
    Fx[2] = -sn * s[3]; Fx[3] = cs; Fx[6] = cs * s[3]; Fx[7] = sn; Fu[4] = 1; Fu[7] = 1;
  } else {
    double L = d->lr + d->lf, t = tan(u[1]), beta = atan2(d->lr * t, L);
    double db = d->lr * L * (1 + t * t) / (L * L + d->lr * d->lr * t * t);
    double sn = sin(beta + s[3]), cs = cos(beta + s[3]), v = s[2];
    Fx[2] = cs; Fx[3] = -v * sn; Fx[6] = sn; Fx[7] = v * cs; Fx[14] = sin(beta) / d->lr;
    Fu[1] = -v * sn * db; Fu[3] = v * cs * db; Fu[4] = 1; Fu[7] = v * cos(beta) * db / d->lr;
  }
}
/* discrete_dynamics(RK2) + its Jacobian (local_quantities.jl:13,26) */
static void rk2(const Work* w, const double* s, const double* u, double* xn, double* A, double* B) {
  double f0[4], sm[4], f1[4], dt = w->dt, h = dt / 2;
  dyn_f(w, s, u, f0);
  for (int c = 0; c < 4; c++) sm[c] = s[c] + (f0[c] * dt) / 2;
  dyn_f(w, sm, u, f1);
  for (int c = 0; c < 4; c++) xn[c] = s[c] + f1[c] * dt;
  if (!A) return;
  double Fx0[16], Fu0[8], Fxm[16], Fum[8];
  dyn_jac(w, s, u, Fx0, Fu0); dyn_jac(w, sm, u, Fxm, Fum);
  for (int r = 0; r < 4; r++) {
    for (int c = 0; c < 4; c++) {
      double acc = 0;
      for (int q = 0; q < 4; q++) acc += Fxm[r * 4 + q] * ((q == c) + h * Fx0[q * 4 + c]);
      A[r * 4 + c] = (r == c) + dt * acc;
    }
    for (int j = 0; j < 2; j++) {
      double acc = 0;
      for (int q = 0; q < 4; q++) acc += Fxm[r * 4 + q] * (h * Fu0[q * 2 + j]);
      B[r * 2 + j] = dt * (acc + Fum[r * 2 + j]);
    }
  }
}
static void rk3(const Work* w, const double* s, const double* u, double* xn) {
  double k1[4], k2[4], k3[4], t[4], dt = w->dt;
  dyn_f(w, s, u, k1);
  for (int c = 0; c < 4; c++) { k1[c] *= dt; t[c] = s[c] + k1[c] / 2; }
  dyn_f(w, t, u, k2);
  for (int c = 0; c < 4; c++) { k2[c] *= dt; t[c] = s[c] - k1[c] + 2 * k2[c]; }
  dyn_f(w, t, u, k3);
  for (int c = 0; c < 4; c++) { k3[c] *= dt; xn[c] = s[c] + (k1[c] + 4 * k2[c] + k3[c]) / 6; }
}

/* ---- index maps (time-major band order) --------------------------------------------------------------------- */
#define ROW_DYN(s, a) ((s) * w->b + (a))
#define ROW_U(s, idx) ((s) * w->b + w->n + (idx))
#define ROW_X(s, i, a) ((s) * w->b + w->n + w->m + (i) * w->n + (a))       /* opt_i x at knot s+1 */
#define COL_L(s, i, a) ((s) * w->b + (i) * w->n + (a))
#define COL_U(s, idx) ((s) * w->b + w->p * w->n + (idx))
#define COL_X(s, a) ((s) * w->b + w->p * w->n + w->m + (a))                /* x at knot s+1 */
static inline void jadd(Work* w, int r, int c, double v) { w->band[(size_t)r * w->wd + (c - r + w->kl)] += v; }

/* ---- constraints: one row = (value c, sparse gradient) ------------------------------------------------------ */
typedef struct { double c; int nnz; int idx[4]; double g[4]; int is_control; } ConRow;

/* enumerate the AL rows of stage s (state rows of knot s+1, control rows of knot s) at iterate (X,U) */
static int con_rows(const Work* w, const double* X, const double* U, int s, ConRow* out) {
  const agb_problem_desc* d = w->d;
  const int p = w->p, n = w->n, m = w->m, k = s + 1;
  const double* x = X + k * n; const double* u = U + s * m;
  int r = 0;
  for (int i = 0; i < p; i++) {
    for (int j = 0; j < p; j++) if (j != i && d->col_radius[i][j] > 0) {          /* CollisionConstraint */
      double dx = x[i] - x[j], dy = x[p + i] - x[p + j], rad = d->col_radius[i][j];
      ConRow* q = &out[r++]; q->is_control = 0; q->c = rad * rad - (dx * dx + dy * dy); q->nnz = 4;
      q->idx[0] = i; q->g[0] = -2 * dx; q->idx[1] = p + i; q->g[1] = -2 * dy;
      q->idx[2] = j; q->g[2] = 2 * dx;  q->idx[3] = p + j; q->g[3] = 2 * dy;
    }
    for (int g = 0; g < d->has_state_bound[i]; g++) {                              /* StateBoundConstraint convals, in the order added */
      for (int a = 0; a < n; a++) if (isfinite(d->x_max[i][a]) && d->x_max_con[i][a] == g) { ConRow* q = &out[r++]; q->is_control = 0; q->c = x[a] - d->x_max[i][a]; q->nnz = 1; q->idx[0] = a; q->g[0] = 1; }
      for (int a = 0; a < n; a++) if (isfinite(d->x_min[i][a]) && d->x_min_con[i][a] == g) { ConRow* q = &out[r++]; q->is_control = 0; q->c = d->x_min[i][a] - x[a]; q->nnz = 1; q->idx[0] = a; q->g[0] = -1; }
    }
    double px = x[i], py = x[p + i];
    for (int t = 0; t < d->n_walls[i]; t++) {                                      /* WallConstraint */
      const double* wl = d->walls[i][t];
      int left = (px - wl[0]) * (wl[2] - wl[0]) + (py - wl[1]) * (wl[3] - wl[1]) > 0;
      int right = (px - wl[2]) * (wl[0] - wl[2]) + (py - wl[3]) * (wl[1] - wl[3]) > 0;
      double msk = (left && right) ? 1.0 : 0.0;
      ConRow* q = &out[r++]; q->is_control = 0; q->c = ((px - wl[0]) * wl[4] + (py - wl[1]) * wl[5]) * msk; q->nnz = 2;
      q->idx[0] = i; q->g[0] = msk * wl[4]; q->idx[1] = p + i; q->g[1] = msk * wl[5];
    }
    for (int t = 0; t < d->n_circles[i]; t++) {                                    /* CircleConstraint */
      const double* cl = d->circles[i][t];
      double ex = px - cl[0], ey = py - cl[1];
      ConRow* q = &out[r++]; q->is_control = 0; q->c = cl[2] * cl[2] - ex * ex - ey * ey; q->nnz = 2;
      q->idx[0] = i; q->g[0] = -2 * ex; q->idx[1] = p + i; q->g[1] = -2 * ey;
    }
  }
  if (d->has_control_bound) {                                                      /* ControlBoundConstraint */
    for (int a = 0; a < m; a++) if (isfinite(d->u_max[a])) { ConRow* q = &out[r++]; q->is_control = 1; q->c = u[a] - d->u_max[a]; q->nnz = 1; q->idx[0] = a; q->g[0] = 1; }
    for (int a = 0; a < m; a++) if (isfinite(d->u_min[a])) { ConRow* q = &out[r++]; q->is_control = 1; q->c = d->u_min[a] - u[a]; q->nnz = 1; q->idx[0] = a; q->g[0] = -1; }
  }
  return r;
}

/* ---- residual! (+ regularize_residual!) and, if jac != 0, residual_jacobian! (+ regularisation) -------------- */
typedef struct { double sum, opt, dyn, con, sta; } Norms;

static Norms assemble(Work* w, const double* X, const double* U, const double* L, const double* Xref, const double* Uref,
                      double reg_res, int jac, double reg_jac) {
  const agb_problem_desc* d = w->d;
  const int p = w->p, n = w->n, m = w->m, N = w->N, K = w->K, S = w->S;
  double* res = w->res;
  memset(res, 0, S * sizeof(double));
  if (jac) memset(w->band, 0, (size_t)S * w->wd * sizeof(double));
  Norms nm = {0, 0, 0, 0, 0};
  ConRow rows[MAXROW];
  for (int s = 0; s < K; s++) {
    /* dynamics: A_s, B_s per player; defect row; A^T λ, B^T λ, -λ contributions (global_quantities.jl:43-63) */
    for (int i = 0; i < p; i++) {
      double st[4], u[2], xn[4], A[16], B[8];
      for (int c = 0; c < 4; c++) st[c] = X[s * n + c * p + i];
      for (int j = 0; j < 2; j++) u[j] = U[s * m + j * p + i];
      rk2(w, st, u, xn, A, B);
      for (int c = 0; c < 4; c++) res[ROW_DYN(s, c * p + i)] += xn[c] - X[(s + 1) * n + c * p + i];
      for (int pl = 0; pl < p; pl++) {                              /* every player pl sees A_i, and B_i if pl == i */
        const double* lam = L + (pl * K + s) * n;
        if (s >= 1) for (int c = 0; c < 4; c++) { double acc = 0; for (int q = 0; q < 4; q++) acc += A[q * 4 + c] * lam[q * p + i]; res[ROW_X(s - 1, pl, c * p + i)] += acc; }
        if (pl == i) for (int j = 0; j < 2; j++) { double acc = 0; for (int q = 0; q < 4; q++) acc += B[q * 2 + j] * lam[q * p + i]; res[ROW_U(s, j * p + i)] += acc; }
        for (int c = 0; c < 4; c++) res[ROW_X(s, pl, c * p + i)] -= lam[c * p + i];
        if (jac) {
          if (s >= 1) for (int c = 0; c < 4; c++) for (int q = 0; q < 4; q++) jadd(w, ROW_X(s - 1, pl, c * p + i), COL_L(s, pl, q * p + i), A[q * 4 + c]);
          if (pl == i) for (int j = 0; j < 2; j++) for (int q = 0; q < 4; q++) jadd(w, ROW_U(s, j * p + i), COL_L(s, pl, q * p + i), B[q * 2 + j]);
          for (int c = 0; c < 4; c++) jadd(w, ROW_X(s, pl, c * p + i), COL_L(s, pl, c * p + i), -1.0);
        }
      }
      if (jac) {
        for (int c = 0; c < 4; c++) {
          if (s >= 1) for (int q = 0; q < 4; q++) jadd(w, ROW_DYN(s, c * p + i), COL_X(s - 1, q * p + i), A[c * 4 + q]);
          for (int j = 0; j < 2; j++) jadd(w, ROW_DYN(s, c * p + i), COL_U(s, j * p + i), B[c * 2 + j]);
          jadd(w, ROW_DYN(s, c * p + i), COL_X(s, c * p + i), -1.0);
        }
      }
    }
    /* cost at knot k = s+1 (x rows) and stage s (u rows): LQR + collision cost, dt-scaled (objective.jl) */
    const int k = s + 1;
    const double dtx = (k < N - 1) ? w->dt : 1.0;
    const double* x = X + k * n; const double* u = U + s * m;
    for (int i = 0; i < p; i++) {
      for (int c = 0; c < 4; c++) {
        int a = c * p + i;
        res[ROW_X(s, i, a)] += dtx * w->Q[a] * (x[a] - w->xf[a]);
        if (jac) jadd(w, ROW_X(s, i, a), COL_X(s, a), dtx * w->Q[a]);
      }
      for (int j = 0; j < 2; j++) {
        int a = j * p + i;
        res[ROW_U(s, a)] += w->dt * w->R[a] * (u[a] - w->uf[a]);
        if (jac) jadd(w, ROW_U(s, a), COL_U(s, a), w->dt * w->R[a]);
      }
      if (d->has_collision_cost) for (int j = 0; j < p; j++) if (j != i) {        /* objective.jl:134-173 */
        double dx = x[i] - x[j], dy = x[p + i] - x[p + j], dn = sqrt(dx * dx + dy * dy), rr = d->cc_radius[i], mu = d->cc_mu[i];
        if (fmax(0.0, rr - dn) > 0.0) {
          double eps = 1e-10, en = eps * sqrt((double)n);
          double gx = mu * (rr * (eps + dx) / (en + dn) - dx), gy = mu * (rr * (eps + dy) / (en + dn) - dy);
          res[ROW_X(s, i, i)] -= gx * dtx; res[ROW_X(s, i, p + i)] -= gy * dtx;
          res[ROW_X(s, i, j)] += gx * dtx; res[ROW_X(s, i, p + j)] += gy * dtx;
          if (jac) {
            double dn3 = dn * dn * dn, h[2][2] = {{mu * (1 - rr / dn + rr * dx * dx / dn3), mu * (rr * dx * dy / dn3)},
                                                   {mu * (rr * dx * dy / dn3), mu * (1 - rr / dn + rr * dy * dy / dn3)}};
            int pi[2] = {i, p + i}, pj[2] = {j, p + j};
            for (int a = 0; a < 2; a++) for (int c = 0; c < 2; c++) {
              jadd(w, ROW_X(s, i, pi[a]), COL_X(s, pi[c]), dtx * h[a][c]); jadd(w, ROW_X(s, i, pi[a]), COL_X(s, pj[c]), -dtx * h[a][c]);
              jadd(w, ROW_X(s, i, pj[a]), COL_X(s, pi[c]), -dtx * h[a][c]); jadd(w, ROW_X(s, i, pj[a]), COL_X(s, pj[c]), dtx * h[a][c]);
            }
          }
        }
      }
    }
    /* AL terms (constraint_derivatives.jl; expansion pinned by test/constraints/constraint_derivatives.jl:28-34) */
    int nr = con_rows(w, X, U, s, rows);
    for (int r = 0; r < nr; r++) {
      ConRow* q = &rows[r];
      double lam = w->lam[s * w->nrow + r], mu = w->mu[s * w->nrow + r];
      double wt = ((q->c >= 0) || (lam > 0)) ? mu : 0.0, g = lam + wt * q->c;
      if (q->is_control) {
        nm.con = fmax(nm.con, q->c);
        for (int e = 0; e < q->nnz; e++) {                      /* every player takes its own control components */
          res[ROW_U(s, q->idx[e])] += q->g[e] * g;
          if (jac) for (int f = 0; f < q->nnz; f++) if (q->idx[e] % p == q->idx[f] % p) jadd(w, ROW_U(s, q->idx[e]), COL_U(s, q->idx[f]), wt * q->g[e] * q->g[f]);
        }
      } else {
        int i = w->row_owner[r];
        nm.sta = fmax(nm.sta, q->c);
        for (int e = 0; e < q->nnz; e++) {
          res[ROW_X(s, i, q->idx[e])] += q->g[e] * g;
          if (jac) for (int f = 0; f < q->nnz; f++) jadd(w, ROW_X(s, i, q->idx[e]), COL_X(s, q->idx[f]), wt * q->g[e] * q->g[f]);
        }
      }
    }
    /* regularisation (global_quantities.jl:67-86 residual, :176-193 Jacobian) */
    for (int i = 0; i < p; i++) {
      if (reg_res != 0.0) {
        for (int a = 0; a < n; a++) res[ROW_X(s, i, a)] += reg_res * (X[k * n + a] - Xref[k * n + a]);
        for (int j = 0; j < 2; j++) res[ROW_U(s, j * p + i)] += reg_res * (U[s * m + j * p + i] - Uref[s * m + j * p + i]);
      }
      if (jac) {
        for (int a = 0; a < n; a++) jadd(w, ROW_X(s, i, a), COL_X(s, a), reg_jac);
        for (int j = 0; j < 2; j++) jadd(w, ROW_U(s, j * p + i), COL_U(s, j * p + i), reg_jac);
      }
    }
  }
  for (int s = 0; s < K; s++) {
    for (int a = 0; a < n; a++) { double v = fabs(res[ROW_DYN(s, a)]); nm.sum += v; nm.dyn = fmax(nm.dyn, v); }
    for (int a = n; a < w->b; a++) { double v = fabs(res[s * w->b + a]); nm.sum += v; nm.opt = fmax(nm.opt, v); }
  }
  w->n_eval++;
  return nm;
}

/* ---- Δtraj = −(lu(jac) \ res): band LU with partial pivoting (row-stored band, width 2kl+ku+1) ---------------- */
static int band_solve(Work* w) {
  const int S = w->S, kl = w->kl, ku = w->ku, wd = w->wd;
  double* B = w->band; double* x = w->rhs;
  for (int r = 0; r < S; r++) x[r] = -w->res[r];
#define BE(r, c) B[(size_t)(r) * wd + ((c) - (r) + kl)]
  for (int j = 0; j < S; j++) {
    int rmax = j + kl < S - 1 ? j + kl : S - 1, cmax = j + ku + kl < S - 1 ? j + ku + kl : S - 1;
    int pr = j; double best = fabs(BE(j, j));
    for (int r = j + 1; r <= rmax; r++) { double v = fabs(BE(r, j)); if (v > best) { best = v; pr = r; } }
    if (!(best > 0.0)) return 1;
    if (pr != j) {
      for (int c = j; c <= cmax; c++) { double t = BE(j, c); BE(j, c) = BE(pr, c); BE(pr, c) = t; }
      double t = x[j]; x[j] = x[pr]; x[pr] = t;
    }
    const double inv = 1.0 / BE(j, j);
    for (int r = j + 1; r <= rmax; r++) {
      double f = BE(r, j) * inv;
      if (f == 0.0) continue;
      for (int c = j + 1; c <= cmax; c++) BE(r, c) -= f * BE(j, c);
      x[r] -= f * x[j];
    }
  }
  for (int j = S - 1; j >= 0; j--) {
    int cmax = j + ku + kl < S - 1 ? j + ku + kl : S - 1;
    double acc = x[j];
    for (int c = j + 1; c <= cmax; c++) acc -= BE(j, c) * x[c];
    x[j] = acc / BE(j, j);
  }
#undef BE
  return 0;
}

static void scatter_step(Work* w) {          /* set_traj!: solution vector → ΔX, ΔU, ΔΛ (primal_dual_traj.jl:46-75) */
  for (int s = 0; s < w->K; s++) {
    for (int a = 0; a < w->n; a++) w->dX[(s + 1) * w->n + a] = w->rhs[COL_X(s, a)];
    for (int a = 0; a < w->m; a++) w->dU[s * w->m + a] = w->rhs[COL_U(s, a)];
    for (int i = 0; i < w->p; i++) for (int a = 0; a < w->n; a++) w->dL[(i * w->K + s) * w->n + a] = w->rhs[COL_L(s, i, a)];
  }
}
static void axpy_traj(Work* w, double alpha, double* Xo, double* Uo, double* Lo) {   /* update_traj! (:109-128) */
  for (int q = w->n; q < w->N * w->n; q++) Xo[q] = w->X[q] + alpha * w->dX[q];
  for (int q = 0; q < w->n; q++) Xo[q] = w->X[q];
  for (int q = 0; q < w->K * w->m; q++) Uo[q] = w->U[q] + alpha * w->dU[q];
  for (int q = 0; q < w->p * w->K * w->n; q++) Lo[q] = w->L[q] + alpha * w->dL[q];
}

static void setup_rows(Work* w) {
  const agb_problem_desc* d = w->d;
  int r = 0;
  for (int i = 0; i < w->p; i++) {
    int r0 = r;
    for (int j = 0; j < w->p; j++) if (j != i && d->col_radius[i][j] > 0) r++;
    if (d->has_state_bound[i]) for (int a = 0; a < w->n; a++) { if (isfinite(d->x_max[i][a])) r++; }
    if (d->has_state_bound[i]) for (int a = 0; a < w->n; a++) { if (isfinite(d->x_min[i][a])) r++; }
    r += d->n_walls[i] + d->n_circles[i];
    for (int q = r0; q < r; q++) w->row_owner[q] = i;
  }
  w->nrow_state = r;
  if (d->has_control_bound) for (int a = 0; a < w->m; a++) { if (isfinite(d->u_max[a])) r++; if (isfinite(d->u_min[a])) r++; }
  for (int q = w->nrow_state; q < r; q++) w->row_owner[q] = -1;
  w->nrow = r;
}

/* newton_solve!(prob) for one instance.  stats[10] as in include/algames_b200.h; returns the status code. */
static void log_rec(Work* w, const Norms* r, double delta, int kout, int l) {
  if (w->hist && w->n_rec < w->hist_max) {
    double* h = w->hist + (size_t)w->n_rec * AGB_NHIST;
    h[0] = kout; h[1] = r->sum / (double)w->S; h[2] = r->dyn; h[3] = r->con; h[4] = r->sta; h[5] = r->opt; h[6] = delta; h[7] = l; h[8] = -1; h[9] = 0;
  }
  w->n_rec++;
}

static int solve_one(Work* w, double* stats) {
  const agb_options* o = w->o;
  const int n = w->n, m = w->m, p = w->p, K = w->K;
  const double Sd = (double)w->S;
  for (int s = 0; s < K; s++) for (int i = 0; i < p; i++) {                      /* rollout!(RK3) (:17) */
    double st[4], u[2], xn[4];
    for (int c = 0; c < 4; c++) st[c] = w->X[s * n + c * p + i];
    u[0] = w->U[s * m + i]; u[1] = w->U[s * m + p + i];
    rk3(w, st, u, xn);
    for (int c = 0; c < 4; c++) w->X[(s + 1) * n + c * p + i] = xn[c];
  }
  if (o->dual_reset) for (int q = 0; q < K * w->nrow; q++) { w->lam[q] = 0; w->mu[q] = o->rho_0; }
  Norms rec = {0, 0, 0, 0, 0};
  double delta = 0; int outer = 0, failed = 0, last_exit = AGB_MAX_OUTER;
  w->n_newton = 0; w->n_eval = 0; w->n_rec = 0;
  ConRow rows[MAXROW];
  for (int kout = 1; kout <= o->outer_iter; kout++) {
    outer = kout;
    int ls_count = 0;
    last_exit = AGB_MAX_OUTER;
    for (int l = 1; l <= o->inner_iter; l++) {
      double l2 = (double)l * l, reg = o->reg_0 * (l2 * l2);
      rec = assemble(w, w->X, w->U, w->L, w->X, w->U, 0.0, 1, reg);              /* residual! + residual_jacobian! */
      double res_norm = rec.sum / Sd;
      log_rec(w, &rec, delta, kout, l);                                          /* record!(stats, …) (:75) */
      delta = 0;
      if (!(rec.sum == rec.sum) || isinf(rec.sum)) { if (!failed) failed = AGB_NONFINITE; break; }
      if (rec.opt < o->eps_opt) break;
      if (band_solve(w)) { if (!failed) failed = AGB_SINGULAR; }
      w->n_newton++;
      scatter_step(w);
      double alpha = 1.0; int j = 1;                                             /* line_search (:105-125) */
      while (j < o->ls_iter) {
        axpy_traj(w, alpha, w->Xt, w->Ut, w->Lt);
        Norms t = assemble(w, w->Xt, w->Ut, w->Lt, w->X, w->U, o->regularize ? reg : 0.0, 0, 0.0);
        if (t.sum / Sd <= (1.0 - alpha * o->beta) * res_norm) break;
        alpha *= o->alpha_decrease; j++;
      }
      ls_count = (j == o->ls_iter) ? ls_count + 1 : 0;
      double acc = 0;
      for (int q = n; q < w->N * n; q++) acc += fabs(w->dX[q]);
      for (int q = 0; q < K * m; q++) acc += fabs(w->dU[q]);
      axpy_traj(w, alpha, w->X, w->U, w->L);
      delta = alpha * acc / (double)(K * (n + m));
      if (delta < o->delta_min) { last_exit = AGB_STALLED; break; }
      if (ls_count >= 1) { last_exit = AGB_LINE_SEARCH_FAILED; break; }
      if (!(delta == delta)) { if (!failed) failed = AGB_NONFINITE; break; }
    }
    if (failed) break;
    if (kout == o->outer_iter || (rec.dyn < o->eps_dyn && rec.con < o->eps_con && rec.sta < o->eps_sta && rec.opt < o->eps_opt)) break;
    for (int s = 0; s < K; s++) {                                                /* evaluate! + dual_update! + penalty_update! */
      int nr = con_rows(w, w->X, w->U, s, rows);
      for (int r = 0; r < nr; r++) {
        int own = w->row_owner[r];
        double a = own >= 0 ? o->alphax_dual[own] : o->alpha_dual;
        double* lam = &w->lam[s * w->nrow + r]; double* mu = &w->mu[s * w->nrow + r];
        *lam = fmin(fmax(*lam + a * (*mu) * rows[r].c, 0.0), o->lambda_max);
        *mu = fmin(fmax(o->rho_increase * (*mu), 0.0), o->rho_max);
      }
    }
  }
  rec = assemble(w, w->X, w->U, w->L, w->X, w->U, 0.0, 0, 0.0);
  log_rec(w, &rec, delta, outer, 0);                                             /* :63 */
  int finite = (rec.sum == rec.sum) && !isinf(rec.sum);
  int conv = finite && rec.dyn < o->eps_dyn && rec.con < o->eps_con && rec.sta < o->eps_sta && rec.opt < o->eps_opt;
  stats[0] = rec.sum / Sd; stats[1] = rec.dyn; stats[2] = rec.con; stats[3] = rec.sta; stats[4] = rec.opt; stats[5] = delta;
  stats[6] = w->n_newton; stats[7] = outer; stats[8] = w->n_eval; stats[9] = failed != 0;
  return conv ? AGB_CONVERGED : (failed ? failed : (!finite ? AGB_NONFINITE : last_exit));
}

/* Batched entry point (ctypes).  Layouts as in include/algames_b200.h; nthreads <= 0 → all cores. Returns threads used. */
typedef struct {
  const agb_problem_desc* d; const agb_options* o; int batch;
  const double *x0, *xf, *Q, *R, *uf, *Z0, *L0; double *Z, *L, *stats; int* status;
  atomic_int next;
  const double *lam0, *mu0;      /* [B][K][nrow] initial AL multipliers / penalties (used when !dual_reset), or NULL */
  double *lam_out, *mu_out;      /* [B][K][nrow] or NULL */
  double* hist; int* hist_count; int hist_max;   /* [B][hist_max][AGB_NHIST], [B], or NULL */
} Job;

static void* worker(void* arg) {
  Job* jb = (Job*)arg;
  const agb_problem_desc* d = jb->d; const agb_options* o = jb->o;
  const int p = d->p, n = 4 * p, m = 2 * p, N = d->N, K = N - 1, b = p * n + m + n, S = K * b;
  Work w; memset(&w, 0, sizeof w);
  w.d = d; w.o = o; w.p = p; w.n = n; w.m = m; w.N = N; w.K = K; w.b = b; w.S = S; w.dt = d->dt;
  w.kl = 2 * n - 1; w.ku = p * n + n - 1; w.wd = 2 * w.kl + w.ku + 1;
  setup_rows(&w);
  size_t nd = (size_t)3 * (N * n + N * m + p * K * n) + 2 * (size_t)K * (w.nrow + 1) + 2 * (size_t)S + (size_t)S * w.wd;
  double* mem = (double*)calloc(nd, sizeof(double)); double* q = mem;
  w.X = q; q += N * n; w.U = q; q += N * m; w.L = q; q += p * K * n;
  w.Xt = q; q += N * n; w.Ut = q; q += N * m; w.Lt = q; q += p * K * n;
  w.dX = q; q += N * n; w.dU = q; q += N * m; w.dL = q; q += p * K * n;
  w.lam = q; q += K * (w.nrow + 1); w.mu = q; q += K * (w.nrow + 1); w.res = q; q += S; w.rhs = q; q += S; w.band = q;
  for (;;) {
    int inst = atomic_fetch_add(&jb->next, 1);
    if (inst >= jb->batch) break;
    w.xf = jb->xf + (size_t)inst * n; w.Q = jb->Q + (size_t)inst * n; w.R = jb->R + (size_t)inst * m; w.uf = jb->uf + (size_t)inst * m;
    const double* z0 = jb->Z0 + (size_t)inst * N * (n + m);
    for (int k = 0; k < N; k++) { memcpy(w.X + k * n, z0 + k * (n + m), n * sizeof(double)); memcpy(w.U + k * m, z0 + k * (n + m) + n, m * sizeof(double)); }
    memcpy(w.L, jb->L0 + (size_t)inst * p * K * n, (size_t)p * K * n * sizeof(double));
    memcpy(w.X, jb->x0 + (size_t)inst * n, n * sizeof(double));
    for (int t = 0; t < K * w.nrow; t++) { w.lam[t] = 0; w.mu[t] = o->rho_0; }
    if (jb->lam0) memcpy(w.lam, jb->lam0 + (size_t)inst * K * w.nrow, (size_t)K * w.nrow * sizeof(double));
    if (jb->mu0) memcpy(w.mu, jb->mu0 + (size_t)inst * K * w.nrow, (size_t)K * w.nrow * sizeof(double));
    w.hist = jb->hist ? jb->hist + (size_t)inst * jb->hist_max * AGB_NHIST : NULL; w.hist_max = jb->hist_max;
    jb->status[inst] = solve_one(&w, jb->stats + (size_t)inst * AGB_NSTATS);
    if (jb->hist_count) jb->hist_count[inst] = w.n_rec;
    if (jb->lam_out) memcpy(jb->lam_out + (size_t)inst * K * w.nrow, w.lam, (size_t)K * w.nrow * sizeof(double));
    if (jb->mu_out) memcpy(jb->mu_out + (size_t)inst * K * w.nrow, w.mu, (size_t)K * w.nrow * sizeof(double));
    double* z = jb->Z + (size_t)inst * N * (n + m);
    for (int k = 0; k < N; k++) { memcpy(z + k * (n + m), w.X + k * n, n * sizeof(double)); memcpy(z + k * (n + m) + n, w.U + k * m, m * sizeof(double)); }
    memcpy(jb->L + (size_t)inst * p * K * n, w.L, (size_t)p * K * n * sizeof(double));
  }
  free(mem);
  return NULL;
}

static int run_job(Job* jb, int nthreads) {
  if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (nthreads > jb->batch) nthreads = jb->batch;
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, worker, jb);
  worker(jb);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th);
  return nthreads;
}

int ago_newton_solve(const agb_problem_desc* d, const agb_options* o, int batch, int nthreads,
                     const double* x0, const double* xf, const double* Q, const double* R, const double* uf,
                     const double* Z0, const double* L0, double* Z, double* L, double* stats, int* status) {
  Job jb = {d, o, batch, x0, xf, Q, R, uf, Z0, L0, Z, L, stats, status, 0, NULL, NULL, NULL, NULL, NULL, NULL, 0};
  return run_job(&jb, nthreads);
}

/* Same, with the AL multipliers / penalties going in (warm start, dual_reset = false) and coming out, and the optional
 * per-record history; any of lam0, mu0, lam_out, mu_out, hist, hist_count may be NULL. */
int ago_newton_solve_ex(const agb_problem_desc* d, const agb_options* o, int batch, int nthreads,
                        const double* x0, const double* xf, const double* Q, const double* R, const double* uf,
                        const double* Z0, const double* L0, const double* lam0, const double* mu0,
                        double* Z, double* L, double* lam_out, double* mu_out, double* stats, int* status,
                        double* hist, int* hist_count, int hist_max) {
  Job jb = {d, o, batch, x0, xf, Q, R, uf, Z0, L0, Z, L, stats, status, 0, lam0, mu0, lam_out, mu_out, hist, hist_count, hist_max};
  return run_job(&jb, nthreads);
}

int ago_nrow(const agb_problem_desc* d) {
  Work w; memset(&w, 0, sizeof w);
  w.d = d; w.p = d->p; w.n = 4 * d->p; w.m = 2 * d->p;
  setup_rows(&w);
  return w.nrow;
}
