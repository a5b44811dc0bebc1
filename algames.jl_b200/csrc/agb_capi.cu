// agb_capi.cu — kernels + C ABI of libalgames_b200.so (see include/algames_b200.h for the contract).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <new>
#include <cstdlib>
#include <stddef.h>
#include <string>
#include "agb_kernels.cuh"
#include "agb_active_set.cuh"

using namespace agb;

namespace agb {   // agb_band.cu
void launch_band_solve(const DevDesc* dd, const agb_options& o, const Buffers& g, int inst0, int batch, int only_status, int grid, cudaStream_t st);
void launch_band_op(const DevDesc* dd, const agb_options& o, const Buffers& g, const OpArgs& a, int batch, int grid, cudaStream_t st);
size_t band_scratch_doubles(const DevDesc& d);
size_t band_window_bytes(const DevDesc& d, size_t limit);
int band_prepare(size_t bytes, int device);
}

// internal stage-major layout → reference row ("vertical", mode 0) or column ("horizontal", mode 1) order
__global__ void agb_export_kernel(const double* __restrict__ in, double* __restrict__ out, int batch, int P, int K, int mode) {
  const int n = 4 * P, m = 2 * P, b = P * n + m + n, Sz = K * b;
  const size_t total = (size_t)batch * Sz;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int q = (int)(t % Sz);
    const size_t inst = t / Sz;
    const int s = q / b, e = q - s * b;
    int dst;
    if (e < P * n) {                 // rx(i,s,a): row (opt_i x_{s+2}) / column λ_{i,s+1}
      const int i = e / n, a = e - i * n;
      dst = (mode == 0) ? (i * K + s) * (n + 2) + a : s * b + n + m + i * n + a;
    } else if (e < P * n + m) {      // ru(s, idx=(j,i))
      const int idx = e - P * n, j = idx / P, i = idx - j * P;
      dst = (mode == 0) ? (i * K + s) * (n + 2) + n + j : s * b + n + i * 2 + j;
    } else {                         // rd(s,a): row dyn_s / column x_{s+2}
      const int a = e - P * n - m;
      dst = (mode == 0) ? P * K * (n + 2) + s * n + a : s * b + a;
    }
    out[inst * Sz + dst] = in[t];
  }
}

// init_traj! with shift s (primal_dual_traj.jl:29-44) on device: the resident solution moves s knots earlier,
// the tail comes from the fresh arrays (zeros if null).
__global__ void agb_shift_kernel(const double* __restrict__ Z, const double* __restrict__ L, const double* __restrict__ Zf,
                                 const double* __restrict__ Lf, double* __restrict__ Z0, double* __restrict__ L0,
                                 int batch, int P, int n, int m, int N, int shift) {
  const int K = N - 1, zs = N * (n + m), ls = P * K * n;
  const size_t total = (size_t)batch * (zs + ls);
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t inst = t / (zs + ls);
    const int q = (int)(t % (zs + ls));
    if (q < zs) {
      const int k = q / (n + m), e = q - k * (n + m);
      const size_t o = inst * zs;
      Z0[o + q] = (k + shift < N) ? Z[o + (size_t)(k + shift) * (n + m) + e] : (Zf ? Zf[o + q] : 0.0);
    } else {
      const int r = q - zs, i = r / (K * n), rem = r - i * K * n, k = rem / n, e = rem - k * n;
      const size_t o = inst * ls;
      L0[o + r] = (k + shift < K) ? L[o + (size_t)(i * K + k + shift) * n + e] : (Lf ? Lf[o + r] : 0.0);
    }
  }
}

// x0 <- x_{1+shift} of the resident solution (+ disturbance)
__global__ void agb_advance_kernel(const double* __restrict__ Z, const double* __restrict__ dist, double* __restrict__ x0,
                                   int batch, int n, int m, int N, int shift) {
  const size_t total = (size_t)batch * n;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t inst = t / n; const int a = (int)(t % n);
    x0[t] = Z[(inst * N + shift) * (n + m) + a] + (dist ? dist[t] : 0.0);
  }
}

__global__ void agb_broadcast_kernel(double* dst, const double* src, int batch, int len) {
  const size_t total = (size_t)batch * len;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) dst[t] = src[t % len];
}
__global__ void agb_fill_kernel(double* dst, double v, size_t total) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) dst[t] = v;
}

// FP64 vector peak: 8 independent DFMA chains per thread, everything in registers (agb_measure_fp64_peak)
__global__ void __launch_bounds__(1024, 2) agb_fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1e-3, x2 = x0 + 2e-3, x3 = x0 + 3e-3, x4 = x0 + 4e-3, x5 = x0 + 5e-3, x6 = x0 + 6e-3, x7 = x0 + 7e-3;
#pragma unroll 1
  for (int it = 0; it < iters; it += 8) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (r == 12345.678) out[0] = r;          // never true: keeps the chains alive
}

// =============================================================================================================
// Host side
// =============================================================================================================
struct agb_handle {
  int device = 0, batch = 0;
  agb_problem_desc desc;
  DevDesc hd;
  DevDesc* dd = nullptr;
  cudaStream_t stream = nullptr;
  static constexpr int kMaxChunks = 32;
  cudaStream_t chunk_stream[kMaxChunks] = {};   // agb_solve_from_host pipeline
#ifndef AGB_EMULATE
  // agb_solve_from_host as a CUDA graph: the ~11 copies / launches per chunk are captured once per (buffers, options, chunking)
  // and replayed with one cudaGraphLaunch, which takes the host's API time off the step and lets the pipeline run finer chunks
  struct HostKey { const void* ptr[9]; agb_options o; int chunks; int epoch; };
  HostKey fh_key{}, fh_seen{};
  bool fh_valid = false, fh_seen_valid = false;
  cudaGraphExec_t fh_exec = nullptr;
  cudaEvent_t fh_fork = nullptr, fh_join[kMaxChunks] = {};
  int fh_launches = 0, fh_epoch = 0;
#endif
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  // device buffers
  double *x0 = nullptr, *xf = nullptr, *Q = nullptr, *R = nullptr, *uf = nullptr;
  double *Z0 = nullptr, *L0 = nullptr, *Z = nullptr, *L = nullptr, *conlam = nullptr, *conmu = nullptr, *D = nullptr, *KUg = nullptr, *stats = nullptr;
  int* status = nullptr;
  double* hist = nullptr; int* hist_count = nullptr; int hist_max = 0;   // agb_set_history
  double* Hpg = nullptr; int hpg_stride = 0;                              // big layout: pair / self Hessian blocks
  double* results = nullptr; size_t results_doubles = 0;   // owns Z, L, stats, status
  // band solver (agb_band.cuh): scratch slots (one per resident CTA); conlam0 / conmu0 keep the multipliers a warm-started
  // solve began with, for the fallback re-solve of AGB_SINGULAR instances
  double* band = nullptr; size_t band_stride = 0; int band_slots = 0; int band_win = 0;
  double *conlam0 = nullptr, *conmu0 = nullptr;
  bool fallback = true;
  int force_singular = 0;
  double* stage = nullptr; size_t stage_bytes = 0;     // scratch for exported outputs
  double* stage2 = nullptr; size_t stage2_bytes = 0;
  double* stage3 = nullptr;                            // disturbance staging of agb_mpc_advance
  double* stage_mpc = nullptr; size_t stage_mpc_bytes = 0;   // agb_mpc_run: disturbances in, per-re-solve records out
  long long launches = 0;
  size_t smem_bytes = 0;
  cudaEvent_t ev_async = nullptr;                      // orders agb_newton_solve_async (caller stream) against h->stream
  // multi-GPU gather (agb_peer_* / agb_allgather)
  int nranks = 0, rank = -1;
  size_t slab_off[AGB_MAX_RANKS + 1] = {};             // byte offset of every rank's result slab in a gather buffer
  unsigned char* gather = nullptr;                     // this rank's gather buffer [slab_off[nranks]]
  unsigned char* peer_gather[AGB_MAX_RANKS] = {};      // every rank's gather buffer as seen from this process / device
  int peer_device[AGB_MAX_RANKS] = {};                 // in-process peers: their device ordinal; -1 = IPC-mapped pointer
  bool peer_ipc_opened[AGB_MAX_RANKS] = {};
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_ready = nullptr;
  std::string err;
};

static thread_local std::string g_create_err;

static int fail(agb_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_err = msg;
  return code;
}
#define AGB_CUDA(h, call)                                                                         \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) return fail(h, AGB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

static int build_dev_desc(const agb_problem_desc* d, DevDesc* o, std::string* why) {
  memset(o, 0, sizeof(*o));
  if (d->p < 1 || d->p > AGB_MAX_P) { *why = "p must be in 1..4"; return AGB_EINVAL; }
  if (d->model < 0 || d->model > 3) { *why = "unknown model"; return AGB_EINVAL; }
  if (d->solver != AGB_SOLVER_AUTO && d->solver != AGB_SOLVER_BAND) { *why = "unknown solver"; return AGB_EINVAL; }
  if (d->model == AGB_MODEL_DOUBLE_INTEGRATOR && d->d != 2 && d->d != 3) { *why = "DoubleIntegratorGame: d must be 2 or 3"; return AGB_EUNSUPPORTED; }
  if (d->N < 2) { *why = "N must be >= 2"; return AGB_EINVAL; }
  if (!(d->dt > 0)) { *why = "dt must be positive"; return AGB_EINVAL; }
  const bool di3 = d->model == AGB_MODEL_DOUBLE_INTEGRATOR && d->d == 3;          // 3-D double integrator: band solver
  const int ni = d->model == AGB_MODEL_QUADROTOR ? 12 : (di3 ? 6 : 4), mi = d->model == AGB_MODEL_QUADROTOR ? 4 : (di3 ? 3 : 2);
  const int p = d->p, n = ni * p, m = mi * p;
  o->ni = ni; o->mi = mi;
  o->model = d->model; o->p = p; o->n = n; o->m = m; o->N = d->N; o->K = d->N - 1;
  o->quad_mass = d->quad_mass; o->spherical = d->spherical_collision ? 1 : 0;
  if (d->model == AGB_MODEL_QUADROTOR && !(d->quad_mass > 0)) { *why = "QuadrotorGame: mass must be positive"; return AGB_EINVAL; }
  o->use_band = (d->model == AGB_MODEL_QUADROTOR || di3 || d->spherical_collision || d->solver == AGB_SOLVER_BAND) ? 1 : 0;
  for (int i = 0; i < p; i++) if (d->n_walls3d[i] > 0 || d->n_cylinders[i] > 0) o->use_band = 1;
  o->kl = (2 * n - 1 > n + m) ? 2 * n - 1 : n + m;
  o->ku = (p * n + n - 1 > p * n + m) ? p * n + n - 1 : p * n + m;
  o->wd = 2 * o->kl + o->ku + 1;
  o->b = p * n + m + n; o->S = o->K * o->b;
  o->dt = d->dt; o->lf = d->lf; o->lr = d->lr;
  if (d->model == AGB_MODEL_BICYCLE && !(d->lr > 0)) { *why = "BicycleGame: lr must be positive"; return AGB_EINVAL; }
  o->has_cc = d->has_collision_cost ? 1 : 0;
  o->npairs = p * (p - 1);
  for (int i = 0; i < p; i++) { o->cc_radius[i] = d->cc_radius[i]; o->cc_mu[i] = d->cc_mu[i]; }
  int row = 0, ncw = 0;
  for (int i = 0; i < AGB_MAX_P; i++) {
    for (int j = 0; j < AGB_MAX_P; j++) o->col_row[i][j] = -1;
    for (int a = 0; a < AGB_MAX_N; a++) { o->sbmax_row[i][a] = -1; o->sbmin_row[i][a] = -1; }
  }
  for (int i = 0; i < p; i++) {
    o->srow_off[i] = row;
    for (int j = 0; j < p; j++) {
      if (j != i && d->col_radius[i][j] > 0) { o->col_radius[i][j] = d->col_radius[i][j]; o->col_row[i][j] = row++; }
    }
    o->sb_shift[i] = row - ncw;                 // this player's state-bound rows (if any) start here
    if (d->has_state_bound[i]) {
      const int ncon = d->has_state_bound[i];
      if (ncon < 0 || ncon > 2 * AGB_MAX_N) { *why = "has_state_bound: conval count out of range"; return AGB_EINVAL; }
      for (int a = 0; a < n; a++) {
        const bool fmx = isfinite(d->x_max[i][a]), fmn = isfinite(d->x_min[i][a]);
        if ((fmx && (d->x_max_con[i][a] < 0 || d->x_max_con[i][a] >= ncon)) || (fmn && (d->x_min_con[i][a] < 0 || d->x_min_con[i][a] >= ncon))) {
          *why = "x_max_con / x_min_con: conval index out of range"; return AGB_EINVAL;
        }
        // the reference checks x_max >= x_min inside one StateBoundConstraint (state_bound_constraint.jl:62-68)
        const bool same = !(fmx && fmn) || d->x_max_con[i][a] == d->x_min_con[i][a];
        if (same && !(d->x_max[i][a] >= d->x_min[i][a])) { *why = "Upper bounds must be greater than or equal to lower bounds"; return AGB_EINVAL; }
      }
      o->sb_ncon[i] = ncon;
      for (int a = 0; a < n; a++) {
        o->has_x_max[i][a] = isfinite(d->x_max[i][a]) ? 1 : 0; o->has_x_min[i][a] = isfinite(d->x_min[i][a]) ? 1 : 0;
        o->x_max_con[i][a] = d->x_max_con[i][a]; o->x_min_con[i][a] = d->x_min_con[i][a];
      }
      const int row0 = row;
      for (int g = 0; g < ncon; g++) {           // conval order, each: finite x_max rows then finite x_min rows
        for (int a = 0; a < n; a++) if (isfinite(d->x_max[i][a]) && d->x_max_con[i][a] == g) { o->x_max[i][a] = d->x_max[i][a]; o->sbmax_row[i][a] = row++; }
        for (int a = 0; a < n; a++) if (isfinite(d->x_min[i][a]) && d->x_min_con[i][a] == g) { o->x_min[i][a] = d->x_min[i][a]; o->sbmin_row[i][a] = row++; }
      }
      ncw += row - row0;
    }
    if (d->n_walls[i] < 0 || d->n_walls[i] > AGB_MAX_WALLS || d->n_circles[i] < 0 || d->n_circles[i] > AGB_MAX_CIRCLES) {
      *why = "too many walls / circles"; return AGB_EINVAL;
    }
    o->n_walls[i] = d->n_walls[i]; o->wall_row[i] = row; row += d->n_walls[i];
    for (int q = 0; q < d->n_walls[i]; q++) for (int e = 0; e < 6; e++) o->walls[i][q][e] = d->walls[i][q][e];
    o->n_circles[i] = d->n_circles[i]; o->circle_row[i] = row; row += d->n_circles[i];
    for (int q = 0; q < d->n_circles[i]; q++) for (int e = 0; e < 3; e++) o->circles[i][q][e] = d->circles[i][q][e];
    if (d->n_walls3d[i] < 0 || d->n_walls3d[i] > AGB_MAX_WALLS || d->n_cylinders[i] < 0 || d->n_cylinders[i] > AGB_MAX_WALLS) {
      *why = "too many 3-D walls / cylinders"; return AGB_EINVAL;
    }
    if ((d->n_walls3d[i] > 0 || d->n_cylinders[i] > 0) && ni < 3) { *why = "3-D walls and cylinders need three position components"; return AGB_EINVAL; }
    o->n_walls3d[i] = d->n_walls3d[i]; o->wall3d_row[i] = row; row += d->n_walls3d[i];
    for (int q = 0; q < d->n_walls3d[i]; q++) for (int e = 0; e < 12; e++) o->walls3d[i][q][e] = d->walls3d[i][q][e];
    o->n_cyl[i] = d->n_cylinders[i]; o->cyl_row[i] = row; row += d->n_cylinders[i];
    for (int q = 0; q < d->n_cylinders[i]; q++) {
      for (int e = 0; e < 6; e++) o->cyl[i][q][e] = d->cylinders[i][q][e];
      const int ax = (int)d->cylinders[i][q][3];
      if (ax < 0 || ax > 2 || (double)ax != d->cylinders[i][q][3]) { *why = "cylinder axis must be 0, 1 or 2"; return AGB_EINVAL; }
    }
  }
  for (int i = p; i <= AGB_MAX_P; i++) o->srow_off[i] = row;
  o->nrow_state = row;
  o->has_pairs = o->has_cc; o->has_self = 0; o->has_sb = 0; o->has_cb = d->has_control_bound ? 1 : 0;
  for (int i = 0; i < p; i++) if (d->has_state_bound[i]) o->has_sb = 1;
  for (int i = 0; i < p; i++) {
    for (int j = 0; j < p; j++) if (o->col_row[i][j] >= 0) o->has_pairs = 1;
    if (o->n_walls[i] > 0 || o->n_circles[i] > 0) o->has_self = 1;
  }
  for (int idx = 0; idx < AGB_MAX_M; idx++) { o->ub_row[idx] = -1; o->lb_row[idx] = -1; }
  o->cb_shift = row - ncw;
  if (d->has_control_bound) {
    for (int idx = 0; idx < m; idx++) {
      if (!(d->u_max[idx] >= d->u_min[idx])) { *why = "Upper bounds must be greater than or equal to lower bounds"; return AGB_EINVAL; }
    }
    for (int idx = 0; idx < m; idx++) if (isfinite(d->u_max[idx])) { o->u_max[idx] = d->u_max[idx]; o->ub_row[idx] = row++; }
    for (int idx = 0; idx < m; idx++) if (isfinite(d->u_min[idx])) { o->u_min[idx] = d->u_min[idx]; o->lb_row[idx] = row++; }
  }
  o->nrow = row;
  o->nrow_control = row - o->nrow_state;
  ncw += o->nrow_control;
  o->ncw = ncw;
  // shared-memory layout: small (everything resident) or big (L, CL, CM, Hp, Hs in global memory, agb_internal.h)
  const int N = o->N, K = o->K, W = m + n + 1;
  auto layout = [&](bool big) {
    int off = 0;
    auto take = [&](int cnt) { int r = off; off += (cnt + 1) & ~1; return r; };
    o->o_X = take(N * n); o->o_U = take(N * m); o->o_L = take(big ? 0 : p * K * n); o->o_R = take(K * o->b);
    o->o_AB = take(d->model == AGB_MODEL_DOUBLE_INTEGRATOR ? 0 : K * p * 16);
    o->o_CL = take(big ? 0 : K * o->nrow); o->o_CM = take(big ? 0 : K * o->nrow);
    o->o_CW = take((o->has_sb || o->has_cb) ? K * o->ncw : 0);     // pair / wall / circle weights are folded into Hp / Hs
    o->o_Hp = take((o->has_pairs && !big) ? N * o->npairs * 3 : 0); o->o_Hs = take((o->has_self && !big) ? N * p * 3 : 0);
    o->o_P = take(p * n * n); o->o_Sv = take(p * n); o->o_Y = take(m * (n + 1)); o->o_Aug = take(m * W);
    o->o_KU = o->o_Aug;                          // one-stage gain buffer of the best-response factorisation (Aug is unused there)
    o->o_Base = take(p * n * (n + 1)); o->o_W = take(p * n * m); o->o_Ta = take(2 * p * (4 * p * p + n));   // Hm: double-buffered per-stage H^x blocks (Inst::HmS per player)
    {  // Gp / Gs live only inside one residual evaluation, the factorisation scratch (P … Hm) only inside kkt_solve: alias them
      const int need = (o->has_pairs ? N * o->npairs * 2 : 0) + (o->has_self ? N * p * 2 : 0);
      const int have = off - o->o_P;
      if (need > have) take(need - have);
      o->o_Gp = o->o_P; o->o_Gs = o->o_P + (o->has_pairs ? N * o->npairs * 2 : 0);
    }
    // the forward sweep lays a ring of fwd_ring_depth(p) closed-loop blocks (Inst::ACS = n*(n+2) doubles) and its flags
    // over [o_P, end of Hm)
    {
      const int need = fwd_ring_depth(p) * n * (n + 2) + 8;
      if (off - o->o_P < need) take(need - (off - o->o_P));
    }
    o->o_par = take(2 * n + 2 * m); o->o_red = take(6 * kMaxWarps);
    o->smem_doubles = off;
    o->big = big ? 1 : 0;
  };
  if (o->use_band) { o->smem_doubles = 0; o->big = 0; return AGB_OK; }     // no structured-kernel form: band solver only
  layout(false);
  // layout choice (DevDesc::big: 0 small / big storage with 1: 2, 2: 4, 3: 3 CTAs per SM):
  //   small if it gives 4 CTAs per SM; else big storage with as many CTAs per SM as its footprint allows (4 at 128
  //   registers, 3 at 168, 2 at 255).  4-player kernels exist only as layout 1.
  const size_t four_per_sm = (228u * 1024u - 4u * 1024u) / 4u;
  const char* force = getenv("AGB_FORCE_BIG_LAYOUT");         // test hook: "1" / "2" / "3" force that layout on 3-player instances
  const size_t small_bytes = (size_t)o->smem_doubles * sizeof(double);
  if (p >= 4) { layout(true); o->big = 1; }
  else if (p == 3 && force && force[0] >= '1' && force[0] <= '3') { layout(true); o->big = force[0] - '0'; }
  else if (p == 3 && small_bytes > four_per_sm) {
    layout(true);
    const size_t big_bytes = (size_t)o->smem_doubles * sizeof(double), three_per_sm = (228u * 1024u - 3u * 1024u) / 3u;
    o->big = big_bytes <= four_per_sm ? 2 : (big_bytes <= three_per_sm ? 3 : 1);
  }
  return AGB_OK;
}

namespace agb {
cudaError_t set_attr(int p, int big, int model, size_t smem) {
  switch (p) { case 1: return set_attr_p1(model, smem); case 2: return set_attr_p2(model, smem);
               case 3: return big == 1 ? set_attr_p3b(model, smem) : (big == 2 ? set_attr_p3m(model, smem) : (big == 3 ? set_attr_p3t(model, smem) : set_attr_p3(model, smem))); default: return set_attr_p4(model, smem); }
}
void launch_solve(int p, int big, const LaunchArgs& L) {
  switch (p) { case 1: launch_solve_p1(L); break; case 2: launch_solve_p2(L); break;
               case 3: if (big == 1) launch_solve_p3b(L); else if (big == 2) launch_solve_p3m(L); else if (big == 3) launch_solve_p3t(L); else launch_solve_p3(L); break; default: launch_solve_p4(L); }
}
void launch_ibr(int p, int big, const LaunchArgs& L) {
  switch (p) { case 1: launch_ibr_p1(L); break; case 2: launch_ibr_p2(L); break;
               case 3: if (big == 1) launch_ibr_p3b(L); else if (big == 2) launch_ibr_p3m(L); else if (big == 3) launch_ibr_p3t(L); else launch_ibr_p3(L); break; default: launch_ibr_p4(L); }
}
void launch_op(int p, int big, const LaunchArgs& L) {
  switch (p) { case 1: launch_op_p1(L); break; case 2: launch_op_p2(L); break;
               case 3: if (big == 1) launch_op_p3b(L); else if (big == 2) launch_op_p3m(L); else if (big == 3) launch_op_p3t(L); else launch_op_p3(L); break; default: launch_op_p4(L); }
}
}  // namespace agb

extern "C" {

void agb_default_options(agb_options* o) {   // src/struct/options.jl:5-116
  memset(o, 0, sizeof(*o));
  o->reg_0 = 1e-3; o->regularize = 1; o->alpha_decrease = 0.5; o->beta = 0.01; o->ls_iter = 25; o->delta_min = 1e-9;
  o->rho_0 = 1.0; o->rho_increase = 10.0; o->rho_max = 1e7; o->lambda_max = 1e7; o->alpha_dual = 1.0;
  for (int i = 0; i < AGB_MAX_P; i++) o->alphax_dual[i] = 1.0;
  o->active_set_tolerance = 1e-4;
  o->eps_dyn = o->eps_sta = o->eps_con = o->eps_opt = 1e-3;
  o->outer_iter = 7; o->inner_iter = 20; o->dual_reset = 1;
}

int agb_sizes_of(const agb_problem_desc* d, agb_sizes* out) {
  if (!d || !out) return fail(nullptr, AGB_EINVAL, "null argument");
  DevDesc t; std::string why;
  int rc = build_dev_desc(d, &t, &why);
  if (rc) return fail(nullptr, rc, why);
  out->n = t.n; out->m = t.m; out->p = t.p; out->N = t.N; out->S = t.S;
  out->nrow = t.nrow; out->nrow_state = t.nrow_state; out->nrow_control = t.nrow_control;
  return AGB_OK;
}

const char* agb_last_error(const agb_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }


void agb_destroy(agb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  void* ptrs[] = {h->dd, h->x0, h->xf, h->Q, h->R, h->uf, h->Z0, h->L0, h->results, h->conlam, h->conmu, h->D, h->KUg, h->stage, h->stage2, h->stage3, h->stage_mpc, h->hist, h->hist_count, h->Hpg, h->band, h->conlam0, h->conmu0};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (int r = 0; r < h->nranks; r++) if (h->peer_ipc_opened[r]) cudaIpcCloseMemHandle(h->peer_gather[r]);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  if (h->gather) cudaFree(h->gather);
  if (h->ev_ready) cudaEventDestroy(h->ev_ready);
  if (h->ev_async) cudaEventDestroy(h->ev_async);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (cudaStream_t cs : h->chunk_stream) if (cs) cudaStreamDestroy(cs);
#ifndef AGB_EMULATE
  if (h->fh_exec) cudaGraphExecDestroy(h->fh_exec);
  if (h->fh_fork) cudaEventDestroy(h->fh_fork);
  for (cudaEvent_t e : h->fh_join) if (e) cudaEventDestroy(e);
#endif
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

static int alloc_d(agb_handle* h, double** p, size_t count) {
  AGB_CUDA(h, cudaMalloc((void**)p, (count ? count : 1) * sizeof(double)));
  AGB_CUDA(h, cudaMemsetAsync(*p, 0, (count ? count : 1) * sizeof(double), h->stream));
  return AGB_OK;
}

static inline int grid_for(size_t total) { size_t g = (total + 255) / 256; return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g)); }

int agb_create(const agb_problem_desc* desc, int batch, int device, agb_handle** out) {
  if (!desc || !out) return fail(nullptr, AGB_EINVAL, "null argument");
  *out = nullptr;
  if (batch < 1) return fail(nullptr, AGB_EINVAL, "batch must be >= 1");
  DevDesc t; std::string why;
  int rc = build_dev_desc(desc, &t, &why);
  if (rc) return fail(nullptr, rc, why);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(nullptr, AGB_ECUDA, "no CUDA device: libalgames_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(nullptr, AGB_EINVAL, "bad device ordinal");
  agb_handle* h = new (std::nothrow) agb_handle();
  if (!h) return fail(nullptr, AGB_ENOMEM, "out of host memory");
  h->device = device; h->batch = batch; h->desc = *desc; h->hd = t;
  h->smem_bytes = (size_t)t.smem_doubles * sizeof(double);
#define CK(call) do { int rc_ = (call); if (rc_) { g_create_err = h->err; agb_destroy(h); return rc_; } } while (0)
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_create_err = std::string(#call) + ": " + cudaGetErrorString(e_); agb_destroy(h); return AGB_ECUDA; } } while (0)
  CKC(cudaSetDevice(device));
  int max_smem = 0;
  CKC(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  if (!t.use_band && (int)h->smem_bytes > max_smem) {
    char buf[160];
    snprintf(buf, sizeof buf, "instance needs %zu B of shared memory per CTA, device allows %d B", h->smem_bytes, max_smem);
    g_create_err = buf; agb_destroy(h); return AGB_EUNSUPPORTED;
  }
  // the opt-in limit is a per-kernel attribute shared by every handle using the same template instance: always raise it to
  // the device maximum, so a later handle with a smaller footprint can never lower it under an earlier one's launches
  if (!t.use_band) CKC(agb::set_attr(t.p, t.big, t.model, (size_t)max_smem));
  CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CKC(cudaEventCreate(&h->ev0));
  CKC(cudaEventCreate(&h->ev1));
  CKC(cudaEventCreateWithFlags(&h->ev_async, cudaEventDisableTiming));
  CKC(cudaMalloc((void**)&h->dd, sizeof(DevDesc)));
  CKC(cudaMemcpyAsync(h->dd, &h->hd, sizeof(DevDesc), cudaMemcpyHostToDevice, h->stream));
  const size_t B = batch, n = t.n, m = t.m, N = t.N, K = t.K, p = t.p;
  CK(alloc_d(h, &h->x0, B * n)); CK(alloc_d(h, &h->xf, B * n)); CK(alloc_d(h, &h->Q, B * n));
  CK(alloc_d(h, &h->R, B * m)); CK(alloc_d(h, &h->uf, B * m));
  CK(alloc_d(h, &h->Z0, B * N * (n + m))); CK(alloc_d(h, &h->L0, B * p * K * n));
  {  // results slab: [Z | L | stats | status] in one allocation (one collective gathers everything)
    const size_t zs = B * N * (n + m), ls = B * p * K * n, ss = B * AGB_NSTATS, is = (B + 1) / 2;
    h->results_doubles = zs + ls + ss + is;
    CK(alloc_d(h, &h->results, h->results_doubles));
    h->Z = h->results; h->L = h->Z + zs; h->stats = h->L + ls; h->status = (int*)(h->stats + ss);
  }
  CK(alloc_d(h, &h->conlam, B * K * t.nrow)); CK(alloc_d(h, &h->conmu, B * K * t.nrow));
  CK(alloc_d(h, &h->D, B * t.S)); CK(alloc_d(h, &h->KUg, B * K * m * (n + 2)));       // rows padded to n+2 (Inst::n1p)
  if (t.big) { h->hpg_stride = N * t.npairs * 3 + N * p * 3; CK(alloc_d(h, &h->Hpg, B * (size_t)h->hpg_stride)); }
  {  // band solver scratch: one slot per resident CTA — the whole device for band-only schemas, a few slots for the fallback
    const char* e = getenv("AGB_BAND_FALLBACK");         // "0": AGB_SINGULAR instances of the structured kernels are returned as they are
    h->fallback = !(e && e[0] == '0');
    if (const char* f = getenv("AGB_TEST_FORCE_SINGULAR")) h->force_singular = atoi(f);
    if (t.use_band || h->fallback) {
      int sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
      const char* w = getenv("AGB_BAND_WINDOW");           // "0": eliminate in the global band (the first form of the solver; A/B and tests)
      h->band_win = (w && w[0] == '0') ? 0 : (int)agb::band_window_bytes(t, (size_t)max_smem - 1024);
      const int occ = agb::band_prepare((size_t)h->band_win, device);
      const int want = t.use_band ? occ * sms : 32;
      h->band_slots = batch < want ? batch : want;
      h->band_stride = (agb::band_scratch_doubles(t) + 1) & ~(size_t)1;
      CK(alloc_d(h, &h->band, (size_t)h->band_slots * h->band_stride));
      if (!t.use_band) { CK(alloc_d(h, &h->conlam0, B * K * t.nrow)); CK(alloc_d(h, &h->conmu0, B * K * t.nrow)); }
    }
  }
  // descriptor defaults broadcast to every instance; μ starts at 1 (Altro ALConVal default)
  double tmp[2 * AGB_MAX_N + 2 * AGB_MAX_M];
  double* dtmp = nullptr;
  CKC(cudaMalloc((void**)&dtmp, sizeof tmp));
  memcpy(tmp, desc->xf, n * sizeof(double)); memcpy(tmp + n, desc->Q, n * sizeof(double));
  memcpy(tmp + 2 * n, desc->R, m * sizeof(double)); memcpy(tmp + 2 * n + m, desc->uf, m * sizeof(double));
  CKC(cudaMemcpyAsync(dtmp, tmp, sizeof tmp, cudaMemcpyHostToDevice, h->stream));
  AGB_LAUNCH(agb_broadcast_kernel, grid_for(B * n), 256, 0, h->stream, h->xf, dtmp, batch, (int)n);
  AGB_LAUNCH(agb_broadcast_kernel, grid_for(B * n), 256, 0, h->stream, h->Q, dtmp + n, batch, (int)n);
  AGB_LAUNCH(agb_broadcast_kernel, grid_for(B * m), 256, 0, h->stream, h->R, dtmp + 2 * n, batch, (int)m);
  AGB_LAUNCH(agb_broadcast_kernel, grid_for(B * m), 256, 0, h->stream, h->uf, dtmp + 2 * n + m, batch, (int)m);
  if (t.nrow) AGB_LAUNCH(agb_fill_kernel, grid_for(B * K * t.nrow), 256, 0, h->stream, h->conmu, 1.0, B * K * t.nrow);
  h->launches += 5;
  CKC(cudaStreamSynchronize(h->stream));
  cudaFree(dtmp);
  CKC(cudaGetLastError());
#undef CK
#undef CKC
  *out = h;
  return AGB_OK;
}

int agb_get_sizes(const agb_handle* h, agb_sizes* out) {
  if (!h || !out) return AGB_EINVAL;
  out->n = h->hd.n; out->m = h->hd.m; out->p = h->hd.p; out->N = h->hd.N; out->S = h->hd.S;
  out->nrow = h->hd.nrow; out->nrow_state = h->hd.nrow_state; out->nrow_control = h->hd.nrow_control;
  return AGB_OK;
}

static Buffers buffers_of(agb_handle* h) {
  Buffers g;
  g.x0 = h->x0; g.xf = h->xf; g.Q = h->Q; g.R = h->R; g.uf = h->uf; g.Z0 = h->Z0; g.L0 = h->L0; g.Z = h->Z; g.L = h->L;
  g.conlam = h->conlam; g.conmu = h->conmu; g.D = h->D; g.KUg = h->KUg; g.stats = h->stats; g.status = h->status;
  g.hist = h->hist; g.hist_count = h->hist_count; g.hist_max = h->hist_max;
  g.Hpg = h->Hpg; g.hpg_stride = h->hpg_stride;
  g.band = h->band; g.band_stride = h->band_stride; g.band_slots = h->band_slots; g.band_win = h->band_win; g.conlam0 = nullptr; g.conmu0 = nullptr;
  g.force_singular = h->force_singular;
  return g;
}

static int h2d(agb_handle* h, double* dst, const double* src, size_t count) {
  if (!src || !count) return AGB_OK;
  AGB_CUDA(h, cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return AGB_OK;
}
static int d2h(agb_handle* h, double* dst, const double* src, size_t count) {
  if (!dst || !count) return AGB_OK;
  AGB_CUDA(h, cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  return AGB_OK;
}
#define AGB_TRY(call) do { int rc_ = (call); if (rc_) return rc_; } while (0)

int agb_set_instance_params(agb_handle* h, const double* x0, const double* xf, const double* Q, const double* R, const double* uf) {
  if (!h) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = h->batch, n = h->hd.n, m = h->hd.m;
  AGB_TRY(h2d(h, h->x0, x0, B * n)); AGB_TRY(h2d(h, h->xf, xf, B * n)); AGB_TRY(h2d(h, h->Q, Q, B * n));
  AGB_TRY(h2d(h, h->R, R, B * m)); AGB_TRY(h2d(h, h->uf, uf, B * m));
  AGB_CUDA(h, cudaStreamSynchronize(h->stream));
  return AGB_OK;
}

int agb_set_initial(agb_handle* h, const double* Z0, const double* L0, const double* conlam, const double* conmu) {
  if (!h) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = h->batch, zs = (size_t)h->hd.N * (h->hd.n + h->hd.m), ls = (size_t)h->hd.p * h->hd.K * h->hd.n, cs = (size_t)h->hd.K * h->hd.nrow;
  AGB_TRY(h2d(h, h->Z0, Z0, B * zs)); AGB_TRY(h2d(h, h->L0, L0, B * ls));
  if (Z0) AGB_CUDA(h, cudaMemcpyAsync(h->Z, h->Z0, B * zs * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  if (L0) AGB_CUDA(h, cudaMemcpyAsync(h->L, h->L0, B * ls * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  AGB_TRY(h2d(h, h->conlam, conlam, B * cs)); AGB_TRY(h2d(h, h->conmu, conmu, B * cs));
  AGB_CUDA(h, cudaStreamSynchronize(h->stream));
  return AGB_OK;
}

int agb_get_state(agb_handle* h, double* Z, double* L, double* conlam, double* conmu) {
  if (!h) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = h->batch, zs = (size_t)h->hd.N * (h->hd.n + h->hd.m), ls = (size_t)h->hd.p * h->hd.K * h->hd.n, cs = (size_t)h->hd.K * h->hd.nrow;
  AGB_TRY(d2h(h, Z, h->Z, B * zs)); AGB_TRY(d2h(h, L, h->L, B * ls));
  AGB_TRY(d2h(h, conlam, h->conlam, B * cs)); AGB_TRY(d2h(h, conmu, h->conmu, B * cs));
  AGB_CUDA(h, cudaStreamSynchronize(h->stream));
  return AGB_OK;
}

static int ensure_stage(agb_handle* h, double** buf, size_t* cur, size_t bytes) {
  if (*cur >= bytes) return AGB_OK;
  if (*buf) { cudaFree(*buf); *buf = nullptr; *cur = 0; }
  cudaError_t e = cudaMalloc((void**)buf, bytes);
  if (e != cudaSuccess) return fail(h, AGB_ENOMEM, std::string("cudaMalloc staging: ") + cudaGetErrorString(e));
  *cur = bytes;
  return AGB_OK;
}

int agb_shift_initial(agb_handle* h, int s, const double* Zfresh, const double* Lfresh) {
  if (!h || s < 0) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = h->batch, zs = (size_t)h->hd.N * (h->hd.n + h->hd.m), ls = (size_t)h->hd.p * h->hd.K * h->hd.n;
  double *zf = nullptr, *lf = nullptr;
  if (Zfresh) { AGB_TRY(ensure_stage(h, &h->stage, &h->stage_bytes, B * zs * sizeof(double))); zf = h->stage; AGB_TRY(h2d(h, zf, Zfresh, B * zs)); }
  if (Lfresh) { AGB_TRY(ensure_stage(h, &h->stage2, &h->stage2_bytes, B * ls * sizeof(double))); lf = h->stage2; AGB_TRY(h2d(h, lf, Lfresh, B * ls)); }
  AGB_LAUNCH(agb_shift_kernel, grid_for(B * (zs + ls)), 256, 0, h->stream, h->Z, h->L, zf, lf, h->Z0, h->L0, h->batch, h->hd.p, h->hd.n, h->hd.m, h->hd.N, s);
  h->launches++;
  AGB_CUDA(h, cudaMemcpyAsync(h->Z, h->Z0, B * zs * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  AGB_CUDA(h, cudaMemcpyAsync(h->L, h->L0, B * ls * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  AGB_CUDA(h, cudaStreamSynchronize(h->stream));
  AGB_CUDA(h, cudaGetLastError());
  return AGB_OK;
}

int agb_mpc_advance(agb_handle* h, int s, const double* disturbance, const double* Zfresh, const double* Lfresh) {
  if (!h || s < 1 || s >= h->hd.N) return fail(h, AGB_EINVAL, "shift must be in 1..N-1");
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = h->batch, n = h->hd.n;
  double* dd = nullptr;
  if (disturbance) {
    if (!h->stage3) AGB_CUDA(h, cudaMalloc((void**)&h->stage3, B * n * sizeof(double)));
    dd = h->stage3;
    AGB_TRY(h2d(h, dd, disturbance, B * n));
  }
  AGB_LAUNCH(agb_advance_kernel, grid_for(B * n), 256, 0, h->stream, h->Z, dd, h->x0, h->batch, h->hd.n, h->hd.m, h->hd.N, s);
  h->launches++;
  return agb_shift_initial(h, s, Zfresh, Lfresh);
}

// Device-resident variant: disturbance_dev is a DEVICE pointer [B][n] (or NULL), the tail of the shifted iterate is zero;
// nothing is copied from the host and nothing synchronises — 200 × (agb_newton_solve_async + agb_mpc_advance_async) is one
// uninterrupted stream of kernels.
int agb_mpc_advance_async(agb_handle* h, int s, const double* disturbance_dev) {
  if (!h || s < 1 || s >= h->hd.N) return fail(h, AGB_EINVAL, "shift must be in 1..N-1");
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = h->batch, n = h->hd.n, zs = (size_t)h->hd.N * (h->hd.n + h->hd.m), ls = (size_t)h->hd.p * h->hd.K * h->hd.n;
  AGB_LAUNCH(agb_advance_kernel, grid_for(B * n), 256, 0, h->stream, h->Z, disturbance_dev, h->x0, h->batch, h->hd.n, h->hd.m, h->hd.N, s);
  AGB_LAUNCH(agb_shift_kernel, grid_for(B * (zs + ls)), 256, 0, h->stream, h->Z, h->L, (const double*)nullptr, (const double*)nullptr, h->Z0, h->L0, h->batch, h->hd.p, h->hd.n, h->hd.m, h->hd.N, s);
  h->launches += 2;
  AGB_CUDA(h, cudaMemcpyAsync(h->Z, h->Z0, B * zs * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  AGB_CUDA(h, cudaMemcpyAsync(h->L, h->L0, B * ls * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  AGB_CUDA(h, cudaGetLastError());
  return AGB_OK;
}

static int launch_solve_range(agb_handle* h, const agb_options* o, cudaStream_t st, int lo, int hi, int slot0, int nslots, const MpcArgs* mp = nullptr);
static int launch_solve(agb_handle* h, const agb_options* o, cudaStream_t st);
static int finish(agb_handle* h);

// The whole receding-horizon loop in ONE launch of the solve kernel (MpcArgs, agb_kernels.cuh): stream b's CTA runs its `resolves`
// re-solves back to back, so a stream that needs 140 Newton steps at one re-solve delays nobody else — the step-wise loop pays the
// slowest stream at every step.  The last advance goes through agb_mpc_advance_async, which leaves the handle (Z, L, Z0, L0, x0)
// exactly as the step-wise loop does.  Band-solver schemas run the step-wise loop here, with the same outputs.
int agb_mpc_run_async(agb_handle* h, const agb_options* o, int resolves, int s, const double* disturbance_dev, double* stats_dev,
                      int* status_dev, double* xs_dev, void* stream) {
  if (!h || !o || !stats_dev || !status_dev) return AGB_EINVAL;
  if (resolves < 1) return fail(h, AGB_EINVAL, "resolves must be >= 1");
  if (s < 1 || s >= h->hd.N) return fail(h, AGB_EINVAL, "shift must be in 1..N-1");
  if (o->ls_iter < 1 || o->outer_iter < 1 || o->inner_iter < 1) return fail(h, AGB_EINVAL, "outer_iter, inner_iter, ls_iter must be >= 1");
  AGB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
  const size_t B = h->batch, n = h->hd.n;
  if (st != h->stream) {
    AGB_CUDA(h, cudaEventRecord(h->ev_async, h->stream));
    AGB_CUDA(h, cudaStreamWaitEvent(st, h->ev_async, 0));
  }
  if (h->hd.use_band) {
    if (st != h->stream) return fail(h, AGB_EUNSUPPORTED, "band-solver schemas run the receding-horizon loop on the handle's own stream");
    agb_options ow = *o;
    for (int t = 0; t < resolves; t++) {
      AGB_TRY(launch_solve(h, t == 0 ? o : &ow, st));
      ow.dual_reset = 0;
      AGB_CUDA(h, cudaMemcpyAsync(stats_dev + (size_t)t * B * AGB_NSTATS, h->stats, B * AGB_NSTATS * sizeof(double), cudaMemcpyDeviceToDevice, st));
      AGB_CUDA(h, cudaMemcpyAsync(status_dev + (size_t)t * B, h->status, B * sizeof(int), cudaMemcpyDeviceToDevice, st));
      if (t + 1 < resolves) {
        AGB_TRY(agb_mpc_advance_async(h, s, disturbance_dev ? disturbance_dev + (size_t)t * B * n : nullptr));
        if (xs_dev) AGB_CUDA(h, cudaMemcpyAsync(xs_dev + (size_t)t * B * n, h->x0, B * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
      }
    }
  } else {
    MpcArgs mp;
    mp.resolves = resolves; mp.shift = s; mp.dist = disturbance_dev; mp.stats = stats_dev; mp.status = status_dev; mp.xs = xs_dev;
    AGB_TRY(launch_solve_range(h, o, st, 0, h->batch, 0, h->band_slots, &mp));
  }
  if (st != h->stream) {
    AGB_CUDA(h, cudaEventRecord(h->ev_async, st));
    AGB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_async, 0));
  }
  AGB_TRY(agb_mpc_advance_async(h, s, disturbance_dev ? disturbance_dev + (size_t)(resolves - 1) * B * n : nullptr));   // on the handle's stream
  if (h->hd.use_band && xs_dev)
    AGB_CUDA(h, cudaMemcpyAsync(xs_dev + (size_t)(resolves - 1) * B * n, h->x0, B * n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  return AGB_OK;
}

// Host-buffer form: disturbance [resolves][B][n] (or NULL) in, per-re-solve stats / status / executed states out.
int agb_mpc_run(agb_handle* h, const agb_options* o, int resolves, int s, const double* disturbance, double* stats_out,
                int* status_out, double* xs_out) {
  if (!h || !o) return AGB_EINVAL;
  if (resolves < 1) return fail(h, AGB_EINVAL, "resolves must be >= 1");
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = h->batch, n = h->hd.n, R = (size_t)resolves;
  const size_t nd = disturbance ? R * B * n : 0, ns = R * B * AGB_NSTATS, nx = xs_out ? R * B * n : 0;
  const size_t status_doubles = (R * B * sizeof(int) + sizeof(double) - 1) / sizeof(double);
  AGB_TRY(ensure_stage(h, &h->stage_mpc, &h->stage_mpc_bytes, (nd + ns + nx + status_doubles) * sizeof(double)));
  double* dd = disturbance ? h->stage_mpc : nullptr;
  double* sd = h->stage_mpc + nd;
  double* xd = xs_out ? sd + ns : nullptr;
  int* td = (int*)(sd + ns + nx);
  AGB_TRY(h2d(h, dd, disturbance, nd));
  AGB_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  AGB_TRY(agb_mpc_run_async(h, o, resolves, s, dd, sd, td, xd, h->stream));
  AGB_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  h->timed = true;
  AGB_TRY(d2h(h, stats_out, sd, ns));
  AGB_TRY(d2h(h, xs_out, xd, nx));
  if (status_out) AGB_CUDA(h, cudaMemcpyAsync(status_out, td, R * B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  return finish(h);
}

void* agb_get_stream(agb_handle* h) { return h ? (void*)h->stream : nullptr; }

int agb_join_stream(agb_handle* h, void* stream) {
  if (!h) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  if ((cudaStream_t)stream == h->stream) return AGB_OK;
  AGB_CUDA(h, cudaEventRecord(h->ev_async, h->stream));
  AGB_CUDA(h, cudaStreamWaitEvent((cudaStream_t)stream, h->ev_async, 0));
  return AGB_OK;
}

static int launch_op(agb_handle* h, const agb_options* o, const OpArgs& a) {
  agb_options od;
  if (o) od = *o; else agb_default_options(&od);
  if (h->hd.use_band) {
    if (a.player >= 0 || !(a.op == OP_ROLLOUT || a.op == OP_RESIDUAL || a.op == OP_JAC_DENSE || a.op == OP_KKT_SOLVE || a.op == OP_EVAL_CON))
      return fail(h, AGB_EUNSUPPORTED, "this entry point is not available for schemas solved by the band solver (QuadrotorGame, 3-D constraints, AGB_SOLVER_BAND): "
                                       "use agb_rollout / agb_residual / agb_residual_jacobian_dense / agb_kkt_solve / agb_evaluate_constraints and the solves");
    const int grid = h->batch < h->band_slots ? h->batch : h->band_slots;
    agb::launch_band_op(h->dd, od, buffers_of(h), a, h->batch, grid, h->stream);
    h->launches++;
    AGB_CUDA(h, cudaGetLastError());
    return AGB_OK;
  }
  LaunchArgs L;
  L.model = h->hd.model; L.grid = h->batch; L.smem = h->smem_bytes; L.stream = h->stream; L.dd = h->dd; L.o = od;
  memset(&L.io, 0, sizeof L.io);
  L.g = buffers_of(h); L.a = a; L.batch = h->batch;
  agb::launch_op(h->hd.p, h->hd.big, L);
  h->launches++;
  AGB_CUDA(h, cudaGetLastError());
  return AGB_OK;
}
static int finish(agb_handle* h) {
  AGB_CUDA(h, cudaStreamSynchronize(h->stream));
  AGB_CUDA(h, cudaGetLastError());
  return AGB_OK;
}
static OpArgs op_args(int op) { OpArgs a; memset(&a, 0, sizeof a); a.op = op; a.player = -1; return a; }

int agb_rollout(agb_handle* h) {
  if (!h) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  AGB_TRY(launch_op(h, nullptr, op_args(OP_ROLLOUT)));
  return finish(h);
}

static int export_to_host(agb_handle* h, const double* internal, double* host_out, int mode) {
  const size_t B = h->batch, S = h->hd.S;
  AGB_TRY(ensure_stage(h, &h->stage2, &h->stage2_bytes, B * S * sizeof(double)));
  AGB_LAUNCH(agb_export_kernel, grid_for(B * S), 256, 0, h->stream, internal, h->stage2, h->batch, h->hd.p, h->hd.K, mode);
  h->launches++;
  return d2h(h, host_out, h->stage2, B * S);
}

int agb_residual(agb_handle* h, double reg_x, double reg_u, double alpha, double* res_out, double* norms_out) {
  if (!h) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = h->batch, S = h->hd.S;
  AGB_TRY(ensure_stage(h, &h->stage, &h->stage_bytes, (B * S + B * 5) * sizeof(double)));
  OpArgs a = op_args(OP_RESIDUAL);
  a.reg_x = reg_x; a.reg_u = reg_u; a.alpha = alpha; a.out0 = h->stage; a.out1 = h->stage + B * S;
  AGB_TRY(launch_op(h, nullptr, a));
  if (res_out) { if (h->hd.use_band) AGB_TRY(d2h(h, res_out, h->stage, B * S)); else AGB_TRY(export_to_host(h, h->stage, res_out, 0)); }
  AGB_TRY(d2h(h, norms_out, h->stage + B * S, B * 5));
  return finish(h);
}

int agb_residual_jacobian_dense(agb_handle* h, double reg_x, double reg_u, double* J_out) {
  if (!h || !J_out) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = h->batch, S = h->hd.S;
  const size_t bytes = B * S * S * sizeof(double);
  if (bytes > ((size_t)8 << 30)) return fail(h, AGB_EUNSUPPORTED, "dense Jacobian export is for small cases only (> 8 GiB requested)");
  AGB_TRY(ensure_stage(h, &h->stage, &h->stage_bytes, bytes));
  AGB_CUDA(h, cudaMemsetAsync(h->stage, 0, bytes, h->stream));
  OpArgs a = op_args(OP_JAC_DENSE);
  a.reg_x = reg_x; a.reg_u = reg_u; a.out0 = h->stage;
  AGB_TRY(launch_op(h, nullptr, a));
  AGB_TRY(d2h(h, J_out, h->stage, B * S * S));
  return finish(h);
}

int agb_kkt_solve(agb_handle* h, double reg_x, double reg_u, double* dtraj_out) {
  if (!h) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  OpArgs a = op_args(OP_KKT_SOLVE);
  a.reg_x = reg_x; a.reg_u = reg_u;
  if (h->hd.use_band) {                              // the band solver writes the step in the reference's column order itself
    const size_t B = h->batch, S = h->hd.S;
    AGB_TRY(ensure_stage(h, &h->stage2, &h->stage2_bytes, B * S * sizeof(double)));
    a.out0 = h->stage2;
    AGB_TRY(launch_op(h, nullptr, a));
    AGB_TRY(d2h(h, dtraj_out, h->stage2, B * S));
    return finish(h);
  }
  AGB_TRY(launch_op(h, nullptr, a));
  if (dtraj_out) AGB_TRY(export_to_host(h, h->D, dtraj_out, 1));
  return finish(h);
}

int agb_line_search(agb_handle* h, const agb_options* o, double reg_x, double reg_u, double* alpha_out, int* j_out) {
  if (!h || !o) return AGB_EINVAL;
  (void)reg_u;
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = h->batch;
  AGB_TRY(ensure_stage(h, &h->stage, &h->stage_bytes, B * (sizeof(double) + sizeof(int))));
  OpArgs a = op_args(OP_LINE_SEARCH);
  a.reg_x = reg_x; a.reg_u = reg_u; a.out0 = h->stage; a.iout = (int*)(h->stage + B);
  AGB_TRY(launch_op(h, o, a));
  AGB_TRY(d2h(h, alpha_out, h->stage, B));
  if (j_out) AGB_CUDA(h, cudaMemcpyAsync(j_out, a.iout, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  return finish(h);
}

int agb_update_traj(agb_handle* h, const double* alpha, double* delta_step_out) {
  if (!h || !alpha) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = h->batch;
  AGB_TRY(ensure_stage(h, &h->stage, &h->stage_bytes, 2 * B * sizeof(double)));
  AGB_TRY(h2d(h, h->stage, alpha, B));
  OpArgs a = op_args(OP_UPDATE);
  a.in0 = h->stage; a.out0 = h->stage + B;
  AGB_TRY(launch_op(h, nullptr, a));
  AGB_TRY(d2h(h, delta_step_out, h->stage + B, B));
  return finish(h);
}

static int simple_op(agb_handle* h, const agb_options* o, int op) {
  if (!h || !o) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  AGB_TRY(launch_op(h, o, op_args(op)));
  return finish(h);
}
int agb_dual_update(agb_handle* h, const agb_options* o) { return simple_op(h, o, OP_DUAL_UPDATE); }
int agb_penalty_update(agb_handle* h, const agb_options* o) { return simple_op(h, o, OP_PENALTY_UPDATE); }
int agb_reset_duals_penalties(agb_handle* h, const agb_options* o) { return simple_op(h, o, OP_RESET); }

int agb_evaluate_constraints(agb_handle* h, double* c_out) {
  if (!h || !c_out) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t cnt = (size_t)h->batch * h->hd.K * h->hd.nrow;
  if (!cnt) return AGB_OK;
  AGB_TRY(ensure_stage(h, &h->stage, &h->stage_bytes, cnt * sizeof(double)));
  OpArgs a = op_args(OP_EVAL_CON);
  a.out0 = h->stage;
  AGB_TRY(launch_op(h, nullptr, a));
  AGB_TRY(d2h(h, c_out, h->stage, cnt));
  return finish(h);
}

int agb_active_set(agb_handle* h, double tol, unsigned char* active_out) {
  if (!h || !active_out) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t cnt = (size_t)h->batch * h->hd.K * h->hd.nrow;
  if (!cnt) return AGB_OK;
  AGB_TRY(ensure_stage(h, &h->stage, &h->stage_bytes, cnt));
  OpArgs a = op_args(OP_ACTIVE_SET);
  a.tol = tol; a.bout = (unsigned char*)h->stage;
  AGB_TRY(launch_op(h, nullptr, a));
  AGB_CUDA(h, cudaMemcpyAsync(active_out, h->stage, cnt, cudaMemcpyDeviceToHost, h->stream));
  return finish(h);
}

// ---- active-set analysis (src/active_set/*.jl): bordered residual / Jacobian, active masks, null space -------------------------
struct DevBuf {                       // scratch of one call (an analysis entry point, not the solve path): freed on return
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};

static int as_check(agb_handle* h, int* Sv, int* Sh) {
  if (h->hd.use_band) return fail(h, AGB_EUNSUPPORTED, "the active-set analysis covers the planar models (collision avoidance between position pairs)");
  const int npo = h->hd.p * (h->hd.p - 1);
  *Sv = h->hd.S + h->hd.K * npo / 2; *Sh = h->hd.S + h->hd.K * npo;            // active_set_core.jl:81-82
  return AGB_OK;
}

int agb_active_set_sizes(agb_handle* h, int* Sv_out, int* Sh_out) {
  if (!h || !Sv_out || !Sh_out) return AGB_EINVAL;
  return as_check(h, Sv_out, Sh_out);
}

// device-side assembly shared by the entry points below: res [B][Sv], jac [B][Sv][Sh], vmask [B][Sv], hmask [B][Sh]
// (any of them may be null); masked = 0 marks every pair active
static int as_assemble(agb_handle* h, double tol, int masked, double* res, double* jac, unsigned char* vmask, unsigned char* hmask) {
  int Sv, Sh;
  AGB_TRY(as_check(h, &Sv, &Sh));
  const size_t B = h->batch, S = h->hd.S;
  DevBuf act;
  if (masked && (vmask || hmask)) {
    const size_t cnt = B * h->hd.K * h->hd.nrow;
    if (act.alloc(cnt) != cudaSuccess) return fail(h, AGB_ENOMEM, "cudaMalloc (active flags)");
    OpArgs a = op_args(OP_ACTIVE_SET);
    a.tol = tol; a.bout = (unsigned char*)act.p;
    AGB_TRY(launch_op(h, nullptr, a));
  }
  if (res) {
    AGB_CUDA(h, cudaMemsetAsync(res, 0, B * Sv * sizeof(double), h->stream));
    AGB_TRY(ensure_stage(h, &h->stage, &h->stage_bytes, (B * S + B * 5) * sizeof(double)));
    AGB_TRY(ensure_stage(h, &h->stage2, &h->stage2_bytes, B * S * sizeof(double)));
    OpArgs a = op_args(OP_RESIDUAL);
    a.out0 = h->stage; a.out1 = h->stage + B * S;
    AGB_TRY(launch_op(h, nullptr, a));
    AGB_LAUNCH(agb_export_kernel, grid_for(B * S), 256, 0, h->stream, h->stage, h->stage2, h->batch, h->hd.p, h->hd.K, 0);
    h->launches++;
    AGB_CUDA(h, cudaMemcpy2DAsync(res, Sv * sizeof(double), h->stage2, S * sizeof(double), S * sizeof(double), B, cudaMemcpyDeviceToDevice, h->stream));
  }
  if (jac) {
    AGB_CUDA(h, cudaMemsetAsync(jac, 0, B * Sv * Sh * sizeof(double), h->stream));
    AGB_TRY(ensure_stage(h, &h->stage, &h->stage_bytes, B * S * S * sizeof(double)));
    AGB_CUDA(h, cudaMemsetAsync(h->stage, 0, B * S * S * sizeof(double), h->stream));
    OpArgs a = op_args(OP_JAC_DENSE);                       // residual_jacobian!(prob, pdtraj), unregularised (:141-142)
    a.out0 = h->stage;
    AGB_TRY(launch_op(h, nullptr, a));
    for (size_t b = 0; b < B; b++)
      AGB_CUDA(h, cudaMemcpy2DAsync(jac + b * Sv * Sh, Sh * sizeof(double), h->stage + b * S * S, S * sizeof(double), S * sizeof(double), S,
                                    cudaMemcpyDeviceToDevice, h->stream));
  }
  if (vmask) AGB_CUDA(h, cudaMemsetAsync(vmask, 1, B * Sv, h->stream));
  if (hmask) AGB_CUDA(h, cudaMemsetAsync(hmask, 1, B * Sh, h->stream));
  AGB_LAUNCH(agb_as_border_kernel, grid_for(B * h->hd.K * h->hd.p * (h->hd.p - 1)), 256, 0, h->stream, h->dd, h->Z,
             (const unsigned char*)act.p, h->batch, jac, res, vmask, hmask);
  h->launches++;
  AGB_CUDA(h, cudaStreamSynchronize(h->stream));            // `act` is released on return
  AGB_CUDA(h, cudaGetLastError());
  return AGB_OK;
}

static int as_too_big(agb_handle* h, int Sv, int Sh) {
  if ((size_t)h->batch * Sv * Sh * sizeof(double) > ((size_t)8 << 30))
    return fail(h, AGB_EUNSUPPORTED, "dense active-set Jacobian is for small cases only (> 8 GiB requested)");
  return AGB_OK;
}

int agb_active_set_residual(agb_handle* h, double* res_out) {
  if (!h || !res_out) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  int Sv, Sh;
  AGB_TRY(as_check(h, &Sv, &Sh));
  DevBuf res;
  if (res.alloc((size_t)h->batch * Sv * sizeof(double)) != cudaSuccess) return fail(h, AGB_ENOMEM, "cudaMalloc (active-set residual)");
  AGB_TRY(as_assemble(h, 0.0, 0, (double*)res.p, nullptr, nullptr, nullptr));
  AGB_TRY(d2h(h, res_out, (double*)res.p, (size_t)h->batch * Sv));
  return finish(h);
}

int agb_active_set_jacobian_dense(agb_handle* h, double* jac_out) {
  if (!h || !jac_out) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  int Sv, Sh;
  AGB_TRY(as_check(h, &Sv, &Sh));
  AGB_TRY(as_too_big(h, Sv, Sh));
  DevBuf jac;
  if (jac.alloc((size_t)h->batch * Sv * Sh * sizeof(double)) != cudaSuccess) return fail(h, AGB_ENOMEM, "cudaMalloc (active-set Jacobian)");
  AGB_TRY(as_assemble(h, 0.0, 0, nullptr, (double*)jac.p, nullptr, nullptr));
  AGB_TRY(d2h(h, jac_out, (double*)jac.p, (size_t)h->batch * Sv * Sh));
  return finish(h);
}

int agb_active_set_masks(agb_handle* h, double tol, unsigned char* vmask_out, unsigned char* hmask_out) {
  if (!h || !vmask_out || !hmask_out) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  int Sv, Sh;
  AGB_TRY(as_check(h, &Sv, &Sh));
  const size_t B = h->batch;
  DevBuf mk;
  if (mk.alloc(B * (Sv + Sh)) != cudaSuccess) return fail(h, AGB_ENOMEM, "cudaMalloc (active-set masks)");
  unsigned char* vm = (unsigned char*)mk.p; unsigned char* hm = vm + B * Sv;
  AGB_TRY(as_assemble(h, tol, 1, nullptr, nullptr, vm, hm));
  AGB_CUDA(h, cudaMemcpyAsync(vmask_out, vm, B * Sv, cudaMemcpyDeviceToHost, h->stream));
  AGB_CUDA(h, cudaMemcpyAsync(hmask_out, hm, B * Sh, cudaMemcpyDeviceToHost, h->stream));
  return finish(h);
}

int agb_update_nullspace(agb_handle* h, double tol, double atol, int max_dim, double* null_out, int* dim_out) {
  if (!h || !dim_out || max_dim < 0 || (max_dim > 0 && !null_out)) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  int Sv, Sh;
  AGB_TRY(as_check(h, &Sv, &Sh));
  AGB_TRY(as_too_big(h, Sv, Sh));
  const size_t B = h->batch;
  DevBuf jac, work, iwork, mk, nul, dim;
  if (jac.alloc(B * Sv * Sh * sizeof(double)) != cudaSuccess || work.alloc(B * ((size_t)Sv * Sh + Sv) * sizeof(double)) != cudaSuccess ||
      iwork.alloc(B * (Sv + 2 * (size_t)Sh) * sizeof(int)) != cudaSuccess || mk.alloc(B * (Sv + Sh)) != cudaSuccess ||
      nul.alloc(B * (size_t)max_dim * Sh * sizeof(double)) != cudaSuccess || dim.alloc(B * sizeof(int)) != cudaSuccess)
    return fail(h, AGB_ENOMEM, "cudaMalloc (null-space workspace)");
  unsigned char* vm = (unsigned char*)mk.p; unsigned char* hm = vm + B * Sv;
  AGB_TRY(as_assemble(h, tol, 1, nullptr, (double*)jac.p, vm, hm));
  AGB_LAUNCH(agb_as_nullspace_kernel, (int)(B < 148 * 4 ? B : 148 * 4), kAsThreads, 0, h->stream, h->batch, Sv, Sh, (const double*)jac.p,
             (const unsigned char*)vm, (const unsigned char*)hm, (double*)work.p, (int*)iwork.p, atol, max_dim, (double*)nul.p, (int*)dim.p);
  h->launches++;
  AGB_TRY(d2h(h, null_out, (double*)nul.p, B * (size_t)max_dim * Sh));
  AGB_CUDA(h, cudaMemcpyAsync(dim_out, dim.p, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  return finish(h);
}

int agb_band_info(const agb_handle* h, int* window_bytes_out, int* slots_out) {
  if (!h) return AGB_EINVAL;
  if (window_bytes_out) *window_bytes_out = h->band_win;
  if (slots_out) *slots_out = h->band_slots;
  return AGB_OK;
}

int agb_debug_gain_solve(agb_handle* h, const double* aug, double* aug_out, int* ok_out) {
  if (!h || !aug || !aug_out) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = h->batch, cnt = B * h->hd.m * (h->hd.m + h->hd.n + 1);
  AGB_TRY(ensure_stage(h, &h->stage, &h->stage_bytes, (2 * cnt + B) * sizeof(double)));
  AGB_TRY(h2d(h, h->stage, aug, cnt));
  OpArgs a = op_args(OP_GAIN_SOLVE);
  a.in0 = h->stage; a.out0 = h->stage + cnt; a.iout = (int*)(h->stage + 2 * cnt);
  AGB_TRY(launch_op(h, nullptr, a));
  AGB_TRY(d2h(h, aug_out, h->stage + cnt, cnt));
  if (ok_out) AGB_CUDA(h, cudaMemcpyAsync(ok_out, a.iout, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  return finish(h);
}

// newton_solve! of instances [lo, hi) on stream st: the band solver for band-only schemas; otherwise the structured kernel,
// followed (fallback) by a band-solver launch that re-solves, from the same initial iterate, the instances it left
// AGB_SINGULAR.  `slot0 / nslots`: the band scratch slots this launch may use (concurrent chunks get disjoint slots).
static int launch_solve_range(agb_handle* h, const agb_options* o, cudaStream_t st, int lo, int hi, int slot0, int nslots, const MpcArgs* mp) {
  Buffers g = buffers_of(h);
  if (h->hd.use_band) {
    const int grid = (hi - lo) < nslots ? (hi - lo) : nslots;
    g.band = h->band + (size_t)slot0 * h->band_stride; g.band_slots = nslots;
    agb::launch_band_solve(h->dd, *o, g, lo, hi, -1, grid, st);
    h->launches++;
    AGB_CUDA(h, cudaGetLastError());
    return AGB_OK;
  }
  const size_t cs = (size_t)h->hd.K * h->hd.nrow;
  const bool fb = h->fallback && h->band != nullptr;
  if (fb && !o->dual_reset && cs)  {                        // warm start: keep what the solve began with for the fallback
    AGB_CUDA(h, cudaMemcpyAsync(h->conlam0 + lo * cs, h->conlam + lo * cs, (size_t)(hi - lo) * cs * sizeof(double), cudaMemcpyDeviceToDevice, st));
    AGB_CUDA(h, cudaMemcpyAsync(h->conmu0 + lo * cs, h->conmu + lo * cs, (size_t)(hi - lo) * cs * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  LaunchArgs L;
  L.model = h->hd.model; L.grid = hi - lo; L.smem = h->smem_bytes; L.stream = st; L.dd = h->dd; L.o = *o;
  memset(&L.io, 0, sizeof L.io);
  L.g = g; memset(&L.a, 0, sizeof L.a); L.batch = hi; L.inst0 = lo;
  memset(&L.mp, 0, sizeof L.mp);
  if (mp) L.mp = *mp;
  agb::launch_solve(h->hd.p, h->hd.big, L);
  h->launches++;
  AGB_CUDA(h, cudaGetLastError());
  if (fb && !mp) {
    g.band = h->band + (size_t)slot0 * h->band_stride; g.band_slots = nslots;
    if (!o->dual_reset && cs) { g.conlam0 = h->conlam0; g.conmu0 = h->conmu0; }
    const int grid = (hi - lo) < nslots ? (hi - lo) : nslots;
    agb::launch_band_solve(h->dd, *o, g, lo, hi, AGB_SINGULAR, grid, st);
    h->launches++;
    AGB_CUDA(h, cudaGetLastError());
  }
  return AGB_OK;
}
static int launch_solve(agb_handle* h, const agb_options* o, cudaStream_t st) {
  return launch_solve_range(h, o, st, 0, h->batch, 0, h->band_slots);
}

int agb_newton_solve_async(agb_handle* h, const agb_options* o, void* stream) {
  if (!h || !o) return AGB_EINVAL;
  if (o->ls_iter < 1 || o->outer_iter < 1 || o->inner_iter < 1) return fail(h, AGB_EINVAL, "outer_iter, inner_iter, ls_iter must be >= 1");
  AGB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (st != h->stream) {                 // the solve reads what earlier calls enqueued on the handle's stream …
    AGB_CUDA(h, cudaEventRecord(h->ev_async, h->stream));
    AGB_CUDA(h, cudaStreamWaitEvent(st, h->ev_async, 0));
  }
  AGB_TRY(launch_solve(h, o, st));
  if (st != h->stream) {                 // … and every later call on the handle sees its results
    AGB_CUDA(h, cudaEventRecord(h->ev_async, st));
    AGB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_async, 0));
  }
  return AGB_OK;
}

int agb_newton_solve_batch(agb_handle* h, const agb_options* o, double* Z_out, double* L_out, double* conlam_out,
                           double* conmu_out, double* stats_out, int* status_out) {
  if (!h || !o) return AGB_EINVAL;
  if (o->ls_iter < 1 || o->outer_iter < 1 || o->inner_iter < 1) return fail(h, AGB_EINVAL, "outer_iter, inner_iter, ls_iter must be >= 1");
  AGB_CUDA(h, cudaSetDevice(h->device));
  AGB_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  AGB_TRY(launch_solve(h, o, h->stream));
  AGB_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  h->timed = true;
  const size_t B = h->batch, zs = (size_t)h->hd.N * (h->hd.n + h->hd.m), ls = (size_t)h->hd.p * h->hd.K * h->hd.n, cs = (size_t)h->hd.K * h->hd.nrow;
  AGB_TRY(d2h(h, Z_out, h->Z, B * zs)); AGB_TRY(d2h(h, L_out, h->L, B * ls));
  AGB_TRY(d2h(h, conlam_out, h->conlam, B * cs)); AGB_TRY(d2h(h, conmu_out, h->conmu, B * cs));
  AGB_TRY(d2h(h, stats_out, h->stats, B * AGB_NSTATS));
  if (status_out) AGB_CUDA(h, cudaMemcpyAsync(status_out, h->status, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  return finish(h);
}

int agb_ibr_newton_solve_batch(agb_handle* h, const agb_options* o, const agb_ibr_options* io, double* Z_out, double* L_out,
                               double* conlam_out, double* conmu_out, double* stats_out, int* status_out) {
  if (!h || !o || !io) return AGB_EINVAL;
  if (h->hd.use_band) return fail(h, AGB_EUNSUPPORTED, "iterative best response is not available for schemas solved by the band solver");
  if (o->ls_iter < 1 || o->outer_iter < 1 || o->inner_iter < 1 || io->ibr_iter < 1)
    return fail(h, AGB_EINVAL, "outer_iter, inner_iter, ls_iter, ibr_iter must be >= 1");
  unsigned seen = 0;
  for (int i = 0; i < h->hd.p; i++) {
    if (io->ordering[i] < 0 || io->ordering[i] >= h->hd.p || (seen >> io->ordering[i] & 1u)) return fail(h, AGB_EINVAL, "ordering must be a permutation of the players");
    seen |= 1u << io->ordering[i];
  }
  AGB_CUDA(h, cudaSetDevice(h->device));
  LaunchArgs L;
  L.model = h->hd.model; L.grid = h->batch; L.smem = h->smem_bytes; L.stream = h->stream; L.dd = h->dd; L.o = *o; L.io = *io;
  L.g = buffers_of(h); memset(&L.a, 0, sizeof L.a); L.batch = h->batch;
  AGB_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  agb::launch_ibr(h->hd.p, h->hd.big, L);
  h->launches++;
  AGB_CUDA(h, cudaGetLastError());
  AGB_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  h->timed = true;
  const size_t B = h->batch, zs = (size_t)h->hd.N * (h->hd.n + h->hd.m), ls = (size_t)h->hd.p * h->hd.K * h->hd.n, cs = (size_t)h->hd.K * h->hd.nrow;
  AGB_TRY(d2h(h, Z_out, h->Z, B * zs)); AGB_TRY(d2h(h, L_out, h->L, B * ls));
  AGB_TRY(d2h(h, conlam_out, h->conlam, B * cs)); AGB_TRY(d2h(h, conmu_out, h->conmu, B * cs));
  AGB_TRY(d2h(h, stats_out, h->stats, B * AGB_NSTATS));
  if (status_out) AGB_CUDA(h, cudaMemcpyAsync(status_out, h->status, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  return finish(h);
}

int agb_ibr_residual(agb_handle* h, int player, double reg_x, double reg_u, double alpha, double* res_out, double* norms_out) {
  if (!h || player < 0 || player >= h->hd.p) return fail(h, AGB_EINVAL, "bad player index");
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = h->batch, S = h->hd.S;
  AGB_TRY(ensure_stage(h, &h->stage, &h->stage_bytes, (B * S + B * 5) * sizeof(double)));
  OpArgs a = op_args(OP_RESIDUAL);
  a.player = player; a.reg_x = reg_x; a.reg_u = reg_u; a.alpha = alpha; a.out0 = h->stage; a.out1 = h->stage + B * S;
  AGB_TRY(launch_op(h, nullptr, a));
  if (res_out) AGB_TRY(export_to_host(h, h->stage, res_out, 0));
  AGB_TRY(d2h(h, norms_out, h->stage + B * S, B * 5));
  return finish(h);
}

int agb_ibr_kkt_solve(agb_handle* h, int player, double reg_x, double reg_u, double* dtraj_out) {
  if (!h || player < 0 || player >= h->hd.p) return fail(h, AGB_EINVAL, "bad player index");
  AGB_CUDA(h, cudaSetDevice(h->device));
  OpArgs a = op_args(OP_KKT_SOLVE);
  a.player = player; a.reg_x = reg_x; a.reg_u = reg_u;
  AGB_TRY(launch_op(h, nullptr, a));
  if (dtraj_out) AGB_TRY(export_to_host(h, h->D, dtraj_out, 1));
  return finish(h);
}

// the chunk pipeline of agb_solve_from_host: chunk c = [H2D inputs -> solve -> D2H results] on its own stream
static int enqueue_host_chunks(agb_handle* h, const agb_options* o, int chunks, const double* x0, const double* Z0, const double* L0,
                               double* Z_out, double* L_out, double* conlam_out, double* conmu_out, double* stats_out, int* status_out) {
  const int B = h->batch, n = h->hd.n;
  const size_t zs = (size_t)h->hd.N * (h->hd.n + h->hd.m), ls = (size_t)h->hd.p * h->hd.K * h->hd.n, cs = (size_t)h->hd.K * h->hd.nrow;
  for (int c = 0; c < chunks; c++) {
    cudaStream_t st = h->chunk_stream[c];
    const int lo = (int)((long long)B * c / chunks), hi = (int)((long long)B * (c + 1) / chunks), cnt = hi - lo;
    if (cnt <= 0) continue;
    AGB_CUDA(h, cudaMemcpyAsync(h->x0 + (size_t)lo * n, x0 + (size_t)lo * n, (size_t)cnt * n * sizeof(double), cudaMemcpyHostToDevice, st));
    AGB_CUDA(h, cudaMemcpyAsync(h->Z0 + lo * zs, Z0 + lo * zs, cnt * zs * sizeof(double), cudaMemcpyHostToDevice, st));
    AGB_CUDA(h, cudaMemcpyAsync(h->L0 + lo * ls, L0 + lo * ls, cnt * ls * sizeof(double), cudaMemcpyHostToDevice, st));
    {
      const int per = h->band_slots / chunks > 0 ? h->band_slots / chunks : 1;        // disjoint band scratch slots per chunk
      AGB_TRY(launch_solve_range(h, o, st, lo, hi, (c * per) % (h->band_slots > 0 ? h->band_slots : 1), per));
    }
    if (Z_out) AGB_CUDA(h, cudaMemcpyAsync(Z_out + lo * zs, h->Z + lo * zs, cnt * zs * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (L_out) AGB_CUDA(h, cudaMemcpyAsync(L_out + lo * ls, h->L + lo * ls, cnt * ls * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (conlam_out && cs) AGB_CUDA(h, cudaMemcpyAsync(conlam_out + lo * cs, h->conlam + lo * cs, cnt * cs * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (conmu_out && cs) AGB_CUDA(h, cudaMemcpyAsync(conmu_out + lo * cs, h->conmu + lo * cs, cnt * cs * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (stats_out) AGB_CUDA(h, cudaMemcpyAsync(stats_out + (size_t)lo * AGB_NSTATS, h->stats + (size_t)lo * AGB_NSTATS, (size_t)cnt * AGB_NSTATS * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (status_out) AGB_CUDA(h, cudaMemcpyAsync(status_out + lo, h->status + lo, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  return AGB_OK;
}

#ifndef AGB_EMULATE
static bool host_pinned(const void* p) {                 // page-locked (or managed) host memory: async copies of it can be captured
  if (!p) return true;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
#endif

int agb_solve_from_host(agb_handle* h, const agb_options* o, const double* x0, const double* Z0, const double* L0,
                        double* Z_out, double* L_out, double* conlam_out, double* conmu_out, double* stats_out, int* status_out) {
  if (!h || !o || !x0 || !Z0 || !L0) return AGB_EINVAL;
  if (o->ls_iter < 1 || o->outer_iter < 1 || o->inner_iter < 1) return fail(h, AGB_EINVAL, "outer_iter, inner_iter, ls_iter must be >= 1");
  AGB_CUDA(h, cudaSetDevice(h->device));
  AGB_CUDA(h, cudaStreamSynchronize(h->stream));
  const int B = h->batch;
  bool graph_ok = false;
#ifndef AGB_EMULATE
  {
    const char* e = getenv("AGB_HOST_GRAPH");          // "0": always enqueue call by call
    graph_ok = !(e && e[0] == '0') && !h->hd.use_band && host_pinned(x0) && host_pinned(Z0) && host_pinned(L0) && host_pinned(Z_out) &&
               host_pinned(L_out) && host_pinned(conlam_out) && host_pinned(conmu_out) && host_pinned(stats_out) && host_pinned(status_out);
  }
#endif
  // each chunk: H2D -> solve -> D2H on its own stream.  Measured on config B, batch 1024 (converged instances/s end to end): call by
  // call 1 / 2 / 4 / 6 / 8 / 12 chunks = 296 / 332 / 339 / 346 / 345 / 325 k; replayed as a graph 8 / 12 / 16 / 24 / 32 chunks =
  // 348 / 346 / 345 / 342 / 334 k — finer chunks lose on the device side (small kernels and copies), not on the host's API time
  int chunks = B >= 1024 ? 8 : (B >= 512 ? 4 : 1);
  if (h->hd.use_band) chunks = 1;                     // band-only schemas: a solve takes ~100 ms, the copies microseconds, and splitting
                                                      // the scratch slots over chunks only adds waves (measured: 4 chunks = 3.5 x slower)
  if (const char* e = getenv("AGB_HOST_CHUNKS")) {    // tuning hook
    const int v = atoi(e);
    if (v >= 1 && v <= agb_handle::kMaxChunks) chunks = v < B ? v : B;
  }
  if (h->band_slots > 0 && chunks > h->band_slots) chunks = h->band_slots;           // every chunk needs its own band scratch slot
  for (int c = 0; c < chunks; c++)
    if (!h->chunk_stream[c]) AGB_CUDA(h, cudaStreamCreateWithFlags(&h->chunk_stream[c], cudaStreamNonBlocking));
#ifndef AGB_EMULATE
  if (graph_ok) {
    agb_handle::HostKey key;
    memset(&key, 0, sizeof key);
    const void* ptrs[9] = {x0, Z0, L0, Z_out, L_out, conlam_out, conmu_out, stats_out, status_out};
    memcpy(key.ptr, ptrs, sizeof ptrs); key.o = *o; key.chunks = chunks; key.epoch = h->fh_epoch;
    if (h->fh_valid && memcmp(&key, &h->fh_key, sizeof key) == 0) {                    // the captured pipeline, one launch
      AGB_CUDA(h, cudaGraphLaunch(h->fh_exec, h->stream));
      h->launches += h->fh_launches;
      AGB_CUDA(h, cudaStreamSynchronize(h->stream));
      return AGB_OK;
    }
    if (h->fh_seen_valid && memcmp(&key, &h->fh_seen, sizeof key) == 0) {              // second call with these buffers: capture it
      if (!h->fh_fork) AGB_CUDA(h, cudaEventCreateWithFlags(&h->fh_fork, cudaEventDisableTiming));
      for (int c = 0; c < chunks; c++) if (!h->fh_join[c]) AGB_CUDA(h, cudaEventCreateWithFlags(&h->fh_join[c], cudaEventDisableTiming));
      if (h->fh_exec) { cudaGraphExecDestroy(h->fh_exec); h->fh_exec = nullptr; }
      h->fh_valid = false;
      cudaGraph_t graph = nullptr;
      const int launches0 = h->launches;
      AGB_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
      int rc = AGB_OK;
      cudaError_t ce = cudaEventRecord(h->fh_fork, h->stream);
      for (int c = 0; c < chunks && ce == cudaSuccess; c++) ce = cudaStreamWaitEvent(h->chunk_stream[c], h->fh_fork, 0);
      if (ce == cudaSuccess) rc = enqueue_host_chunks(h, o, chunks, x0, Z0, L0, Z_out, L_out, conlam_out, conmu_out, stats_out, status_out);
      for (int c = 0; c < chunks && ce == cudaSuccess; c++) {
        ce = cudaEventRecord(h->fh_join[c], h->chunk_stream[c]);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(h->stream, h->fh_join[c], 0);
      }
      const cudaError_t ee = cudaStreamEndCapture(h->stream, &graph);                  // always ends the capture, also after an error
      h->fh_launches = h->launches - launches0;
      h->launches = launches0;
      if (rc == AGB_OK && ce == cudaSuccess && ee == cudaSuccess && graph &&
          cudaGraphInstantiate(&h->fh_exec, graph, 0) == cudaSuccess) {
        cudaGraphDestroy(graph);
        h->fh_key = key; h->fh_valid = true;
        AGB_CUDA(h, cudaGraphLaunch(h->fh_exec, h->stream));
        h->launches += h->fh_launches;
        AGB_CUDA(h, cudaStreamSynchronize(h->stream));
        return AGB_OK;
      }
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();                                                              // capture failed: fall through to the call-by-call path
      h->fh_seen_valid = false;
    } else {
      h->fh_seen = key; h->fh_seen_valid = true;
    }
  }
#endif
  AGB_TRY(enqueue_host_chunks(h, o, chunks, x0, Z0, L0, Z_out, L_out, conlam_out, conmu_out, stats_out, status_out));
  for (int c = 0; c < chunks; c++) if (h->chunk_stream[c]) AGB_CUDA(h, cudaStreamSynchronize(h->chunk_stream[c]));
  AGB_CUDA(h, cudaGetLastError());
  return AGB_OK;
}

int agb_get_device_view(agb_handle* h, agb_device_view* out) {
  if (!h || !out) return AGB_EINVAL;
  out->Z_dev = h->Z; out->L_dev = h->L; out->conlam_dev = h->conlam; out->conmu_dev = h->conmu; out->stats_dev = h->stats;
  out->status_dev = h->status; out->x0_dev = h->x0; out->Z0_dev = h->Z0; out->L0_dev = h->L0;
  out->results_dev = h->results; out->results_bytes = (unsigned long long)h->results_doubles * sizeof(double);
  return AGB_OK;
}

int agb_set_history(agb_handle* h, int max_records) {
  if (!h || max_records < 0) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
#ifndef AGB_EMULATE
  h->fh_epoch++;                                       // the history buffers are kernel arguments of a captured pipeline
#endif
  AGB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->hist) { cudaFree(h->hist); h->hist = nullptr; }
  if (h->hist_count) { cudaFree(h->hist_count); h->hist_count = nullptr; }
  h->hist_max = 0;
  if (max_records == 0) return AGB_OK;
  const size_t B = h->batch;
  AGB_CUDA(h, cudaMalloc((void**)&h->hist, B * max_records * AGB_NHIST * sizeof(double)));
  AGB_CUDA(h, cudaMalloc((void**)&h->hist_count, B * sizeof(int)));
  AGB_CUDA(h, cudaMemsetAsync(h->hist, 0, B * max_records * AGB_NHIST * sizeof(double), h->stream));
  AGB_CUDA(h, cudaMemsetAsync(h->hist_count, 0, B * sizeof(int), h->stream));
  AGB_CUDA(h, cudaStreamSynchronize(h->stream));
  h->hist_max = max_records;
  return AGB_OK;
}

int agb_get_history(agb_handle* h, double* hist_out, int* count_out) {
  if (!h || !hist_out || !count_out) return AGB_EINVAL;
  if (!h->hist) return fail(h, AGB_EINVAL, "agb_get_history: no history (call agb_set_history first)");
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = h->batch;
  AGB_CUDA(h, cudaMemcpyAsync(hist_out, h->hist, B * h->hist_max * AGB_NHIST * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  AGB_CUDA(h, cudaMemcpyAsync(count_out, h->hist_count, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  AGB_CUDA(h, cudaStreamSynchronize(h->stream));
  return AGB_OK;
}

long long agb_launch_count(const agb_handle* h) { return h ? h->launches : 0; }

float agb_last_solve_ms(agb_handle* h) {
  if (!h || !h->timed) return -1.0f;
  cudaSetDevice(h->device);
  if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0f;
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0f;
  return ms;
}

// ---- ABI layout check (bindings mirror the structs by hand) ------------------------------------------------------
static void abi_words(int* w) {
  int k = 0;
  w[k++] = (int)sizeof(agb_problem_desc); w[k++] = (int)offsetof(agb_problem_desc, dt); w[k++] = (int)offsetof(agb_problem_desc, Q);
  w[k++] = (int)offsetof(agb_problem_desc, col_radius); w[k++] = (int)offsetof(agb_problem_desc, has_state_bound);
  w[k++] = (int)offsetof(agb_problem_desc, walls); w[k++] = (int)offsetof(agb_problem_desc, circles); w[k++] = (int)offsetof(agb_problem_desc, x_max_con);
  w[k++] = (int)offsetof(agb_problem_desc, quad_mass); w[k++] = (int)offsetof(agb_problem_desc, solver);
  w[k++] = (int)sizeof(agb_options); w[k++] = (int)offsetof(agb_options, alphax_dual); w[k++] = (int)offsetof(agb_options, eps_dyn);
  w[k++] = (int)offsetof(agb_options, dual_reset);
  w[k++] = (int)sizeof(agb_ibr_options); w[k++] = (int)offsetof(agb_ibr_options, delta_min);
  w[k++] = (int)sizeof(agb_sizes); w[k++] = (int)sizeof(agb_device_view);
  w[k++] = AGB_MAX_P; w[k++] = AGB_MAX_N; w[k++] = AGB_MAX_M; w[k++] = AGB_MAX_WALLS; w[k++] = AGB_MAX_CIRCLES;
  w[k++] = AGB_NSTATS; w[k++] = AGB_NHIST; w[k++] = AGB_IPC_BYTES;
  static_assert(AGB_ABI_WORDS == 26, "update abi_words");
}
static const char* const kAbiNames[AGB_ABI_WORDS] = {
  "sizeof(agb_problem_desc)", "offsetof(agb_problem_desc, dt)", "offsetof(agb_problem_desc, Q)", "offsetof(agb_problem_desc, col_radius)",
  "offsetof(agb_problem_desc, has_state_bound)", "offsetof(agb_problem_desc, walls)", "offsetof(agb_problem_desc, circles)",
  "offsetof(agb_problem_desc, x_max_con)", "offsetof(agb_problem_desc, quad_mass)", "offsetof(agb_problem_desc, solver)", "sizeof(agb_options)", "offsetof(agb_options, alphax_dual)", "offsetof(agb_options, eps_dyn)",
  "offsetof(agb_options, dual_reset)", "sizeof(agb_ibr_options)", "offsetof(agb_ibr_options, delta_min)", "sizeof(agb_sizes)",
  "sizeof(agb_device_view)", "AGB_MAX_P", "AGB_MAX_N", "AGB_MAX_M", "AGB_MAX_WALLS", "AGB_MAX_CIRCLES", "AGB_NSTATS", "AGB_NHIST", "AGB_IPC_BYTES"};

int agb_abi_layout(int* layout_out, int count) {
  if (!layout_out || count < AGB_ABI_WORDS) return fail(nullptr, AGB_EINVAL, "agb_abi_layout: need room for AGB_ABI_WORDS ints");
  abi_words(layout_out);
  return AGB_OK;
}
int agb_abi_check(const int* layout, int count) {
  if (!layout || count != AGB_ABI_WORDS) return fail(nullptr, AGB_EINVAL, "agb_abi_check: expected AGB_ABI_WORDS = 26 entries (binding built against another header?)");
  int mine[AGB_ABI_WORDS];
  abi_words(mine);
  for (int k = 0; k < AGB_ABI_WORDS; k++) {
    if (layout[k] != mine[k]) {
      char buf[200];
      snprintf(buf, sizeof buf, "ABI mismatch: %s is %d in the library, %d in the binding", kAbiNames[k], mine[k], layout[k]);
      return fail(nullptr, AGB_EINVAL, buf);
    }
  }
  return AGB_OK;
}

// ---- per-knot violation vectors ----------------------------------------------------------------------------------
int agb_violations(agb_handle* h, double* dyn_out, double* con_out, double* sta_out, double* opt_out) {
  if (!h) return AGB_EINVAL;
  AGB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = h->batch, N = h->hd.N, K = h->hd.K, per = 4 * N - 2;
  AGB_TRY(ensure_stage(h, &h->stage, &h->stage_bytes, B * per * sizeof(double)));
  OpArgs a = op_args(OP_VIOLATIONS);
  a.out0 = h->stage;
  AGB_TRY(launch_op(h, nullptr, a));
  const size_t pitch = per * sizeof(double);
  if (dyn_out) AGB_CUDA(h, cudaMemcpy2DAsync(dyn_out, K * sizeof(double), h->stage, pitch, K * sizeof(double), B, cudaMemcpyDeviceToHost, h->stream));
  if (con_out) AGB_CUDA(h, cudaMemcpy2DAsync(con_out, K * sizeof(double), h->stage + K, pitch, K * sizeof(double), B, cudaMemcpyDeviceToHost, h->stream));
  if (sta_out) AGB_CUDA(h, cudaMemcpy2DAsync(sta_out, N * sizeof(double), h->stage + 2 * K, pitch, N * sizeof(double), B, cudaMemcpyDeviceToHost, h->stream));
  if (opt_out) AGB_CUDA(h, cudaMemcpy2DAsync(opt_out, N * sizeof(double), h->stage + 2 * K + N, pitch, N * sizeof(double), B, cudaMemcpyDeviceToHost, h->stream));
  return finish(h);
}

// ---- multi-GPU: result slabs of all ranks on every rank, pushed over peer memory by the copy engines ----------------
static size_t slab_bytes_for(const agb_handle* h, size_t B) {
  const size_t zs = B * h->hd.N * (h->hd.n + h->hd.m), ls = B * h->hd.p * h->hd.K * h->hd.n, ss = B * AGB_NSTATS, is = (B + 1) / 2;
  return (zs + ls + ss + is) * sizeof(double);
}

int agb_peer_init(agb_handle* h, int nranks, int rank, const int* batches) {
  if (!h || !batches || nranks < 1 || nranks > AGB_MAX_RANKS || rank < 0 || rank >= nranks) return fail(h, AGB_EINVAL, "agb_peer_init: bad rank layout");
  if (batches[rank] != h->batch) return fail(h, AGB_EINVAL, "agb_peer_init: batches[rank] must equal this handle's batch");
  if (h->gather) return fail(h, AGB_EINVAL, "agb_peer_init: already initialised");
  AGB_CUDA(h, cudaSetDevice(h->device));
  h->slab_off[0] = 0;
  for (int r = 0; r < nranks; r++) {
    if (batches[r] < 1) return fail(h, AGB_EINVAL, "agb_peer_init: every rank needs at least one instance");
    h->slab_off[r + 1] = h->slab_off[r] + slab_bytes_for(h, (size_t)batches[r]);
  }
  if (h->slab_off[rank + 1] - h->slab_off[rank] != h->results_doubles * sizeof(double)) return fail(h, AGB_EINVAL, "agb_peer_init: slab size mismatch");
  AGB_CUDA(h, cudaMalloc((void**)&h->gather, h->slab_off[nranks]));
  AGB_CUDA(h, cudaMemsetAsync(h->gather, 0, h->slab_off[nranks], h->stream));
  AGB_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  AGB_CUDA(h, cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
  AGB_CUDA(h, cudaStreamSynchronize(h->stream));
  h->nranks = nranks; h->rank = rank;
  for (int r = 0; r < nranks; r++) { h->peer_gather[r] = nullptr; h->peer_device[r] = -1; h->peer_ipc_opened[r] = false; }
  h->peer_gather[rank] = h->gather; h->peer_device[rank] = h->device;
  return AGB_OK;
}

int agb_peer_export(agb_handle* h, unsigned char* ipc_out) {
  if (!h || !ipc_out || !h->gather) return fail(h, AGB_EINVAL, "agb_peer_export: call agb_peer_init first");
  static_assert(sizeof(cudaIpcMemHandle_t) == AGB_IPC_BYTES, "AGB_IPC_BYTES");
  AGB_CUDA(h, cudaSetDevice(h->device));
  cudaIpcMemHandle_t mh;
  AGB_CUDA(h, cudaIpcGetMemHandle(&mh, h->gather));
  memcpy(ipc_out, &mh, AGB_IPC_BYTES);
  return AGB_OK;
}

int agb_peer_connect(agb_handle* h, const unsigned char* ipc_all) {
  if (!h || !ipc_all || !h->gather) return fail(h, AGB_EINVAL, "agb_peer_connect: call agb_peer_init first");
  AGB_CUDA(h, cudaSetDevice(h->device));
  for (int r = 0; r < h->nranks; r++) {
    if (r == h->rank || h->peer_gather[r]) continue;
    cudaIpcMemHandle_t mh;
    memcpy(&mh, ipc_all + (size_t)r * AGB_IPC_BYTES, AGB_IPC_BYTES);
    void* p = nullptr;
    AGB_CUDA(h, cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
    h->peer_gather[r] = (unsigned char*)p; h->peer_device[r] = -1; h->peer_ipc_opened[r] = true;
  }
  return AGB_OK;
}

int agb_peer_connect_local(agb_handle* const* handles, int nranks) {
  if (!handles || nranks < 1 || nranks > AGB_MAX_RANKS) return fail(nullptr, AGB_EINVAL, "agb_peer_connect_local: bad arguments");
  for (int r = 0; r < nranks; r++)
    if (!handles[r] || handles[r]->nranks != nranks || handles[r]->rank != r || !handles[r]->gather)
      return fail(nullptr, AGB_EINVAL, "agb_peer_connect_local: handles[r] must be rank r of an nranks layout (agb_peer_init)");
  for (int r = 0; r < nranks; r++) {
    agb_handle* h = handles[r];
    AGB_CUDA(h, cudaSetDevice(h->device));
    for (int q = 0; q < nranks; q++) {
      if (q == r) continue;
      if (handles[q]->device != h->device) {
        int can = 0;
        AGB_CUDA(h, cudaDeviceCanAccessPeer(&can, h->device, handles[q]->device));
        if (can) {
          cudaError_t e = cudaDeviceEnablePeerAccess(handles[q]->device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(h, AGB_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
          (void)cudaGetLastError();
        }
      }
      h->peer_gather[q] = handles[q]->gather; h->peer_device[q] = handles[q]->device;
    }
  }
  return AGB_OK;
}

int agb_allgather(agb_handle* h, void* stream) {
  if (!h || !h->gather) return fail(h, AGB_EINVAL, "agb_allgather: call agb_peer_init and connect the peers first");
  for (int r = 0; r < h->nranks; r++) if (!h->peer_gather[r]) return fail(h, AGB_EINVAL, "agb_allgather: peers not connected");
  AGB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t after = stream ? (cudaStream_t)stream : h->stream;
  AGB_CUDA(h, cudaEventRecord(h->ev_ready, after));
  AGB_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_ready, 0));
  const size_t bytes = h->results_doubles * sizeof(double), off = h->slab_off[h->rank];
  for (int q = 0; q < h->nranks; q++) {
    const int r = (h->rank + q) % h->nranks;             // start with the own copy, then ring order: spreads the link load
    unsigned char* dst = h->peer_gather[r] + off;
    if (h->peer_device[r] >= 0 && h->peer_device[r] != h->device)
      AGB_CUDA(h, cudaMemcpyPeerAsync(dst, h->peer_device[r], h->results, h->device, bytes, h->copy_stream));
    else
      AGB_CUDA(h, cudaMemcpyAsync(dst, h->results, bytes, cudaMemcpyDeviceToDevice, h->copy_stream));
  }
  return AGB_OK;
}

int agb_allgather_wait(agb_handle* h) {
  if (!h || !h->gather) return fail(h, AGB_EINVAL, "agb_allgather_wait: call agb_peer_init first");
  AGB_CUDA(h, cudaSetDevice(h->device));
  AGB_CUDA(h, cudaStreamSynchronize(h->copy_stream));
  return AGB_OK;
}

int agb_gathered_view(agb_handle* h, void** gathered_dev, unsigned long long* offsets_out) {
  if (!h || !h->gather) return fail(h, AGB_EINVAL, "agb_gathered_view: call agb_peer_init first");
  if (gathered_dev) *gathered_dev = h->gather;
  if (offsets_out) for (int r = 0; r <= h->nranks; r++) offsets_out[r] = (unsigned long long)h->slab_off[r];
  return AGB_OK;
}

int agb_unpack_gathered(agb_handle* h, int src_rank, double* Z, double* L, double* stats, int* status) {
  if (!h || !h->gather || src_rank < 0 || src_rank >= h->nranks) return fail(h, AGB_EINVAL, "agb_unpack_gathered: bad rank");
  AGB_CUDA(h, cudaSetDevice(h->device));
  // batch of the source rank from its slab size: slab(B) = 8·(B·(zs+ls+NSTATS) + ceil(B/2))
  const size_t zs1 = (size_t)h->hd.N * (h->hd.n + h->hd.m), ls1 = (size_t)h->hd.p * h->hd.K * h->hd.n;
  const size_t bytes = h->slab_off[src_rank + 1] - h->slab_off[src_rank];
  size_t B = bytes / (8 * (zs1 + ls1 + AGB_NSTATS));
  while (B > 0 && slab_bytes_for(h, B) > bytes) B--;
  if (B == 0 || slab_bytes_for(h, B) != bytes) return fail(h, AGB_EINVAL, "agb_unpack_gathered: inconsistent slab size");
  const double* base = (const double*)(h->gather + h->slab_off[src_rank]);
  AGB_TRY(d2h(h, Z, base, B * zs1)); AGB_TRY(d2h(h, L, base + B * zs1, B * ls1));
  AGB_TRY(d2h(h, stats, base + B * (zs1 + ls1), B * AGB_NSTATS));
  if (status) AGB_CUDA(h, cudaMemcpyAsync(status, base + B * (zs1 + ls1 + AGB_NSTATS), B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  return finish(h);
}

int agb_create_sharded(const agb_problem_desc* desc, int batch, int ndev, const int* devices, agb_handle** handles_out, int* ndev_out) {
  if (!desc || !handles_out || !ndev_out) return fail(nullptr, AGB_EINVAL, "null argument");
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) return fail(nullptr, AGB_ECUDA, "no CUDA device: libalgames_b200 has no CPU fallback");
  if (ndev <= 0) { ndev = have; devices = nullptr; }
  if (ndev > AGB_MAX_RANKS) return fail(nullptr, AGB_EINVAL, "too many devices");
  if (batch < ndev) return fail(nullptr, AGB_EINVAL, "agb_create_sharded: batch must be >= the number of devices");
  int batches[AGB_MAX_RANKS];
  for (int r = 0; r < ndev; r++) batches[r] = batch / ndev + (r < batch % ndev ? 1 : 0);      // ranks < batch % ndev get one extra
  int rc = AGB_OK, made = 0;
  for (int r = 0; r < ndev && rc == AGB_OK; r++) {
    rc = agb_create(desc, batches[r], devices ? devices[r] : r, &handles_out[r]);
    if (rc == AGB_OK) { made++; rc = agb_peer_init(handles_out[r], ndev, r, batches); if (rc) g_create_err = handles_out[r]->err; }
  }
  if (rc == AGB_OK) rc = agb_peer_connect_local(handles_out, ndev);
  if (rc != AGB_OK) { for (int r = 0; r < made; r++) { agb_destroy(handles_out[r]); handles_out[r] = nullptr; } return rc; }
  *ndev_out = ndev;
  return AGB_OK;
}

// ---- measured FP64 vector peak -------------------------------------------------------------------------------------
int agb_measure_fp64_peak(int device, int iters, double* tflops_out, float* ms_out) {
  if (!tflops_out || iters < 8) return fail(nullptr, AGB_EINVAL, "agb_measure_fp64_peak: bad arguments");
#ifdef AGB_EMULATE
  return fail(nullptr, AGB_EUNSUPPORTED, "agb_measure_fp64_peak: needs a CUDA device");
#endif
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(nullptr, AGB_ECUDA, "no such CUDA device");
  AGB_CUDA(nullptr, cudaSetDevice(device));
  int sms = 0;
  AGB_CUDA(nullptr, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  double* out = nullptr;
  AGB_CUDA(nullptr, cudaMalloc((void**)&out, sizeof(double)));
  cudaEvent_t e0, e1;
  AGB_CUDA(nullptr, cudaEventCreate(&e0)); AGB_CUDA(nullptr, cudaEventCreate(&e1));
  iters = (iters + 7) & ~7;
  const int grid = sms * 2, block = 1024;
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {                      // first repetition warms up; best of the rest
    cudaEventRecord(e0, 0);
    AGB_LAUNCH(agb_fp64_peak_kernel, grid, block, 0, 0, out, iters, 0.999999, 1e-7);
    cudaEventRecord(e1, 0);
    AGB_CUDA(nullptr, cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  AGB_CUDA(nullptr, cudaGetLastError());
  *tflops_out = 2.0 * 8.0 * (double)iters * (double)block * (double)grid / ((double)best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  return AGB_OK;
}

}  // extern "C"
