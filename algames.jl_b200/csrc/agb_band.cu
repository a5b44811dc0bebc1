// agb_band.cu — kernels and launchers of the band solver (agb_band.cuh): the general newton_solve! for every schema, and
// the fallback for instances the structured kernels return as AGB_SINGULAR.
#include <cstdio>
#include "agb_band.cuh"
#include "agb_kernels.cuh"

namespace agb {

using band::Ctx;
using band::Norms;
constexpr int kBandThreads = 128;     // band_solve_window reproduces the 128 partial sums of the global path's back substitution

// time-major band numbering → the reference's vertical (row) / horizontal (column) order (core/newton_core.jl:40-89)
__device__ __forceinline__ int ref_row(const Ctx& C, int r) {
  const int s = r / C.b, e = r % C.b;
  if (e < C.n) return C.p * C.K * (C.n + C.mi) + s * C.n + e;
  if (e < C.n + C.m) { const int idx = e - C.n, i = idx % C.p, j = idx / C.p; return (i * C.K + s) * (C.n + C.mi) + C.n + j; }
  const int t = e - C.n - C.m, i = t / C.n, a = t % C.n;
  return (i * C.K + s) * (C.n + C.mi) + a;
}
__device__ __forceinline__ int ref_col(const Ctx& C, int c) {
  const int s = c / C.b, e = c % C.b;
  if (e < C.p * C.n) return s * C.b + C.n + C.m + e;
  if (e < C.p * C.n + C.m) { const int idx = e - C.p * C.n, i = idx % C.p, j = idx / C.p; return s * C.b + C.n + i * C.mi + j; }
  return s * C.b + (e - C.p * C.n - C.m);
}

__device__ void band_load(Ctx& C, const Buffers& g, int inst, const double* Zsrc, const double* Lsrc) {
  const int n = C.n, m = C.m;
  C.xf = g.xf + (size_t)inst * n; C.Q = g.Q + (size_t)inst * n; C.R = g.R + (size_t)inst * m; C.uf = g.uf + (size_t)inst * m;
  C.lam = g.conlam + (size_t)inst * C.K * C.nrow; C.mu = g.conmu + (size_t)inst * C.K * C.nrow;
  const double* z = Zsrc + (size_t)inst * C.N * (n + m);
  for (int q = C.tid; q < C.N * (n + m); q += C.nt) {
    const int k = q / (n + m), e = q % (n + m);
    if (e < n) C.X[k * n + e] = z[q]; else C.U[k * m + (e - n)] = z[q];
  }
  const double* l = Lsrc + (size_t)inst * C.p * C.K * n;
  for (int q = C.tid; q < C.p * C.K * n; q += C.nt) C.L[q] = l[q];
  __syncthreads();
  for (int a = C.tid; a < n; a += C.nt) C.X[a] = g.x0[(size_t)inst * n + a];            // x_1 ← x0 (primal_dual_traj.jl:42)
  __syncthreads();
}
__device__ void band_store(Ctx& C, const Buffers& g, int inst) {
  const int n = C.n, m = C.m;
  double* z = g.Z + (size_t)inst * C.N * (n + m);
  for (int q = C.tid; q < C.N * (n + m); q += C.nt) {
    const int k = q / (n + m), e = q % (n + m);
    z[q] = (e < n) ? C.X[k * n + e] : C.U[k * m + (e - n)];
  }
  double* l = g.L + (size_t)inst * C.p * C.K * n;
  for (int q = C.tid; q < C.p * C.K * n; q += C.nt) l[q] = C.L[q];
}

// newton_solve!(prob) (solver_methods.jl:5-65) for every instance — or, with only_status >= 0, for the instances a previous
// kernel left with that status (the structured solver's AGB_SINGULAR ones), from the same initial iterate.
__global__ void __launch_bounds__(kBandThreads, 3) agb_band_newton_kernel(const DevDesc* __restrict__ dd, agb_options o, Buffers g, int inst0, int batch, int only_status) {
  __shared__ double red[64];
  AGB_DYN_SMEM(sm);
  if ((int)blockIdx.x >= g.band_slots) return;
  if (only_status >= 0) {                              // fallback launch: normally nothing to do — one coalesced scan of the range's statuses
    int none = 1;
    for (int i = inst0 + (int)threadIdx.x; i < batch; i += (int)blockDim.x) none &= (g.status[i] != only_status);
    if (__syncthreads_and(none)) return;
  }
  Ctx C;
  C.bind(dd, g.band + (size_t)blockIdx.x * g.band_stride, red);
  C.win = g.band_win > 0 ? sm : nullptr;
  const int K = C.K, n = C.n, m = C.m;
  const double Sd = (double)C.S;
#ifdef AGB_BAND_TIMING
  const long long bt_k0 = clock64();
#endif
  for (int inst = inst0 + blockIdx.x; inst < batch; inst += gridDim.x) {   // instances [inst0, batch)
    __syncthreads();
    if (only_status >= 0 && g.status[inst] != only_status) continue;
    band_load(C, g, inst, g.Z0, g.L0);
    if (only_status >= 0 && g.conlam0 != nullptr) {                          // fallback: the multipliers the first attempt started from
      for (int q = C.tid; q < K * C.nrow; q += C.nt) { C.lam[q] = g.conlam0[(size_t)inst * K * C.nrow + q]; C.mu[q] = g.conmu0[(size_t)inst * K * C.nrow + q]; }
      __syncthreads();
    }
    C.rollout();                                                              // :17
    if (o.dual_reset) {                                                       // :25 reset!(game_con)
      for (int q = C.tid; q < K * C.nrow; q += C.nt) { C.lam[q] = 0.0; C.mu[q] = o.rho_0; }
      __syncthreads();
    }
    int n_newton = 0, n_eval = 0, outer_done = 0, failed = 0, n_rec = 0, last_exit = AGB_MAX_OUTER;
    auto log_record = [&](const Norms& r, double dlt, int kk, int ll) {          // record!(stats, …) (statistics.jl:44-57)
      if (g.hist != nullptr && C.tid == 0 && n_rec < g.hist_max) {
        double* h = g.hist + ((size_t)inst * g.hist_max + n_rec) * AGB_NHIST;
        h[0] = (double)kk; h[1] = r.sum / Sd; h[2] = r.dyn; h[3] = r.con; h[4] = r.sta; h[5] = r.opt; h[6] = dlt; h[7] = (double)ll; h[8] = -1.0; h[9] = 0.0;
      }
      n_rec++;
    };
    double delta = 0.0;
    Norms rec = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int kout = 1; kout <= o.outer_iter; kout++) {                        // :30
      outer_done = kout;
      int ls_count = 0;
      last_exit = AGB_MAX_OUTER;
      for (int l = 1; l <= o.inner_iter; l++) {                               // :38
        const double l2 = (double)l * (double)l, reg = o.reg_0 * (l2 * l2);   // :39
        rec = C.assemble(C.X, C.U, C.L, C.X, C.U, 0.0, true, reg);            // residual! + residual_jacobian! (:73-86)
        n_eval++;
        const double res_norm = rec.sum / Sd;
        log_record(rec, delta, kout, l);
        delta = 0.0;
        if (!(rec.sum == rec.sum) || isinf(rec.sum)) { if (!failed) failed = AGB_NONFINITE; break; }
        if (rec.opt < o.eps_opt) break;                                       // :80-82
        if (!C.band_solve()) { if (!failed) failed = AGB_SINGULAR; break; }   // :87
        n_newton++;
        C.scatter_step();
        double alpha = 1.0; int j = 1;                                        // line_search (:105-125)
        while (j < o.ls_iter) {
          C.axpy_traj(alpha, C.Xt, C.Ut, C.Lt);
          const Norms t = C.assemble(C.Xt, C.Ut, C.Lt, C.X, C.U, o.regularize ? reg : 0.0, false, 0.0);
          n_eval++;
          if (t.sum / Sd <= (1.0 - alpha * o.beta) * res_norm) break;
          alpha *= o.alpha_decrease; j++;
        }
        ls_count = (j == o.ls_iter) ? ls_count + 1 : 0;                       // :92-93
        const double acc = C.delta_sum();
        C.axpy_traj(alpha, C.X, C.U, C.L);                                    // :94 (taken even when the search failed)
        delta = alpha * acc / (double)(K * (n + m));                          // Δ_step (primal_dual_traj.jl:130-147)
        if (delta < o.delta_min) { last_exit = AGB_STALLED; break; }          // :96-98
        if (ls_count >= 1) { last_exit = AGB_LINE_SEARCH_FAILED; break; }     // :43
        if (!(delta == delta)) { if (!failed) failed = AGB_NONFINITE; break; }
      }
      if (failed) break;
      if (kout == o.outer_iter || (rec.dyn < o.eps_dyn && rec.con < o.eps_con && rec.sta < o.eps_sta && rec.opt < o.eps_opt)) break;   // :49-55
      C.dual_penalty_update(o);                                               // :57-61
    }
    rec = C.assemble(C.X, C.U, C.L, C.X, C.U, 0.0, false, 0.0);               // final record (:63)
    n_eval++;
    log_record(rec, delta, outer_done, 0);
    if (g.hist != nullptr && C.tid == 0) g.hist_count[inst] = n_rec;
    const bool finite = (rec.sum == rec.sum) && !isinf(rec.sum);
    const bool conv = finite && rec.dyn < o.eps_dyn && rec.con < o.eps_con && rec.sta < o.eps_sta && rec.opt < o.eps_opt;
    band_store(C, g, inst);
    if (C.tid == 0) {
      double* st = g.stats + (size_t)inst * AGB_NSTATS;
      st[0] = rec.sum / Sd; st[1] = rec.dyn; st[2] = rec.con; st[3] = rec.sta; st[4] = rec.opt;
      st[5] = delta; st[6] = (double)n_newton; st[7] = (double)outer_done; st[8] = (double)n_eval; st[9] = (double)(failed != 0) + (only_status >= 0 ? 2.0 : 0.0);
      g.status[inst] = conv ? AGB_CONVERGED : (failed ? failed : (!finite ? AGB_NONFINITE : last_exit));
    }
  }
#ifdef AGB_BAND_TIMING
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    band::g_bt[15] += clock64() - bt_k0;
    printf("BAND_TIMING pass1 %lld zero %lld pass2 %lld norms %lld elim %lld backsub %lld total %lld\n", band::g_bt[0], band::g_bt[1], band::g_bt[2],
           band::g_bt[3], band::g_bt[4], band::g_bt[5], band::g_bt[15]);
    printf("BAND_TIMING_ELIM search %lld setup %lld update %lld tail+barrier %lld\n", band::g_bt[6], band::g_bt[7], band::g_bt[8], band::g_bt[9]);
  }
#endif
}

// per-function entry points of the band solver on the resident batch (parity tests; outputs in the reference's orders)
__global__ void __launch_bounds__(kBandThreads) agb_band_op_kernel(const DevDesc* __restrict__ dd, agb_options o, Buffers g, OpArgs a, int batch) {
  __shared__ double red[64];
  AGB_DYN_SMEM(sm);
  if ((int)blockIdx.x >= g.band_slots) return;
  Ctx C;
  C.bind(dd, g.band + (size_t)blockIdx.x * g.band_stride, red);
  C.win = g.band_win > 0 ? sm : nullptr;
  const int Sz = C.S;
  for (int inst = blockIdx.x; inst < batch; inst += gridDim.x) {
    __syncthreads();
    band_load(C, g, inst, g.Z, g.L);
    double* D = g.D + (size_t)inst * Sz;                                      // last Newton step, time-major band numbering
    switch (a.op) {
      case OP_ROLLOUT: C.rollout(); band_store(C, g, inst); break;
      case OP_RESIDUAL: {
        Norms r;
        if (a.alpha == 0.0) r = C.assemble(C.X, C.U, C.L, C.X, C.U, 0.0, false, 0.0);
        else {
          for (int q = C.tid; q < Sz; q += C.nt) C.rhs[q] = D[q];
          __syncthreads();
          C.scatter_step();
          C.axpy_traj(a.alpha, C.Xt, C.Ut, C.Lt);
          r = C.assemble(C.Xt, C.Ut, C.Lt, C.X, C.U, a.reg_x, false, 0.0);
        }
        if (a.out0) for (int q = C.tid; q < Sz; q += C.nt) a.out0[(size_t)inst * Sz + ref_row(C, q)] = C.res[q];
        if (a.out1 && C.tid == 0) { double* nr = a.out1 + (size_t)inst * 5; nr[0] = r.sum / (double)Sz; nr[1] = r.dyn; nr[2] = r.con; nr[3] = r.sta; nr[4] = r.opt; }
      } break;
      case OP_JAC_DENSE: {
        C.assemble(C.X, C.U, C.L, C.X, C.U, 0.0, true, a.reg_x);
        double* J = a.out0 + (size_t)inst * Sz * Sz;                          // zeroed by the caller
        for (size_t q = C.tid; q < (size_t)Sz * C.wd; q += C.nt) {
          const int r = (int)(q / C.wd), c = r + (int)(q % C.wd) - C.kl;
          if (c >= 0 && c < Sz && C.bandm[q] != 0.0) J[(size_t)ref_row(C, r) * Sz + ref_col(C, c)] = C.bandm[q];
        }
      } break;
      case OP_KKT_SOLVE: {
        C.assemble(C.X, C.U, C.L, C.X, C.U, 0.0, true, a.reg_x);
        const bool ok = C.band_solve();
        for (int q = C.tid; q < Sz; q += C.nt) { D[q] = C.rhs[q]; if (a.out0) a.out0[(size_t)inst * Sz + ref_col(C, q)] = C.rhs[q]; }
        if (a.iout && C.tid == 0) a.iout[inst] = ok ? 0 : 1;
      } break;
      case OP_EVAL_CON: {
        for (int item = C.tid; item < C.K * (1 + C.p); item += C.nt) {
          const int grp = item % (1 + C.p), s = item / (1 + C.p);
          double* out = a.out0 + ((size_t)inst * C.K + s) * C.nrow;
          if (grp == 0) C.control_rows(C.U + s * C.m, [&](int row, double c, int, const int*, const double*) { out[row] = c; });
          else C.state_rows(grp - 1, C.X + (s + 1) * C.n, [&](int row, double c, int, const int*, const double*) { out[row] = c; });
        }
      } break;
      default: break;
    }
  }
  (void)o;
}

void launch_band_solve(const DevDesc* dd, const agb_options& o, const Buffers& g, int inst0, int batch, int only_status, int grid, cudaStream_t st) {
  AGB_LAUNCH(agb_band_newton_kernel, grid, kBandThreads, g.band_win, st, dd, o, g, inst0, batch, only_status);
}
void launch_band_op(const DevDesc* dd, const agb_options& o, const Buffers& g, const OpArgs& a, int batch, int grid, cudaStream_t st) {
  AGB_LAUNCH(agb_band_op_kernel, grid, kBandThreads, g.band_win, st, dd, o, g, a, batch);
}
// Shared-memory bytes of band_solve_window's window: (kl+2) row slots of (kl+ku+2 | 1) doubles + the two slot maps; 0 when a row
// does not fit the per-lane register cache (8 x 32 entries) or the window exceeds `limit` (then the band is eliminated in global memory).
size_t band_window_bytes(const DevDesc& d, size_t limit) {
  const size_t CW = (size_t)d.kl + d.ku + 1, WR = (size_t)d.kl + 2, WS = ((CW + 1 + 31) / 32) * 32 + 1;   // as band_solve_window<NCH>
  if (CW + 1 > 256) return 0;
  size_t RW = 32; while (RW < CW + 1) RW <<= 1;
  const size_t bytes = (WR * WS + WR + RW + 4 * (kBandThreads / 32 + 1)) * sizeof(double) + 16;   // window | slot maps | x ring | pivot candidates
  return bytes <= limit ? bytes : 0;
}
// Opt the two band kernels in to `bytes` of dynamic shared memory (once per device: the attribute is only ever raised) and return
// how many CTAs of the solve kernel fit one SM.
int band_prepare(size_t bytes, int device) {
  static size_t raised[64] = {0};
  if (device >= 0 && device < 64 && bytes > raised[device]) {
    cudaFuncSetAttribute((const void*)agb_band_newton_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    cudaFuncSetAttribute((const void*)agb_band_op_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    raised[device] = bytes;
  }
  int occ = 2;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, agb_band_newton_kernel, kBandThreads, bytes);
  return occ > 0 ? occ : 1;
}
size_t band_scratch_doubles(const DevDesc& d) {
  const size_t N = d.N, K = d.K, n = d.n, m = d.m, p = d.p, S = d.S;
  return 3 * (N * n + N * m + p * K * n) + 2 * S + K * p * d.ni * (d.ni + d.mi) + K * n + S * (size_t)d.wd + 8;
}

}  // namespace agb
