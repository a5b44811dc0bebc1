// agb_kernels.cuh — the two instance kernels (templated on player count and model) and their launchers.
#pragma once
#include <cuda_runtime.h>
#include "agb_solver.cuh"

namespace agb {

// Kernel launch / dynamic shared memory go through two macros so that the test-only CTA emulator under tests/emu/
// can compile this very file with g++ (AGB_EMULATE); the shipped library is always the nvcc build.
#ifndef AGB_EMULATE
#define AGB_DYN_SMEM(name) extern __shared__ __align__(16) double name[]
#define AGB_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

// resident CTAs per SM the register allocation is capped for, per layout (DevDesc::big; see the table above the launchers)
#ifndef AGB_LAY2_CTAS
#define AGB_LAY2_CTAS 4
#endif
#define AGB_MIN_CTAS(LAY) ((LAY) == 1 ? 2 : ((LAY) == 3 ? 3 : ((LAY) == 2 ? AGB_LAY2_CTAS : 4)))

// =============================================================================================================
// Kernels
// =============================================================================================================
// newton_solve!(prob) for every instance of the batch (solver_methods.jl:5-65); one CTA per instance.
template <int P, int MODEL, int LAY>
__global__ void __launch_bounds__(threads_for(P), AGB_MIN_CTAS(LAY)) agb_newton_solve_kernel(const DevDesc* __restrict__ dd, agb_options o, Buffers g, int inst0, int batch, MpcArgs mp) {
  AGB_DYN_SMEM(sm);
  Inst<P, MODEL, (LAY != 0)> I;
  I.bind(dd, sm);
  constexpr int n = Inst<P, MODEL, (LAY != 0)>::n, kThreads = threads_for(P);
  const int K = I.K;
  const double S = (double)(K * Inst<P, MODEL, (LAY != 0)>::b);
  for (int inst = inst0 + blockIdx.x; inst < batch; inst += gridDim.x) {   // instances [inst0, batch)
    __syncthreads();
    if (g.force_singular > 0 && inst % g.force_singular == 0) {                          // test hook of the band fallback
      if (I.tid == 0) { g.status[inst] = AGB_SINGULAR; for (int q = 0; q < AGB_NSTATS; q++) g.stats[(size_t)inst * AGB_NSTATS + q] = 0.0; }
      continue;
    }
    I.bind_instance(g, inst);
    I.load_params(g, inst);
    I.load_iterate(g.Z0, g.L0, inst, &g);
    __syncthreads();
    for (int a = I.tid; a < n; a += kThreads) I.X[a] = g.x0[(size_t)inst * n + a];     // x_1 ← x0 (primal_dual_traj.jl:42)
#ifdef AGB_PHASE_TIMING
#define AGB_PROFK(k) do { const long long t_ = clock64(); I.prof[k] += t_ - I.prof_t; I.prof_t = t_; } while (0)
#else
#define AGB_PROFK(k) do { } while (0)
#endif
    // Receding horizon (mp.resolves > 0): this CTA runs ALL re-solves of its stream back to back — solve, apply the first
    // control (x0 ← x_{1+s} + disturbance), shift the iterate by s knots in shared memory, re-solve keeping multipliers and
    // penalties (Options.shift / dual_reset = false, options.jl:16-17) — with no launch boundary and no other stream to wait for.
    for (int t_mpc = 0;; t_mpc++) {
    __syncthreads();
    AGB_PROFK(11);
    I.rollout();                                                                        // :17
    AGB_PROFK(15);
    if (o.dual_reset && t_mpc == 0) I.reset_duals_penalties(o);                         // :25
    AGB_PROFK(11);
    int n_newton = 0, n_eval = 0, outer_done = 0, failed = 0, n_rec = 0;
    // record!(stats, …) (statistics.jl:44-57): optional per-instance log of every record of the solve
    auto log_record = [&](const Acc& r, double dlt, int kk, int ll) {
      if (g.hist != nullptr && I.tid == 0) {
        if (n_rec < g.hist_max) {
          double* hrec = g.hist + ((size_t)inst * g.hist_max + n_rec) * AGB_NHIST;
          hrec[0] = (double)kk; hrec[1] = r.sum / S; hrec[2] = r.dyn; hrec[3] = r.con; hrec[4] = r.sta; hrec[5] = r.opt;
          hrec[6] = dlt; hrec[7] = (double)ll; hrec[8] = -1.0; hrec[9] = 0.0;
        }
      }
      n_rec++;
    };
    bool kept = false;               // R's successor (rows at the accepted trial point = the current iterate) is in Rtrial
    Acc kept_rec = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double delta = 0.0;
    Acc rec = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    bool rec_current = false;        // rec was evaluated at the resident iterate with the resident multipliers (nothing moved since)
    int last_exit = AGB_MAX_OUTER;   // how the last inner loop ended (status of a non-converged instance)
    for (int kout = 1; kout <= o.outer_iter; kout++) {                                  // :30
      outer_done = kout;
      int ls_count = 0;
      last_exit = AGB_MAX_OUTER;
      for (int l = 1; l <= o.inner_iter; l++) {                                         // :38
        const double l2 = (double)l * (double)l;
        const double reg = o.reg_0 * (l2 * l2);                                         // :39
        // ---- inner_iteration (:67-103)
        if (kept) { I.load_kept_residual(); rec = kept_rec; kept = false; AGB_PROFK(10); }   // accepted trial point: already evaluated
        else { rec = I.template residual<false>(0.0, 0.0, 0.0, I.R); n_eval++; AGB_PROFK(0); }   // :73-75 (the reg terms vanish at Z)
        rec_current = true;
        const double res_norm = rec.sum / S;                                            // :76
        log_record(rec, delta, kout, l);                                                // :75
        delta = 0.0;
        if (!(rec.sum == rec.sum) || isinf(rec.sum)) { if (!failed) failed = AGB_NONFINITE; break; }
        if (rec.opt < o.eps_opt) break;                                                 // :80-82
        if (!I.kkt_solve(reg, reg)) { if (!failed) failed = AGB_SINGULAR; }             // :84-88
        n_newton++;
        double alpha; int j;
        kept = I.line_search(o, reg, res_norm, alpha, j, n_eval, &kept_rec);            // :91
        AGB_PROFK(1);
        ls_count = (j == o.ls_iter) ? ls_count + 1 : 0;                                 // :92-93
        delta = I.update_traj(alpha);                                                   // :94-95 (taken even when the search failed)
        rec_current = false;
        AGB_PROFK(9);
        if (delta < o.delta_min) { last_exit = AGB_STALLED; break; }                    // :96-98
        if (ls_count >= 1) { last_exit = AGB_LINE_SEARCH_FAILED; break; }               // :43
        if (!(delta == delta)) { if (!failed) failed = AGB_NONFINITE; break; }
      }
      if (failed) break;
      if (kout == o.outer_iter || (rec.dyn < o.eps_dyn && rec.con < o.eps_con && rec.sta < o.eps_sta && rec.opt < o.eps_opt))
        break;                                                                          // :49-55
      I.dual_update(o);                                                                 // :57-58
      I.penalty_update(o);                                                              // :61
      kept = false; rec_current = false;                                                // multipliers changed
    }
    AGB_PROFK(11);
    if (rec_current) { }                                                                // :63 final record — the inner loop left on its ϵ_opt test: same point, same numbers
    else if (kept) { I.load_kept_residual(); rec = kept_rec; }
    else { rec = I.template residual<false>(0.0, 0.0, 0.0, I.R); n_eval++; }
    AGB_PROFK(13);
    log_record(rec, delta, outer_done, 0);                                              // :63
    if (g.hist != nullptr && I.tid == 0) g.hist_count[inst] = n_rec;
    const bool finite = (rec.sum == rec.sum) && !isinf(rec.sum);
    const bool conv = finite && rec.dyn < o.eps_dyn && rec.con < o.eps_con && rec.sta < o.eps_sta && rec.opt < o.eps_opt;
    if (mp.resolves > 0) {
      if (I.tid == 0) {                                                                 // this re-solve's record: [resolves][batch][…]
        double* st = mp.stats + ((size_t)t_mpc * batch + inst) * AGB_NSTATS;
        st[0] = rec.sum / S; st[1] = rec.dyn; st[2] = rec.con; st[3] = rec.sta; st[4] = rec.opt;
        st[5] = delta; st[6] = (double)n_newton; st[7] = (double)outer_done; st[8] = (double)n_eval; st[9] = (double)(failed != 0);
        mp.status[(size_t)t_mpc * batch + inst] = conv ? AGB_CONVERGED : (failed ? failed : (!finite ? AGB_NONFINITE : last_exit));
      }
      if (mp.xs != nullptr)                                                             // executed state x_{1+s} + disturbance
        for (int a = I.tid; a < n; a += kThreads)
          mp.xs[((size_t)t_mpc * batch + inst) * n + a] = I.X[mp.shift * n + a] + (mp.dist ? mp.dist[((size_t)t_mpc * batch + inst) * n + a] : 0.0);
      if (t_mpc + 1 < mp.resolves) {                                                    // (the host applies the last advance)
        __syncthreads();
        AGB_PROFK(11);
        I.mpc_shift(mp.shift, mp.dist ? mp.dist + ((size_t)t_mpc * batch + inst) * n : nullptr);
        AGB_PROFK(14);
        continue;
      }
    }
    I.store_iterate(g.Z, g.L, inst);
    I.store_duals(g, inst);
    AGB_PROFK(11);
#ifdef AGB_PHASE_TIMING
    if (g.hist != nullptr && g.hist_max >= 2 && I.tid == 0)
      for (int q = 0; q < 16; q++) g.hist[(size_t)inst * g.hist_max * AGB_NHIST + q] = (double)I.prof[q];
#endif
    if (I.tid == 0) {
      double* st = g.stats + (size_t)inst * AGB_NSTATS;
      st[0] = rec.sum / S; st[1] = rec.dyn; st[2] = rec.con; st[3] = rec.sta; st[4] = rec.opt;
      st[5] = delta; st[6] = (double)n_newton; st[7] = (double)outer_done; st[8] = (double)n_eval; st[9] = (double)(failed != 0);
      g.status[inst] = conv ? AGB_CONVERGED : (failed ? failed : (!finite ? AGB_NONFINITE : last_exit));
    }
    break;
    }   // re-solves of this stream
  }
}

// ibr_newton_solve!(prob; ibr_opts) for every instance (solver_methods.jl:133-224); one CTA per instance.
template <int P, int MODEL, int LAY>
__global__ void __launch_bounds__(threads_for(P), AGB_MIN_CTAS(LAY)) agb_ibr_solve_kernel(const DevDesc* __restrict__ dd, agb_options o,
                                                                                          agb_ibr_options io, Buffers g, int batch) {
  AGB_DYN_SMEM(sm);
  Inst<P, MODEL, (LAY != 0)> I;
  I.bind(dd, sm);
  constexpr int n = Inst<P, MODEL, (LAY != 0)>::n, kThreads = threads_for(P);
  const int K = I.K;
  for (int inst = blockIdx.x; inst < batch; inst += gridDim.x) {
    __syncthreads();
    I.bind_instance(g, inst);
    I.load_params(g, inst);
    I.load_iterate(g.Z0, g.L0, inst, &g);
    __syncthreads();
    for (int a = I.tid; a < n; a += kThreads) I.X[a] = g.x0[(size_t)inst * n + a];
    __syncthreads();
    I.rollout();                                                                        // :145
    int n_newton = 0, n_eval = 0, failed = 0, sweeps = 0, n_rec = 0;
    // record!(stats, prob, model, game_con, pdtraj, t_elap, Δ, k, i) (statistics.jl:59-72): player-i violations next to the
    // FULL-game residual norm; that extra residual! is evaluated only while a history buffer is set
    auto log_record = [&](const Acc& r, double dlt, int kk, int ll, int player, int sweep) {
      if (g.hist == nullptr) return;
      I.pl = -1;
      const Acc full = I.template residual<false>(0.0, 0.0, 0.0, nullptr);    // rows are not stored (R keeps player i's)
      I.pl = player;
      if (I.tid == 0 && n_rec < g.hist_max) {
        double* hrec = g.hist + ((size_t)inst * g.hist_max + n_rec) * AGB_NHIST;
        hrec[0] = (double)kk; hrec[1] = full.sum / (double)(K * Inst<P, MODEL, (LAY != 0)>::b); hrec[2] = r.dyn; hrec[3] = r.con; hrec[4] = r.sta;
        hrec[5] = r.opt; hrec[6] = dlt; hrec[7] = (double)ll; hrec[8] = (double)player; hrec[9] = (double)sweep;
      }
      n_rec++;
    };
    unsigned change = (1u << P) - 1u;                                                   // Δ_change = trues(p) (:150)
    double delta = 0.0, dmax = 0.0;                                                     // dmax = maximum(stats.Δ_traj)
    for (int q = 0; q < io.ibr_iter && !failed; q++) {                                  // :151
      sweeps = q + 1;
      for (int id = 0; id < P && !failed; id++) {
        const int i = io.ordering[id];
        I.pl = i;
        // ---- ibr_newton_solve!(prob, i) (:168-224)
        if (o.dual_reset) {                                                             // :179-183
          I.reset_duals_penalties(o);
          for (int t = I.tid; t < P * K * n; t += kThreads) I.L[t] = 0.0;               // reset_duals!(pdtraj)
          __syncthreads();
        }
        const double S = I.res_size();
        delta = 0.0;
        bool kept = false;
        int out = 0;
        Acc rec = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, kept_rec = rec;
        for (int kout = 1; kout <= o.outer_iter; kout++) {
          out = kout;
          int ls_count = 0;
          for (int l = 1; l <= o.inner_iter; l++) {
            const double l2 = (double)l * (double)l;
            const double reg = o.reg_0 * (l2 * l2);
            if (kept) { I.load_kept_residual(); rec = kept_rec; kept = false; }
            else { rec = I.template residual<false>(0.0, 0.0, 0.0, I.R); n_eval++; }    // ibr_inner_iteration (:226-265)
            const double res_norm = rec.sum / S;
            log_record(rec, delta, kout, l, i, q + 1);                                  // :237
            delta = 0.0;
            if (!(rec.sum == rec.sum) || isinf(rec.sum)) { if (!failed) failed = AGB_NONFINITE; break; }
            if (rec.opt < o.eps_opt) break;
            if (!I.kkt_solve(reg, reg)) { if (!failed) failed = AGB_SINGULAR; }
            n_newton++;
            double alpha; int j;
            kept = I.line_search(o, reg, res_norm, alpha, j, n_eval, &kept_rec);
            ls_count = (j == o.ls_iter) ? ls_count + 1 : 0;
            delta = I.update_traj(alpha);
            dmax = fmax(dmax, delta);
            if (delta < o.delta_min) break;
            if (ls_count >= 1) break;
            if (!(delta == delta)) { if (!failed) failed = AGB_NONFINITE; break; }
          }
          if (failed) break;
          if (kout == o.outer_iter || (rec.dyn < o.eps_dyn && rec.con < o.eps_con && rec.sta < o.eps_sta && rec.opt < o.eps_opt))
            break;
          I.dual_update(o);
          I.penalty_update(o);
          kept = false;
        }
        if (g.hist != nullptr && !failed) {                                             // final record!(…, out, i) (:222)
          if (kept) { I.load_kept_residual(); rec = kept_rec; kept = false; }
          else rec = I.template residual<false>(0.0, 0.0, 0.0, I.R);
          log_record(rec, delta, out, 0, i, q + 1);
        }
        if (!(io.delta_min > dmax)) change |= (1u << i); else change &= ~(1u << i);     // :156
      }
      if (change == 0u) break;                                                          // :163
    }
    I.pl = -1;
    Acc rec = I.template residual<false>(0.0, 0.0, 0.0, I.R);                           // residual!(prob, prob.pdtraj) (:160)
    n_eval++;
    const double S = I.res_size();
    const bool finite = (rec.sum == rec.sum) && !isinf(rec.sum);
    const bool conv = finite && rec.dyn < o.eps_dyn && rec.con < o.eps_con && rec.sta < o.eps_sta && rec.opt < o.eps_opt;
    I.store_iterate(g.Z, g.L, inst);
    I.store_duals(g, inst);
    if (I.tid == 0) {
      double* st = g.stats + (size_t)inst * AGB_NSTATS;
      st[0] = rec.sum / S; st[1] = rec.dyn; st[2] = rec.con; st[3] = rec.sta; st[4] = rec.opt;
      st[5] = delta; st[6] = (double)n_newton; st[7] = (double)sweeps; st[8] = (double)n_eval; st[9] = (double)(failed != 0);
      g.status[inst] = conv ? AGB_CONVERGED : (failed ? failed : (!finite ? AGB_NONFINITE : AGB_MAX_OUTER));
    }
    if (g.hist != nullptr && I.tid == 0) g.hist_count[inst] = n_rec;
  }
}

// Per-function entry points on the resident batch (parity tests and stand-alone use of the exported reference API).
template <int P, int MODEL, int LAY>
__global__ void __launch_bounds__(threads_for(P)) agb_op_kernel(const DevDesc* __restrict__ dd, agb_options o, Buffers g, OpArgs a, int batch) {
  AGB_DYN_SMEM(sm);
  Inst<P, MODEL, (LAY != 0)> I;
  I.bind(dd, sm);
  constexpr int n = Inst<P, MODEL, (LAY != 0)>::n, m = Inst<P, MODEL, (LAY != 0)>::m, b = Inst<P, MODEL, (LAY != 0)>::b, kThreads = threads_for(P);
  const int K = I.K, Sz = K * b, nrow = I.nrow;
  for (int inst = blockIdx.x; inst < batch; inst += gridDim.x) {
    __syncthreads();
    I.bind_instance(g, inst);
    I.load_params(g, inst);
    I.load_iterate(g.Z, g.L, inst, &g);
    __syncthreads();
    for (int q = I.tid; q < n; q += kThreads) I.X[q] = g.x0[(size_t)inst * n + q];
    __syncthreads();
    double* D = g.D + (size_t)inst * Sz;
    I.pl = a.player;
    const double S = I.res_size();
    switch (a.op) {
      case OP_ROLLOUT: {
        I.rollout();
        I.store_iterate(g.Z, g.L, inst);
      } break;
      case OP_RESIDUAL: {
        Acc r;
        double* out = a.out0 ? a.out0 + (size_t)inst * Sz : nullptr;
        if (a.alpha == 0.0) {
          r = I.template residual<false>(0.0, 0.0, 0.0, I.R);
          if (out) for (int q = I.tid; q < Sz; q += kThreads) out[q] = I.R[q];
        } else {
          for (int q = I.tid; q < Sz; q += kThreads) I.R[q] = D[q];
          __syncthreads();
          r = I.template residual<true>(a.alpha, a.reg_x, a.reg_u, out);
        }
        if (a.out1 && I.tid == 0) {
          double* nr = a.out1 + (size_t)inst * 5;
          nr[0] = r.sum / S; nr[1] = r.dyn; nr[2] = r.con; nr[3] = r.sta; nr[4] = r.opt;
        }
      } break;
      case OP_JAC_DENSE: {
        // residual_jacobian! + regularize_residual_jacobian! written block by block in the reference's
        // (vertical, horizontal) order (core/newton_core.jl:40-89); J_out must be zeroed by the caller.
        I.template residual<false>(0.0, 0.0, 0.0, I.R);
        __syncthreads();
        double* J = a.out0 + (size_t)inst * Sz * Sz;
        const int vdyn = P * K * (n + 2);
        for (int item = I.tid; item < P * K; item += kThreads) {
          const int s = item % K, i = item / K, k = s + 1;
          const size_t vx = (size_t)(i * K + s) * (n + 2), vu = vx + n;
          for (int r = 0; r < n; r++) {
            double* row = J + (vx + r) * Sz;
            for (int c = 0; c < n; c++) {
              row[s * b + c] = I.h_entry(i, k, r, c, a.reg_x);                                               // (opt_i x_k, x_k)
            }
            row[s * b + n + m + i * n + r] = -1.0;                               // (opt_i x_k, λ_{i,k-1}) = −I
            if (k < K) {
              const int cr = r / P, ir = r % P;
              for (int q = 0; q < 4; q++) row[k * b + n + m + i * n + q * P + ir] = I.Ael(k, ir, q, cr);   // A_kᵀ
            }
          }
          for (int j = 0; j < 2; j++) {
            double* row = J + (vu + j) * Sz;
            row[s * b + n + i * 2 + j] = I.hu_entry(s, j * P + i, a.reg_u);       // (opt_i u_ik, u_ik)
            for (int q = 0; q < 4; q++) row[s * b + n + m + i * n + q * P + i] = I.Bel(s, i, q, j);          // B_iᵀ
          }
        }
        for (int item = I.tid; item < K * n; item += kThreads) {
          const int r = item % n, s = item / n;
          const int cr = r / P, ir = r % P;
          double* row = J + (size_t)(vdyn + s * n + r) * Sz;
          if (s > 0) for (int q = 0; q < 4; q++) row[(s - 1) * b + q * P + ir] = I.Ael(s, ir, cr, q);        // A_s
          for (int j = 0; j < 2; j++) row[s * b + n + ir * 2 + j] = I.Bel(s, ir, cr, j);                    // B_i
          row[s * b + r] = -1.0;                                                                            // −I
        }
      } break;
      case OP_KKT_SOLVE: {
        I.template residual<false>(0.0, 0.0, 0.0, I.R);
        __syncthreads();
        const bool ok = I.kkt_solve(a.reg_x, a.reg_u);
        for (int q = I.tid; q < Sz; q += kThreads) D[q] = I.R[q];
        if (a.iout && I.tid == 0) a.iout[inst] = ok ? 0 : 1;
      } break;
      case OP_LINE_SEARCH: {
        for (int q = I.tid; q < Sz; q += kThreads) I.R[q] = D[q];
        __syncthreads();
        Acc r0 = I.template residual<true>(0.0, 0.0, 0.0, nullptr);
        double alpha; int j, ne = 0;
        const double reg = a.reg_x;
        I.line_search(o, reg, r0.sum / S, alpha, j, ne);
        if (I.tid == 0) { a.out0[inst] = alpha; a.iout[inst] = j; }
      } break;
      case OP_UPDATE: {
        for (int q = I.tid; q < Sz; q += kThreads) I.R[q] = D[q];
        __syncthreads();
        const double delta = I.update_traj(a.in0[inst]);
        __syncthreads();
        I.store_iterate(g.Z, g.L, inst);
        if (a.out0 && I.tid == 0) a.out0[inst] = delta;
      } break;
      case OP_DUAL_UPDATE: { I.dual_update(o); I.store_duals(g, inst); } break;
      case OP_PENALTY_UPDATE: { I.penalty_update(o); I.store_duals(g, inst); } break;
      case OP_RESET: { I.reset_duals_penalties(o); I.store_duals(g, inst); } break;
      case OP_EVAL_CON: {
        for (int q = I.tid; q < K * nrow; q += kThreads) a.out0[(size_t)inst * K * nrow + q] = I.con_value(q / nrow, q % nrow);
      } break;
      case OP_ACTIVE_SET: {
        for (int q = I.tid; q < K * nrow; q += kThreads) {
          const double c = I.con_value(q / nrow, q % nrow);
          a.bout[(size_t)inst * K * nrow + q] = ((c >= -a.tol) || (I.CL[q] > 0.0)) ? 1 : 0;
        }
      } break;
      case OP_VIOLATIONS: {                 // per-knot vectors of struct/violations.jl (full game, or player a.player)
        I.template residual<false>(0.0, 0.0, 0.0, I.R);
        __syncthreads();
        const int N = K + 1;
        double* o_dyn = a.out0 + (size_t)inst * (4 * N - 2);
        double* o_con = o_dyn + K; double* o_sta = o_con + K; double* o_opt = o_sta + N;
        const int pl = I.pl;
        for (int s = I.tid; s < K; s += kThreads) {
          double v = 0.0;
          for (int q = 0; q < n; q++) if (pl < 0 || q % P == pl) v = fmax(v, fabs(I.R[s * b + Inst<P, MODEL, (LAY != 0)>::OD + q]));
          o_dyn[s] = v;
          double c = 0.0;
          if (pl < 0) { for (int r = I.d->nrow_state; r < nrow; r++) c = fmax(c, I.con_value(s, r)); }
          else { for (int j = 0; j < 2; j++) { const int local = j * P + pl; if (local < I.d->nrow_control) c = fmax(c, I.con_value(s, I.d->nrow_state + local)); } }
          o_con[s] = c;
        }
        for (int k = I.tid; k < N; k += kThreads) {
          double v = 0.0, w = 0.0;
          if (k > 0) {
            for (int r = 0; r < I.d->nrow_state; r++) if (pl < 0 || I.row_player(r) == pl) v = fmax(v, I.con_value(k - 1, r));
            for (int q = 0; q < P * n; q++) if (pl < 0 || q / n == pl) w = fmax(w, fabs(I.R[(k - 1) * b + q]));
          }
          if (k < K) for (int q = 0; q < m; q++) if (pl < 0 || q % P == pl) w = fmax(w, fabs(I.R[k * b + P * n + q]));
          o_sta[k] = v; o_opt[k] = w;
        }
      } break;
      case OP_GAIN_SOLVE: {
        constexpr int W = Inst<P, MODEL, (LAY != 0)>::W;
        for (int q = I.tid; q < m * W; q += kThreads) I.Aug[q] = a.in0[(size_t)inst * m * W + q];
        __syncthreads();
        int ok = 1;
        if (I.warp == 0) ok = I.gj_warp(I.Aug) ? 1 : 0;
        __syncthreads();
        for (int q = I.tid; q < m * W; q += kThreads) a.out0[(size_t)inst * m * W + q] = I.Aug[q];
        if (I.tid == 0) a.iout[inst] = ok;
      } break;
      default: break;
    }
  }
}


// ---- launchers: one translation unit per player count (agb_kernels_p<P>.cu) instantiates these ---------------------
struct LaunchArgs {
  int model, grid;
  size_t smem;
  cudaStream_t stream;
  const DevDesc* dd;
  agb_options o;
  agb_ibr_options io;
  Buffers g;
  OpArgs a;
  int batch;
  int inst0;             // first instance of the launch (newton_solve only; chunked host pipeline)
  MpcArgs mp;            // receding-horizon loop inside the solve kernel (resolves <= 1: one plain solve)
};

template <int P, int MODEL, int LAY> inline cudaError_t set_attr_pm(size_t smem) {
  cudaError_t e = cudaFuncSetAttribute((const void*)agb_newton_solve_kernel<P, MODEL, LAY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute((const void*)agb_ibr_solve_kernel<P, MODEL, LAY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute((const void*)agb_op_kernel<P, MODEL, LAY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
template <int P, int LAY> inline cudaError_t set_attr_p(int model, size_t smem) {
  switch (model) {
    case AGB_MODEL_DOUBLE_INTEGRATOR: return set_attr_pm<P, AGB_MODEL_DOUBLE_INTEGRATOR, LAY>(smem);
    case AGB_MODEL_UNICYCLE: return set_attr_pm<P, AGB_MODEL_UNICYCLE, LAY>(smem);
    default: return set_attr_pm<P, AGB_MODEL_BICYCLE, LAY>(smem);
  }
}
template <int P, int MODEL, int LAY> inline void launch_solve_pm(const LaunchArgs& L) {
  auto kfn = agb_newton_solve_kernel<P, MODEL, LAY>;
  AGB_LAUNCH(kfn, L.grid, threads_for(P), L.smem, L.stream, L.dd, L.o, L.g, L.inst0, L.batch, L.mp);
}
template <int P, int MODEL, int LAY> inline void launch_ibr_pm(const LaunchArgs& L) {
  auto kfn = agb_ibr_solve_kernel<P, MODEL, LAY>;
  AGB_LAUNCH(kfn, L.grid, threads_for(P), L.smem, L.stream, L.dd, L.o, L.io, L.g, L.batch);
}
template <int P, int LAY> inline void launch_ibr_p(const LaunchArgs& L) {
  switch (L.model) {
    case AGB_MODEL_DOUBLE_INTEGRATOR: launch_ibr_pm<P, AGB_MODEL_DOUBLE_INTEGRATOR, LAY>(L); break;
    case AGB_MODEL_UNICYCLE: launch_ibr_pm<P, AGB_MODEL_UNICYCLE, LAY>(L); break;
    default: launch_ibr_pm<P, AGB_MODEL_BICYCLE, LAY>(L); break;
  }
}
template <int P, int MODEL, int LAY> inline void launch_op_pm(const LaunchArgs& L) {
  auto kfn = agb_op_kernel<P, MODEL, LAY>;
  AGB_LAUNCH(kfn, L.grid, threads_for(P), L.smem, L.stream, L.dd, L.o, L.g, L.a, L.batch);
}
template <int P, int LAY> inline void launch_solve_p(const LaunchArgs& L) {
  switch (L.model) {
    case AGB_MODEL_DOUBLE_INTEGRATOR: launch_solve_pm<P, AGB_MODEL_DOUBLE_INTEGRATOR, LAY>(L); break;
    case AGB_MODEL_UNICYCLE: launch_solve_pm<P, AGB_MODEL_UNICYCLE, LAY>(L); break;
    default: launch_solve_pm<P, AGB_MODEL_BICYCLE, LAY>(L); break;
  }
}
template <int P, int LAY> inline void launch_op_p(const LaunchArgs& L) {
  switch (L.model) {
    case AGB_MODEL_DOUBLE_INTEGRATOR: launch_op_pm<P, AGB_MODEL_DOUBLE_INTEGRATOR, LAY>(L); break;
    case AGB_MODEL_UNICYCLE: launch_op_pm<P, AGB_MODEL_UNICYCLE, LAY>(L); break;
    default: launch_op_pm<P, AGB_MODEL_BICYCLE, LAY>(L); break;
  }
}

// layouts (DevDesc::big): 0 = small, 4 CTAs/SM (agb_kernels_p1..p3.cu); 1 = big storage, 2 CTAs/SM, up to 255 registers
// (agb_kernels_p3b.cu, agb_kernels_p4.cu); 2 / 3 = big storage, 4 CTAs/SM at 128 registers / 3 CTAs/SM at 168 (agb_kernels_p3m.cu, agb_kernels_p3t.cu: mid-size 3-player games)
cudaError_t set_attr(int p, int big, int model, size_t smem);
void launch_solve(int p, int big, const LaunchArgs& L);
void launch_op(int p, int big, const LaunchArgs& L);
void launch_ibr(int p, int big, const LaunchArgs& L);
#define AGB_DECLARE_P(PP)                                   \
  cudaError_t set_attr_p##PP(int model, size_t smem);      \
  void launch_solve_p##PP(const LaunchArgs& L);            \
  void launch_op_p##PP(const LaunchArgs& L);              \
  void launch_ibr_p##PP(const LaunchArgs& L);
AGB_DECLARE_P(1) AGB_DECLARE_P(2) AGB_DECLARE_P(3) AGB_DECLARE_P(3b) AGB_DECLARE_P(3m) AGB_DECLARE_P(3t) AGB_DECLARE_P(4)

}  // namespace agb
