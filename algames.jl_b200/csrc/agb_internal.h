// agb_internal.h — device-side problem descriptor shared by the kernels and the C-ABI layer.
// Not part of the public ABI (include/algames_b200.h is).
#pragma once
#include "../../include/algames_b200.h"

namespace agb {

// one CTA of 4 warps owns one game instance.  Instances whose working set would leave a single CTA per SM (4 players; long
// horizons with many constraint rows) use the "big" layout: duals, AL multipliers/penalties and the pair/self Hessian
// blocks live in L2-resident global memory instead of shared memory, so that two CTAs fit per SM.
#ifdef __CUDACC__
#define AGB_HD __host__ __device__
#else
#define AGB_HD
#endif
AGB_HD constexpr int threads_for(int) { return 128; }
// stages of closed-loop blocks in the forward sweep's shared-memory ring (laid over the factorisation scratch, which
// build_dev_desc sizes for it): as deep as the scratch of a p-player game allows
AGB_HD constexpr int fwd_ring_depth(int p) { return p == 1 ? 3 : (p == 2 ? 6 : (p == 3 ? 8 : 12)); }
constexpr int kMaxWarps = 8;

// Flattened, index-resolved form of agb_problem_desc, lives in device global memory (read through L1).
// Index conventions follow the reference's component-major joint layout (src/dynamics/unicycle.jl:18-20):
// state comp c of player i -> c*p+i, control comp j of player i -> j*p+i.
struct DevDesc {
  int model, p, n, m, N, K, b, S;      // K = N-1 stages, b = p*n+m+n rows per stage, S = K*b (problem_size.jl:22)
  int ni, mi;                          // states / controls per player (4, 2; QuadrotorGame 12, 4)
  int use_band;                        // 1: the schema has no structured-kernel form (band solver only)
  double dt, lf, lr;
  double quad_mass;                    // QuadrotorGame
  int spherical;                       // collision avoidance on three position components
  int n_walls3d[AGB_MAX_P], wall3d_row[AGB_MAX_P];
  double walls3d[AGB_MAX_P][AGB_MAX_WALLS][12];
  int n_cyl[AGB_MAX_P], cyl_row[AGB_MAX_P];
  double cyl[AGB_MAX_P][AGB_MAX_WALLS][6];
  int sb_ncon[AGB_MAX_P];              // StateBound convals of player i (band solver: rows enumerated conval by conval)
  int x_max_con[AGB_MAX_P][AGB_MAX_N], x_min_con[AGB_MAX_P][AGB_MAX_N];
  int has_x_max[AGB_MAX_P][AGB_MAX_N], has_x_min[AGB_MAX_P][AGB_MAX_N];
  // band solver: KKT band in time-major order, rows [dyn_s | opt u_s | opt x_{s+1}], columns [λ_s | u_s | x_{s+1}]
  int kl, ku, wd;                      // lower / upper bandwidth, stored row width 2·kl + ku + 1 (room for the pivoting fill)
  // objective (objective.jl:84-100)
  int has_cc, npairs;
  int has_pairs, has_self, has_sb, has_cb;  // + any state bound; any control bound                  // any pair term (collision cost / avoidance); any wall / circle
  double cc_radius[AGB_MAX_P], cc_mu[AGB_MAX_P];
  // state-constraint rows of player i at one knot, canonical order:
  // [collision j!=i ascending | state bound max rows | min rows | walls | circles]
  double col_radius[AGB_MAX_P][AGB_MAX_P];
  int col_row[AGB_MAX_P][AGB_MAX_P];       // row inside the stage row block, -1 = no constraint
  int sbmax_row[AGB_MAX_P][AGB_MAX_N];     // row of the x_max bound on joint comp a owned by player i, -1 = none
  int sbmin_row[AGB_MAX_P][AGB_MAX_N];
  double x_max[AGB_MAX_P][AGB_MAX_N], x_min[AGB_MAX_P][AGB_MAX_N];
  int n_walls[AGB_MAX_P], wall_row[AGB_MAX_P];
  double walls[AGB_MAX_P][AGB_MAX_WALLS][6];
  int n_circles[AGB_MAX_P], circle_row[AGB_MAX_P];
  double circles[AGB_MAX_P][AGB_MAX_CIRCLES][3];
  int srow_off[AGB_MAX_P + 1];             // first state row of player i; srow_off[p] = nrow_state
  // control-bound rows (control_bound_constraint.jl:33-35): finite u_max rows then finite u_min rows
  int ub_row[AGB_MAX_M], lb_row[AGB_MAX_M];   // row inside the stage row block, -1 = none
  double u_max[AGB_MAX_M], u_min[AGB_MAX_M];
  int nrow_state, nrow_control, nrow;
  // CW (active weights) is kept for bound rows only: compact index = row - sb_shift[i] (state bounds of player i) or
  // row - cb_shift (control bounds); ncw compact rows per stage
  int sb_shift[AGB_MAX_P], cb_shift, ncw;
  // shared-memory layout (offsets in doubles)
  int o_X, o_U, o_L, o_R, o_KU, o_AB, o_CL, o_CM, o_CW, o_Gp, o_Hp, o_Gs, o_Hs, o_P, o_Sv, o_Y, o_Aug, o_Base, o_W, o_Ta, o_par, o_red;
  int smem_doubles;
  int big;                                 // layout: 0 small; 1, 2, 3: L, CL, CM, Hp, Hs in global memory (2, 4, 3 CTAs per SM)
};

// Per-batch device buffers (all FP64 unless noted); layouts as in include/algames_b200.h.
struct Buffers {
  const double* x0;   // [B][n]
  const double* xf;   // [B][n]
  const double* Q;    // [B][n]
  const double* R;    // [B][m]
  const double* uf;   // [B][m]
  const double* Z0;   // [B][N][n+m]  initial iterate as init_traj! leaves it
  const double* L0;   // [B][p][K][n]
  double* Z;          // resident iterate / result
  double* L;
  double* conlam;     // [B][K][nrow]
  double* conmu;
  double* D;          // [B][S] Newton step, internal stage-major layout
  double* KUg;        // [B][K][m][n+2] (rows padded to even length) feedback gains of the stage-wise factorisation (L2-resident scratch)
  double* stats;      // [B][AGB_NSTATS]
  int* status;        // [B]
  double* hist;       // [B][hist_max][AGB_NHIST] record!(stats, …) log, or nullptr
  int* hist_count;    // [B]
  int hist_max;
  double* band;       // band solver scratch: [slots][band_stride]
  size_t band_stride;
  int band_slots;
  int band_win;       // bytes of the shared-memory elimination window per CTA (band_window_bytes), 0: eliminate in the global band
  const double* conlam0;   // multipliers / penalties the solve started from (fallback re-solves with dual_reset = 0), or nullptr
  const double* conmu0;
  int force_singular; // test hook (AGB_TEST_FORCE_SINGULAR=k): the structured solve kernel reports every k-th instance as AGB_SINGULAR
  double* Hpg;        // big layout only: [B][N·p(p-1)·3 + N·p·3] pair / self Hessian blocks
  int hpg_stride;
};

// Receding-horizon loop run inside the solve kernel (agb_mpc_run): every stream's CTA does `resolves` × (newton_solve!, advance by
// `shift` knots) on its own.  All pointers are device memory; resolves == 0 means one plain solve and nothing else here is read.
struct MpcArgs {
  int resolves, shift;
  const double* dist;   // [resolves][B][n] disturbance added to the executed state, or nullptr
  double* stats;        // [resolves][B][AGB_NSTATS]
  int* status;          // [resolves][B]
  double* xs;           // [resolves][B][n] executed states (x0 of the next re-solve), or nullptr
};

enum Op {
  OP_ROLLOUT = 0, OP_RESIDUAL, OP_JAC_DENSE, OP_KKT_SOLVE, OP_LINE_SEARCH, OP_UPDATE,
  OP_DUAL_UPDATE, OP_PENALTY_UPDATE, OP_RESET, OP_EVAL_CON, OP_ACTIVE_SET, OP_GAIN_SOLVE, OP_VIOLATIONS
};

struct OpArgs {
  int op;
  double reg_x, reg_u, alpha;
  double tol;
  double* out0;          // device staging, meaning depends on op
  double* out1;
  int* iout;
  unsigned char* bout;
  const double* in0;     // e.g. per-instance alpha for OP_UPDATE
  int player;            // iterative best response: player whose problem the op addresses, -1 = the full game
};

}  // namespace agb
