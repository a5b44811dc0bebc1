/*
 * algames_b200.h — C ABI of libalgames_b200.so: the batched ALGAMES Newton/KKT +
 * augmented-Lagrangian solve on NVIDIA B200 (sm_100a).
 *
 * The reference (RoboticExplorationLab/Algames.jl) has no FFI boundary of its own
 * (SURVEY.md F8): the drop-in boundary is its exported Julia API.  Every entry point
 * below replaces one exported reference function on a *batch* of independent
 * GameProblems that share one schema (model, players, horizon, constraint lists) and
 * differ in x0 / targets / weights.  The Julia-side binding (`ccall`) is shown in
 * INTEGRATION.md and shipped in julia/AlgamesB200.jl.
 *
 * Conventions
 *  - plain pointers and sizes only; all arrays are C row-major, FP64, caller-owned host
 *    memory unless the name says `_dev`.  A Julia Array{Float64} with reversed dims has
 *    exactly this layout.
 *  - every function returns 0 on success, <0 on error (agb_last_error() gives the text);
 *    nothing throws across the boundary; one bad instance never aborts the batch
 *    (per-instance `status`).
 *  - calls on one handle must be serialised by the caller; host-buffer calls are
 *    synchronous (return after the stream has drained).
 *  - state / control / multiplier vectors use the REFERENCE layouts:
 *      x, xf, Q   [n]  component-major joint state   (src/dynamics/unicycle.jl:18-20)
 *      u, uf, R   [m]  component-major joint control
 *      Z     [B][N][n+m]      knot-major primal trajectory, z_k = [x_k; u_k]       (struct/primal_dual_traj.jl:5-20)
 *      L     [B][p][N-1][n]   dynamics multipliers λ_{i,k}                          (primal_dual_traj.jl:9)
 *      res   [B][S]           KKT residual, reference row ("vertical") order        (core/newton_core.jl:40-63)
 *      dtraj [B][S]           Newton step, reference column ("horizontal") order    (core/newton_core.jl:65-89)
 *      conlam/conmu [B][N-1][nrow]  AL multipliers/penalties: stage k holds the state-constraint
 *                 rows of knot k+1 — per player [collision j≠i ascending | state bounds: per conval max rows, min rows |
 *                 walls | circles | 3-D walls | cylinders] — followed by the control-bound rows of knot k [u_max rows, u_min rows]
 *                 (finite bounds only, reference component order; control_bound_constraint.jl:33-35).
 *
 * Environment variables read by the library (test / tuning hooks, never needed in production):
 *   AGB_FORCE_BIG_LAYOUT=1|2|3  agb_create: force a big-storage shared-memory layout on 3-player instances (parity tests
 *                               run every layout on small games with it)
 *   AGB_HOST_CHUNKS=1..32       agb_solve_from_host: number of copy/solve pipeline chunks (default 8 for batch >= 1024)
 *   AGB_HOST_GRAPH=0            agb_solve_from_host: never replay the pipeline as a CUDA graph (default: captured on the second call
 *                               with the same page-locked buffers and options, replayed afterwards)
 *   AGB_BAND_FALLBACK=0         agb_create: do not re-solve AGB_SINGULAR instances of the structured kernels with the band solver
 *   AGB_TEST_FORCE_SINGULAR=k   agb_create: the structured solve reports every k-th instance as AGB_SINGULAR (fallback tests)
 *   AGB_BAND_WINDOW=0           agb_create: the band solver eliminates in the global-memory band, not in its shared-memory window
 */
#ifndef ALGAMES_B200_H
#define ALGAMES_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define AGB_MAX_P 4      /* players                                                        */
#define AGB_MAX_N 48     /* joint state dim   n = ni·p  (ni = 4; QuadrotorGame: 12)        */
#define AGB_MAX_M 16     /* joint control dim m = mi·p  (mi = 2; QuadrotorGame: 4)         */
#define AGB_MAX_WALLS 8  /* walls per player              */
#define AGB_MAX_CIRCLES 8
#define AGB_NSTATS 10

enum { AGB_MODEL_DOUBLE_INTEGRATOR = 0, AGB_MODEL_UNICYCLE = 1, AGB_MODEL_BICYCLE = 2, AGB_MODEL_QUADROTOR = 3 };

/* Two solvers share the ABI.  The STRUCTURED kernels (stage-wise block elimination in shared memory) cover
 * the planar models with 4 states / 2 controls per player and planar constraints.  The BAND solver (explicit KKT band in
 * global memory, LU with partial pivoting — the reference's own formulation, solver_methods.jl:87) covers every schema:
 * QuadrotorGame, the 3-D constraints, and it re-solves, transparently, any instance on which the structured elimination
 * meets a singular stage system (status AGB_SINGULAR) although the KKT matrix itself is regular.  agb_problem_desc.solver
 * picks: AGB_SOLVER_AUTO (structured where the schema allows, band as fallback / for the rest) or AGB_SOLVER_BAND. */
enum { AGB_SOLVER_AUTO = 0, AGB_SOLVER_BAND = 1 };

/* per-instance status written by agb_newton_solve_batch.  Codes 1-3 all mean "outer_iter exhausted with a
 * tolerance unmet" (the reference keeps iterating through failed line searches and stalls, solver_methods.jl:38-55); they
 * tell how the LAST inner loop ended.  AGB_IS_NOT_CONVERGED / AGB_IS_NUMERICAL_FAILURE group them. */
enum {
  AGB_CONVERGED = 0,          /* final record: dyn, con, sta, opt maxima all below their ϵ (solver_methods.jl:49-53)     */
  AGB_MAX_OUTER = 1,          /* not converged; last inner loop ran out of inner_iter or stopped on opt < ϵ_opt (:80-82)  */
  AGB_LINE_SEARCH_FAILED = 2, /* not converged; last inner loop ended on a failed line search, j == ls_iter (:43, :92-93) */
  AGB_STALLED = 3,            /* not converged; last inner loop ended on Δ_traj < Δ_min (:96-98)                          */
  AGB_SINGULAR = 4,           /* zero / non-finite pivot in the structured KKT factorisation and in its pivoted fallback  */
  AGB_NONFINITE = 5           /* non-finite residual or step (NaN / Inf inputs)                                           */
};
#define AGB_IS_NOT_CONVERGED(s) ((s) >= AGB_MAX_OUTER && (s) <= AGB_STALLED)
#define AGB_IS_NUMERICAL_FAILURE(s) ((s) >= AGB_SINGULAR)

/* error codes */
enum { AGB_OK = 0, AGB_EINVAL = -1, AGB_ECUDA = -2, AGB_ENOMEM = -3, AGB_EUNSUPPORTED = -4 };

/* Problem schema shared by the whole batch.
 * Mirrors GameProblem's constructor arguments (src/problem/problem.jl:35-53):
 * model (dynamics/*.jl), GameObjective (objective/objective.jl:12-35, :84-100),
 * GameConstraintValues + adders (constraints/constraints_methods.jl:5-195).           */
typedef struct agb_problem_desc {
  int model;                 /* AGB_MODEL_*                                              */
  int p;                     /* players (1..4); n = ni·p, m = mi·p (ni, mi = 4, 2; quadrotor 12, 4) */
  int d;                     /* DoubleIntegratorGame dimension: 2 (structured kernels) or 3 (band solver)      */
  int N;                     /* knots                                                    */
  double dt;
  double lf, lr;             /* BicycleGame (dynamics/bicycle.jl:15)                     */
  /* LQR defaults, joint component-major (objective.jl:24-28 expand_vector); may be
   * overridden per instance with agb_set_instance_params                              */
  double Q[AGB_MAX_N], R[AGB_MAX_M], xf[AGB_MAX_N], uf[AGB_MAX_M];
  /* add_collision_cost!(game_obj, radius, μ)  (objective.jl:84-100)                   */
  int has_collision_cost;
  double cc_radius[AGB_MAX_P], cc_mu[AGB_MAX_P];
  /* add_collision_avoidance!: pair radius r_i+r_j for the ordered pair (i,j); 0 = none */
  double col_radius[AGB_MAX_P][AGB_MAX_P];
  /* add_control_bound!: ±INFINITY = no bound                                          */
  int has_control_bound;
  double u_max[AGB_MAX_M], u_min[AGB_MAX_M];
  /* add_state_bound!(game_con, i, x_max, x_min): bound on the JOINT state, owned by player i.
   * has_state_bound[i] = number of StateBoundConstraint convals of player i (0 = none); several convals
   * (add_velocity_bound! gives every player one per bounded velocity, velocity_constraint.jl:13-28) are
   * merged component-wise into x_max / x_min, x_max_con / x_min_con (end of this struct) naming the
   * 0-based conval that owns each finite entry.  Two convals of one player bounding the same component
   * from the same side cannot be expressed (the host layers reject that schema).                      */
  int has_state_bound[AGB_MAX_P];
  double x_max[AGB_MAX_P][AGB_MAX_N], x_min[AGB_MAX_P][AGB_MAX_N];
  /* add_wall_constraint!: (x1,y1,x2,y2,xv,yv) per wall per player                      */
  int n_walls[AGB_MAX_P];
  double walls[AGB_MAX_P][AGB_MAX_WALLS][6];
  /* add_circle_constraint!: (xc,yc,r)                                                  */
  int n_circles[AGB_MAX_P];
  double circles[AGB_MAX_P][AGB_MAX_CIRCLES][3];
  /* owning conval (0 .. has_state_bound[i]-1) of each finite x_max / x_min entry; all zero for a single
   * add_state_bound! per player.  Rows follow the reference's conval order: for each conval in the order
   * it was added, its finite x_max rows then its finite x_min rows (state_bound_constraint.jl:33-35).   */
  int x_max_con[AGB_MAX_P][AGB_MAX_N], x_min_con[AGB_MAX_P][AGB_MAX_N];
  /* ---- QuadrotorGame and the 3-D constraints (dynamics/quadrotor.jl, constraints_methods.jl:45-81, :201-285) ---------- */
  double quad_mass;          /* QuadrotorGame(mass = 0.5); inertia, rotor geometry and constants are the reference's (:22-29) */
  int spherical_collision;   /* add_spherical_collision_avoidance!: col_radius acts on the first THREE state components  */
  /* add_wall_constraint!(game_con, i, walls::Vector{Wall3D}): corner points p1, p2, p3 and outward normal v per wall      */
  int n_walls3d[AGB_MAX_P];
  double walls3d[AGB_MAX_P][AGB_MAX_WALLS][12];
  /* add_wall_constraint!(game_con, i, walls::Vector{CylinderWall}): base point p(3), axis (0,1,2 = x,y,z), length, radius */
  int n_cylinders[AGB_MAX_P];
  double cylinders[AGB_MAX_P][AGB_MAX_WALLS][6];
  int solver;                /* AGB_SOLVER_AUTO / AGB_SOLVER_BAND                                                          */
} agb_problem_desc;

/* Live fields of Options (src/struct/options.jl:5-116; dead fields omitted). */
typedef struct agb_options {
  double reg_0;                 /* :27  */
  int regularize;               /* :21  */
  double alpha_decrease;        /* :37  */
  double beta;                  /* :40  */
  int ls_iter;                  /* :43  */
  double delta_min;             /* :46  */
  double rho_0;                 /* :50  initial penalty μ0 */
  double rho_increase;          /* :56  ϕ  */
  double rho_max;               /* :59  μ_max */
  double lambda_max;            /* :62  */
  double alpha_dual;            /* :65  */
  double alphax_dual[AGB_MAX_P];/* :68  */
  double active_set_tolerance;  /* :71 (only used by agb_active_set) */
  double eps_dyn, eps_sta, eps_con, eps_opt; /* :75-84 */
  int outer_iter, inner_iter;   /* :88, :91 */
  int dual_reset;               /* :115 */
} agb_options;

/* IBROptions (src/struct/options.jl:123-136); players in `ordering` are 0-based here. */
typedef struct agb_ibr_options {
  int ibr_iter;                 /* :126 (default 100) */
  int ordering[AGB_MAX_P];      /* :129 (default 0,1,2,…) */
  double delta_min;             /* :132 (default 1e-9) */
} agb_ibr_options;

typedef struct agb_handle agb_handle;

/* Fills *o with the reference defaults (options.jl). */
void agb_default_options(agb_options* o);

/* Sizes derived from a descriptor (problem_size.jl:22 and the row schema above). */
typedef struct agb_sizes { int n, m, p, N, S, nrow, nrow_state, nrow_control; } agb_sizes;
int agb_sizes_of(const agb_problem_desc* d, agb_sizes* out);

/* Layout check for bindings that mirror the structs above by hand (ctypes, Julia): `layout` lists, in this order,
 *   sizeof(agb_problem_desc), offsetof(.., dt), offsetof(.., Q), offsetof(.., col_radius), offsetof(.., has_state_bound),
 *   offsetof(.., walls), offsetof(.., circles), offsetof(.., x_max_con), offsetof(.., quad_mass), offsetof(.., solver),
 *   sizeof(agb_options), offsetof(.., alphax_dual), offsetof(.., eps_dyn), offsetof(.., dual_reset),
 *   sizeof(agb_ibr_options), offsetof(.., delta_min), sizeof(agb_sizes), sizeof(agb_device_view),
 *   AGB_MAX_P, AGB_MAX_N, AGB_MAX_M, AGB_MAX_WALLS, AGB_MAX_CIRCLES, AGB_NSTATS, AGB_NHIST, AGB_IPC_BYTES
 * (AGB_ABI_WORDS = 26 values).  Returns AGB_OK when every entry equals the library's own, else AGB_EINVAL with
 * agb_last_error(NULL) naming the first mismatch.  agb_abi_layout writes the library's values (for diagnostics). */
#define AGB_ABI_WORDS 26
int agb_abi_check(const int* layout, int count);
int agb_abi_layout(int* layout_out, int count);

/* GameProblem(...) for a batch: allocates all device state.  device = CUDA ordinal.   */
int agb_create(const agb_problem_desc* desc, int batch, int device, agb_handle** out);
void agb_destroy(agb_handle* h);
const char* agb_last_error(const agb_handle* h);   /* h may be NULL (creation errors)  */
int agb_get_sizes(const agb_handle* h, agb_sizes* out);

/* Per-instance x0 [B][n] and optional overrides (NULL keeps the descriptor default):
 * xf [B][n], Q [B][n], R [B][m], uf [B][m].                                            */
int agb_set_instance_params(agb_handle* h, const double* x0, const double* xf,
                            const double* Q, const double* R, const double* uf);

/* Initial iterate as init_traj! leaves it (primal_dual_traj.jl:29-44; x_1 is overwritten
 * with x0 by the solver): Z0 [B][N][n+m], L0 [B][p][N-1][n].  conlam/conmu [B][N-1][nrow]
 * may be NULL (zeros / rho_0 are used when dual_reset, else the resident values).      */
int agb_set_initial(agb_handle* h, const double* Z0, const double* L0,
                    const double* conlam, const double* conmu);
int agb_get_state(agb_handle* h, double* Z, double* L, double* conlam, double* conmu);

/* MPC warm start = init_traj! with shift s (primal_dual_traj.jl:34-41): the resident
 * solution is shifted by s knots; the tail is filled from Zfresh/Lfresh (same shapes
 * as Z0/L0; only the last s knots are read).  New x0 comes from agb_set_instance_params. */
int agb_shift_initial(agb_handle* h, int s, const double* Zfresh, const double* Lfresh);
/* One receding-horizon step, entirely on device: x0 <- x_{1+s} of the resident solution (+ disturbance [B][n], may be
 * NULL), then agb_shift_initial(s, Zfresh, Lfresh).  Together with agb_newton_solve_async(dual_reset = 0) this is the MPC
 * loop Options.shift / Options.dual_reset exist for (struct/options.jl:16-17, :114-115); no host round trip per step. */
int agb_mpc_advance(agb_handle* h, int s, const double* disturbance, const double* Zfresh, const double* Lfresh);
/* The same step with no host involvement at all: disturbance_dev is a DEVICE pointer [B][n] (or NULL), the tail knots are
 * zero, nothing synchronises.  agb_newton_solve_async + agb_mpc_advance_async, repeated, is one uninterrupted stream of
 * kernels (the ordering rules of agb_newton_solve_async apply). */
int agb_mpc_advance_async(agb_handle* h, int s, const double* disturbance_dev);
/* The whole receding-horizon loop of every stream in ONE kernel launch (config D: Options.shift = s, dual_reset = false after
 * the first solve; struct/options.jl:16-17, :114-115; init_traj! with s, primal_dual_traj.jl:34-41).  Stream b's CTA runs
 * `resolves` x (newton_solve!, x0 <- x_{1+s} + disturbance[t][b], shift by s knots, multipliers and penalties carried) back to
 * back in shared memory: no launch boundary, and a stream whose re-solve runs long holds nobody else.  Bit for bit the same
 * results as `resolves` x (agb_newton_solve_async, agb_mpc_advance_async); afterwards the handle holds the last solution advanced
 * once, exactly as that loop leaves it.  o->dual_reset applies to the first re-solve only.  A re-solve the structured kernel
 * reports AGB_SINGULAR is recorded as such and the stream carries on from it (the step-wise loop would re-solve it with the band
 * solver).  All pointers are DEVICE memory: disturbance_dev [resolves][B][n] or NULL; stats_dev [resolves][B][AGB_NSTATS];
 * status_dev [resolves][B]; xs_dev [resolves][B][n] executed states (the x0 of re-solve t+1), or NULL.  stream: a cudaStream_t, or
 * NULL for the handle's own.  QuadrotorGame / 3-D schemas (band solver) run the step-wise loop behind the same call. */
int agb_mpc_run_async(agb_handle* h, const agb_options* o, int resolves, int s, const double* disturbance_dev, double* stats_dev,
                      int* status_dev, double* xs_dev, void* stream);
/* The same with HOST buffers (disturbance may be NULL; stats_out / status_out / xs_out may be NULL), synchronous. */
int agb_mpc_run(agb_handle* h, const agb_options* o, int resolves, int s, const double* disturbance, double* stats_out,
                int* status_out, double* xs_out);
/* The handle's own stream (a cudaStream_t): pass it to agb_newton_solve_async to keep a device-resident loop on ONE stream. */
void* agb_get_stream(agb_handle* h);
/* Makes `stream` (a cudaStream_t) wait, on the device, for everything enqueued so far on the handle's own stream. */
int agb_join_stream(agb_handle* h, void* stream);

/* ---- per-function entry points (operate on the resident batch; parity tests) -------- */
/* rollout!(RK3, model, traj)                                  solver_methods.jl:17     */
int agb_rollout(agb_handle* h);
/* residual! + regularize_residual! at (Z + alpha·Δ) relative to Z, Δ = last agb_kkt_solve
 * step (alpha = 0 ⇒ residual!(prob, pdtraj)).  res_out [B][S] or NULL.
 * norms_out [B][5] or NULL: ‖res‖₁/S, dyn, con, sta, opt maxima (violations.jl:18-168). */
int agb_residual(agb_handle* h, double reg_x, double reg_u, double alpha,
                 double* res_out, double* norms_out);
/* residual_jacobian! + regularize_residual_jacobian!, dense [B][S][S] (small cases).    */
int agb_residual_jacobian_dense(agb_handle* h, double reg_x, double reg_u, double* J_out);
/* Δtraj = −(lu(jac) \ res)                                    solver_methods.jl:87-88  */
int agb_kkt_solve(agb_handle* h, double reg_x, double reg_u, double* dtraj_out);
/* line_search(prob, res_norm)                                 solver_methods.jl:105-125 */
int agb_line_search(agb_handle* h, const agb_options* o, double reg_x, double reg_u,
                    double* alpha_out, int* j_out);
/* update_traj!(pdtraj, pdtraj, α, Δpdtraj) + Δ_step           primal_dual_traj.jl:109-147 */
int agb_update_traj(agb_handle* h, const double* alpha, double* delta_step_out);
/* evaluate! + dual_update!, penalty_update!, reset!           constraints_methods.jl:295-440 */
int agb_dual_update(agb_handle* h, const agb_options* o);
int agb_penalty_update(agb_handle* h, const agb_options* o);
int agb_reset_duals_penalties(agb_handle* h, const agb_options* o);
/* constraint values c [B][N-1][nrow] at the resident iterate (evaluate!), and the active-set
 * predicate (c >= -tol) | (λ > 0) as 0/1 bytes (update_active_set!, constraints_methods.jl:396-415) */
int agb_evaluate_constraints(agb_handle* h, double* c_out);
int agb_active_set(agb_handle* h, double tol, unsigned char* active_out);

/* ---- Active-set analysis of the resident iterate (src/active_set/*.jl; off the solve path, planar models) ----------------------
 * ActiveSetCore (active_set_core.jl:57-160) borders the Newton system with one ROW per unordered collision pair i<j and knot
 * k = 2..N — stamp (:v,:col,i,j,k), row S + (k-2)·p(p-1)/2 + index of (i,j) — and one COLUMN per ordered pair i!=j and knot —
 * stamp (:h,:col,i,j,k), column S + (k-2)·p(p-1) + index of (i,j).  Sv = S + p(p-1)(N-1)/2 rows, Sh = S + p(p-1)(N-1) columns
 * (:81-82); the first S rows / columns are residual! / residual_jacobian! in the reference's order. */
int agb_active_set_sizes(agb_handle* h, int* Sv_out, int* Sh_out);
/* residual!(ascore, prob, pdtraj) (active_set_methods.jl:96-124): [KKT residual ; c_ij(x_k) of the pairs i<j].  res_out [B][Sv]. */
int agb_active_set_residual(agb_handle* h, double* res_out);
/* residual_jacobian!(ascore, prob, pdtraj) (:131-170): the unregularised KKT Jacobian bordered by grad c_ij' in column (h,i,j,k) on
 * the rows (opt_i, x_k), and by grad c_ij in row (v,i,j,k), i<j, on the columns x_k.  Dense, row-major: jac_out [B][Sv][Sh]. */
int agb_active_set_jacobian_dense(agb_handle* h, double* jac_out);
/* active_vertical_mask! / active_horizontal_mask! (:28-74): 1 for the Newton rows / columns and for the pairs whose collision row
 * is active — c >= -tol or lambda > 0 (active(game_con, stamp), :5-26).  vmask_out [B][Sv], hmask_out [B][Sh]. */
int agb_active_set_masks(agb_handle* h, double tol, unsigned char* vmask_out, unsigned char* hmask_out);
/* update_nullspace!(ascore, prob, pdtraj) (:173-184): an orthonormal basis of nullspace(jac[vmask, hmask]), scattered to the hmask
 * rows (add_matrix!, active_set_core.jl:29-42).  One CTA per instance: Gauss-Jordan with complete pivoting on the dense masked
 * matrix, rank = number of pivots > atol (the reference passes atol = 1e-20 to LinearAlgebra.nullspace), modified Gram-Schmidt.
 * The basis is unique only up to rotation: compare subspaces.  With atol below eps * max|jac| an exactly rank-deficient matrix
 * yields n_cols - min(n_rows, n_cols) vectors, the count the reference's SVD produces (round-off-level singular values pass its
 * atol test).  null_out [B][max_dim][Sh] (vector d of instance b at
 * null_out + (b*max_dim + d)*Sh), dim_out [B] = the dimension found (vectors beyond max_dim are not written). */
int agb_update_nullspace(agb_handle* h, double tol, double atol, int max_dim, double* null_out, int* dim_out);

/* Diagnostic: how this handle runs the band solver (lu(jac) \ res on the explicit KKT band, solver_methods.jl:87) — bytes of the
 * shared-memory elimination window per CTA (0: the window does not fit or AGB_BAND_WINDOW=0, the band is eliminated in global
 * memory) and the number of resident CTA slots with band scratch (0: structured kernels only, no fallback). */
int agb_band_info(const agb_handle* h, int* window_bytes_out, int* slots_out);

/* Diagnostic: runs the kernel's m x (m+n+1) gain-system solver (threshold-pivoted Gauss-Jordan with partial-pivoting
 * fallback) on caller-supplied systems aug [B][m][m+n+1]; the reduced systems come back in place of the
 * input layout in aug_out, ok_out[B] = 0 when a pivot was zero / non-finite.  Unit-test hook for the pivoting logic. */
int agb_debug_gain_solve(agb_handle* h, const double* aug, double* aug_out, int* ok_out);

/* ---- the hot path: newton_solve!(prob) for every instance ---------------------------- */
/* Host-buffer form.  All outputs may be NULL.  Z_out [B][N][n+m], L_out [B][p][N-1][n],
 * conlam_out/conmu_out [B][N-1][nrow], stats_out [B][AGB_NSTATS] =
 *   {‖res‖₁/S, dyn, con, sta, opt (final record), last Δ, Newton steps, outer iterations,
 *    residual evaluations, spare}, status_out [B].                                      */
int agb_newton_solve_batch(agb_handle* h, const agb_options* o,
                           double* Z_out, double* L_out, double* conlam_out, double* conmu_out,
                           double* stats_out, int* status_out);

/* ---- iterative best response (the reference's second solver, solver_methods.jl:133-289) --------------------------- */
/* ibr_newton_solve!(prob; ibr_opts) for every instance: sweeps over the players in `ordering`, each player solving its own
 * augmented-Lagrangian optimal-control problem with the others' strategies fixed (ibr_newton_solve!(prob, i), :168-224).
 * Outputs as agb_newton_solve_batch; stats = full-game record at the returned iterate, Newton steps and residual
 * evaluations summed over all best responses, stats[7] = IBR sweeps executed; status is taken on that full-game record. */
int agb_ibr_newton_solve_batch(agb_handle* h, const agb_options* o, const agb_ibr_options* io,
                               double* Z_out, double* L_out, double* conlam_out, double* conmu_out,
                               double* stats_out, int* status_out);
/* ibr_residual! + regularize_ibr_residual! of player `player` (0-based) at (Z + alpha·Δ): the rows of the other players
 * are returned as zero (the reference masks them out, newton_core.jl:205-245); norms_out [B][5] = masked ‖res‖₁ / mask
 * length and the player-i violations of record!(…, i) (statistics.jl:59-72). */
int agb_ibr_residual(agb_handle* h, int player, double reg_x, double reg_u, double alpha,
                     double* res_out, double* norms_out);
/* Δtraj[horiz_mask] = −(lu(jac[verti_mask, horiz_mask]) \ res[verti_mask]), zeros elsewhere (solver_methods.jl:248-250). */
int agb_ibr_kkt_solve(agb_handle* h, int player, double reg_x, double reg_u, double* dtraj_out);

/* One-shot host form: x0 [B][n] and the initial iterate Z0 / L0 (as agb_set_initial) go in, results come out, with the
 * batch cut into chunks whose host->device copy, solve and device->host copy are pipelined on separate streams (pass
 * page-locked buffers for the copies to overlap).  Equivalent to agb_set_instance_params(x0) + agb_set_initial(Z0, L0) +
 * agb_newton_solve_batch, except that the per-function "resident iterate" is the result, and o->dual_reset should be 1
 * (the multipliers start from the handle's resident values otherwise).  Outputs may be NULL.  With page-locked buffers the
 * second call with the same buffer addresses and options captures the whole pipeline as a CUDA graph and later calls replay it
 * with one launch (an MPC or Monte-Carlo loop refills the same buffers); agb_set_history invalidates the capture. */
int agb_solve_from_host(agb_handle* h, const agb_options* o, const double* x0, const double* Z0, const double* L0,
                        double* Z_out, double* L_out, double* conlam_out, double* conmu_out,
                        double* stats_out, int* status_out);

/* Device-resident form: enqueue the solve on `stream` (a cudaStream_t, 0 = legacy default)
 * with no host copies and no host synchronisation; results stay in the handle's device buffers.  Ordering: the solve
 * waits (on the device) for everything already enqueued on the handle's own stream, and every later call on the handle
 * (agb_get_state, agb_mpc_advance, agb_shift_initial, agb_residual, …, agb_allgather with a NULL stream) waits for the
 * solve — so agb_newton_solve_async + agb_mpc_advance is a correct device-resident MPC loop without host syncs. */
int agb_newton_solve_async(agb_handle* h, const agb_options* o, void* stream);

/* Raw device pointers of the resident results (for NCCL all-gather / zero-copy consumers). */
typedef struct agb_device_view {
  double* Z_dev;      /* [B][N][n+m]    */
  double* L_dev;      /* [B][p][N-1][n] */
  double* conlam_dev; /* [B][N-1][nrow] */
  double* conmu_dev;
  double* stats_dev;  /* [B][AGB_NSTATS] */
  int* status_dev;    /* [B] */
  double* x0_dev;     /* [B][n] */
  double* Z0_dev;     /* initial iterate [B][N][n+m] */
  double* L0_dev;
  /* Z_dev, L_dev, stats_dev and status_dev are consecutive slices of ONE allocation, so a single collective on
   * [results_dev, results_dev + results_bytes) moves every result of the batch (the path's only all-gather). */
  void* results_dev;
  unsigned long long results_bytes;
} agb_device_view;
int agb_get_device_view(agb_handle* h, agb_device_view* out);

/* Full convergence history of newton_solve!: one record per record!(stats, …) call of the reference
 * (src/struct/statistics.jl:44-57, called from inner_iteration src/problem/solver_methods.jl:75 and once more at :63).
 * agb_set_history(h, max_records) makes every later agb_newton_solve_batch / agb_solve_from_host keep up to max_records
 * records per instance on the device (0 switches the log off and frees it); agb_get_history copies them out:
 * hist_out [B][max_records][AGB_NHIST] = {outer iteration k, res = ‖res‖₁/S, dyn_vio.max, con_vio.max, sta_vio.max,
 * opt_vio.max, Δ_traj, inner iteration l (0 for the final record), player, sweep}, count_out [B] = records the solve
 * produced (may exceed max_records: later records were dropped).  newton_solve: player = -1, sweep = 0.
 * agb_ibr_newton_solve_batch logs record!(stats, …, k, i) (statistics.jl:59-72, called at solver_methods.jl:222 and :237):
 * res = FULL-game ‖res‖₁/S, the four maxima restricted to player i (violations.jl:28-37, 69-82, 116-134, 170-183),
 * player = i (0-based), sweep = IBR iteration q (1-based); the extra full-game residual is evaluated only while a
 * history buffer is set. */
#define AGB_NHIST 10
int agb_set_history(agb_handle* h, int max_records);
int agb_get_history(agb_handle* h, double* hist_out, int* count_out);

/* Per-knot violation vectors of the resident iterate (struct/violations.jl:5-16, 44-55, 84-96, 136-148): what
 * dynamics_violation / control_violation / state_violation / optimality_violation return in `.vio`:
 * dyn_out [B][N-1], con_out [B][N-1], sta_out [B][N], opt_out [B][N] (knot 1 of sta is always 0; opt of knot 1 holds the
 * u rows only, of knot N the x rows only).  Any pointer may be NULL.  The maxima are stats[1..4] of the solve. */
int agb_violations(agb_handle* h, double* dyn_out, double* con_out, double* sta_out, double* opt_out);

/* ---- multi-GPU: the batch is split contiguously over ranks — one agb_handle per GPU, in one process or in
 * one process per GPU — solves need no communication, and ONE all-gather collects every rank's result slab
 * [Z | L | stats | status] (agb_device_view.results_dev) on every rank.  The gather is a push over NVLink peer memory on
 * the copy engines (cudaMemcpyPeerAsync / IPC-mapped peer buffers): it takes no SM from solves that are still running.
 *   agb_peer_init        rank layout: batches[r] = instances of rank r (this handle's batch must equal batches[rank]);
 *                        allocates this rank's gather buffer [sum_r slab(r)].
 *   agb_peer_export      64-byte CUDA IPC handle of that buffer (one process per GPU; exchange them out of band,
 *                        e.g. torch.distributed.all_gather) — then agb_peer_connect with all ranks' handles.
 *   agb_peer_connect_local  all ranks live in THIS process: pass their handles (enables peer access between the devices).
 *   agb_allgather        enqueue: push this rank's slab into every rank's gather buffer, ordered after `stream`
 *                        (NULL = the handle's own stream, i.e. after every solve enqueued through the handle).
 *   agb_allgather_wait   host-blocks until this rank's pushes have landed.  A barrier between the ranks (or waiting on
 *                        every local handle) then makes every gather buffer complete.
 *   agb_gathered_view    device pointer of the gather buffer and the byte offset of every rank's slab in it.
 *   agb_unpack_gathered  copies rank r's slab out of the local gather buffer into host arrays (any may be NULL).
 *   agb_create_sharded   convenience for a single-process host (Julia): splits `batch` over `ndev` devices
 *                        (devices == NULL: ordinals 0..ndev-1; ndev <= 0: all visible devices), creates the handles and
 *                        runs agb_peer_init + agb_peer_connect_local.  handles_out must hold ndev (or device-count)
 *                        pointers; *ndev_out receives the count.  Destroy every handle with agb_destroy. */
#define AGB_IPC_BYTES 64
#define AGB_MAX_RANKS 64
int agb_peer_init(agb_handle* h, int nranks, int rank, const int* batches);
int agb_peer_export(agb_handle* h, unsigned char* ipc_out /* [AGB_IPC_BYTES] */);
int agb_peer_connect(agb_handle* h, const unsigned char* ipc_all /* [nranks][AGB_IPC_BYTES] */);
int agb_peer_connect_local(agb_handle* const* handles, int nranks);
int agb_allgather(agb_handle* h, void* stream);
int agb_allgather_wait(agb_handle* h);
int agb_gathered_view(agb_handle* h, void** gathered_dev, unsigned long long* offsets_out /* [nranks+1] bytes */);
int agb_unpack_gathered(agb_handle* h, int src_rank, double* Z, double* L, double* stats, int* status);
int agb_create_sharded(const agb_problem_desc* desc, int batch, int ndev, const int* devices,
                       agb_handle** handles_out, int* ndev_out);

/* FP64 vector peak of the device, measured: a register-resident kernel of independent DFMA chains (8 per thread, 1024
 * threads per CTA, 2 CTAs per SM), `iters` FMAs per chain; returns TFLOP/s (2 flops per DFMA) through *tflops_out and
 * the kernel time through *ms_out.  bench.py uses it as the denominator of roofline.frac (the solve is FP64-issue
 * bound, not HBM bound). */
int agb_measure_fp64_peak(int device, int iters, double* tflops_out, float* ms_out);

/* Number of kernels this library has launched on the handle since creation. */
long long agb_launch_count(const agb_handle* h);
/* Device time of the last agb_newton_solve_* kernel (CUDA events on its own stream), ms;
 * synchronises.                                                                         */
float agb_last_solve_ms(agb_handle* h);

#ifdef __cplusplus
}
#endif
#endif
