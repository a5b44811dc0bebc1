// agb_solver.cuh — device code of the batched ALGAMES Newton/KKT + augmented-Lagrangian solve (sm_100a, FP64).
//
// One CTA (128 threads; 256 for 4 players) owns one game instance; the whole iterate, the KKT right-hand side / Newton step, the
// feedback gains of the stage-wise factorisation and the AL multipliers stay in shared memory for the entire
// newton_solve! loop (reference: src/problem/solver_methods.jl:5-125).  The KKT system is never
// formed: in time-major order it is block tridiagonal, and its block-LU with the pivot order
// (λ_k via the -I blocks, then u_k through S_k = Hu + BᵀPB with partial pivoting, then x_k) is the game Riccati
// recursion implemented in kkt_solve() below.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>
#include "agb_internal.h"

namespace agb {

#define AGB_FULL 0xffffffffu

// 16-byte asynchronous global → shared copies (the emulator build copies synchronously)
#ifndef AGB_EMULATE
typedef unsigned SmemAddr;                                   // 32-bit shared-window address
__device__ __forceinline__ SmemAddr smem_addr(double* p) { return (SmemAddr)__cvta_generic_to_shared(p); }
__device__ __forceinline__ SmemAddr smem_off(SmemAddr a, int doubles) { return a + 8u * (unsigned)doubles; }
__device__ __forceinline__ void cp_async16(SmemAddr dst, const double* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}
#else
typedef double* SmemAddr;
inline SmemAddr smem_addr(double* p) { return p; }
inline SmemAddr smem_off(SmemAddr a, int doubles) { return a + doubles; }
inline void cp_async16(SmemAddr dst, const double* gmem_src) { dst[0] = gmem_src[0]; dst[1] = gmem_src[1]; }
#endif
__device__ __forceinline__ void cp_async_commit() {
#ifndef AGB_EMULATE
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int NPENDING> __device__ __forceinline__ void cp_async_wait() {
#ifndef AGB_EMULATE
  asm volatile("cp.async.wait_group %0;" ::"n"(NPENDING) : "memory");
#endif
}

// Handle of one array inside the CTA's dynamic shared memory: a 32-bit offset (in doubles) from the window base, so that
// the compiler addresses it as [offset + constant] instead of carrying 64-bit generic pointers through the kernel.
#ifndef AGB_EMULATE
__device__ __forceinline__ double* smem_base() { extern __shared__ __align__(16) double agb_smem_[]; return agb_smem_; }
#else
inline double* smem_base() { return emu::st().smem; }
#endif
struct SP {
  int off;
  __device__ __forceinline__ operator double*() const { return smem_base() + off; }
};
// Same interface for an array that lives in global memory (the "big" instance layout, agb_internal.h).
struct GPh {
  double* p;
  __device__ __forceinline__ operator double*() const { return p; }
};
template <bool BIG> struct SpillSel { typedef SP type; };
template <> struct SpillSel<true> { typedef GPh type; };

// 1-D bulk asynchronous copies (the TMA engine's cp.async.bulk, SASS UBLKCP): one elected thread moves a whole contiguous slab
// global → shared, completion counted in bytes on an mbarrier every thread then waits on; shared → global as a bulk group.
// Slabs must be 16-byte aligned multiples of 16 bytes (callers fall back to element loops otherwise).
#ifndef AGB_EMULATE
__device__ __forceinline__ unsigned smem32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem32(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(void* bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem32(dst)), "l"(src), "r"(bytes), "r"(smem32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) { asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem32(src)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_commit_wait() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#else
inline void mbar_init(void*) {}
inline void mbar_expect(void*, unsigned) {}
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, void*) { memcpy(dst, src, bytes); }
inline void mbar_wait(void*, unsigned) {}
inline void bulk_s2g(void* dst, const void* src, unsigned bytes) { memcpy(dst, src, bytes); }
inline void bulk_commit_wait() {}
inline void fence_async_smem() {}
#endif

// Flags between the warps of one CTA (producer / consumer forward sweep): release-store, acquire-load, and a spin hint
// (the fiber emulator must yield inside a spin loop; the GPU backs off a little so that pollers leave issue slots free).
#ifndef AGB_EMULATE
__device__ __forceinline__ void flag_store(int* p, int v) { asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory"); }
__device__ __forceinline__ int flag_load(const int* p) { int v; asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory"); return v; }
__device__ __forceinline__ void spin_hint(bool relaxed) { if (relaxed) __nanosleep(64); }
#else
inline void flag_store(int* p, int v) { *(volatile int*)p = v; }
inline int flag_load(const int* p) { return *(volatile const int*)p; }
inline void spin_hint(bool) { emu::yield(emu::RUN); }
#endif

// Optional critical-path profile (-DAGB_PHASE_TIMING, profiling builds only): thread 0 accumulates clock64() deltas per
// phase between block barriers; the solve kernel writes the 16 counters over the instance's history log.
#ifdef AGB_PHASE_TIMING
#define AGB_PROF(k) do { const long long t_ = clock64(); prof[k] += t_ - prof_t; prof_t = t_; } while (0)
#else
#define AGB_PROF(k) do { } while (0)
#endif

struct Acc {            // norms of one residual evaluation (statistics.jl:44-57, violations.jl:18-168)
  double sum, opt, dyn, con, sta;
  double psum;          // Σ|row| without the proximal terms (trial evaluations that are kept, see Inst::residual)
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
template <int P, int MODEL, bool BIG = false>
struct Inst {
  static constexpr int n = 4 * P, m = 2 * P, b = P * n + m + n, W = m + n + 1, KUS = m * (n + 1), n1 = n + 1;
  // rows of the gain scratch in global memory are padded to an even length (16-byte row alignment for 128-bit loads)
  static constexpr int n1p = n + 2, KUSP = m * n1p;
  // forward sweep: closed-loop blocks [A − B K | rd − B κ] (n rows of n1p) staged through a shared-memory ring of FWD_D stages
  static constexpr int ACS = n * n1p, FWD_D = fwd_ring_depth(P);
  // per-stage H^x_{i,s} in compact form: dense 2P x 2P position block + diagonal, per player (kkt_solve phases 2/3)
  static constexpr int HmS = 4 * P * P + n;
  static constexpr int HB = 32 + ((P * n + 31) / 32) * 32;      // first thread of the H builders (after GJ warp + row threads)
  static constexpr int OX = 0, OU = P * n, OD = P * n + m;     // offsets inside one stage of R: [rx(p·n) | ru(m) | rd(n)]
  static constexpr int NP = P * (P - 1);                       // ordered player pairs
  static constexpr int kThreads = threads_for(P);
  // Structural non-zeros of the per-player RK2 Jacobians (bit q of the mask <-> At[q] / Bt[q], layouts below):
  //   DoubleIntegrator  A = [I dt·I; 0 I], B = [dt²/2·I; dt·I]   (constants)
  //   Unicycle          At rows x,y only;  B = [Bp (2x2); dt·I]
  //   Bicycle           additionally ∂ψ⁺/∂v and the δ-column / ψ-row of B
  static constexpr unsigned AT_NZ = MODEL == AGB_MODEL_DOUBLE_INTEGRATOR ? 0x09u : (MODEL == AGB_MODEL_UNICYCLE ? 0x0Fu : 0x4Fu);
  static constexpr unsigned BT_NZ = MODEL == AGB_MODEL_DOUBLE_INTEGRATOR ? 0x99u : (MODEL == AGB_MODEL_UNICYCLE ? 0x9Fu : 0xDFu);
  template <int CC> static __device__ __forceinline__ double at_dot(const double At[8], double x0, double x1, double x2, double x3) {
    double v = 0.0;                                            // Σ_q At[q][CC]·x_q over structural non-zeros
    if ((AT_NZ >> (0 + CC)) & 1u) v += At[0 + CC] * x0;
    if ((AT_NZ >> (2 + CC)) & 1u) v += At[2 + CC] * x1;
    if ((AT_NZ >> (4 + CC)) & 1u) v += At[4 + CC] * x2;
    if ((AT_NZ >> (6 + CC)) & 1u) v += At[6 + CC] * x3;
    return v;
  }
  template <int CC> static __device__ __forceinline__ double bt_dot(const double Bt[8], double x0, double x1, double x2, double x3) {
    double v = 0.0;
    if ((BT_NZ >> (0 + CC)) & 1u) v += Bt[0 + CC] * x0;
    if ((BT_NZ >> (2 + CC)) & 1u) v += Bt[2 + CC] * x1;
    if ((BT_NZ >> (4 + CC)) & 1u) v += Bt[4 + CC] * x2;
    if ((BT_NZ >> (6 + CC)) & 1u) v += Bt[6 + CC] * x3;
    return v;
  }
  static __device__ __forceinline__ double at_dot_sel(int cc, const double At[8], double x0, double x1, double x2, double x3) {
    return cc == 0 ? at_dot<0>(At, x0, x1, x2, x3) : at_dot<1>(At, x0, x1, x2, x3);
  }
  static __device__ __forceinline__ double bt_dot_sel(int cc, const double Bt[8], double x0, double x1, double x2, double x3) {
    return cc == 0 ? bt_dot<0>(Bt, x0, x1, x2, x3) : bt_dot<1>(Bt, x0, x1, x2, x3);
  }
  // row r of At / Bt times a 2-vector: At[r][0]·y0 + At[r][1]·y1
  template <int RR> static __device__ __forceinline__ double at_row(const double At[8], double y0, double y1) {
    double v = 0.0;
    if ((AT_NZ >> (2 * RR)) & 1u) v += At[2 * RR] * y0;
    if ((AT_NZ >> (2 * RR + 1)) & 1u) v += At[2 * RR + 1] * y1;
    return v;
  }
  template <int RR> static __device__ __forceinline__ double bt_row(const double Bt[8], double y0, double y1) {
    double v = 0.0;
    if ((BT_NZ >> (2 * RR)) & 1u) v += Bt[2 * RR] * y0;
    if ((BT_NZ >> (2 * RR + 1)) & 1u) v += Bt[2 * RR + 1] * y1;
    return v;
  }

  const DevDesc* __restrict__ d;
  int N, K, nrow, ncw, has_cc, has_pairs, has_self, has_sb, has_cb;
  static constexpr int model = MODEL;
  double dt;
  SP X, U, R, KU, AB, CW, Gp, Gs, Pm, Sv, Ym, Aug, Base, Wm, Hm, xf, Q, Rw, uf, red;
  typename SpillSel<BIG>::type L, CL, CM, Hp, Hs;      // shared memory, or global memory in the big layout
  double* KUg;            // this instance's slice of Buffers::KUg (global)
  double* Rtrial;         // global scratch [S]: un-regularised residual rows of the last trial point (reused if accepted)
  bool keep;              // trial evaluation also produces what the next inner iteration needs (rows, Hessian blocks)
  int pl;                 // iterative best response: the player whose problem is being solved (-1: the full game)
  int tid, lane, warp;
  unsigned bulk_phase;    // parity of the bulk-copy mbarrier (red[47]; the reductions use red[0..23])
#ifdef AGB_PHASE_TIMING
  long long prof_t, prof[16];
#endif

  __device__ void bind(const DevDesc* dd, double* sm) {
    d = dd; N = dd->N; K = dd->K; nrow = dd->nrow; ncw = dd->ncw; has_cc = dd->has_cc; has_sb = dd->has_sb; has_cb = dd->has_cb;
    has_pairs = dd->has_pairs; has_self = dd->has_self; dt = dd->dt;
    (void)sm;
    X.off = dd->o_X; U.off = dd->o_U; R.off = dd->o_R; KU.off = dd->o_KU; AB.off = dd->o_AB;
    CW.off = dd->o_CW; Gp.off = dd->o_Gp; Gs.off = dd->o_Gs;
    if constexpr (!BIG) { L.off = dd->o_L; CL.off = dd->o_CL; CM.off = dd->o_CM; Hp.off = dd->o_Hp; Hs.off = dd->o_Hs; }
    else { L.p = nullptr; CL.p = nullptr; CM.p = nullptr; Hp.p = nullptr; Hs.p = nullptr; }
    Pm.off = dd->o_P; Sv.off = dd->o_Sv; Ym.off = dd->o_Y; Aug.off = dd->o_Aug; Base.off = dd->o_Base;
    Wm.off = dd->o_W; Hm.off = dd->o_Ta; xf.off = dd->o_par; Q.off = xf.off + n; Rw.off = Q.off + n; uf.off = Rw.off + m; red.off = dd->o_red;
    tid = threadIdx.x; lane = tid & 31; warp = tid >> 5; KUg = nullptr; pl = -1; Rtrial = nullptr; keep = false;
    bulk_phase = 0;
    if constexpr (!BIG) { if (tid == 0) mbar_init((double*)red + 47); }      // (the kernels' instance loops start with a block barrier)
#ifdef AGB_PHASE_TIMING
    for (int q = 0; q < 16; q++) prof[q] = 0;
    prof_t = clock64();
#endif
  }
  __device__ void bind_instance(const Buffers& g, int inst) {
    KUg = g.KUg + (size_t)inst * K * KUSP; Rtrial = g.D + (size_t)inst * K * b;
    if constexpr (BIG) {      // the result / multiplier buffers have exactly the shared-memory layouts: work in place
      L.p = g.L + (size_t)inst * P * K * n;
      CL.p = g.conlam + (size_t)inst * K * nrow; CM.p = g.conmu + (size_t)inst * K * nrow;
      Hp.p = g.Hpg + (size_t)inst * g.hpg_stride; Hs.p = Hp.p + N * NP * 3;
    }
  }

  // ---- iterate accessors; TRIAL reads Z + alpha·Δ with Δ held in R (update_traj!, primal_dual_traj.jl:109-128)
  template <bool TRIAL> __device__ __forceinline__ double xg(int k, int a, double alpha) const {
    double v = X[k * n + a];
    if (TRIAL) { if (k > 0) v = fma(alpha, R[(k - 1) * b + OD + a], v); }     // same expression as update_traj
    return v;
  }
  template <bool TRIAL> __device__ __forceinline__ double ug(int s, int idx, double alpha) const {
    double v = U[s * m + idx];
    if (TRIAL) v = fma(alpha, R[s * b + OU + idx], v);
    return v;
  }
  template <bool TRIAL> __device__ __forceinline__ double lg(int i, int s, int a, double alpha) const {
    double v = L[(i * K + s) * n + a];
    if (TRIAL) v = fma(alpha, R[s * b + OX + i * n + a], v);
    return v;
  }

  // Per-player discrete Jacobians of stage s at the last expansion point.  All in-scope models have dynamics that do
  // not depend on position, so A_i = I + At with At non-zero only in columns 2,3: At[r*2 + (c-2)]; Bt[r*2 + j].
  // DoubleIntegrator: constants, never stored.
  __device__ __forceinline__ void loadAB(int s, int i, double At[8], double Bt[8]) const {
    if constexpr (MODEL == AGB_MODEL_DOUBLE_INTEGRATOR) {
#pragma unroll
      for (int q = 0; q < 8; q++) { At[q] = 0.0; Bt[q] = 0.0; }
      At[0 * 2 + 0] = dt; At[1 * 2 + 1] = dt;
      Bt[0 * 2 + 0] = dt * (dt / 2); Bt[1 * 2 + 1] = dt * (dt / 2); Bt[2 * 2 + 0] = dt; Bt[3 * 2 + 1] = dt;
    } else {
      const double2* src = reinterpret_cast<const double2*>(AB + (s * P + i) * 16);
#pragma unroll
      for (int q = 0; q < 4; q++) { double2 v = src[q]; At[2 * q] = v.x; At[2 * q + 1] = v.y; }
#pragma unroll
      for (int q = 0; q < 4; q++) { double2 v = src[4 + q]; Bt[2 * q] = v.x; Bt[2 * q + 1] = v.y; }
    }
  }
  __device__ __forceinline__ double Ael(int s, int i, int r, int c) const {   // A_i[r][c] of stage s (cold paths only)
    double At[8], Bt[8]; loadAB(s, i, At, Bt);
    double v = (r == c) ? 1.0 : 0.0;
    if (c >= 2) {
#pragma unroll
      for (int q = 0; q < 8; q++) if (q == r * 2 + (c - 2)) v += At[q];
    }
    return v;
  }
  __device__ __forceinline__ double Bel(int s, int i, int r, int j) const {   // B_i[r][j] of stage s (cold paths only)
    double At[8], Bt[8]; loadAB(s, i, At, Bt);
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < 8; q++) if (q == r * 2 + j) v = Bt[q];
    return v;
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
  // pre-pass A, per (stage, player): RK2 step and its Jacobian blocks; dynamics rows.
  template <bool TRIAL> __device__ void pass_dyn(double alpha, double* Rout, Acc& acc) {
    for (int item = tid; item < K * P; item += kThreads) {
      int s = item / P, i = item - s * P;
      double st[4], u[2], xn[4];
#pragma unroll
      for (int c = 0; c < 4; c++) st[c] = xg<TRIAL>(s, c * P + i, alpha);
#pragma unroll
      for (int j = 0; j < 2; j++) u[j] = ug<TRIAL>(s, j * P + i, alpha);
      if (model == AGB_MODEL_DOUBLE_INTEGRATOR) {
        rk2_only(model, dt, d->lf, d->lr, st, u, xn);
      } else {
        double A[16], B[8];
        rk2_jac(model, dt, d->lf, d->lr, st, u, xn, A, B);
        double* dst = AB + (s * P + i) * 16;
#pragma unroll
        for (int r = 0; r < 4; r++) {
          dst[r * 2 + 0] = A[r * 4 + 2] - (r == 2 ? 1.0 : 0.0);
          dst[r * 2 + 1] = A[r * 4 + 3] - (r == 3 ? 1.0 : 0.0);
        }
#pragma unroll
        for (int q = 0; q < 8; q++) dst[8 + q] = B[q];
      }
#pragma unroll
      for (int c = 0; c < 4; c++) {
        double r = xn[c] - xg<TRIAL>(s + 1, c * P + i, alpha);      // local_quantities.jl:13
        acc.sum += fabs(r);
        if (pl < 0 || i == pl) acc.dyn = fmax(acc.dyn, fabs(r));
        if (Rout) Rout[s * b + OD + c * P + i] = r;
      }
    }
  }

  // pre-pass B, per (knot k = 1..K, ordered pair (i,j)): collision cost (objective.jl:134-173) and collision-avoidance
  // constraint of player i against j.  Gp = gradient contribution to rows (opt_i, pos_i) (rows (opt_i, pos_j) get −Gp);
  // Hp = symmetric 2x2 block B with H[pos_i,pos_i] = H[pos_j,pos_j] = +B, H[pos_i,pos_j] = H[pos_j,pos_i] = −B.
  template <bool TRIAL> __device__ void pass_pairs(double alpha, Acc& acc) {
    if (!has_pairs) return;
    constexpr int NPd = NP > 0 ? NP : 1, P1 = P > 1 ? P - 1 : 1;     // (single-player games have no pairs)
    for (int item = tid; item < K * NP; item += kThreads) {
      const int pr = item % NPd, k = item / NPd + 1, s = k - 1;
      const int i = pr / P1, jj = pr - i * P1, j = jj < i ? jj : jj + 1;
      const double dtx = (k < K) ? dt : 1.0;                   // terminal knot is not dt-scaled (objective test :52-64)
      const double dx = xg<TRIAL>(k, i, alpha) - xg<TRIAL>(k, j, alpha);
      const double dy = xg<TRIAL>(k, P + i, alpha) - xg<TRIAL>(k, P + j, alpha);
      const double d2 = dx * dx + dy * dy;
      double gx = 0.0, gy = 0.0, h00 = 0.0, h01 = 0.0, h11 = 0.0;
      if (has_cc) {
        const double dn = sqrt(d2);
        const double rr = d->cc_radius[i], mu = d->cc_mu[i];
        if (fmax(0.0, rr - dn) > 0.0) {
          const double eps = 1e-10, eps_norm = eps * sqrt((double)n);
          const double q = rr / (eps_norm + dn);
          gx = -dtx * (mu * (q * (eps + dx) - dx));            // q[pxi] = −g
          gy = -dtx * (mu * (q * (eps + dy) - dy));
          if (!TRIAL || keep) {
            const double idn = 1.0 / dn, t = rr * idn, t3 = t * idn * idn;
            h00 = dtx * mu * (1.0 - t + t3 * dx * dx);
            h01 = dtx * mu * (t3 * dx * dy);
            h11 = dtx * mu * (1.0 - t + t3 * dy * dy);
          }
        }
      }
      const int row = d->col_row[i][j];
      if (row >= 0) {                                          // CollisionConstraint: c = r² − ‖xi−xj‖², ∇c[pos_i] = −2d
        const double rad = d->col_radius[i][j];
        const double cv = rad * rad - d2;
        double w; const double g = al_row(s, row, cv, w);
        gx -= 2.0 * dx * g; gy -= 2.0 * dy * g;
        if (pl < 0 || i == pl) acc.sta = fmax(acc.sta, cv);
        if (!TRIAL || keep) { const double w4 = 4.0 * w; h00 += w4 * dx * dx; h01 += w4 * dx * dy; h11 += w4 * dy * dy; }
      }
      double* gp = Gp + (k * NP + pr) * 2; gp[0] = gx; gp[1] = gy;
      if (!TRIAL || keep) { double* hp = Hp + (k * NP + pr) * 3; hp[0] = h00; hp[1] = h01; hp[2] = h11; }
    }
  }

  // pre-pass C, per (knot, player): walls and circles acting on the player's own position.
  template <bool TRIAL> __device__ void pass_self(double alpha, Acc& acc) {
    if (!has_self) return;
    for (int item = tid; item < K * P; item += kThreads) {
      const int i = item % P, k = item / P + 1, s = k - 1;
      const double px = xg<TRIAL>(k, i, alpha), py = xg<TRIAL>(k, P + i, alpha);
      double gx = 0.0, gy = 0.0, h00 = 0.0, h01 = 0.0, h11 = 0.0;
      const int nw = d->n_walls[i];
      for (int q = 0; q < nw; q++) {                           // WallConstraint (constraints/wall_constraint.jl:56-89)
        const double* wl = d->walls[i][q];
        const double x1 = wl[0], y1 = wl[1], x2 = wl[2], y2 = wl[3], xv = wl[4], yv = wl[5];
        const bool left = (px - x1) * (x2 - x1) + (py - y1) * (y2 - y1) > 0.0;
        const bool right = (px - x2) * (x1 - x2) + (py - y2) * (y1 - y2) > 0.0;
        const double msk = (left && right) ? 1.0 : 0.0;
        const double cv = ((px - x1) * xv + (py - y1) * yv) * msk;
        double w; const double g = al_row(s, d->wall_row[i] + q, cv, w);
        gx += msk * xv * g; gy += msk * yv * g;
        if (pl < 0 || i == pl) acc.sta = fmax(acc.sta, cv);
        const double wm = w * msk;
        h00 += wm * xv * xv; h01 += wm * xv * yv; h11 += wm * yv * yv;
      }
      const int nc = d->n_circles[i];
      for (int q = 0; q < nc; q++) {                           // CircleConstraint: c = r² − (x−xc)² − (y−yc)²
        const double* cl = d->circles[i][q];
        const double ex = px - cl[0], ey = py - cl[1];
        const double cv = cl[2] * cl[2] - ex * ex - ey * ey;
        double w; const double g = al_row(s, d->circle_row[i] + q, cv, w);
        gx -= 2.0 * ex * g; gy -= 2.0 * ey * g;
        if (pl < 0 || i == pl) acc.sta = fmax(acc.sta, cv);
        const double w4 = 4.0 * w;
        h00 += w4 * ex * ex; h01 += w4 * ex * ey; h11 += w4 * ey * ey;
      }
      double* gs = Gs + (k * P + i) * 2; gs[0] = gx; gs[1] = gy;
      if (!TRIAL || keep) { double* hs = Hs + (k * P + i) * 3; hs[0] = h00; hs[1] = h01; hs[2] = h11; }
    }
  }

  // one element of player i's stationarity row block w.r.t. x at knot k (1..K), joint comp a
  // HALF: 0 = the caller guarantees a position component (c < 2), 1 = a velocity / heading component (c >= 2), -1 = unknown
  template <bool TRIAL, int HALF = -1> __device__ double xrow_elem(int i, int k, int a, double alpha, double reg_x, Acc& acc, double& plain) {
    const int c = a / P, ia = a - c * P, s = k - 1;
    const bool is_pos = HALF < 0 ? (c < 2) : (HALF == 0);
    const double xa = xg<TRIAL>(k, a, alpha);
    double v = 0.0;
    if (ia == i) v = ((k < K) ? dt : 1.0) * Q[a] * (xa - xf[a]);      // LQR gradient (objective.jl:24-32)
    if (is_pos) {
      if (has_pairs) {
        if (ia == i) {
#pragma unroll
          for (int jj = 0; jj < P - 1; jj++) v += Gp[(k * NP + i * (P - 1) + jj) * 2 + c];
        } else {
          v -= Gp[(k * NP + pair_index(i, ia)) * 2 + c];
        }
      }
      if (has_self && ia == i) v += Gs[(k * P + i) * 2 + c];
    }
    if (has_sb) {                                                     // StateBoundConstraint (state_bound_constraint.jl:80-92)
      int row = d->sbmax_row[i][a];
      if (row >= 0) {
        const double cv = xa - d->x_max[i][a];
        double w; const double g = al_row(s, row, cv, w);
        v += g; acc.sta = fmax(acc.sta, cv); CW[s * ncw + row - d->sb_shift[i]] = w;
      }
      row = d->sbmin_row[i][a];
      if (row >= 0) {
        const double cv = d->x_min[i][a] - xa;
        double w; const double g = al_row(s, row, cv, w);
        v -= g; acc.sta = fmax(acc.sta, cv); CW[s * ncw + row - d->sb_shift[i]] = w;
      }
    }
    if (k < K) {                                                      // + A_kᵀ λ_{i,k}   (global_quantities.jl:45-53)
      v += lg<TRIAL>(i, k, a, alpha);
      if (!is_pos) {
        double At[8], Bt[8]; loadAB(k, ia, At, Bt);
        v += at_dot_sel(c - 2, At, lg<TRIAL>(i, k, ia, alpha), lg<TRIAL>(i, k, P + ia, alpha),
                        lg<TRIAL>(i, k, 2 * P + ia, alpha), lg<TRIAL>(i, k, 3 * P + ia, alpha));
      }
    }
    v -= lg<TRIAL>(i, k - 1, a, alpha);                               // − λ_{i,k−1}
    plain = v;
    if (TRIAL) v += reg_x * (alpha * R[s * b + OD + a]);              // regularize_residual! (:67-86)
    return v;
  }

  // one element of player i's stationarity row block w.r.t. u_{i,s}: own control comp j
  template <bool TRIAL> __device__ double urow_elem(int i, int s, int j, double alpha, double reg_u, Acc& acc, double& plain) {
    const int idx = j * P + i;
    const double ua = ug<TRIAL>(s, idx, alpha);
    double v = dt * Rw[idx] * (ua - uf[idx]);
    double At[8], Bt[8]; loadAB(s, i, At, Bt);
    v += bt_dot_sel(j, Bt, lg<TRIAL>(i, s, i, alpha), lg<TRIAL>(i, s, P + i, alpha), lg<TRIAL>(i, s, 2 * P + i, alpha),
                    lg<TRIAL>(i, s, 3 * P + i, alpha));
    if (has_cb) {                                                     // ControlBoundConstraint (control_bound_constraint.jl:94-106)
      int row = d->ub_row[idx];
      if (row >= 0) {
        const double cv = ua - d->u_max[idx];
        double w; const double g = al_row(s, row, cv, w);
        v += g; if (pl < 0) acc.con = fmax(acc.con, cv); CW[s * ncw + row - d->cb_shift] = w;
      }
      row = d->lb_row[idx];
      if (row >= 0) {
        const double cv = d->u_min[idx] - ua;
        double w; const double g = al_row(s, row, cv, w);
        v -= g; if (pl < 0) acc.con = fmax(acc.con, cv); CW[s * ncw + row - d->cb_shift] = w;
      }
    }
    plain = v;
    if (TRIAL) v += reg_u * (alpha * R[s * b + OU + idx]);
    return v;
  }

  __device__ Acc block_reduce(Acc a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a.sum += __shfl_xor_sync(AGB_FULL, a.sum, o);
      a.psum += __shfl_xor_sync(AGB_FULL, a.psum, o);
      a.opt = fmax(a.opt, __shfl_xor_sync(AGB_FULL, a.opt, o));
      a.dyn = fmax(a.dyn, __shfl_xor_sync(AGB_FULL, a.dyn, o));
      a.con = fmax(a.con, __shfl_xor_sync(AGB_FULL, a.con, o));
      a.sta = fmax(a.sta, __shfl_xor_sync(AGB_FULL, a.sta, o));
    }
    __syncthreads();            // red[] may still be read from the previous reduction
    if (lane == 0) { double* r = red + warp * 6; r[0] = a.sum; r[1] = a.opt; r[2] = a.dyn; r[3] = a.con; r[4] = a.sta; r[5] = a.psum; }
    __syncthreads();
    Acc t = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) {
      const double* r = red + w * 6;
      t.sum += r[0]; t.opt = fmax(t.opt, r[1]); t.dyn = fmax(t.dyn, r[2]); t.con = fmax(t.con, r[3]); t.sta = fmax(t.sta, r[4]); t.psum += r[5];
    }
    return t;
  }

  // rows of the system being solved: the full game, or player pl's best-response problem (newton_core.jl:205-245)
  __device__ __forceinline__ double res_size() const { return pl < 0 ? (double)(K * b) : (double)(K * (2 * n + 2)); }

  // Full residual evaluation.  !TRIAL: at Z, rows stored to Rout (= R).  TRIAL: at Z + alpha·Δ (Δ in R) with the proximal
  // terms of regularize_residual!; rows go to Rout if non-null (may be global memory), norms are always returned.
  // NaN-safe maxima: fmax drops NaNs, so a non-finite residual is caught through `sum`.
  // With `plain` (trial evaluations of the line search): Rout receives the UN-regularised rows and *plain their norms, so
  // that an accepted trial point needs no re-evaluation at the next inner iteration.
  template <bool TRIAL> __device__ Acc residual(double alpha, double reg_x, double reg_u, double* Rout, Acc* plain = nullptr) {
    Acc acc = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    keep = TRIAL && plain != nullptr;
    pass_dyn<TRIAL>(alpha, Rout, acc);
    acc.psum = acc.sum;               // dynamics rows carry no proximal term
    pass_pairs<TRIAL>(alpha, acc);
    pass_self<TRIAL>(alpha, acc);
    __syncthreads();
    // x rows in two passes — position components (pair / wall / circle terms), then velocity / heading components (Aᵀλ
    // terms) — so that the lanes of a warp follow one path; stage-major item order, compile-time divisors only
    const int nxh = P * K * 2 * P;
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
      for (int item = tid; item < nxh; item += kThreads) {
        const int s = item / (P * 2 * P), rem = item - s * (P * 2 * P);
        const int i = rem / (2 * P), a = half * 2 * P + (rem - i * (2 * P));
        if (pl >= 0 && i != pl) { if (Rout) Rout[s * b + OX + i * n + a] = 0.0; continue; }    // IBR: rows of player pl only
        double pv;
        double v = half == 0 ? xrow_elem<TRIAL, 0>(i, s + 1, a, alpha, reg_x, acc, pv) : xrow_elem<TRIAL, 1>(i, s + 1, a, alpha, reg_x, acc, pv);
        acc.sum += fabs(v);
        acc.psum += fabs(pv);
        if (keep) v = pv;
        acc.opt = fmax(acc.opt, fabs(v));
        if (Rout) Rout[s * b + OX + i * n + a] = v;
      }
    }
    const int nu = K * m;
    for (int item = tid; item < nu; item += kThreads) {
      int idx = item % m, s = item / m;
      int j = idx / P, i = idx - j * P;
      if (pl >= 0 && i != pl) { if (Rout) Rout[s * b + OU + idx] = 0.0; continue; }
      double pv;
      double v = urow_elem<TRIAL>(i, s, j, alpha, reg_u, acc, pv);
      acc.sum += fabs(v);
      acc.psum += fabs(pv);
      if (keep) v = pv;
      acc.opt = fmax(acc.opt, fabs(v));
      if (Rout) Rout[s * b + OU + idx] = v;
    }
    if (pl >= 0 && has_cb && (!TRIAL || keep)) {
      // control_violation(game_con, pdtraj, i) (violations.jl:69-82) indexes the stacked bound rows with pu[i]: rows
      // pu[i] of the control-bound conval, whatever bound they belong to — restated as is (at the trial point too when
      // its record is kept for the next inner iteration)
      for (int item = tid; item < K * 2; item += kThreads) {
        const int j = item & 1, s = item >> 1, local = j * P + pl, row = d->nrow_state + local;
        if (local < d->nrow_control) {
          for (int idx = 0; idx < m; idx++) {
            if (d->ub_row[idx] == row) acc.con = fmax(acc.con, ug<TRIAL>(s, idx, alpha) - d->u_max[idx]);
            if (d->lb_row[idx] == row) acc.con = fmax(acc.con, d->u_min[idx] - ug<TRIAL>(s, idx, alpha));
          }
        }
      }
    }
    Acc t = block_reduce(acc);
    if (plain) { *plain = t; plain->sum = t.psum; }
    keep = false;
    return t;
  }

  // ---------------------------------------------------------------------------------------------------------
  // residual_jacobian! blocks (global_quantities.jl:109-193) at the last expansion point (uses Hp, Hs, CW, AB)
  // ---------------------------------------------------------------------------------------------------------
  // H^x_{i,k}[a][bq] for position comps a = (ca,ia), bq = (cb,ib), ca,cb < 2
  __device__ __forceinline__ double hpos_entry(int i, int k, int ca, int ia, int cb, int ib) const {
    const int e = ca + cb;
    double v = 0.0;
    if (ia == ib) {
      if (ia == i) {
        if (has_pairs) {
#pragma unroll
          for (int jj = 0; jj < P - 1; jj++) v += Hp[(k * NP + i * (P - 1) + jj) * 3 + e];
        }
        if (has_self) v += Hs[(k * P + i) * 3 + e];
      } else if (has_pairs) {
        v = Hp[(k * NP + pair_index(i, ia)) * 3 + e];
      }
    } else if (has_pairs) {
      if (ia == i) v = -Hp[(k * NP + pair_index(i, ib)) * 3 + e];
      else if (ib == i) v = -Hp[(k * NP + pair_index(i, ia)) * 3 + e];
    }
    return v;
  }
  // diagonal part of H^x_{i,k}: dt·Q_i + state-bound terms + reg.x
  __device__ __forceinline__ double hd_entry(int i, int k, int a, double reg_x) const {
    const int ia = a % P, s = k - 1;
    double v = reg_x;
    if (ia == i) v += ((k < K) ? dt : 1.0) * Q[a];
    if (has_sb) {
      int row = d->sbmax_row[i][a]; if (row >= 0) v += CW[s * ncw + row - d->sb_shift[i]];
      row = d->sbmin_row[i][a];     if (row >= 0) v += CW[s * ncw + row - d->sb_shift[i]];
    }
    return v;
  }
  __device__ __forceinline__ double h_entry(int i, int k, int a, int bq, double reg_x) const {
    double v = (a == bq) ? hd_entry(i, k, a, reg_x) : 0.0;
    if (a < 2 * P && bq < 2 * P) v += hpos_entry(i, k, a / P, a % P, bq / P, bq % P);
    return v;
  }
  // H^u diagonal of stage s, joint control comp idx: dt·R + control-bound terms + reg.u
  __device__ __forceinline__ double hu_entry(int s, int idx, double reg_u) const {
    double v = dt * Rw[idx] + reg_u;
    if (has_cb) {
      int row = d->ub_row[idx]; if (row >= 0) v += CW[s * ncw + row - d->cb_shift];
      row = d->lb_row[idx];     if (row >= 0) v += CW[s * ncw + row - d->cb_shift];
    }
    return v;
  }

  // 1/x to (almost) full double precision without the slow-path branches of a correctly rounded division:
  // rcp.approx.ftz.f64 (≈ 20 bits) + two Newton steps.  x = 0 / inf / NaN propagate as inf / 0 / NaN like 1/x would.
  static __device__ __forceinline__ double fast_rcp(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0); r = fma(r, e, r);
    e = fma(-x, r, 1.0); r = fma(r, e, r);
    return r;
#else
    return 1.0 / x;
#endif
  }

  // |x| as an ordered integer: the high word of a non-negative double is monotone in its value, so magnitude tests run on
  // the integer pipe — the FP64 pipe of the scheduler that hosts this warp is the scarce resource of the whole solve
  // (measured: every variant of this routine that issued more FP64 instructions was slower, whatever its dependency chain).
  static __device__ __forceinline__ int abs_hi(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x) & 0x7fffffff;
#else
    long long u; memcpy(&u, &x, 8); return (int)((u >> 32) & 0x7fffffff);
#endif
  }

  static __device__ __forceinline__ int imax(int x, int y) { return x > y ? x : y; }

  // Gauss-Jordan on the m x W augmented gain system, one column per lane (warp 0).  Threshold pivoting like the
  // reference's UMFPACK: the diagonal pivot is kept while |a_tt| >= 2^-7·max_r |a_rt| (no row exchange, no index
  // shuffles); only if some step violates the threshold is the system re-solved from shared memory with full partial
  // pivoting.  Returns false on a zero / non-finite pivot.
  //
  // Default form: 2 x 2 BLOCK pivots (m = 2P is even): the two pivot columns of a block step travel through shared memory in one
  // round trip and the block [p00 p01; p10 p11] is inverted explicitly (one determinant, one reciprocal), so the m-pivot chain of
  // [store column -> load -> reciprocal -> scale -> eliminate] becomes m/2 links of barely longer length, with fewer instructions
  // on this warp (the stage loop's critical path).  The threshold test bounds the same quantity as the scalar form,
  // the size of the elimination multipliers: max|inverse entry| * max|entry below the block| <= 2^7, on the integer pipe; a zero /
  // non-finite determinant or a violated bound sends the system to the partial-pivoting fallback.  -DAGB_GJ_SCALAR keeps the
  // scalar-pivot form (A/B measurements).
  __device__ bool gj_warp(double* aug, double* kg = nullptr) {
    double a[m];
    const bool act = lane < W;
#pragma unroll
    for (int r = 0; r < m; r++) a[r] = act ? aug[r * W + lane] : 0.0;
    bool weak = false;                                   // warp-uniform: evaluated on the broadcast pivot column
#ifndef AGB_GJ_SCALAR
#pragma unroll
    for (int t = 0; t < m; t += 2) {
      double2* piv = reinterpret_cast<double2*>(Ym + t * m);                // columns t, t+1 (Ym is free between phases 1 and 3)
      if (lane == t || lane == t + 1) {
        double2* dst = piv + (lane - t) * (m / 2);
#pragma unroll
        for (int r = 0; r < m; r += 2) dst[r / 2] = make_double2(a[r], a[r + 1]);
      }
      __syncwarp();
      double f0[m], f1[m];
#pragma unroll
      for (int r = 0; r < m; r += 2) {
        const double2 v = piv[r / 2], w = piv[m / 2 + r / 2];
        f0[r] = v.x; f0[r + 1] = v.y; f1[r] = w.x; f1[r + 1] = w.y;
      }
      const double p00 = f0[t], p10 = f0[t + 1], p01 = f1[t], p11 = f1[t + 1];
      const double det = fma(p00, p11, -(p01 * p10));
      const int hd = abs_hi(det);
      if ((unsigned)(hd - 0x00100000) >= 0x7fe00000u) weak = true;          // zero / subnormal / inf / NaN determinant
      const double rd = fast_rcp(det);
      const double i00 = p11 * rd, i01 = -(p01 * rd), i10 = -(p10 * rd), i11 = p00 * rd;
      if (t + 2 < m) {                                                       // multipliers of the rows below the block stay <= 2^7
        int hi_i = imax(imax(abs_hi(i00), abs_hi(i01)), imax(abs_hi(i10), abs_hi(i11)));
        int hi_f = 0;
#pragma unroll
        for (int r = t + 2; r < m; r++) hi_f = imax(hi_f, imax(abs_hi(f0[r]), abs_hi(f1[r])));
        if (hi_f > 0 && hi_i + hi_f - (1023 << 20) > ((1023 + 7) << 20)) weak = true;
      }
      const double at0 = fma(i01, a[t + 1], i00 * a[t]), at1 = fma(i11, a[t + 1], i10 * a[t]);
#pragma unroll
      for (int r = 0; r < m; r++) if (r != t && r != t + 1) a[r] = fma(-f1[r], at1, fma(-f0[r], at0, a[r]));
      a[t] = at0; a[t + 1] = at1;
    }
#else
#pragma unroll
    for (int t = 0; t < m; t++) {
      // pivot column t → every lane, through shared memory (Ym is free between phases 1 and 3; one slot per step)
      double2* piv = reinterpret_cast<double2*>(Ym + t * m);
      if (lane == t) {
#pragma unroll
        for (int r = 0; r < m; r += 2) piv[r / 2] = make_double2(a[r], a[r + 1]);
      }
      __syncwarp();
      double f[m];
#pragma unroll
      for (int r = 0; r < m; r += 2) { const double2 v = piv[r / 2]; f[r] = v.x; f[r + 1] = v.y; }
      const double pv = f[t];
      f[t] = 0.0;
      const int hp = abs_hi(pv);
      if ((unsigned)(hp - 0x00100000) >= 0x7fe00000u) weak = true;          // zero / subnormal / inf / NaN pivot
      const int lim = hp + (7 << 20);                                       // 2^7·|pv|
#pragma unroll
      for (int r = t + 1; r < m; r++) if (abs_hi(f[r]) > lim) weak = true;
      const double at = a[t] * fast_rcp(pv);
#pragma unroll
      for (int r = 0; r < m; r++) if (r != t) a[r] = fma(-f[r], at, a[r]);
      a[t] = at;
    }
#endif
    if (!weak) {
      if (act) {
#pragma unroll
        for (int r = 0; r < m; r++) aug[r * W + lane] = a[r];
        if (kg != nullptr && lane >= m) {                                   // K | κ → the L2-resident gain scratch, from registers
#pragma unroll
          for (int r = 0; r < m; r++) kg[r * n1p + (lane - m)] = a[r];
        }
      }
      return true;
    }
    const bool ok = gj_warp_pivoted(aug);
    if (kg != nullptr && lane >= m && act) {
#pragma unroll
      for (int r = 0; r < m; r++) kg[r * n1p + (lane - m)] = aug[r * W + lane];
    }
    return ok;
  }

  // full partial pivoting (row exchanges), used when the threshold test above fails
  __device__ bool gj_warp_pivoted(double* aug) {
#ifdef AGB_DEBUG_COUNT_FALLBACK
    { static long cnt = 0; if (lane == 0) { cnt++; if ((cnt & (cnt - 1)) == 0) fprintf(stderr, "gj_warp_pivoted call #%ld\n", cnt); } }
#endif
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
      double f[m];
#pragma unroll
      for (int r = 0; r < m; r++) f[r] = (r != t) ? __shfl_sync(AGB_FULL, a[r], t) : 0.0;
      if (!(fabs(pv) > 0.0) || isinf(pv)) ok = false;
      const double at = a[t] * (1.0 / pv);
#pragma unroll
      for (int r = 0; r < m; r++) if (r != t) a[r] = fma(-f[r], at, a[r]);
      a[t] = at;
    }
    if (act) {
#pragma unroll
      for (int r = 0; r < m; r++) aug[r * W + lane] = a[r];
    }
    return ok;
  }

  // Y_i = B_iᵀ P_i (+ affine column B_iᵀ s_i) for stage s, from the P rows of player i's own states
  __device__ void compute_Y(int s) {
    for (int item = tid; item < P * n1; item += kThreads) {
      const int col = item % n1, i = item / n1;
      double At[8], Bt[8]; loadAB(s, i, At, Bt);
      double pv[4];
#pragma unroll
      for (int q = 0; q < 4; q++) pv[q] = (col < n) ? Pm[(i * n + q * P + i) * n + col] : Sv[i * n + q * P + i];
      Ym[(0 * P + i) * n1 + col] = bt_dot<0>(Bt, pv[0], pv[1], pv[2], pv[3]);
      Ym[(1 * P + i) * n1 + col] = bt_dot<1>(Bt, pv[0], pv[1], pv[2], pv[3]);
    }
  }

  // ---------------------------------------------------------------------------------------------------------
  // Iterative best response (solver_methods.jl:226-265, global_quantities.jl:288-365): Newton step of player pl's own
  // optimal-control problem, Δtraj[horiz_mask] = −(lu(jac[verti_mask, horiz_mask]) \ res[verti_mask]).  Unknowns x,
  // u_pl, λ_pl; the other players' controls and multipliers do not move.  Same stage-wise factorisation as kkt_solve
  // with one value matrix P (n x n) and a 2 x 2 gain system per stage; R holds the masked residual on entry (rows of
  // the other players zero) and the step on exit.
  // ---------------------------------------------------------------------------------------------------------
  __device__ bool kkt_solve_ibr(double reg_x, double reg_u) {
    const int i = pl;
    int ok = 1;
    for (int item = tid; item < m * n1; item += kThreads) KU[item] = 0.0;       // gain rows of the other players stay zero
    for (int item = tid; item < n * n; item += kThreads) Pm[item] = h_entry(i, K, item / n, item % n, reg_x);
    for (int a = tid; a < n; a += kThreads) Sv[a] = R[(K - 1) * b + OX + i * n + a];
    __syncthreads();
    for (int col = tid; col < n1; col += kThreads) {
      double At[8], Bt[8]; loadAB(K - 1, i, At, Bt);
      double pv[4];
#pragma unroll
      for (int q = 0; q < 4; q++) pv[q] = (col < n) ? Pm[(q * P + i) * n + col] : Sv[q * P + i];
      Ym[col] = bt_dot<0>(Bt, pv[0], pv[1], pv[2], pv[3]);
      Ym[n1 + col] = bt_dot<1>(Bt, pv[0], pv[1], pv[2], pv[3]);
    }
    __syncthreads();
    for (int s = K - 1; s >= 0; s--) {
      const double* Rs = R + s * b;
      constexpr int NK = n1, NB = P * n, NW = P * 2, NA = P;
      // ---- phase A: gains (one column per thread, 2 x 2 solve with partial pivoting) and the K-independent parts
      for (int item = tid; item < NK + NB + NW + NA; item += kThreads) {
        if (item < NK) {
          const int col = item;
          double At[8], Bt[8]; loadAB(s, i, At, Bt);
          const double* y0 = Ym; const double* y1 = Ym + n1;
          double s00 = bt_dot<0>(Bt, y0[i], y0[P + i], y0[2 * P + i], y0[3 * P + i]) + hu_entry(s, i, reg_u);
          double s01 = bt_dot<1>(Bt, y0[i], y0[P + i], y0[2 * P + i], y0[3 * P + i]);
          double s10 = bt_dot<0>(Bt, y1[i], y1[P + i], y1[2 * P + i], y1[3 * P + i]);
          double s11 = bt_dot<1>(Bt, y1[i], y1[P + i], y1[2 * P + i], y1[3 * P + i]) + hu_entry(s, P + i, reg_u);
          double r0, r1;
          if (col < n) {
            const int c2 = col / P, i2 = col - c2 * P;
            r0 = y0[col]; r1 = y1[col];
            if (c2 >= 2) {
              double At2[8], Bt2[8]; loadAB(s, i2, At2, Bt2);
              r0 += at_dot_sel(c2 - 2, At2, y0[i2], y0[P + i2], y0[2 * P + i2], y0[3 * P + i2]);
              r1 += at_dot_sel(c2 - 2, At2, y1[i2], y1[P + i2], y1[2 * P + i2], y1[3 * P + i2]);
            }
          } else {
            r0 = y0[n] + Rs[OU + i]; r1 = y1[n] + Rs[OU + P + i];
            for (int a2 = 0; a2 < n; a2++) { r0 += y0[a2] * Rs[OD + a2]; r1 += y1[a2] * Rs[OD + a2]; }
          }
          if (fabs(s10) > fabs(s00)) { double t = s00; s00 = s10; s10 = t; t = s01; s01 = s11; s11 = t; t = r0; r0 = r1; r1 = t; }
          if (!(fabs(s00) > 0.0) || isinf(s00)) ok = 0;
          const double f = s10 / s00, d11 = s11 - f * s01;
          if (!(fabs(d11) > 0.0) || isinf(d11)) ok = 0;
          const double k1 = (r1 - f * r0) / d11, k0 = (r0 - s01 * k1) / s00;
          KU[i * n1 + col] = k0; KU[(P + i) * n1 + col] = k1;
        } else if (s > 0) {
          const int it = item - NK;
          int kind, col, i2;                               // 0: Base column, 1: W column, 2: affine
          if (it < NB) { kind = 0; col = it % n; i2 = it / n; }
          else if (it < NB + NW) { kind = 1; col = (it - NB) & 1; i2 = (it - NB) >> 1; }
          else { kind = 2; col = n; i2 = it - NB - NW; }
          const double* p0 = Pm + (0 * P + i2) * n;
          const double* p1 = Pm + (1 * P + i2) * n;
          const double* p2 = Pm + (2 * P + i2) * n;
          const double* p3 = Pm + (3 * P + i2) * n;
          double T0, T1, T2, T3;
          if (kind == 1) {                                  // (P B_pl)[(q,i2)][j]
            double At[8], Bt[8]; loadAB(s, i, At, Bt);
            T0 = bt_dot_sel(col, Bt, p0[i], p0[P + i], p0[2 * P + i], p0[3 * P + i]);
            T1 = bt_dot_sel(col, Bt, p1[i], p1[P + i], p1[2 * P + i], p1[3 * P + i]);
            T2 = bt_dot_sel(col, Bt, p2[i], p2[P + i], p2[2 * P + i], p2[3 * P + i]);
            T3 = bt_dot_sel(col, Bt, p3[i], p3[P + i], p3[2 * P + i], p3[3 * P + i]);
          } else if (kind == 0) {                           // (P A)[(q,i2)][col]
            const int c2 = col / P, i3 = col - c2 * P;
            T0 = p0[col]; T1 = p1[col]; T2 = p2[col]; T3 = p3[col];
            if (c2 >= 2) {
              double At[8], Bt[8]; loadAB(s, i3, At, Bt);
              T0 += at_dot_sel(c2 - 2, At, p0[i3], p0[P + i3], p0[2 * P + i3], p0[3 * P + i3]);
              T1 += at_dot_sel(c2 - 2, At, p1[i3], p1[P + i3], p1[2 * P + i3], p1[3 * P + i3]);
              T2 += at_dot_sel(c2 - 2, At, p2[i3], p2[P + i3], p2[2 * P + i3], p2[3 * P + i3]);
              T3 += at_dot_sel(c2 - 2, At, p3[i3], p3[P + i3], p3[2 * P + i3], p3[3 * P + i3]);
            }
          } else {                                          // P rd + s
            T0 = Sv[i2]; T1 = Sv[P + i2]; T2 = Sv[2 * P + i2]; T3 = Sv[3 * P + i2];
            for (int a2 = 0; a2 < n; a2++) {
              const double e = Rs[OD + a2];
              T0 += p0[a2] * e; T1 += p1[a2] * e; T2 += p2[a2] * e; T3 += p3[a2] * e;
            }
          }
          double At2[8], Bt2[8]; loadAB(s, i2, At2, Bt2);
          const double o[4] = {T0, T1, T2 + at_dot<0>(At2, T0, T1, T2, T3), T3 + at_dot<1>(At2, T0, T1, T2, T3)};
          if (kind == 1) {
#pragma unroll
            for (int c = 0; c < 4; c++) Wm[(c * P + i2) * 2 + col] = o[c];
          } else if (kind == 0) {
#pragma unroll
            for (int c = 0; c < 4; c++) Base[(c * P + i2) * n1 + col] = o[c] + h_entry(i, s, c * P + i2, col, reg_x);
          } else {
#pragma unroll
            for (int c = 0; c < 4; c++) Base[(c * P + i2) * n1 + n] = o[c] + R[(s - 1) * b + OX + i * n + c * P + i2];
          }
        }
      }
      __syncthreads();
      for (int item = tid; item < m * n1; item += kThreads) KUg[s * KUSP + (item / n1) * n1p + item % n1] = KU[item];
      if (s == 0) break;
      // ---- phase B: P ← Base − W K,  s ← base − W κ,  Y ← B_{s−1}ᵀ P
      for (int item = tid; item < P * n1; item += kThreads) {
        const int col = item % n1, i2 = item / n1;
        const double k0 = KU[i * n1 + col], k1 = KU[(P + i) * n1 + col];
        double pn[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const int a = c * P + i2;
          pn[c] = Base[a * n1 + col] - Wm[a * 2] * k0 - Wm[a * 2 + 1] * k1;
        }
        if (i2 == i) {
          double At[8], Bt[8]; loadAB(s - 1, i, At, Bt);
          Ym[col] = bt_dot<0>(Bt, pn[0], pn[1], pn[2], pn[3]);
          Ym[n1 + col] = bt_dot<1>(Bt, pn[0], pn[1], pn[2], pn[3]);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) { const int a = c * P + i2; if (col < n) Pm[a * n + col] = pn[c]; else Sv[a] = pn[c]; }
      }
      __syncthreads();
    }
    __syncthreads();
    ok = __syncthreads_and(ok);
    forward_and_costate(reg_x);
    return ok != 0;
  }

  // H^x_{i,s} of every player in compact dense form (position block [2P][2P], then the diagonal), one thread per
  // 2x2 position block (i; i2, ib); laid out one stage ahead of its use by spare warps of the factorisation
  __device__ void build_hm(double* hmb, int s, double reg_x, int t0, int stride) {
    for (int blk = t0; blk >= 0 && blk < P * P * P; blk += stride) {
      const int i = blk / (P * P), rem = blk - i * P * P, i2 = rem / P, ib = rem - i2 * P;
      double h0 = 0.0, h1 = 0.0, h2 = 0.0;
      if (i2 == ib) {
        if (i2 == i) {
          if (has_pairs) {
#pragma unroll
            for (int jj = 0; jj < P - 1; jj++) { const double* hp = Hp + (s * NP + i * (P - 1) + jj) * 3; h0 += hp[0]; h1 += hp[1]; h2 += hp[2]; }
          }
          if (has_self) { const double* hs = Hs + (s * P + i) * 3; h0 += hs[0]; h1 += hs[1]; h2 += hs[2]; }
        } else if (has_pairs) {
          const double* hp = Hp + (s * NP + pair_index(i, i2)) * 3; h0 = hp[0]; h1 = hp[1]; h2 = hp[2];
        }
      } else if (has_pairs && (i2 == i || ib == i)) {
        const double* hp = Hp + (s * NP + pair_index(i, i2 == i ? ib : i2)) * 3; h0 = -hp[0]; h1 = -hp[1]; h2 = -hp[2];
      }
      double* hm = hmb + i * HmS;
      hm[i2 * (2 * P) + ib] = h0; hm[i2 * (2 * P) + P + ib] = h1;
      hm[(P + i2) * (2 * P) + ib] = h1; hm[(P + i2) * (2 * P) + P + ib] = h2;
      if (i2 == ib) {
#pragma unroll
        for (int c = 0; c < 4; c++) hm[4 * P * P + c * P + i2] = hd_entry(i, s, c * P + i2, reg_x);
      }
    }
  }

  // ---------------------------------------------------------------------------------------------------------
  // Δtraj = −(lu(jac) \ res)  (solver_methods.jl:87-88): R holds res on entry, Δ on exit (same stage-major slots:
  // rx(i,s) → Δλ_{i,s}, ru(s) → Δu_s, rd(s) → Δx_{s+1}).  Returns false on a singular / non-finite pivot.
  //
  // Per stage s (backward), with P_i, s_i of knot s+1 and Y = B_sᵀ P already in shared memory:
  //   phase 1 (all warps)   Aug = [Hu + Y B | Y A | Y rd + Bᵀ s + ru];   t_i = P_i rd + s_i
  //   phase 2 (warp 0)      Gauss-Jordan → K, κ                      ‖  (warps 1-3) everything of the P update that does
  //                         not depend on K:  Base_i = H_{i,s} + Aᵀ P_i A,  base_i = r^x_{i,s} + Aᵀ t_i,  W_i = Aᵀ P_i B
  //   phase 3 (all warps)   P_i ← Base_i − W_i K,  s_i ← base_i − W_i κ,  Y ← B_{s−1}ᵀ P_i
  // ---------------------------------------------------------------------------------------------------------
  __device__ bool kkt_solve(double reg_x, double reg_u) {
    if (pl >= 0) return kkt_solve_ibr(reg_x, reg_u);
    int ok = 1;
    AGB_PROF(11);
    // terminal knot: P_i = H_{i,N}, s_i = r^x_{i,N}
    for (int item = tid; item < P * n * n; item += kThreads) {
      int bq = item % n, r = item / n;
      int a = r % n, i = r / n;
      Pm[item] = h_entry(i, K, a, bq, reg_x);
    }
    for (int item = tid; item < P * n; item += kThreads) {
      int a = item % n, i = item / n;
      Sv[item] = R[(K - 1) * b + OX + i * n + a];
    }
    if (K - 1 > 0) build_hm(Hm + ((K - 1) & 1) * P * HmS, K - 1, reg_x, tid, kThreads);
    __syncthreads();
    compute_Y(K - 1);
    __syncthreads();
    AGB_PROF(2);
    for (int s = K - 1; s >= 0; s--) {
      const double* Rs = R + s * b;
      // ---- phase 1: Aug = [Hu + Y B | Y A | Y rd + (Bᵀ s) + ru], items grouped by kind so that warps stay convergent
      for (int item = tid; item < m * W; item += kThreads) {
        double v;
        int r, col;
        if (item < m * m) {                                  // S block
          r = item / m; col = item - r * m;
          const double* yr = Ym + r * n1;
          const int j2 = col / P, i2 = col - j2 * P;
          double At[8], Bt[8]; loadAB(s, i2, At, Bt);
          v = bt_dot_sel(j2, Bt, yr[i2], yr[P + i2], yr[2 * P + i2], yr[3 * P + i2]);
          if (col == r) v += hu_entry(s, r, reg_u);
        } else if (item < m * m + m * n) {                   // Y A block
          const int it = item - m * m;
          r = it / n; const int a2 = it - r * n; col = m + a2;
          const double* yr = Ym + r * n1;
          const int c2 = a2 / P, i2 = a2 - c2 * P;
          v = yr[a2];
          if (c2 >= 2) {
            double At[8], Bt[8]; loadAB(s, i2, At, Bt);
            v += at_dot_sel(c2 - 2, At, yr[i2], yr[P + i2], yr[2 * P + i2], yr[3 * P + i2]);
          }
        } else {                                             // affine column
          r = item - m * m - m * n; col = m + n;
          const double* yr = Ym + r * n1;
          double v0 = yr[n] + Rs[OU + r], v1 = 0.0, v2 = 0.0, v3 = 0.0;
#pragma unroll
          for (int a2 = 0; a2 < n; a2 += 4) {
            v0 += yr[a2] * Rs[OD + a2]; v1 += yr[a2 + 1] * Rs[OD + a2 + 1];
            v2 += yr[a2 + 2] * Rs[OD + a2 + 2]; v3 += yr[a2 + 3] * Rs[OD + a2 + 3];
          }
          v = (v0 + v1) + (v2 + v3);
        }
        Aug[r * W + col] = v;
      }
      __syncthreads();
      AGB_PROF(3);
      // ---- phase 2
      if (warp == 0) {
#ifdef AGB_PHASE_TIMING
        const long long gj0 = clock64();
#endif
        if (!gj_warp(Aug, KUg + s * KUSP)) ok = 0;                           // K, κ stay in Aug for phase 3
#ifdef AGB_PHASE_TIMING
        prof[12] += clock64() - gj0;
#endif
      } else if (s > 0) {
        // one thread per row a = (c,i2) of player i's Base_i = H_{i,s} + Aᵀ P_i A, W_i = Aᵀ P_i B and base_i:
        // X = (Aᵀ P_i)[a,:] is a combination of at most four rows of P_i held in registers; A and B are block diagonal
        // per player, so the columns of player i3 need only X[(·,i3)], with the models' zero structure compiled out
        for (int row = tid - 32; row < P * n; row += kThreads - 32) {
          const int i = row / n, a = row - i * n, c = a / P, i2 = a - c * P;
          const double* Pi = Pm + i * n * n;
          const double* sv = Sv + i * n;
          double X[n];
          {
            const double2* pr = reinterpret_cast<const double2*>(Pi + a * n);
#pragma unroll
            for (int q = 0; q < n / 2; q++) { const double2 v = pr[q]; X[2 * q] = v.x; X[2 * q + 1] = v.y; }
          }
          double xs = sv[a];
          if (c >= 2) {
            if constexpr (MODEL == AGB_MODEL_DOUBLE_INTEGRATOR) {            // Aᵀ[(c,i2)][(c-2,i2)] = dt
              const double2* pq = reinterpret_cast<const double2*>(Pi + ((c - 2) * P + i2) * n);
#pragma unroll
              for (int q = 0; q < n / 2; q++) { const double2 v = pq[q]; X[2 * q] = fma(dt, v.x, X[2 * q]); X[2 * q + 1] = fma(dt, v.y, X[2 * q + 1]); }
              xs = fma(dt, sv[(c - 2) * P + i2], xs);
            } else {
              double At2[8], Bt2[8]; loadAB(s, i2, At2, Bt2);
#pragma unroll
              for (int q = 0; q < 4; q++) {
                if (((AT_NZ >> (2 * q)) & 3u) != 0u) {                       // row q of Ã has a structural non-zero
                  const double cf = (c == 2) ? At2[2 * q] : At2[2 * q + 1];
                  const double2* pq = reinterpret_cast<const double2*>(Pi + (q * P + i2) * n);
#pragma unroll
                  for (int q2 = 0; q2 < n / 2; q2++) { const double2 v = pq[q2]; X[2 * q2] = fma(cf, v.x, X[2 * q2]); X[2 * q2 + 1] = fma(cf, v.y, X[2 * q2 + 1]); }
                  xs = fma(cf, sv[q * P + i2], xs);
                }
              }
            }
          }
          double* brow = Base + (i * n + a) * n1;
          double* wrow = Wm + (i * n + a) * m;
          {                                                                  // affine: r^x_{i,s}[a] + (Aᵀ(P_i rd + s_i))[a]
            const double2* rd2 = reinterpret_cast<const double2*>(Rs + OD);
            double v0 = xs + R[(s - 1) * b + OX + i * n + a], v1 = 0.0;
#pragma unroll
            for (int q = 0; q < n / 2; q++) { const double2 e = rd2[q]; v0 = fma(X[2 * q], e.x, v0); v1 = fma(X[2 * q + 1], e.y, v1); }
            brow[n] = v0 + v1;
          }
          const double* hm = Hm + (s & 1) * P * HmS + i * HmS;            // H^x_{i,s}, laid out during the previous stage
          const double hd = hm[4 * P * P + a];
          const double* hrow = hm + (c < 2 ? a : 0) * (2 * P);
          const double hmask = (c < 2) ? 1.0 : 0.0;                       // position rows only (branch-free below)
#pragma unroll
          for (int i3 = 0; i3 < P; i3++) {
            double At3[8], Bt3[8]; loadAB(s, i3, At3, Bt3);
            const double x0 = X[i3], x1 = X[P + i3], x2 = X[2 * P + i3], x3 = X[3 * P + i3];
            double o0 = x0, o1 = x1;
            double o2 = x2 + at_dot<0>(At3, x0, x1, x2, x3), o3 = x3 + at_dot<1>(At3, x0, x1, x2, x3);
            o0 = fma(hmask, hrow[i3], o0); o1 = fma(hmask, hrow[P + i3], o1);
            brow[i3] = o0; brow[P + i3] = o1; brow[2 * P + i3] = o2; brow[3 * P + i3] = o3;
            wrow[i3] = bt_dot<0>(Bt3, x0, x1, x2, x3);
            wrow[P + i3] = bt_dot<1>(Bt3, x0, x1, x2, x3);
          }
          brow[a] += hd;                                                     // diagonal of H^x_{i,s} (same thread wrote brow[a])
        }
        // the remaining warps lay out H^x_{i,s-1} for the next stage
        if (s - 1 > 0) build_hm(Hm + ((s - 1) & 1) * P * HmS, s - 1, reg_x, tid - HB, kThreads - HB);
      }
      __syncthreads();
      AGB_PROF(4);
      if (s == 0) break;
      // ---- phase 3
      {
        const double* ku = Aug + m;                         // K | κ: columns m.. of the reduced system
        for (int item = tid; item < P * P * n1; item += kThreads) {
          const int col = item % n1, t = item / n1;
          const int i2 = t % P, i = t / P;
          double kc[m];
#pragma unroll
          for (int r = 0; r < m; r++) kc[r] = ku[r * W + col];
          double pn[4];
#ifdef AGB_P3_PREFETCH
          // all four row groups are fetched before the first product, so that one shared-memory latency is exposed instead of four
          double2 wv[4][m / 2]; double bs[4];
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const int a = c * P + i2;
            const double2* w = reinterpret_cast<const double2*>(Wm + (i * n + a) * m);
            bs[c] = Base[(i * n + a) * n1 + col];
#pragma unroll
            for (int r = 0; r < m / 2; r++) wv[c][r] = w[r];
          }
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const int a = c * P + i2;
            double v0 = bs[c], v1 = 0.0;
#pragma unroll
            for (int r = 0; r < m; r += 2) { v0 -= wv[c][r / 2].x * kc[r]; v1 -= wv[c][r / 2].y * kc[r + 1]; }
            pn[c] = v0 + v1;
            if (col < n) Pm[(i * n + a) * n + col] = pn[c]; else Sv[i * n + a] = pn[c];
          }
#else
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const int a = c * P + i2;
            const double2* w = reinterpret_cast<const double2*>(Wm + (i * n + a) * m);      // rows of W are 16-byte aligned (m even)
            double v0 = Base[(i * n + a) * n1 + col], v1 = 0.0;
#pragma unroll
            for (int r = 0; r < m; r += 2) { const double2 wv = w[r / 2]; v0 -= wv.x * kc[r]; v1 -= wv.y * kc[r + 1]; }
            pn[c] = v0 + v1;
            if (col < n) Pm[(i * n + a) * n + col] = pn[c]; else Sv[i * n + a] = pn[c];
          }
#endif
          if (i2 == i) {                                    // Y for the next stage (s-1)
            double At[8], Bt[8]; loadAB(s - 1, i, At, Bt);
            Ym[(0 * P + i) * n1 + col] = bt_dot<0>(Bt, pn[0], pn[1], pn[2], pn[3]);
            Ym[(1 * P + i) * n1 + col] = bt_dot<1>(Bt, pn[0], pn[1], pn[2], pn[3]);
          }
        }
      }
      __syncthreads();
      AGB_PROF(5);
    }
    ok = __syncthreads_and(ok);
    forward_and_costate(reg_x);
    return ok != 0;
  }

  // g_{i,k}[a] = r^x_{i,k}[a] + (H_{i,k} Δx_k)[a], k = s+1 (the opt-x row of player i without its multiplier terms)
  template <int HALF = -1> __device__ __forceinline__ double costate_g(int i, int s, int a, double reg_x) const {
    const int k = s + 1, c = a / P, ia = a - c * P;
    const bool is_pos = HALF < 0 ? (c < 2) : (HALF == 0);
    const double* dx = R + s * b + OD;
    double v = R[s * b + OX + i * n + a] + hd_entry(i, k, a, reg_x) * dx[a];
    if (is_pos) {
      if (ia == i) {
        if (has_pairs) {
#pragma unroll
          for (int jj = 0; jj < P - 1; jj++) {
            const int j = jj < i ? jj : jj + 1;
            const double* hp = Hp + (k * NP + i * (P - 1) + jj) * 3;
            v += hp[c] * (dx[i] - dx[j]) + hp[c + 1] * (dx[P + i] - dx[P + j]);
          }
        }
        if (has_self) { const double* hs = Hs + (k * P + i) * 3; v += hs[c] * dx[i] + hs[c + 1] * dx[P + i]; }
      } else if (has_pairs) {
        const double* hp = Hp + (k * NP + pair_index(i, ia)) * 3;
        v -= hp[c] * (dx[i] - dx[ia]) + hp[c + 1] * (dx[P + i] - dx[P + ia]);
      }
    }
    return v;
  }

  // forward sweep + costate recursion shared by the game and the best-response factorisations
  __device__ void forward_and_costate(double reg_x) {
    // ---- forward sweep.  The recurrence Δx_{s+1} = (A − B K_s) Δx_s + (rd_s − B κ_s) is a serial chain over the stages, and a
    // dependent FP64 operation costs ≈ 20 cycles, so the chain is kept as short as it can be: the closed-loop blocks
    // Acl_s = [A − B K_s | rd_s − B κ_s] are PRODUCED by warps 1.. (gains back from the L2-resident scratch, one warp per
    // stage, round robin) into a shared-memory ring laid over the factorisation's scratch (P, s, Y, Aug, Base …: free at this
    // point), and CONSUMED by warp 0, whose lane a owns component a: one row · Δx_s product in four accumulators, one store,
    // one warp barrier per stage.  ready[slot] = s + 1 publishes stage s, done = s + 1 frees its slot (release / acquire
    // flags in shared memory).  The controls Δu_s = −K_s Δx_s − κ_s do not feed the chain: they are formed afterwards by all
    // threads in parallel.
    {
      double* const ring = Pm;
      int* const flags = reinterpret_cast<int*>(ring + FWD_D * ACS);      // ready[FWD_D], done
      if (tid <= FWD_D) flags[tid] = 0;
      __syncthreads();
      constexpr int NW = kThreads / 32 - 1;                               // producer warps
      if (warp > 0) {
        for (int s = warp - 1; s < K; s += NW) {
          const int slot = s % FWD_D;
          const double* kg = KUg + s * KUSP;
          // gains first (L2 latency), then wait for the slot
          constexpr int NQ = (P * n1 + 31) / 32;
          double k0[NQ], k1[NQ];
          int it[NQ];
#pragma unroll
          for (int q = 0; q < NQ; q++) {
            it[q] = lane + 32 * q;
            const int i = it[q] / n1, col = it[q] - i * n1;
            const bool on = it[q] < P * n1;
            k0[q] = on ? kg[(0 * P + i) * n1p + col] : 0.0;
            k1[q] = on ? kg[(1 * P + i) * n1p + col] : 0.0;
          }
          while (s - flag_load(flags + FWD_D) >= FWD_D) spin_hint(true);
          double* dst = ring + slot * ACS;
#pragma unroll
          for (int q = 0; q < NQ; q++) {
            if (it[q] < P * n1) {
              const int i = it[q] / n1, col = it[q] - i * n1;
              double At[8], Bt[8]; loadAB(s, i, At, Bt);
              const int c2 = col / P, i2 = col - c2 * P;
              const bool own = (col < n) && (i2 == i);
#pragma unroll
              for (int c = 0; c < 4; c++) {
                const int a = c * P + i;
                double base;
                if (col == n) base = R[s * b + OD + a];
                else {
                  base = (own && c2 == c) ? 1.0 : 0.0;
                  if (own && c2 == 2) base += At[c * 2];
                  if (own && c2 == 3) base += At[c * 2 + 1];
                }
                dst[a * n1p + col] = (base - Bt[c * 2] * k0[q]) - Bt[c * 2 + 1] * k1[q];
              }
            }
          }
          __syncwarp();
          if (lane == 0) flag_store(flags + slot, s + 1);
        }
      } else {
        const bool act = lane < n;
        const int a = act ? lane : 0;
        for (int s = 0; s < K; s++) {
          const int slot = s % FWD_D;
          while (flag_load(flags + slot) != s + 1) spin_hint(false);
          const double2* row = reinterpret_cast<const double2*>(ring + slot * ACS + a * n1p);
          double v0 = row[n / 2].x, v1 = 0.0, v2 = 0.0, v3 = 0.0;          // affine column
          if (s > 0) {
            const double2* dx = reinterpret_cast<const double2*>(R + (s - 1) * b + OD);
#pragma unroll
            for (int q = 0; q < n / 2; q += 2) {
              const double2 r0 = row[q], r1 = row[q + 1], x0 = dx[q], x1 = dx[q + 1];
              v0 = fma(r0.x, x0.x, v0); v1 = fma(r0.y, x0.y, v1); v2 = fma(r1.x, x1.x, v2); v3 = fma(r1.y, x1.y, v3);
            }
          }
          const double v = (v0 + v1) + (v2 + v3);
          if (act) R[s * b + OD + a] = v;                                  // rd_s → Δx_{s+1} (its producer has read rd_s)
          __syncwarp();                                                    // Δx_{s+1} visible; every lane is done with the slot
          if (lane == 0) flag_store(flags + FWD_D, s + 1);
        }
      }
      __syncthreads();
      // Δu_s = −K_s Δx_s − κ_s for every stage (Δx_0 = 0: x_1 is not a variable)
      for (int item = tid; item < K * m; item += kThreads) {
        const int s = item / m, r = item - s * m;
        const double2* kr = reinterpret_cast<const double2*>(KUg + s * KUSP + r * n1p);
        double u0 = kr[n / 2].x, u1 = 0.0, u2 = 0.0, u3 = 0.0;
        if (s > 0) {
          const double2* dx = reinterpret_cast<const double2*>(R + (s - 1) * b + OD);
#pragma unroll
          for (int q = 0; q < n / 2; q += 2) {
            const double2 g0 = kr[q], g1 = kr[q + 1], x0 = dx[q], x1 = dx[q + 1];
            u0 = fma(g0.x, x0.x, u0); u1 = fma(g0.y, x0.y, u1); u2 = fma(g1.x, x1.x, u2); u3 = fma(g1.y, x1.y, u3);
          }
        }
        R[s * b + OU + r] = -((u0 + u1) + (u2 + u3));
      }
    }
    __syncthreads();
    AGB_PROF(6);
    // ---- costate: Δλ_{i,k−1} = g_{i,k} + A_kᵀ Δλ_{i,k} with g_{i,k} = H_{i,k} Δx_k + r^x_{i,k} (exactly the opt-x rows).
    // One warp per player, Δλ in registers (lane a owns component a); g of the next step is evaluated ahead of the
    // dependent shuffle → Aᵀ chain.
#pragma unroll 1
    for (int half = 0; half < 2; half++) {                  // position components first, then velocity / heading: convergent warps
      for (int item = tid; item < P * K * 2 * P; item += kThreads) {
        const int s = item / (P * 2 * P), rem = item - s * (P * 2 * P);
        const int i = rem / (2 * P), a = half * 2 * P + (rem - i * (2 * P));
        double v = 0.0;                                     // IBR: Δλ_j = 0 for j != pl
        if (pl < 0 || i == pl) v = half == 0 ? costate_g<0>(i, s, a, reg_x) : costate_g<1>(i, s, a, reg_x);
        R[s * b + OX + i * n + a] = v;
      }
    }
    __syncthreads();
    AGB_PROF(7);
    // One thread per (player i, state block of player j): the four components of Δλ_i on player j's states stay in
    // registers (A is block diagonal per player), so a step is load g → add → at most two dependent FMAs → store, with
    // no shuffles; the next step's g is fetched ahead of the chain.
    if (tid < P * P && (pl < 0 || tid / P == pl)) {
      const int i = tid / P, j = tid - i * P;
      double* gi = R + OX + i * n + j;                   // component c of the block: gi[s·b + c·P]
      double ln[4] = {0.0, 0.0, 0.0, 0.0}, g[4];
#pragma unroll
      for (int c = 0; c < 4; c++) g[c] = gi[(K - 1) * b + c * P];
      for (int s = K - 1; s >= 0; s--) {
        double gn[4] = {0.0, 0.0, 0.0, 0.0};
        if (s > 0) {
#pragma unroll
          for (int c = 0; c < 4; c++) gn[c] = gi[(s - 1) * b + c * P];
        }
        double v[4];
        if (s < K - 1) {
          double At[8], Bt[8]; loadAB(s + 1, j, At, Bt);
          v[0] = g[0] + ln[0]; v[1] = g[1] + ln[1];
          v[2] = (g[2] + ln[2]) + at_dot<0>(At, ln[0], ln[1], ln[2], ln[3]);
          v[3] = (g[3] + ln[3]) + at_dot<1>(At, ln[0], ln[1], ln[2], ln[3]);
        } else {
#pragma unroll
          for (int c = 0; c < 4; c++) v[c] = g[c];
        }
#pragma unroll
        for (int c = 0; c < 4; c++) { gi[s * b + c * P] = v[c]; ln[c] = v[c]; g[c] = gn[c]; }
      }
    }
    __syncthreads();
    AGB_PROF(8);
  }


  // line_search (solver_methods.jl:105-125): returns alpha, j through references; n_eval counts residual evaluations.
  // With `accepted_rec` every trial also leaves its un-regularised rows in Rtrial (global) and, on acceptance, its norms
  // in *accepted_rec: the accepted point IS the next iterate, so the next inner iteration can skip its residual!.
  __device__ bool line_search(const agb_options& o, double reg, double res_norm, double& alpha, int& j, int& n_eval,
                              Acc* accepted_rec = nullptr) {
    const double S = res_size();
    const double rr = o.regularize ? reg : 0.0;
    alpha = 1.0; j = 1;
    while (j < o.ls_iter) {
      Acc plain;
      Acc t = accepted_rec ? residual<true>(alpha, rr, rr, Rtrial, &plain) : residual<true>(alpha, rr, rr, nullptr);
      n_eval++;
      const double trial = t.sum / S;
      if (trial <= (1.0 - alpha * o.beta) * res_norm) {
        if (accepted_rec) *accepted_rec = plain;
        return true;
      }
      alpha *= o.alpha_decrease;
      j++;
    }
    return false;
  }

  // R <- rows kept by the accepted trial evaluation
  __device__ void load_kept_residual() {
    // K·b is even and both arrays are 16-byte aligned: 128-bit copies, four loads in flight per thread (L2 latency)
    const double2* src = reinterpret_cast<const double2*>(Rtrial);
    double2* dst = reinterpret_cast<double2*>((double*)R);
    const int nv = (K * b) >> 1;
    int q = tid;
    for (; q + 3 * kThreads < nv; q += 4 * kThreads) {
      const double2 v0 = src[q], v1 = src[q + kThreads], v2 = src[q + 2 * kThreads], v3 = src[q + 3 * kThreads];
      dst[q] = v0; dst[q + kThreads] = v1; dst[q + 2 * kThreads] = v2; dst[q + 3 * kThreads] = v3;
    }
    for (; q < nv; q += kThreads) dst[q] = src[q];
    __syncthreads();
  }

  // update_traj!(pdtraj, pdtraj, α, Δpdtraj) + Δ_step (primal_dual_traj.jl:109-147)
  __device__ double update_traj(double alpha) {
    double loc = 0.0;
    // duals: Λ_{i,s} += α·Δλ_{i,s} (stage-major items; four independent read-modify-writes in flight — Λ may live in L2)
    {
      double* Lp = L;
      const int nl = K * P * n;
      int item = tid;
      for (; item + 3 * kThreads < nl; item += 4 * kThreads) {
        int idx[4]; double dv[4], lv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int it = item + u * kThreads, s = it / (P * n), rem = it - s * (P * n), i = rem / n, a = rem - i * n;
          idx[u] = (i * K + s) * n + a; dv[u] = R[s * b + OX + rem];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) lv[u] = Lp[idx[u]];
#pragma unroll
        for (int u = 0; u < 4; u++) Lp[idx[u]] = fma(alpha, dv[u], lv[u]);
      }
      for (; item < nl; item += kThreads) {
        const int s = item / (P * n), rem = item - s * (P * n), i = rem / n, a = rem - i * n;
        double* t = &Lp[(i * K + s) * n + a]; *t = fma(alpha, R[s * b + OX + rem], *t);
      }
    }
    // primals: u_s += α·Δu_s, x_{s+1} += α·Δx_{s+1}; Δ_step accumulates |Δu| + |Δx| (primal_dual_traj.jl:130-147)
    for (int item = tid; item < K * (m + n); item += kThreads) {
      const int s = item / (m + n), q = item - s * (m + n);
      const double dv = R[s * b + OU + q];
      if (q < m) { double* t = &U[s * m + q]; *t = fma(alpha, dv, *t); }
      else { double* t = &X[(s + 1) * n + (q - m)]; *t = fma(alpha, dv, *t); }
      loc += fabs(dv);
    }
    Acc a = {loc, 0.0, 0.0, 0.0, 0.0, 0.0};
    a = block_reduce(a);
    return a.sum * alpha / (double)(K * (n + m));
  }

  // rollout!(RK3, model, traj): independent per player (separable dynamics)
  __device__ void rollout() {
    if constexpr (MODEL == AGB_MODEL_UNICYCLE) {
      // Unicycle: (θ, v) evolve from the controls alone and the position derivative depends on (θ, v) only, so the serial chain
      // of rk3_step splits into three passes with the same arithmetic (same operations in the same order per component):
      //   A  2P threads   θ_k, v_k for every knot            (a chain of additions, no trigonometry)
      //   B  P·K threads  the position increment of stage k   (the three sincos of the RK3 stages, all stages in parallel)
      //   C  2P threads   x_{k+1} = x_k + increment           (a chain of additions)
      if (tid < 2 * P) {
        const int i = tid % P, c = 2 + tid / P, j = tid / P;
        double st = X[c * P + i];
        for (int s = 0; s < K; s++) {
          const double u = U[s * m + j * P + i];
          const double k1 = u * dt, k2 = u * dt, k3 = u * dt;
          st = st + (k1 + 4 * k2 + k3) / 6;
          X[(s + 1) * n + c * P + i] = st;
        }
      }
      __syncthreads();
      for (int item = tid; item < P * K; item += kThreads) {
        const int i = item % P, s = item / P;
        const double th = X[s * n + 2 * P + i], v = X[s * n + 3 * P + i], w = U[s * m + i], a = U[s * m + P + i];
        const double k1t = w * dt, k1v = a * dt;
        double sn, cs;
        sincos(th, &sn, &cs);
        const double k1x = (cs * v) * dt, k1y = (sn * v) * dt;
        const double tb = th + k1t / 2, vb = v + k1v / 2;
        sincos(tb, &sn, &cs);
        const double k2x = (cs * vb) * dt, k2y = (sn * vb) * dt, k2t = w * dt, k2v = a * dt;
        const double tc = th - k1t + 2 * k2t, vc = v - k1v + 2 * k2v;
        sincos(tc, &sn, &cs);
        const double k3x = (cs * vc) * dt, k3y = (sn * vc) * dt;
        X[(s + 1) * n + 0 * P + i] = (k1x + 4 * k2x + k3x) / 6;          // increments, summed by pass C
        X[(s + 1) * n + 1 * P + i] = (k1y + 4 * k2y + k3y) / 6;
      }
      __syncthreads();
      if (tid < 2 * P) {
        const int i = tid % P, c = tid / P;
        double st = X[c * P + i];
        for (int s = 0; s < K; s++) { st = st + X[(s + 1) * n + c * P + i]; X[(s + 1) * n + c * P + i] = st; }
      }
      __syncthreads();
      return;
    }
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
  static __device__ __forceinline__ bool bulk_ok(const void* gp, unsigned bytes) { return bytes > 0 && (bytes & 15u) == 0 && (((size_t)gp) & 15u) == 0; }
  // Instance slab → shared memory.  Λ, the AL multipliers and the penalties have the same layout in HBM and in shared
  // memory: thread 0 issues them as bulk copies on ONE mbarrier phase (a phase must not complete twice before every waiter
  // has seen it, so the three slabs share a transaction and the next phase starts an instance later, many block barriers
  // away); z_k = [x_k; u_k] is de-interleaved by element loops that run beside the bulk copies.
  __device__ void load_iterate(const double* Zg, const double* Lg, int inst, const Buffers* gd = nullptr) {
    const double* __restrict__ z = Zg + (size_t)inst * N * (n + m);
    const double* __restrict__ l = Lg + (size_t)inst * P * K * n;
    const size_t o = (size_t)inst * K * nrow;
    const unsigned lbytes = (unsigned)(P * K * n) * 8u, cbytes = (unsigned)(K * nrow) * 8u;
    bool lb = false, cb = false;
    if constexpr (!BIG) {
      lb = bulk_ok(l, lbytes);
      cb = gd != nullptr && bulk_ok(gd->conlam + o, cbytes) && bulk_ok(gd->conmu + o, cbytes);
      if ((lb || cb) && tid == 0) {
        mbar_expect((double*)red + 47, (lb ? lbytes : 0u) + (cb ? 2u * cbytes : 0u));
        if (lb) bulk_g2s((double*)L, l, lbytes, (double*)red + 47);
        if (cb) { bulk_g2s((double*)CL, gd->conlam + o, cbytes, (double*)red + 47); bulk_g2s((double*)CM, gd->conmu + o, cbytes, (double*)red + 47); }
      }
    }
#pragma unroll 4
    for (int item = tid; item < N * (n + m); item += kThreads) {
      const int q = item % (n + m), k = item / (n + m);
      const double v = __ldg(z + item);
      if (q < n) X[k * n + q] = v; else U[k * m + (q - n)] = v;
    }
    if (!lb) {
#pragma unroll 4
      for (int item = tid; item < P * K * n; item += kThreads) L[item] = __ldg(l + item);
    }
    if (gd != nullptr && !cb) {
      for (int item = tid; item < K * nrow; item += kThreads) { CL[item] = gd->conlam[o + item]; CM[item] = gd->conmu[o + item]; }
    }
    if (lb || cb) { mbar_wait((double*)red + 47, bulk_phase); bulk_phase ^= 1u; }
  }
  __device__ void store_iterate(double* Zg, double* Lg, int inst) {
    double* z = Zg + (size_t)inst * N * (n + m);
    double* l = Lg + (size_t)inst * P * K * n;
    const unsigned lbytes = (unsigned)(P * K * n) * 8u;
    bool bulk = false;
    if constexpr (!BIG) {
      bulk = bulk_ok(l, lbytes);
      if (bulk) {                                            // writers order their shared-memory stores before the async proxy reads them
        fence_async_smem();
        __syncthreads();
        if (tid == 0) bulk_s2g(l, (double*)L, lbytes);
      }
    }
    for (int item = tid; item < N * (n + m); item += kThreads) {
      const int q = item % (n + m), k = item / (n + m);
      z[item] = (q < n) ? X[k * n + q] : U[k * m + (q - n)];
    }
    if (bulk) { if (tid == 0) bulk_commit_wait(); }          // (Λ's shared copy is not rewritten before the next block barrier)
    else for (int item = tid; item < P * K * n; item += kThreads) l[item] = L[item];
  }
  // Receding-horizon advance inside the CTA (the in-kernel form of agb_advance_kernel + agb_shift_kernel, agb_capi.cu): the iterate
  // moves s knots forward — z_k ← z_{k+s}, λ_k ← λ_{k+s}, zero tail (init_traj! with s, primal_dual_traj.jl:34-41) — and
  // x_1 ← x_{1+s} + disturbance.  One thread per component walks its column in ascending k, so the shift is in place.
  __device__ void mpc_shift(int s, const double* __restrict__ dist) {
    const double x0n = (tid < n) ? X[s * n + tid] + (dist ? dist[tid] : 0.0) : 0.0;
    shift_array((double*)X, N * n, s * n);
    shift_array((double*)U, N * m, s * m);
    for (int i = 0; i < P; i++) shift_array((double*)L + i * K * n, K * n, s * n);
    if (tid < n) X[tid] = x0n;
  }
  // a[item] ← a[item + by] (zero past the end), in ascending chunks of 4·kThreads elements: every thread reads its elements of the
  // chunk into registers, block barrier, writes.  A later chunk reads only elements no earlier chunk wrote (by ≥ 1).
  __device__ __forceinline__ void shift_array(double* a, int len, int by) {
    for (int c0 = 0; c0 < len; c0 += 4 * kThreads) {
      double v[4];
#pragma unroll
      for (int q = 0; q < 4; q++) { const int item = c0 + tid + q * kThreads; v[q] = (item + by < len) ? a[item + by] : 0.0; }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; q++) { const int item = c0 + tid + q * kThreads; if (item < len) a[item] = v[q]; }
    }
  }
  __device__ void load_duals(const Buffers& g, int inst) {
    const size_t o = (size_t)inst * K * nrow;
    for (int item = tid; item < K * nrow; item += kThreads) { CL[item] = g.conlam[o + item]; CM[item] = g.conmu[o + item]; }
  }
  __device__ void store_duals(const Buffers& g, int inst) {
    const size_t o = (size_t)inst * K * nrow;
    const unsigned cbytes = (unsigned)(K * nrow) * 8u;
    if constexpr (!BIG) {
      if (bulk_ok(g.conlam + o, cbytes) && bulk_ok(g.conmu + o, cbytes)) {
        fence_async_smem();
        __syncthreads();
        if (tid == 0) { bulk_s2g(g.conlam + o, (double*)CL, cbytes); bulk_s2g(g.conmu + o, (double*)CM, cbytes); bulk_commit_wait(); }
        return;
      }
    }
    for (int item = tid; item < K * nrow; item += kThreads) { g.conlam[o + item] = CL[item]; g.conmu[o + item] = CM[item]; }
  }
};

}  // namespace agb
