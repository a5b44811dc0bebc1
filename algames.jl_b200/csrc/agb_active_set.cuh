// agb_active_set.cuh — active-set analysis of the resident iterate (reference: src/active_set/*.jl), included by agb_capi.cu.
//
// The reference borders the Newton system with one row per unordered collision pair and knot and one column per ordered
// pair and knot (ActiveSetCore, active_set_core.jl:57-160), masks the rows / columns of inactive pairs
// (active_set_methods.jl:28-74) and takes the null space of the masked matrix (update_nullspace!, :173-184) — the directions
// along which the equilibrium can move.  Off the solve path: one CTA per instance works on a dense matrix in global memory.
#pragma once

namespace agb {

// unordered / ordered pair numbering of ActiveSetCore's stamps (active_set_core.jl:118-126, :151-159)
__host__ __device__ inline int as_unordered_index(int P, int i, int j) { return i * P - i * (i + 1) / 2 + (j - i - 1); }      // i < j
__host__ __device__ inline int as_ordered_index(int P, int i, int j) { return i * (P - 1) + (j < i ? j : j - 1); }            // i != j

// Border of the KKT Jacobian and tail of the residual: one thread per (instance, stage, ordered pair).
//   jac[(opt_i x_k), (h,col,i,j,k)] += ∇c_ijᵀ      (active_set_methods.jl:154-156)
//   jac[(v,col,i,j,k), (x_k)]       += ∇c_ij  i<j   (:157-161)         c_ij = r_ij² − ‖p_i − p_j‖²
//   res[(v,col,i,j,k)]              += c_ij   i<j   (:113-121)
//   vmask / hmask: 1 for the Newton rows / columns and for the pairs whose collision row is active (`act`, agb_active_set;
//   all pairs when act == nullptr)                                     (:28-74)
__global__ void agb_as_border_kernel(const DevDesc* __restrict__ dd, const double* __restrict__ Z, const unsigned char* __restrict__ act,
                                     int batch, double* jac, double* res, unsigned char* vmask, unsigned char* hmask) {
  const int P = dd->p, n = dd->n, m = dd->m, N = dd->N, K = dd->K, b = dd->b, S = dd->S, nrow = dd->nrow;
  const int npo = P * (P - 1), npu = npo / 2, Sv = S + K * npu, Sh = S + K * npo;
  const size_t total = (size_t)batch * K * npo;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int inst = (int)(t / ((size_t)K * npo));
    const int rem = (int)(t % ((size_t)K * npo)), s = rem / npo, op = rem % npo;
    const int i = op / (P - 1), jj = op % (P - 1), j = jj < i ? jj : jj + 1;
    const int k = s + 1;                                                    // 0-based knot of x_{k} (the reference's k = s + 2)
    const double* z = Z + ((size_t)inst * N + k) * (n + m);
    const double dx = z[0 * P + i] - z[0 * P + j], dy = z[1 * P + i] - z[1 * P + j];
    const double rad = dd->col_radius[i][j];
    const int crow = dd->col_row[i][j];
    const bool on = crow >= 0 && (act == nullptr || act[((size_t)inst * K + s) * nrow + crow] != 0);
    const int hcol = S + s * npo + op;
    if (hmask) hmask[(size_t)inst * Sh + hcol] = on ? 1 : 0;
    const int cidx[4] = {0 * P + i, 1 * P + i, 0 * P + j, 1 * P + j};
    const double g[4] = {-2.0 * dx, -2.0 * dy, 2.0 * dx, 2.0 * dy};
    if (jac && crow >= 0) {
      double* J = jac + (size_t)inst * Sv * Sh;
      const size_t vx = (size_t)(i * K + s) * (n + 2);                      // rows (opt_i, x_k) in the reference's vertical order
      for (int q = 0; q < 4; q++) J[(vx + cidx[q]) * Sh + hcol] += g[q];
    }
    if (i < j) {
      const int vrow = S + s * npu + as_unordered_index(P, i, j);
      if (vmask) vmask[(size_t)inst * Sv + vrow] = on ? 1 : 0;
      if (crow >= 0) {
        if (jac) { double* J = jac + (size_t)inst * Sv * Sh; for (int q = 0; q < 4; q++) J[(size_t)vrow * Sh + s * b + cidx[q]] += g[q]; }
        if (res) res[(size_t)inst * Sv + vrow] += rad * rad - (dx * dx + dy * dy);
      }
    }
  }
}

// Block-wide helpers of the null-space kernel (256 threads)
constexpr int kAsThreads = 256;
__device__ inline double as_block_sum(double v, double* buf) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(AGB_FULL, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) buf[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < kAsThreads / 32; w++) r += buf[w];
  return r;
}

// nullspace(jac[vmask, hmask]) for every instance (active_set_methods.jl:173-184).  Gauss–Jordan elimination with COMPLETE pivoting
// on the masked matrix (r x c, global memory): the pivot search of step t+1 rides on the update pass of step t; elimination
// stops when every remaining entry is <= atol (the reference's `nullspace(djac, atol = 1e-20)` drops singular values <= atol),
// which leaves rank pivots and c − rank free columns.  Free column f gives the null vector x_f = 1, x_p = −M[p][f] / M[p][p];
// the vectors are scattered to the hmask rows (add_matrix!, active_set_core.jl:29-42) and orthonormalised (two passes of modified
// Gram–Schmidt), as LinearAlgebra.nullspace returns an orthonormal basis.  null_out [B][max_dim][Sh], dim_out [B] (the dimension
// even when it exceeds max_dim; then only the first max_dim vectors are written, not orthonormalised against the missing ones).
__global__ void __launch_bounds__(kAsThreads) agb_as_nullspace_kernel(int batch, int Sv, int Sh, const double* __restrict__ jac,
                                                                      const unsigned char* __restrict__ vmask, const unsigned char* __restrict__ hmask,
                                                                      double* work, int* iwork, double atol, int max_dim, double* null_out, int* dim_out) {
  __shared__ double sbuf[kAsThreads / 32];
  __shared__ double s_pmax[kAsThreads / 32];
  __shared__ int s_pi[kAsThreads / 32], s_pj[kAsThreads / 32];
  __shared__ int s_r, s_c;
  const int tid = threadIdx.x, nt = kAsThreads, lane = tid & 31, warp = tid >> 5, nw = nt / 32;
  for (int inst = blockIdx.x; inst < batch; inst += gridDim.x) {
    double* M = work + (size_t)inst * ((size_t)Sv * Sh + Sv);
    double* fcol = M + (size_t)Sv * Sh;
    int* rows = iwork + (size_t)inst * (Sv + 2 * Sh);
    int* cols = rows + Sv;
    int* cperm = cols + Sh;
    const double* J = jac + (size_t)inst * Sv * Sh;
    __syncthreads();
    if (tid == 0) {
      int r = 0, c = 0;
      for (int q = 0; q < Sv; q++) if (vmask[(size_t)inst * Sv + q]) rows[r++] = q;
      for (int q = 0; q < Sh; q++) if (hmask[(size_t)inst * Sh + q]) { cperm[c] = c; cols[c++] = q; }
      s_r = r; s_c = c;
    }
    __syncthreads();
    const int r = s_r, c = s_c;
    // gather the masked matrix and find the first pivot
    double lmax = -1.0; int li = 0, lj = 0;
    for (int a = warp; a < r; a += nw) {
      const double* src = J + (size_t)rows[a] * Sh;
      for (int q = lane; q < c; q += 32) {
        const double v = src[cols[q]];
        M[(size_t)a * c + q] = v;
        const double av = fabs(v);
        if (av > lmax) { lmax = av; li = a; lj = q; }
      }
    }
    int rank = 0;
    double amax = 0.0;
    const int tmax = r < c ? r : c;
    for (int t = 0;; t++) {
      // block arg-max of (lmax, li, lj)
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(AGB_FULL, lmax, o);
        const int oi = __shfl_xor_sync(AGB_FULL, li, o), oj = __shfl_xor_sync(AGB_FULL, lj, o);
        if (ov > lmax || (ov == lmax && (oi < li || (oi == li && oj < lj)))) { lmax = ov; li = oi; lj = oj; }
      }
      __syncthreads();
      if (lane == 0) { s_pmax[warp] = lmax; s_pi[warp] = li; s_pj[warp] = lj; }
      __syncthreads();
      double pmax = s_pmax[0]; int pi = s_pi[0], pj = s_pj[0];
      for (int w = 1; w < nw; w++) {
        const double ov = s_pmax[w]; const int oi = s_pi[w], oj = s_pj[w];
        if (ov > pmax || (ov == pmax && (oi < pi || (oi == pi && oj < pj)))) { pmax = ov; pi = oi; pj = oj; }
      }
      if (t == 0) amax = pmax;
      if (t >= tmax || !(pmax > atol)) { rank = t; break; }
      if (pi != t) for (int q = tid; q < c; q += nt) { const double v = M[(size_t)t * c + q]; M[(size_t)t * c + q] = M[(size_t)pi * c + q]; M[(size_t)pi * c + q] = v; }
      __syncthreads();
      if (pj != t) {
        for (int a = tid; a < r; a += nt) { const double v = M[(size_t)a * c + t]; M[(size_t)a * c + t] = M[(size_t)a * c + pj]; M[(size_t)a * c + pj] = v; }
        if (tid == 0) { const int v = cperm[t]; cperm[t] = cperm[pj]; cperm[pj] = v; }
      }
      __syncthreads();
      const double piv = M[(size_t)t * c + t];
      for (int a = tid; a < r; a += nt) fcol[a] = (a == t) ? 0.0 : M[(size_t)a * c + t] / piv;
      __syncthreads();
      lmax = -1.0; li = 0; lj = 0;
      const double* prow = M + (size_t)t * c;
      for (int a = warp; a < r; a += nw) {
        if (a == t) continue;
        const double f = fcol[a];
        double* row = M + (size_t)a * c;
        if (lane == 0) row[t] = 0.0;
        if (f != 0.0) {
          for (int q = t + 1 + lane; q < c; q += 32) {
            const double v = fma(-f, prow[q], row[q]);
            row[q] = v;
            if (a > t) { const double av = fabs(v); if (av > lmax) { lmax = av; li = a; lj = q; } }
          }
        } else if (a > t) {
          for (int q = t + 1 + lane; q < c; q += 32) { const double av = fabs(row[q]); if (av > lmax) { lmax = av; li = a; lj = q; } }
        }
      }
      __syncthreads();
    }
    // An EXACT rank deficiency (structurally zero border rows: coincident players have grad c = 0) leaves exact zeros here, while the
    // reference's SVD returns those singular values at round-off level, eps * sigma_max > atol = 1e-20, and counts them as rank: its
    // null.mat then has c - min(r, c) columns (test/active_set/active_set_methods.jl:123 expects exactly that).  Same count here —
    // the first c - min(r, c) of the c - rank null vectors — whenever atol is below what an SVD can resolve.
    const int dim = (rank < tmax && atol < 2.220446049250313e-16 * amax) ? c - tmax : c - rank;
    if (tid == 0) dim_out[inst] = dim;
    const int nd = dim < max_dim ? dim : max_dim;
    double* NV = null_out + (size_t)inst * max_dim * Sh;
    for (int q = tid; q < nd * Sh; q += nt) NV[q] = 0.0;
    __syncthreads();
    for (int d = 0; d < nd; d++) {
      const int f = rank + d;
      double* x = NV + (size_t)d * Sh;
      if (tid == 0) x[cols[cperm[f]]] = 1.0;
      for (int p = tid; p < rank; p += nt) x[cols[cperm[p]]] = -M[(size_t)p * c + f] / M[(size_t)p * c + p];
    }
    __syncthreads();
    // orthonormal basis: modified Gram–Schmidt, two passes
    for (int d = 0; d < nd; d++) {
      double* x = NV + (size_t)d * Sh;
      for (int pass = 0; pass < 2; pass++) {
        for (int u = 0; u < d; u++) {
          const double* y = NV + (size_t)u * Sh;
          double part = 0.0;
          for (int q = tid; q < Sh; q += nt) part += x[q] * y[q];
          const double dot = as_block_sum(part, sbuf);
          for (int q = tid; q < Sh; q += nt) x[q] -= dot * y[q];
          __syncthreads();
        }
      }
      double part = 0.0;
      for (int q = tid; q < Sh; q += nt) part += x[q] * x[q];
      const double nrm = sqrt(as_block_sum(part, sbuf));
      if (nrm > 0.0) for (int q = tid; q < Sh; q += nt) x[q] /= nrm;
      __syncthreads();
    }
  }
}

}  // namespace agb
