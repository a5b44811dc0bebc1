// Accuracy of rcp.approx.ftz.f64 (MUFU.RCP64H) on sm_100a: max relative error over mantissas in [1, 2) and a few exponents.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(double* out, int n) {
  double worst = 0, worst1 = 0, worst2 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    for (int e = -3; e <= 3; e += 3) {
      double x = ldexp(1.0 + (double)i / n, e * 100);
      double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
      worst = fmax(worst, fabs(r * x - 1.0));
      double e1 = fma(-x, r, 1.0), r1 = fma(r, e1, r);
      worst1 = fmax(worst1, fabs(fma(r1, x, -1.0)));
      double e2 = fma(-x, r1, 1.0), r2 = fma(r1, e2, r1);
      worst2 = fmax(worst2, fabs(fma(r2, x, -1.0)));
    }
  }
  atomicMax((unsigned long long*)&out[0], __double_as_longlong(worst));
  atomicMax((unsigned long long*)&out[1], __double_as_longlong(worst1));
  atomicMax((unsigned long long*)&out[2], __double_as_longlong(worst2));
}
int main() {
  double* d; cudaMalloc(&d, 24); cudaMemset(d, 0, 24);
  k<<<296, 256>>>(d, 1 << 24);
  double h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  printf("rcp.approx.ftz.f64 max rel err %.3e (2^%.1f); after 1 Newton step %.3e; after 2 %.3e\n", h[0], log2(h[0]), h[1], h[2]);
  return 0;
}
