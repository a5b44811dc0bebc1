// Stand-alone check of the 1-D bulk-copy helpers of agb_solver.cuh (cp.async.bulk + mbarrier) on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define AGB_FULL 0xffffffffu
__device__ __forceinline__ unsigned smem32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem32(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(void* bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem32(dst)), "l"(src), "r"(bytes), "r"(smem32(bar)) : "memory");
}
__device__ __forceinline__ int mbar_wait(void* bar, unsigned parity) {
  unsigned ok; int spins = 0;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem32(bar)), "r"(parity) : "memory");
  } while (!ok && ++spins < 1000000);
  return ok ? spins : -1;
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) { asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem32(src)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_commit_wait() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void k(const double* src, double* dst, int n, int* status, int bar_slot) {
  extern __shared__ __align__(16) double sm[];
  double* bar = sm + bar_slot;
  unsigned phase = 0;
  if (threadIdx.x == 0) mbar_init(bar);
  __syncthreads();
  for (int rep = 0; rep < 3; rep++) {
    if (threadIdx.x == 0) { mbar_expect(bar, n * 8); bulk_g2s(sm, src + (size_t)blockIdx.x * n, n * 8, bar); }
    const int sp = mbar_wait(bar, phase); phase ^= 1;
    if (threadIdx.x == 0) status[blockIdx.x * 3 + rep] = sp;
    if (sp < 0) return;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] += 1.0;
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) { bulk_s2g(dst + (size_t)blockIdx.x * n, sm, n * 8); bulk_commit_wait(); }
    __syncthreads();
  }
}
int main() {
  const int n = 1404, B = 8, slot = 7000;        // 11232 bytes, like Λ of config B; barrier far behind the data like red[47]
  double *s, *d; int* st;
  cudaMalloc(&s, B * n * 8); cudaMalloc(&d, B * n * 8); cudaMalloc(&st, B * 3 * 4);
  double* h = new double[B * n]; for (int i = 0; i < B * n; i++) h[i] = i;
  cudaMemcpy(s, h, B * n * 8, cudaMemcpyHostToDevice); cudaMemset(d, 0, B * n * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 57072);
  k<<<B, 128, 57072>>>(s, d, n, st, slot);
  cudaError_t e = cudaDeviceSynchronize();
  int hs[B * 3]; cudaMemcpy(hs, st, sizeof hs, cudaMemcpyDeviceToHost);
  double* r = new double[B * n]; cudaMemcpy(r, d, B * n * 8, cudaMemcpyDeviceToHost);
  int bad = 0; for (int i = 0; i < B * n; i++) if (r[i] != h[i] + 1.0) bad++;
  printf("cuda: %s; spins %d %d %d ... ; mismatches %d of %d\n", cudaGetErrorString(e), hs[0], hs[1], hs[2], bad, B * n);
  return 0;
}
