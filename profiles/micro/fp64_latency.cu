// Single-warp FP64 issue behaviour on B200 (sm_100a): dependent-chain latency and independent-stream throughput of
// DFMA / DMUL / DADD, MUFU.RCP64H, a shared-memory store→load round trip and a 64-bit shuffle, timed with clock64() inside
// one warp.  Variants: 1 warp per SM; 4 warps with ids 0..3 (one per scheduler, if warp w sits on scheduler w mod 4); 4
// warps with ids 0,4,8,12 (all on one scheduler under that mapping).  Build: nvcc -arch=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

#define N_OPS 512

template <int ILP, int KIND>   // KIND 0 dfma, 1 dmul, 2 dadd
__device__ __forceinline__ long long run_fp(double* sink, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_OPS / 8; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        if (KIND == 0) x[i] = fma(x[i], a, b);
        else if (KIND == 1) x[i] = x[i] * a;
        else x[i] = x[i] + b;
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  if (s == 123.456) *sink = s;
  return t1 - t0;
}

__device__ __forceinline__ long long run_rcp(double* sink) {
  double x = 1.0 + threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_OPS / 8; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
  }
  long long t1 = clock64();
  if (x == 123.456) *sink = x;
  return t1 - t0;
}

__device__ __forceinline__ long long run_smem(double* sink) {
  __shared__ double buf[64];
  double x = threadIdx.x;
  const int lane = threadIdx.x & 31;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_OPS / 8; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) { buf[lane] = x; __syncwarp(); x = buf[(lane + 1) & 31] ; __syncwarp(); }
  }
  long long t1 = clock64();
  if (x == 123.456) *sink = x;
  return t1 - t0;
}

__device__ __forceinline__ long long run_shfl(double* sink) {
  double x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_OPS / 8; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
  }
  long long t1 = clock64();
  if (x == 123.456) *sink = x;
  return t1 - t0;
}

__global__ void bench(long long* out, double* sink, unsigned warp_mask, double a, double b) {
  const int w = threadIdx.x >> 5;
  if (!((warp_mask >> w) & 1u)) return;
  long long r[16];
  r[0] = run_fp<1, 0>(sink, a, b);  r[1] = run_fp<2, 0>(sink, a, b);  r[2] = run_fp<4, 0>(sink, a, b);  r[3] = run_fp<8, 0>(sink, a, b);
  r[4] = run_fp<1, 1>(sink, a, b);  r[5] = run_fp<4, 1>(sink, a, b);  r[6] = run_fp<1, 2>(sink, a, b);  r[7] = run_fp<4, 2>(sink, a, b);
  r[8] = run_rcp(sink); r[9] = run_smem(sink); r[10] = run_shfl(sink); r[11] = run_fp<16, 0>(sink, a, b);
  if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) for (int k = 0; k < 12; k++) out[w * 16 + k] = r[k];
}

int main() {
  long long* out; double* sink;
  cudaMalloc(&out, 64 * 16 * sizeof(long long)); cudaMalloc(&sink, 8);
  const char* names[12] = {"dfma chain (ILP1)", "dfma ILP2", "dfma ILP4", "dfma ILP8", "dmul chain", "dmul ILP4", "dadd chain", "dadd ILP4",
                           "rcp.approx.f64 chain", "smem st->ld (+2 syncwarp)", "shfl.f64 chain", "dfma ILP16"};
  const int ilp[12] = {1, 2, 4, 8, 1, 4, 1, 4, 1, 1, 1, 16};
  struct { const char* what; unsigned mask; int threads; int grid; } cfg[] = {
    {"1 warp on the SM", 0x1u, 32, 1}, {"4 warps, ids 0-3", 0xFu, 128, 1}, {"4 warps, ids 0,4,8,12", 0x1111u, 512, 1},
    {"8 warps ids 0-7", 0xFFu, 256, 1}, {"16 warps ids 0-15", 0xFFFFu, 512, 1}, {"1 warp per SM on all SMs", 0x1u, 32, 148}};
  for (auto& c : cfg) {
    long long h[64 * 16];
    cudaMemset(out, 0, sizeof h);
    bench<<<c.grid, c.threads>>>(out, sink, c.mask, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    bench<<<c.grid, c.threads>>>(out, sink, c.mask, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    printf("== %s (cycles per warp-instruction, warp 0 / last warp)\n", c.what);
    int lastw = 0; for (int w = 0; w < 32; w++) if ((c.mask >> w) & 1u) lastw = w;
    for (int k = 0; k < 12; k++)
      printf("  %-28s %7.2f  %7.2f\n", names[k], (double)h[k] / (N_OPS * ilp[k]), (double)h[lastw * 16 + k] / (N_OPS * ilp[k]));
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
