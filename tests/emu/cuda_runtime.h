// tests/emu/cuda_runtime.h — TEST-ONLY shim: lets g++ compile algames.jl_b200/csrc/agb_capi.cu into a CPU
// functional model of the CUDA kernels (one CTA = 128 cooperative fibers, barriers and warp shuffles emulated).
// It exists so the kernel logic can be checked against the oracle in the CPU-only test tier; it is never built
// into, loaded by, or reachable from the shipped package (libalgames_b200.so is always the nvcc build and has
// no CPU fallback).
#pragma once
#ifndef AGB_EMULATE
#error "the emulation shim must only be used with -DAGB_EMULATE"
#endif
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <functional>
#include <vector>
#include <cstdio>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(x)
#define __shared__ static      /* CTAs run one at a time on the emulator: a function-local static is the block's shared memory */

struct emu_dim3 { unsigned x, y, z; };
struct alignas(16) double2 { double x, y; };
inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
extern emu_dim3 threadIdx, blockIdx, gridDim, blockDim;

namespace emu {
enum Wait { RUN = 0, BLOCK_BAR, WARP_BAR, DONE };
struct Fiber { ucontext_t ctx; int wait; int pred; char* stack; };
struct State {
  ucontext_t sched;
  std::vector<Fiber> f;
  int cur = 0;
  int and_result = 1;
  unsigned long long slots[64][32];
  double* smem = nullptr;
  const std::function<void()>* body = nullptr;
};
State& st();
void yield(int wait);
void launch(int grid, int block, size_t smem_bytes, const std::function<void()>& body);
}  // namespace emu

#define AGB_DYN_SMEM(name) double* name = emu::st().smem
#define AGB_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch((int)(grid), (int)(block), (size_t)(smem), [&]() { kern(__VA_ARGS__); })

template <class T> inline T __ldg(const T* p) { return *p; }
inline void __syncthreads() { emu::st().f[emu::st().cur].pred = 1; emu::yield(emu::BLOCK_BAR); }
inline int __syncthreads_and(int p) { emu::st().f[emu::st().cur].pred = p ? 1 : 0; emu::yield(emu::BLOCK_BAR); return emu::st().and_result; }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::yield(emu::WARP_BAR); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  emu::State& s = emu::st();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long raw = 0; memcpy(&raw, &v, sizeof(T));
  s.slots[warp][lane] = raw;
  emu::yield(emu::WARP_BAR);
  raw = s.slots[warp][src & 31];
  emu::yield(emu::WARP_BAR);
  T r; memcpy(&r, &raw, sizeof(T));
  return r;
}
inline int __any_sync(unsigned, int pred) {
  emu::State& s = emu::st();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  s.slots[warp][lane] = pred ? 1ull : 0ull;
  emu::yield(emu::WARP_BAR);
  int any = 0;
  for (int l = 0; l < 32; l++) any |= (int)s.slots[warp][l];
  emu::yield(emu::WARP_BAR);
  return any;
}
template <class T> inline T __shfl_xor_sync(unsigned m, T v, int x) { return __shfl_sync(m, v, (int)((threadIdx.x & 31) ^ x)); }

// ---- the slice of the CUDA runtime API agb_capi.cu uses, on host memory ---------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97, cudaFuncAttributeMaxDynamicSharedMemorySize = 8,
       cudaEventDisableTiming = 2, cudaDevAttrMultiProcessorCount = 16, cudaErrorPeerAccessAlreadyEnabled = 704,
       cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
// AGB_EMU_DEVICES=n makes the shim report n (identical, host-memory) devices, for the in-process multi-GPU gather tests
inline cudaError_t cudaGetDeviceCount(int* n) { const char* e = getenv("AGB_EMU_DEVICES"); *n = e ? atoi(e) : 1; if (*n < 1) *n = 1; return 0; }
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return 801; }      // no inter-process mapping on the shim
inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return 801; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return 0; }
inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 1; return 0; }
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return 0; }
inline cudaError_t cudaEventCreateWithFlags(void** e, unsigned) { *e = nullptr; return 0; }
inline cudaError_t cudaStreamWaitEvent(void*, void*, unsigned) { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 232448; return 0; }
inline cudaError_t cudaFuncSetAttribute(const void*, int, int) { return 0; }
template <class F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaMalloc(void** p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t) { memset(p, v, bytes); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t bytes, cudaMemcpyKind, cudaStream_t) { memmove(d, s, bytes); return 0; }
inline cudaError_t cudaMemcpyPeerAsync(void* d, int, const void* s, int, size_t bytes, cudaStream_t) { memmove(d, s, bytes); return 0; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < height; r++) memmove((char*)d + r * dpitch, (const char*)s + r * spitch, width);
  return 0;
}
