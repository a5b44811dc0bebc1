// tests/emu/emu.cpp — TEST-ONLY cooperative-fiber CTA emulator (see cuda_runtime.h in this directory).
#include "cuda_runtime.h"

emu_dim3 threadIdx, blockIdx, gridDim, blockDim;

namespace emu {
static State g_state;
State& st() { return g_state; }
static const size_t kStack = 256 * 1024;

static void trampoline() {
  State& s = st();
  (*s.body)();
  s.f[s.cur].wait = DONE;
  swapcontext(&s.f[s.cur].ctx, &s.sched);
}

void yield(int wait) {
  State& s = st();
  s.f[s.cur].wait = wait;
  swapcontext(&s.f[s.cur].ctx, &s.sched);
}

void launch(int grid, int block, size_t smem_bytes, const std::function<void()>& body) {
  State& s = st();
  s.body = &body;
  gridDim = {(unsigned)grid, 1, 1};
  blockDim = {(unsigned)block, 1, 1};
  std::vector<char> stacks((size_t)block * kStack);
  std::vector<double> smem(smem_bytes / sizeof(double) + 2);
  s.smem = smem.data();
  s.f.assign(block, Fiber());
  for (int bx = 0; bx < grid; bx++) {
    blockIdx = {(unsigned)bx, 0, 0};
    for (int t = 0; t < block; t++) {
      Fiber& f = s.f[t];
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = stacks.data() + (size_t)t * kStack;
      f.ctx.uc_stack.ss_size = kStack;
      f.ctx.uc_link = &s.sched;
      f.wait = RUN; f.pred = 1;
      makecontext(&f.ctx, trampoline, 0);
    }
    for (;;) {
      bool ran = false;
      for (int t = 0; t < block; t++) {
        if (s.f[t].wait != RUN) continue;
        ran = true;
        s.cur = t;
        threadIdx = {(unsigned)t, 0, 0};
        swapcontext(&s.sched, &s.f[t].ctx);
      }
      int done = 0, atblock = 0;
      for (int t = 0; t < block; t++) { done += s.f[t].wait == DONE; atblock += s.f[t].wait == BLOCK_BAR; }
      if (done == block) break;
      bool released = false;
      if (atblock > 0 && atblock + done == block) {
        int a = 1;
        for (int t = 0; t < block; t++) if (s.f[t].wait == BLOCK_BAR) a &= s.f[t].pred;
        s.and_result = a;
        for (int t = 0; t < block; t++) if (s.f[t].wait == BLOCK_BAR) s.f[t].wait = RUN;
        released = true;
      } else {
        for (int w = 0; w * 32 < block; w++) {
          int hi = (w + 1) * 32 < block ? (w + 1) * 32 : block, wd = 0, ww = 0;
          for (int t = w * 32; t < hi; t++) { wd += s.f[t].wait == DONE; ww += s.f[t].wait == WARP_BAR; }
          if (ww > 0 && ww + wd == hi - w * 32) {
            for (int t = w * 32; t < hi; t++) if (s.f[t].wait == WARP_BAR) s.f[t].wait = RUN;
            released = true;
          }
        }
      }
      if (!released && !ran) { fprintf(stderr, "emu: deadlock (divergent barrier) in block %d\n", bx); abort(); }
    }
  }
  s.smem = nullptr;
}
}  // namespace emu
