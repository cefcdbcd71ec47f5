// warpsim.cpp — scheduler of the test-only SIMT emulator (see warpsim.h).
#include "warpsim.h"

wsim_dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace wsim {

Block* g_block = nullptr;
bool g_reverse = false;
static const size_t kStack = 512 * 1024;

void set_reverse(bool r) { g_reverse = r; }

static void trampoline() {
  g_block->body();
  g_block->cur->state = DONE;
  swapcontext(&g_block->cur->ctx, &g_block->sched);
}

static void run_fiber(Block& B, Fiber& f) {
  B.cur = &f;
  threadIdx.x = f.tid;
  f.state = RUNNABLE;
  swapcontext(&B.sched, &f.ctx);
}

static void run_block(Block& B) {
  const unsigned nwarps = (B.nthreads + 31) / 32;
  long stall = 0;
  for (;;) {
    bool all_done = true, all_at_block = true, progressed = false;
    for (unsigned w = 0; w < nwarps; w++) {
      const unsigned lo = w * 32, hi = lo + 32 < B.nthreads ? lo + 32 : B.nthreads;
      for (;;) {  // rounds of this warp until every live lane waits at a block barrier (or spins on another warp)
        bool ran = false;
        for (unsigned k = 0; k < hi - lo; k++) {
          Fiber& f = B.fibers[g_reverse ? hi - 1 - k : lo + k];
          if (f.state == RUNNABLE || f.state == SPIN) { const int was = f.state; run_fiber(B, f); if (!(was == SPIN && f.state == SPIN)) ran = true; }
        }
        if (ran) progressed = true;
        int n_warp = 0, n_block = 0, n_spin = 0, n_live = 0;
        for (unsigned t = lo; t < hi; t++) {
          const int s = B.fibers[t].state;
          if (s == DONE) continue;
          n_live++;
          n_warp += s == AT_WARP; n_block += s == AT_BLOCK; n_spin += s == SPIN;
        }
        if (n_live == 0) break;
        if (n_warp == n_live) { for (unsigned t = lo; t < hi; t++) if (B.fibers[t].state == AT_WARP) B.fibers[t].state = RUNNABLE; continue; }
        if (n_block == n_live) break;
        if (n_spin > 0) { if (ran) continue; break; }  // spinning lanes made no progress: let the other warps run
        fprintf(stderr, "warpsim: divergent synchronisation in block %u warp %u (%d at warp barrier, %d at block barrier, %d live)\n",
                blockIdx.x, w, n_warp, n_block, n_live);
        abort();
      }
    }
    for (auto& f : B.fibers) { if (f.state != DONE) { all_done = false; if (f.state != AT_BLOCK) all_at_block = false; } }
    if (all_done) return;
    if (all_at_block) { for (auto& f : B.fibers) if (f.state == AT_BLOCK) f.state = RUNNABLE; stall = 0; continue; }
    if (!progressed && ++stall > 1000) { fprintf(stderr, "warpsim: deadlock in block %u\n", blockIdx.x); abort(); }
  }
}

void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()>& body) {
  gridDim.x = grid; blockDim.x = block;
  Block B;
  B.nthreads = block;
  B.body = body;
  B.fibers.resize(block);
  B.slots.assign((size_t)((block + 31) / 32) * 2 * 32, 0);
  std::vector<unsigned char> dyn(smem_bytes + 64);
  B.dyn = reinterpret_cast<unsigned char*>(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
  for (auto& f : B.fibers) f.stack = (char*)malloc(kStack);
  g_block = &B;
  for (unsigned b = 0; b < grid; b++) {
    blockIdx.x = b;
    memset(B.dyn, 0xFF, smem_bytes);
    B.bor[0] = B.bor[1] = B.bor[2] = 0;
    for (unsigned t = 0; t < block; t++) {
      Fiber& f = B.fibers[t];
      f.state = RUNNABLE; f.tid = t; f.wpar = 0; f.bcall = 0;
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack;
      f.ctx.uc_stack.ss_size = kStack;
      f.ctx.uc_link = nullptr;
      makecontext(&f.ctx, trampoline, 0);
    }
    run_block(B);
  }
  for (auto& f : B.fibers) free(f.stack);
  g_block = nullptr;
}

}  // namespace wsim
