// warpsim.h — TEST INFRASTRUCTURE, not product.  A tiny SIMT emulator that lets the device code in
// judo_b200/csrc/*.cuh compile with g++ (-DB2_HOST_SIM) and run on the CPU, one thread block at a time, so that the
// kernels' logic can be checked against the oracle in the `-m "not gpu"` test tier (this container has no GPU).
//
// * every CUDA thread of a block is a ucontext fiber with its own stack; fibers switch only at synchronisation points
//   (__syncwarp, __syncthreads, warp collectives, mbarrier waits), exactly where CUDA allows lanes to exchange data;
// * warp collectives (__shfl_*_sync, __ballot_sync, __all_sync) exchange values through per-warp slots around a warp barrier;
// * lanes of a warp can be run in FORWARD or REVERSE order between barriers: a missing __syncwarp / __syncthreads shows up as a
//   result that depends on the order (tests run both and compare);
// * dynamic shared memory is filled with 0xFF bytes (NaNs) so reads of uninitialised shared memory surface.
//
// Nothing under judo_b200/ includes or links this file; the shipped library is built by nvcc without B2_HOST_SIM.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <string>
#include <vector>

// (after every standard header: libstdc++ itself spells attributes with these names)
#define __device__
#define __constant__ static const
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static
using std::max;
using std::min;


struct wsim_dim3 { unsigned x = 1, y = 1, z = 1; };
extern wsim_dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace wsim {

enum { RUNNABLE = 0, AT_WARP = 1, AT_BLOCK = 2, DONE = 3, SPIN = 4 };

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  int state = RUNNABLE;
  unsigned tid = 0;
  int wpar = 0;   // parity of this lane's next warp collective
  long bcall = 0; // index of this thread's next __syncthreads_or
};

struct Block {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  Fiber* cur = nullptr;
  unsigned nthreads = 0;
  std::vector<uint64_t> slots;  // [warp][parity][lane]
  int bor[3] = {0, 0, 0};
  unsigned char* dyn = nullptr;
  std::function<void()> body;
};

extern Block* g_block;
extern bool g_reverse;

inline void yield(int state) {
  Fiber* f = g_block->cur;
  f->state = state;
  swapcontext(&f->ctx, &g_block->sched);
}
inline void warp_barrier() { yield(AT_WARP); }
inline void block_barrier() { yield(AT_BLOCK); }
inline void spin_yield() { yield(SPIN); }
inline unsigned char* dyn_smem() { return g_block->dyn; }

template <class T>
inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <class T>
inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

// all lanes of the calling warp publish `bits`; returns the base of the 32 published slots (valid until the next-but-one collective)
inline const uint64_t* exchange(uint64_t bits) {
  Fiber* f = g_block->cur;
  const unsigned warp = f->tid >> 5, lane = f->tid & 31;
  uint64_t* base = g_block->slots.data() + ((size_t)warp * 2 + f->wpar) * 32;
  base[lane] = bits;
  f->wpar ^= 1;
  warp_barrier();
  return base;
}
inline unsigned live_mask() {  // lanes of the calling warp that exist and have not exited
  Fiber* f = g_block->cur;
  const unsigned w0 = (f->tid >> 5) << 5;
  unsigned m = 0;
  for (unsigned l = 0; l < 32 && w0 + l < g_block->nthreads; l++)
    if (g_block->fibers[w0 + l].state != DONE) m |= 1u << l;
  return m;
}

void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()>& body);
void set_reverse(bool r);

}  // namespace wsim

// ------------------------------------------------------------------ CUDA intrinsics used by judo_b200/csrc
inline void __syncwarp(unsigned = 0xffffffffu) { wsim::warp_barrier(); }
inline void __syncthreads() { wsim::block_barrier(); }
inline int __syncthreads_or(int pred) {
  wsim::Fiber* f = wsim::g_block->cur;
  const long k = f->bcall++;
  wsim::g_block->bor[(k + 1) % 3] = 0;  // every thread has finished reading call k-2's slot before anyone reaches call k
  if (pred) wsim::g_block->bor[k % 3] = 1;
  wsim::block_barrier();
  return wsim::g_block->bor[k % 3];
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src) { const uint64_t* s = wsim::exchange(wsim::to_bits(v)); return wsim::from_bits<T>(s[src & 31]); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int m) {
  const unsigned lane = wsim::g_block->cur->tid & 31;
  const uint64_t* s = wsim::exchange(wsim::to_bits(v));
  return wsim::from_bits<T>(s[(lane ^ m) & 31]);
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned d) {
  const unsigned lane = wsim::g_block->cur->tid & 31;
  const uint64_t* s = wsim::exchange(wsim::to_bits(v));
  return lane >= d ? wsim::from_bits<T>(s[lane - d]) : v;
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned d) {
  const unsigned lane = wsim::g_block->cur->tid & 31;
  const uint64_t* s = wsim::exchange(wsim::to_bits(v));
  return lane + d < 32 ? wsim::from_bits<T>(s[lane + d]) : v;
}
inline unsigned __ballot_sync(unsigned, int pred) {
  const uint64_t* s = wsim::exchange(pred ? 1 : 0);
  const unsigned live = wsim::live_mask();
  unsigned m = 0;
  for (int l = 0; l < 32; l++) if ((live >> l & 1) && s[l]) m |= 1u << l;
  return m;
}
inline int __all_sync(unsigned, int pred) {
  const uint64_t* s = wsim::exchange(pred ? 1 : 0);
  const unsigned live = wsim::live_mask();
  for (int l = 0; l < 32; l++) if ((live >> l & 1) && !s[l]) return 0;
  return 1;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
inline void __threadfence() {}
inline void __threadfence_system() {}
inline void __nanosleep(unsigned) { wsim::spin_yield(); }
template <class T>
inline T __ldcg(const T* p) { return *p; }
template <class T>
inline T __ldg(const T* p) { return *p; }
template <class T, class U>
inline T atomicAdd(T* p, U v) { T old = *p; *p = old + (T)v; return old; }
template <class T, class U>
inline T atomicMax(T* p, U v) { T old = *p; if ((T)v > old) *p = (T)v; return old; }
inline unsigned atomicInc(unsigned* p, unsigned lim) { unsigned old = *p; *p = old >= lim ? 0 : old + 1; return old; }
inline long long clock64() { return 0; }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
inline void sincospi(double x, double* s, double* c) { sincos(3.14159265358979323846 * x, s, c); }
using std::isfinite;

typedef void* cudaStream_t;
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }
