// epilogue.cuh — optimizer update fused into the tail of the rollout kernel.
//
// Every warp reduces the rollouts it owns to a small partial while rewards and knots are still in registers
// (warp shuffles only), publishes it, and takes a ticket; the LAST warp to arrive combines all partials in index
// order (deterministic) and writes either the final nominal knots (single GPU) or this rank's partial for the
// all_gather (multi-GPU).  One kernel launch per plan step.
//
//   MPPI  (judo/optimizers/mppi.py:61-82):  partial = [beta, S, V[KNU]],  beta = min cost,
//         S = sum exp(-(c-beta)/T),  V = sum exp(-(c-beta)/T) * knots;  combine rescales to the global beta.
//   CEM   (judo/optimizers/cem.py:76-92) / PS (ps.py:52-65) / trace elites (controller.py:341):
//         partial = top-k (reward, index) pairs; the final stage picks the global top-k and reads those knots.
#pragma once
#include "common.cuh"

namespace b2 {

enum { EP_NONE = -1, EP_MPPI = 0, EP_CEM = 1, EP_PS = 2 };
constexpr int EP_MAXK = 8;  // max(num_elites, trace elites) the fused path supports; larger -> separate kernels

struct PlanEpilogue {
  int optimizer;          // EP_*
  int k;                  // best rollouts to list (descending, ties: higher index first), 0..EP_MAXK
  int k_cem;              // elites entering CEM's mean/var (<= k)
  int n_trace;            // elites whose sensors go to elite_sens (<= k): the caller's n_elite
  int finalize;           // 1: write nominal/sigma/elite;  0: write rank partials for the exchange
  int index_offset;       // global index of this launch's rollout 0
  double temperature, sigma_min, sigma_max;
  double* warp_mppi;      // scratch [nwarps][2+KNU]
  double* warp_topk;      // scratch [nwarps][k+1][2]   (slot k: PS argmax, ties -> lower index)
  unsigned int* ticket;   // zero-initialised; reset by the last warp
  double* nominal;        // [KNU]  (finalize)
  double* sigma;          // [KNU]  (finalize, CEM)
  double* elite;          // [k] global indices as doubles, -1 padded (finalize)
  double* elite_knots;    // [k][KNU] the elite candidates themselves, or NULL (finalize)
  // elite traces (thread-per-rollout kernels, finalize == 1): with trace_q set every rollout keeps its positions after each step
  // (N, H, NQ) in HBM, and the last warp evaluates the sensors of the k elite rollouts from them into elite_sens (k, H, NS) — what
  // Controller.update_traces reads (controller.py:323-363) without a second launch that re-simulates the elites
  double* trace_q;
  double* elite_sens;
  double* rank_mppi;      // [2+KNU]            (!finalize)
  double* rank_topk;      // CEM: [k][2+KNU]; PS: [1][2+KNU]   (!finalize)
  // peer exchange (finalize == 2, MPPI): the last warp writes this rank's partial straight into every peer's exchange buffer
  // over NVLink (P2P stores + system-scope release flag), waits for the world_size flags of the current epoch in its OWN buffer
  // and finishes the update — the collective is part of the rollout kernel, no NCCL call and no extra launch on the data path.
  int world, rank;
  unsigned long long epoch;
  double* peer[8];        // peer[g]: exchange buffer of rank g mapped into this process (peer[rank] is local)
  unsigned long long* stamps;  // optional (finalize == 2): %globaltimer at [0] kernel entry (block 0), [1] partial published, [2] all peers seen, [3] line-up kernel exit, [4] exchange branch done
};
constexpr int EP_XCHG_FLAGS = 16;                  // u64 flags in front of the partial slots: [0..8) step epochs, [8..16) align epochs
constexpr int EP_XCHG_STRIDE = 2 + 96;             // doubles per rank slot (beta, S, V[<=96])
constexpr size_t EP_XCHG_LL = 8 * EP_XCHG_FLAGS + 2 * 8 * (size_t)EP_XCHG_STRIDE * 8;     // byte offset of the flag-in-data slots (MPPI)
constexpr size_t EP_XCHG_BYTES = EP_XCHG_LL + 2 * 8 * (size_t)EP_XCHG_STRIDE * 16;          // slots double-buffered by epoch parity

// Flag-in-data exchange (the "LL" protocol of NCCL): every double travels as ONE 16-byte store {lo, epoch, hi, epoch}.  NVLink delivers
// 8-byte units atomically, so a reader that sees both epoch words has the value: no fence between data and flag on the writer, no
// separate flag to poll and no second round trip for the data on the reader.
struct __align__(16) LLWord { unsigned lo, f0, hi, f1; };
#ifdef B2_HOST_SIM
__device__ __forceinline__ void ll_store(LLWord* p, double v, unsigned epoch) {
  unsigned long long b; memcpy(&b, &v, 8);
  p->lo = (unsigned)b; p->f0 = epoch; p->hi = (unsigned)(b >> 32); p->f1 = epoch;
}
__device__ __forceinline__ LLWord ll_load(const LLWord* p) { return *p; }
#else
__device__ __forceinline__ void ll_store(LLWord* p, double v, unsigned epoch) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)b), "r"(epoch), "r"((unsigned)(b >> 32)), "r"(epoch) : "memory");
}
__device__ __forceinline__ LLWord ll_load(const LLWord* p) {
  LLWord w;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w.lo), "=r"(w.f0), "=r"(w.hi), "=r"(w.f1) : "l"(p) : "memory");
  return w;
}
#endif
__device__ __forceinline__ bool ll_ready(const LLWord& w, unsigned epoch) { return w.f0 == epoch && w.f1 == epoch; }
__device__ __forceinline__ double ll_value(const LLWord& w) { return __longlong_as_double((long long)(((unsigned long long)w.hi << 32) | w.lo)); }
__device__ __forceinline__ LLWord* ll_slot(double* buf, unsigned long long epoch, int src_rank) {
  return reinterpret_cast<LLWord*>(reinterpret_cast<char*>(buf) + EP_XCHG_LL) + (size_t)((epoch & 1) * 8 + src_rank) * EP_XCHG_STRIDE;
}

#ifdef B2_HOST_SIM
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) { *p = v; }
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) { return *p; }
#else
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
#endif

__device__ __forceinline__ unsigned long long global_ns() {
#ifdef B2_HOST_SIM
  return 0;
#else
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
#endif
}

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
  return v;
}
__device__ __forceinline__ double wmin(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, shfl_xor_d(v, o));
  return v;
}
__device__ __forceinline__ bool better(double ra, long long ia, double rb, long long ib, bool prefer_high) {
  if (ia < 0) return false;
  if (ib < 0) return true;
  if (ra != rb) return ra > rb;
  return prefer_high ? ia > ib : ia < ib;
}
// warp arg-best of (r, i) under `better`; every lane gets the winner
__device__ __forceinline__ void warp_best(double& r, long long& i, bool prefer_high) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double r2 = shfl_xor_d(r, o);
    long long i2 = __shfl_xor_sync(0xffffffffu, i, o);
    if (better(r2, i2, r, i, prefer_high)) { r = r2; i = i2; }
  }
}

// Raise this rank's epoch flag in every peer's buffer (after the partial stores) and wait, bounded, for all world_size flags of
// the epoch in OUR buffer.  Returns false on time-out (a peer never launched): the caller writes NaN instead of hanging the GPU.
__device__ inline bool peer_signal_and_wait(const PlanEpilogue& ep, int lane) {
  __threadfence_system();
  __syncwarp();
  if (ep.stamps && lane == 0) ep.stamps[1] = global_ns();
  if (lane < ep.world) st_release_sys(reinterpret_cast<unsigned long long*>(ep.peer[lane]) + ep.rank, ep.epoch);
  bool ok = true;
  if (lane < ep.world) {
    const unsigned long long* fl = reinterpret_cast<const unsigned long long*>(ep.peer[ep.rank]) + lane;
    int spins = 0;
    while (ld_acquire_sys(fl) < ep.epoch) { __nanosleep(64); if (++spins > (1 << 22)) { ok = false; break; } }
  }
  const bool all_ok = __all_sync(0xffffffffu, ok);
  if (ep.stamps && lane == 0) ep.stamps[2] = global_ns();
  return all_ok;
}

// Line the ranks up WITHOUT exchanging data: raise this rank's align flag in every peer buffer and wait for all of them.  bench.py runs it
// between the L2 flush and the start event so that the timed step measures the exchange, not how differently long the flushes took.
__global__ void exchange_align_kernel(PlanEpilogue ep) {
  const int lane = threadIdx.x & 31;
  __threadfence_system();
  if (lane < ep.world) st_release_sys(reinterpret_cast<unsigned long long*>(ep.peer[lane]) + 8 + ep.rank, ep.epoch);
  if (lane < ep.world) {
    const unsigned long long* fl = reinterpret_cast<const unsigned long long*>(ep.peer[ep.rank]) + 8 + lane;
    int spins = 0;
    while (ld_acquire_sys(fl) < ep.epoch) { __nanosleep(32); if (++spins > (1 << 22)) break; }
  }
  __syncwarp();
  if (ep.stamps && lane == 0) ep.stamps[3] = global_ns();  // [3]: the line-up kernel's exit (bench: start of the launch gap)
}

// Final stage, run by one full warp after all partials are visible.  knots: this launch's (N, KNU) candidates.
// Loads are issued in batches of independent __ldcg's (the stage is one warp deep: L2 latency, not bandwidth, is the cost).
template <int MAXKNU>
__device__ inline void epilogue_final(const PlanEpilogue& ep, int nparts, int KNU, const double* __restrict__ knots,
                                      long long (&elite_local)[EP_MAXK], int& n_elite_local) {
  const int lane = threadIdx.x & 31;
  n_elite_local = 0;
  if (ep.optimizer == EP_MPPI) {
    const int stride = 2 + KNU;
    double b = INFINITY;
    for (int p = lane; p < nparts; p += 32) b = fmin(b, __ldcg(ep.warp_mppi + (size_t)p * stride));
    const double beta = wmin(b);
    double s = 0, v[MAXKNU];
#pragma unroll
    for (int j = 0; j < MAXKNU; j++) v[j] = 0;
    for (int p = lane; p < nparts; p += 32) {
      const double* q = ep.warp_mppi + (size_t)p * stride;
      const double bp = __ldcg(q), sp = __ldcg(q + 1);
      double x[MAXKNU];
#pragma unroll
      for (int j = 0; j < MAXKNU; j++) x[j] = j < KNU ? __ldcg(q + 2 + j) : 0.0;
      const double sc = sp > 0 ? exp(-(bp - beta) / ep.temperature) : 0.0;
      s += sp * sc;
#pragma unroll
      for (int j = 0; j < MAXKNU; j++) v[j] += x[j] * sc;
    }
    const double S = wsum(s);
    if (ep.finalize == 2) {
      // ---- fused exchange: publish [beta, S, V] into every peer's slot for this rank as flag-in-data words (one-way NVLink stores,
      // no fence, no separate flag), then gather the world_size partials from OUR buffer: lane g polls rank g's (beta, S), lane j
      // polls column j of every rank's V — all loads of a poll round are in flight together, so the exchange costs one NVLink
      // store latency plus one local poll round instead of fence + flag + dependent reads.
      double vj[MAXKNU];
#pragma unroll
      for (int j = 0; j < MAXKNU; j++) vj[j] = j < KNU ? wsum(v[j]) : 0.0;
      const unsigned ep32 = (unsigned)ep.epoch;
      for (int g = 0; g < ep.world; g++) {
        LLWord* slot = ll_slot(ep.peer[g], ep.epoch, ep.rank);
        if (lane == 0) ll_store(slot, beta, ep32);
        if (lane == 1) ll_store(slot + 1, S, ep32);
#pragma unroll
        for (int j = 0; j < MAXKNU; j++) if (j < KNU && lane == ((j + 2) & 31)) ll_store(slot + 2 + j, vj[j], ep32);
      }
      if (ep.stamps && lane == 0) ep.stamps[1] = global_ns();
      constexpr int NT = (MAXKNU + 31) / 32;
      LLWord wb, ws, wv[NT][8];
      bool ok = false;
      for (int spins = 0; spins < (1 << 22); spins++) {
        bool ready = true;
        if (lane < ep.world) {
          const LLWord* src = ll_slot(ep.peer[ep.rank], ep.epoch, lane);
          wb = ll_load(src); ws = ll_load(src + 1);
        }
#pragma unroll
        for (int t = 0; t < NT; t++)
#pragma unroll
          for (int g = 0; g < 8; g++)
            if (g < ep.world && lane + 32 * t < KNU) wv[t][g] = ll_load(ll_slot(ep.peer[ep.rank], ep.epoch, g) + 2 + lane + 32 * t);
        if (lane < ep.world) ready = ll_ready(wb, ep32) && ll_ready(ws, ep32);
#pragma unroll
        for (int t = 0; t < NT; t++)
#pragma unroll
          for (int g = 0; g < 8; g++)
            if (g < ep.world && lane + 32 * t < KNU) ready = ready && ll_ready(wv[t][g], ep32);
        if (__all_sync(0xffffffffu, ready)) { ok = true; break; }
        __nanosleep(32);
      }
      if (ep.stamps && lane == 0) ep.stamps[2] = global_ns();
      // ---- combine (lane g holds rank g's beta / S; every rank sums the ranks in the same order: identical nominal everywhere)
      const double bg = lane < ep.world ? ll_value(wb) : INFINITY;
      const double sg = lane < ep.world ? ll_value(ws) : 0.0;
      const double bmin = wmin(bg);
      const double scg = sg > 0 ? exp(-(bg - bmin) / ep.temperature) : 0.0;
      const double Stot = wsum(sg * scg);
      double sc[8];
#pragma unroll
      for (int g = 0; g < 8; g++) sc[g] = __shfl_sync(0xffffffffu, scg, g);
#pragma unroll
      for (int t = 0; t < NT; t++) {
        double acc = 0;
#pragma unroll
        for (int g = 0; g < 8; g++)
          if (g < ep.world && lane + 32 * t < KNU) acc += ll_value(wv[t][g]) * sc[g];
        if (lane + 32 * t < KNU) ep.nominal[lane + 32 * t] = ok ? acc / Stot : __longlong_as_double(0x7ff8000000000000ll);
      }
      if (ep.stamps && lane == 0) ep.stamps[4] = global_ns();
    } else {
      double* out = ep.finalize ? ep.nominal : ep.rank_mppi + 2;
#pragma unroll
      for (int j = 0; j < MAXKNU; j++) {
        if (j < KNU) {
          const double t = wsum(v[j]);
          if (lane == 0) out[j] = ep.finalize ? t / S : t;
        }
      }
      if (!ep.finalize && lane == 0) { ep.rank_mppi[0] = beta; ep.rank_mppi[1] = S; }
    }
  }
  if (ep.k > 0 || ep.optimizer == EP_PS) {
    const int k = ep.k, slots = k + 1;
    // global top-k (descending; ties: higher index first) over nparts*k candidates.  Each lane streams its share of the
    // candidates (independent L2 loads) through a register-resident sorted list, then k warp-wide pops merge the 32 lists.
    double lr[EP_MAXK]; long long li[EP_MAXK];
#pragma unroll
    for (int e = 0; e < EP_MAXK; e++) { lr[e] = -INFINITY; li[e] = -1; }
    const int ncand = nparts * k;
    for (int cnd = lane; cnd < ncand; cnd += 32) {
      const int pidx = cnd / k, slot = cnd - pidx * k;
      const double* q = ep.warp_topk + ((size_t)pidx * slots + slot) * 2;
      double r = __ldcg(q); long long i = (long long)__ldcg(q + 1);
      if (i < 0) continue;
#pragma unroll
      for (int e = 0; e < EP_MAXK; e++) {  // insertion: carry the displaced entry down the list
        if (better(r, i, lr[e], li[e], true)) { const double tr = lr[e]; const long long ti = li[e]; lr[e] = r; li[e] = i; r = tr; i = ti; }
      }
    }
    double er[EP_MAXK]; long long ei[EP_MAXK];
#pragma unroll
    for (int e = 0; e < EP_MAXK; e++) { er[e] = -INFINITY; ei[e] = -1; }
    int ne = 0;
#pragma unroll
    for (int e = 0; e < EP_MAXK; e++) {
      if (e < k) {
        double br = lr[0]; long long bi = li[0];
        warp_best(br, bi, true);
        if (bi >= 0) {
          er[e] = br; ei[e] = bi; ne = e + 1;
          if (li[0] == bi) {  // my head won: pop it
#pragma unroll
            for (int t = 0; t + 1 < EP_MAXK; t++) { lr[t] = lr[t + 1]; li[t] = li[t + 1]; }
            lr[EP_MAXK - 1] = -INFINITY; li[EP_MAXK - 1] = -1;
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < EP_MAXK; e++) elite_local[e] = ei[e];
    n_elite_local = ne;
    // PS: first maximum (ties -> lower index)
    double pr = -INFINITY; long long pi = -1;
    if (ep.optimizer == EP_PS) {
      for (int p = lane; p < nparts; p += 32) {
        const double* q = ep.warp_topk + ((size_t)p * slots + k) * 2;
        double r = __ldcg(q); long long i = (long long)__ldcg(q + 1);
        if (better(r, i, pr, pi, false)) { pr = r; pi = i; }
      }
      warp_best(pr, pi, false);
    }
    if (ep.finalize == 2 && ep.optimizer != EP_MPPI) {
      // ---- fused exchange of the rank's best candidates: kout x [reward, global index, knots] into every peer's slot
      const int kout = ep.optimizer == EP_CEM ? ep.k_cem : 1, stride = 2 + KNU;
      for (int g = 0; g < ep.world; g++) {
        double* slot = ep.peer[g] + EP_XCHG_FLAGS + (size_t)((ep.epoch & 1) * 8 + ep.rank) * EP_XCHG_STRIDE;
        for (int e = 0; e < kout; e++) {
          const bool have = ep.optimizer == EP_CEM ? e < ne : pi >= 0;
          const long long loc = ep.optimizer == EP_CEM ? ei[e] : pi;
          const double rr = ep.optimizer == EP_CEM ? er[e] : pr;
          if (lane == 0) { slot[e * stride] = have ? rr : -INFINITY; slot[e * stride + 1] = have ? (double)(loc + ep.index_offset) : -1.0; }
          for (int j = lane; j < KNU; j += 32) slot[e * stride + 2 + j] = have ? __ldcg(knots + (size_t)loc * KNU + j) : 0.0;
        }
      }
      const bool ok = peer_signal_and_wait(ep, lane);
      // ---- global selection over world*kout candidates held in OUR buffer (<= 64: two per lane)
      const volatile double* own = ep.peer[ep.rank] + EP_XCHG_FLAGS + (size_t)(ep.epoch & 1) * 8 * EP_XCHG_STRIDE;
      const int ncand = ep.world * kout;
      const bool hi_first = ep.optimizer == EP_CEM;
      double cr[2]; long long ci[2]; int cl[2];
#pragma unroll
      for (int t = 0; t < 2; t++) {
        const int cnd = lane + 32 * t;
        cr[t] = -INFINITY; ci[t] = -1; cl[t] = -1;
        if (cnd < ncand) {
          const int g = cnd / kout, e = cnd - g * kout;
          cl[t] = g * EP_XCHG_STRIDE + e * stride;
          cr[t] = own[cl[t]]; ci[t] = (long long)own[cl[t] + 1];
        }
      }
      int wloc[EP_MAXK];
      int nw = 0;
      bool used[2] = {false, false};
      for (int e = 0; e < kout; e++) {
        double br = -INFINITY; long long bi = -1;
#pragma unroll
        for (int t = 0; t < 2; t++) if (!used[t] && better(cr[t], ci[t], br, bi, hi_first)) { br = cr[t]; bi = ci[t]; }
        warp_best(br, bi, hi_first);
        if (bi < 0) break;
        int myloc = -1;
#pragma unroll
        for (int t = 0; t < 2; t++) if (!used[t] && ci[t] == bi) { used[t] = true; myloc = cl[t]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) myloc = max(myloc, __shfl_xor_sync(0xffffffffu, myloc, o));
        wloc[nw++] = myloc;
      }
      const double qnan = __longlong_as_double(0x7ff8000000000000ll);
      for (int j = lane; j < KNU; j += 32) {
        double mean = 0;
        for (int e = 0; e < nw; e++) mean += own[wloc[e] + 2 + j];
        mean = nw ? mean / nw : 0.0;
        double var = 0;
        for (int e = 0; e < nw; e++) { const double d = own[wloc[e] + 2 + j] - mean; var += d * d; }
        var = nw ? var / nw : 0.0;
        ep.nominal[j] = ok ? mean : qnan;
        if (ep.optimizer == EP_CEM && ep.sigma) ep.sigma[j] = ok ? fmin(fmax(sqrt(var), ep.sigma_min), ep.sigma_max) : qnan;
      }
    } else if (ep.finalize) {
      if (ep.elite) for (int e = lane; e < k; e += 32) ep.elite[e] = e < ne ? (double)(ei[e] + ep.index_offset) : -1.0;
      if (ep.elite_knots)
        for (int e = 0; e < ne; e++)
          for (int j = lane; j < KNU; j += 32) ep.elite_knots[(size_t)e * KNU + j] = __ldcg(knots + (size_t)ei[e] * KNU + j);
      if (ep.optimizer == EP_CEM) {
        const int nc = min(ne, ep.k_cem);
        for (int j = lane; j < KNU; j += 32) {
          double mean = 0;
          for (int e = 0; e < nc; e++) mean += __ldcg(knots + (size_t)ei[e] * KNU + j);
          mean = nc ? mean / nc : 0.0;
          double var = 0;
          for (int e = 0; e < nc; e++) { double d = __ldcg(knots + (size_t)ei[e] * KNU + j) - mean; var += d * d; }
          var = nc ? var / nc : 0.0;
          ep.nominal[j] = mean;
          ep.sigma[j] = fmin(fmax(sqrt(var), ep.sigma_min), ep.sigma_max);
        }
      } else if (ep.optimizer == EP_PS) {
        for (int j = lane; j < KNU; j += 32) ep.nominal[j] = pi >= 0 ? __ldcg(knots + (size_t)pi * KNU + j) : 0.0;
      }
    } else if (ep.optimizer == EP_PS) {  // rank partial: 1 x [reward, global index, knots]
      double* o = ep.rank_topk;
      if (lane == 0) { o[0] = pi >= 0 ? pr : -INFINITY; o[1] = pi >= 0 ? (double)(pi + ep.index_offset) : -1.0; }
      for (int j = lane; j < KNU; j += 32) o[2 + j] = pi >= 0 ? __ldcg(knots + (size_t)pi * KNU + j) : 0.0;
    } else {  // rank partial: k x [reward, global index, knots]
      const int stride = 2 + KNU;
      for (int e = 0; e < k; e++) {
        double* o = ep.rank_topk + (size_t)e * stride;
        if (lane == 0) { o[0] = e < ne ? er[e] : -INFINITY; o[1] = e < ne ? (double)(ei[e] + ep.index_offset) : -1.0; }
        for (int j = lane; j < KNU; j += 32) o[2 + j] = e < ne ? __ldcg(knots + (size_t)ei[e] * KNU + j) : 0.0;
      }
    }
  }
}

// Publish this warp's partial (already written by the caller), take a ticket, and run the final stage in the last warp.
// Returns true in the warp that ran the final stage; elite_local / n_elite_local then hold the LOCAL indices of the listed elites.
template <int MAXKNU>
__device__ inline bool epilogue_commit(const PlanEpilogue& ep, int nparts, int KNU, const double* __restrict__ knots,
                                       long long (&elite_local)[EP_MAXK], int& n_elite_local) {
  __threadfence();
  __syncwarp();
  unsigned t = 0;
  if ((threadIdx.x & 31) == 0) t = atomicAdd(ep.ticket, 1u);
  t = __shfl_sync(0xffffffffu, t, 0);
  if (t == (unsigned)nparts - 1) {
    __threadfence();
    epilogue_final<MAXKNU>(ep, nparts, KNU, knots, elite_local, n_elite_local);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) *ep.ticket = 0;
    return true;
  }
  return false;
}

// Thread-per-rollout flavour: lane owns rollout `n` (valid or not) with reward r and knots kn[0..KNU).
template <int MAXKNU>
__device__ inline bool epilogue_thread_per_rollout(const PlanEpilogue& ep, bool valid, int n_local, double r, const double (&kn)[MAXKNU],
                                                   int KNU, int warp_global, int nwarps, const double* __restrict__ knots,
                                                   long long (&elite_local)[EP_MAXK], int& n_elite_local) {
  const int lane = threadIdx.x & 31;
  if (ep.optimizer == EP_MPPI) {
    const double c = valid ? -r : INFINITY;
    const double beta = wmin(c);
    const double w = valid ? exp(-(c - beta) / ep.temperature) : 0.0;
    const double S = wsum(w);
    double* out = ep.warp_mppi + (size_t)warp_global * (2 + KNU);
    if (lane == 0) { out[0] = beta; out[1] = S; }
#pragma unroll
    for (int j = 0; j < MAXKNU; j++) {
      if (j < KNU) {
        double v = wsum(w * kn[j]);
        if (lane == 0) out[2 + j] = v;
      }
    }
  }
  if (ep.k > 0 || ep.optimizer == EP_PS) {
    const int slots = ep.k + 1;
    double* out = ep.warp_topk + (size_t)warp_global * slots * 2;
    bool taken = !valid;
    for (int e = 0; e < ep.k; e++) {
      double br = taken ? -INFINITY : r; long long bi = taken ? -1 : n_local;
      warp_best(br, bi, true);
      if (bi == n_local && !taken) taken = true;
      if (lane == 0) { out[2 * e] = bi >= 0 ? br : -INFINITY; out[2 * e + 1] = (double)bi; }
    }
    double br = valid ? r : -INFINITY; long long bi = valid ? n_local : -1;
    warp_best(br, bi, false);
    if (lane == 0) { out[2 * ep.k] = bi >= 0 ? br : -INFINITY; out[2 * ep.k + 1] = (double)bi; }
  }
  return epilogue_commit<MAXKNU>(ep, nwarps, KNU, knots, elite_local, n_elite_local);
}

}  // namespace b2
