// kernels.cuh — the fused rollout kernel for the thread-per-rollout tasks and the optimizer-update reductions.
#pragma once
#include "epilogue.cuh"
#include "sampling.cuh"
#include "small_tasks.cuh"

namespace b2 {

// =========================================================================================== fused rollout kernel
// COST=false (contract A, RolloutBackend.rollout):   in = controls (N,H,NU) -> states (N,H,NX), sensors (N,H,NS)
// COST=true  (contract B, fused plan):               in = knots (N,K,NU), basis (H,K) -> cost (N,H) f32, reward (N)
//
// One thread per rollout.  The block's slice of the knots array and the spline basis are staged to shared memory
// with 1-D TMA bulk copies (cp.async.bulk + mbarrier) issued by thread 0; the per-step cost row of every rollout is
// staged in shared memory and written back as coalesced 128-bit stores, so HBM sees only the algorithmic bytes.
template <class Task, bool COST, int MAXK>
__global__ void __launch_bounds__(128) rollout_kernel(const typename Task::Consts c, const double* __restrict__ x0, int x0_batched,
                                                      const double* __restrict__ in, int N, int H, int K,
                                                      const double* __restrict__ basis, const double* __restrict__ cost_params,
                                                      double* __restrict__ states, double* __restrict__ sensors,
                                                      float* __restrict__ cost_NH, double* __restrict__ reward_N,
                                                      const PlanEpilogue ep, const SampleSpec smp) {
  constexpr int NU = Task::NU, NX = Task::NX, NS = Task::NS;
  B2_DYNAMIC_SMEM(unsigned char, smem_raw);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int n0 = blockIdx.x * nthr, n = n0 + tid;
  const int nblk = min(nthr, N - n0);

  if (ep.stamps && blockIdx.x == 0 && tid == 0) ep.stamps[0] = global_ns();
  typename Task::State s;
  if constexpr (COST) {
    // shared layout: [mbarrier 16B][basis H*K doubles (padded to 16B)][knots nthr*K*NU doubles][cost tile nthr*(H+1) floats]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    double* sB = reinterpret_cast<double*>(smem_raw + 16);
    const int nB = H * K, nBpad = (nB + 1) & ~1;
    double* sK = sB + nBpad;
    float* sC = reinterpret_cast<float*>(sK + (size_t)nthr * K * NU);
    const unsigned bytesB = (unsigned)(nB * sizeof(double)), bytesK = (unsigned)((size_t)nblk * K * NU * sizeof(double));
    const double* gK = in + (size_t)n0 * K * NU;
    const bool want_knots = !smp.enabled;  // sampled candidates are generated in registers below
    const bool tma_ok = (bytesB % 16 == 0) && (!want_knots || bytesK % 16 == 0) && ((reinterpret_cast<uintptr_t>(basis) & 15) == 0) &&
                        (!want_knots || (reinterpret_cast<uintptr_t>(gK) & 15) == 0);
    if (tma_ok) {
      if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
      __syncthreads();
      if (tid == 0) {
        mbar_expect_tx(bar, bytesB + (want_knots ? bytesK : 0u));
        tma_bulk_g2s(sB, basis, bytesB, bar);
        if (want_knots) tma_bulk_g2s(sK, gK, bytesK, bar);
      }
      mbar_wait(bar, 0);
    } else {
      for (int i = tid; i < nB; i += nthr) sB[i] = basis[i];
      if (want_knots) for (int i = tid; i < nblk * K * NU; i += nthr) sK[i] = gK[i];
      __syncthreads();
    }
    double cp[Task::NCOST];
#pragma unroll
    for (int i = 0; i < Task::NCOST; i++) cp[i] = cost_params[i];
    double kn[MAXK * NU];
    double reward = 0;
#pragma unroll
    for (int i = 0; i < MAXK * NU; i++) kn[i] = 0;
    if (n < N) {
      if (smp.enabled) {
        // on-device sampling: Philox keyed by the GLOBAL rollout index -> identical candidates however N is sharded
        const long long gn = (long long)n + ep.index_offset;
        const int KNU = K * NU;
        if (smp.enabled == 2) {  // host-drawn normals, uploaded ahead of the step: assemble this rollout's candidate from its row
          const double* zr = smp.z + (size_t)(gn > 0 ? gn - 1 : 0) * KNU;
#pragma unroll
          for (int e = 0; e < MAXK * NU; e++) if (e < KNU) kn[e] = sample_element_hostz(smp, gn, e, NU, gn > 0 ? __ldg(zr + e) : 0.0);
        } else {
#pragma unroll
          for (int p2 = 0; p2 < (MAXK * NU + 1) / 2; p2++) {
            if (2 * p2 < KNU) {
              double z0, z1;
              normal_pair(smp, gn, p2, &z0, &z1);
              kn[2 * p2] = sample_element(smp, gn, 2 * p2, NU, z0);
              if (2 * p2 + 1 < MAXK * NU && 2 * p2 + 1 < KNU) kn[2 * p2 + 1] = sample_element(smp, gn, 2 * p2 + 1, NU, z1);
            }
          }
        }
#pragma unroll
        for (int e = 0; e < MAXK * NU; e++) if (e < KNU) smp.knots_out[(size_t)n * KNU + e] = kn[e];
      } else {
#pragma unroll
        for (int k = 0; k < MAXK; k++)
#pragma unroll
          for (int j = 0; j < NU; j++) kn[k * NU + j] = k < K ? sK[(size_t)tid * K * NU + k * NU + j] : 0.0;
      }
      Task::load(s, x0 + (x0_batched ? (size_t)n * NX : 0));
      double total = 0;
      for (int t = 0; t < H; t++) {
        double u[NU];
#pragma unroll
        for (int j = 0; j < NU; j++) u[j] = 0;
#pragma unroll
        for (int k = 0; k < MAXK; k++) {
          if (k < K) {
            double b = sB[t * K + k];
#pragma unroll
            for (int j = 0; j < NU; j++) u[j] += b * kn[k * NU + j];
          }
        }
        const double ct = Task::step_cost(c, s, u, cp);
        if (ep.trace_q) {  // positions after step t: the elites' trace sensors are evaluated from them by the last warp (below)
#pragma unroll
          for (int j = 0; j < Task::NQ; j++) ep.trace_q[((size_t)t * N + n) * Task::NQ + j] = s.q[j];  // time-major: a warp's stores coalesce
        }
        total += ct;
        if (cost_NH) sC[(size_t)tid * (H + 1) + t] = (float)ct;
      }
      reward = Task::finish(total, H);
      reward_N[n] = reward;
    }
    if (ep.optimizer != EP_NONE || ep.k > 0) {
      long long el[EP_MAXK];
      int ne = 0;
      const bool last = epilogue_thread_per_rollout<MAXK * NU>(ep, n < N, n, reward, kn, K * NU, blockIdx.x * (nthr >> 5) + (tid >> 5),
                                                               gridDim.x * (nthr >> 5), smp.enabled ? smp.knots_out : in, el, ne);
      if (last && ep.elite_sens && ep.trace_q) {
        // T1 (controller.py:323-363): sensors of the elite rollouts at every step, from the captured positions.  MuJoCo evaluates
        // position sensors in mj_forward, i.e. at the PRE-step state: step 0 sees x0, step t the positions after step t-1.
        // (Every rollout's trace_q stores precede its warp's ticket (threadfence + atomic), so they are visible here; __ldcg reads L2.)
        const int lane = tid & 31;
        const int nt = min(ne, ep.n_trace);
        // four (elite, step) items per lane and round, their L2 loads issued together: the stage is one warp deep, so a round costs one
        // L2 round trip whatever it carries (one item per round made this tail ~15 us at 5 elites x 64 steps)
        for (int base = 0; base < nt * H; base += 128) {
          double q[4][Task::NQ];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int idx = base + 32 * u + lane;
            if (idx < nt * H) {
              const int e = idx / H, t = idx - e * H;
              long long r = el[0];
#pragma unroll
              for (int i = 1; i < EP_MAXK; i++) if (i == e) r = el[i];
#pragma unroll
              for (int j = 0; j < Task::NQ; j++) q[u][j] = t == 0 ? x0[j] : __ldcg(ep.trace_q + ((size_t)(t - 1) * N + r) * Task::NQ + j);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int idx = base + 32 * u + lane;
            if (idx < nt * H) {
              const int e = idx / H, t = idx - e * H;
              double sens[NS];
              Task::sensors(c, q[u], sens);
#pragma unroll
              for (int j = 0; j < NS; j++) ep.elite_sens[((size_t)e * H + t) * NS + j] = sens[j];
            }
          }
        }
      }
    }
    if (cost_NH) {
      __syncthreads();
      // coalesced write-back of the block's (nblk, H) tile: consecutive threads write consecutive floats
      float* g = cost_NH + (size_t)n0 * H;
      const int tot = nblk * H;
      if ((H & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        for (int i = tid * 4; i < tot; i += nthr * 4) {
          int r = i / H, col = i - r * H;
          const float* src = sC + (size_t)r * (H + 1) + col;
          float4 v = make_float4(src[0], src[1], src[2], src[3]);
          *reinterpret_cast<float4*>(g + i) = v;
        }
      } else {
        for (int i = tid; i < tot; i += nthr) { int r = i / H, col = i - r * H; g[i] = sC[(size_t)r * (H + 1) + col]; }
      }
    }
  } else {
    if (n >= N) return;
    Task::load(s, x0 + (x0_batched ? (size_t)n * NX : 0));
    const double* uc = in + (size_t)n * H * NU;
    double* so = states + (size_t)n * H * NX;
    double* se = sensors ? sensors + (size_t)n * H * NS : nullptr;
    for (int t = 0; t < H; t++) {
      double u[NU], sens[NS];
#pragma unroll
      for (int j = 0; j < NU; j++) u[j] = uc[t * NU + j];
      Task::step(c, s, u, se ? sens : nullptr);
      double x[NX];
      Task::store(s, x);
      if constexpr (NX % 2 == 0) {
#pragma unroll
        for (int j = 0; j < NX; j += 2) *reinterpret_cast<double2*>(so + (size_t)t * NX + j) = make_double2(x[j], x[j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < NX; j++) so[(size_t)t * NX + j] = x[j];
      }
      if (se) {
#pragma unroll
        for (int j = 0; j < NS; j += 2) *reinterpret_cast<double2*>(se + (size_t)t * NS + j) = make_double2(sens[j], sens[j + 1]);
      }
    }
  }
}

// reward from given trajectories (Task.reward for contract-A callers): one thread per rollout
template <class Task>
__global__ void reward_kernel(const double* __restrict__ states, const double* __restrict__ controls, int N, int H,
                              const double* __restrict__ cost_params, double* __restrict__ reward_N) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double cp[Task::NCOST];
#pragma unroll
  for (int i = 0; i < Task::NCOST; i++) cp[i] = cost_params[i];
  double total = 0;
  for (int t = 0; t < H; t++) {
    typename Task::State s;
    Task::load(s, states + ((size_t)n * H + t) * Task::NX);
    total += Task::cost(cp, s, controls + ((size_t)n * H + t) * Task::NU);
  }
  reward_N[n] = Task::finish(total, H);
}

template <class Task>
inline size_t rollout_cost_smem(int threads, int H, int K, bool want_cost) {
  size_t nB = ((size_t)H * K + 1) & ~(size_t)1;
  return 16 + nB * 8 + (size_t)threads * K * Task::NU * 8 + (want_cost ? (size_t)threads * (H + 1) * 4 : 0);
}

// =========================================================================================== reductions
__device__ inline double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ inline double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide reductions through one value per warp in shared memory; result broadcast to all threads
__device__ inline double block_min(double v, double* sh) {
  v = warp_min(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = l < nw ? sh[l] : INFINITY;
  r = warp_min(r);
  return r;
}
__device__ inline double block_sum(double v, double* sh) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = l < nw ? sh[l] : 0.0;
  r = warp_sum(r);
  return r;
}

// MPPI partial (mppi.py:61-82 restricted to this block's rollouts):
//   beta_b = min(-r); w_n = exp(-(c_n - beta_b)/T); S_b = sum w; V_b[j] = sum_n w_n knots[n][j]
// partial[b] = [beta_b, S_b, V_b[KNU]].  grid = number of partials, blockDim = 256.
__global__ void __launch_bounds__(256) mppi_partial_kernel(const double* __restrict__ knots, const double* __restrict__ rewards, int N,
                                                           int KNU, double temperature, double* __restrict__ partial) {
  __shared__ double sh[32];
  B2_DYNAMIC_SMEM(double, sw);  // weights of this block's chunk
  const int nb = gridDim.x, b = blockIdx.x;
  const int chunk = (N + nb - 1) / nb, lo = b * chunk, hi = min(N, lo + chunk), cnt = max(0, hi - lo);
  double m = INFINITY;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) m = fmin(m, -rewards[lo + i]);
  const double beta = block_min(m, sh);
  double s = 0;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    double w = exp(-((-rewards[lo + i]) - beta) / temperature);
    sw[i] = w;
    s += w;
  }
  const double S = block_sum(s, sh);
  __syncthreads();
  double* out = partial + (size_t)b * (2 + KNU);
  if (threadIdx.x == 0) { out[0] = cnt ? beta : INFINITY; out[1] = cnt ? S : 0.0; }
  // V_b[j]: thread t owns column j = t % KNU and every `groups`-th rollout; fixed-order tree through shared memory
  double* sred = sw + cnt;  // blockDim.x doubles
  const int groups = KNU <= (int)blockDim.x ? (int)blockDim.x / KNU : 1;
  for (int j0 = 0; j0 < KNU; j0 += blockDim.x) {
    const int width = min(KNU - j0, (int)blockDim.x);
    const int g = threadIdx.x / width, j = j0 + threadIdx.x % width;
    double acc = 0;
    if (g < groups)
      for (int i = g; i < cnt; i += groups) acc += sw[i] * knots[(size_t)(lo + i) * KNU + j];
    __syncthreads();
    sred[threadIdx.x] = acc;
    __syncthreads();
    if ((int)threadIdx.x < width) {
      double v = 0;
      for (int gg = 0; gg < groups; gg++) v += sred[gg * width + threadIdx.x];
      out[2 + j0 + threadIdx.x] = v;
    }
  }
}

// Combine MPPI partials (from blocks and/or ranks): rescale every partial to the global beta.
__global__ void mppi_combine_kernel(const double* __restrict__ partials, int np, int KNU, double temperature,
                                    double* __restrict__ nominal) {
  __shared__ double sh[32];
  double m = INFINITY;
  for (int i = threadIdx.x; i < np; i += blockDim.x) m = fmin(m, partials[(size_t)i * (2 + KNU)]);
  const double beta = block_min(m, sh);
  double s = 0;
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    const double* p = partials + (size_t)i * (2 + KNU);
    if (p[1] > 0) s += p[1] * exp(-(p[0] - beta) / temperature);
  }
  const double S = block_sum(s, sh);
  for (int j = threadIdx.x; j < KNU; j += blockDim.x) {
    double v = 0;
    for (int i = 0; i < np; i++) {
      const double* p = partials + (size_t)i * (2 + KNU);
      if (p[1] > 0) v += p[2 + j] * exp(-(p[0] - beta) / temperature);
    }
    nominal[j] = v / S;
  }
}

// rows[e] = knots[idx[e] - index_offset] for the elite list (leap path; the fused kernels do this in their epilogue)
__global__ void gather_rows_kernel(const double* __restrict__ knots, const double* __restrict__ idx, int index_offset, int KNU,
                                   double* __restrict__ rows) {
  const long long i = (long long)idx[blockIdx.x] - index_offset;
  for (int j = threadIdx.x; j < KNU; j += blockDim.x) rows[(size_t)blockIdx.x * KNU + j] = i >= 0 ? knots[(size_t)i * KNU + j] : 0.0;
}

// total order for elite selection: larger reward first; ties by index (higher first for CEM's flipped argsort,
// lower first for PS's argmax)
__device__ inline bool key_better(double ra, long long ia, double rb, long long ib, int prefer_high) {
  if (ra != rb) return ra > rb;
  return prefer_high ? ia > ib : ia < ib;
}

// top-k of this block's chunk -> partial[b] = k x [reward, global index, knots[KNU]]; missing entries get index -1
__global__ void __launch_bounds__(256) topk_partial_kernel(const double* __restrict__ knots, const double* __restrict__ rewards, int N,
                                                           int KNU, int k, int index_offset, int prefer_high,
                                                           double* __restrict__ partial) {
  __shared__ double sr[256];
  __shared__ long long si[256];
  const int nb = gridDim.x, b = blockIdx.x;
  const int chunk = (N + nb - 1) / nb, lo = b * chunk, hi = min(N, lo + chunk);
  double prev_r = INFINITY;
  long long prev_i = prefer_high ? (1LL << 62) : -1;
  bool have_prev = false;
  double* out = partial + (size_t)b * k * (2 + KNU);
  for (int e = 0; e < k; e++) {
    double br = -INFINITY;
    long long bi = -1;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      double r = rewards[i];
      if (have_prev && !key_better(prev_r, prev_i, r, i, prefer_high)) continue;  // must rank strictly after the previous pick
      if (bi < 0 || key_better(r, i, br, bi, prefer_high)) { br = r; bi = i; }
    }
    sr[threadIdx.x] = br; si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        double r2 = sr[threadIdx.x + o]; long long i2 = si[threadIdx.x + o];
        if (i2 >= 0 && (si[threadIdx.x] < 0 || key_better(r2, i2, sr[threadIdx.x], si[threadIdx.x], prefer_high))) { sr[threadIdx.x] = r2; si[threadIdx.x] = i2; }
      }
      __syncthreads();
    }
    br = sr[0]; bi = si[0];
    __syncthreads();
    double* o = out + (size_t)e * (2 + KNU);
    if (threadIdx.x == 0) { o[0] = bi >= 0 ? br : -INFINITY; o[1] = bi >= 0 ? (double)(bi + index_offset) : -1.0; }
    for (int j = threadIdx.x; j < KNU; j += blockDim.x) o[2 + j] = bi >= 0 ? knots[(size_t)bi * KNU + j] : 0.0;
    prev_r = br; prev_i = bi; have_prev = true;
    if (bi < 0) { prev_r = -INFINITY; }
  }
}

// final top-k over np*k candidates; nominal = mean of elites; sigma = clip(sqrt(var, ddof 0)) (cem.py:88-91);
// PS is k=1 (ps.py:64-65).  elite_idx (k doubles) receives the chosen global indices in descending order.
__global__ void __launch_bounds__(256) topk_combine_kernel(const double* __restrict__ partials, int np, int KNU, int k, int prefer_high,
                                                           double sigma_min, double sigma_max, double* __restrict__ nominal,
                                                           double* __restrict__ sigma, double* __restrict__ elite_idx) {
  __shared__ double sr[256];
  __shared__ long long si[256];
  __shared__ int chosen[256];
  __shared__ int nch;
  const int ncand = np * k, stride = 2 + KNU;
  double prev_r = INFINITY;
  long long prev_i = 0;
  bool have_prev = false;
  if (threadIdx.x == 0) nch = 0;
  __syncthreads();
  for (int e = 0; e < k && e < 256; e++) {
    double br = -INFINITY; long long bi = -1; int bc = -1;
    for (int cidx = threadIdx.x; cidx < ncand; cidx += blockDim.x) {
      const double* p = partials + (size_t)cidx * stride;
      long long gi = (long long)p[1];
      if (gi < 0) continue;
      if (have_prev && !key_better(prev_r, prev_i, p[0], gi, prefer_high)) continue;
      if (bi < 0 || key_better(p[0], gi, br, bi, prefer_high)) { br = p[0]; bi = gi; bc = cidx; }
    }
    sr[threadIdx.x] = br; si[threadIdx.x] = bi >= 0 ? ((bi << 20) | (long long)bc) : -1;  // pack candidate slot (ncand < 2^20)
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        double r2 = sr[threadIdx.x + o]; long long p2 = si[threadIdx.x + o];
        long long p1 = si[threadIdx.x];
        if (p2 >= 0 && (p1 < 0 || key_better(r2, p2 >> 20, sr[threadIdx.x], p1 >> 20, prefer_high))) { sr[threadIdx.x] = r2; si[threadIdx.x] = p2; }
      }
      __syncthreads();
    }
    long long pk = si[0];
    br = sr[0];
    __syncthreads();
    if (pk < 0) break;
    if (threadIdx.x == 0) { chosen[nch] = (int)(pk & ((1 << 20) - 1)); if (elite_idx) elite_idx[nch] = (double)(pk >> 20); nch++; }
    prev_r = br; prev_i = pk >> 20; have_prev = true;
    __syncthreads();
  }
  __syncthreads();
  const int ne = nch;
  if (elite_idx) for (int e = ne + threadIdx.x; e < k; e += blockDim.x) elite_idx[e] = -1.0;
  for (int j = threadIdx.x; j < KNU; j += blockDim.x) {
    double mean = 0;
    for (int e = 0; e < ne; e++) mean += partials[(size_t)chosen[e] * stride + 2 + j];
    mean = ne ? mean / ne : 0.0;
    nominal[j] = mean;
    if (sigma) {
      double var = 0;
      for (int e = 0; e < ne; e++) { double dlt = partials[(size_t)chosen[e] * stride + 2 + j] - mean; var += dlt * dlt; }
      var = ne ? var / ne : 0.0;
      sigma[j] = fmin(fmax(sqrt(var), sigma_min), sigma_max);
    }
  }
}

}  // namespace b2
