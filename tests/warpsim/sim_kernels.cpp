// sim_kernels.cpp — TEST INFRASTRUCTURE: C entry points that run the device code of judo_b200/csrc on the CPU emulator
// (warpsim.h) so the no-GPU test tier can check the kernels' logic against the oracle.  Not part of libb200mpc.so.
#include "leap.cuh"

using namespace b2;


extern "C" int sim_leap_nconsts() { return (int)(sizeof(LeapModel) / sizeof(double)); }
// the kernel's optional counters (B200MPC_LEAP_PROF on the GPU): switch them on for the next launches / read and clear them
static int g_sim_leap_prof = 0;
extern "C" void sim_leap_set_prof(int on) { g_sim_leap_prof = on ? 1 : 0; }
extern "C" void sim_leap_read_prof(unsigned long long* out20) { for (int i = 0; i < 20; i++) { out20[i] = g_leap_prof[i]; g_leap_prof[i] = 0; }
  for (int i = 20; i < 24; i++) g_leap_prof[i] = 0; }

// contract A through leap_rollout_kernel<false>; wpb warps per block, sync_mode as B200MPC_LEAP_SYNC, reverse = lane order
extern "C" int sim_leap_rollout(const double* consts, const double* x0, int batched, const double* controls, int N, int H, double* states,
                                double* sensors, int wpb, int sync_mode, int reverse) {
  const LeapModel* m = reinterpret_cast<const LeapModel*>(consts);
  const size_t ws = leap_wstride(0, 0, H);
  wsim::set_reverse(reverse != 0);
  wsim::launch((N + wpb - 1) / wpb, 32 * wpb, wpb * ws, [&] {
    leap_rollout_kernel<false>(m, x0, batched, controls, N, H, 0, nullptr, nullptr, states, sensors, nullptr, nullptr, (int)ws, (sync_mode << 8) | g_sim_leap_prof,
                               SampleSpec{}, 0);
  });
  return 0;
}

// contract B through leap_rollout_kernel<true>
extern "C" int sim_leap_plan_costs(const double* consts, const double* x0, const double* knots, int N, int K, const double* basis, int H,
                                   const double* params, float* cost_NH, double* reward_N, int wpb, int sync_mode, int reverse, double* trace_out) {
  const LeapModel* m = reinterpret_cast<const LeapModel*>(consts);
  const size_t ws = leap_wstride(1, K, H);
  wsim::set_reverse(reverse != 0);
  wsim::launch((N + wpb - 1) / wpb, 32 * wpb, wpb * ws, [&] {
    leap_rollout_kernel<true>(m, x0, 0, knots, N, H, K, basis, params, nullptr, nullptr, cost_NH, reward_N, (int)ws, (sync_mode << 8) | g_sim_leap_prof, SampleSpec{}, 0, trace_out);
  });
  return 0;
}

// ------------------------------------------------------------------ fr3_pick
#include "fr3.cuh"

extern "C" int sim_fr3_nconsts() { return (int)(sizeof(Fr3Model) / sizeof(double)); }
extern "C" int sim_fr3_work_bytes() { return (int)sizeof(Fr3Work); }

extern "C" int sim_fr3_rollout(const double* consts, const double* x0, int batched, const double* controls, int N, int H, double* states,
                               double* sensors, int wpb, int sync_mode, int reverse) {
  const Fr3Model* m = reinterpret_cast<const Fr3Model*>(consts);
  const size_t ws = fr3_wstride(0, 0, H);
  wsim::set_reverse(reverse != 0);
  wsim::launch((N + wpb - 1) / wpb, 32 * wpb, wpb * ws, [&] {
    fr3_rollout_kernel<false>(m, x0, batched, controls, N, H, 0, nullptr, nullptr, states, sensors, nullptr, nullptr, (int)ws, sync_mode, SampleSpec{}, 0);
  });
  return 0;
}

extern "C" int sim_fr3_plan_costs(const double* consts, const double* x0, const double* knots, int N, int K, const double* basis, int H,
                                  const double* params, float* cost_NH, double* reward_N, int wpb, int sync_mode, int reverse, double* trace_out) {
  const Fr3Model* m = reinterpret_cast<const Fr3Model*>(consts);
  const size_t ws = fr3_wstride(1, K, H);
  wsim::set_reverse(reverse != 0);
  wsim::launch((N + wpb - 1) / wpb, 32 * wpb, wpb * ws, [&] {
    fr3_rollout_kernel<true>(m, x0, 0, knots, N, H, K, basis, params, nullptr, nullptr, cost_NH, reward_N, (int)ws, sync_mode, SampleSpec{}, 0, trace_out);
  });
  return 0;
}

extern "C" int sim_fr3_reward(const double* states, const double* sensors, int N, int H, const double* params, double* reward_N) {
  wsim::set_reverse(false);
  wsim::launch((N + 127) / 128, 128, 0, [&] { fr3_reward_kernel(states, sensors, N, H, params, reward_N); });
  return 0;
}

extern "C" int sim_leap_work_bytes() { return (int)sizeof(LeapWork); }

// ------------------------------------------------------------------ thread-per-rollout tasks + fused optimizer epilogue
#include "kernels.cuh"

template <class Task>
static int sim_small_plan_step(const double* consts, const double* x0, const double* knots, int N, int K, const double* basis, int H,
                               const double* params, int optimizer, const double* opt_params, int n_elite, int threads, float* cost_NH,
                               double* reward_N, double* nominal, double* sigma, double* elite, int reverse) {
  typename Task::Consts c;
  memcpy(&c, consts, sizeof(c));
  const int KNU = K * Task::NU, grid = (N + threads - 1) / threads, nw = grid * (threads / 32);
  PlanEpilogue ep;
  memset(&ep, 0, sizeof(ep));
  ep.optimizer = optimizer;
  const int k_cem = optimizer == EP_CEM ? (int)opt_params[0] : 0;
  ep.k = n_elite > k_cem ? n_elite : k_cem;
  ep.k_cem = k_cem;
  ep.finalize = 1;
  if (optimizer == EP_MPPI) ep.temperature = opt_params[0];
  if (optimizer == EP_CEM) { ep.sigma_min = opt_params[1]; ep.sigma_max = opt_params[2]; }
  std::vector<double> wm((size_t)nw * (2 + KNU) + 1), wt((size_t)nw * (ep.k + 1) * 2 + 1);
  unsigned int ticket = 0;
  ep.warp_mppi = wm.data(); ep.warp_topk = wt.data(); ep.ticket = &ticket;
  ep.nominal = nominal; ep.sigma = sigma; ep.elite = elite;
  const size_t smem = rollout_cost_smem<Task>(threads, H, K, cost_NH != nullptr);
  wsim::set_reverse(reverse != 0);
  wsim::launch(grid, threads, smem, [&] {
    if (K <= 4) rollout_kernel<Task, true, 4>(c, x0, 0, knots, N, H, K, basis, params, nullptr, nullptr, cost_NH, reward_N, ep, SampleSpec{});
    else rollout_kernel<Task, true, 8>(c, x0, 0, knots, N, H, K, basis, params, nullptr, nullptr, cost_NH, reward_N, ep, SampleSpec{});
  });
  return 0;
}

// The same fused launch with the candidates assembled IN the kernel from host-drawn normals (SampleSpec.enabled == 2: what
// b200mpc_controller_step launches when the step's block of normals was speculated and uploaded ahead of it).  z: ((N-1), K*nu).
template <class Task>
static int sim_small_plan_step_hostz(const double* consts, const double* x0, const double* z, const double* nominal_in, const double* sigma_in,
                                     const double* lo, const double* hi, int N, int K, const double* basis, int H, const double* params,
                                     double temperature, int threads, double* reward_N, double* nominal, double* knots_out) {
  typename Task::Consts c;
  memcpy(&c, consts, sizeof(c));
  const int KNU = K * Task::NU, grid = (N + threads - 1) / threads, nw = grid * (threads / 32);
  PlanEpilogue ep;
  memset(&ep, 0, sizeof(ep));
  ep.optimizer = EP_MPPI; ep.finalize = 1; ep.temperature = temperature;
  std::vector<double> wm((size_t)nw * (2 + KNU) + 1), wt((size_t)nw * 2 + 1);
  unsigned int ticket = 0;
  ep.warp_mppi = wm.data(); ep.warp_topk = wt.data(); ep.ticket = &ticket; ep.nominal = nominal;
  SampleSpec smp{};
  smp.enabled = 2; smp.z = z; smp.nominal = nominal_in; smp.sigma = sigma_in; smp.lo = lo; smp.hi = hi; smp.knots_out = knots_out;
  const size_t smem = rollout_cost_smem<Task>(threads, H, K, false);
  wsim::set_reverse(false);
  wsim::launch(grid, threads, smem, [&] {
    if (K <= 4) rollout_kernel<Task, true, 4>(c, x0, 0, knots_out, N, H, K, basis, params, nullptr, nullptr, nullptr, reward_N, ep, smp);
    else rollout_kernel<Task, true, 8>(c, x0, 0, knots_out, N, H, K, basis, params, nullptr, nullptr, nullptr, reward_N, ep, smp);
  });
  return 0;
}
extern "C" int sim_plan_step_hostz(int task, const double* consts, const double* x0, const double* z, const double* nominal_in,
                                   const double* sigma_in, const double* lo, const double* hi, int N, int K, const double* basis, int H,
                                   const double* params, double temperature, int threads, double* reward_N, double* nominal, double* knots_out) {
  if (task == 0) return sim_small_plan_step_hostz<CartpoleTask>(consts, x0, z, nominal_in, sigma_in, lo, hi, N, K, basis, H, params, temperature, threads, reward_N, nominal, knots_out);
  return sim_small_plan_step_hostz<CylinderPushTask>(consts, x0, z, nominal_in, sigma_in, lo, hi, N, K, basis, H, params, temperature, threads, reward_N, nominal, knots_out);
}

// task: 0 cartpole, 1 cylinder_push; optimizer: 0 mppi, 1 cem, 2 ps; one launch = rollout + cost + fused optimizer update
extern "C" int sim_plan_step(int task, const double* consts, const double* x0, const double* knots, int N, int K, const double* basis, int H,
                             const double* params, int optimizer, const double* opt_params, int n_elite, int threads, float* cost_NH,
                             double* reward_N, double* nominal, double* sigma, double* elite, int reverse) {
  if (task == 0) return sim_small_plan_step<CartpoleTask>(consts, x0, knots, N, K, basis, H, params, optimizer, opt_params, n_elite, threads, cost_NH, reward_N, nominal, sigma, elite, reverse);
  return sim_small_plan_step<CylinderPushTask>(consts, x0, knots, N, K, basis, H, params, optimizer, opt_params, n_elite, threads, cost_NH, reward_N, nominal, sigma, elite, reverse);
}

template <class Task>
static int sim_small_rollout(const double* consts, const double* x0, int batched, const double* controls, int N, int H, double* states, double* sensors, int threads) {
  typename Task::Consts c;
  memcpy(&c, consts, sizeof(c));
  PlanEpilogue ep;
  memset(&ep, 0, sizeof(ep));
  ep.optimizer = EP_NONE;
  wsim::set_reverse(false);
  wsim::launch((N + threads - 1) / threads, threads, 0, [&] {
    rollout_kernel<Task, false, 1>(c, x0, batched, controls, N, H, 0, nullptr, nullptr, states, sensors, nullptr, nullptr, ep, SampleSpec{});
  });
  return 0;
}
extern "C" int sim_rollout(int task, const double* consts, const double* x0, int batched, const double* controls, int N, int H, double* states,
                           double* sensors, int threads) {
  if (task == 0) return sim_small_rollout<CartpoleTask>(consts, x0, batched, controls, N, H, states, sensors, threads);
  return sim_small_rollout<CylinderPushTask>(consts, x0, batched, controls, N, H, states, sensors, threads);
}

// ------------------------------------------------------------------ self-test kernels of the emulator itself (tests/test_warpsim_selftest.py)
namespace selftest {

__global__ void warp_collectives(const double* in, double* sum_out, unsigned* ballot_out, double* scan_out, int* all_out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gw = blockIdx.x * (blockDim.x >> 5) + w;
  double v = in[gw * 32 + lane], s = v;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) sum_out[gw] = s;
  const unsigned b = __ballot_sync(0xffffffffu, v > 0.5);
  if (lane == 0) { ballot_out[gw] = b; all_out[gw] = __popc(b); }
  double p = v;
  for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, p, o); if (lane >= o) p += t; }
  scan_out[gw * 32 + lane] = p;
  const int a = __all_sync(0xffffffffu, v >= 0.0);
  if (lane == 1) all_out[gw] += 1000 * a;
}

// block reduction through static shared memory, then a lock-step loop in which warp w needs w + 1 rounds
__global__ void block_sync(const double* in, double* out, int* rounds_out) {
  __shared__ double part[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double s = in[blockIdx.x * blockDim.x + threadIdx.x];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) part[w] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0; for (int i = 0; i < nw; i++) t += part[i]; out[blockIdx.x] = t; }
  int mine = 0, rounds = 0;
  bool done = false;
  for (int it = 0; it < 100; it++) {
    if (!__syncthreads_or(done ? 0 : 1)) break;
    rounds++;
    if (done) continue;
    if (++mine > w) done = true;
  }
  if (lane == 0) rounds_out[blockIdx.x * nw + w] = 100 * rounds + mine;
}

// lane L writes slot L, then reads its neighbour's slot: correct only with the barrier in between
__global__ void neighbour(double* out, int with_barrier) {
  __shared__ double slot[32];
  const int lane = threadIdx.x & 31;
  slot[lane] = -1.0;
  __syncwarp();
  slot[lane] = (double)lane;
  if (with_barrier) __syncwarp();
  out[lane] = slot[(lane + 1) & 31];
}

__global__ void tma_copy(const double* src, double* out, int n) {
  B2_DYNAMIC_SMEM(unsigned char, raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(raw);
  double* buf = reinterpret_cast<double*>(raw + 16);
  if (threadIdx.x == 0) { b2::mbar_init(bar, 1); b2::fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { b2::mbar_expect_tx(bar, n * 8); b2::tma_bulk_g2s(buf, src, n * 8, bar); }
  b2::mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = 2 * buf[i];
}

}  // namespace selftest

extern "C" void sim_selftest_warp(const double* in, int nblocks, int warps, double* sum_out, unsigned* ballot_out, double* scan_out, int* all_out, int reverse) {
  wsim::set_reverse(reverse != 0);
  wsim::launch(nblocks, 32 * warps, 0, [&] { selftest::warp_collectives(in, sum_out, ballot_out, scan_out, all_out); });
}
extern "C" void sim_selftest_block(const double* in, int nblocks, int warps, double* out, int* rounds_out, int reverse) {
  wsim::set_reverse(reverse != 0);
  wsim::launch(nblocks, 32 * warps, 0, [&] { selftest::block_sync(in, out, rounds_out); });
}
extern "C" void sim_selftest_neighbour(double* out, int with_barrier, int reverse) {
  wsim::set_reverse(reverse != 0);
  wsim::launch(1, 32, 0, [&] { selftest::neighbour(out, with_barrier); });
}
extern "C" void sim_selftest_tma(const double* src, double* out, int n, int reverse) {
  wsim::set_reverse(reverse != 0);
  wsim::launch(1, 64, 16 + n * 8, [&] { selftest::tma_copy(src, out, n); });
}
