// sim_kernels.cpp — TEST INFRASTRUCTURE: C entry points that run the device code of judo_b200/csrc on the CPU emulator
// (warpsim.h) so the no-GPU test tier can check the kernels' logic against the oracle.  Not part of libb200mpc.so.
#include "leap.cuh"

using namespace b2;

static size_t leap_wstride(int cost_mode, int K, int H) {
  (void)H;
  size_t w = ((sizeof(LeapWork) + 15) & ~(size_t)15) + 16 + (cost_mode ? (size_t)K * LEAP_NU * sizeof(double) : 0);
  return (w + 15) & ~(size_t)15;
}

extern "C" int sim_leap_nconsts() { return (int)(sizeof(LeapModel) / sizeof(double)); }

// contract A through leap_rollout_kernel<false>; wpb warps per block, sync_mode as B200MPC_LEAP_SYNC, reverse = lane order
extern "C" int sim_leap_rollout(const double* consts, const double* x0, int batched, const double* controls, int N, int H, double* states,
                                double* sensors, int wpb, int sync_mode, int reverse) {
  const LeapModel* m = reinterpret_cast<const LeapModel*>(consts);
  const size_t ws = leap_wstride(0, 0, H);
  wsim::set_reverse(reverse != 0);
  wsim::launch((N + wpb - 1) / wpb, 32 * wpb, wpb * ws, [&] {
    leap_rollout_kernel<false>(m, x0, batched, controls, N, H, 0, nullptr, nullptr, states, sensors, nullptr, nullptr, (int)ws, sync_mode << 8,
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
    leap_rollout_kernel<true>(m, x0, 0, knots, N, H, K, basis, params, nullptr, nullptr, cost_NH, reward_N, (int)ws, sync_mode << 8, SampleSpec{}, 0, trace_out);
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
