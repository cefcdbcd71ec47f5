// sim_kernels.cpp — TEST INFRASTRUCTURE: C entry points that run the device code of judo_b200/csrc on the CPU emulator
// (warpsim.h) so the no-GPU test tier can check the kernels' logic against the oracle.  Not part of libb200mpc.so.
#include "leap.cuh"

using namespace b2;

static size_t leap_wstride(int cost_mode, int K, int H) {
  size_t w = ((sizeof(LeapWork) + 15) & ~(size_t)15) + 16 + (cost_mode ? ((size_t)K * LEAP_NU + (size_t)H * K) * sizeof(double) : 0);
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
                                   const double* params, float* cost_NH, double* reward_N, int wpb, int sync_mode, int reverse) {
  const LeapModel* m = reinterpret_cast<const LeapModel*>(consts);
  const size_t ws = leap_wstride(1, K, H);
  wsim::set_reverse(reverse != 0);
  wsim::launch((N + wpb - 1) / wpb, 32 * wpb, wpb * ws, [&] {
    leap_rollout_kernel<true>(m, x0, 0, knots, N, H, K, basis, params, nullptr, nullptr, cost_NH, reward_N, (int)ws, sync_mode << 8, SampleSpec{}, 0);
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
                                  const double* params, float* cost_NH, double* reward_N, int wpb, int sync_mode, int reverse) {
  const Fr3Model* m = reinterpret_cast<const Fr3Model*>(consts);
  const size_t ws = fr3_wstride(1, K, H);
  wsim::set_reverse(reverse != 0);
  wsim::launch((N + wpb - 1) / wpb, 32 * wpb, wpb * ws, [&] {
    fr3_rollout_kernel<true>(m, x0, 0, knots, N, H, K, basis, params, nullptr, nullptr, cost_NH, reward_N, (int)ws, sync_mode, SampleSpec{}, 0);
  });
  return 0;
}

extern "C" int sim_fr3_reward(const double* states, const double* sensors, int N, int H, const double* params, double* reward_N) {
  wsim::set_reverse(false);
  wsim::launch((N + 127) / 128, 128, 0, [&] { fr3_reward_kernel(states, sensors, N, H, params, reward_N); });
  return 0;
}

extern "C" int sim_leap_work_bytes() { return (int)sizeof(LeapWork); }
