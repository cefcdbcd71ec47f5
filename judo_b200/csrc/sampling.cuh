// sampling.cuh — on-device candidate sampling ("perf mode" of SURVEY.md §8a S1-S3 / hard part 5).
//
// The reference draws  noised = nominal + sigma * np.random.randn(N-1, K, nu)  with row 0 the un-noised nominal
// (judo/optimizers/mppi.py:58-59, cem.py:73-74, ps.py:49-50) and then clips to the actuator range
// (judo/controller/controller.py:253-257).  Seed-parity mode keeps that on the host (NumPy's MT19937 stream cannot be
// reproduced on the fly on a GPU); this mode generates the SAME distribution inside the rollout kernel with a counter-based
// Philox4x32-10 generator keyed by (seed, plan-step counter, GLOBAL rollout index, element), so results are invariant to how
// the rollouts are sharded over blocks and GPUs, and nothing but the (K, nu) nominal crosses PCIe.
#pragma once
#include "common.cuh"

namespace b2 {

struct SampleSpec {
  int enabled;
  unsigned long long seed, counter;  // counter: plan-step index
  const double* nominal;             // (K*nu)
  const double* sigma;               // (K*nu) per-element standard deviation (ramp and CEM state already applied)
  const double* lo;                  // (nu) clip range (may hold -inf / +inf)
  const double* hi;                  // (nu)
  double* knots_out;                 // (N, K*nu) the generated candidates (read back by the epilogue and, on request, the host)
  // enabled == 2: the normals are the HOST's (NumPy's legacy stream, drawn and uploaded ahead of the step): ((N-1), K*nu) in the
  // reference's draw order; the kernel only assembles clip(nominal + sigma * z) — seed parity without the candidates crossing PCIe
  const double* z;
};

__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two independent standard normals for (global rollout n, element pair p) — Box-Muller on two 53-bit uniforms
__device__ __forceinline__ void normal_pair(const SampleSpec& s, long long n, int p, double* z0, double* z1) {
  unsigned r[4];
  philox4x32_10((unsigned)(n & 0xffffffffll), (unsigned)(n >> 32), (unsigned)p, (unsigned)(s.counter & 0xffffffffull),
                (unsigned)(s.seed & 0xffffffffull), (unsigned)(s.seed >> 32) ^ (unsigned)(s.counter >> 32), r);
  const double u1 = ((double)(((unsigned long long)r[0] << 21) ^ (unsigned long long)(r[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);  // (0,1)
  const double u2 = ((double)(((unsigned long long)r[2] << 21) ^ (unsigned long long)(r[3] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
  const double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  *z0 = rad * cs; *z1 = rad * sn;
}

// candidate element e of global rollout n:  clip(nominal + sigma * z, lo, hi); rollout 0 is the un-noised nominal
__device__ __forceinline__ double sample_element(const SampleSpec& s, long long n, int e, int nu, double z) {
  const double v = s.nominal[e] + (n == 0 ? 0.0 : s.sigma[e] * z);
  const int j = e % nu;
  return fmin(fmax(v, s.lo[j]), s.hi[j]);
}

// the same element from a host-drawn normal, rounded exactly as NumPy rounds `nominal + sigma * noise` (two roundings, no fused
// multiply-add) and clipped as np.clip does (compare + select: NaN stays NaN)
__device__ __forceinline__ double sample_element_hostz(const SampleSpec& s, long long n, int e, int nu, double z) {
#ifdef B2_HOST_SIM
  double v = z * s.sigma[e];
  v = v + s.nominal[e];
#else
  double v = __dadd_rn(__dmul_rn(z, s.sigma[e]), s.nominal[e]);
#endif
  if (n == 0) v = s.nominal[e];
  const int j = e % nu;
  const double l = s.lo[j], h = s.hi[j];
  v = v < l ? l : v;
  v = v > h ? h : v;
  return v;
}

}  // namespace b2
