// host_glue.cpp — CPU side of the Controller.update_action fast path (pure host code, no CUDA; linked into libb200mpc.so).
//
//   * mt19937_normals : NumPy's LEGACY global stream (np.random.randn = RandomState.standard_normal: MT19937 -> two 53-bit doubles ->
//                       polar Box-Muller with rejection, the second value of each pair returned first) reproduced bit for bit, working
//                       directly on the bit generator's state (numpy/random/src/mt19937, legacy-distributions.c:legacy_gauss).  The
//                       reference samples its candidates from that stream (judo/optimizers/mppi.py:58, cem.py:73, ps.py:49), so identical
//                       seeds must give identical candidates; this routine is ~2x faster than numpy's scalar loop because the stages are
//                       batched (bulk word generation, vectorisable conversion / compaction / sqrt-div, one libm log per accepted pair).
//   * candidates      : row 0 = nominal, rows 1.. = nominal + sigma * z, clipped (mppi.py:58-59 + controller.py:253-258).
//   * spline_basis    : the (H, K) matrix of scipy's interp1d(kind = zero | linear | cubic) (controller.py:382-401), same formulas as
//                       judo_b200/spline.py.
//   * trace_segments  : Controller.update_traces' (elite, sensor, step) line segments (controller.py:341-363).
// Compiled with -ffp-contract=off: every product/sum below is rounded exactly as NumPy rounds it.
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <vector>

// SIMD variants of the hot loops, chosen at load time by the dynamic linker (the library is built on one CPU and run on another)
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define B2_SIMD_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define B2_SIMD_CLONES
#endif

namespace b2host {

// ------------------------------------------------------------------ MT19937 (state layout of numpy's mt19937_state: key[624], pos)
static constexpr int MT_N = 624, MT_M = 397;

static inline void mt_regenerate(uint32_t* k) {
  uint32_t y;
  int i;
  for (i = 0; i < MT_N - MT_M; i++) { y = (k[i] & 0x80000000u) | (k[i + 1] & 0x7fffffffu); k[i] = k[i + MT_M] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu); }
  for (; i < MT_N - 1; i++) { y = (k[i] & 0x80000000u) | (k[i + 1] & 0x7fffffffu); k[i] = k[i + (MT_M - MT_N)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu); }
  y = (k[MT_N - 1] & 0x80000000u) | (k[0] & 0x7fffffffu);
  k[MT_N - 1] = k[MT_M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
}
static inline uint32_t mt_temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

// next `nw` tempered 32-bit outputs of the stream
B2_SIMD_CLONES static void mt_words(uint32_t* key, int* ppos, uint32_t* out, size_t nw) {
  int pos = *ppos;
  size_t got = 0;
  while (got < nw) {
    if (pos >= MT_N) { mt_regenerate(key); pos = 0; }
    size_t take = (size_t)(MT_N - pos);
    if (take > nw - got) take = nw - got;
    for (size_t j = 0; j < take; j++) out[got + j] = mt_temper(key[pos + j]);
    got += take;
    pos += (int)take;
  }
  *ppos = pos;
}

// One slice [lo, hi) of a round's attempts: legacy_double conversion, acceptance test, order-preserving compaction to the front of
// the slice, one libm log per accepted pair, the two normals of each pair into pairs[2 * lo ...].  Returns the accepted count.
B2_SIMD_CLONES static size_t normals_slice(const uint32_t* w, double* x1, double* x2, double* r2, double* lg, double* pairs, size_t lo, size_t hi) {
  // legacy_double: (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53; then 2 x - 1 for both coordinates
  for (size_t a = lo; a < hi; a++) {
    const double u = ((double)(int32_t)(w[4 * a] >> 5) * 67108864.0 + (double)(int32_t)(w[4 * a + 1] >> 6)) / 9007199254740992.0;
    const double v = ((double)(int32_t)(w[4 * a + 2] >> 5) * 67108864.0 + (double)(int32_t)(w[4 * a + 3] >> 6)) / 9007199254740992.0;
    const double p = 2.0 * u - 1.0, q = 2.0 * v - 1.0;
    x1[a] = p; x2[a] = q; r2[a] = p * p + q * q;
  }
  size_t acc = lo;  // branch-free compaction of the accepted attempts (r2 < 1 and r2 != 0)
  for (size_t a = lo; a < hi; a++) {
    const double r = r2[a];
    x1[acc] = x1[a]; x2[acc] = x2[a]; r2[acc] = r;
    acc += !(r >= 1.0 || r == 0.0);
  }
  for (size_t a = lo; a < acc; a++) lg[a] = log(r2[a]);  // the same libm log numpy calls
  for (size_t a = lo; a < acc; a++) {
    const double f = sqrt(-2.0 * lg[a] / r2[a]);
    pairs[2 * a] = f * x2[a];       // legacy_gauss returns f * x2 first and keeps f * x1 for the next call
    pairs[2 * a + 1] = f * x1[a];
  }
  return acc - lo;
}

// out[0..n): the next n legacy normals, n EVEN, the generator's has_gauss flag clear on entry (and left clear).
// Attempts are processed in rounds of exactly as many as there are pairs still missing, so the stream is never consumed past the
// last accepted attempt (a rejected attempt costs 4 words, exactly as in legacy_gauss).
void mt19937_normals(uint32_t* key, int* pos, double* out, size_t n) {
  constexpr size_t CH = 8192;  // attempts per round (scratch stays L2-resident)
  static thread_local std::vector<uint32_t> wbuf(4 * CH);
  static thread_local std::vector<double> X1(CH), X2(CH), R2(CH), LG(CH);
  uint32_t* w = wbuf.data();
  double *x1 = X1.data(), *x2 = X2.data(), *r2 = R2.data(), *lg = LG.data();
  size_t i = 0;
  while (i < n) {
    const size_t need = (n - i) / 2;
    const size_t na = need < CH ? need : CH;
    mt_words(key, pos, w, 4 * na);
    i += 2 * normals_slice(w, x1, x2, r2, lg, out + i, 0, na);
  }
}

// ------------------------------------------------------------------ candidates
// z: (N-1) * KNU normals in the reference's draw order.  knots[0] = clip(nominal); knots[1 + r] = clip(nominal + sigma * z[r]) with the two
// roundings of NumPy's `nominal + sigma * noise` (base.py:_noised) and np.clip's min(max(x, lo), hi).
B2_SIMD_CLONES void assemble_candidates(const double* z, const double* nominal, const double* sigma, const double* lo, const double* hi, int N, int K, int nu,
                         double* knots) {
  const int KNU = K * nu;
  for (int e = 0; e < KNU; e++) {
    const double l = lo[e % nu], h = hi[e % nu];
    double v = nominal[e];
    v = v < l ? l : v;
    v = v > h ? h : v;
    knots[e] = v;
  }
  if (KNU <= 64) {
    double l64[64], h64[64];
    for (int e = 0; e < KNU; e++) { l64[e] = lo[e % nu]; h64[e] = hi[e % nu]; }
    for (int r = 1; r < N; r++) {
      const double* zr = z + (size_t)(r - 1) * KNU;
      double* kr = knots + (size_t)r * KNU;
      for (int e = 0; e < KNU; e++) {
        double v = zr[e] * sigma[e];
        v = v + nominal[e];
        v = v < l64[e] ? l64[e] : v;   // NaN stays NaN, as in np.clip
        v = v > h64[e] ? h64[e] : v;
        kr[e] = v;
      }
    }
  } else {
    for (int r = 1; r < N; r++)
      for (int e = 0; e < KNU; e++) {
        double v = z[(size_t)(r - 1) * KNU + e] * sigma[e];
        v = v + nominal[e];
        const double l = lo[e % nu], h = hi[e % nu];
        v = v < l ? l : v;
        v = v > h ? h : v;
        knots[(size_t)r * KNU + e] = v;
      }
  }
}

// ------------------------------------------------------------------ spline basis (judo_b200/spline.py)
// order: 0 zero, 1 linear, 2 cubic (not-a-knot).  t (K) strictly increasing, q (H).  B (H, K) row-major.  Returns 0, or 1 on bad input.
int spline_basis(int order, const double* t, int K, const double* q, int H, double* B) {
  if (K < 1 || H < 0 || order < 0 || order > 2 || (order == 2 && K < 4) || (order >= 1 && K < 2)) return 1;
  memset(B, 0, sizeof(double) * (size_t)H * K);
  double C[12][12];
  if (order == 2) {
    if (K > 12) return 1;
    // second-derivative operator of the not-a-knot cubic spline: solve A m = R y  (spline.py:_cubic_not_a_knot_matrix)
    double A[12][12] = {{0}}, R[12][12] = {{0}}, h[12];
    for (int i = 0; i + 1 < K; i++) h[i] = t[i + 1] - t[i];
    for (int i = 1; i + 1 < K; i++) {
      A[i][i - 1] = h[i - 1]; A[i][i] = 2 * (h[i - 1] + h[i]); A[i][i + 1] = h[i];
      R[i][i - 1] = 6 / h[i - 1]; R[i][i] = -6 / h[i - 1] - 6 / h[i]; R[i][i + 1] = 6 / h[i];
    }
    A[0][0] = h[1]; A[0][1] = -(h[0] + h[1]); A[0][2] = h[0];
    A[K - 1][K - 3] = h[K - 2]; A[K - 1][K - 2] = -(h[K - 3] + h[K - 2]); A[K - 1][K - 1] = h[K - 3];
    // Gaussian elimination with partial pivoting on [A | R]
    for (int c = 0; c < K; c++) {
      int p = c;
      for (int r = c + 1; r < K; r++) if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
      if (A[p][c] == 0) return 1;
      if (p != c) for (int j = 0; j < K; j++) { double s = A[c][j]; A[c][j] = A[p][j]; A[p][j] = s; s = R[c][j]; R[c][j] = R[p][j]; R[p][j] = s; }
      for (int r = c + 1; r < K; r++) {
        const double f = A[r][c] / A[c][c];
        if (f == 0) continue;
        for (int j = c; j < K; j++) A[r][j] -= f * A[c][j];
        for (int j = 0; j < K; j++) R[r][j] -= f * R[c][j];
      }
    }
    for (int j = 0; j < K; j++)
      for (int r = K - 1; r >= 0; r--) {
        double s = R[r][j];
        for (int c = r + 1; c < K; c++) s -= A[r][c] * C[c][j];
        C[r][j] = s / A[r][r];
      }
  }
  for (int i = 0; i < H; i++) {
    const double x = q[i];
    int seg = -1;  // index of the last knot <= x (np.searchsorted(t, x, side="right") - 1)
    while (seg + 1 < K && t[seg + 1] <= x) seg++;
    double* row = B + (size_t)i * K;
    if (order == 0) { row[seg < 0 ? 0 : seg] = 1.0; continue; }
    const bool below = seg < 0, above = x > t[K - 1];
    if (seg < 0) seg = 0;
    if (seg > K - 2) seg = K - 2;
    const double t0 = t[seg], t1 = t[seg + 1], h = t1 - t0;
    double a = (t1 - x) / h, b = (x - t0) / h;
    if (below) { a = 1.0; b = 0.0; }
    if (above) { a = 0.0; b = 1.0; }
    row[seg] = a; row[seg + 1] = b;
    if (order == 2) {
      const double ca = (a * a * a - a) * h * h / 6.0, cb = (b * b * b - b) * h * h / 6.0;
      for (int k = 0; k < K; k++) row[k] += ca * C[seg][k] + cb * C[seg + 1][k];
    }
  }
  return 0;
}

// ------------------------------------------------------------------ trace segments (controller.py:341-363)
// sens (ne, H, ns): sensor trajectories of the elite rollouts; cols (3 * nts): sensordata columns of the trace sensors in sensor order.
// out (ne * nts * (H - 1), 2, 3): segment i of sensor s of elite e sits at row (e * nts + s) * (H - 1) + i.
void trace_segments(const double* sens, int ne, int H, int ns, const int* cols, int nts, double* out) {
  const int size = H - 1;
  for (int e = 0; e < ne; e++)
    for (int s = 0; s < nts; s++) {
      double* o = out + (size_t)(e * nts + s) * size * 6;
      const int c0 = cols[3 * s], c1 = cols[3 * s + 1], c2 = cols[3 * s + 2];
      for (int i = 0; i < size; i++) {
        const double* a = sens + ((size_t)e * H + i) * ns;
        const double* b = a + ns;
        o[6 * i] = a[c0]; o[6 * i + 1] = a[c1]; o[6 * i + 2] = a[c2];
        o[6 * i + 3] = b[c0]; o[6 * i + 4] = b[c1]; o[6 * i + 5] = b[c2];
      }
    }
}

}  // namespace b2host
