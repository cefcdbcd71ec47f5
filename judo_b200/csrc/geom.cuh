// geom.cuh — small 3-D vector helpers and the box/sphere collision routines shared by the articulated-body kernels
// (leap.cuh, fr3.cuh).  Same algorithms (and operation order) as the oracle's collide_* functions.
#pragma once
#include "common.cuh"

namespace b2 {

// Rollout-steps in which a warp-per-rollout kernel found more contacts than its per-step buffer holds (leap: 30, fr3: 48) and had to
// drop the surplus.  Queried through b200mpc_contact_overflows(): truncation is never silent.  The counter is PER HANDLE: it lives in the
// 16 bytes the host allocates behind the handle's device-resident model table (leap_create / fr3_create).
#ifdef B2_HOST_SIM
static unsigned long long g_contact_overflow_sim;  // the CPU emulator passes bare model tables: one process-wide counter there
template <class Model>
__device__ __forceinline__ unsigned long long* contact_overflow_counter(const Model*) { return &g_contact_overflow_sim; }
#else
template <class Model>
__device__ __forceinline__ unsigned long long* contact_overflow_counter(const Model* m) {
  return reinterpret_cast<unsigned long long*>(const_cast<Model*>(m + 1));
}
#endif

// ------------------------------------------------------------------ small vector helpers (same op order as the oracle)
__device__ __forceinline__ double ldot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void lcross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ double lnorm3(const double* a) { return sqrt(ldot3(a, a)); }
__device__ __forceinline__ double lnormalize3(double* a) {
  double n = lnorm3(a);
  if (n < B2_MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; return 0; }
  const double inv = 1.0 / n;  // one reciprocal instead of three fp64 divisions (each is a ~40-instruction sequence)
  a[0] *= inv; a[1] *= inv; a[2] *= inv;
  return n;
}
__device__ __forceinline__ void lquat_mul(double* r, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
__device__ __forceinline__ void lquat_normalize(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < B2_MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  const double inv = 1.0 / n;
  q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}
__device__ __forceinline__ void lquat2mat(double* m, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
__device__ __forceinline__ void lmat_vec(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2], z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ void lmatT_vec(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2], z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ void lmat_mul(double* r, const double* a, const double* b) {
  double t[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
#pragma unroll
  for (int i = 0; i < 9; i++) r[i] = t[i];
}

// ------------------------------------------------------------------ collision (reduced geometry; same routines as the oracle)
struct LRaw { double dist, pos[3], normal[3]; };

// Scratch of one box-box call: the clipped polygon and its double buffer.  Callers choose where it lives: the leap kernel hands out a slice
// of its per-warp shared-memory work area (no local-memory stack frame, no DRAM traffic), fr3 a local array.
constexpr int L_POLY_MAX = 9;  // a quadrilateral clipped by four half-planes has at most 8 vertices
struct LBoxScratch { double poly[L_POLY_MAX][2], out[L_POLY_MAX][2]; };

// v[idx] for a runtime idx in 0..2 without indexing a local array (which would force it into local memory)
__device__ __forceinline__ double sel3(const double* v, int idx) { return idx == 0 ? v[0] : (idx == 1 ? v[1] : v[2]); }

__device__ __noinline__ int l_clip_poly(double (*poly)[2], double (*out)[2], int n, int axis, double sign, double lim) {
  int no = 0;
  for (int i = 0; i < n; i++) {
    const double* a = poly[i];
    const double* b = poly[(i + 1) % n];
    double da = sign * a[axis] - lim, db = sign * b[axis] - lim;
    if (da <= 0) { out[no][0] = a[0]; out[no][1] = a[1]; no++; }
    if ((da < 0 && db > 0) || (da > 0 && db < 0)) {
      double t = da / (da - db);
      out[no][0] = a[0] + t * (b[0] - a[0]); out[no][1] = a[1] + t * (b[1] - a[1]); no++;
    }
    if (no >= L_POLY_MAX - 1) break;
  }
  for (int i = 0; i < no; i++) { poly[i][0] = out[i][0]; poly[i][1] = out[i][1]; }
  return no;
}

// Sphere vs sphere (the oracle's collide_sphere_sphere): normal from geom1 to geom2, position midway between the surfaces.
__device__ inline int l_sphere_sphere(const double* p1, double r1, const double* p2, double r2, double margin, LRaw* out) {
  double dv[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  const double dn = lnorm3(dv);
  if (dn - r1 - r2 >= margin) return 0;
  if (dn < B2_MINVAL) { out->normal[0] = 1; out->normal[1] = 0; out->normal[2] = 0; }
  else {
#pragma unroll
    for (int k = 0; k < 3; k++) out->normal[k] = dv[k] / dn;
  }
  out->dist = dn - r1 - r2;
#pragma unroll
  for (int k = 0; k < 3; k++) out->pos[k] = p1[k] + out->normal[k] * (r1 + 0.5 * out->dist);
  return 1;
}

// Conservative pre-filters in front of the divergent contact-generation pass: they never reject a touching pair.
// box-box: the 6 face axes of the separating-axis test (slack 1e-9); sphere-box: the exact sphere-box distance.
__device__ inline bool l_box_box_may_touch(const double* p1, const double* m1, const double* s1, const double* p2, const double* m2, const double* s2) {
  const double dc[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  double t[3], R[3][3];
  lmatT_vec(t, m1, dc);  // centre of box 2 in the frame of box 1
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) R[i][j] = fabs(m1[i] * m2[j] + m1[3 + i] * m2[3 + j] + m1[6 + i] * m2[6 + j]) + 1e-12;
#pragma unroll
  for (int i = 0; i < 3; i++)
    if (fabs(t[i]) - (s1[i] + s2[0] * R[i][0] + s2[1] * R[i][1] + s2[2] * R[i][2]) >= 1e-9) return false;
  double t2[3];
  lmatT_vec(t2, m2, dc);
#pragma unroll
  for (int j = 0; j < 3; j++)
    if (fabs(t2[j]) - (s2[j] + s1[0] * R[0][j] + s1[1] * R[1][j] + s1[2] * R[2][j]) >= 1e-9) return false;
  return true;
}
__device__ inline bool l_sphere_box_may_touch(const double* ps, double rs, const double* pb, const double* mb, const double* sb) {
  const double dc[3] = {ps[0] - pb[0], ps[1] - pb[1], ps[2] - pb[2]};
  double t[3], d2 = 0;
  lmatT_vec(t, mb, dc);
#pragma unroll
  for (int k = 0; k < 3; k++) { const double ex = fabs(t[k]) - sb[k]; if (ex > 0) d2 += ex * ex; }
  const double rr = rs + 1e-9;
  return d2 < rr * rr;
}

__device__ __noinline__ int l_sphere_box(const double* ps, double rs, const double* pb, const double* mb, const double* sb, double margin, LRaw* out) {
  double rel[3], loc[3], cl[3];
#pragma unroll
  for (int k = 0; k < 3; k++) rel[k] = ps[k] - pb[k];
  lmatT_vec(loc, mb, rel);
  int inside = 1;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    cl[k] = fmin(fmax(loc[k], -sb[k]), sb[k]);
    if (cl[k] != loc[k]) inside = 0;
  }
  double nl[3], dist;
  if (!inside) {
    double dv[3] = {loc[0] - cl[0], loc[1] - cl[1], loc[2] - cl[2]};
    double dn = lnorm3(dv);
    if (dn - rs >= margin) return 0;
#pragma unroll
    for (int k = 0; k < 3; k++) nl[k] = -dv[k] / dn;
    dist = dn - rs;
  } else {
    int ax = 0; double best = 1e300;
#pragma unroll
    for (int k = 0; k < 3; k++) { double g = sb[k] - fabs(loc[k]); if (g < best) { best = g; ax = k; } }
#pragma unroll
    for (int k = 0; k < 3; k++) {  // (no runtime index into the local arrays: they stay in registers)
      nl[k] = k == ax ? (loc[k] >= 0 ? -1.0 : 1.0) : 0.0;
      if (k == ax) cl[k] = loc[k] >= 0 ? sb[k] : -sb[k];
    }
    dist = -best - rs;
  }
  lmat_vec(out->normal, mb, nl);
  double clw[3];
  lmat_vec(clw, mb, cl);
#pragma unroll
  for (int k = 0; k < 3; k++) out->pos[k] = pb[k] + clw[k] - out->normal[k] * (-0.5 * dist);
  out->dist = dist;
  return 1;
}

// overlap (optional): set to 1 when no separating axis was found (the boxes interpenetrate), 0 otherwise
__device__ __noinline__ int l_box_box(const double* p1, const double* m1, const double* s1, const double* p2, const double* m2, const double* s2,
                                double margin, LRaw* out, int maxout, LBoxScratch* scr, int* overlap = nullptr) {
  if (overlap) *overlap = 0;
  double R[3][3], AR[3][3], t[3], d12[3];
#pragma unroll
  for (int k = 0; k < 3; k++) d12[k] = p2[k] - p1[k];
  lmatT_vec(t, m1, d12);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      R[i][j] = m1[i] * m2[j] + m1[3 + i] * m2[3 + j] + m1[6 + i] * m2[6 + j];
      AR[i][j] = fabs(R[i][j]) + 1e-12;
    }
  double best = -1e300; int code = -1; double bsign = 1;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    double sep = fabs(t[i]) - (s1[i] + s2[0] * AR[i][0] + s2[1] * AR[i][1] + s2[2] * AR[i][2]);
    if (sep >= margin) return 0;
    if (sep > best) { best = sep; code = i; bsign = t[i] >= 0 ? 1 : -1; }
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    double tj = t[0] * R[0][j] + t[1] * R[1][j] + t[2] * R[2][j];
    double sep = fabs(tj) - (s2[j] + s1[0] * AR[0][j] + s1[1] * AR[1][j] + s1[2] * AR[2][j]);
    if (sep >= margin) return 0;
    if (sep > best) { best = sep; code = 3 + j; bsign = tj >= 0 ? 1 : -1; }
  }
  // edge-edge axes L = A_i x B_j in closed form (orthonormal frames): with i1 = i+1, i2 = i+2 (mod 3), same for j,
  //   d12 . L = t[i2] R[i1][j] - t[i1] R[i2][j],   r1 = s1[i1] |R[i2][j]| + s1[i2] |R[i1][j]|,   r2 = s2[j1] |R[i][j2]| + s2[j2] |R[i][j1]|,
  //   |L|^2 = R[i1][j]^2 + R[i2][j]^2  — one rsqrt per axis, no square root / division chain
  double ebest = -1e300; int ecode = -1; double esign = 1;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const double len2 = R[i1][j] * R[i1][j] + R[i2][j] * R[i2][j];
      if (len2 < 1e-16) continue;
      const double td = t[i2] * R[i1][j] - t[i1] * R[i2][j];
      const double rr = s1[i1] * fabs(R[i2][j]) + s1[i2] * fabs(R[i1][j]) + s2[j1] * fabs(R[i][j2]) + s2[j2] * fabs(R[i][j1]);
      const double sep = (fabs(td) - rr) * rsqrt(len2);
      if (sep >= margin) return 0;
      if (sep > ebest) { ebest = sep; ecode = 6 + 3 * i + j; esign = td >= 0 ? 1 : -1; }
    }
  if (overlap) *overlap = 1;
  if (ecode >= 0 && ebest > best + 1e-6 + 0.05 * fabs(best)) {
    int i = (ecode - 6) / 3, j = (ecode - 6) % 3;
    double n[3];
    {
      const double a1[3] = {m1[i], m1[3 + i], m1[6 + i]}, a2[3] = {m2[j], m2[3 + j], m2[6 + j]};
      lcross3(n, a1, a2);
      const double sc = esign * rsqrt(ldot3(n, n));
      n[0] *= sc; n[1] *= sc; n[2] *= sc;
    }
    double c1[3] = {p1[0], p1[1], p1[2]}, c2[3] = {p2[0], p2[1], p2[2]};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (k != i) { double a[3] = {m1[k], m1[3 + k], m1[6 + k]}; double sg = ldot3(a, n) > 0 ? 1 : -1;
#pragma unroll
        for (int q = 0; q < 3; q++) c1[q] += sg * s1[k] * a[q]; }
      if (k != j) { double a[3] = {m2[k], m2[3 + k], m2[6 + k]}; double sg = ldot3(a, n) > 0 ? -1 : 1;
#pragma unroll
        for (int q = 0; q < 3; q++) c2[q] += sg * s2[k] * a[q]; }
    }
    double u1[3] = {m1[i], m1[3 + i], m1[6 + i]}, u2[3] = {m2[j], m2[3 + j], m2[6 + j]}, w0[3];
#pragma unroll
    for (int k = 0; k < 3; k++) w0[k] = c1[k] - c2[k];
    double b = ldot3(u1, u2), dd = ldot3(u1, w0), e = ldot3(u2, w0), den = 1 - b * b;
    double sc = den > 1e-12 ? (b * e - dd) / den : 0, tc = den > 1e-12 ? (e - b * dd) / den : 0;
    sc = fmin(fmax(sc, -s1[i]), s1[i]); tc = fmin(fmax(tc, -s2[j]), s2[j]);
    out[0].dist = ebest;
#pragma unroll
    for (int k = 0; k < 3; k++) { out[0].normal[k] = n[k]; out[0].pos[k] = 0.5 * ((c1[k] + sc * u1[k]) + (c2[k] + tc * u2[k])); }
    return 1;
  }
  const double *pr, *mr, *sr, *pi, *mi, *si;
  int raxis; double nsign;
  const int ref_is_1 = code < 3;
  if (ref_is_1) { pr = p1; mr = m1; sr = s1; pi = p2; mi = m2; si = s2; raxis = code; nsign = bsign; }
  else { pr = p2; mr = m2; sr = s2; pi = p1; mi = m1; si = s1; raxis = code - 3; nsign = -bsign; }
  double nref[3] = {mr[raxis] * nsign, mr[3 + raxis] * nsign, mr[6 + raxis] * nsign};
  int iaxis = 0; double imin = 1e300, isign = 1;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double a[3] = {mi[k], mi[3 + k], mi[6 + k]}, dd = ldot3(a, nref);
    if (-fabs(dd) < imin) { imin = -fabs(dd); iaxis = k; isign = dd > 0 ? -1 : 1; }
  }
  const int iu = (iaxis + 1) % 3, iv = (iaxis + 2) % 3, ru = (raxis + 1) % 3, rv = (raxis + 2) % 3;
  double fc[3];
#pragma unroll
  for (int k = 0; k < 3; k++) fc[k] = pi[k] + isign * si[iaxis] * mi[3 * k + iaxis];
  double (*poly)[2] = scr->poly;
  int n = 4;
  // the incident face's corners (+u+v, -u+v, -u-v, +u-v) in the reference box's frame; corners 0, 1 and 3 span the face plane
  double l0[3] = {0, 0, 0}, l1[3] = {0, 0, 0}, l2[3] = {0, 0, 0};
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const double su = (c == 0 || c == 3) ? 1.0 : -1.0, sv = c < 2 ? 1.0 : -1.0;
    double vert[3], loc[3];
#pragma unroll
    for (int k = 0; k < 3; k++) vert[k] = fc[k] + su * si[iu] * mi[3 * k + iu] + sv * si[iv] * mi[3 * k + iv] - pr[k];
    lmatT_vec(loc, mr, vert);
    poly[c][0] = sel3(loc, ru); poly[c][1] = sel3(loc, rv);
    if (c == 0) { l0[0] = loc[0]; l0[1] = loc[1]; l0[2] = loc[2]; }
    if (c == 1) { l1[0] = loc[0]; l1[1] = loc[1]; l1[2] = loc[2]; }
    if (c == 3) { l2[0] = loc[0]; l2[1] = loc[1]; l2[2] = loc[2]; }
  }
  const double l0u = sel3(l0, ru), l0v = sel3(l0, rv), l0h = sel3(l0, raxis);
  const double e1u = sel3(l1, ru) - l0u, e1v = sel3(l1, rv) - l0v, e1h = sel3(l1, raxis) - l0h;
  const double e2u = sel3(l2, ru) - l0u, e2v = sel3(l2, rv) - l0v, e2h = sel3(l2, raxis) - l0h;
  const double det = e1u * e2v - e1v * e2u;
  const double idet = fabs(det) > 1e-14 ? 1.0 / det : 0.0;
  const double sru = sr[ru], srv = sr[rv];
  // clipping is the identity when the whole incident face projects inside the reference face (the common case on large faces)
  bool inside = true;
#pragma unroll
  for (int c = 0; c < 4; c++) inside = inside && fabs(poly[c][0]) <= sru && fabs(poly[c][1]) <= srv;
  if (!inside) {
    n = l_clip_poly(poly, scr->out, n, 0, 1, sru);
    if (n) n = l_clip_poly(poly, scr->out, n, 0, -1, sru);
    if (n) n = l_clip_poly(poly, scr->out, n, 1, 1, srv);
    if (n) n = l_clip_poly(poly, scr->out, n, 1, -1, srv);
  }
  int nc = 0;
  for (int c = 0; c < n && nc < maxout; c++) {
    double du = poly[c][0] - l0u, dv = poly[c][1] - l0v, h;
    if (fabs(det) > 1e-14) {
      double a = (du * e2v - dv * e2u) * idet, b = (e1u * dv - e1v * du) * idet;
      h = l0h + a * e1h + b * e2h;
    } else h = l0h;
    double dist = nsign * h - sr[raxis];
    if (dist >= margin) continue;
    double loc[3], wpt[3];
    const double hh = h - 0.5 * dist * nsign;
#pragma unroll
    for (int k = 0; k < 3; k++) loc[k] = k == ru ? poly[c][0] : (k == rv ? poly[c][1] : hh);
    lmat_vec(wpt, mr, loc);
    out[nc].dist = dist;
#pragma unroll
    for (int k = 0; k < 3; k++) { out[nc].pos[k] = pr[k] + wpt[k]; out[nc].normal[k] = ref_is_1 ? nref[k] : -nref[k]; }
    nc++;
  }
  return nc;
}

__device__ inline void l_make_frame(double* frame) {
  double* x = frame; double* y = frame + 3; double* z = frame + 6;
  lnormalize3(x);
  if (fabs(x[1]) < 0.5) { y[0] = 0; y[1] = 1; y[2] = 0; } else { y[0] = 0; y[1] = 0; y[2] = 1; }
  double dd = ldot3(x, y);
  for (int k = 0; k < 3; k++) y[k] -= dd * x[k];
  lnormalize3(y);
  lcross3(z, x, y);
}


// ------------------------------------------------------------------ signed box-box distance (distance sensors)
// Penetrating boxes: minus the penetration depth (largest of the 15 separating-axis values, exact for boxes);
// separated boxes: exact Euclidean distance by GJK on the Minkowski difference.  want_value = false stops after the SAT
// stage (callers that only need the sign): the return value is then a lower bound of the distance with the right sign.
__device__ __forceinline__ void l_box_support(const double* p, const double* m, const double* s, const double* dir, double* out) {
  out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const double ax[3] = {m[a], m[3 + a], m[6 + a]};
    const double sg = ldot3(ax, dir) >= 0 ? s[a] : -s[a];
    out[0] += sg * ax[0]; out[1] += sg * ax[1]; out[2] += sg * ax[2];
  }
}
__device__ __forceinline__ void l_cp3(double* d, const double* s) { d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; }
__device__ inline void l_closest_segment(double (*W)[3], int* n, double* v) {
  const double *a = W[0], *b = W[1];
  const double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
  double t = -ldot3(a, ab);
  const double den = ldot3(ab, ab);
  if (t <= 0 || den <= 0) { *n = 1; l_cp3(v, a); return; }
  if (t >= den) { l_cp3(W[0], b); *n = 1; l_cp3(v, W[0]); return; }
  t /= den;
  for (int k = 0; k < 3; k++) v[k] = a[k] + t * ab[k];
}
__device__ __noinline__ void l_closest_triangle(double (*W)[3], int* n, double* v) {
  double a[3], b[3], c[3], ab[3], ac[3], bc[3];
  l_cp3(a, W[0]); l_cp3(b, W[1]); l_cp3(c, W[2]);
  for (int k = 0; k < 3; k++) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; bc[k] = c[k] - b[k]; }
  const double d1 = -ldot3(ab, a), d2 = -ldot3(ac, a);
  if (d1 <= 0 && d2 <= 0) { *n = 1; l_cp3(v, a); return; }
  const double d3 = -ldot3(ab, b), d4 = -ldot3(ac, b);
  if (d3 >= 0 && d4 <= d3) { l_cp3(W[0], b); *n = 1; l_cp3(v, b); return; }
  const double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) {
    const double t = d1 / (d1 - d3);
    *n = 2;
    for (int k = 0; k < 3; k++) v[k] = a[k] + t * ab[k];
    return;
  }
  const double d5 = -ldot3(ab, c), d6 = -ldot3(ac, c);
  if (d6 >= 0 && d5 <= d6) { l_cp3(W[0], c); *n = 1; l_cp3(v, c); return; }
  const double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) {
    const double t = d2 / (d2 - d6);
    l_cp3(W[1], c); *n = 2;
    for (int k = 0; k < 3; k++) v[k] = a[k] + t * ac[k];
    return;
  }
  const double va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    const double t = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    l_cp3(W[0], b); l_cp3(W[1], c); *n = 2;
    for (int k = 0; k < 3; k++) v[k] = b[k] + t * bc[k];
    return;
  }
  const double den = 1.0 / (va + vb + vc), tv = vb * den, tw = vc * den;
  *n = 3;
  for (int k = 0; k < 3; k++) v[k] = a[k] + ab[k] * tv + ac[k] * tw;
}
// returns 1 when the origin is inside the tetrahedron (the sets intersect)
__device__ __noinline__ int l_closest_tetrahedron(double (*W)[3], int* n, double* v) {
  const int F[4][4] = {{0, 1, 2, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {1, 3, 2, 0}};
  double best = 1e300, bv[3] = {0, 0, 0}, BW[3][3], P[4][3];
  int bn = 0, outside_any = 0;
  for (int i = 0; i < 4; i++) l_cp3(P[i], W[i]);
  for (int f = 0; f < 4; f++) {
    const double *a = P[F[f][0]], *b = P[F[f][1]], *c = P[F[f][2]], *dd = P[F[f][3]];
    double ab[3], ac[3], nrm[3];
    for (int k = 0; k < 3; k++) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; }
    lcross3(nrm, ab, ac);
    const double so = -ldot3(a, nrm);
    const double ad[3] = {dd[0] - a[0], dd[1] - a[1], dd[2] - a[2]};
    const double sd = ldot3(ad, nrm);
    if (!(so * sd > 0)) outside_any = 1;  // all four faces are evaluated: flat tetrahedra make the sign tests alone unreliable
    double T[3][3], tv[3];
    int tn = 3;
    l_cp3(T[0], a); l_cp3(T[1], b); l_cp3(T[2], c);
    l_closest_triangle(T, &tn, tv);
    const double dist = ldot3(tv, tv);
    if (dist < best) { best = dist; bn = tn; l_cp3(bv, tv); for (int i = 0; i < 3; i++) l_cp3(BW[i], T[i]); }
  }
  if (!outside_any) return 1;
  *n = bn; l_cp3(v, bv);
  for (int i = 0; i < bn; i++) l_cp3(W[i], BW[i]);
  return 0;
}
__device__ __noinline__ double l_box_box_distance(const double* p1, const double* m1, const double* s1, const double* p2, const double* m2,
                                                const double* s2, double cutoff, bool want_value) {
  double d12[3], best = -1e300, R[3][3], t[3];
  for (int k = 0; k < 3; k++) d12[k] = p2[k] - p1[k];
  lmatT_vec(t, m1, d12);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) R[i][j] = m1[i] * m2[j] + m1[3 + i] * m2[3 + j] + m1[6 + i] * m2[6 + j];
  // the 15 separating-axis values in closed form (see l_box_box)
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double sep = fabs(t[i]) - (s1[i] + s2[0] * fabs(R[i][0]) + s2[1] * fabs(R[i][1]) + s2[2] * fabs(R[i][2]));
    if (sep > best) best = sep;
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const double tj = t[0] * R[0][j] + t[1] * R[1][j] + t[2] * R[2][j];
    const double sep = fabs(tj) - (s2[j] + s1[0] * fabs(R[0][j]) + s1[1] * fabs(R[1][j]) + s1[2] * fabs(R[2][j]));
    if (sep > best) best = sep;
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const double len2 = R[i1][j] * R[i1][j] + R[i2][j] * R[i2][j];
      if (len2 < 1e-16) continue;
      const double td = t[i2] * R[i1][j] - t[i1] * R[i2][j];
      const double rr = s1[i1] * fabs(R[i2][j]) + s1[i2] * fabs(R[i1][j]) + s2[j1] * fabs(R[i][j2]) + s2[j2] * fabs(R[i][j1]);
      const double sep = (fabs(td) - rr) * rsqrt(len2);
      if (sep > best) best = sep;
    }
  if (best <= 0) return best;
  if (best >= cutoff) return cutoff;
  if (!want_value) return best;
  double W[4][3], v[3] = {-d12[0], -d12[1], -d12[2]}, vv_prev = 1e300;
  int n = 0;
  if (ldot3(v, v) < 1e-30) { v[0] = 1; v[1] = v[2] = 0; }
  for (int it = 0; it < 64; it++) {
    const double nd[3] = {-v[0], -v[1], -v[2]};
    double a[3], b[3], w[3];
    l_box_support(p1, m1, s1, nd, a);
    l_box_support(p2, m2, s2, v, b);
    for (int k = 0; k < 3; k++) w[k] = a[k] - b[k];
    const double vv = ldot3(v, v);
    if (n > 0 && vv - ldot3(v, w) <= 1e-13 * vv) break;
    if (n > 0 && vv >= vv_prev * (1 - 1e-15)) break;  // parallel faces: the simplex cycles between equivalent vertices; v no longer shortens
    if (n > 0) vv_prev = vv;  // (the initial v, the centre difference, is not a point of the simplex yet)
    bool dup = false;
    for (int i = 0; i < n; i++) if (W[i][0] == w[0] && W[i][1] == w[1] && W[i][2] == w[2]) dup = true;
    if (dup) break;
    l_cp3(W[n++], w);
    if (n == 1) l_cp3(v, W[0]);
    else if (n == 2) l_closest_segment(W, &n, v);
    else if (n == 3) l_closest_triangle(W, &n, v);
    else if (l_closest_tetrahedron(W, &n, v)) return best;
    if (ldot3(v, v) < 1e-30) return best;
  }
  const double dist = lnorm3(v);
  return dist < cutoff ? dist : cutoff;
}

}  // namespace b2
