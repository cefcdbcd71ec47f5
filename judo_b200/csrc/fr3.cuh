// fr3.cuh — fr3_pick: warp-per-rollout reduced articulated-body integrator + phase-switched cost.
//
// One warp owns one rollout: 16 qpos / 15 dofs (free object + 7 arm hinges + 2 finger slides), state and all per-step work
// arrays in shared memory.  Every stage of MuJoCo's mj_step that judo/models/xml/fr3_pick.xml switches on is restated and spread
// over the 32 lanes: kinematics of the serial chain, composite-inertia mass matrix (lane per arm dof), RNE bias forces,
// collision + signed distances of the 21 box pairs (lane per pair; GJK for the distance sensors), constraint rows
// (joint equality, friction loss, joint limits, pyramidal contacts kept as 3 frame rows + a 3x3 weight per contact),
// mj_makeImpedance, the primal Newton solver (dense 15x15 Hessian, warp Cholesky, register line search), position servos with
// joint-level force clamps, implicitfast integration.  The collision GEOMETRY is reduced to the model's box geoms
// (judo_b200/tasks/fr3_pick.py); the constraint/solver pipeline is the full one.
#pragma once
#include "epilogue.cuh"
#include "geom.cuh"
#include "sampling.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

namespace b2 {

// solver-loop reciprocals / roots without the library's special-case branches (common.cuh); B2_FR3_LIBMATH restores the library calls
#ifdef B2_FR3_LIBMATH
#define F_RSQRT(x) rsqrt(x)
#define F_SQRT(x) sqrt(x)
#define F_RCP(x) (1.0 / (x))
#else
#define F_RSQRT(x) fast_rsqrt_pos(x)
#define F_SQRT(x) fast_sqrt_nonneg(x)
#define F_RCP(x) fast_rcp_pos(x)
#endif

constexpr int FR_NQ = 16, FR_NV = 15, FR_NU = 8, FR_NS = 14, FR_NX = 31, FR_NCOST = 23;
constexpr int FR_NTRACE = 6;  // doubles per step kept by the fused kernel's trace capture: trace_object, trace_grasp_site
constexpr int FB = 11;        // moving bodies: 0 object, 1..7 links, 8 hand (welded to link 7), 9 left finger, 10 right finger
constexpr int FNA = 9;        // arm dofs (global dof = 6 + j): 7 hinges, 2 finger slides
constexpr int FNPAD = 10, FNPAIR = 21;  // pairs: 0 table-object, 1..10 table-pad, 11..20 object-pad
constexpr int FMAXCON = 48;   // contacts kept per step (4 pyramid rows each)
constexpr int FNS = 19;       // scalar rows in FIXED slots: 0 equality, 1 + j friction loss of arm dof j, 10 + j joint limit of arm dof j
constexpr int FMAXS = 20;
constexpr int FLD = 16;       // leading dimension of the dense matrices
#ifndef B2_FULLMASK
#define B2_FULLMASK 0xffffffffu
#endif

// All-double POD; field order == judo_b200/tasks/fr3_pick.py:fr3_consts.
struct Fr3Model {
  double dt, gravity[3], impratio, tolerance, ls_tolerance, meaninertia, iterations, ls_iterations;
  double base_pos[3], base_quat[4];
  double body_pos[FB][3], body_quat[FB][4], body_ipos[FB][3], body_imat[FB][9], body_mass[FB], body_inertia[FB][3], body_invw[FB];
  double jnt_axis[FB][3];
  double obj_inertia[3];
  double dof_damping[FR_NV], dof_armature[FR_NV], dof_frictionloss[FR_NV], dof_invw[FR_NV];
  double fr_solref[2], fr_solimp[5];
  double lim_lo[FNA], lim_hi[FNA], lim_margin, lim_solref[2], lim_solimp[5];
  double eq_solref[2], eq_solimp[5];
  double kp[FR_NU], kv[FR_NU], ctrl_lo[FR_NU], ctrl_hi[FR_NU];
  double frc_limited[FNA], frc_lo[FNA], frc_hi[FNA];
  double table_pos[3], table_mat[9], table_size[3], obj_size[3];
  double pad_body[FNPAD], pad_pos[FNPAD][3], pad_size[FNPAD][3];
  double con_mu[3], con_solref[3][2], con_solimp[3][5];
  double site_pos[3];
  double cutoff;
};

// per-warp shared-memory work area (31.7 KB: seven rollouts per SM, i.e. N = 1024 is a single wave on 148 SMs)
struct Fr3Work {
  double qpos[FR_NQ], qvel[FR_NV], warm[FR_NV], ctrl[FR_NU];
  double xpos[FB][3], xmat[FB][9], xipos[FB][3], Iw[FB][6];
  double anchor[FNA][3], axis[FNA][3];
  double M[FNA][FNA + 1], Ld[FLD];            // arm mass matrix; reciprocal pivots of the arm factor
  double H[FR_NV][FLD], Hd[FLD];              // Newton Hessian / factor; outside the solver rows 0..8 hold the arm factor of M (+ h D)
  double qfrc_bias[FR_NV], qfrc_smooth[FR_NV], qacc_smooth[FR_NV], qacc[FR_NV], qfrc_constraint[FR_NV];
  double Ma[FR_NV], grad[FR_NV], search[FR_NV], Mv[FR_NV];
  double sD[FMAXS], sR[FMAXS], saref[FMAXS], sjar[FMAXS], sforce[FMAXS], sfl[FMAXS], ssign[FMAXS];
  int sstate[FMAXS];
  double cdist[FMAXCON];
  double cgeo[FMAXCON][12];     // while rows are built: contact point (3) + frame (9); in the solver: aref (3), jar (3), force (3)
  double cJ[FMAXCON][3][FR_NV]; // frame-row Jacobians (normal, tangent 1, tangent 2)
  double cD[FMAXCON], cmu[FMAXCON];
  int ccls[FMAXCON], cbody[FMAXCON];
  double pdist[FNPAIR + 3], sens[FR_NS];
  int ncon;
  double cost, gauss;
};
constexpr int CG_AREF = 0, CG_JAR = 3, CG_FORCE = 6;  // solver-phase layout of cgeo[c]

// lower-triangle entry e -> (row, column) of the 15x15 Hessian, packed as row * 16 + column
__device__ const unsigned char FR3_TRI[FR_NV * (FR_NV + 1) / 2] = {
    0x00, 0x10, 0x11, 0x20, 0x21, 0x22, 0x30, 0x31, 0x32, 0x33, 0x40, 0x41, 0x42, 0x43, 0x44, 0x50, 0x51, 0x52, 0x53, 0x54, 0x55, 0x60, 0x61, 0x62,
    0x63, 0x64, 0x65, 0x66, 0x70, 0x71, 0x72, 0x73, 0x74, 0x75, 0x76, 0x77, 0x80, 0x81, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x90, 0x91, 0x92,
    0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0xa0, 0xa1, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb0, 0xb1, 0xb2, 0xb3, 0xb4, 0xb5,
    0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xbb, 0xc0, 0xc1, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xcb, 0xcc, 0xd0, 0xd1, 0xd2, 0xd3, 0xd4,
    0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xdb, 0xdc, 0xdd, 0xe0, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xeb, 0xec, 0xed, 0xee};

// optional phase timers (B200MPC_FR3_PROF=1): clock64 deltas accumulated by lane 0
__device__ unsigned long long g_fr3_prof[16];
#define FPROF_T() (prof ? clock64() : 0)
#define FPROF_ADD(slot, t0) do { if (prof && lane == 0) atomicAdd(&g_fr3_prof[slot], (unsigned long long)(clock64() - (t0))); } while (0)

enum { FST_SATISFIED = 0, FST_QUADRATIC = 1, FST_LINEARNEG = 2, FST_LINEARPOS = 3, FST_INACTIVE = 4 };

__device__ __forceinline__ double fwsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(B2_FULLMASK, v, o);
  return v;
}
// two independent sums reduced together: the shuffle rounds interleave, so the pair costs the latency of one reduction
__device__ __forceinline__ void fwsum2(double& a, double& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ta = __shfl_xor_sync(B2_FULLMASK, a, o), tb = __shfl_xor_sync(B2_FULLMASK, b, o);
    a += ta; b += tb;
  }
}

// ------------------------------------------------------------------ kinematics (mj_kinematics + mj_comPos)
// The arm is a serial chain, but composing rigid transforms is associative: lane c < 9 builds the parent-relative transform of chain
// body 1 + c (links 1..7, hand, left finger), an inclusive warp scan (4 shuffle rounds) turns them into world transforms; lane 9 hangs the
// right finger off the hand's result; lane 10 is the free object.  (Rotation matrices are composed instead of quaternions: the
// difference to the oracle's sequential quaternion recursion is rounding-level.)
__device__ __forceinline__ void fr3_xf_compose(double* R, double* p, const double* Ra, const double* pa) {  // (R, p) <- (Ra, pa) o (R, p)
  double t[3], Rn[9];
  lmat_vec(t, Ra, p);
  lmat_mul(Rn, Ra, R);
#pragma unroll
  for (int k = 0; k < 3; k++) p[k] = pa[k] + t[k];
#pragma unroll
  for (int k = 0; k < 9; k++) R[k] = Rn[k];
}
__device__ inline void fr3_kinematics(const Fr3Model* __restrict__ m, Fr3Work* W, int lane) {
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
  const int b = lane + 1;  // body of this lane (lanes 0..9)
  if (lane < 10) {
    double quat[4] = {m->body_quat[b][0], m->body_quat[b][1], m->body_quat[b][2], m->body_quat[b][3]};
    if (b <= 7) {  // hinge about the local axis through the body origin
      double sn, cs, dq[4], nq[4];
      sincos(0.5 * W->qpos[7 + lane], &sn, &cs);
      dq[0] = cs; dq[1] = sn * m->jnt_axis[b][0]; dq[2] = sn * m->jnt_axis[b][1]; dq[3] = sn * m->jnt_axis[b][2];
      lquat_mul(nq, quat, dq);
#pragma unroll
      for (int k = 0; k < 4; k++) quat[k] = nq[k];
    }
    lquat_normalize(quat);
    lquat2mat(R, quat);
#pragma unroll
    for (int k = 0; k < 3; k++) p[k] = m->body_pos[b][k];
    if (b >= 9) {  // slide along the local axis
      double ax[3];
      lmat_vec(ax, R, m->jnt_axis[b]);
      const double q = W->qpos[7 + b - 2];
#pragma unroll
      for (int k = 0; k < 3; k++) p[k] += ax[k] * q;
    }
    if (lane == 0) {  // the chain hangs off the (static) base frame
      double Rb[9];
      lquat2mat(Rb, m->base_quat);
      fr3_xf_compose(R, p, Rb, m->base_pos);
    }
  } else if (lane == 10) {
    lquat_normalize(W->qpos + 3);
#pragma unroll
    for (int k = 0; k < 3; k++) W->xpos[0][k] = W->qpos[k];
    lquat2mat(W->xmat[0], W->qpos + 3);
  }
  // inclusive scan over the chain lanes 0..8
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) {
    double Ra[9], pa[3];
#pragma unroll
    for (int k = 0; k < 9; k++) Ra[k] = __shfl_up_sync(B2_FULLMASK, R[k], d);
#pragma unroll
    for (int k = 0; k < 3; k++) pa[k] = __shfl_up_sync(B2_FULLMASK, p[k], d);
    if (lane >= d && lane < 9) fr3_xf_compose(R, p, Ra, pa);
  }
  {  // right finger: parent = hand = lane 7
    double Ra[9], pa[3];
#pragma unroll
    for (int k = 0; k < 9; k++) Ra[k] = __shfl_sync(B2_FULLMASK, R[k], 7);
#pragma unroll
    for (int k = 0; k < 3; k++) pa[k] = __shfl_sync(B2_FULLMASK, p[k], 7);
    if (lane == 9) fr3_xf_compose(R, p, Ra, pa);
  }
  if (lane < 10) {
#pragma unroll
    for (int k = 0; k < 3; k++) W->xpos[b][k] = p[k];
#pragma unroll
    for (int k = 0; k < 9; k++) W->xmat[b][k] = R[k];
    if (b != 8) {  // joint anchor (= body origin) and world axis of arm dof j
      const int j = b <= 7 ? b - 1 : b - 2;
      double ax[3];
      lmat_vec(ax, R, m->jnt_axis[b]);
#pragma unroll
      for (int k = 0; k < 3; k++) { W->anchor[j][k] = p[k]; W->axis[j][k] = ax[k]; }
    }
  }
  __syncwarp();
  if (lane < FB) {
    const int bb = lane;
    double t[3], im[9];
    lmat_vec(t, W->xmat[bb], m->body_ipos[bb]);
#pragma unroll
    for (int k = 0; k < 3; k++) W->xipos[bb][k] = W->xpos[bb][k] + t[k];
    lmat_mul(im, W->xmat[bb], m->body_imat[bb]);
    int e = 0;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = r; c < 3; c++) {
        double sacc = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) sacc += im[3 * r + k] * m->body_inertia[bb][k] * im[3 * c + k];
        W->Iw[bb][e++] = sacc;  // xx xy xz yy yz zz
      }
  }
  __syncwarp();
}

__device__ __forceinline__ void fIw_mul(double* r, const double* I6, const double* v) {
  r[0] = I6[0] * v[0] + I6[1] * v[1] + I6[2] * v[2];
  r[1] = I6[1] * v[0] + I6[3] * v[1] + I6[4] * v[2];
  r[2] = I6[2] * v[0] + I6[4] * v[1] + I6[5] * v[2];
}

// arm dof j moves bodies [fr3_sub_lo(j), fr3_sub_hi(j)]
__device__ __forceinline__ int fr3_sub_lo(int j) { return j < 7 ? j + 1 : j + 2; }
__device__ __forceinline__ int fr3_sub_hi(int j) { return j < 7 ? 10 : j + 2; }

// ------------------------------------------------------------------ bias forces of the arm (mj_rne with zero acceleration)
// Newton-Euler in world coordinates (base: w = 0, a = -g).  Down the chain every quantity is a running sum whose increments depend
// only on the parent's totals, so the forward pass is three warp prefix sums over the chain lanes 0..8 (body 1 + lane; lane 9 = right
// finger, child of the hand on lane 7):  w_b = w_p + a qd;  alpha_b = alpha_p + w_p x (a qd);
// acc_b (body origin) = acc_p + alpha_p x r + w_p x (w_p x r) [+ 2 w_p x (a qd) for a slide],  r = x_b - x_p.
// Then lane per body: wrench about the world origin; lane per dof: projection of the subtree wrench on the joint axis.
__device__ __forceinline__ void fr3_scan3(double* v, int lane) {  // inclusive prefix sum of a 3-vector over lanes 0..8
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) {
#pragma unroll
    for (int k = 0; k < 3; k++) { const double t = __shfl_up_sync(B2_FULLMASK, v[k], d); if (lane >= d && lane < 9) v[k] += t; }
  }
}
__device__ inline void fr3_rne_bias(const Fr3Model* __restrict__ m, Fr3Work* W, int lane) {
  const int b = lane + 1;
  const bool on = lane < 10, hinge = lane < 7, slide = lane == 8 || lane == 9;
  const int j = hinge ? lane : lane - 1;  // arm dof of this body (hand: unused)
  double u[3] = {0, 0, 0};                // joint velocity a qd (hinge: angular, slide: linear)
  if (on && (hinge || slide)) {
    const double qd = W->qvel[6 + j];
#pragma unroll
    for (int k = 0; k < 3; k++) u[k] = W->axis[j][k] * qd;
  }
  double w[3] = {0, 0, 0};
  if (hinge) { w[0] = u[0]; w[1] = u[1]; w[2] = u[2]; }
  fr3_scan3(w, lane);
  double wp[3];   // parent's angular velocity
  {
    const double h7[3] = {__shfl_sync(B2_FULLMASK, w[0], 7), __shfl_sync(B2_FULLMASK, w[1], 7), __shfl_sync(B2_FULLMASK, w[2], 7)};
    if (lane == 9) { w[0] = h7[0]; w[1] = h7[1]; w[2] = h7[2]; }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) wp[k] = hinge ? w[k] - u[k] : w[k];
  double al[3] = {0, 0, 0}, dal[3] = {0, 0, 0};
  if (hinge) lcross3(dal, wp, u);
#pragma unroll
  for (int k = 0; k < 3; k++) al[k] = dal[k];
  fr3_scan3(al, lane);
  {
    const double h7[3] = {__shfl_sync(B2_FULLMASK, al[0], 7), __shfl_sync(B2_FULLMASK, al[1], 7), __shfl_sync(B2_FULLMASK, al[2], 7)};
    if (lane == 9) { al[0] = h7[0]; al[1] = h7[1]; al[2] = h7[2]; }
  }
  double alp[3];  // parent's angular acceleration
#pragma unroll
  for (int k = 0; k < 3; k++) alp[k] = al[k] - dal[k];
  double ac[3] = {0, 0, 0};
  if (on) {
    const double* xp = lane == 0 ? m->base_pos : (lane == 9 ? W->xpos[8] : W->xpos[b - 1]);
    const double r[3] = {W->xpos[b][0] - xp[0], W->xpos[b][1] - xp[1], W->xpos[b][2] - xp[2]};
    double t[3], t2[3], t3[3];
    lcross3(t, wp, r); lcross3(t2, wp, t); lcross3(t3, alp, r);
#pragma unroll
    for (int k = 0; k < 3; k++) ac[k] = t3[k] + t2[k];
    if (slide) {
      lcross3(t, wp, u);
#pragma unroll
      for (int k = 0; k < 3; k++) ac[k] += 2 * t[k];
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) ac[k] -= m->gravity[k];
    }
  }
  double dac[3] = {ac[0], ac[1], ac[2]};
  fr3_scan3(ac, lane);
  {
    const double h7[3] = {__shfl_sync(B2_FULLMASK, ac[0], 7), __shfl_sync(B2_FULLMASK, ac[1], 7), __shfl_sync(B2_FULLMASK, ac[2], 7)};
    if (lane == 9) { ac[0] = h7[0] + dac[0]; ac[1] = h7[1] + dac[1]; ac[2] = h7[2] + dac[2]; }
  }
  // wrench of this body about the world origin -> scratch rows of H (free outside the solver)
  if (on) {
    double rc[3], t[3], t2[3], t3[3], F[3], Iwa[3], Iww[3], n[3];
#pragma unroll
    for (int k = 0; k < 3; k++) rc[k] = W->xipos[b][k] - W->xpos[b][k];
    lcross3(t, w, rc); lcross3(t2, w, t); lcross3(t3, al, rc);
#pragma unroll
    for (int k = 0; k < 3; k++) F[k] = m->body_mass[b] * (ac[k] + t3[k] + t2[k]);
    fIw_mul(Iwa, W->Iw[b], al);
    fIw_mul(Iww, W->Iw[b], w);
    lcross3(t, w, Iww);
    lcross3(n, W->xipos[b], F);
#pragma unroll
    for (int k = 0; k < 3; k++) { W->H[b][k] = F[k]; W->H[b][3 + k] = Iwa[k] + t[k] + n[k]; }
  }
  __syncwarp();
  if (lane < FNA) {
    const int jj = lane, lo = fr3_sub_lo(jj), hi = fr3_sub_hi(jj);
    double F[3] = {0, 0, 0}, N0[3] = {0, 0, 0};
    for (int bb = lo; bb <= hi; bb++) {
#pragma unroll
      for (int k = 0; k < 3; k++) { F[k] += W->H[bb][k]; N0[k] += W->H[bb][3 + k]; }
    }
    if (jj < 7) {
      double t[3];
      lcross3(t, W->anchor[jj], F);
      W->qfrc_bias[6 + jj] = W->axis[jj][0] * (N0[0] - t[0]) + W->axis[jj][1] * (N0[1] - t[1]) + W->axis[jj][2] * (N0[2] - t[2]);
    } else W->qfrc_bias[6 + jj] = ldot3(W->axis[jj], F);
  }
  __syncwarp();
}

// ------------------------------------------------------------------ mass matrix (mj_crb) and bias forces (mj_rne)
// lanes 0..8: column j of the arm mass matrix from the composite inertia of the subtree dof j moves;
// lane 9: RNE with zero acceleration down and up the arm; lane 10: the free object.
__device__ inline void fr3_mass_and_bias(const Fr3Model* __restrict__ m, Fr3Work* W, int lane) {
  if (lane < FNA) {
    const int j = lane, lo = fr3_sub_lo(j), hi = fr3_sub_hi(j);
    double mc = 0, c[3] = {0, 0, 0};
    for (int b = lo; b <= hi; b++) {
      mc += m->body_mass[b];
#pragma unroll
      for (int k = 0; k < 3; k++) c[k] += m->body_mass[b] * W->xipos[b][k];
    }
#pragma unroll
    for (int k = 0; k < 3; k++) c[k] /= mc;
    double Ic[6] = {0, 0, 0, 0, 0, 0};
    for (int b = lo; b <= hi; b++) {
      const double d[3] = {W->xipos[b][0] - c[0], W->xipos[b][1] - c[1], W->xipos[b][2] - c[2]};
      const double mb = m->body_mass[b], dd = ldot3(d, d);
      Ic[0] += W->Iw[b][0] + mb * (dd - d[0] * d[0]); Ic[1] += W->Iw[b][1] - mb * d[0] * d[1]; Ic[2] += W->Iw[b][2] - mb * d[0] * d[2];
      Ic[3] += W->Iw[b][3] + mb * (dd - d[1] * d[1]); Ic[4] += W->Iw[b][4] - mb * d[1] * d[2];
      Ic[5] += W->Iw[b][5] + mb * (dd - d[2] * d[2]);
    }
    // momentum of the composite body under unit velocity of dof j: linear h, angular (about c) Lc
    double h[3], Lc[3] = {0, 0, 0};
    if (j < 7) {
      const double r[3] = {c[0] - W->anchor[j][0], c[1] - W->anchor[j][1], c[2] - W->anchor[j][2]};
      double t[3];
      lcross3(t, W->axis[j], r);
#pragma unroll
      for (int k = 0; k < 3; k++) h[k] = mc * t[k];
      fIw_mul(Lc, Ic, W->axis[j]);
    } else {
#pragma unroll
      for (int k = 0; k < 3; k++) h[k] = mc * W->axis[j][k];
    }
    for (int i = 0; i < FNA; i++) {
      double v = 0;
      if (i < 7 && i <= j) {  // hinge ancestor (or j itself)
        const double r[3] = {c[0] - W->anchor[i][0], c[1] - W->anchor[i][1], c[2] - W->anchor[i][2]};
        double t[3];
        lcross3(t, r, h);
        v = W->axis[i][0] * (Lc[0] + t[0]) + W->axis[i][1] * (Lc[1] + t[1]) + W->axis[i][2] * (Lc[2] + t[2]);
      } else if (i == j) v = ldot3(W->axis[i], h);
      else continue;  // not an ancestor: filled by symmetry (or zero between the two fingers)
      if (i == j) v += m->dof_armature[6 + j];
      W->M[i][j] = v;
      W->M[j][i] = v;
    }
    if (j == 7) { W->M[7][8] = 0; W->M[8][7] = 0; }
  }
  if (lane == 10) {
    const double wl[3] = {W->qvel[3], W->qvel[4], W->qvel[5]};
    const double Iwl[3] = {m->obj_inertia[0] * wl[0], m->obj_inertia[1] * wl[1], m->obj_inertia[2] * wl[2]};
    double g[3];
    lcross3(g, wl, Iwl);
#pragma unroll
    for (int k = 0; k < 3; k++) { W->qfrc_bias[k] = -m->body_mass[0] * m->gravity[k]; W->qfrc_bias[3 + k] = g[k]; }
  }
  fr3_rne_bias(m, W, lane);
}

// y_i = (M x)_i: object block is diagonal (mass, principal inertia in the body frame), arm block dense
__device__ __forceinline__ double fr3_mulM_row(const Fr3Model* __restrict__ m, const Fr3Work* W, const double* x, int i) {
  if (i < 3) return m->body_mass[0] * x[i];
  if (i < 6) return m->obj_inertia[i - 3] * x[i];
  const double* row = W->M[i - 6];
  double s = 0;
#pragma unroll
  for (int k = 0; k < FNA; k++) s += row[k] * x[6 + k];
  return s;
}

// ------------------------------------------------------------------ dense warp Cholesky: lane i owns row i
// In place on the lower triangle of A (n x n, leading dimension FLD); dinv[k] = 1 / L_kk.
__device__ inline void warp_chol(double (*A)[FLD], double* dinv, int n, int lane) {
  for (int k = 0; k < n; k++) {
    double d = A[k][k];
    if (d < B2_MINVAL) d = B2_MINVAL;
    const double rs = F_RSQRT(d);  // (d is clamped to B2_MINVAL above: positive, normal)
    double l = 0;
    if (lane == k) dinv[k] = rs;
    if (lane > k && lane < n) { l = A[lane][k] * rs; A[lane][k] = l; }
    __syncwarp();
    if (lane > k && lane < n)
      for (int j = k + 1; j <= lane; j++) A[lane][j] -= l * A[j][k];
    __syncwarp();
  }
}
// x <- (L L^T)^-1 x with x_lane in a register (lanes >= n pass 0); column-oriented substitutions, one shuffle per column
__device__ inline double warp_chol_solve(const double (*A)[FLD], const double* dinv, int n, double x, int lane) {
  const double di = lane < n ? dinv[lane] : 0.0;
  for (int k = 0; k < n; k++) {
    const double yk = __shfl_sync(B2_FULLMASK, x * di, k);
    if (lane == k) x = yk;
    else if (lane > k && lane < n) x -= A[lane][k] * yk;
  }
  for (int k = n - 1; k >= 0; k--) {
    const double xk = __shfl_sync(B2_FULLMASK, x * di, k);
    if (lane == k) x = xk;
    else if (lane < k) x -= A[k][lane] * xk;
  }
  return x;
}

// Same for a BLOCK-DIAGONAL matrix made of an object block (rows 0..5) and an arm block (rows 6..14): the two blocks are independent, so
// their columns are eliminated side by side — 9 sequential columns instead of 15.
__device__ inline void warp_chol2(double (*A)[FLD], double* dinv, int lane) {
  const int base = lane < 6 ? 0 : 6, nloc = lane < 6 ? 6 : (lane < FR_NV ? 9 : 0);
  for (int k = 0; k < 9; k++) {
    const bool act = k < nloc;
    const int col = act ? base + k : 0;
    double d = A[col][col];
    if (d < B2_MINVAL) d = B2_MINVAL;
    const double rs = F_RSQRT(d);  // (d is clamped to B2_MINVAL above: positive, normal)
    double l = 0;
    if (act && lane == col) dinv[col] = rs;
    if (act && lane > col) { l = A[lane][col] * rs; A[lane][col] = l; }
    __syncwarp();
    if (act && lane > col)
      for (int j = col + 1; j <= lane; j++) A[lane][j] -= l * A[j][col];
    __syncwarp();
  }
}
__device__ inline double warp_chol2_solve(const double (*A)[FLD], const double* dinv, double x, int lane) {
  const int base = lane < 6 ? 0 : 6, nloc = lane < 6 ? 6 : (lane < FR_NV ? 9 : 0);
  const double di = lane < FR_NV ? dinv[lane] : 0.0;
  for (int k = 0; k < 9; k++) {
    const bool act = k < nloc;
    const int col = act ? base + k : 0;
    const double yk = __shfl_sync(B2_FULLMASK, x * di, col);
    if (act) { if (lane == col) x = yk; else if (lane > col) x -= A[lane][col] * yk; }
  }
  for (int k = 8; k >= 0; k--) {
    const bool act = k < nloc;
    const int col = act ? base + k : 0;
    const double xk = __shfl_sync(B2_FULLMASK, x * di, col);
    if (act) { if (lane == col) x = xk; else if (lane < col && lane >= base) x -= A[col][lane] * xk; }
  }
  return x;
}

// arm block: factorise M (+ diag(add)) into rows 0..8 of W->H (free outside the solver); object block is diagonal
__device__ inline void fr3_factor_arm(Fr3Work* W, const double* add /* [FNA] or null */, int lane) {
  if (lane < FNA)
    for (int j = 0; j <= lane; j++) W->H[lane][j] = W->M[lane][j] + ((add && j == lane) ? add[lane] : 0.0);
  __syncwarp();
  warp_chol(W->H, W->Ld, FNA, lane);
}

// ------------------------------------------------------------------ collision + distance sensors, lane per box pair
__device__ inline void fr3_pair_geoms(const Fr3Model* __restrict__ m, const Fr3Work* W, int p, const double** p1, const double** m1,
                                      const double** s1, double* p2, const double** m2, const double** s2, int* cls, int* body) {
  if (p == 0) {  // table - object
    *p1 = m->table_pos; *m1 = m->table_mat; *s1 = m->table_size;
    p2[0] = W->xpos[0][0]; p2[1] = W->xpos[0][1]; p2[2] = W->xpos[0][2];
    *m2 = W->xmat[0]; *s2 = m->obj_size; *cls = 0; *body = 0;
    return;
  }
  const int k = (p - 1) % FNPAD, b = (int)m->pad_body[k];
  double t[3];
  lmat_vec(t, W->xmat[b], m->pad_pos[k]);
  p2[0] = W->xpos[b][0] + t[0]; p2[1] = W->xpos[b][1] + t[1]; p2[2] = W->xpos[b][2] + t[2];
  *m2 = W->xmat[b]; *s2 = m->pad_size[k]; *body = b;
  if (p <= FNPAD) { *p1 = m->table_pos; *m1 = m->table_mat; *s1 = m->table_size; *cls = 1; }
  else { *p1 = W->xpos[0]; *m1 = W->xmat[0]; *s1 = m->obj_size; *cls = 2; }
}

// dist_mode: 0 = every distance sensor exactly (contract A); 1 = what the cost needs: the SIGN of the finger-table distances (taken
// from the contact routine's separating-axis stage, no extra work); 2 = as 1 plus the object-table VALUE (PLACE phase)
// Narrow-phase scratch of one box pair; the FNPAIR of them alias the contact Jacobians cJ, which are only built after the collision pass,
// so the contact routines touch no local memory (the kernel used to carry a 1.1 KB stack frame per thread).
struct Fr3ColScratch { double p2[3]; LRaw raw[8]; LBoxScratch box; };
static_assert(sizeof(Fr3ColScratch) * FNPAIR <= sizeof(double) * FMAXCON * 3 * FR_NV, "collision scratch must fit into cJ");

__device__ inline void fr3_collision(const Fr3Model* __restrict__ m, Fr3Work* W, int lane, int dist_mode) {
  Fr3ColScratch* S = reinterpret_cast<Fr3ColScratch*>(&W->cJ[0][0][0]) + (lane < FNPAIR ? lane : 0);
  const LRaw* raw = S->raw;
  int n = 0, cls = 0, body = 0;
  if (lane < FNPAIR) {
    const double *p1, *m1, *s1, *m2, *s2;
    double* p2 = S->p2;
    int overlap = 0;
    fr3_pair_geoms(m, W, lane, &p1, &m1, &s1, p2, &m2, &s2, &cls, &body);
    n = l_box_box(p1, m1, s1, p2, m2, s2, 0.0, S->raw, 8, &S->box, &overlap);
    double dist = m->cutoff;
    if (dist_mode == 0 || (dist_mode == 2 && lane == 0)) dist = l_box_box_distance(p1, m1, s1, p2, m2, s2, m->cutoff, true);
    else if (lane >= 1 && lane <= FNPAD) dist = overlap ? -1.0 : 1.0;
    W->pdist[lane] = dist;
  }
  // ordered slot allocation (pair order == the oracle's): exclusive prefix of n over the lanes
  int pre = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(B2_FULLMASK, pre, o); if (lane >= o) pre += v; }
  const int total = __shfl_sync(B2_FULLMASK, pre, 31);
  pre -= n;
  for (int i = 0; i < n; i++) {
    const int slot = pre + i;
    if (slot < FMAXCON) {
      W->cdist[slot] = raw[i].dist;
#pragma unroll
      for (int k = 0; k < 3; k++) { W->cgeo[slot][k] = raw[i].pos[k]; W->cgeo[slot][3 + k] = raw[i].normal[k]; }
      l_make_frame(W->cgeo[slot] + 3);
      W->ccls[slot] = cls; W->cbody[slot] = body;
    }
  }
  if (lane == 0) { W->ncon = total < FMAXCON ? total : FMAXCON; if (total > FMAXCON) atomicAdd(contact_overflow_counter(m), 1ull); }
  __syncwarp();
  // sensors: 5 body-pair distances (min over the pads, clipped to +-cutoff), ee z axis, object position, grasp site
  if (lane < 5) {
    const int lo = lane == 0 ? 11 : lane == 1 ? 16 : lane == 2 ? 1 : lane == 3 ? 6 : 0, cnt = lane == 4 ? 1 : 5;
    double best = m->cutoff;
    for (int k = 0; k < cnt; k++) best = fmin(best, W->pdist[lo + k]);
    W->sens[lane] = fmin(fmax(best, -m->cutoff), m->cutoff);
  } else if (lane < 8) W->sens[lane] = W->xmat[8][3 * (lane - 5) + 2];
  else if (lane < 11) W->sens[lane] = W->xpos[0][lane - 8];
  else if (lane < 14) {
    const int k = lane - 11;
    W->sens[lane] = W->xpos[8][k] + W->xmat[8][3 * k] * m->site_pos[0] + W->xmat[8][3 * k + 1] * m->site_pos[1] + W->xmat[8][3 * k + 2] * m->site_pos[2];
  }
  __syncwarp();
}

// ------------------------------------------------------------------ constraint rows (mj_makeConstraint + mj_makeImpedance)
// one out-of-line copy of the impedance sigmoid / (K, B) conversion (pow() and the fp64 divisions are hundreds of instructions each:
// inlining them at every call site bloats a kernel whose instruction footprint already exceeds the instruction caches)
__device__ __noinline__ double fr3_impedance(const double* solimp, double pos, double margin) { return impedance(solimp, pos, margin); }
__device__ __noinline__ void fr3_KB(const double* solref, const double* solimp, double dt, double* K, double* B) {
  double ref0 = solref[0];
  const double ref1 = solref[1], dmax = fmin(fmax(solimp[1], B2_MINIMP), B2_MAXIMP);
  if (ref0 > 0) {
    if (ref0 < 2 * dt) ref0 = 2 * dt;
    *K = 1 / fmax(B2_MINVAL, dmax * dmax * ref0 * ref0 * ref1 * ref1);
    *B = 2 / fmax(B2_MINVAL, dmax * ref0);
  } else { *K = -ref0 / fmax(B2_MINVAL, dmax * dmax); *B = -ref1 / fmax(B2_MINVAL, dmax); }
}
// J x for scalar row r: row 0 is the equality q13 - q14, rows 1..9 / 10..18 act on arm dof r - 1 / r - 10
__device__ __forceinline__ double fr3_srow_dot(const Fr3Work* W, int r, const double* x) {
  return r == 0 ? x[13] - x[14] : W->ssign[r] * x[r < 10 ? 5 + r : r - 4];
}
__device__ __forceinline__ double fr3_cJ_dot(const double* J, int cls, const double* x) {
  double s = 0;
  if (cls != 1) {
#pragma unroll
    for (int i = 0; i < 6; i++) s += J[i] * x[i];
  }
  if (cls != 0) {
#pragma unroll
    for (int i = 6; i < FR_NV; i++) s += J[i] * x[i];
  }
  return s;
}

__device__ inline void fr3_make_constraint(const Fr3Model* __restrict__ m, Fr3Work* W, int lane) {
  // scalar rows in fixed slots: lane 31 -> equality (row 0); lane j < 9 -> friction loss (row 1 + j) and joint limit (row 10 + j;
  // lower or upper side, whichever is violated — never both; FST_INACTIVE when neither)
  if (lane == 31) {
    double K, B;
    fr3_KB(m->eq_solref, m->eq_solimp, m->dt, &K, &B);
    const double pos = W->qpos[14] - W->qpos[15], vel = W->qvel[13] - W->qvel[14];
    const double imp = fr3_impedance(m->eq_solimp, pos, 0.0);
    const double R = fmax(B2_MINVAL, (1 - imp) * (m->dof_invw[13] + m->dof_invw[14]) / imp);
    W->sR[0] = R; W->sD[0] = 1 / R; W->sfl[0] = 0; W->ssign[0] = 1; W->sstate[0] = FST_QUADRATIC;
    W->saref[0] = -B * vel - K * imp * pos;
  }
  if (lane < FNA) {
    const int rf = 1 + lane, rl = 10 + lane, dof = 6 + lane;
    double K, B;
    fr3_KB(m->fr_solref, m->fr_solimp, m->dt, &K, &B);
    double imp = fr3_impedance(m->fr_solimp, 0.0, 0.0);
    double R = fmax(B2_MINVAL, (1 - imp) * m->dof_invw[dof] / imp);
    W->ssign[rf] = 1; W->sfl[rf] = m->dof_frictionloss[dof]; W->sstate[rf] = FST_QUADRATIC;
    W->sR[rf] = R; W->sD[rf] = 1 / R; W->saref[rf] = -B * W->qvel[dof];
    const double q = W->qpos[7 + lane];
    const double dlo = q - m->lim_lo[lane], dhi = m->lim_hi[lane] - q;
    double pos = 0, sign = 0;
    if (dlo < m->lim_margin) { pos = dlo; sign = 1; }
    else if (dhi < m->lim_margin) { pos = dhi; sign = -1; }
    W->ssign[rl] = sign; W->sfl[rl] = 0; W->sforce[rl] = 0; W->sjar[rl] = 0;
    if (sign != 0) {
      fr3_KB(m->lim_solref, m->lim_solimp, m->dt, &K, &B);
      imp = fr3_impedance(m->lim_solimp, pos, m->lim_margin);
      R = fmax(B2_MINVAL, (1 - imp) * m->dof_invw[dof] / imp);
      W->sR[rl] = R; W->sD[rl] = 1 / R; W->sstate[rl] = FST_SATISFIED;
      W->saref[rl] = -B * sign * W->qvel[dof] - K * imp * (pos - m->lim_margin);
    } else { W->sR[rl] = 1; W->sD[rl] = 0; W->saref[rl] = 0; W->sstate[rl] = FST_INACTIVE; }
  }
  const int ncon = W->ncon;
  // contact frame-row Jacobians: lane per (contact, axis);  J = frame_a . (Jp(body2) - Jp(body1)) at the contact point
  for (int e = lane; e < 3 * ncon; e += 32) {
    const int c = e / 3, a = e - 3 * c, cls = W->ccls[c], b = W->cbody[c];
    const double* p = W->cgeo[c];
    const double* fr = W->cgeo[c] + 3 + 3 * a;
    double* J = W->cJ[c][a];
    // object: geom2 of class 0 (+), geom1 of class 2 (-), absent from class 1
    const double so = cls == 0 ? 1.0 : cls == 2 ? -1.0 : 0.0;
    const double ro[3] = {p[0] - W->xpos[0][0], p[1] - W->xpos[0][1], p[2] - W->xpos[0][2]};
#pragma unroll
    for (int k = 0; k < 3; k++) J[k] = so * fr[k];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double ax[3] = {W->xmat[0][k], W->xmat[0][3 + k], W->xmat[0][6 + k]};
      double cr[3];
      lcross3(cr, ax, ro);
      J[3 + k] = so * ldot3(fr, cr);
    }
    // pad on finger body b (9 or 10): geom2 of classes 1 and 2 (+): the 7 hinges and that finger's slide
#pragma unroll
    for (int j = 0; j < FNA; j++) {
      double v = 0;
      if (cls != 0) {
        if (j < 7) {
          const double rr[3] = {p[0] - W->anchor[j][0], p[1] - W->anchor[j][1], p[2] - W->anchor[j][2]};
          double cr[3];
          lcross3(cr, W->axis[j], rr);
          v = ldot3(fr, cr);
        } else if (j + 2 == b) v = ldot3(fr, W->axis[j]);
      }
      J[6 + j] = v;
    }
  }
  __syncwarp();
  // per contact: impedance, common pyramid-edge R, reference accelerations in the (n, t1, t2) basis
  for (int c = lane; c < ncon; c += 32) {
    const int cls = W->ccls[c], b = W->cbody[c];
    const double mu = m->con_mu[cls];
    double K, B;
    fr3_KB(m->con_solref[cls], m->con_solimp[cls], m->dt, &K, &B);
    const double imp = fr3_impedance(m->con_solimp[cls], W->cdist[c], 0.0);
    const double tran = (cls == 1 ? 0.0 : m->body_invw[0]) + (cls == 0 ? 0.0 : m->body_invw[b]);
    const double R0 = fmax(B2_MINVAL, (1 - imp) * (tran + mu * mu * tran) / imp);
    const double R1 = R0 / fmax(B2_MINVAL, m->impratio);
    const double mureg = mu * sqrt(R1 / R0);
    const double Rpy = 2 * mureg * mureg * R0;
    W->cD[c] = 1 / Rpy; W->cmu[c] = mu;
    const double vn = fr3_cJ_dot(W->cJ[c][0], cls, W->qvel), v1 = fr3_cJ_dot(W->cJ[c][1], cls, W->qvel), v2 = fr3_cJ_dot(W->cJ[c][2], cls, W->qvel);
    // (the contact point / frame in cgeo[c] are dead once the Jacobians exist: the slot now holds aref, jar, force)
    W->cgeo[c][CG_AREF] = -B * vn - K * imp * W->cdist[c];
    W->cgeo[c][CG_AREF + 1] = -B * v1;
    W->cgeo[c][CG_AREF + 2] = -B * v2;
  }
  __syncwarp();
}

// ------------------------------------------------------------------ Newton solver (mj_solNewton, primal, pyramidal cones)
// scalar row at x: returns cost, writes force and state (row 0: equality, rows 1..9: friction loss, rest: limits)
__device__ __forceinline__ double fr3_srow_eval(const Fr3Work* W, int r, double x, double* force, int* state) {
  const double D = W->sD[r];
  if (r == 0) { *force = -D * x; *state = FST_QUADRATIC; return 0.5 * D * x * x; }
  if (r < 10) {
    const double f = W->sfl[r], R = W->sR[r];
    if (x <= -R * f) { *force = f; *state = FST_LINEARNEG; return -0.5 * R * f * f - f * x; }
    if (x >= R * f) { *force = -f; *state = FST_LINEARPOS; return -0.5 * R * f * f + f * x; }
    *force = -D * x; *state = FST_QUADRATIC; return 0.5 * D * x * x;
  }
  if (*state == FST_INACTIVE) { *force = 0; return 0; }
  if (x < 0) { *force = -D * x; *state = FST_QUADRATIC; return 0.5 * D * x * x; }
  *force = 0; *state = FST_SATISFIED; return 0;
}
// one pyramidal contact at residuals r = (rn, rt1, rt2): the 4 edge rows are rn +- mu rt_a.  Returns the cost; writes the force in the
// (n, t1, t2) basis.
__device__ __forceinline__ double fr3_contact_eval(double D, double mu, const double* r, double* f) {
  double cost = 0, fn = 0, ft[2] = {0, 0};
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const double sgn = s == 0 ? 1.0 : -1.0;
      const double x = r[0] + sgn * mu * r[1 + a];
      if (x < 0) { const double fr = -D * x; cost += 0.5 * D * x * x; fn += fr; ft[a] += sgn * mu * fr; }
    }
  f[0] = fn; f[1] = ft[0]; f[2] = ft[1];
  return cost;
}
// 3x3 weight (nn, nt1, nt2, t1t1, t2t2) of the contact's Hessian term J^T W J in the (n, t1, t2) basis, from the active edge rows
__device__ __forceinline__ bool fr3_contact_weight(double D, double mu, const double* r, double* w) {
  double cnt[2] = {0, 0}, sg[2] = {0, 0};
#pragma unroll
  for (int a = 0; a < 2; a++) {
    if (r[0] + mu * r[1 + a] < 0) { cnt[a] += 1; sg[a] += 1; }
    if (r[0] - mu * r[1 + a] < 0) { cnt[a] += 1; sg[a] -= 1; }
  }
  w[0] = D * (cnt[0] + cnt[1]); w[1] = D * mu * sg[0]; w[2] = D * mu * sg[1]; w[3] = D * mu * mu * cnt[0]; w[4] = D * mu * mu * cnt[1];
  return cnt[0] + cnt[1] > 0;
}

// jar = J qacc - aref for every row; Ma = M qacc
__device__ __noinline__ void fr3_set_point(const Fr3Model* __restrict__ m, Fr3Work* W, const double* qacc, int lane) {
  if (lane < FR_NV) W->Ma[lane] = fr3_mulM_row(m, W, qacc, lane);
  const int ncon = W->ncon;
  if (lane < FNS) W->sjar[lane] = fr3_srow_dot(W, lane, qacc) - W->saref[lane];
  for (int e = lane; e < 3 * ncon; e += 32) {
    const int c = e / 3, a = e - 3 * c;
    W->cgeo[c][CG_JAR + a] = fr3_cJ_dot(W->cJ[c][a], W->ccls[c], qacc) - W->cgeo[c][CG_AREF + a];
  }
  __syncwarp();
}

// cost / forces / states at the current jar, gradient, Gauss term
__device__ __noinline__ void fr3_constraint_update(const Fr3Model* __restrict__ m, Fr3Work* W, const double* qacc, int lane) {
  const int ncon = W->ncon;
  double cost = 0;
  if (lane < FNS) cost += fr3_srow_eval(W, lane, W->sjar[lane], &W->sforce[lane], &W->sstate[lane]);
  for (int c = lane; c < ncon; c += 32) cost += fr3_contact_eval(W->cD[c], W->cmu[c], W->cgeo[c] + CG_JAR, W->cgeo[c] + CG_FORCE);
  cost = fwsum(cost);
  __syncwarp();
  double g = 0;
  if (lane < FR_NV) {
    const int i = lane;
    double f = 0;
    if (i >= 6) f = W->sforce[i - 5] + W->ssign[i + 4] * W->sforce[i + 4];  // friction-loss row 1 + j, limit row 10 + j (j = i - 6)
    if (i == 13) f += W->sforce[0]; else if (i == 14) f -= W->sforce[0];
    for (int c = 0; c < ncon; c++) {
      const int cls = W->ccls[c];
      if ((i < 6 && cls == 1) || (i >= 6 && cls == 0)) continue;
      const double* fc = W->cgeo[c] + CG_FORCE;
      f += W->cJ[c][0][i] * fc[0] + W->cJ[c][1][i] * fc[1] + W->cJ[c][2][i] * fc[2];
    }
    W->qfrc_constraint[i] = f;
    W->grad[i] = W->Ma[i] - W->qfrc_smooth[i] - f;
    g = (W->Ma[i] - W->qfrc_smooth[i]) * (qacc[i] - W->qacc_smooth[i]);
  }
  g = fwsum(g);
  if (lane == 0) { W->gauss = 0.5 * g; W->cost = 0.5 * g + cost; }
  __syncwarp();
}

// Newton direction: search = -H^-1 grad,  H = M + sum_rows D j j^T + sum_contacts Jc^T Wc Jc  (dense 15x15, lane per entry, warp Cholesky)
__device__ inline void fr3_newton_direction(const Fr3Model* __restrict__ m, Fr3Work* W, int lane) {
  const int ncon = W->ncon;
  // every lane owns up to four entries (i >= j) of the lower triangle: mass matrix + scalar rows first ...
  constexpr int NE = FR_NV * (FR_NV + 1) / 2;
  int ei[4], ej[4];
  double h[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int e = lane + 32 * q;
    const int i = e < NE ? FR3_TRI[e] >> 4 : 0, j = e < NE ? FR3_TRI[e] & 15 : 0;
    double v = 0;
    if (i < 6) { if (i == j) v = i < 3 ? m->body_mass[0] : m->obj_inertia[i - 3]; }
    else if (j >= 6) v = W->M[i - 6][j - 6];
    if (i == j && i >= 6) {
      if (W->sstate[i - 5] == FST_QUADRATIC) v += W->sD[i - 5];
      if (W->sstate[i + 4] == FST_QUADRATIC) v += W->sD[i + 4];
      if (i >= 13) v += W->sD[0];
    } else if (i == 14 && j == 13) v -= W->sD[0];
    ei[q] = i; ej[q] = j; h[q] = v;
  }
  // ... then the contacts: the 3x3 weight of a contact is evaluated once per lane and applied to the lane's entries
  for (int c = 0; c < ncon; c++) {
    double w[5];
    if (!fr3_contact_weight(W->cD[c], W->cmu[c], W->cgeo[c] + CG_JAR, w)) continue;  // no active edge
    const int cls = W->ccls[c];
    const double (*J)[FR_NV] = W->cJ[c];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int i = ei[q], j = ej[q];
      if ((j < 6 && cls == 1) || (i >= 6 && cls == 0)) continue;  // i >= j: both dofs must be touched by the contact
      const double ni = J[0][i], nj = J[0][j], ai = J[1][i], aj = J[1][j], bi = J[2][i], bj = J[2][j];
      h[q] += w[0] * ni * nj + w[1] * (ni * aj + ai * nj) + w[2] * (ni * bj + bi * nj) + w[3] * ai * aj + w[4] * bi * bj;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; q++) if (lane + 32 * q < NE) W->H[ei[q]][ej[q]] = h[q];
  __syncwarp();
  // object-pad contacts are the only coupling between the object block and the arm block
  bool coupled = false;
  for (int c = lane; c < ncon; c += 32) coupled = coupled || W->ccls[c] == 2;
  coupled = __any_sync(B2_FULLMASK, coupled);
  double x;
  if (coupled) {
    warp_chol(W->H, W->Hd, FR_NV, lane);
    x = warp_chol_solve(W->H, W->Hd, FR_NV, lane < FR_NV ? -W->grad[lane] : 0.0, lane);
  } else {
    warp_chol2(W->H, W->Hd, lane);
    x = warp_chol2_solve(W->H, W->Hd, lane < FR_NV ? -W->grad[lane] : 0.0, lane);
  }
  if (lane < FR_NV) W->search[lane] = x;
  __syncwarp();
}

// Line search state in REGISTERS: lane r owns scalar row r (<= 19) and lane c owns contacts c and c + 32 (ncon <= 48).
struct Fr3LS {
  double rjar, rjv, rD, rR, rfl;
  double cjar[2][3], cjv[2][3], cD[2], cmu[2];
  int kind;  // -1 none / inactive limit, 0 equality, 1 friction, 2 limit
  int nc;
};
__device__ __forceinline__ void fr3_ls_eval(const Fr3LS& L, double alpha, double g1, double g2, double* d1, double* d2) {
  double p1 = 0, p2 = 0;
  if (L.kind >= 0) {
    const double x = L.rjar + alpha * L.rjv;
    bool quad = L.kind == 0;
    if (L.kind == 1) {
      const double lim = L.rR * L.rfl;
      if (x <= -lim) p1 -= L.rfl * L.rjv;
      else if (x >= lim) p1 += L.rfl * L.rjv;
      else quad = true;
    } else if (L.kind == 2) quad = x < 0;
    if (quad) { p1 += L.rD * x * L.rjv; p2 += L.rD * L.rjv * L.rjv; }
  }
#pragma unroll
  for (int s = 0; s < 2; s++)
    if (s < L.nc) {
      const double mu = L.cmu[s], D = L.cD[s];
      const double rn = L.cjar[s][0] + alpha * L.cjv[s][0], r1 = L.cjar[s][1] + alpha * L.cjv[s][1], r2 = L.cjar[s][2] + alpha * L.cjv[s][2];
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int sg = 0; sg < 2; sg++) {
          const double sgn = sg == 0 ? mu : -mu;
          const double x = rn + sgn * (a == 0 ? r1 : r2), v = L.cjv[s][0] + sgn * L.cjv[s][1 + a];
          if (x < 0) { p1 += D * x * v; p2 += D * v * v; }
        }
    }
  fwsum2(p1, p2);
  *d1 = g1 + alpha * g2 + p1;
  *d2 = g2 + p2;
}

// exact line search: safeguarded 1-D Newton on the convex piecewise-quadratic cost along the search direction.  The residuals (jar)
// and their directional derivatives (jv) of the rows / contacts this lane owns are in L (registers); returns the step alpha.
__device__ inline double fr3_line_search(const Fr3Model* __restrict__ m, Fr3Work* W, int lane, Fr3LS& L) {
  double g1 = 0, g2 = 0, sn = 0, gs = 0;
  if (lane < FR_NV) {
    g1 = W->search[lane] * (W->Ma[lane] - W->qfrc_smooth[lane]); g2 = W->search[lane] * W->Mv[lane];
    sn = W->search[lane] * W->search[lane]; gs = W->grad[lane] * W->search[lane];
  }
  fwsum2(g1, g2);
  fwsum2(sn, gs);
  const double snorm = F_SQRT(sn);
  // derivatives at alpha = 0 without touching the rows: d1(0) = grad . search and, for the Newton direction, d2(0) = -d1(0)
  double d1 = gs, d2 = -d1, lo = 0, hi = -1, alpha;
  if (snorm < B2_MINVAL) return 0;
  const double gtol = m->tolerance * m->ls_tolerance * snorm * m->meaninertia * FR_NV;
  if (d1 >= 0 || d2 <= 0) return 0;
  alpha = 1.0;
  double prev_step = 1e300;
  const int iters = (int)m->ls_iterations;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    fr3_ls_eval(L, alpha, g1, g2, &d1, &d2);
    if (fabs(d1) < gtol) return alpha;
    if (d1 < 0) lo = alpha; else hi = alpha;
    double next = d2 > 0 ? alpha - d1 * F_RCP(fmax(d2, 1e-200)) : -1;
    if (hi < 0) { if (!(next > lo)) next = 2 * alpha + B2_MINVAL; }
    else if (!(next > lo && next < hi && fabs(next - alpha) < 0.5 * prev_step)) next = 0.5 * (lo + hi);
    if (next == alpha) return alpha;
    prev_step = fabs(next - alpha);
    alpha = next;
  }
  return lo > 0 ? lo : alpha;
}

// mj_fwdConstraint.  Called by ALL warps of the block; with sync_mode >= 3 the Newton iterations of the block's warps run in
// lock-step (block barrier per iteration, finished warps idle) so the iteration body is fetched once for all of them.
__device__ inline void fr3_fwd_constraint(const Fr3Model* __restrict__ m, Fr3Work* W, int lane, bool active, int sync_mode, int prof) {
  bool done = !active;
  double scale = 0;
  if (!done) {
    // warm start: keep qacc_warmstart unless qacc_smooth has lower cost.  qacc_smooth is evaluated FIRST so that in the common case
    // (the warm start wins) the row residuals / forces / gradient left behind are already those of the chosen point
    fr3_set_point(m, W, W->qacc_smooth, lane);
    fr3_constraint_update(m, W, W->qacc_smooth, lane);
    const double cs = W->cost;
    __syncwarp();
    fr3_set_point(m, W, W->warm, lane);
    fr3_constraint_update(m, W, W->warm, lane);
    const double cw = W->cost;
    __syncwarp();
    if (lane < FR_NV) W->qacc[lane] = cw > cs ? W->qacc_smooth[lane] : W->warm[lane];
    __syncwarp();
    if (cw > cs) {
      fr3_set_point(m, W, W->qacc, lane);
      fr3_constraint_update(m, W, W->qacc, lane);
    }
    scale = 1.0 / (m->meaninertia * FR_NV);
  }
  const int iters = (int)m->iterations;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    if (sync_mode >= 3) { if (!__syncthreads_or(done ? 0 : 1)) break; }
    else if (done) break;
    if (done) continue;
    double gn = lane < FR_NV ? W->grad[lane] * W->grad[lane] : 0.0;
    gn = fwsum(gn);
    if (scale * F_SQRT(gn) < m->tolerance) { done = true; continue; }
    long long t1 = FPROF_T();
    fr3_newton_direction(m, W, lane);
    FPROF_ADD(8, t1); t1 = FPROF_T();
    if (prof && lane == 0) atomicAdd(&g_fr3_prof[10], 1ull);
    const int ncon = W->ncon;
    if (lane < FR_NV) W->Mv[lane] = fr3_mulM_row(m, W, W->search, lane);
    __syncwarp();
    // row data of the line search, held in registers by the lane that owns the row / contact
    Fr3LS L;
    L.kind = lane >= FNS ? -1 : lane == 0 ? 0 : lane < 10 ? 1 : (W->sstate[lane] == FST_INACTIVE ? -1 : 2);
    L.rjar = L.rjv = L.rD = L.rR = L.rfl = 0;
    if (L.kind >= 0) { L.rjar = W->sjar[lane]; L.rjv = fr3_srow_dot(W, lane, W->search); L.rD = W->sD[lane]; L.rR = W->sR[lane]; L.rfl = W->sfl[lane]; }
    L.nc = 0;
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const int c = lane + 32 * s;
#pragma unroll
      for (int a = 0; a < 3; a++) { L.cjar[s][a] = 0; L.cjv[s][a] = 0; }
      L.cD[s] = 0; L.cmu[s] = 0;
      if (c < ncon) {
        L.nc = s + 1;
#pragma unroll
        for (int a = 0; a < 3; a++) { L.cjar[s][a] = W->cgeo[c][CG_JAR + a]; L.cjv[s][a] = fr3_cJ_dot(W->cJ[c][a], W->ccls[c], W->search); }
        L.cD[s] = W->cD[c]; L.cmu[s] = W->cmu[c];
      }
    }
    const double alpha = fr3_line_search(m, W, lane, L);
    FPROF_ADD(9, t1); t1 = FPROF_T();
    if (alpha == 0) { done = true; continue; }
    const double oldcost = W->cost;
    __syncwarp();
    if (lane < FR_NV) { W->qacc[lane] += alpha * W->search[lane]; W->Ma[lane] += alpha * W->Mv[lane]; }
    if (L.kind >= 0) W->sjar[lane] = L.rjar + alpha * L.rjv;
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const int c = lane + 32 * s;
      if (c < ncon) {
#pragma unroll
        for (int a = 0; a < 3; a++) W->cgeo[c][CG_JAR + a] = L.cjar[s][a] + alpha * L.cjv[s][a];
      }
    }
    __syncwarp();
    fr3_constraint_update(m, W, W->qacc, lane);
    FPROF_ADD(7, t1);
    const double newcost = W->cost;
    __syncwarp();
    if (scale * (oldcost - newcost) < m->tolerance) done = true;
  }
}

// ------------------------------------------------------------------ one mj_step
__device__ inline void fr3_step(const Fr3Model* __restrict__ m, Fr3Work* W, int lane, int dist_mode, bool active, int sync_mode, int prof) {
  long long t0 = FPROF_T();
  if (active) {
    fr3_kinematics(m, W, lane);
    FPROF_ADD(0, t0); t0 = FPROF_T();
    fr3_mass_and_bias(m, W, lane);
    FPROF_ADD(1, t0);
  }
  if (sync_mode >= 2) __syncthreads();
  t0 = FPROF_T();
  if (active) {
    fr3_collision(m, W, lane, dist_mode);
    FPROF_ADD(2, t0); t0 = FPROF_T();
    fr3_make_constraint(m, W, lane);
    FPROF_ADD(3, t0); t0 = FPROF_T();
    // passive + actuation -> qfrc_smooth; qacc_smooth = M^-1 qfrc_smooth
    if (lane < FR_NV) {
      const int i = lane;
      double f = -m->dof_damping[i] * W->qvel[i] - W->qfrc_bias[i];
      if (i >= 6 && i <= 13) {
        const int a = i - 6;
        const double u = fmin(fmax(W->ctrl[a], m->ctrl_lo[a]), m->ctrl_hi[a]);
        double fa = m->kp[a] * u - m->kp[a] * W->qpos[7 + a] - m->kv[a] * W->qvel[i];
        if (m->frc_limited[a] != 0) fa = fmin(fmax(fa, m->frc_lo[a]), m->frc_hi[a]);
        f += fa;
      }
      W->qfrc_smooth[i] = f;
    }
    fr3_factor_arm(W, nullptr, lane);
    double x = 0;
    if (lane < 3) x = W->qfrc_smooth[lane] / m->body_mass[0];
    else if (lane < 6) x = W->qfrc_smooth[lane] / m->obj_inertia[lane - 3];
    const double xa = warp_chol_solve(W->H, W->Ld, FNA, lane < FNA ? W->qfrc_smooth[6 + lane] : 0.0, lane);
    if (lane < 6) W->qacc_smooth[lane] = x;
    if (lane < FNA) W->qacc_smooth[6 + lane] = xa;
    __syncwarp();
    FPROF_ADD(4, t0);
  }
  if (sync_mode >= 2) __syncthreads();
  t0 = FPROF_T();
  fr3_fwd_constraint(m, W, lane, active, sync_mode, prof);
  __syncwarp();
  if (!active) return;
  FPROF_ADD(5, t0); t0 = FPROF_T();
  // implicitfast: (M + h (damping + kv)) qacc = qfrc_smooth + qfrc_constraint on the arm block; the object block has no
  // velocity-dependent forces, so its qacc is the solver's; then the semi-implicit advance
  const double h = m->dt;
  if (lane < FNA) W->Mv[lane] = h * (m->dof_damping[6 + lane] + (lane < FR_NU ? m->kv[lane] : 0.0));
  __syncwarp();
  fr3_factor_arm(W, W->Mv, lane);
  const double qa = warp_chol_solve(W->H, W->Ld, FNA, lane < FNA ? W->qfrc_smooth[6 + lane] + W->qfrc_constraint[6 + lane] : 0.0, lane);
  double qo = 0;
  if (lane < 3) qo = (W->qfrc_smooth[lane] + W->qfrc_constraint[lane]) / m->body_mass[0];
  else if (lane < 6) qo = (W->qfrc_smooth[lane] + W->qfrc_constraint[lane]) / m->obj_inertia[lane - 3];
  __syncwarp();
  if (lane < FR_NV) W->warm[lane] = W->qacc[lane];
  if (lane < 6) W->qvel[lane] += h * qo;
  if (lane < FNA) W->qvel[6 + lane] += h * qa;
  __syncwarp();
  if (lane < 3) W->qpos[lane] += h * W->qvel[lane];
  else if (lane == 3) {
    double w[3] = {W->qvel[3], W->qvel[4], W->qvel[5]};
    const double ang = h * lnormalize3(w);
    double sn, cs, dq[4], nq[4];
    sincos(0.5 * ang, &sn, &cs);
    dq[0] = cs; dq[1] = sn * w[0]; dq[2] = sn * w[1]; dq[3] = sn * w[2];
    lquat_mul(nq, W->qpos + 3, dq);
    lquat_normalize(nq);
    for (int k = 0; k < 4; k++) W->qpos[3 + k] = nq[k];
  } else if (lane >= 6 && lane < FR_NV) W->qpos[lane + 1] += h * W->qvel[lane];
  __syncwarp();
  FPROF_ADD(6, t0);
}

// per-step cost (fr3_pick.py:225-311): phase term + global terms, from the POST-step state and this step's sensordata.
// params: [phase, w_lift_close, w_lift_height, w_move_goal, w_move_close, w_place_table, w_place_goal, w_upright, w_coll, w_qvel,
//          w_open, goal_x, goal_y, pick_height, q_home(9)];  decay = linspace(1, 0, H)[t]
__device__ inline double fr3_cost(const double* p, const double* qpos, const double* qvel, const double* sens, double decay) {
  const int phase = (int)p[0];
  const double gx = sens[11] - qpos[0], gy = sens[12] - qpos[1], gz = sens[13] - qpos[2];
  const double grasp = gx * gx + gy * gy + gz * gz;
  const double ex = qpos[0] - p[11], ey = qpos[1] - p[12];
  const double goal = sqrt(ex * ex + ey * ey);
  double c;
  if (phase == 0) { const double e = qpos[2] - p[13]; c = p[1] * grasp + p[2] * e * e; }
  else if (phase == 1) c = p[3] * goal + p[4] * grasp;
  else if (phase == 2) c = p[5] * sens[4] + p[6] * goal;
  else {
    double s = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) { const double e = qpos[7 + k] - p[14 + k]; s += e * e; }
    c = sqrt(s);
  }
  const double ux = sens[5], uy = sens[6], uz = sens[7] + 1.0;
  const double touching = (sens[2] <= 0.0 || sens[3] <= 0.0) ? 1.0 : 0.0;
  double v2 = 0;
#pragma unroll
  for (int k = 0; k < FR_NV; k++) v2 += qvel[k] * qvel[k];
  const double go = qpos[15] - 0.04;
  c += p[7] * sqrt(ux * ux + uy * uy + uz * uz) - p[8] * (1.0 - touching) + p[9] * decay * sqrt(v2) + p[10] * go * go;
  return c;
}

// ------------------------------------------------------------------ kernels
// COST: in = knots (N,K,8), basis (H,K) -> reward (N) [+ cost (N,H) f32];  !COST: in = controls (N,H,8) -> states, sensors
template <bool COST>
__global__ void __launch_bounds__(256) fr3_rollout_kernel(const Fr3Model* __restrict__ m, const double* __restrict__ x0, int x0_batched,
                                                          const double* __restrict__ in, int N, int H, int K, const double* __restrict__ basis,
                                                          const double* __restrict__ cost_params, double* __restrict__ states,
                                                          double* __restrict__ sensors, float* __restrict__ cost_NH, double* __restrict__ reward_N,
                                                          int wstride, int sync_mode, const SampleSpec smp, int index_offset,
                                                          double* __restrict__ trace_out = nullptr /* COST: (N, H, 6) or null */) {
  const int prof = sync_mode >> 8;
  sync_mode &= 255;
  B2_DYNAMIC_SMEM(unsigned char, fsm_all);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int n = blockIdx.x * wpb + wib;
  const bool active = n < N;
  unsigned char* fsm = fsm_all + (size_t)wib * wstride;
  Fr3Work* W = reinterpret_cast<Fr3Work*>(fsm);
  if (active) {
    const double* xs = x0 + (x0_batched ? (size_t)n * FR_NX : 0);
    if (lane < FR_NQ) W->qpos[lane] = xs[lane];
    if (lane < FR_NV) { W->qvel[lane] = xs[FR_NQ + lane]; W->warm[lane] = 0; }
  }
  __syncwarp();
  if constexpr (COST) {
    // this rollout's knots (K*8 doubles) are staged behind the work area with one TMA bulk copy; the (H,K) basis is shared by
    // every rollout and stays in global memory (K cached loads per lane and step) so that seven work areas fit one SM
    uint64_t* bar = reinterpret_cast<uint64_t*>(fsm + ((sizeof(Fr3Work) + 15) & ~(size_t)15));
    double* sK = reinterpret_cast<double*>(bar + 2);
    const int dist_mode = (int)cost_params[0] == 2 ? 2 : 1;  // only the PLACE phase reads the object-table distance value
    if (active) {
      const unsigned bytesK = (unsigned)(K * FR_NU * sizeof(double));
      const double* gK = in + (size_t)n * K * FR_NU;
      if (!smp.enabled) {
        if ((reinterpret_cast<uintptr_t>(gK) & 15) == 0) {
          if (lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
          __syncwarp();
          if (lane == 0) { mbar_expect_tx(bar, bytesK); tma_bulk_g2s(sK, gK, bytesK, bar); }
          mbar_wait(bar, 0);
        } else {
          for (int i = lane; i < K * FR_NU; i += 32) sK[i] = gK[i];
          __syncwarp();
        }
      } else {
        const long long gn = (long long)n + index_offset;
        const int KNU = K * FR_NU;
        for (int p2 = lane; 2 * p2 < KNU; p2 += 32) {
          double z0, z1;
          normal_pair(smp, gn, p2, &z0, &z1);
          const double a = sample_element(smp, gn, 2 * p2, FR_NU, z0);
          sK[2 * p2] = a; smp.knots_out[(size_t)n * KNU + 2 * p2] = a;
          if (2 * p2 + 1 < KNU) { const double b = sample_element(smp, gn, 2 * p2 + 1, FR_NU, z1); sK[2 * p2 + 1] = b; smp.knots_out[(size_t)n * KNU + 2 * p2 + 1] = b; }
        }
        __syncwarp();
      }
    }
    double total = 0;
#pragma unroll 1
    for (int t = 0; t < H; t++) {
      if (sync_mode >= 1) __syncthreads();
      if (active && lane < FR_NU) {
        double u = 0;
        for (int k = 0; k < K; k++) u += __ldg(basis + t * K + k) * sK[k * FR_NU + lane];
        W->ctrl[lane] = u;
      }
      __syncwarp();
      fr3_step(m, W, lane, dist_mode, active, sync_mode, prof);
      if (trace_out && active && lane < FR_NTRACE) trace_out[((size_t)n * H + t) * FR_NTRACE + lane] = W->sens[8 + lane];  // T1: trace sensors of every rollout
      if (active && lane == 0) {
        double cp[FR_NCOST];
#pragma unroll
        for (int i = 0; i < FR_NCOST; i++) cp[i] = cost_params[i];
        const double decay = H > 1 ? 1.0 - (double)t / (double)(H - 1) : 1.0;
        const double ct = fr3_cost(cp, W->qpos, W->qvel, W->sens, decay);
        total += ct;
        if (cost_NH) cost_NH[(size_t)n * H + t] = (float)ct;
      }
    }
    if (active && lane == 0) reward_N[n] = -total;
  } else {
#pragma unroll 1
    for (int t = 0; t < H; t++) {
      if (sync_mode >= 1) __syncthreads();
      if (active && lane < FR_NU) W->ctrl[lane] = in[((size_t)n * H + t) * FR_NU + lane];
      __syncwarp();
      fr3_step(m, W, lane, 0, active, sync_mode, prof);
      if (!active) continue;
      double* so = states + ((size_t)n * H + t) * FR_NX;
      if (lane < FR_NQ) so[lane] = W->qpos[lane];
      if (lane < FR_NV) so[FR_NQ + lane] = W->qvel[lane];
      if (sensors && lane < FR_NS) sensors[((size_t)n * H + t) * FR_NS + lane] = W->sens[lane];
    }
  }
}

// Task.reward from given trajectories (contract A callers): one thread per rollout
__global__ void fr3_reward_kernel(const double* __restrict__ states, const double* __restrict__ sensors, int N, int H,
                                  const double* __restrict__ cost_params, double* __restrict__ reward_N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double cp[FR_NCOST];
#pragma unroll
  for (int i = 0; i < FR_NCOST; i++) cp[i] = cost_params[i];
  double total = 0;
  for (int t = 0; t < H; t++) {
    const double* s = states + ((size_t)n * H + t) * FR_NX;
    const double decay = H > 1 ? 1.0 - (double)t / (double)(H - 1) : 1.0;
    total += fr3_cost(cp, s, s + FR_NQ, sensors + ((size_t)n * H + t) * FR_NS, decay);
  }
  reward_N[n] = -total;
}

inline size_t fr3_wstride(int cost_mode, int K, int H) {
  (void)H;
  size_t w = ((sizeof(Fr3Work) + 15) & ~(size_t)15) + 16 + (cost_mode ? (size_t)K * FR_NU * sizeof(double) : 0);
  return (w + 15) & ~(size_t)15;
}

// ------------------------------------------------------------------ host side
#ifndef B2_HOST_SIM
inline int fr3_create(Fr3Model** out, const double* consts, size_t n, std::string* err) {
  if (n != sizeof(Fr3Model) / sizeof(double)) { *err = "wrong number of task constants"; return 1; }
  Fr3Model* d = nullptr;
  if (cudaMalloc(&d, sizeof(Fr3Model) + 16) != cudaSuccess) { *err = "cudaMalloc failed"; return 1; }  // + the handle's contact-overflow counter (geom.cuh)
  if (cudaMemset(d + 1, 0, 16) != cudaSuccess) { cudaFree(d); *err = "cudaMemset failed"; return 1; }
  if (cudaMemcpy(d, consts, sizeof(Fr3Model), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); *err = "cudaMemcpy failed"; return 1; }
  *out = d;
  return 0;
}
inline void fr3_prof_dump() {
  unsigned long long h[16];
  if (cudaMemcpyFromSymbol(h, g_fr3_prof, sizeof(h)) != cudaSuccess) return;
  const char* names[11] = {"kinematics", "mass+bias", "collision", "constraints", "smooth", "solver(total)", "integrate", "  update", "  direction", "  linesearch", "newton iters"};
  for (int i = 0; i < 11; i++) fprintf(stderr, "fr3_prof %-14s %llu\n", names[i], h[i]);
  memset(h, 0, sizeof(h));
  cudaMemcpyToSymbol(g_fr3_prof, h, sizeof(h));
}
inline void fr3_destroy(Fr3Model* m) { if (getenv("B200MPC_FR3_PROF")) fr3_prof_dump(); cudaFree(m); }

inline int fr3_launch(const Fr3Model* m, int cost_mode, const double* d_x0, int batched, const double* d_in, int N, int H, int K,
                      const double* d_basis, const double* d_params, double* d_states, double* d_sensors, float* d_cost, double* d_reward,
                      const PlanEpilogue& ep, const SampleSpec& smp, cudaStream_t st, std::string* err, double* d_trace = nullptr) {
  const char* sm_env = getenv("B200MPC_FR3_SYNC");
  const int sync_mode = (sm_env ? atoi(sm_env) : 3) | ((getenv("B200MPC_FR3_PROF") ? 1 : 0) << 8);
  const size_t wstride = fr3_wstride(cost_mode, K, H);
  int wpb = (N + 147) / 148;  // spread the rollouts over the 148 SMs first, then stack warps per SM
  if (wpb < 1) wpb = 1;
  if (wpb > 8) wpb = 8;
  while (wpb > 1 && wpb * wstride > 226 * 1024) wpb--;
  const size_t smem = wpb * wstride;
  if (smem > 227 * 1024) { *err = "horizon/knots too large for the shared-memory tile"; return 1; }
  const int grid = (N + wpb - 1) / wpb;
  cudaError_t e;
  if (cost_mode) {
    e = set_max_dynamic_smem_once((const void*)fr3_rollout_kernel<true>, 3, smem);
    if (e == cudaSuccess)
      fr3_rollout_kernel<true><<<grid, 32 * wpb, smem, st>>>(m, d_x0, batched, d_in, N, H, K, d_basis, d_params, nullptr, nullptr, d_cost, d_reward, (int)wstride, sync_mode, smp, ep.index_offset, d_trace);
  } else {
    e = set_max_dynamic_smem_once((const void*)fr3_rollout_kernel<false>, 2, smem);
    if (e == cudaSuccess)
      fr3_rollout_kernel<false><<<grid, 32 * wpb, smem, st>>>(m, d_x0, batched, d_in, N, H, 0, nullptr, nullptr, d_states, d_sensors, nullptr, nullptr, (int)wstride, sync_mode, smp, 0);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("fr3 launch: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

inline int fr3_reward_launch(const double* d_states, const double* d_sensors, int N, int H, const double* d_params, double* d_reward,
                             cudaStream_t st, std::string* err) {
  fr3_reward_kernel<<<(N + 127) / 128, 128, 0, st>>>(d_states, d_sensors, N, H, d_params, d_reward);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("fr3 reward launch: ") + cudaGetErrorString(e); return 1; }
  return 0;
}
#endif  // !B2_HOST_SIM

}  // namespace b2
