// leap.cuh — leap_cube: warp-per-rollout reduced articulated-body integrator + cost.
//
// One warp owns one rollout: 23 qpos / 22 dofs (free cube + 4 fingers x 4 hinges), state and all per-step work
// arrays live in shared memory (~32 KB per warp -> 7 resident rollouts per SM, one wave at N = 1024).
// Every stage of MuJoCo's mj_step that judo/models/xml/leap_cube.xml switches on is restated and spread over the
// 32 lanes: kinematics (lane per chain), mass matrix + RNE (lane per finger), collision of the cube against the hand
// geoms (lane per geom / per candidate pair), constraint rows (lane per row), the primal Newton solver with elliptic
// cones (lane per row / per contact / per Hessian entry, warp Cholesky), implicitfast integration.
// Collision covers every geom pair MuJoCo's static filters leave (cube-hand AND hand-hand); only the fingertip meshes are replaced
// by spheres; the constraint/solver pipeline is the full one.
#pragma once
#include "epilogue.cuh"
#include "geom.cuh"
#include "sampling.cuh"

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

namespace b2 {

// solver-loop reciprocals / roots without the library's special-case branches (common.cuh); B2_LEAP_LIBMATH restores the library calls
#ifdef B2_LEAP_LIBMATH
#define L_RSQRT(x) rsqrt(x)
#define L_SQRT(x) sqrt(x)
#define L_RCP(x) (1.0 / (x))
#else
#define L_RSQRT(x) fast_rsqrt_pos(x)
#define L_SQRT(x) fast_sqrt_nonneg(x)
#define L_RCP(x) fast_rcp_pos(x)
#endif

constexpr int LEAP_NQ = 23, LEAP_NV = 22, LEAP_NU = 16, LEAP_NS = 31, LEAP_NX = 45, LEAP_NCOST = 9;
constexpr int LEAP_NTRACE = 15;  // doubles per step kept by the fused kernel's trace capture: the 5 framepos trace sensors
constexpr int LB = 17;         // moving bodies: 0 = cube, 1 + 4f + d = link d of finger f
constexpr int LMAXG = 80;      // hand collision geoms
constexpr int LMAXHH = 1664;   // hand-hand candidate pairs (finger-finger, finger-palm, links of one finger: 1 621 for leap_cube)
constexpr int LMAXCON = 40;    // contacts kept per step (3 rows each); the C4 bench scenario peaks at 36 with the hand-hand pairs
constexpr int LMAXFL = 32;     // friction-loss rows (16) + active joint-limit rows (at most one side per joint)
constexpr int LMAXEFC = LMAXFL + 3 * LMAXCON;
constexpr unsigned FULL = 0xffffffffu;

// All-double POD; field order == judo_b200/tasks/leap_cube.py:leap_consts.
struct LeapModel {
  double dt, gravity[3], impratio, tolerance, ls_tolerance, meaninertia, iterations, ls_iterations;
  double base_pos[5][3], base_quat[5][4];
  double body_pos[LB][3], body_quat[LB][4], body_ipos[LB][3], body_imat[LB][9], body_mass[LB], body_inertia[LB][3], body_invw[LB];
  double jnt_axis[LB][3], jnt_pos[LB][3], qpos0[LEAP_NQ];
  double cube_Irot[9];
  double dof_damping[LEAP_NV], dof_frictionloss[LEAP_NV], dof_invw[LEAP_NV];
  double nfr, fr_dof[LEAP_NV], fr_solref[2], fr_solimp[5];
  double limited[16], lim_lo[16], lim_hi[16], lim_margin, lim_solref[2], lim_solimp[5];
  double kp[16], kv[16], ctrllimited[16], ctrl_lo[16], ctrl_hi[16];
  double ngeom, geom_type[LMAXG], geom_body[LMAXG], geom_pos[LMAXG][3], geom_mat[LMAXG][9], geom_size[LMAXG][3], geom_rbound[LMAXG], geom_mu[LMAXG];
  double cube_size[3], cube_rbound, con_solref[2], con_solimp[5];
  double site_body[5], site_pos[5][3];
  double fr_row[LEAP_NV];  // friction-loss row of each dof (-1: none)
  double geom_fr[LMAXG];   // sliding friction of each hand geom (hand-hand contacts mix by max)
  // hand-hand pairs that survive MuJoCo's static filters, body pair by body pair: 16-bit codes g1 * 256 + g2, four to a double (the
  // table is streamed once per step by every warp: 3.2 KB instead of 13 KB, L1 is ~28 KB next to 227 KB of shared memory)
  double nhh, hh_pair[LMAXHH / 4];
};

// per-warp shared-memory work area
struct LeapWork {
  double qpos[LEAP_NQ], qvel[LEAP_NV], warm[LEAP_NV], ctrl[LEAP_NU];
  double xpos[LB][3], xmat[LB][9], xanchor[LB][3], xaxis[LB][3];  // (xquat / xipos / Iw are dead after the mass matrix: they alias Jc, see below)
  double Mc[6];          // cube block: 3 translational masses are Mc[0..2]; rotational block is model.cube_Irot
  double Mf[4][4][4];    // finger blocks
  // Newton Hessian in arrow form (dof order: 6 cube dofs, then 4 fingers x 4): cube block, finger blocks, coupling
  double Hcc[6][6], Hcf[4][4][6], Hff[4][4][4], xc[6], yf[4][4];
  double qfrc_bias[LEAP_NV], qfrc_smooth[LEAP_NV], qacc_smooth[LEAP_NV], qacc[LEAP_NV], qfrc_constraint[LEAP_NV];
  double Ma[LEAP_NV], grad[LEAP_NV], search[LEAP_NV], Mv[LEAP_NV], tmp[LEAP_NV];
  double cdist[LMAXCON], cpos[LMAXCON][3], cframe[LMAXCON][9], cmu[LMAXCON], cfri[LMAXCON];  // (cframe doubles as the cone Hessian, see leap_cHc)
  double cDm[LMAXCON];   // D0 / (mu^2 (1 + mu^2)) of the contact's middle (cone) zone: fixed during the solve, divided once per step
  double Jc[3 * LMAXCON][10];  // compressed contact rows: 6 cube dofs + 4 dofs of the touched finger
  double eD[LMAXEFC], eR[LMAXEFC], earef[LMAXEFC], ejar[LMAXEFC], ejv[LMAXEFC], eforce[LMAXEFC];
  double efloss[LMAXFL], esign[LMAXFL];  // friction-loss / limit rows only
  int estate[LMAXEFC], edof[LMAXFL];
  // bodies of geom1 / geom2 (-1 static, 0 cube, 1 + 4f + d link d of finger f) and where their Jacobian entries sit in the compressed row:
  // cfa = first block Jc[.][0..5] (-1: the cube's 6 dofs, f >= 0: finger f's 4 dofs, -2: empty), cfinger = second block Jc[.][6..9] (finger, -1: empty)
  int cbA[LMAXCON], cbB[LMAXCON], cfa[LMAXCON], cfinger[LMAXCON];
  int limrow[16][2];   // efc row of joint j's lower / upper limit, -1 when inactive
  int ncon, nefc, nfl, ncand, solver_iter, ncross;  // ncross: contacts between two different fingers (they break the arrow structure)
  double cost, gauss;
};

// Narrow-phase scratch of ONE candidate pair (geom pose, raw contacts, clipping polygons).  LCOL_LANES of them alias the solver arrays
// cHc / cDm / Jc of the work area, which are dead until the constraint rows are built: the collision routines then touch no local memory
// (the kernel had a 1 152-byte stack frame per thread = 38 MB of local memory per launch, 17 MB of it reaching DRAM).
struct LeapColScratch { double gp[3], gm[9], gp2[3], gm2[9]; LRaw raw[8]; LBoxScratch box; };
constexpr int LCOL_LANES = 8;
static_assert(sizeof(LeapColScratch) * LCOL_LANES <= sizeof(double) * (LMAXCON + 3 * LMAXCON * 10), "collision scratch must fit into cDm + Jc");
static_assert(offsetof(LeapWork, Jc) == offsetof(LeapWork, cDm) + sizeof(double) * LMAXCON, "cDm, Jc must be contiguous");

// Arrays with disjoint lifetimes share storage (the work area decides how many rollouts fit on an SM):
//  * xquat / xipos / Iw are written by the kinematics and last read by the mass matrix + bias pass; Jc is written when the constraint rows
//    are built, later in the same step;
//  * a contact's frame is last read when its Jacobian rows are built; its 3x3 cone Hessian is first written by the solver after that;
//  * the solver's row arrays eD .. eforce are dead during collision detection: geom centres and candidate lists live there.
__device__ __forceinline__ double (*leap_xquat(LeapWork* W))[4] { return reinterpret_cast<double(*)[4]>(&W->Jc[0][0]); }
__device__ __forceinline__ double (*leap_xipos(LeapWork* W))[3] { return reinterpret_cast<double(*)[3]>(&W->Jc[0][0] + 4 * LB); }
__device__ __forceinline__ double (*leap_Iw(LeapWork* W))[6] { return reinterpret_cast<double(*)[6]>(&W->Jc[0][0] + 7 * LB); }
__device__ __forceinline__ double* leap_cHc(LeapWork* W, int c) { return W->cframe[c]; }
__device__ __forceinline__ const double* leap_cHc(const LeapWork* W, int c) { return W->cframe[c]; }
struct LeapColLists { double gc[LMAXG][4]; unsigned short hc[512], hf[512]; int cand[LMAXG]; };  // geom centre + bounding radius, hand-hand candidates, cube candidates
static_assert(sizeof(LeapColLists) <= sizeof(double) * 6 * LMAXEFC, "collision lists must fit into eD .. eforce");
static_assert(offsetof(LeapWork, eforce) == offsetof(LeapWork, eD) + sizeof(double) * 5 * LMAXEFC, "eD .. eforce must be contiguous");
__device__ __forceinline__ LeapColLists* leap_col_lists(LeapWork* W) { return reinterpret_cast<LeapColLists*>(&W->eD[0]); }

// optional phase timers (clock64 deltas accumulated by lane 0): kin, mass+bias, collision, constraints, smooth, solver, integrate,
// and inside the solver: update, direction (H + Cholesky + solve), line search, #newton iterations
__device__ unsigned long long g_leap_prof[24];  // [16..18]: dense Newton directions, hand-hand contacts, of those inside one finger / against the palm; [19..20]: hand-hand pairs past the bounding spheres / past the pre-filter; [21..23]: slowest block (cycles), sum over blocks, blocks
// per-block counters (prof mode): [0] block cycles, [1] lock-step Newton iterations the block walked, [2] iterations its warps were active in,
// [3] dense directions, [4] contacts (sum over warps and steps), [5] cube narrow-phase rounds, [6] hand-hand narrow-phase rounds, [7] launches,
// [8] line-search evaluations, [9] of those in iterations with finger-finger contacts, [10] cycles of those iterations (direction + line search + update)
__device__ unsigned long long g_leap_blk[160][12];
#define LPROF_BLK(slot, v) do { if (prof && lane == 0 && blockIdx.x < 160) atomicAdd(&g_leap_blk[blockIdx.x][slot], (unsigned long long)(v)); } while (0)
#define LPROF_T() (prof ? clock64() : 0)
#define LPROF_ADD(slot, t0) do { if (prof && lane == 0) atomicAdd(&g_leap_prof[slot], (unsigned long long)(clock64() - (t0))); } while (0)

enum { LST_SATISFIED = 0, LST_QUADRATIC = 1, LST_LINEARNEG = 2, LST_LINEARPOS = 3, LST_CONE = 4 };

// ------------------------------------------------------------------ kinematics (mj_kinematics + mj_comPos)
__device__ inline void leap_kinematics(const LeapModel* __restrict__ m, LeapWork* W, int lane) {
  double(*xquat)[4] = leap_xquat(W);
  double(*xipos)[3] = leap_xipos(W);
  double(*Iw)[6] = leap_Iw(W);
  if (lane == 0) {  // chain 0: the free cube
    lquat_normalize(W->qpos + 3);
#pragma unroll
    for (int k = 0; k < 3; k++) { W->xpos[0][k] = W->qpos[k]; W->xanchor[0][k] = W->qpos[k]; W->xaxis[0][k] = (k == 2); }
#pragma unroll
    for (int k = 0; k < 4; k++) xquat[0][k] = W->qpos[3 + k];
    lquat2mat(W->xmat[0], xquat[0]);
  } else if (lane <= 4) {  // chains 1..4: fingers hanging off the static palm
    const int f = lane - 1;
    double ppos[3], pquat[4], pmat[9];
#pragma unroll
    for (int k = 0; k < 3; k++) ppos[k] = m->base_pos[lane][k];
#pragma unroll
    for (int k = 0; k < 4; k++) pquat[k] = m->base_quat[lane][k];
    lquat2mat(pmat, pquat);
    for (int d = 0; d < 4; d++) {
      const int b = 1 + 4 * f + d;
      double pos[3], quat[4], t[3], mat[9], axis[3], off[3], anchor[3], dq[4], nq[4];
      lmat_vec(t, pmat, m->body_pos[b]);
#pragma unroll
      for (int k = 0; k < 3; k++) pos[k] = ppos[k] + t[k];
      lquat_mul(quat, pquat, m->body_quat[b]);
      lquat2mat(mat, quat);
      lmat_vec(axis, mat, m->jnt_axis[b]);
      lmat_vec(off, mat, m->jnt_pos[b]);
      const double q = W->qpos[7 + 4 * f + d] - m->qpos0[7 + 4 * f + d];
      double sn, cs;
      sincos(0.5 * q, &sn, &cs);
#pragma unroll
      for (int k = 0; k < 3; k++) anchor[k] = pos[k] + off[k];
      dq[0] = cs; dq[1] = sn * m->jnt_axis[b][0]; dq[2] = sn * m->jnt_axis[b][1]; dq[3] = sn * m->jnt_axis[b][2];
      lquat_mul(nq, quat, dq);
      lquat2mat(mat, nq);
      lmat_vec(off, mat, m->jnt_pos[b]);
#pragma unroll
      for (int k = 0; k < 3; k++) { pos[k] = anchor[k] - off[k]; W->xanchor[b][k] = anchor[k]; W->xaxis[b][k] = axis[k]; }
      lquat_normalize(nq);
      lquat2mat(mat, nq);
#pragma unroll
      for (int k = 0; k < 3; k++) { W->xpos[b][k] = pos[k]; ppos[k] = pos[k]; }
#pragma unroll
      for (int k = 0; k < 4; k++) { xquat[b][k] = nq[k]; pquat[k] = nq[k]; }
#pragma unroll
      for (int k = 0; k < 9; k++) { W->xmat[b][k] = mat[k]; pmat[k] = mat[k]; }
    }
  }
  __syncwarp();
  if (lane < LB) {  // inertial frames and world inertia tensors
    const int b = lane;
    double t[3], im[9];
    lmat_vec(t, W->xmat[b], m->body_ipos[b]);
#pragma unroll
    for (int k = 0; k < 3; k++) xipos[b][k] = W->xpos[b][k] + t[k];
    lmat_mul(im, W->xmat[b], m->body_imat[b]);
    int e = 0;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = r; c < 3; c++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) s += im[3 * r + k] * m->body_inertia[b][k] * im[3 * c + k];
        Iw[b][e++] = s;  // xx xy xz yy yz zz
      }
  }
  __syncwarp();
}

__device__ __forceinline__ void Iw_mul(double* r, const double* I6, const double* v) {
  r[0] = I6[0] * v[0] + I6[1] * v[1] + I6[2] * v[2];
  r[1] = I6[1] * v[0] + I6[3] * v[1] + I6[4] * v[2];
  r[2] = I6[2] * v[0] + I6[4] * v[1] + I6[5] * v[2];
}

// ------------------------------------------------------------------ mass matrix (mj_crb) and bias forces (mj_rne), lane per chain
__device__ inline void leap_mass_and_bias(const LeapModel* __restrict__ m, LeapWork* W, int lane) {
  const double(*xipos)[3] = leap_xipos(W);
  const double(*Iw)[6] = leap_Iw(W);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) W->Mc[k] = m->body_mass[0];
    // cube bias: gravity on the translational dofs; gyroscopic torque in the body frame
    double wl[3] = {W->qvel[3], W->qvel[4], W->qvel[5]}, Iwl[3], g[3];
    lmat_vec(Iwl, m->cube_Irot, wl);
    lcross3(g, wl, Iwl);
#pragma unroll
    for (int k = 0; k < 3; k++) { W->qfrc_bias[k] = -m->body_mass[0] * m->gravity[k]; W->qfrc_bias[3 + k] = g[k]; }
  } else if (lane <= 4) {
    const int f = lane - 1, b0 = 1 + 4 * f;
    // M_ij = sum_{b >= j} m_b (a_i x r_ib).(a_j x r_jb) + a_j . Iw_b a_i     (i <= j, r_ib = com_b - anchor_i)
    double c[4][4][3], Ia[4][4][3];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int b = i; b < 4; b++) {
        double r[3];
#pragma unroll
        for (int k = 0; k < 3; k++) r[k] = xipos[b0 + b][k] - W->xanchor[b0 + i][k];
        lcross3(c[i][b], W->xaxis[b0 + i], r);
        Iw_mul(Ia[i][b], Iw[b0 + b], W->xaxis[b0 + i]);
      }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = i; j < 4; j++) {
        double s = 0;
#pragma unroll
        for (int b = j; b < 4; b++) s += m->body_mass[b0 + b] * ldot3(c[i][b], c[j][b]) + ldot3(W->xaxis[b0 + j], Ia[i][b]);
        W->Mf[f][i][j] = s; W->Mf[f][j][i] = s;
      }
    // RNE with zero acceleration, classical Newton-Euler in world coordinates (base: w = 0, a = -g)
    double Wv[3] = {0, 0, 0}, A[3] = {0, 0, 0}, P[3], V[3] = {0, 0, 0}, Ac[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { P[k] = m->base_pos[lane][k]; Ac[k] = -m->gravity[k]; }
    double F[4][3], N0[4][3];
    for (int d = 0; d < 4; d++) {
      const int b = b0 + d;
      double r[3], t[3], t2[3];
#pragma unroll
      for (int k = 0; k < 3; k++) r[k] = W->xanchor[b][k] - P[k];
      lcross3(t, Wv, r);
#pragma unroll
      for (int k = 0; k < 3; k++) V[k] += t[k];
      lcross3(t2, Wv, t);
      lcross3(t, A, r);
#pragma unroll
      for (int k = 0; k < 3; k++) { Ac[k] += t[k] + t2[k]; P[k] = W->xanchor[b][k]; }
      const double qd = W->qvel[6 + 4 * f + d];
      double u[3] = {W->xaxis[b][0] * qd, W->xaxis[b][1] * qd, W->xaxis[b][2] * qd};
      lcross3(t, Wv, u);
#pragma unroll
      for (int k = 0; k < 3; k++) { A[k] += t[k]; Wv[k] += u[k]; }
      // move the reference point to the body origin (kept as P for the next link), then evaluate at the COM
#pragma unroll
      for (int k = 0; k < 3; k++) r[k] = W->xpos[b][k] - P[k];
      lcross3(t, Wv, r);
      lcross3(t2, Wv, t);
#pragma unroll
      for (int k = 0; k < 3; k++) V[k] += t[k];
      lcross3(t, A, r);
#pragma unroll
      for (int k = 0; k < 3; k++) { Ac[k] += t[k] + t2[k]; P[k] = W->xpos[b][k]; }
      double ac[3], Iwa[3], Iww[3], n[3];
#pragma unroll
      for (int k = 0; k < 3; k++) r[k] = xipos[b][k] - W->xpos[b][k];
      lcross3(t, Wv, r); lcross3(t2, Wv, t); lcross3(t, A, r);
#pragma unroll
      for (int k = 0; k < 3; k++) { ac[k] = Ac[k] + t[k] + t2[k]; F[d][k] = m->body_mass[b] * ac[k]; }
      Iw_mul(Iwa, Iw[b], A);
      Iw_mul(Iww, Iw[b], Wv);
      lcross3(t, Wv, Iww);
      lcross3(n, xipos[b], F[d]);
#pragma unroll
      for (int k = 0; k < 3; k++) N0[d][k] = Iwa[k] + t[k] + n[k];
    }
    for (int d = 3; d >= 0; d--) {
      const int b = b0 + d;
      double t[3], nn[3];
      lcross3(t, W->xanchor[b], F[d]);
#pragma unroll
      for (int k = 0; k < 3; k++) nn[k] = N0[d][k] - t[k];
      W->qfrc_bias[6 + 4 * f + d] = ldot3(W->xaxis[b], nn);
      if (d > 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { F[d - 1][k] += F[d][k]; N0[d - 1][k] += N0[d][k]; }
      }
    }
  }
  __syncwarp();
}

// y = M x (block diagonal), lane per dof
__device__ __forceinline__ double leap_mulM_row(const LeapModel* __restrict__ m, const LeapWork* W, const double* x, int i) {
  if (i < 3) return W->Mc[i] * x[i];
  if (i < 6) { const double* R = m->cube_Irot + 3 * (i - 3); return R[0] * x[3] + R[1] * x[4] + R[2] * x[5]; }
  const int f = (i - 6) >> 2, r = (i - 6) & 3;
  const double* row = W->Mf[f][r];
  const double* xx = x + 6 + 4 * f;
  return row[0] * xx[0] + row[1] * xx[1] + row[2] * xx[2] + row[3] * xx[3];
}

// x <- (M + diag(add))^-1 x, exploiting the block structure: cube (lane 0) and one finger per lane (1..4).
// Small dense Cholesky per block, same recurrences as the oracle's dense factorisation restricted to the block.
template <int N>
__device__ __forceinline__ void small_chol_solve(double (&A)[N][N], double* x) {
  double L[N][N];  // the diagonal holds RECIPROCAL pivots (one rsqrt per pivot; every later division becomes a multiplication)
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) {
      double s = A[i][j];
#pragma unroll
      for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
      if (i == j) { if (s < B2_MINVAL) s = B2_MINVAL; L[i][i] = L_RSQRT(s); } else L[i][j] = s * L[j][j];
    }
#pragma unroll
  for (int i = 0; i < N; i++) { double s = x[i];
#pragma unroll
    for (int k = 0; k < i; k++) s -= L[i][k] * x[k];
    x[i] = s * L[i][i]; }
#pragma unroll
  for (int i = N - 1; i >= 0; i--) { double s = x[i];
#pragma unroll
    for (int k = i + 1; k < N; k++) s -= L[k][i] * x[k];
    x[i] = s * L[i][i]; }
}
__device__ inline void leap_block_solve(const LeapModel* __restrict__ m, LeapWork* W, const double* add, double* x, int lane) {
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) x[k] = x[k] / (W->Mc[k] + (add ? add[k] : 0.0));
    double A[3][3], v[3] = {x[3], x[4], x[5]};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) A[i][j] = m->cube_Irot[3 * i + j] + ((i == j && add) ? add[3 + i] : 0.0);
    small_chol_solve<3>(A, v);
    x[3] = v[0]; x[4] = v[1]; x[5] = v[2];
  } else if (lane <= 4) {
    const int f = lane - 1;
    double A[4][4], v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { v[i] = x[6 + 4 * f + i];
#pragma unroll
      for (int j = 0; j < 4; j++) A[i][j] = W->Mf[f][i][j] + ((i == j && add) ? add[6 + 4 * f + i] : 0.0); }
    small_chol_solve<4>(A, v);
#pragma unroll
    for (int i = 0; i < 4; i++) x[6 + 4 * f + i] = v[i];
  }
  __syncwarp();
}

// ------------------------------------------------------------------ collision (routines in geom.cuh)
// world pose of hand geom g (static geoms are stored in the world frame)
__device__ __forceinline__ void leap_geom_pose(const LeapModel* __restrict__ m, const LeapWork* W, int g, double* gp, double* gm) {
  const int b = (int)m->geom_body[g];
  if (b < 0) {
    for (int k = 0; k < 3; k++) gp[k] = m->geom_pos[g][k];
    for (int k = 0; k < 9; k++) gm[k] = m->geom_mat[g][k];
  } else {
    double t[3];
    lmat_vec(t, W->xmat[b], m->geom_pos[g]);
    for (int k = 0; k < 3; k++) gp[k] = W->xpos[b][k] + t[k];
    lmat_mul(gm, W->xmat[b], m->geom_mat[g]);
  }
}

// Append this round's raw contacts (n per lane) behind the ncon already stored, in lane order (= pair order): exclusive prefix of n
// over the lanes.  bA / bB: bodies of geom1 / geom2.  Returns the new running total (which may exceed LMAXCON: the surplus is dropped).
__device__ __forceinline__ int leap_store_contacts(LeapWork* W, int lane, int n, const LRaw* raw, int ncon, int bA, int bB, double mu) {
  int pre = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += v; }
  const int total = __shfl_sync(FULL, pre, 31);
  pre -= n;
  for (int i = 0; i < n; i++) {
    const int slot = ncon + pre + i;
    if (slot < LMAXCON) {
      W->cdist[slot] = raw[i].dist;
      for (int k = 0; k < 3; k++) { W->cpos[slot][k] = raw[i].pos[k]; W->cframe[slot][k] = raw[i].normal[k]; }
      l_make_frame(W->cframe[slot]);
      const int fA = bA >= 1 ? (bA - 1) >> 2 : -1, fB = bB >= 1 ? (bB - 1) >> 2 : -1;
      W->cbA[slot] = bA; W->cbB[slot] = bB;
      if (bA == 0 || bB == 0) { W->cfa[slot] = -1; W->cfinger[slot] = bA == 0 ? fB : fA; }        // cube against a finger or the palm
      else if (fA >= 0 && fB >= 0 && fA != fB) { W->cfa[slot] = fA; W->cfinger[slot] = fB; }      // two different fingers
      else { W->cfa[slot] = -2; W->cfinger[slot] = fA >= 0 ? fA : fB; }                           // one finger: against itself or the palm
      W->cfri[slot] = mu;
    }
  }
  return ncon + total;
}

__device__ inline void leap_collision(const LeapModel* __restrict__ m, LeapWork* W, int lane, int prof = 0) {
  const int ng = (int)m->ngeom;
  long long tc0 = LPROF_T();
  const double* cp = W->xpos[0];
  // the solver's row arrays are dead until the constraint rows are built: they hold the world centres (+ bounding radii) of the hand
  // geoms and the candidate lists
  LeapColLists* CL = leap_col_lists(W);
  double(*gc)[4] = CL->gc;
  // broad phase: bounding spheres against the cube's; ordered compaction keeps the geom order of the pair list
  int base = 0;
  for (int g0 = 0; g0 < ng; g0 += 32) {
    const int g = g0 + lane;
    bool keep = false;
    if (g < ng) {
      const int b = (int)m->geom_body[g];
      double gp[3];
      if (b < 0) { gp[0] = m->geom_pos[g][0]; gp[1] = m->geom_pos[g][1]; gp[2] = m->geom_pos[g][2]; }
      else { double t[3]; lmat_vec(t, W->xmat[b], m->geom_pos[g]); gp[0] = W->xpos[b][0] + t[0]; gp[1] = W->xpos[b][1] + t[1]; gp[2] = W->xpos[b][2] + t[2]; }
      gc[g][0] = gp[0]; gc[g][1] = gp[1]; gc[g][2] = gp[2]; gc[g][3] = m->geom_rbound[g];
      double dc[3] = {gp[0] - cp[0], gp[1] - cp[1], gp[2] - cp[2]};
      keep = !(lnorm3(dc) > m->geom_rbound[g] + m->cube_rbound);  // the oracle's bounding-sphere test
      if (keep) {
        // conservative pre-filter (never rejects a touching pair): the 6 face axes of the SAT / exact sphere-box distance,
        // so that only a handful of pairs reach the divergent contact-generation pass
        double t[3];
        lmatT_vec(t, W->xmat[0], dc);  // geom centre in the cube frame
        if ((int)m->geom_type[g] == 6) {
          double gm[9], R[3][3];
          if (b < 0) { for (int k = 0; k < 9; k++) gm[k] = m->geom_mat[g][k]; } else lmat_mul(gm, W->xmat[b], m->geom_mat[g]);
          const double* m1 = W->xmat[0];
          for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) R[i][j] = fabs(m1[i] * gm[j] + m1[3 + i] * gm[3 + j] + m1[6 + i] * gm[6 + j]) + 1e-12;
          const double* s2 = m->geom_size[g];
          for (int i = 0; i < 3 && keep; i++)
            if (fabs(t[i]) - (m->cube_size[i] + s2[0] * R[i][0] + s2[1] * R[i][1] + s2[2] * R[i][2]) >= 1e-9) keep = false;
          if (keep) {
            double t2[3];
            lmatT_vec(t2, gm, dc);
            for (int j = 0; j < 3 && keep; j++)
              if (fabs(t2[j]) - (s2[j] + m->cube_size[0] * R[0][j] + m->cube_size[1] * R[1][j] + m->cube_size[2] * R[2][j]) >= 1e-9) keep = false;
          }
        } else {
          double d2 = 0;
          for (int k = 0; k < 3; k++) { const double ex = fabs(t[k]) - m->cube_size[k]; if (ex > 0) d2 += ex * ex; }
          const double rr = m->geom_size[g][0] + 1e-9;
          if (d2 >= rr * rr) keep = false;
        }
      }
    }
    const unsigned mask = __ballot_sync(FULL, keep);
    if (keep) CL->cand[base + __popc(mask & ((1u << lane) - 1))] = g;
    base += __popc(mask);
  }
  __syncwarp();
  const int ncand = base;
  LPROF_ADD(11, tc0); tc0 = LPROF_T();
  if (prof && lane == 0) atomicAdd(&g_leap_prof[13], (unsigned long long)ncand);
  int ncon = 0;
  LeapColScratch* scr_all = reinterpret_cast<LeapColScratch*>(&W->cDm[0]);
  for (int c0 = 0; c0 < ncand; c0 += LCOL_LANES) {  // LCOL_LANES pairs per round (2.7 candidates per step on average), one per lane
    const int ci = c0 + lane;
    LeapColScratch* S = scr_all + (lane < LCOL_LANES ? lane : 0);
    int n = 0, g = 0, b = -1, swap = 0;
    if (lane < LCOL_LANES && ci < ncand) {
      g = CL->cand[ci];
      b = (int)m->geom_body[g];
      leap_geom_pose(m, W, g, S->gp, S->gm);
      if ((int)m->geom_type[g] == 6) n = l_box_box(cp, W->xmat[0], m->cube_size, S->gp, S->gm, m->geom_size[g], 0.0, S->raw, 8, &S->box);  // geom1 = cube
      else { n = l_sphere_box(S->gp, m->geom_size[g][0], cp, W->xmat[0], m->cube_size, 0.0, S->raw); swap = 1; }          // geom1 = sphere
    }
    ncon = leap_store_contacts(W, lane, n, S->raw, ncon, swap ? b : 0, swap ? 0 : b, m->geom_mu[g]);
  }
  LPROF_ADD(12, tc0); tc0 = LPROF_T();
  // ---- hand-hand pairs.  (1) bounding spheres of every listed pair (squared form with a relative slack: a superset of the oracle's
  // survivors, the contact routines decide), (2) the conservative pre-filters on the survivors, (3) contact generation, LCOL_LANES
  // pairs per round.  Typical step: ~14 of 1 621 pairs pass (1), none pass (2).
  const int nhh = (int)m->nhh;
  unsigned short* hc = CL->hc;
  unsigned short* hf = CL->hf;
  constexpr int HCAP = 512;
  const unsigned short* hp = reinterpret_cast<const unsigned short*>(m->hh_pair);
  int nh = 0;
  // four pairs per lane and round, all table loads of a round in flight together, the next round's prefetched: the loop is bound by
  // load latency (global table -> shared geom centres), not by arithmetic
  int nxt[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { const int p = 32 * k + lane; nxt[k] = p < nhh ? (int)__ldg(hp + p) : -1; }
  for (int p0 = 0; p0 < nhh; p0 += 128) {
    int code[4];
#pragma unroll
    for (int k = 0; k < 4; k++) code[k] = nxt[k];
#pragma unroll
    for (int k = 0; k < 4; k++) { const int p = p0 + 128 + 32 * k + lane; nxt[k] = p < nhh ? (int)__ldg(hp + p) : -1; }
    bool keep[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      keep[k] = false;
      if (code[k] >= 0) {
        const double* a = gc[code[k] >> 8];
        const double* b = gc[code[k] & 255];
        const double dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2], rr = a[3] + b[3];
        keep[k] = dx * dx + dy * dy + dz * dz <= rr * rr * (1 + 1e-12);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const unsigned mask = __ballot_sync(FULL, keep[k]);
      const int at = nh + __popc(mask & ((1u << lane) - 1));
      if (keep[k] && at < HCAP) hc[at] = (unsigned short)code[k];
      nh += __popc(mask);
    }
  }
  __syncwarp();
  bool hh_overflow = nh > HCAP;
  if (hh_overflow) nh = HCAP;
  int nf = 0;
  for (int c0 = 0; c0 < nh; c0 += 32) {
    const int ci = c0 + lane;
    bool keep = false;
    int code = 0;
    if (ci < nh) {
      code = hc[ci];
      const int g1 = code >> 8, g2 = code & 255;
      const bool box1 = (int)m->geom_type[g1] == 6, box2 = (int)m->geom_type[g2] == 6;
      if (box1 && box2) {
        double p1[3], m1[9], p2[3], m2[9];
        leap_geom_pose(m, W, g1, p1, m1);
        leap_geom_pose(m, W, g2, p2, m2);
        keep = l_box_box_may_touch(p1, m1, m->geom_size[g1], p2, m2, m->geom_size[g2]);
      } else if (box1 != box2) {
        const int gs = box1 ? g2 : g1, gb = box1 ? g1 : g2;
        double pb[3], mb[9];
        leap_geom_pose(m, W, gb, pb, mb);
        keep = l_sphere_box_may_touch(gc[gs], m->geom_size[gs][0], pb, mb, m->geom_size[gb]);
      } else keep = true;  // sphere-sphere: the bounding spheres ARE the geoms
    }
    const unsigned mask = __ballot_sync(FULL, keep);
    if (keep) hf[nf + __popc(mask & ((1u << lane) - 1))] = (unsigned short)code;
    nf += __popc(mask);
  }
  __syncwarp();
  for (int c0 = 0; c0 < nf; c0 += LCOL_LANES) {
    const int ci = c0 + lane;
    LeapColScratch* S = scr_all + (lane < LCOL_LANES ? lane : 0);
    int n = 0, bA = -1, bB = -1;
    double mu = 0;
    if (lane < LCOL_LANES && ci < nf) {
      const int code = hf[ci];
      int g1 = code >> 8, g2 = code & 255;
      if ((int)m->geom_type[g1] > (int)m->geom_type[g2]) { const int t = g1; g1 = g2; g2 = t; }  // the oracle orders a pair by geom type (sphere < box)
      bA = (int)m->geom_body[g1]; bB = (int)m->geom_body[g2];
      mu = fmax(m->geom_fr[g1], m->geom_fr[g2]);
      leap_geom_pose(m, W, g1, S->gp, S->gm);
      leap_geom_pose(m, W, g2, S->gp2, S->gm2);
      const bool box1 = (int)m->geom_type[g1] == 6, box2 = (int)m->geom_type[g2] == 6;
      if (box1 && box2) n = l_box_box(S->gp, S->gm, m->geom_size[g1], S->gp2, S->gm2, m->geom_size[g2], 0.0, S->raw, 8, &S->box);
      else if (box2) n = l_sphere_box(S->gp, m->geom_size[g1][0], S->gp2, S->gm2, m->geom_size[g2], 0.0, S->raw);
      else n = l_sphere_sphere(S->gp, m->geom_size[g1][0], S->gp2, m->geom_size[g2][0], 0.0, S->raw);
    }
    ncon = leap_store_contacts(W, lane, n, S->raw, ncon, bA, bB, mu);
  }
  __syncwarp();
  const int nkept = ncon < LMAXCON ? ncon : LMAXCON;
  int ncross = 0;
  for (int c0 = 0; c0 < nkept; c0 += 32) ncross += __popc(__ballot_sync(FULL, c0 + lane < nkept && W->cfa[c0 + lane] >= 0));
  if (lane == 0) {
    W->ncon = nkept; W->ncross = ncross;
    if (ncon > LMAXCON || hh_overflow) atomicAdd(contact_overflow_counter(m), 1ull);
  }
  __syncwarp();
  LPROF_ADD(15, tc0);
  if (prof && lane == 0) {
    atomicAdd(&g_leap_prof[14], (unsigned long long)ncon);
    int nh2 = 0, none = 0;
    for (int c = 0; c < nkept; c++) { if (W->cbA[c] != 0 && W->cbB[c] != 0) { nh2++; if (W->cfa[c] == -2) none++; } }
    atomicAdd(&g_leap_prof[17], (unsigned long long)nh2); atomicAdd(&g_leap_prof[18], (unsigned long long)none);
    atomicAdd(&g_leap_prof[19], (unsigned long long)nh); atomicAdd(&g_leap_prof[20], (unsigned long long)nf);
    LPROF_BLK(4, nkept); LPROF_BLK(5, (ncand + LCOL_LANES - 1) / LCOL_LANES); LPROF_BLK(6, (nf + LCOL_LANES - 1) / LCOL_LANES);
  }
}

// ------------------------------------------------------------------ constraint rows (mj_makeConstraint + mj_makeImpedance)
__device__ __forceinline__ double leap_Jrow_dot(const LeapWork* W, int crow, const double* x) {
  const double* J = W->Jc[crow];
  const int fa = W->cfa[crow / 3], f = W->cfinger[crow / 3];
  double s = 0;
  if (fa == -1) s = J[0] * x[0] + J[1] * x[1] + J[2] * x[2] + J[3] * x[3] + J[4] * x[4] + J[5] * x[5];
  else if (fa >= 0) { const double* xx = x + 6 + 4 * fa; s = J[0] * xx[0] + J[1] * xx[1] + J[2] * xx[2] + J[3] * xx[3]; }
  if (f >= 0) { const double* xx = x + 6 + 4 * f; s += J[6] * xx[0] + J[7] * xx[1] + J[8] * xx[2] + J[9] * xx[3]; }
  return s;
}

__device__ inline void leap_make_constraint(const LeapModel* __restrict__ m, LeapWork* W, int lane) {
  const int nfr = (int)m->nfr;
  // friction-loss rows
  if (lane < nfr) {
    const int dof = (int)m->fr_dof[lane];
    W->edof[lane] = dof; W->esign[lane] = 1; W->efloss[lane] = m->dof_frictionloss[dof];
  }
  // joint-limit rows: joint-major, lower side before upper side
  bool lo_act = false, hi_act = false;
  double dlo = 0, dhi = 0;
  if (lane < 16 && m->limited[lane] != 0) {
    const double q = W->qpos[7 + lane];
    dlo = q - m->lim_lo[lane]; dhi = m->lim_hi[lane] - q;
    lo_act = dlo < m->lim_margin; hi_act = dhi < m->lim_margin;
  }
  const unsigned mlo = __ballot_sync(FULL, lo_act), mhi = __ballot_sync(FULL, hi_act);
  const unsigned below = (1u << lane) - 1;
  const int rbase = nfr + __popc(mlo & below) + __popc(mhi & below);
  if (lo_act) { W->edof[rbase] = 6 + lane; W->esign[rbase] = 1; W->efloss[rbase] = 0; W->ejar[rbase] = dlo; }
  if (hi_act) { const int r = rbase + (lo_act ? 1 : 0); W->edof[r] = 6 + lane; W->esign[r] = -1; W->efloss[r] = 0; W->ejar[r] = dhi; }
  if (lane < 16) { W->limrow[lane][0] = lo_act ? rbase : -1; W->limrow[lane][1] = hi_act ? rbase + (lo_act ? 1 : 0) : -1; }
  const int nfl = nfr + __popc(mlo) + __popc(mhi);
  const int ncon = W->ncon;
  __syncwarp();
  // contact Jacobians, lane per (contact, frame axis): J = frame^T (Jp(body2) - Jp(body1)) in the compressed layout of the work area
  // (first block: the cube's 6 dofs or a finger's 4, second block: a finger's 4; a contact inside ONE finger sums both sides into
  // the second block)
  for (int e = lane; e < 3 * ncon; e += 32) {
    const int c = e / 3, a = e - 3 * c;
    const double* fr = W->cframe[c] + 3 * a;
    const double* p = W->cpos[c];
    double* J = W->Jc[e];
#pragma unroll
    for (int k = 0; k < 10; k++) J[k] = 0;
    const int fa = W->cfa[c];
#pragma unroll
    for (int side = 0; side < 2; side++) {
      const int b = side ? W->cbB[c] : W->cbA[c];
      const double sgn = side ? 1.0 : -1.0;
      if (b == 0) {
        // cube: translation columns are the identity, rotation columns are (body axis) x (p - xpos)
        double r[3] = {p[0] - W->xpos[0][0], p[1] - W->xpos[0][1], p[2] - W->xpos[0][2]};
#pragma unroll
        for (int k = 0; k < 3; k++) J[k] = sgn * fr[k];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          double ax[3] = {W->xmat[0][k], W->xmat[0][3 + k], W->xmat[0][6 + k]}, cr[3];
          lcross3(cr, ax, r);
          J[3 + k] = sgn * ldot3(fr, cr);
        }
      } else if (b >= 1) {
        const int f = (b - 1) >> 2, dep = (b - 1) & 3, off = fa == f ? 0 : 6;
#pragma unroll
        for (int d = 0; d < 4; d++) {
          if (d <= dep) {
            const int bb = 1 + 4 * f + d;
            double rr[3] = {p[0] - W->xanchor[bb][0], p[1] - W->xanchor[bb][1], p[2] - W->xanchor[bb][2]}, cr[3];
            lcross3(cr, W->xaxis[bb], rr);
            J[off + d] += sgn * ldot3(fr, cr);
          }
        }
      }
    }
  }
  if (lane == 0) { W->nfl = nfl; W->nefc = nfl + 3 * ncon; }
  __syncwarp();
  // R, aref per row (mj_makeImpedance); contact rows get their cone adjustment afterwards
  const int nefc = nfl + 3 * ncon;
  for (int r = lane; r < nefc; r += 32) {
    const double *solref, *solimp;
    double pos, margin = 0, vel, diagA;
    bool friction_row = false;
    if (r < nfr) { solref = m->fr_solref; solimp = m->fr_solimp; pos = 0; vel = W->qvel[W->edof[r]]; diagA = m->dof_invw[W->edof[r]]; friction_row = true; }
    else if (r < nfl) { solref = m->lim_solref; solimp = m->lim_solimp; pos = W->ejar[r]; margin = m->lim_margin; vel = W->esign[r] * W->qvel[W->edof[r]]; diagA = m->dof_invw[W->edof[r]]; }
    else {
      const int e = r - nfl, c = e / 3;
      solref = m->con_solref; solimp = m->con_solimp; pos = W->cdist[c];
      vel = leap_Jrow_dot(W, e, W->qvel);
      const int bA = W->cbA[c], bB = W->cbB[c];
      diagA = (bA >= 0 ? m->body_invw[bA] : 0.0) + (bB >= 0 ? m->body_invw[bB] : 0.0);
      friction_row = (e - 3 * c) > 0;
    }
    double ref0 = solref[0], ref1 = solref[1];
    const double dmax = fmin(fmax(solimp[1], B2_MINIMP), B2_MAXIMP);
    const double imp = impedance(solimp, pos, margin);
    double K, B;
    if (ref0 > 0) {
      if (ref0 < 2 * m->dt) ref0 = 2 * m->dt;
      K = 1 / fmax(B2_MINVAL, dmax * dmax * ref0 * ref0 * ref1 * ref1);
      B = 2 / fmax(B2_MINVAL, dmax * ref0);
    } else { K = -ref0 / fmax(B2_MINVAL, dmax * dmax); B = -ref1 / fmax(B2_MINVAL, dmax); }
    if (friction_row) K = 0;
    W->eR[r] = fmax(B2_MINVAL, (1 - imp) * diagA / imp);
    W->earef[r] = -B * vel - K * imp * (pos - margin);
  }
  __syncwarp();
  for (int c = lane; c < ncon; c += 32) {
    const int i = nfl + 3 * c;
    const double R0 = W->eR[i];
    const double R1 = R0 / fmax(B2_MINVAL, m->impratio);
    W->eR[i + 1] = R1;
    W->cmu[c] = W->cfri[c] * sqrt(R1 / R0);
    W->eR[i + 2] = R1 * W->cfri[c] * W->cfri[c] / (W->cfri[c] * W->cfri[c]);
  }
  __syncwarp();
  for (int r = lane; r < nefc; r += 32) W->eD[r] = 1 / W->eR[r];
  __syncwarp();
  for (int c = lane; c < ncon; c += 32) { const double mu = W->cmu[c]; W->cDm[c] = W->eD[nfl + 3 * c] / (mu * mu * (1 + mu * mu)); }
  __syncwarp();
}

// ------------------------------------------------------------------ Newton solver (mj_solNewton, primal, elliptic cones)
// one elliptic contact at x[3]: returns cost, writes force[3], state, optional 3x3 cone Hessian
__device__ __noinline__ double leap_cone_eval(const LeapWork* W, int c, int row0, const double* x, double* force, int* state, double* Hc) {
  const double mu = W->cmu[c], f1 = W->cfri[c], f2 = W->cfri[c];
  const double D0 = W->eD[row0];
  const double U0 = x[0] * mu, U1 = x[1] * f1, U2 = x[2] * f2;
  const double N = U0, T = L_SQRT(U1 * U1 + U2 * U2);
  double cost = 0;
  if (N >= mu * T) { force[0] = force[1] = force[2] = 0; *state = LST_SATISFIED; }
  else if (mu * N + T <= 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) { force[k] = -W->eD[row0 + k] * x[k]; cost += 0.5 * W->eD[row0 + k] * x[k] * x[k]; }
    *state = LST_QUADRATIC;
  } else {
    const double Dm = W->cDm[c], NmT = N - mu * T;
    cost = 0.5 * Dm * NmT * NmT;
    force[0] = -Dm * NmT * mu;
    const double Tinv = T > B2_MINVAL ? L_RCP(T) : 0;
    force[1] = -force[0] * Tinv * U1 * f1;
    force[2] = -force[0] * Tinv * U2 * f2;
    *state = LST_CONE;
    if (Hc) {
      const double S[3] = {mu, f1, f2}, U[3] = {U0, U1, U2};
      double HU[9];
      const double Ti = Tinv;
      HU[0] = Dm;
#pragma unroll
      for (int j = 1; j < 3; j++) HU[j] = HU[3 * j] = -Dm * mu * U[j] * Ti;
#pragma unroll
      for (int j = 1; j < 3; j++)
#pragma unroll
        for (int k = 1; k < 3; k++)
          HU[3 * j + k] = Dm * mu * mu * U[j] * U[k] * Ti * Ti - Dm * mu * NmT * ((j == k ? Ti : 0) - U[j] * U[k] * Ti * Ti * Ti);
#pragma unroll
      for (int j = 0; j < 3; j++)
#pragma unroll
        for (int k = 0; k < 3; k++) Hc[3 * j + k] = S[j] * HU[3 * j + k] * S[k];
    }
  }
  return cost;
}

// friction-loss / limit row at x: returns cost, writes force and state
__device__ __forceinline__ double leap_row_eval(const LeapWork* W, int r, int nfr, double x, double* force, int* state) {
  const double D = W->eD[r];
  if (r < nfr) {
    const double f = W->efloss[r], R = W->eR[r];
    if (x <= -R * f) { *force = f; *state = LST_LINEARNEG; return -0.5 * R * f * f - f * x; }
    if (x >= R * f) { *force = -f; *state = LST_LINEARPOS; return -0.5 * R * f * f + f * x; }
    *force = -D * x; *state = LST_QUADRATIC; return 0.5 * D * x * x;
  }
  if (x < 0) { *force = -D * x; *state = LST_QUADRATIC; return 0.5 * D * x * x; }
  *force = 0; *state = LST_SATISFIED; return 0;
}

__device__ __forceinline__ double lwsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// two independent sums reduced together: the shuffle rounds interleave, so the pair costs the latency of one reduction
__device__ __forceinline__ void lwsum2(double& a, double& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ta = __shfl_xor_sync(FULL, a, o), tb = __shfl_xor_sync(FULL, b, o);
    a += ta; b += tb;
  }
}

// jar = J qacc - aref for every row; Ma = M qacc
__device__ __noinline__ void leap_set_point(const LeapModel* __restrict__ m, LeapWork* W, const double* qacc, int lane) {
  if (lane < LEAP_NV) W->Ma[lane] = leap_mulM_row(m, W, qacc, lane);
  const int nfl = W->nfl, nefc = W->nefc;
  for (int r = lane; r < nefc; r += 32) {
    const double v = r < nfl ? W->esign[r] * qacc[W->edof[r]] : leap_Jrow_dot(W, r - nfl, qacc);
    W->ejar[r] = v - W->earef[r];
  }
  __syncwarp();
}

// cost / forces / states at the current jar (+ gradient and, optionally, the Newton Hessian)
__device__ __noinline__ void leap_constraint_update(const LeapModel* __restrict__ m, LeapWork* W, const double* qacc, bool want_h, int lane) {
  const int nfr = (int)m->nfr, nfl = W->nfl, ncon = W->ncon;
  double cost = 0;
  for (int r = lane; r < nfl; r += 32) cost += leap_row_eval(W, r, nfr, W->ejar[r], &W->eforce[r], &W->estate[r]);
  for (int c = lane; c < ncon; c += 32) {
    const int r0 = nfl + 3 * c;
    cost += leap_cone_eval(W, c, r0, W->ejar + r0, W->eforce + r0, &W->estate[r0], want_h ? leap_cHc(W, c) : nullptr);
    W->estate[r0 + 1] = W->estate[r0 + 2] = W->estate[r0];
  }
  cost = lwsum(cost);
  __syncwarp();
  // qfrc_constraint = J^T force, gradient, Gauss term — lane per dof
  double g = 0;
  if (lane < LEAP_NV) {
    const int i = lane;
    double f = 0;
    if (i >= 6) {  // friction-loss row of dof i is row fr_row[i] (O(1) lookups instead of a search over the rows)
      const int fr = (int)m->fr_row[i];
      if (fr >= 0) f += W->eforce[fr];
      const int rl = W->limrow[i - 6][0], rh = W->limrow[i - 6][1];
      if (rl >= 0) f += W->eforce[rl];
      if (rh >= 0) f -= W->eforce[rh];
    }
    const int fi = i >= 6 ? (i - 6) >> 2 : -1, ki = i < 6 ? i : 6 + ((i - 6) & 3);
    for (int c = 0; c < ncon; c++) {
      if (i < 6 ? W->cfa[c] != -1 : W->cfinger[c] != fi) continue;
#pragma unroll
      for (int a = 0; a < 3; a++) f += W->Jc[3 * c + a][ki] * W->eforce[nfl + 3 * c + a];
    }
    if (i >= 6 && W->ncross > 0)  // contacts between two fingers keep the first finger's entries in the first block
      for (int c = 0; c < ncon; c++) {
        if (W->cfa[c] != fi) continue;
#pragma unroll
        for (int a = 0; a < 3; a++) f += W->Jc[3 * c + a][ki - 6] * W->eforce[nfl + 3 * c + a];
      }
    W->qfrc_constraint[i] = f;
    W->grad[i] = W->Ma[i] - W->qfrc_smooth[i] - f;
    g = (W->Ma[i] - W->qfrc_smooth[i]) * (qacc[i] - W->qacc_smooth[i]);
  }
  g = lwsum(g);
  if (lane == 0) { W->gauss = 0.5 * g; W->cost = 0.5 * g + cost; }
  __syncwarp();
}

// lower-triangle index tables packed 3 bits per entry: 6x6 (21 entries) and 4x4 (10 entries)
#define TRI6_I 0x5b6db2491b6d2448ull  /* 0,1,1,2,2,2,3,3,3,3,4,4,4,4,4,5,5,5,5,5,5 */
#define TRI6_J 0x58d111a21a211040ull /* 0,0,1,0,1,2,0,1,2,3,0,1,2,3,4,0,1,2,3,4,5 */
#define TRI4_I 0x1b6d2448ull         /* 0,1,1,2,2,2,3,3,3,3 */
#define TRI4_J 0x1a211040ull         /* 0,0,1,0,1,2,0,1,2,3 */
__device__ __forceinline__ int tri_get(unsigned long long tab, int e) { return (int)((tab >> (3 * e)) & 7ull); }

// Tail shared by the two direction routines: with X = L_F^-1 B in Hcf (16 x 6, rows = finger dofs) and y = L_F^-1 (-grad_F) in yf,
// Schur complement of the cube block, its Cholesky and both substitutions in one lane -> xc = search[0..5].
__device__ __forceinline__ void leap_schur_cube(const LeapModel* __restrict__ m, LeapWork* W, int lane) {
  // Schur complement S = C - sum_f X_f^T X_f (lanes 0..20) and reduced right-hand side (lanes 21..26)
  if (lane < 21) {
    const int i = tri_get(TRI6_I, lane), j = tri_get(TRI6_J, lane);
    double sacc = W->Hcc[i][j];
#pragma unroll
    for (int f = 0; f < 4; f++)
#pragma unroll
      for (int r = 0; r < 4; r++) sacc -= W->Hcf[f][r][i] * W->Hcf[f][r][j];
    W->Hcc[i][j] = sacc;
  } else if (lane < 27) {
    const int cc = lane - 21;
    double sacc = -W->grad[cc];
#pragma unroll
    for (int f = 0; f < 4; f++)
#pragma unroll
      for (int r = 0; r < 4; r++) sacc -= W->Hcf[f][r][cc] * W->yf[f][r];
    W->xc[cc] = sacc;
  }
  __syncwarp();
  // cube block: Cholesky + both substitutions in one lane (6x6)
  if (lane == 0) {
    double L[6][6], x[6];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) {
        double sacc = W->Hcc[i][j];
#pragma unroll
        for (int k = 0; k < j; k++) sacc -= L[i][k] * L[j][k];
        if (i == j) { if (sacc < B2_MINVAL) sacc = B2_MINVAL; L[i][i] = L_RSQRT(sacc); } else L[i][j] = sacc * L[j][j];
      }
#pragma unroll
    for (int i = 0; i < 6; i++) { double sacc = W->xc[i];
#pragma unroll
      for (int k = 0; k < i; k++) sacc -= L[i][k] * x[k];
      x[i] = sacc * L[i][i]; }
#pragma unroll
    for (int i = 5; i >= 0; i--) { double sacc = x[i];
#pragma unroll
      for (int k = i + 1; k < 6; k++) sacc -= L[k][i] * x[k];
      x[i] = sacc * L[i][i]; }
#pragma unroll
    for (int i = 0; i < 6; i++) { W->xc[i] = x[i]; W->search[i] = x[i]; }
  }
  __syncwarp();
}

// Finger part of the Newton direction when two DIFFERENT fingers touch each other (~0.8 % of the Newton iterations of the C4 bench
// scenario, but the rollouts they sit in are the slowest of their blocks): the finger blocks couple, so the 16 finger dofs are
// factorised as ONE dense block by the warp (lane per row, left-looking Cholesky) instead of four 4 x 4 blocks; cube block, couplings
// and the assembly of everything but the cross-finger terms stay those of the arrow form.  The 16 x 16 matrix (packed lower triangle)
// aliases the kinematics arrays, which are dead once the constraint rows are built.
static_assert(16 * 17 / 2 <= LB * 18, "packed finger Hessian must fit into xpos .. xaxis");
static_assert(offsetof(LeapWork, Mc) == offsetof(LeapWork, xpos) + sizeof(double) * LB * 18, "xpos .. xaxis must be contiguous");
#define LF(i, j) Fp[(i) * ((i) + 1) / 2 + (j)] /* lower triangle, row-packed */
__device__ __noinline__ void leap_coupled_fingers(const LeapModel* __restrict__ m, LeapWork* W, int lane) {
  const int nfl = W->nfl, ncon = W->ncon;
  double* Fp = &W->xpos[0][0];
  double* dinv = W->tmp;
  double* Bm = &W->Hcf[0][0][0];  // (16, 6): row 4 f + r
  double* yv = &W->yf[0][0];      // (16)
  // block diagonal from the assembled finger blocks, zeros elsewhere (lane i writes row i)
  if (lane < 16) {
    double* row = &LF(lane, 0);
#pragma unroll
    for (int j = 0; j < 16; j++)
      if (j <= lane) row[j] = (j >> 2) == (lane >> 2) ? W->Hff[lane >> 2][lane & 3][j & 3] : 0.0;
  }
  __syncwarp();
  // cross-finger contacts: first finger's own block (lanes 0..9) and the coupling block (lanes 10..25); the second finger's own block
  // was accumulated by the arrow assembly
  for (int c = 0; c < ncon; c++) {
    const int fa = W->cfa[c];
    if (fa < 0) continue;
    const int st = W->estate[nfl + 3 * c];
    if (st == LST_SATISFIED) continue;
    const int fb = W->cfinger[c];
    double Wm[9];
    if (st == LST_QUADRATIC) {
#pragma unroll
      for (int k = 0; k < 9; k++) Wm[k] = 0;
      Wm[0] = W->eD[nfl + 3 * c]; Wm[4] = W->eD[nfl + 3 * c + 1]; Wm[8] = W->eD[nfl + 3 * c + 2];
    } else {
#pragma unroll
      for (int k = 0; k < 9; k++) Wm[k] = leap_cHc(W, c)[k];
    }
    const double(*J)[10] = &W->Jc[3 * c];
    int si = -1, sj = 0, row = 0, col = 0;
    if (lane < 10) { si = tri_get(TRI4_I, lane); sj = tri_get(TRI4_J, lane); row = 4 * fa + si; col = 4 * fa + sj; }
    else if (lane < 26) { const int r = (lane - 10) >> 2, q = (lane - 10) & 3; si = 6 + r; sj = q; row = 4 * fb + r; col = 4 * fa + q; }
    if (si >= 0) {
      double h = 0;
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) h += Wm[3 * a + b] * J[a][si] * J[b][sj];
      if (row >= col) LF(row, col) += h; else LF(col, row) += h;
    }
    __syncwarp();
  }
  // Left-looking Cholesky of the finger block, lane i < 16 owns row i.  The six cube-coupling columns of B and the right-hand side ride
  // along as EXTRA ROWS (lanes 16..21: B^T, lane 22: -grad_F): what the elimination leaves in them is X^T = (L^-1 B)^T and
  // y^T = (L^-1 (-grad_F))^T, so the two forward substitutions cost nothing beyond the 16 column steps.  dinv[k] = 1 / L_kk.
  double* Lx = Fp + 136;       // (6, 16): row cc = column cc of B
  double* gr = Fp + 136 + 96;  // (16)
  static_assert(136 + 96 + 16 <= LB * 18, "finger block + coupling rows + rhs must fit into xpos .. xaxis");
  for (int e = lane; e < 96; e += 32) { const int cc = e >> 4, k = e & 15; Lx[16 * cc + k] = Bm[6 * k + cc]; }
  if (lane < 16) gr[lane] = -W->grad[6 + lane];
  __syncwarp();
  double* myrow = lane < 16 ? &LF(lane, 0) : (lane < 22 ? Lx + 16 * (lane - 16) : gr);
  for (int k = 0; k < 16; k++) {
    double sres = 0;
    if (lane >= k && lane < 23) {
      const double* rk = &LF(k, 0);
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      int j = 0;
      for (; j + 3 < k; j += 4) { s0 += myrow[j] * rk[j]; s1 += myrow[j + 1] * rk[j + 1]; s2 += myrow[j + 2] * rk[j + 2]; s3 += myrow[j + 3] * rk[j + 3]; }
      for (; j < k; j++) s0 += myrow[j] * rk[j];
      sres = myrow[k] - ((s0 + s1) + (s2 + s3));
    }
    double d = __shfl_sync(FULL, sres, k);
    if (d < B2_MINVAL) d = B2_MINVAL;
    const double rs = L_RSQRT(d);
    if (lane == k) dinv[k] = rs;
    else if (lane > k && lane < 23) myrow[k] = sres * rs;
    __syncwarp();
  }
  const double di = lane < 16 ? dinv[lane] : 0.0;
  double x;
  for (int e = lane; e < 96; e += 32) { const int cc = e >> 4, k = e & 15; Bm[6 * k + cc] = Lx[16 * cc + k]; }  // X = L^-1 B, (16, 6)
  if (lane < 16) yv[lane] = gr[lane];
  __syncwarp();
  leap_schur_cube(m, W, lane);
  // fingers: x_F = L^-T (y - X x_C)
  x = 0;
  if (lane < 16) {
    x = yv[lane];
#pragma unroll
    for (int cc = 0; cc < 6; cc++) x -= Bm[6 * lane + cc] * W->xc[cc];
  }
  for (int k = 15; k >= 0; k--) {
    const double xk = __shfl_sync(FULL, x * di, k);
    if (lane == k) x = xk;
    else if (lane < k) x -= LF(k, lane) * xk;
  }
  if (lane < 16) W->search[6 + lane] = x;
  __syncwarp();
}
#undef LF

// Newton direction: search = -H^-1 grad, H = M + sum_rows D J^T J (+ elliptic cone blocks).
// Contacts only involve the cube, so H is an ARROW matrix: cube block C (6x6), four independent finger blocks F_f (4x4)
// and couplings B_f (4x6).  Eliminating the fingers first:  L_f = chol(F_f),  X_f = L_f^-1 B_f,
// S = C - sum_f X_f^T X_f,  L_S = chol(S)  — sequential depth 4 + 6 instead of 22, compact code, tiny shared footprint.
__device__ inline void leap_newton_direction(const LeapModel* __restrict__ m, LeapWork* W, int lane) {
  const int nfr = (int)m->nfr, nfl = W->nfl, ncon = W->ncon;
  // ---- assemble: mass matrix
  for (int e = lane; e < 36; e += 32) {
    const int i = e / 6, j = e - 6 * i;
    W->Hcc[i][j] = (i < 3) ? (i == j ? W->Mc[i] : 0.0) : (j >= 3 ? m->cube_Irot[3 * (i - 3) + (j - 3)] : 0.0);
  }
  for (int e = lane; e < 96; e += 32) (&W->Hcf[0][0][0])[e] = 0.0;
  for (int e = lane; e < 64; e += 32) (&W->Hff[0][0][0])[e] = (&W->Mf[0][0][0])[e];
  __syncwarp();
  // friction-loss rows, then limit rows, in their quadratic zone: D on the diagonal (two passes: a dof can own both kinds)
  for (int pass = 0; pass < 2; pass++) {
    const int r = lane;
    const bool mine = pass == 0 ? r < nfr : (r >= nfr && r < nfl);
    if (mine && W->estate[r] == LST_QUADRATIC) {
      const int dof = W->edof[r];
      if (dof < 6) W->Hcc[dof][dof] += W->eD[r]; else W->Hff[(dof - 6) >> 2][(dof - 6) & 3][(dof - 6) & 3] += W->eD[r];
    }
    __syncwarp();
  }
  // contacts: J^T Wc J with Wc = diag(D) (quadratic zone) or the 3x3 cone Hessian
  for (int c = 0; c < ncon; c++) {
    const int st = W->estate[nfl + 3 * c];
    if (st == LST_SATISFIED) continue;
    const int fc = W->cfinger[c];
    double Wm[9];
    if (st == LST_QUADRATIC) {
#pragma unroll
      for (int k = 0; k < 9; k++) Wm[k] = 0;
      Wm[0] = W->eD[nfl + 3 * c]; Wm[4] = W->eD[nfl + 3 * c + 1]; Wm[8] = W->eD[nfl + 3 * c + 2];
    } else {
#pragma unroll
      for (int k = 0; k < 9; k++) Wm[k] = leap_cHc(W, c)[k];
    }
    const double(*J)[10] = &W->Jc[3 * c];
    // pass A: cube-cube (lanes 0..20) and finger-finger (lanes 21..30)
    int ia = -1, ja = -1;
    double* dst = nullptr;
    const bool with_cube = W->cfa[c] == -1;  // (a contact inside one finger / finger-palm only touches that finger's block)
    if (lane < 21) { if (with_cube) { ia = tri_get(TRI6_I, lane); ja = tri_get(TRI6_J, lane); dst = &W->Hcc[ia][ja]; } }
    else if (lane < 31 && fc >= 0) { const int e = lane - 21; ia = 6 + tri_get(TRI4_I, e); ja = 6 + tri_get(TRI4_J, e); dst = &W->Hff[fc][ia - 6][ja - 6]; }
    if (dst) {
      double h = 0;
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) h += Wm[3 * a + b] * J[a][ia] * J[b][ja];
      *dst += h;
    }
    // pass B: finger-cube coupling (lanes 0..23)
    if (with_cube && fc >= 0 && lane < 24) {
      const int r = lane / 6, cc = lane - 6 * r;
      double h = 0;
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) h += Wm[3 * a + b] * J[a][6 + r] * J[b][cc];
      W->Hcf[fc][r][cc] += h;
    }
  }
  __syncwarp();  // (every Hessian entry is accumulated by ONE lane across all contacts: no barrier is needed inside the loop)
  if (W->ncross > 0) { leap_coupled_fingers(m, W, lane); return; }
  // ---- factorise the finger blocks (lane f), keep 1/L_kk on the diagonal slot for the substitutions
  if (lane < 4) {
    const int f = lane;
    double L[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) {
        double sacc = W->Hff[f][i][j];
#pragma unroll
        for (int k = 0; k < j; k++) sacc -= L[i][k] * L[j][k];
        // pivots are kept as RECIPROCALS (one rsqrt per pivot, every later division becomes a multiplication)
        if (i == j) { if (sacc < B2_MINVAL) sacc = B2_MINVAL; L[i][i] = L_RSQRT(sacc); } else L[i][j] = sacc * L[j][j];
      }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) W->Hff[f][i][j] = L[i][j];
    // y_f = L_f^-1 (-grad_f)
    double y[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { double sacc = -W->grad[6 + 4 * f + i];
#pragma unroll
      for (int k = 0; k < i; k++) sacc -= L[i][k] * y[k];
      y[i] = sacc * L[i][i]; W->yf[f][i] = y[i]; }
  }
  __syncwarp();
  // X_f = L_f^-1 B_f, one (finger, cube column) per lane
  if (lane < 24) {
    const int f = lane / 6, cc = lane - 6 * f;
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { double sacc = W->Hcf[f][i][cc];
#pragma unroll
      for (int k = 0; k < i; k++) sacc -= W->Hff[f][i][k] * x[k];
      x[i] = sacc * W->Hff[f][i][i]; }
#pragma unroll
    for (int i = 0; i < 4; i++) W->Hcf[f][i][cc] = x[i];
  }
  __syncwarp();
  leap_schur_cube(m, W, lane);
  // fingers: x_f = L_f^-T (y_f - X_f x_C)
  if (lane < 4) {
    const int f = lane;
    double t[4];
#pragma unroll
    for (int r = 0; r < 4; r++) { double sacc = W->yf[f][r];
#pragma unroll
      for (int cc = 0; cc < 6; cc++) sacc -= W->Hcf[f][r][cc] * W->xc[cc];
      t[r] = sacc; }
#pragma unroll
    for (int i = 3; i >= 0; i--) { double sacc = t[i];
#pragma unroll
      for (int k = i + 1; k < 4; k++) sacc -= W->Hff[f][k][i] * t[k];
      t[i] = sacc * W->Hff[f][i][i]; }
#pragma unroll
    for (int i = 0; i < 4; i++) W->search[6 + 4 * f + i] = t[i];
  }
  __syncwarp();
}

// Line search state held in REGISTERS: lane r owns friction/limit row r (nfl <= 32) and lane c owns contact c, so one evaluation of
// the 1-D cost derivatives is register math plus two warp reductions.  Contacts 32 .. LMAXCON-1 (rare: more than 32 contacts in a step)
// are evaluated from shared memory by the low lanes on top of their own.
struct LeapLS {
  double rjar, rjv, rD, rR, rfl;           // my friction/limit row
  double cjar[3], cjv[3], cD[3], cmu, cfr, cDm;  // my contact
  bool has_row, row_is_friction, has_con;
};

__device__ __forceinline__ void leap_ls_load(const LeapModel* __restrict__ m, const LeapWork* W, int lane, LeapLS& L) {
  const int nfr = (int)m->nfr, nfl = W->nfl;
  L.has_row = lane < nfl; L.row_is_friction = lane < nfr; L.has_con = lane < W->ncon;
  if (L.has_row) { L.rjar = W->ejar[lane]; L.rjv = W->ejv[lane]; L.rD = W->eD[lane]; L.rR = W->eR[lane]; L.rfl = W->efloss[lane]; }
  if (L.has_con) {
    const int r0 = nfl + 3 * lane;
#pragma unroll
    for (int k = 0; k < 3; k++) { L.cjar[k] = W->ejar[r0 + k]; L.cjv[k] = W->ejv[r0 + k]; L.cD[k] = W->eD[r0 + k]; }
    L.cmu = W->cmu[lane]; L.cfr = W->cfri[lane]; L.cDm = W->cDm[lane];
  }
}

// one elliptic contact's contribution to the first / second derivative of the cost along the search direction at step alpha
__device__ __forceinline__ void leap_ls_contact(const double* cjar, const double* cjv, const double* cD, double mu, double f, double Dm, double alpha,
                                                double& p1, double& p2) {
  const double x0 = cjar[0] + alpha * cjv[0], x1 = cjar[1] + alpha * cjv[1], x2 = cjar[2] + alpha * cjv[2];
  const double U1 = x1 * f, U2 = x2 * f;
  const double T2 = U1 * U1 + U2 * U2;
  const double Ti = T2 > B2_MINVAL * B2_MINVAL ? L_RSQRT(T2) : 0;  // one rsqrt instead of a square root and a division
  const double N = x0 * mu, T = T2 * Ti;
  if (N >= mu * T) { /* separating: no force */ }
  else if (mu * N + T <= 0) {
    p1 += cD[0] * x0 * cjv[0] + cD[1] * x1 * cjv[1] + cD[2] * x2 * cjv[2];
    p2 += cD[0] * cjv[0] * cjv[0] + cD[1] * cjv[1] * cjv[1] + cD[2] * cjv[2] * cjv[2];
  } else {
    // s = 0.5 Dm (N - mu T)^2 in the scaled space U = S x; chain rule with dU/dalpha = S jv
    const double NmT = N - mu * T;
    const double v0 = mu * cjv[0], v1 = f * cjv[1], v2 = f * cjv[2];
    const double dT = (U1 * v1 + U2 * v2) * Ti;
    const double dNmT = v0 - mu * dT;
    const double d2T = (v1 * v1 + v2 * v2 - dT * dT) * Ti;  // curvature of |U_T| along a line
    p1 += Dm * NmT * dNmT;
    p2 += Dm * (dNmT * dNmT - NmT * mu * d2T);
  }
}

// first and second derivative of the cost along the search direction at step alpha (all lanes get the result)
__device__ __forceinline__ void leap_ls_eval(const LeapLS& L, const LeapWork* W, int lane, double alpha, double g1, double g2, double* d1, double* d2) {
  double p1 = 0, p2 = 0;
  if (L.has_row) {
    const double x = L.rjar + alpha * L.rjv;
    if (L.row_is_friction) {
      const double lim = L.rR * L.rfl;
      if (x <= -lim) p1 -= L.rfl * L.rjv;
      else if (x >= lim) p1 += L.rfl * L.rjv;
      else { p1 += L.rD * x * L.rjv; p2 += L.rD * L.rjv * L.rjv; }
    } else if (x < 0) { p1 += L.rD * x * L.rjv; p2 += L.rD * L.rjv * L.rjv; }
  }
  if (L.has_con) leap_ls_contact(L.cjar, L.cjv, L.cD, L.cmu, L.cfr, L.cDm, alpha, p1, p2);
  if (W->ncon > 32 && lane + 32 < W->ncon) {
    const int c = lane + 32, r0 = W->nfl + 3 * c;
    leap_ls_contact(W->ejar + r0, W->ejv + r0, W->eD + r0, W->cmu[c], W->cfri[c], W->cDm[c], alpha, p1, p2);
  }
  lwsum2(p1, p2);
  *d1 = g1 + alpha * g2 + p1;
  *d2 = g2 + p2;
}

__device__ inline double leap_line_search(const LeapModel* __restrict__ m, const LeapWork* W, int lane) {
  double g1 = 0, g2 = 0, sn = 0;
  if (lane < LEAP_NV) { g1 = W->search[lane] * (W->Ma[lane] - W->qfrc_smooth[lane]); g2 = W->search[lane] * W->Mv[lane]; sn = W->search[lane] * W->search[lane]; }
  lwsum2(g1, g2);
  double gs = lane < LEAP_NV ? W->grad[lane] * W->search[lane] : 0.0;
  lwsum2(sn, gs);
  const double snorm = L_SQRT(sn);
  if (snorm < B2_MINVAL) return 0;
  const double gtol = m->tolerance * m->ls_tolerance * snorm * m->meaninertia * LEAP_NV;
  LeapLS L;
  leap_ls_load(m, W, lane, L);
  // derivatives at alpha = 0 without evaluating the rows: d1(0) = grad . search, and for the Newton direction H search = -grad
  // gives d2(0) = search^T H search = -d1(0), i.e. the first trial step is the full Newton step
  double d1 = gs, d2 = -d1, lo = 0, hi = -1, alpha;
  if (d1 >= 0 || d2 <= 0) return 0;
  alpha = 1.0;
  double prev_step = 1e300;
  const int iters = (int)m->ls_iterations;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    leap_ls_eval(L, W, lane, alpha, g1, g2, &d1, &d2);
    if (lane == 0) const_cast<LeapWork*>(W)->solver_iter++;  // (diagnostics: line-search evaluations of this call)
    if (fabs(d1) < gtol) return alpha;
    if (d1 < 0) lo = alpha; else hi = alpha;
    double next = d2 > 0 ? alpha - d1 * L_RCP(fmax(d2, 1e-200)) : -1;
    if (hi < 0) { if (!(next > lo)) next = 2 * alpha + B2_MINVAL; }
    else if (!(next > lo && next < hi && fabs(next - alpha) < 0.5 * prev_step)) next = 0.5 * (lo + hi);
    if (next == alpha) return alpha;
    prev_step = fabs(next - alpha);
    alpha = next;
  }
  return lo > 0 ? lo : alpha;
}

// mj_fwdConstraint.  Called by ALL warps of the block; with sync_mode >= 3 the Newton iterations of the block's warps run in
// lock-step (block barrier per iteration, finished warps idle) so that the iteration body is fetched once for all of them.
__device__ inline void leap_fwd_constraint(const LeapModel* __restrict__ m, LeapWork* W, int lane, int prof, bool active, int sync_mode) {
  const int nefc = active ? W->nefc : 0;
  bool done = !active;
  if (active && nefc == 0) {
    if (lane < LEAP_NV) { W->qacc[lane] = W->qacc_smooth[lane]; W->qfrc_constraint[lane] = 0; }
    __syncwarp();
    done = true;
  }
  double scale = 0;
  int nfl = 0;
  if (!done) {
    // warm start: keep qacc_warmstart unless qacc_smooth has lower cost.  qacc_smooth is evaluated FIRST so that in the common case
    // (the warm start wins) the residuals / forces / gradient / cone Hessians left behind are already those of the chosen point
    leap_set_point(m, W, W->qacc_smooth, lane);
    leap_constraint_update(m, W, W->qacc_smooth, false, lane);
    const double cs = W->cost;
    __syncwarp();
    leap_set_point(m, W, W->warm, lane);
    leap_constraint_update(m, W, W->warm, true, lane);
    const double cw = W->cost;
    __syncwarp();
    if (lane < LEAP_NV) W->qacc[lane] = cw > cs ? W->qacc_smooth[lane] : W->warm[lane];
    __syncwarp();
    if (cw > cs) {
      leap_set_point(m, W, W->qacc, lane);
      leap_constraint_update(m, W, W->qacc, true, lane);
    }
    scale = 1.0 / (m->meaninertia * LEAP_NV);
    nfl = W->nfl;
  }
  const int iters = (int)m->iterations;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    if (sync_mode >= 3) { if (!__syncthreads_or(done ? 0 : 1)) break; }
    else if (done) break;
    if (threadIdx.x < 32) LPROF_BLK(1, 1);
    if (done) continue;
    LPROF_BLK(2, 1);
    double gn = lane < LEAP_NV ? W->grad[lane] * W->grad[lane] : 0.0;
    gn = lwsum(gn);
    if (scale * L_SQRT(gn) < m->tolerance) { done = true; continue; }
    long long t1 = LPROF_T();
    const long long t_iter = t1;
    leap_newton_direction(m, W, lane);
    if (W->ncross > 0) { if (prof && lane == 0) atomicAdd(&g_leap_prof[16], 1ull); LPROF_BLK(3, 1); if (prof) { LPROF_BLK(10, clock64() - t1); LPROF_BLK(9, W->ncon); } }
    LPROF_ADD(8, t1); t1 = LPROF_T();
    if (lane < LEAP_NV) W->Mv[lane] = leap_mulM_row(m, W, W->search, lane);
    for (int r = lane; r < nefc; r += 32) W->ejv[r] = r < nfl ? W->esign[r] * W->search[W->edof[r]] : leap_Jrow_dot(W, r - nfl, W->search);
    __syncwarp();
    if (lane == 0) W->solver_iter = 0;
    const double alpha = leap_line_search(m, W, lane);
    if (prof) { __syncwarp(); LPROF_BLK(8, W->solver_iter); if (W->ncross > 0) LPROF_BLK(11, clock64() - t1); }
    LPROF_ADD(9, t1); t1 = LPROF_T();
    if (lane == 0 && prof) atomicAdd(&g_leap_prof[10], 1ull);
    if (alpha == 0) { done = true; continue; }
    const double oldcost = W->cost;
    __syncwarp();
    if (lane < LEAP_NV) { W->qacc[lane] += alpha * W->search[lane]; W->Ma[lane] += alpha * W->Mv[lane]; }
    for (int r = lane; r < nefc; r += 32) W->ejar[r] += alpha * W->ejv[r];
    __syncwarp();
    leap_constraint_update(m, W, W->qacc, true, lane);
    LPROF_ADD(7, t1);
    if (prof && W->ncross > 0) LPROF_BLK(7 + 0 * (int)(clock64() - t_iter), 0);
    const double newcost = W->cost;
    __syncwarp();
    if (scale * (oldcost - newcost) < m->tolerance) done = true;
  }
}

// ------------------------------------------------------------------ one mj_step
// All warps of the block call this together; `active` masks the tail warps.  The two block barriers keep the warps in
// the same code region (see the kernel comment) — they are NOT data dependencies.
__device__ inline void leap_step(const LeapModel* __restrict__ m, LeapWork* W, int lane, double* sens /* global, may be null */, int prof,
                                 bool active, int sync_mode, double* trace = nullptr /* global: the 5 framepos trace sensors only */) {
  long long t0 = LPROF_T();
  if (active) {
    leap_kinematics(m, W, lane);
    LPROF_ADD(0, t0); t0 = LPROF_T();
    leap_mass_and_bias(m, W, lane);
    LPROF_ADD(1, t0);
  }
  if (sync_mode >= 2) __syncthreads();
  t0 = LPROF_T();
  if (active) {
    leap_collision(m, W, lane, prof);
    LPROF_ADD(2, t0); t0 = LPROF_T();
    leap_make_constraint(m, W, lane);
    LPROF_ADD(3, t0); t0 = LPROF_T();
    if (trace && lane < 5) {  // fused mode: only the trace sensors (T1, controller.py:323-363) are kept, for every rollout
      const int b = (int)m->site_body[lane];
      double t[3];
      lmat_vec(t, W->xmat[b], m->site_pos[lane]);
      for (int k = 0; k < 3; k++) trace[3 * lane + k] = W->xpos[b][k] + t[k];
    }
    if (sens) {  // position-stage sensors: 16 jointpos then 5 framepos sites (pre-step state)
      if (lane < 16) sens[lane] = W->qpos[7 + lane];
      if (lane < 5) {
        const int b = (int)m->site_body[lane];
        double t[3];
        lmat_vec(t, W->xmat[b], m->site_pos[lane]);
        for (int k = 0; k < 3; k++) sens[16 + 3 * lane + k] = W->xpos[b][k] + t[k];
      }
    }
    // passive + actuation -> qfrc_smooth; qacc_smooth = M^-1 qfrc_smooth
    if (lane < LEAP_NV) {
      const int i = lane;
      double f = -m->dof_damping[i] * W->qvel[i] - W->qfrc_bias[i];
      if (i >= 6) {
        const int a = i - 6;
        double u = W->ctrl[a];
        if (m->ctrllimited[a] != 0) u = fmin(fmax(u, m->ctrl_lo[a]), m->ctrl_hi[a]);
        f += m->kp[a] * u - m->kp[a] * W->qpos[7 + a] - m->kv[a] * W->qvel[i];
      }
      W->qfrc_smooth[i] = f; W->qacc_smooth[i] = f;
    }
    __syncwarp();
    leap_block_solve(m, W, nullptr, W->qacc_smooth, lane);
    LPROF_ADD(4, t0);
  }
  if (sync_mode >= 2) __syncthreads();
  t0 = LPROF_T();
  leap_fwd_constraint(m, W, lane, prof, active, sync_mode);
  __syncwarp();
  if (!active) return;
  LPROF_ADD(5, t0); t0 = LPROF_T();
  // implicitfast: (M + h (damping + kv)) qacc = qfrc_smooth + qfrc_constraint, then semi-implicit advance
  const double h = m->dt;
  if (lane < LEAP_NV) {
    W->tmp[lane] = h * (m->dof_damping[lane] + (lane >= 6 ? m->kv[lane - 6] : 0.0));
    W->Mv[lane] = W->qfrc_smooth[lane] + W->qfrc_constraint[lane];
  }
  __syncwarp();
  leap_block_solve(m, W, W->tmp, W->Mv, lane);
  if (lane < LEAP_NV) { W->qvel[lane] += h * W->Mv[lane]; W->warm[lane] = W->qacc[lane]; }
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
  } else if (lane >= 6 && lane < LEAP_NV) W->qpos[lane + 1] += h * W->qvel[lane];
  __syncwarp();
  LPROF_ADD(6, t0);
}

// per-step cost (leap_cube.py:76-86): 0.5 w_pos |p - goal|^2 + 0.5 w_rot |log(q* (x) q_goal)|^2 ; params [w_pos, w_rot, goal_quat4, goal_pos3]
__device__ inline double leap_cost(const double* p, const double* qpos) {
  const double dx = qpos[0] - p[6], dy = qpos[1] - p[7], dz = qpos[2] - p[8];
  const double u0 = qpos[3], u1 = -qpos[4], u2 = -qpos[5], u3 = -qpos[6];
  const double v0 = p[2], v1 = p[3], v2 = p[4], v3 = p[5];
  const double w = u0 * v0 - u1 * v1 - u2 * v2 - u3 * v3;
  const double x = u0 * v1 + u1 * v0 + u2 * v3 - u3 * v2;
  const double y = u0 * v2 - u1 * v3 + u2 * v0 + u3 * v1;
  const double z = u0 * v3 + u1 * v2 - u2 * v1 + u3 * v0;
  const double s = sqrt(x * x + y * y + z * z);
  double ax = 1, ay = 0, az = 0;
  if (!(s < 1e-6)) { ax = x / s; ay = y / s; az = z / s; }
  double speed = 2.0 * atan2(s, w);
  if (speed > 3.14159265358979323846) speed -= 2 * 3.14159265358979323846;
  const double rx = ax * speed, ry = ay * speed, rz = az * speed;
  return p[0] * 0.5 * (dx * dx + dy * dy + dz * dz) + p[1] * 0.5 * (rx * rx + ry * ry + rz * rz);
}

// ------------------------------------------------------------------ kernels
// COST: in = knots (N,K,16), basis (H,K) -> reward (N) [+ cost (N,H) f32];  !COST: in = controls (N,H,16) -> states, sensors
// Block = up to 7 warps (7 rollouts, one SM's worth of shared memory).  The warps of a block are re-aligned with a
// __syncthreads() at every time step so that they walk through the same code region together: the kernel's instruction
// footprint far exceeds the SM's instruction caches, and warps drifting apart would each stream it from L2 on their own.
template <bool COST>
__global__ void __launch_bounds__(224) leap_rollout_kernel(const LeapModel* __restrict__ m, const double* __restrict__ x0, int x0_batched,
                                                           const double* __restrict__ in, int N, int H, int K, const double* __restrict__ basis,
                                                           const double* __restrict__ cost_params, double* __restrict__ states,
                                                           double* __restrict__ sensors, float* __restrict__ cost_NH, double* __restrict__ reward_N,
                                                           int wstride, int prof, const SampleSpec smp, int index_offset,
                                                           double* __restrict__ trace_out = nullptr /* COST: (N, H, 15) or null */) {
  const int sync_period = (prof >> 16) > 0 ? (prof >> 16) : 1;  // experiment knob: block barrier every sync_period time steps (sync_mode 1)
  const int sync_mode = (prof >> 8) & 255;
  prof &= 255;
  B2_DYNAMIC_SMEM(unsigned char, lsm_all);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int n = blockIdx.x * wpb + wib;
  const bool active = n < N;
  const long long t_block0 = LPROF_T();
  unsigned char* lsm = lsm_all + (size_t)wib * wstride;
  LeapWork* W = reinterpret_cast<LeapWork*>(lsm);
  if (active) {
    const double* xs = x0 + (x0_batched ? (size_t)n * LEAP_NX : 0);
    if (lane < LEAP_NQ) W->qpos[lane] = xs[lane];
    if (lane < LEAP_NV) { W->qvel[lane] = xs[LEAP_NQ + lane]; W->warm[lane] = 0; }
  }
  __syncwarp();
  if constexpr (COST) {
    // stage this rollout's knots (K*16 doubles) behind the work area with one TMA bulk copy; the (H,K) basis is shared by every
    // rollout and is read from global memory (K cached loads per lane and step) to keep the per-warp shared-memory footprint small
    uint64_t* bar = reinterpret_cast<uint64_t*>(lsm + ((sizeof(LeapWork) + 15) & ~(size_t)15));
    double* sK = reinterpret_cast<double*>(bar + 2);
    float* sC = reinterpret_cast<float*>(sK + K * LEAP_NU);  // this rollout's cost row (H floats), written back coalesced at the end
    if (active) {
      const unsigned bytesK = (unsigned)(K * LEAP_NU * sizeof(double));
      const double* gK = in + (size_t)n * K * LEAP_NU;
      if (!smp.enabled) {
        if ((reinterpret_cast<uintptr_t>(gK) & 15) == 0) {
          if (lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
          __syncwarp();
          if (lane == 0) { mbar_expect_tx(bar, bytesK); tma_bulk_g2s(sK, gK, bytesK, bar); }
          mbar_wait(bar, 0);
        } else {
          for (int i = lane; i < K * LEAP_NU; i += 32) sK[i] = gK[i];
          __syncwarp();
        }
      }
      if (smp.enabled) {  // on-device sampling: lane p draws the normal pair for elements 2p, 2p+1 (Philox keyed by the global index)
        const long long gn = (long long)n + index_offset;
        const int KNU = K * LEAP_NU;
        for (int p2 = lane; 2 * p2 < KNU; p2 += 32) {
          double z0, z1;
          normal_pair(smp, gn, p2, &z0, &z1);
          const double a = sample_element(smp, gn, 2 * p2, LEAP_NU, z0);
          sK[2 * p2] = a; smp.knots_out[(size_t)n * KNU + 2 * p2] = a;
          if (2 * p2 + 1 < KNU) { const double b = sample_element(smp, gn, 2 * p2 + 1, LEAP_NU, z1); sK[2 * p2 + 1] = b; smp.knots_out[(size_t)n * KNU + 2 * p2 + 1] = b; }
        }
        __syncwarp();
      }
    }
    double total = 0;
#pragma unroll 1
    for (int t = 0; t < H; t++) {
      if (sync_mode >= 1 && (sync_period == 1 || t % sync_period == 0)) __syncthreads();
      if (active && lane < LEAP_NU) {
        double u = 0;
        for (int k = 0; k < K; k++) u += __ldg(basis + t * K + k) * sK[k * LEAP_NU + lane];
        W->ctrl[lane] = u;
      }
      __syncwarp();
      leap_step(m, W, lane, nullptr, prof, active, sync_mode, (trace_out && active) ? trace_out + ((size_t)n * H + t) * LEAP_NTRACE : nullptr);
      if (active && lane == 0) {
        double cp[LEAP_NCOST];
#pragma unroll
        for (int i = 0; i < LEAP_NCOST; i++) cp[i] = cost_params[i];
        const double ct = leap_cost(cp, W->qpos);
        total += ct;
        if (cost_NH) sC[t] = (float)ct;
      }
    }
    if (active && lane == 0) reward_N[n] = -(total / H);
    if (prof) {
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned long long dt = (unsigned long long)(clock64() - t_block0);
        atomicMax(&g_leap_prof[21], dt); atomicAdd(&g_leap_prof[22], dt); atomicAdd(&g_leap_prof[23], 1ull);
        if (blockIdx.x < 160) { atomicAdd(&g_leap_blk[blockIdx.x][0], dt); atomicAdd(&g_leap_blk[blockIdx.x][7], 1ull); }
      }
    }
    if (active && cost_NH) {
      __syncwarp();
      for (int i = lane; i < H; i += 32) cost_NH[(size_t)n * H + i] = sC[i];  // one coalesced row per rollout instead of H scalar stores
    }
  } else {
#pragma unroll 1
    for (int t = 0; t < H; t++) {
      if (sync_mode >= 1) __syncthreads();
      if (active && lane < LEAP_NU) W->ctrl[lane] = in[((size_t)n * H + t) * LEAP_NU + lane];
      __syncwarp();
      leap_step(m, W, lane, (sensors && active) ? sensors + ((size_t)n * H + t) * LEAP_NS : nullptr, prof, active, sync_mode);
      if (!active) continue;
      double* so = states + ((size_t)n * H + t) * LEAP_NX;
      if (lane < LEAP_NQ) so[lane] = W->qpos[lane];
      if (lane < LEAP_NV) so[LEAP_NQ + lane] = W->qvel[lane];
    }
  }
}

__global__ void leap_reward_kernel(const double* __restrict__ states, int N, int H, const double* __restrict__ cost_params, double* __restrict__ reward_N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double cp[LEAP_NCOST];
#pragma unroll
  for (int i = 0; i < LEAP_NCOST; i++) cp[i] = cost_params[i];
  double total = 0;
  for (int t = 0; t < H; t++) total += leap_cost(cp, states + ((size_t)n * H + t) * LEAP_NX);
  reward_N[n] = -(total / H);
}

// ------------------------------------------------------------------ host side
// shared memory per warp: work area + mbarrier + (fused mode) this rollout's knots and its cost row
inline size_t leap_wstride(int cost_mode, int K, int H) {
  size_t w = ((sizeof(LeapWork) + 15) & ~(size_t)15) + 16 + (cost_mode ? (size_t)K * LEAP_NU * sizeof(double) + (size_t)H * sizeof(float) : 0);
  return (w + 15) & ~(size_t)15;
}
#ifndef B2_HOST_SIM
inline int leap_create(LeapModel** out, const double* consts, size_t n, std::string* err) {
  if (n != sizeof(LeapModel) / sizeof(double)) { *err = "wrong number of task constants"; return 1; }
  LeapModel* d = nullptr;
  if (cudaMalloc(&d, sizeof(LeapModel) + 16) != cudaSuccess) { *err = "cudaMalloc failed"; return 1; }  // + the handle's contact-overflow counter (geom.cuh)
  if (cudaMemset(d + 1, 0, 16) != cudaSuccess) { cudaFree(d); *err = "cudaMemset failed"; return 1; }
  if (cudaMemcpy(d, consts, sizeof(LeapModel), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); *err = "cudaMemcpy failed"; return 1; }
  *out = d;
  return 0;
}
inline void leap_destroy(LeapModel* m) { cudaFree(m); }
inline void leap_prof_dump() {
  unsigned long long h[24];
  if (cudaMemcpyFromSymbol(h, g_leap_prof, sizeof(h)) != cudaSuccess) return;
  const char* names[24] = {"kinematics", "mass+bias", "collision", "constraints", "smooth", "solver(total)", "integrate", "  update", "  direction", "  linesearch", "newton iters", "  coll:broad", "  coll:narrow", "candidates", "contacts", "  coll:hand-hand", "dense directions", "hand-hand contacts", "  one finger/palm", "hh past spheres", "hh past prefilter", "slowest block", "sum of blocks", "blocks"};
  for (int i = 0; i < 24; i++) fprintf(stderr, "leap_prof %-14s %llu\n", names[i], h[i]);
  static unsigned long long hb[160][12];
  if (cudaMemcpyFromSymbol(hb, g_leap_blk, sizeof(hb)) == cudaSuccess) {
    int order[160];
    for (int i = 0; i < 160; i++) order[i] = i;
    for (int i = 0; i < 160; i++) for (int j = i + 1; j < 160; j++) if (hb[order[j]][0] > hb[order[i]][0]) { int t = order[i]; order[i] = order[j]; order[j] = t; }
    fprintf(stderr, "leap_blk  block: cycles/launch lockstep-iters active-iters dense contacts cube-rounds hh-rounds (per launch; slowest 6, median, fastest)\n");
    const int show[8] = {0, 1, 2, 3, 4, 5, 73, 146};
    for (int q = 0; q < 8; q++) {
      const int b = order[show[q]];
      const double L = hb[b][7] ? (double)hb[b][7] : 1.0;
      fprintf(stderr, "leap_blk %4d: %.0f %.1f %.1f %.1f %.1f %.1f %.1f | ls evals %.1f; finger-finger iterations: contacts %.1f, direction cycles %.0f, line-search cycles %.0f\n", b, hb[b][0] / L, hb[b][1] / L, hb[b][2] / L, hb[b][3] / L, hb[b][4] / L, hb[b][5] / L, hb[b][6] / L, hb[b][8] / L, hb[b][9] / L, hb[b][10] / L, hb[b][11] / L);
    }
    memset(hb, 0, sizeof(hb));
    cudaMemcpyToSymbol(g_leap_blk, hb, sizeof(hb));
  }
  memset(h, 0, sizeof(h));
  cudaMemcpyToSymbol(g_leap_prof, h, sizeof(h));
}
inline int leap_num_partials(int N) { return N; }

inline int leap_launch(const LeapModel* m, int cost_mode, const double* d_x0, int batched, const double* d_in, int N, int H, int K,
                       const double* d_basis, const double* d_params, double* d_states, double* d_sensors, float* d_cost, double* d_reward,
                       const PlanEpilogue& ep, const SampleSpec& smp, cudaStream_t st, std::string* err, double* d_trace = nullptr) {
  const char* sm_env = getenv("B200MPC_LEAP_SYNC");
  const char* sp_env = getenv("B200MPC_LEAP_SYNC_PERIOD");
  const int prof = (getenv("B200MPC_LEAP_PROF") ? 1 : 0) | ((sm_env ? atoi(sm_env) : 3) << 8) | ((sp_env ? atoi(sp_env) : 1) << 16);
  (void)ep;  // the leap path runs the optimizer update as separate reduction kernels (b200mpc.cu)
  const size_t wstride = leap_wstride(cost_mode, K, H);
  int wpb = (N + 147) / 148;  // spread the rollouts over the 148 SMs first, then stack up to 7 warps per SM
  if (wpb < 1) wpb = 1;
  if (wpb > 7) wpb = 7;
  if (const char* wenv = getenv("B200MPC_LEAP_WPB")) wpb = atoi(wenv) > 0 ? atoi(wenv) : wpb;  // experiment knob: warps per block
  while (wpb > 1 && wpb * wstride > 226 * 1024) wpb--;
  const size_t smem = wpb * wstride;
  if (smem > 227 * 1024) { *err = "horizon/knots too large for the shared-memory tile"; return 1; }
  const int grid = (N + wpb - 1) / wpb;
  cudaError_t e;
  if (cost_mode) {
    e = set_max_dynamic_smem_once((const void*)leap_rollout_kernel<true>, 1, smem);
    if (e == cudaSuccess)
      leap_rollout_kernel<true><<<grid, 32 * wpb, smem, st>>>(m, d_x0, batched, d_in, N, H, K, d_basis, d_params, nullptr, nullptr, d_cost, d_reward, (int)wstride, prof, smp, ep.index_offset, d_trace);
  } else {
    e = set_max_dynamic_smem_once((const void*)leap_rollout_kernel<false>, 0, smem);
    if (e == cudaSuccess)
      leap_rollout_kernel<false><<<grid, 32 * wpb, smem, st>>>(m, d_x0, batched, d_in, N, H, 0, nullptr, nullptr, d_states, d_sensors, nullptr, nullptr, (int)wstride, prof, smp, 0);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("leap launch: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

inline int leap_reward_launch(const LeapModel* m, const double* d_states, int N, int H, const double* d_params, double* d_reward, cudaStream_t st,
                              std::string* err) {
  (void)m;
  leap_reward_kernel<<<(N + 127) / 128, 128, 0, st>>>(d_states, N, H, d_params, d_reward);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *err = std::string("leap reward launch: ") + cudaGetErrorString(e); return 1; }
  return 0;
}
#endif  // !B2_HOST_SIM

}  // namespace b2
