// small_tasks.cuh — closed-form forward dynamics + per-step cost for cartpole and cylinder_push.
//
// One CUDA thread owns one rollout (2 and 4 dofs).  Each `step` is one MuJoCo mj_step specialised to the
// features the task's MJCF switches on, with constants taken from the
// compiled table (judo_b200/models/*.json <- /root/reference/judo/models/xml/{cartpole,cylinder_push}.xml).
// `cost` is the per-step term of the task's reward (judo/tasks/cartpole.py:66-71, cylinder_push.py:69-88).
#pragma once
#include "common.cuh"

namespace b2 {

// =========================================================================================== cartpole (D1, C1r)
// All-double POD; field order == judo_b200/consts.py:CARTPOLE_FIELDS.
struct CartpoleConsts {
  double dt, gravity, m_cart, m_pole, l_pole, iyy_pole, damp_cart, damp_pole;
  double kp, ctrllimited, ctrl_lo, ctrl_hi, forcelimited, frc_lo, frc_hi;
  double limited, lim_lo, lim_hi, lim_margin, solref[2], solimp[5], invweight_cart;
  double site_cart[3], site_pole[3];
  double meaninertia, tolerance, ls_tolerance, iterations, ls_iterations;
};

struct CartpoleTask {
  static constexpr int NQ = 2, NV = 2, NU = 1, NS = 6, NX = 4, NCOST = 6;
  using Consts = CartpoleConsts;
  struct State { double q[2], v[2], warm[2]; };

  __device__ static inline void load(State& s, const double* x) {
    s.q[0] = x[0]; s.q[1] = x[1]; s.v[0] = x[2]; s.v[1] = x[3]; s.warm[0] = s.warm[1] = 0;
  }
  __device__ static inline void store(const State& s, double* x) { x[0] = s.q[0]; x[1] = s.q[1]; x[2] = s.v[0]; x[3] = s.v[1]; }

  // cart: slide along x; pole: hinge about +y at the cart origin, COM at (0,0,l) in the pole frame.
  __device__ static inline void step(const Consts& c, State& s, const double* u, double* sens) {
    double sn, cs;
    sincos(s.q[1], &sn, &cs);
    if (sens) {  // framepos sensors are evaluated in mj_forward, i.e. at the pre-step state
      sens[0] = s.q[0] + c.site_cart[0]; sens[1] = c.site_cart[1]; sens[2] = c.site_cart[2];
      sens[3] = s.q[0] + cs * c.site_pole[0] + sn * c.site_pole[2];
      sens[4] = c.site_pole[1];
      sens[5] = -sn * c.site_pole[0] + cs * c.site_pole[2];
    }
    const double h = c.dt, mp = c.m_pole, l = c.l_pole;
    double M[2][2];
    M[0][0] = c.m_cart + mp; M[0][1] = M[1][0] = mp * l * cs; M[1][1] = c.iyy_pole + mp * l * l;
    // joint-limit rows (mj_instantiateLimit): active when dist < margin
    double J[2][2], D[2], aref[2];
    int nefc = 0;
    if (c.limited != 0) {
#pragma unroll
      for (int side = -1; side <= 1; side += 2) {
        double dist = side * ((side < 0 ? c.lim_lo : c.lim_hi) - s.q[0]);
        if (dist < c.lim_margin) {
          double R, ar, jx = -side;
          row_reference(c.solref, c.solimp, h, dist, c.lim_margin, jx * s.v[0], c.invweight_cart, &R, &ar);
          J[nefc][0] = jx; J[nefc][1] = 0; D[nefc] = 1 / R; aref[nefc] = ar; nefc++;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; r++) if (r >= nefc) { J[r][0] = J[r][1] = 0; D[r] = 0; aref[r] = 0; }
    // passive (joint damping), bias (Coriolis + gravity), actuation (position servo with ctrl/force clamps)
    double uc = u[0];
    if (c.ctrllimited != 0) uc = fmin(fmax(uc, c.ctrl_lo), c.ctrl_hi);
    double fa = c.kp * uc - c.kp * s.q[0];
    if (c.forcelimited != 0) fa = fmin(fmax(fa, c.frc_lo), c.frc_hi);
    double bias0 = -mp * l * sn * s.v[1] * s.v[1];
    double bias1 = -mp * c.gravity * l * sn;
    double qfs[2] = {-c.damp_cart * s.v[0] - bias0 + fa, -c.damp_pole * s.v[1] - bias1};
    double L[2][2], qas[2] = {qfs[0], qfs[1]};
    chol<2>(L, M);
    chol_solve<2>(L, qas);
    double qacc[2], qfc[2];
    SolverOpt o{c.meaninertia, c.tolerance, c.ls_tolerance, (int)c.iterations, (int)c.ls_iterations};
    RowSolver<2, 2> sol;
    sol.solve(M, qfs, qas, J, D, aref, nefc, s.warm, o, qacc, qfc);
    // mj_Euler: implicit joint damping, then semi-implicit advance
    double qa[2];
    if (c.damp_cart > 0 || c.damp_pole > 0) {
      double A[2][2] = {{M[0][0] + h * c.damp_cart, M[0][1]}, {M[1][0], M[1][1] + h * c.damp_pole}};
      qa[0] = qfs[0] + qfc[0]; qa[1] = qfs[1] + qfc[1];
      chol<2>(L, A);
      chol_solve<2>(L, qa);
    } else { qa[0] = qacc[0]; qa[1] = qacc[1]; }
    s.v[0] += h * qa[0]; s.v[1] += h * qa[1];
    s.q[0] += h * s.v[0]; s.q[1] += h * s.v[1];
    s.warm[0] = qacc[0]; s.warm[1] = qacc[1];
  }

  // cost params: [w_vertical, w_centered, w_velocity, w_control, p_vertical, p_centered]
  __device__ static inline double cost(const double* p, const State& s, const double* u) {
    double cz = cos(s.q[1]) - 1;
    double vertical = sqrt(cz * cz + p[4] * p[4]) - p[4];
    double centered = sqrt(s.q[0] * s.q[0] + p[5] * p[5]) - p[5];
    double vel = 0.5 * (s.v[0] * s.v[0] + s.v[1] * s.v[1]);
    double ctl = 0.5 * u[0] * u[0];
    return p[0] * vertical + p[1] * centered + p[2] * vel + p[3] * ctl;
  }
  __device__ static inline double finish(double sum, int H) { return -sum; }
};

// =========================================================================================== cylinder_push (D2, C2r)
// Field order == judo_b200/consts.py:CYLINDER_PUSH_FIELDS.
struct CylinderPushConsts {
  double dt, mass_pusher, mass_cart, damp[4];
  double kp, ctrllimited, ctrl_lo, ctrl_hi, forcelimited, frc_lo, frc_hi;
  double r_pusher, r_cart, margin, gap, solref[2], solimp[5], mu, tran, impratio;
  double site_pusher[3], site_cart[3];
  double meaninertia, tolerance, ls_tolerance, iterations, ls_iterations;
};

struct CylinderPushTask {
  static constexpr int NQ = 4, NV = 4, NU = 2, NS = 6, NX = 8, NCOST = 6;
  using Consts = CylinderPushConsts;
  struct State { double q[4], v[4], warm[4]; };

  __device__ static inline void load(State& s, const double* x) {
#pragma unroll
    for (int i = 0; i < 4; i++) { s.q[i] = x[i]; s.v[i] = x[4 + i]; s.warm[i] = 0; }
  }
  __device__ static inline void store(const State& s, double* x) {
#pragma unroll
    for (int i = 0; i < 4; i++) { x[i] = s.q[i]; x[4 + i] = s.v[i]; }
  }

  __device__ static inline void step(const Consts& c, State& s, const double* u, double* sens) {
    if (sens) {
      sens[0] = s.q[0] + c.site_pusher[0]; sens[1] = s.q[1] + c.site_pusher[1]; sens[2] = c.site_pusher[2];
      sens[3] = s.q[2] + c.site_cart[0]; sens[4] = s.q[3] + c.site_cart[1]; sens[5] = c.site_cart[2];
    }
    const double h = c.dt;
    double M[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) M[i][j] = (i == j) ? (i < 2 ? c.mass_pusher : c.mass_cart) : 0.0;
    // collision: two upright discs; normal from geom1 (pusher) to geom2 (cart)
    double J[4][4], D[4], aref[4];
    int nefc = 0;
    double dx = s.q[2] - s.q[0], dy = s.q[3] - s.q[1];
    double dxy = sqrt(dx * dx + dy * dy);
    double dist = dxy - (c.r_pusher + c.r_cart);
    double includemargin = c.margin - c.gap;
    if (dist < c.margin && dxy >= B2_MINVAL && dist < includemargin) {
      double nx = dx / dxy, ny = dy / dxy;
      // mju_makeFrame: second axis from (0,1,0) if |n_y| < 0.5 else (0,0,1); third = n x second.  Only the xy parts
      // of the tangents reach the 4 translational dofs.
      double t1x, t1y, t2x, t2y;
      if (fabs(ny) < 0.5) {
        double yx = -ny * nx, yy = 1 - ny * ny;  // (0,1,0) - n (n.y)
        double yn = sqrt(yx * yx + yy * yy);
        t1x = yx / yn; t1y = yy / yn; t2x = 0; t2y = 0;
      } else { t1x = 0; t1y = 0; t2x = ny; t2y = -nx; }
      double Jn[4] = {-nx, -ny, nx, ny};
      double Jt[2][4] = {{-t1x, -t1y, t1x, t1y}, {-t2x, -t2y, t2x, t2y}};
      double mu = c.mu;
      // rows: (t1,+) (t1,-) (t2,+) (t2,-); R0 from the first row, then the pyramidal adjustment
      double diagA = c.tran + mu * mu * c.tran;
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int sg = 0; sg < 2; sg++) {
          int r = 2 * a + sg;
          double sgn = sg == 0 ? 1.0 : -1.0, vel = 0;
#pragma unroll
          for (int i = 0; i < 4; i++) { J[r][i] = Jn[i] + sgn * mu * Jt[a][i]; vel += J[r][i] * s.v[i]; }
          double R;
          row_reference(c.solref, c.solimp, h, dist, includemargin, vel, diagA, &R, &aref[r]);
          D[r] = R;  // holds R until the cone adjustment below
        }
      double R0 = D[0];
      double R1 = R0 / fmax(B2_MINVAL, c.impratio);
      double mureg = mu * sqrt(R1 / R0);
      double Rpy = 2 * mureg * mureg * R1;
#pragma unroll
      for (int r = 0; r < 4; r++) D[r] = 1 / Rpy;
      nefc = 4;
    } else {
#pragma unroll
      for (int r = 0; r < 4; r++) { D[r] = 0; aref[r] = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) J[r][i] = 0; }
    }
    // smooth forces: damping + position servos on the pusher; gravity does no work on the slides
    double qfs[4], qas[4];
#pragma unroll
    for (int i = 0; i < 4; i++) qfs[i] = -c.damp[i] * s.v[i];
#pragma unroll
    for (int a = 0; a < 2; a++) {
      double uc = u[a];
      if (c.ctrllimited != 0) uc = fmin(fmax(uc, c.ctrl_lo), c.ctrl_hi);
      double f = c.kp * uc - c.kp * s.q[a];
      if (c.forcelimited != 0) f = fmin(fmax(f, c.frc_lo), c.frc_hi);
      qfs[a] += f;
    }
    double L[4][4];
    chol<4>(L, M);
#pragma unroll
    for (int i = 0; i < 4; i++) qas[i] = qfs[i];
    chol_solve<4>(L, qas);
    double qacc[4], qfc[4];
    SolverOpt o{c.meaninertia, c.tolerance, c.ls_tolerance, (int)c.iterations, (int)c.ls_iterations};
    RowSolver<4, 4> sol;
    sol.solve(M, qfs, qas, J, D, aref, nefc, s.warm, o, qacc, qfc);
    double qa[4];
    bool any = false;
#pragma unroll
    for (int i = 0; i < 4; i++) any = any || c.damp[i] > 0;
    if (any) {
      double A[4][4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) A[i][j] = M[i][j];
        A[i][i] += h * c.damp[i];
        qa[i] = qfs[i] + qfc[i];
      }
      chol<4>(L, A);
      chol_solve<4>(L, qa);
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) qa[i] = qacc[i];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) { s.v[i] += h * qa[i]; s.q[i] += h * s.v[i]; s.warm[i] = qacc[i]; }
  }

  // cost params: [w_pusher_proximity, w_pusher_velocity, w_cart_position, pusher_goal_offset, goal_x, goal_y]
  __device__ static inline double cost(const double* p, const State& s, const double* u) {
    double gx = p[4] - s.q[2], gy = p[5] - s.q[3];
    double gn = sqrt(gx * gx + gy * gy);  // no epsilon (cylinder_push.py:76-77)
    double pgx = s.q[2] - p[3] * (gx / gn), pgy = s.q[3] - p[3] * (gy / gn);
    double ex = s.q[0] - pgx, ey = s.q[1] - pgy;
    double prox = 0.5 * (ex * ex + ey * ey);
    double vel = 0.5 * (s.v[0] * s.v[0] + s.v[1] * s.v[1]);
    double goal = 0.5 * (gx * gx + gy * gy);
    return p[0] * prox + p[1] * vel + p[2] * goal;
  }
  __device__ static inline double finish(double sum, int H) { return -sum; }
};

}  // namespace b2
