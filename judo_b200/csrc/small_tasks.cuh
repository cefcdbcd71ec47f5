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
  // sn/cs cache sincos(q[1]) (shared by this step's kinematics and the previous step's cost); the warm start
  // qacc_{t-1} = M_{t-1}^-1 qfrc_smooth_{t-1} of an unconstrained step is kept in factored form (wm01, wq) and only
  // evaluated if a constraint row turns up.
  struct State { double q[2], v[2], warm[2], sn, cs, wm01, wq[2]; bool lazy; int since_exact; };

  __device__ static inline void load(State& s, const double* x) {
    s.q[0] = x[0]; s.q[1] = x[1]; s.v[0] = x[2]; s.v[1] = x[3]; s.warm[0] = s.warm[1] = 0; s.lazy = false;
    s.wm01 = 0; s.wq[0] = s.wq[1] = 0; s.since_exact = 0;
    sincos(s.q[1], &s.sn, &s.cs);
  }
  __device__ static inline void store(const State& s, double* x) { x[0] = s.q[0]; x[1] = s.q[1]; x[2] = s.v[0]; x[3] = s.v[1]; }

  // the two framepos sensors as functions of the positions (sn, cs = sincos(q[1]))
  __device__ static inline void sensors_sc(const Consts& c, double q0, double sn, double cs, double* sens) {
    sens[0] = q0 + c.site_cart[0]; sens[1] = c.site_cart[1]; sens[2] = c.site_cart[2];
    sens[3] = q0 + cs * c.site_pole[0] + sn * c.site_pole[2];
    sens[4] = c.site_pole[1];
    sens[5] = -sn * c.site_pole[0] + cs * c.site_pole[2];
  }
  __device__ static inline void sensors(const Consts& c, const double* q, double* sens) {
    double sn, cs;
    sincos(q[1], &sn, &cs);
    sensors_sc(c, q[0], sn, cs, sens);
  }

  // cart: slide along x; pole: hinge about +y at the cart origin, COM at (0,0,l) in the pole frame.
  __device__ static inline void step(const Consts& c, State& s, const double* u, double* sens) {
    const double sn = s.sn, cs = s.cs;
    if (sens) sensors_sc(c, s.q[0], sn, cs, sens);  // framepos sensors are evaluated in mj_forward, i.e. at the pre-step state
    const double h = c.dt, mp = c.m_pole, l = c.l_pole;
    const double m00 = c.m_cart + mp, m01 = mp * l * cs, m11 = c.iyy_pole + mp * l * l;
    // passive (joint damping), bias (Coriolis + gravity), actuation (position servo with ctrl/force clamps)
    double uc = u[0];
    if (c.ctrllimited != 0) uc = fmin(fmax(uc, c.ctrl_lo), c.ctrl_hi);
    double fa = c.kp * uc - c.kp * s.q[0];
    if (c.forcelimited != 0) fa = fmin(fmax(fa, c.frc_lo), c.frc_hi);
    const double bias0 = -mp * l * sn * s.v[1] * s.v[1];
    const double bias1 = -mp * c.gravity * l * sn;
    const double qfs0 = -c.damp_cart * s.v[0] - bias0 + fa, qfs1 = -c.damp_pole * s.v[1] - bias1;
    // joint-limit rows (mj_instantiateLimit): active when dist < margin
    const double dlo = s.q[0] - c.lim_lo, dhi = c.lim_hi - s.q[0];
    const bool constrained = c.limited != 0 && (dlo < c.lim_margin || dhi < c.lim_margin);
    double qfc0 = 0, qfc1 = 0;
    if (!constrained) {
      // nefc == 0: qacc = qacc_smooth; keep it factored for a possible later warm start
      s.wm01 = m01; s.wq[0] = qfs0; s.wq[1] = qfs1; s.lazy = true;
    } else {
      double M[2][2] = {{m00, m01}, {m01, m11}};
      double J[2][2] = {{0, 0}, {0, 0}}, D[2] = {0, 0}, aref[2] = {0, 0};
      int nefc = 0;
      if (dlo < c.lim_margin) {
        double R;
        row_reference(c.solref, c.solimp, h, dlo, c.lim_margin, s.v[0], c.invweight_cart, &R, &aref[0]);
        J[0][0] = 1; D[0] = 1 / R; nefc = 1;
      }
      if (dhi < c.lim_margin) {
        double R, ar;
        row_reference(c.solref, c.solimp, h, dhi, c.lim_margin, -s.v[0], c.invweight_cart, &R, &ar);
        if (nefc == 0) { J[0][0] = -1; D[0] = 1 / R; aref[0] = ar; } else { J[1][0] = -1; D[1] = 1 / R; aref[1] = ar; }
        nefc++;
      }
      if (s.lazy) {  // materialise the previous step's unconstrained acceleration
        double Mp[2][2] = {{m00, s.wm01}, {s.wm01, m11}}, Lp[2][2];
        chol<2>(Lp, Mp);
        s.warm[0] = s.wq[0]; s.warm[1] = s.wq[1];
        chol_solve<2>(Lp, s.warm);
        s.lazy = false;
      }
      double qfs[2] = {qfs0, qfs1}, qas[2] = {qfs0, qfs1}, L[2][2], qacc[2], qfc[2];
      chol<2>(L, M);
      chol_solve<2>(L, qas);
      SolverOpt o{c.meaninertia, c.tolerance, c.ls_tolerance, (int)c.iterations, (int)c.ls_iterations};
      RowSolver<2, 2> sol;
      sol.solve(M, qfs, qas, J, D, aref, nefc, s.warm, o, qacc, qfc);
      qfc0 = qfc[0]; qfc1 = qfc[1];
      s.warm[0] = qacc[0]; s.warm[1] = qacc[1];
    }
    // mj_Euler: (M + h diag(damping)) qacc = qfrc_smooth + qfrc_constraint — closed-form 2x2 solve — then advance
    const double a00 = m00 + h * c.damp_cart, a11 = m11 + h * c.damp_pole;
    const double r0 = qfs0 + qfc0, r1 = qfs1 + qfc1;
    const double idet = 1.0 / (a00 * a11 - m01 * m01);
    const double qa0 = (a11 * r0 - m01 * r1) * idet, qa1 = (a00 * r1 - m01 * r0) * idet;
    s.v[0] += h * qa0; s.v[1] += h * qa1;
    s.q[0] += h * s.v[0]; s.q[1] += h * s.v[1];
    sincos(s.q[1], &s.sn, &s.cs);
    s.since_exact = 0;
  }

  // cost params: [w_vertical, w_centered, w_velocity, w_control, p_vertical, p_centered]
  __device__ static inline double cost(const double* p, const State& s, const double* u) {
    double cz = s.cs - 1;
    double vertical = sqrt(cz * cz + p[4] * p[4]) - p[4];
    double centered = sqrt(s.q[0] * s.q[0] + p[5] * p[5]) - p[5];
    double vel = 0.5 * (s.v[0] * s.v[0] + s.v[1] * s.v[1]);
    double ctl = 0.5 * u[0] * u[0];
    return p[0] * vertical + p[1] * centered + p[2] * vel + p[3] * ctl;
  }

  // the general path
  __device__ static inline double step_cost_general(const Consts& c, State& s, const double* u, const double* p) {
    step(c, s, u, nullptr);  // limit rows through the Newton solver, exact sincos at the end
    return cost(p, s, u);
  }

  // One step + its cost for the fused kernel, arranged as ONE straight-line block: a lone warp per SM sub-partition issues in order, so
  // what the step costs is the length of its dependent fp64 chain and how well the compiler can interleave the independent chains
  // (dynamics, cost) — every branch in between is a scheduling barrier.  The unconstrained update (no joint-limit row: nefc == 0,
  // qacc = qacc_smooth) and its cost are computed unconditionally; the rare cases redo the step through step() afterwards:
  // a limit row is active, the pole turned by >= 0.5 rad in one step, or 31 incremental rotations have been chained.
  // (sin, cos) of the advanced pole angle come from rotating the cached pair by the increment dq = h * v1 — Taylor series to the first
  // term below 1e-17 for |dq| < 0.5 (sin through dq^15, cos through dq^14), evaluated pairwise (Estrin) to keep the chain short —
  // instead of a full sincos (argument reduction, two polynomials, quadrant selection: the longest link of the old chain).
  __device__ static inline double step_cost(const Consts& c, State& s, const double* u, const double* p) {
    const double sn = s.sn, cs = s.cs;
    const double h = c.dt, mp = c.m_pole, l = c.l_pole;
    const double m00 = c.m_cart + mp, m01 = mp * l * cs, m11 = c.iyy_pole + mp * l * l;
    double uc = u[0];
    if (c.ctrllimited != 0) uc = fmin(fmax(uc, c.ctrl_lo), c.ctrl_hi);
    double fa = c.kp * uc - c.kp * s.q[0];
    if (c.forcelimited != 0) fa = fmin(fmax(fa, c.frc_lo), c.frc_hi);
    const double bias0 = -mp * l * sn * s.v[1] * s.v[1];
    const double bias1 = -mp * c.gravity * l * sn;
    const double qfs0 = -c.damp_cart * s.v[0] - bias0 + fa, qfs1 = -c.damp_pole * s.v[1] - bias1;
    const double dlo = s.q[0] - c.lim_lo, dhi = c.lim_hi - s.q[0];
    const bool constrained = c.limited != 0 && (dlo < c.lim_margin || dhi < c.lim_margin);
    const double a00 = m00 + h * c.damp_cart, a11 = m11 + h * c.damp_pole;
    const double idet = fast_rcp_pos(a00 * a11 - m01 * m01);
    const double qa0 = (a11 * qfs0 - m01 * qfs1) * idet, qa1 = (a00 * qfs1 - m01 * qfs0) * idet;
    const double v0 = s.v[0] + h * qa0, v1 = s.v[1] + h * qa1;
    const double q0 = s.q[0] + h * v0, dq = h * v1, q1 = s.q[1] + dq;
    const double z = dq * dq, z2 = z * z, z4 = z2 * z2;
    // sin(dq) = dq (1 + z S(z)),  S = s3 + s5 z + s7 z^2 + ... + s15 z^6;  cos(dq) = 1 + z C(z),  C = c2 + c4 z + ... + c14 z^6
    const double S01 = fma(z, 1.0 / 120.0, -1.0 / 6.0), S23 = fma(z, 1.0 / 362880.0, -1.0 / 5040.0);
    const double S45 = fma(z, 1.0 / 6227020800.0, -1.0 / 39916800.0), S6 = -1.0 / 1307674368000.0;
    const double Sp = fma(z4, fma(z2, S6, S45), fma(z2, S23, S01));
    const double C01 = fma(z, 1.0 / 24.0, -0.5), C23 = fma(z, 1.0 / 40320.0, -1.0 / 720.0);
    const double C45 = fma(z, 1.0 / 479001600.0, -1.0 / 3628800.0), C6 = -1.0 / 87178291200.0;
    const double Cp = fma(z4, fma(z2, C6, C45), fma(z2, C23, C01));
    const double sd = fma(dq, z * Sp, dq), cd = fma(z, Cp, 1.0);
    const double nsn = fma(sn, cd, cs * sd), ncs = fma(cs, cd, -sn * sd);
    // cost at the new state (cartpole.py:66-71)
    const double cz = ncs - 1;
    const double vertical = fast_sqrt_nonneg(cz * cz + p[4] * p[4]) - p[4];
    const double centered = fast_sqrt_nonneg(q0 * q0 + p[5] * p[5]) - p[5];
    const double ct = p[0] * vertical + p[1] * centered + p[2] * (0.5 * (v0 * v0 + v1 * v1)) + p[3] * (0.5 * u[0] * u[0]);
    if (constrained || !(fabs(dq) < 0.5) || s.since_exact >= 31) return step_cost_general(c, s, u, p);
    s.wm01 = m01; s.wq[0] = qfs0; s.wq[1] = qfs1; s.lazy = true;  // (what step() keeps for a later warm start)
    s.v[0] = v0; s.v[1] = v1; s.q[0] = q0; s.q[1] = q1; s.sn = nsn; s.cs = ncs; s.since_exact++;
    return ct;
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
  // im / ia: reciprocals of the (diagonal) mass matrix and of M + h*damping, computed once per rollout (ready < 0: not yet)
  struct State { double q[4], v[4], warm[4], im[4], ia[4]; bool ready; };

  __device__ static inline void load(State& s, const double* x) {
#pragma unroll
    for (int i = 0; i < 4; i++) { s.q[i] = x[i]; s.v[i] = x[4 + i]; s.warm[i] = 0; }
    s.ready = false;
  }
  __device__ static inline void store(const State& s, double* x) {
#pragma unroll
    for (int i = 0; i < 4; i++) { x[i] = s.q[i]; x[4 + i] = s.v[i]; }
  }

  __device__ static inline void sensors(const Consts& c, const double* q, double* sens) {
    sens[0] = q[0] + c.site_pusher[0]; sens[1] = q[1] + c.site_pusher[1]; sens[2] = c.site_pusher[2];
    sens[3] = q[2] + c.site_cart[0]; sens[4] = q[3] + c.site_cart[1]; sens[5] = c.site_cart[2];
  }

  __device__ static inline void step(const Consts& c, State& s, const double* u, double* sens) {
    if (sens) sensors(c, s.q, sens);
    const double h = c.dt;
    if (!s.ready) {
#pragma unroll
      for (int i = 0; i < 4; i++) { const double mi = i < 2 ? c.mass_pusher : c.mass_cart; s.im[i] = 1.0 / mi; s.ia[i] = 1.0 / (mi + h * c.damp[i]); }
      s.ready = true;
    }
    // smooth forces: damping + position servos on the pusher; gravity does no work on the slides; M is diagonal
    double qfs[4], mass[4] = {c.mass_pusher, c.mass_pusher, c.mass_cart, c.mass_cart};
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
    double qas[4], qfc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; i++) qas[i] = qfs[i] * s.im[i];
    // collision: two upright discs; normal from geom1 (pusher) to geom2 (cart).  Squared test first: no sqrt off-contact
    const double dx = s.q[2] - s.q[0], dy = s.q[3] - s.q[1];
    const double includemargin = c.margin - c.gap, rsum = c.r_pusher + c.r_cart;
    const double reach = rsum + fmin(c.margin, includemargin);
    const double d2 = dx * dx + dy * dy;
    bool contact = false;
    if (reach > 0 && d2 < reach * reach) {
      const double dxy = sqrt(d2), dist = dxy - rsum;
      if (dist < c.margin && dxy >= B2_MINVAL && dist < includemargin) {
        contact = true;
        double M[4][4], J[4][4], D[4], aref[4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) M[i][j] = (i == j) ? mass[i] : 0.0;
        const double nx = dx / dxy, ny = dy / dxy;
        // mju_makeFrame: second axis from (0,1,0) if |n_y| < 0.5 else (0,0,1); third = n x second.  Only the xy parts
        // of the tangents reach the 4 translational dofs.
        double t1x, t1y, t2x, t2y;
        if (fabs(ny) < 0.5) {
          double yx = -ny * nx, yy = 1 - ny * ny;  // (0,1,0) - n (n.y)
          double yn = sqrt(yx * yx + yy * yy);
          t1x = yx / yn; t1y = yy / yn; t2x = 0; t2y = 0;
        } else { t1x = 0; t1y = 0; t2x = ny; t2y = -nx; }
        const double Jn[4] = {-nx, -ny, nx, ny};
        const double Jt[2][4] = {{-t1x, -t1y, t1x, t1y}, {-t2x, -t2y, t2x, t2y}};
        const double mu = c.mu;
        // rows: (t1,+) (t1,-) (t2,+) (t2,-); R0 from the first row, then the pyramidal adjustment
        const double diagA = c.tran + mu * mu * c.tran;
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
          for (int sg = 0; sg < 2; sg++) {
            const int r = 2 * a + sg;
            const double sgn = sg == 0 ? 1.0 : -1.0;
            double vel = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) { J[r][i] = Jn[i] + sgn * mu * Jt[a][i]; vel += J[r][i] * s.v[i]; }
            double R;
            row_reference(c.solref, c.solimp, h, dist, includemargin, vel, diagA, &R, &aref[r]);
            D[r] = R;  // holds R until the cone adjustment below
          }
        const double R0 = D[0];
        const double R1 = R0 / fmax(B2_MINVAL, c.impratio);
        const double mureg = mu * sqrt(R1 / R0);
        const double Rpy = 2 * mureg * mureg * R0;  // = 2 mu^2 R0 / impratio
#pragma unroll
        for (int r = 0; r < 4; r++) D[r] = 1 / Rpy;
        double qacc[4];
        SolverOpt o{c.meaninertia, c.tolerance, c.ls_tolerance, (int)c.iterations, (int)c.ls_iterations};
        RowSolver<4, 4> sol;
        sol.solve(M, qfs, qas, J, D, aref, 4, s.warm, o, qacc, qfc);
#pragma unroll
        for (int i = 0; i < 4; i++) s.warm[i] = qacc[i];
      }
    }
    if (!contact) {
#pragma unroll
      for (int i = 0; i < 4; i++) s.warm[i] = qas[i];  // qacc = qacc_smooth
    }
    // mj_Euler with implicit joint damping (diagonal system), then semi-implicit advance
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const double qa = (qfs[i] + qfc[i]) * s.ia[i];
      s.v[i] += h * qa;
      s.q[i] += h * s.v[i];
    }
  }

  // cost params: [w_pusher_proximity, w_pusher_velocity, w_cart_position, pusher_goal_offset, goal_x, goal_y]
  __device__ static inline double cost(const double* p, const State& s, const double* u) {
    double gx = p[4] - s.q[2], gy = p[5] - s.q[3];
    const double ign = 1.0 / sqrt(gx * gx + gy * gy);  // no epsilon (cylinder_push.py:76-77): inf/nan at the goal, as numpy
    double pgx = s.q[2] - p[3] * (gx * ign), pgy = s.q[3] - p[3] * (gy * ign);
    double ex = s.q[0] - pgx, ey = s.q[1] - pgy;
    double prox = 0.5 * (ex * ex + ey * ey);
    double vel = 0.5 * (s.v[0] * s.v[0] + s.v[1] * s.v[1]);
    double goal = 0.5 * (gx * gx + gy * gy);
    return p[0] * prox + p[1] * vel + p[2] * goal;
  }
  __device__ static inline double step_cost(const Consts& c, State& s, const double* u, const double* p) {
    step(c, s, u, nullptr);
    return cost(p, s, u);
  }
  __device__ static inline double finish(double sum, int H) { return -sum; }
};

}  // namespace b2
