/* mjc.c — ORACLE (test infrastructure, not product). See mjc.h for scope and the parity statement.
 *
 * Each function names the MuJoCo 3.5.0 pipeline stage it restates (names as in MuJoCo's
 * documented mj_step = mj_forward + integrator decomposition) and, where
 * one exists, the reference call site that reaches it.
 */
#include "mjc.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MINVAL 1e-15
#define MINIMP 0.0001
#define MAXIMP 0.9999
#define MINMU 1e-5

static int g_debug = 0;
static long g_ls_evals = 0, g_newton_iters = 0, g_steps = 0;
void mjc_get_counters(long* out) { out[0] = g_ls_evals; out[1] = g_newton_iters; out[2] = g_steps; g_ls_evals = g_newton_iters = g_steps = 0; }
void mjc_set_debug(int v) { g_debug = v; }

/* constraint row types, in MuJoCo's row order */
enum { CT_EQUALITY = 0, CT_FRICTION_DOF = 1, CT_LIMIT_JOINT = 3, CT_CONTACT_PYRAMIDAL = 6, CT_CONTACT_ELLIPTIC = 7 };
/* constraint states */
enum { ST_SATISFIED = 0, ST_QUADRATIC = 1, ST_LINEARNEG = 2, ST_LINEARPOS = 3, ST_CONE = 4 };

struct mjcModel {
  int nq, nv, nu, nbody, njnt, ngeom, nsite, nsensor, nsensordata, npair, neq;
  int integrator, cone, contact_disabled, iterations, ls_iterations;
  double timestep, gravity[3], impratio, tolerance, ls_tolerance, meaninertia;
  double qpos0[MJC_MAXNQ];
  int body_parent[MJC_MAXBODY], body_jntadr[MJC_MAXBODY], body_jntnum[MJC_MAXBODY];
  double body_pos[MJC_MAXBODY][3], body_quat[MJC_MAXBODY][4], body_ipos[MJC_MAXBODY][3], body_iquat[MJC_MAXBODY][4];
  double body_mass[MJC_MAXBODY], body_inertia[MJC_MAXBODY][3], body_invweight0[MJC_MAXBODY][2];
  int jnt_type[MJC_MAXJNT], jnt_body[MJC_MAXJNT], jnt_qposadr[MJC_MAXJNT], jnt_dofadr[MJC_MAXJNT], jnt_limited[MJC_MAXJNT];
  double jnt_pos[MJC_MAXJNT][3], jnt_axis[MJC_MAXJNT][3], jnt_range[MJC_MAXJNT][2], jnt_margin[MJC_MAXJNT];
  double jnt_solref[MJC_MAXJNT][2], jnt_solimp[MJC_MAXJNT][5];
  int jnt_actfrclimited[MJC_MAXJNT];
  double jnt_actfrcrange[MJC_MAXJNT][2];
  int eq_j1[MJC_MAXEQ], eq_j2[MJC_MAXEQ];
  double eq_polycoef[MJC_MAXEQ][5], eq_solref[MJC_MAXEQ][2], eq_solimp[MJC_MAXEQ][5];
  int dof_jnt[MJC_MAXNV];
  double dof_damping[MJC_MAXNV], dof_frictionloss[MJC_MAXNV], dof_armature[MJC_MAXNV], dof_invweight0[MJC_MAXNV];
  double dof_solref[MJC_MAXNV][2], dof_solimp[MJC_MAXNV][5];
  int geom_type[MJC_MAXGEOM], geom_body[MJC_MAXGEOM], geom_condim[MJC_MAXGEOM], geom_priority[MJC_MAXGEOM];
  double geom_size[MJC_MAXGEOM][3], geom_pos[MJC_MAXGEOM][3], geom_quat[MJC_MAXGEOM][4], geom_friction[MJC_MAXGEOM][3];
  double geom_solref[MJC_MAXGEOM][2], geom_solimp[MJC_MAXGEOM][5], geom_margin[MJC_MAXGEOM], geom_gap[MJC_MAXGEOM], geom_solmix[MJC_MAXGEOM];
  int pair_g1[MJC_MAXPAIR], pair_g2[MJC_MAXPAIR];
  int site_body[MJC_MAXSITE];
  double site_pos[MJC_MAXSITE][3];
  int act_dof[MJC_MAXNU], act_ctrllimited[MJC_MAXNU], act_forcelimited[MJC_MAXNU];
  double act_gear[MJC_MAXNU], act_kp[MJC_MAXNU], act_kv[MJC_MAXNU], act_ctrlrange[MJC_MAXNU][2], act_forcerange[MJC_MAXNU][2];
  int sens_type[MJC_MAXSENSOR], sens_obj[MJC_MAXSENSOR], sens_obj2[MJC_MAXSENSOR], sens_adr[MJC_MAXSENSOR];
  double sens_cutoff[MJC_MAXSENSOR];
};

typedef struct {
  double dist, pos[3], frame[9], includemargin, friction[5], solref[2], solimp[5], mu;
  int dim, geom1, geom2, efc_address;
} mjcContact;

typedef struct {
  double qpos[MJC_MAXNQ], qvel[MJC_MAXNV], ctrl[MJC_MAXNU], qacc[MJC_MAXNV], qacc_warmstart[MJC_MAXNV];
  double xpos[MJC_MAXBODY][3], xquat[MJC_MAXBODY][4], xmat[MJC_MAXBODY][9], xipos[MJC_MAXBODY][3], ximat[MJC_MAXBODY][9];
  double xanchor[MJC_MAXJNT][3], xaxis[MJC_MAXJNT][3];
  double geom_xpos[MJC_MAXGEOM][3], geom_xmat[MJC_MAXGEOM][9], site_xpos[MJC_MAXSITE][3];
  double M[MJC_MAXNV * MJC_MAXNV], L[MJC_MAXNV * MJC_MAXNV];
  double qfrc_bias[MJC_MAXNV], qfrc_passive[MJC_MAXNV], qfrc_actuator[MJC_MAXNV], qfrc_smooth[MJC_MAXNV];
  double qacc_smooth[MJC_MAXNV], qfrc_constraint[MJC_MAXNV], actuator_force[MJC_MAXNU];
  double sensordata[MJC_MAXSENSOR * 3];
  int ncon, nefc, solver_iter;
  double solver_grad, solver_cost;
  mjcContact contact[MJC_MAXCON];
  double efc_J[MJC_MAXEFC][MJC_MAXNV];
  double efc_pos[MJC_MAXEFC], efc_margin[MJC_MAXEFC], efc_frictionloss[MJC_MAXEFC], efc_diagApprox[MJC_MAXEFC];
  double efc_D[MJC_MAXEFC], efc_R[MJC_MAXEFC], efc_aref[MJC_MAXEFC], efc_vel[MJC_MAXEFC], efc_force[MJC_MAXEFC];
  int efc_type[MJC_MAXEFC], efc_id[MJC_MAXEFC], efc_state[MJC_MAXEFC];
} mjcData;

/* ------------------------------------------------------------------ small vector helpers */
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline double norm3(const double* a) { return sqrt(dot3(a, a)); }
static inline double normalize3(double* a) {
  double n = norm3(a);
  if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; return 0; }
  a[0] /= n; a[1] /= n; a[2] /= n;
  return n;
}
static void quat_mul(double* r, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void quat_normalize(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static void quat2mat(double* m, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
static inline void mat_vec(double* r, const double* m, const double* v) { /* r = m v (3x3 row-major) */
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
         z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void matT_vec(double* r, const double* m, const double* v) { /* r = m^T v */
  double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2],
         z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void mat_mul(double* r, const double* a, const double* b) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  memcpy(r, t, sizeof t);
}

/* ------------------------------------------------------------------ model blob reader */
typedef struct { const int* ib; const double* db; int ip, dp, ni, nd, err; } Reader;
static int RI(Reader* r) { if (r->ip >= r->ni) { r->err = 1; return 0; } return r->ib[r->ip++]; }
static double RD(Reader* r) { if (r->dp >= r->nd) { r->err = 1; return 0; } return r->db[r->dp++]; }
static void RDV(Reader* r, double* v, int n) { for (int i = 0; i < n; i++) v[i] = RD(r); }

mjcModel* mjc_model_create(const int* ib, int ni, const double* db, int nd) {
  Reader R = {ib, db, 0, 0, ni, nd, 0};
  mjcModel* m = (mjcModel*)calloc(1, sizeof(mjcModel));
  if (!m) return NULL;
  m->nq = RI(&R); m->nv = RI(&R); m->nu = RI(&R); m->nbody = RI(&R); m->njnt = RI(&R); m->ngeom = RI(&R);
  m->nsite = RI(&R); m->nsensor = RI(&R); m->nsensordata = RI(&R); m->npair = RI(&R); m->neq = RI(&R);
  m->integrator = RI(&R); m->cone = RI(&R); m->contact_disabled = RI(&R); m->iterations = RI(&R); m->ls_iterations = RI(&R);
  if (m->nq > MJC_MAXNQ || m->nv > MJC_MAXNV || m->nu > MJC_MAXNU || m->nbody > MJC_MAXBODY || m->njnt > MJC_MAXJNT ||
      m->ngeom > MJC_MAXGEOM || m->nsite > MJC_MAXSITE || m->nsensor > MJC_MAXSENSOR || m->npair > MJC_MAXPAIR || m->neq > MJC_MAXEQ) {
    free(m);
    return NULL;
  }
  m->timestep = RD(&R); RDV(&R, m->gravity, 3); m->impratio = RD(&R); m->tolerance = RD(&R);
  m->ls_tolerance = RD(&R); m->meaninertia = RD(&R);
  RDV(&R, m->qpos0, m->nq);
  for (int i = 0; i < m->nbody; i++) {
    m->body_parent[i] = RI(&R); m->body_jntadr[i] = RI(&R); m->body_jntnum[i] = RI(&R);
    RDV(&R, m->body_pos[i], 3); RDV(&R, m->body_quat[i], 4); RDV(&R, m->body_ipos[i], 3); RDV(&R, m->body_iquat[i], 4);
    m->body_mass[i] = RD(&R); RDV(&R, m->body_inertia[i], 3); RDV(&R, m->body_invweight0[i], 2);
  }
  for (int i = 0; i < m->njnt; i++) {
    m->jnt_type[i] = RI(&R); m->jnt_body[i] = RI(&R); m->jnt_qposadr[i] = RI(&R); m->jnt_dofadr[i] = RI(&R);
    m->jnt_limited[i] = RI(&R); m->jnt_actfrclimited[i] = RI(&R);
    RDV(&R, m->jnt_pos[i], 3); RDV(&R, m->jnt_axis[i], 3); RDV(&R, m->jnt_range[i], 2); m->jnt_margin[i] = RD(&R);
    RDV(&R, m->jnt_solref[i], 2); RDV(&R, m->jnt_solimp[i], 5); RDV(&R, m->jnt_actfrcrange[i], 2);
  }
  for (int i = 0; i < m->nv; i++) {
    m->dof_jnt[i] = RI(&R);
    m->dof_damping[i] = RD(&R); m->dof_frictionloss[i] = RD(&R); m->dof_armature[i] = RD(&R); m->dof_invweight0[i] = RD(&R);
    RDV(&R, m->dof_solref[i], 2); RDV(&R, m->dof_solimp[i], 5);
  }
  for (int i = 0; i < m->ngeom; i++) {
    m->geom_type[i] = RI(&R); m->geom_body[i] = RI(&R); m->geom_condim[i] = RI(&R); m->geom_priority[i] = RI(&R);
    RDV(&R, m->geom_size[i], 3); RDV(&R, m->geom_pos[i], 3); RDV(&R, m->geom_quat[i], 4); RDV(&R, m->geom_friction[i], 3);
    RDV(&R, m->geom_solref[i], 2); RDV(&R, m->geom_solimp[i], 5);
    m->geom_margin[i] = RD(&R); m->geom_gap[i] = RD(&R); m->geom_solmix[i] = RD(&R);
  }
  for (int i = 0; i < m->npair; i++) { m->pair_g1[i] = RI(&R); m->pair_g2[i] = RI(&R); }
  for (int i = 0; i < m->nsite; i++) { m->site_body[i] = RI(&R); RDV(&R, m->site_pos[i], 3); }
  for (int i = 0; i < m->nu; i++) {
    m->act_dof[i] = RI(&R); m->act_ctrllimited[i] = RI(&R); m->act_forcelimited[i] = RI(&R);
    m->act_gear[i] = RD(&R); m->act_kp[i] = RD(&R); m->act_kv[i] = RD(&R);
    RDV(&R, m->act_ctrlrange[i], 2); RDV(&R, m->act_forcerange[i], 2);
  }
  for (int i = 0; i < m->nsensor; i++) {
    m->sens_type[i] = RI(&R); m->sens_obj[i] = RI(&R); m->sens_obj2[i] = RI(&R); m->sens_adr[i] = RI(&R); m->sens_cutoff[i] = RD(&R);
  }
  for (int i = 0; i < m->neq; i++) {
    m->eq_j1[i] = RI(&R); m->eq_j2[i] = RI(&R);
    RDV(&R, m->eq_polycoef[i], 5); RDV(&R, m->eq_solref[i], 2); RDV(&R, m->eq_solimp[i], 5);
  }
  if (R.err || R.ip != ni || R.dp != nd) { free(m); return NULL; }
  return m;
}
void mjc_model_free(mjcModel* m) { free(m); }
int mjc_nq(const mjcModel* m) { return m->nq; }
int mjc_nv(const mjcModel* m) { return m->nv; }
int mjc_nu(const mjcModel* m) { return m->nu; }
int mjc_nsensordata(const mjcModel* m) { return m->nsensordata; }

/* ------------------------------------------------------------------ mj_kinematics + mj_comPos */
static void kinematics(const mjcModel* m, mjcData* d) {
  memset(d->xpos[0], 0, sizeof d->xpos[0]);
  d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
  quat2mat(d->xmat[0], d->xquat[0]);
  for (int i = 1; i < m->nbody; i++) {
    int p = m->body_parent[i];
    double pos[3], quat[4], t[3];
    mat_vec(t, d->xmat[p], m->body_pos[i]);
    for (int k = 0; k < 3; k++) pos[k] = d->xpos[p][k] + t[k];
    quat_mul(quat, d->xquat[p], m->body_quat[i]);
    for (int jj = 0; jj < m->body_jntnum[i]; jj++) {
      int j = m->body_jntadr[i] + jj, qa = m->jnt_qposadr[j];
      if (m->jnt_type[j] == MJC_JNT_FREE) {
        quat_normalize(d->qpos + qa + 3); /* mj_kinematics normalises the free-joint quaternion in place */
        for (int k = 0; k < 3; k++) pos[k] = d->qpos[qa + k];
        for (int k = 0; k < 4; k++) quat[k] = d->qpos[qa + 3 + k];
        for (int k = 0; k < 3; k++) { d->xanchor[j][k] = pos[k]; d->xaxis[j][k] = (k == 2); }
        continue;
      }
      double mat[9], axis[3], off[3];
      quat2mat(mat, quat);
      mat_vec(axis, mat, m->jnt_axis[j]);
      mat_vec(off, mat, m->jnt_pos[j]);
      double q = d->qpos[qa] - m->qpos0[qa];
      if (m->jnt_type[j] == MJC_JNT_SLIDE) {
        for (int k = 0; k < 3; k++) { pos[k] += axis[k] * q; d->xanchor[j][k] = pos[k] + off[k]; d->xaxis[j][k] = axis[k]; }
      } else { /* hinge: rotate about the local axis, keep the anchor fixed */
        double anchor[3], dq[4], nq[4], s = sin(0.5 * q);
        for (int k = 0; k < 3; k++) anchor[k] = pos[k] + off[k];
        dq[0] = cos(0.5 * q); dq[1] = s * m->jnt_axis[j][0]; dq[2] = s * m->jnt_axis[j][1]; dq[3] = s * m->jnt_axis[j][2];
        quat_mul(nq, quat, dq);
        memcpy(quat, nq, sizeof nq);
        quat2mat(mat, quat);
        mat_vec(off, mat, m->jnt_pos[j]);
        for (int k = 0; k < 3; k++) { pos[k] = anchor[k] - off[k]; d->xanchor[j][k] = anchor[k]; d->xaxis[j][k] = axis[k]; }
      }
    }
    quat_normalize(quat);
    memcpy(d->xpos[i], pos, sizeof pos);
    memcpy(d->xquat[i], quat, sizeof quat);
    quat2mat(d->xmat[i], quat);
  }
  for (int i = 0; i < m->nbody; i++) {
    double t[3], im[9];
    mat_vec(t, d->xmat[i], m->body_ipos[i]);
    for (int k = 0; k < 3; k++) d->xipos[i][k] = d->xpos[i][k] + t[k];
    quat2mat(im, m->body_iquat[i]);
    mat_mul(d->ximat[i], d->xmat[i], im);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_body[g];
    double t[3], gm[9];
    mat_vec(t, d->xmat[b], m->geom_pos[g]);
    for (int k = 0; k < 3; k++) d->geom_xpos[g][k] = d->xpos[b][k] + t[k];
    quat2mat(gm, m->geom_quat[g]);
    mat_mul(d->geom_xmat[g], d->xmat[b], gm);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_body[s];
    double t[3];
    mat_vec(t, d->xmat[b], m->site_pos[s]);
    for (int k = 0; k < 3; k++) d->site_xpos[s][k] = d->xpos[b][k] + t[k];
  }
}

/* mj_jac: translational / rotational Jacobian (3 x nv, row-major) of a world point attached to a body. */
static void jac_point(const mjcModel* m, const mjcData* d, int body, const double* point, double* jp, double* jr) {
  int nv = m->nv;
  memset(jp, 0, sizeof(double) * 3 * nv);
  if (jr) memset(jr, 0, sizeof(double) * 3 * nv);
  for (int b = body; b != 0; b = m->body_parent[b]) {
    for (int jj = 0; jj < m->body_jntnum[b]; jj++) {
      int j = m->body_jntadr[b] + jj, da = m->jnt_dofadr[j];
      if (m->jnt_type[j] == MJC_JNT_FREE) {
        double r[3];
        for (int k = 0; k < 3; k++) { jp[k * nv + da + k] = 1; r[k] = point[k] - d->xpos[b][k]; }
        for (int a = 0; a < 3; a++) { /* rotational dofs are about the body-frame axes */
          double ax[3] = {d->xmat[b][a], d->xmat[b][3 + a], d->xmat[b][6 + a]}, c[3];
          cross3(c, ax, r);
          for (int k = 0; k < 3; k++) { jp[k * nv + da + 3 + a] = c[k]; if (jr) jr[k * nv + da + 3 + a] = ax[k]; }
        }
      } else if (m->jnt_type[j] == MJC_JNT_SLIDE) {
        for (int k = 0; k < 3; k++) jp[k * nv + da] = d->xaxis[j][k];
      } else {
        double r[3], c[3];
        for (int k = 0; k < 3; k++) r[k] = point[k] - d->xanchor[j][k];
        cross3(c, d->xaxis[j], r);
        for (int k = 0; k < 3; k++) { jp[k * nv + da] = c[k]; if (jr) jr[k * nv + da] = d->xaxis[j][k]; }
      }
    }
  }
}

/* mj_crb (joint-space inertia; here as sum_b Jp^T m Jp + Jr^T I Jr, mathematically identical to the
 * composite-rigid-body recursion) + armature, then mj_factorM as a dense Cholesky. */
static int cholesky(double* L, const double* A, int n) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      double s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
      if (i == j) { if (s < MINVAL) s = MINVAL; L[i * n + i] = sqrt(s); }
      else L[i * n + j] = s / L[j * n + j];
    }
  return 0;
}
static void chol_solve(const double* L, double* x, int n) { /* in place: x = A^-1 x */
  for (int i = 0; i < n; i++) { double s = x[i]; for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k]; x[i] = s / L[i * n + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = x[i]; for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k]; x[i] = s / L[i * n + i]; }
}
static void mass_matrix(const mjcModel* m, mjcData* d) {
  int nv = m->nv;
  double jp[3 * MJC_MAXNV], jr[3 * MJC_MAXNV];
  memset(d->M, 0, sizeof(double) * nv * nv);
  for (int b = 1; b < m->nbody; b++) {
    if (m->body_mass[b] <= 0) continue;
    int moving = 0;
    for (int p = b; p != 0; p = m->body_parent[p]) if (m->body_jntnum[p] > 0) { moving = 1; break; }
    if (!moving) continue;
    jac_point(m, d, b, d->xipos[b], jp, jr);
    /* Iw = ximat diag(I) ximat^T */
    double Iw[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += d->ximat[b][3 * r + k] * m->body_inertia[b][k] * d->ximat[b][3 * c + k];
        Iw[3 * r + c] = s;
      }
    for (int i = 0; i < nv; i++) {
      double Ij[3];
      double col[3] = {jr[i], jr[nv + i], jr[2 * nv + i]};
      mat_vec(Ij, Iw, col);
      for (int j = 0; j <= i; j++) {
        double s = m->body_mass[b] * (jp[i] * jp[j] + jp[nv + i] * jp[nv + j] + jp[2 * nv + i] * jp[2 * nv + j]);
        s += Ij[0] * jr[j] + Ij[1] * jr[nv + j] + Ij[2] * jr[2 * nv + j];
        d->M[i * nv + j] += s;
      }
    }
  }
  for (int i = 0; i < nv; i++) {
    d->M[i * nv + i] += m->dof_armature[i];
    for (int j = 0; j < i; j++) d->M[j * nv + i] = d->M[i * nv + j];
  }
  cholesky(d->L, d->M, nv);
}

/* mj_rne with zero acceleration: qfrc_bias = C(q,v) v + g(q). Classical Newton-Euler in world coordinates;
 * gravity enters as the fictitious base acceleration -g. */
static void rne_bias(const mjcModel* m, mjcData* d) {
  int nb = m->nbody;
  double w[MJC_MAXBODY][3], al[MJC_MAXBODY][3], vo[MJC_MAXBODY][3], ao[MJC_MAXBODY][3]; /* at body frame origin */
  double F[MJC_MAXBODY][3], N0[MJC_MAXBODY][3];                                           /* subtree wrench about world origin */
  memset(w, 0, sizeof w); memset(al, 0, sizeof al); memset(vo, 0, sizeof vo);
  for (int k = 0; k < 3; k++) ao[0][k] = -m->gravity[k];
  for (int i = 1; i < nb; i++) {
    int p = m->body_parent[i];
    double W[3], A[3], P[3], V[3], Ac[3], r[3], t[3], t2[3];
    memcpy(W, w[p], sizeof W); memcpy(A, al[p], sizeof A); memcpy(P, d->xpos[p], sizeof P);
    memcpy(V, vo[p], sizeof V); memcpy(Ac, ao[p], sizeof Ac);
    for (int jj = 0; jj < m->body_jntnum[i]; jj++) {
      int j = m->body_jntadr[i] + jj, da = m->jnt_dofadr[j];
      if (m->jnt_type[j] == MJC_JNT_FREE) {
        /* linear velocity is the world-frame velocity of the body origin; angular velocity is body-local */
        for (int k = 0; k < 3; k++) { V[k] = d->qvel[da + k]; P[k] = d->xpos[i][k]; }
        mat_vec(W, d->xmat[i], d->qvel + da + 3);
        /* qacc = 0: no change of world linear velocity; local angular velocity constant -> alpha = w x w = 0 */
        for (int k = 0; k < 3; k++) A[k] = 0;
      } else if (m->jnt_type[j] == MJC_JNT_SLIDE) {
        double qd = d->qvel[da], u[3] = {d->xaxis[j][0] * qd, d->xaxis[j][1] * qd, d->xaxis[j][2] * qd};
        cross3(t, W, u);
        for (int k = 0; k < 3; k++) { V[k] += u[k]; Ac[k] += 2 * t[k]; }
      } else {
        /* move the reference point to the hinge anchor, then add the relative rotation */
        for (int k = 0; k < 3; k++) r[k] = d->xanchor[j][k] - P[k];
        cross3(t, W, r);
        for (int k = 0; k < 3; k++) V[k] += t[k];
        cross3(t2, W, t);
        cross3(t, A, r);
        for (int k = 0; k < 3; k++) { Ac[k] += t[k] + t2[k]; P[k] = d->xanchor[j][k]; }
        double qd = d->qvel[da], u[3] = {d->xaxis[j][0] * qd, d->xaxis[j][1] * qd, d->xaxis[j][2] * qd};
        cross3(t, W, u);
        for (int k = 0; k < 3; k++) { A[k] += t[k]; W[k] += u[k]; }
      }
    }
    /* move reference point to the body origin */
    for (int k = 0; k < 3; k++) r[k] = d->xpos[i][k] - P[k];
    cross3(t, W, r);
    cross3(t2, W, t);
    for (int k = 0; k < 3; k++) V[k] += t[k];
    cross3(t, A, r);
    for (int k = 0; k < 3; k++) Ac[k] += t[k] + t2[k];
    memcpy(w[i], W, sizeof W); memcpy(al[i], A, sizeof A); memcpy(vo[i], V, sizeof V); memcpy(ao[i], Ac, sizeof Ac);
  }
  for (int i = 0; i < nb; i++) {
    double r[3], t[3], t2[3], ac[3], Iw_a[3], Iw_w[3], loc[3], n[3];
    for (int k = 0; k < 3; k++) r[k] = d->xipos[i][k] - d->xpos[i][k];
    cross3(t, w[i], r); cross3(t2, w[i], t); cross3(t, al[i], r);
    for (int k = 0; k < 3; k++) ac[k] = ao[i][k] + t[k] + t2[k];
    for (int k = 0; k < 3; k++) F[i][k] = m->body_mass[i] * ac[k];
    matT_vec(loc, d->ximat[i], al[i]);
    for (int k = 0; k < 3; k++) loc[k] *= m->body_inertia[i][k];
    mat_vec(Iw_a, d->ximat[i], loc);
    matT_vec(loc, d->ximat[i], w[i]);
    for (int k = 0; k < 3; k++) loc[k] *= m->body_inertia[i][k];
    mat_vec(Iw_w, d->ximat[i], loc);
    cross3(t, w[i], Iw_w);
    cross3(n, d->xipos[i], F[i]);
    for (int k = 0; k < 3; k++) N0[i][k] = Iw_a[k] + t[k] + n[k];
  }
  for (int i = nb - 1; i >= 1; i--) {
    for (int jj = m->body_jntnum[i] - 1; jj >= 0; jj--) {
      int j = m->body_jntadr[i] + jj, da = m->jnt_dofadr[j];
      if (m->jnt_type[j] == MJC_JNT_FREE) {
        double t[3], nn[3], loc[3];
        cross3(t, d->xpos[i], F[i]);
        for (int k = 0; k < 3; k++) { d->qfrc_bias[da + k] = F[i][k]; nn[k] = N0[i][k] - t[k]; }
        matT_vec(loc, d->xmat[i], nn);
        for (int k = 0; k < 3; k++) d->qfrc_bias[da + 3 + k] = loc[k];
      } else if (m->jnt_type[j] == MJC_JNT_SLIDE) {
        d->qfrc_bias[da] = dot3(d->xaxis[j], F[i]);
      } else {
        double t[3], nn[3];
        cross3(t, d->xanchor[j], F[i]);
        for (int k = 0; k < 3; k++) nn[k] = N0[i][k] - t[k];
        d->qfrc_bias[da] = dot3(d->xaxis[j], nn);
      }
    }
    int p = m->body_parent[i];
    for (int k = 0; k < 3; k++) { F[p][k] += F[i][k]; N0[p][k] += N0[i][k]; }
  }
}

/* ------------------------------------------------------------------ collision (mj_collision) */
/* mju_makeFrame: complete an orthonormal frame from its first axis (the contact normal). */
static void make_frame(double* frame) {
  double* x = frame; double* y = frame + 3; double* z = frame + 6;
  normalize3(x);
  if (fabs(x[1]) < 0.5) { y[0] = 0; y[1] = 1; y[2] = 0; } else { y[0] = 0; y[1] = 0; y[2] = 1; }
  double dd = dot3(x, y);
  for (int k = 0; k < 3; k++) y[k] -= dd * x[k];
  normalize3(y);
  cross3(z, x, y);
}

typedef struct { double dist, pos[3], normal[3]; } RawContact;

/* Upright cylinder vs upright cylinder (axes parallel to world z). MuJoCo has no analytic
 * cylinder-cylinder routine and runs its general convex collider (tolerance 1e-6); for parallel axes with
 * lateral overlap smaller than the axial overlap the penetration direction is the horizontal centre line
 * and the depth is |d_xy| - (r1 + r2) — that closed form is what is restated here (REDUCED). */
static int collide_cylinder_cylinder(const double* p1, const double* m1, const double* s1, const double* p2,
                                     const double* m2, const double* s2, double margin, RawContact* out) {
  (void)m1; (void)m2;
  double dx = p2[0] - p1[0], dy = p2[1] - p1[1];
  double dxy = sqrt(dx * dx + dy * dy);
  double lat = dxy - (s1[0] + s2[0]);
  double zlo = fmax(p1[2] - s1[1], p2[2] - s2[1]), zhi = fmin(p1[2] + s1[1], p2[2] + s2[1]);
  if (zhi - zlo <= 0 || lat >= margin || dxy < MINVAL) return 0;
  out->dist = lat;
  out->normal[0] = dx / dxy; out->normal[1] = dy / dxy; out->normal[2] = 0;
  double mid = s1[0] + 0.5 * lat;
  out->pos[0] = p1[0] + out->normal[0] * mid; out->pos[1] = p1[1] + out->normal[1] * mid; out->pos[2] = 0.5 * (zlo + zhi);
  return 1;
}

/* Sphere vs sphere (mjc_SphereSphere semantics: normal from geom1 to geom2, position midway between the surfaces). */
static int collide_sphere_sphere(const double* p1, double r1, const double* p2, double r2, double margin, RawContact* out) {
  double dv[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  double dn = norm3(dv);
  if (dn - r1 - r2 >= margin) return 0;
  if (dn < MINVAL) { out->normal[0] = 1; out->normal[1] = out->normal[2] = 0; }
  else for (int k = 0; k < 3; k++) out->normal[k] = dv[k] / dn;
  out->dist = dn - r1 - r2;
  for (int k = 0; k < 3; k++) out->pos[k] = p1[k] + out->normal[k] * (r1 + 0.5 * out->dist);
  return 1;
}

/* Sphere vs box (mjc_SphereBox semantics: closest point on the box, inside case pushes out the nearest face). */
static int collide_sphere_box(const double* ps, double rs, const double* pb, const double* mb, const double* sb,
                              double margin, RawContact* out) {
  double rel[3], loc[3], cl[3];
  for (int k = 0; k < 3; k++) rel[k] = ps[k] - pb[k];
  matT_vec(loc, mb, rel);
  int inside = 1;
  for (int k = 0; k < 3; k++) {
    cl[k] = fmin(fmax(loc[k], -sb[k]), sb[k]);
    if (cl[k] != loc[k]) inside = 0;
  }
  double nl[3], dist;
  if (!inside) {
    double dv[3] = {loc[0] - cl[0], loc[1] - cl[1], loc[2] - cl[2]};
    double dn = norm3(dv);
    if (dn - rs >= margin) return 0;
    for (int k = 0; k < 3; k++) nl[k] = -dv[k] / dn; /* from sphere (geom1) toward box (geom2) */
    dist = dn - rs;
  } else {
    int ax = 0; double best = 1e300;
    for (int k = 0; k < 3; k++) { double g = sb[k] - fabs(loc[k]); if (g < best) { best = g; ax = k; } }
    nl[0] = nl[1] = nl[2] = 0; nl[ax] = loc[ax] >= 0 ? -1 : 1;
    cl[ax] = loc[ax] >= 0 ? sb[ax] : -sb[ax];
    dist = -best - rs;
  }
  mat_vec(out->normal, mb, nl);
  double clw[3];
  mat_vec(clw, mb, cl);
  for (int k = 0; k < 3; k++) out->pos[k] = pb[k] + clw[k] - out->normal[k] * (-0.5 * dist);
  out->dist = dist;
  return 1;
}

/* Box vs box: separating-axis test over the 15 axes, then reference-face clipping (face contact, up to 8
 * points) or closest points of the two edges (edge contact). REDUCED restatement of mjc_BoxBox: same contact
 * semantics (normal from geom1 to geom2, dist<0 penetration, pos midway), own clipping order. */
static int clip_poly(double (*poly)[2], int n, int axis, double sign, double lim) {
  double out[16][2];
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
    if (no >= 15) break;
  }
  for (int i = 0; i < no; i++) { poly[i][0] = out[i][0]; poly[i][1] = out[i][1]; }
  return no;
}

static int collide_box_box(const double* p1, const double* m1, const double* s1, const double* p2, const double* m2,
                           const double* s2, double margin, RawContact* out, int maxout) {
  double R[3][3], AR[3][3], t[3], d12[3];
  for (int k = 0; k < 3; k++) d12[k] = p2[k] - p1[k];
  matT_vec(t, m1, d12); /* centre of box 2 in box-1 frame */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      R[i][j] = m1[i] * m2[j] + m1[3 + i] * m2[3 + j] + m1[6 + i] * m2[6 + j]; /* axis_i(1) . axis_j(2) */
      AR[i][j] = fabs(R[i][j]) + 1e-12;
    }
  double best = -1e300; int code = -1; double bsign = 1;
  /* face axes of box 1 */
  for (int i = 0; i < 3; i++) {
    double sep = fabs(t[i]) - (s1[i] + s2[0] * AR[i][0] + s2[1] * AR[i][1] + s2[2] * AR[i][2]);
    if (sep >= margin) return 0;
    if (sep > best) { best = sep; code = i; bsign = t[i] >= 0 ? 1 : -1; }
  }
  /* face axes of box 2 */
  for (int j = 0; j < 3; j++) {
    double tj = t[0] * R[0][j] + t[1] * R[1][j] + t[2] * R[2][j];
    double sep = fabs(tj) - (s2[j] + s1[0] * AR[0][j] + s1[1] * AR[1][j] + s1[2] * AR[2][j]);
    if (sep >= margin) return 0;
    if (sep > best) { best = sep; code = 3 + j; bsign = tj >= 0 ? 1 : -1; }
  }
  /* edge-edge axes; prefer faces unless an edge axis is clearly better (relative 5% + absolute 1e-6 bias) */
  double ebest = -1e300; int ecode = -1; double esign = 1, eaxis[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double a1[3] = {m1[i], m1[3 + i], m1[6 + i]}, a2[3] = {m2[j], m2[3 + j], m2[6 + j]}, ax[3];
      cross3(ax, a1, a2);
      double len = norm3(ax);
      if (len < 1e-8) continue;
      for (int k = 0; k < 3; k++) ax[k] /= len;
      double td = dot3(ax, d12), ra = 0, rb = 0;
      for (int k = 0; k < 3; k++) {
        double c1[3] = {m1[k], m1[3 + k], m1[6 + k]}, c2[3] = {m2[k], m2[3 + k], m2[6 + k]};
        ra += s1[k] * fabs(dot3(ax, c1)); rb += s2[k] * fabs(dot3(ax, c2));
      }
      double sep = fabs(td) - (ra + rb);
      if (sep >= margin) return 0;
      if (sep > ebest) { ebest = sep; ecode = 6 + 3 * i + j; esign = td >= 0 ? 1 : -1; memcpy(eaxis, ax, sizeof ax); }
    }
  if (ecode >= 0 && ebest > best + 1e-6 + 0.05 * fabs(best)) {
    /* edge-edge contact: closest points between the supporting edges */
    int i = (ecode - 6) / 3, j = (ecode - 6) % 3;
    double n[3] = {eaxis[0] * esign, eaxis[1] * esign, eaxis[2] * esign};
    double c1[3], c2[3];
    memcpy(c1, p1, sizeof c1); memcpy(c2, p2, sizeof c2);
    for (int k = 0; k < 3; k++) {
      if (k != i) { double a[3] = {m1[k], m1[3 + k], m1[6 + k]}; double sg = dot3(a, n) > 0 ? 1 : -1; for (int q = 0; q < 3; q++) c1[q] += sg * s1[k] * a[q]; }
      if (k != j) { double a[3] = {m2[k], m2[3 + k], m2[6 + k]}; double sg = dot3(a, n) > 0 ? -1 : 1; for (int q = 0; q < 3; q++) c2[q] += sg * s2[k] * a[q]; }
    }
    double u1[3] = {m1[i], m1[3 + i], m1[6 + i]}, u2[3] = {m2[j], m2[3 + j], m2[6 + j]}, w0[3];
    for (int k = 0; k < 3; k++) w0[k] = c1[k] - c2[k];
    double b = dot3(u1, u2), dd = dot3(u1, w0), e = dot3(u2, w0), den = 1 - b * b;
    double sc = den > 1e-12 ? (b * e - dd) / den : 0, tc = den > 1e-12 ? (e - b * dd) / den : 0;
    sc = fmin(fmax(sc, -s1[i]), s1[i]); tc = fmin(fmax(tc, -s2[j]), s2[j]);
    double q1[3], q2[3];
    for (int k = 0; k < 3; k++) { q1[k] = c1[k] + sc * u1[k]; q2[k] = c2[k] + tc * u2[k]; }
    out[0].dist = ebest;
    memcpy(out[0].normal, n, sizeof n);
    for (int k = 0; k < 3; k++) out[0].pos[k] = 0.5 * (q1[k] + q2[k]);
    return 1;
  }
  /* face contact: reference box owns the axis; incident face = the face of the other box most anti-parallel */
  const double *pr, *mr, *sr, *pi, *mi, *si;
  int raxis; double nsign; /* nsign: reference outward normal sign along its axis */
  int ref_is_1 = code < 3;
  if (ref_is_1) { pr = p1; mr = m1; sr = s1; pi = p2; mi = m2; si = s2; raxis = code; nsign = bsign; }
  else { pr = p2; mr = m2; sr = s2; pi = p1; mi = m1; si = s1; raxis = code - 3; nsign = -bsign; }
  double nref[3] = {mr[raxis] * nsign, mr[3 + raxis] * nsign, mr[6 + raxis] * nsign}; /* outward normal of reference face */
  int iaxis = 0; double imin = 1e300, isign = 1;
  for (int k = 0; k < 3; k++) {
    double a[3] = {mi[k], mi[3 + k], mi[6 + k]}, dd = dot3(a, nref);
    if (-fabs(dd) < imin) { imin = -fabs(dd); iaxis = k; isign = dd > 0 ? -1 : 1; }
  }
  int iu = (iaxis + 1) % 3, iv = (iaxis + 2) % 3, ru = (raxis + 1) % 3, rv = (raxis + 2) % 3;
  double fc[3];
  for (int k = 0; k < 3; k++) fc[k] = pi[k] + isign * si[iaxis] * mi[3 * k + iaxis];
  double poly[16][2], depth3[16];
  (void)depth3;
  int n = 4;
  const double sg[4][2] = {{1, 1}, {-1, 1}, {-1, -1}, {1, -1}};
  double verts[4][3];
  for (int c = 0; c < 4; c++) {
    for (int k = 0; k < 3; k++)
      verts[c][k] = fc[k] + sg[c][0] * si[iu] * mi[3 * k + iu] + sg[c][1] * si[iv] * mi[3 * k + iv] - pr[k];
    double loc[3];
    matT_vec(loc, mr, verts[c]);
    poly[c][0] = loc[ru]; poly[c][1] = loc[rv];
  }
  /* incident-face plane in reference coordinates: height(h) along the reference axis is affine in (u,v) */
  double l0[3], l1[3], l2[3];
  matT_vec(l0, mr, verts[0]); matT_vec(l1, mr, verts[1]); matT_vec(l2, mr, verts[3]);
  double e1u = l1[ru] - l0[ru], e1v = l1[rv] - l0[rv], e1h = l1[raxis] - l0[raxis];
  double e2u = l2[ru] - l0[ru], e2v = l2[rv] - l0[rv], e2h = l2[raxis] - l0[raxis];
  double det = e1u * e2v - e1v * e2u;
  n = clip_poly(poly, n, 0, 1, sr[ru]);
  if (n) n = clip_poly(poly, n, 0, -1, sr[ru]);
  if (n) n = clip_poly(poly, n, 1, 1, sr[rv]);
  if (n) n = clip_poly(poly, n, 1, -1, sr[rv]);
  int nc = 0;
  for (int c = 0; c < n && nc < maxout; c++) {
    double du = poly[c][0] - l0[ru], dv = poly[c][1] - l0[rv], h;
    if (fabs(det) > 1e-14) {
      double a = (du * e2v - dv * e2u) / det, b = (e1u * dv - e1v * du) / det;
      h = l0[raxis] + a * e1h + b * e2h;
    } else h = l0[raxis];
    double dist = nsign * h - sr[raxis]; /* signed distance of the incident point to the reference face */
    if (dist >= margin) continue;
    double loc[3], wpt[3];
    loc[ru] = poly[c][0]; loc[rv] = poly[c][1]; loc[raxis] = h - 0.5 * dist * nsign; /* midway between the surfaces */
    mat_vec(wpt, mr, loc);
    out[nc].dist = dist;
    for (int k = 0; k < 3; k++) { out[nc].pos[k] = pr[k] + wpt[k]; out[nc].normal[k] = ref_is_1 ? nref[k] : -nref[k]; }
    nc++;
  }
  return nc;
}


/* ------------------------------------------------------------------ signed box-box distance (distance sensors)
 * mj_geomDistance restated for the reduced (box-only) geometry: penetrating boxes -> minus the penetration depth
 * (largest separating-axis value over the 15 SAT axes, which is exact for boxes); separated boxes -> the exact
 * Euclidean distance by GJK on the Minkowski difference (Gilbert-Johnson-Keerthi with Ericson's closest-point
 * sub-algorithms for segment / triangle / tetrahedron). */
static void box_support(const double* p, const double* m, const double* s, const double* dir, double* out) {
  for (int k = 0; k < 3; k++) out[k] = p[k];
  for (int a = 0; a < 3; a++) {
    double ax[3] = {m[a], m[3 + a], m[6 + a]};
    double sg = dot3(ax, dir) >= 0 ? s[a] : -s[a];
    for (int k = 0; k < 3; k++) out[k] += sg * ax[k];
  }
}
/* closest point to the origin on segment ab; keeps the supporting vertices in W (n updated) */
static void closest_segment(double (*W)[3], int* n, double* v) {
  const double *a = W[0], *b = W[1];
  double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
  double t = -dot3(a, ab), den = dot3(ab, ab);
  if (t <= 0 || den <= 0) { *n = 1; memcpy(v, a, 3 * sizeof(double)); return; }
  if (t >= den) { memcpy(W[0], b, 3 * sizeof(double)); *n = 1; memcpy(v, W[0], 3 * sizeof(double)); return; }
  t /= den;
  for (int k = 0; k < 3; k++) v[k] = a[k] + t * ab[k];
}
/* closest point to the origin on triangle (Ericson 5.1.5), reducing W to the supporting feature */
static void closest_triangle(double (*W)[3], int* n, double* v) {
  double a[3], b[3], c[3];
  memcpy(a, W[0], sizeof a); memcpy(b, W[1], sizeof b); memcpy(c, W[2], sizeof c);
  double ab[3], ac[3], bc[3];
  for (int k = 0; k < 3; k++) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; bc[k] = c[k] - b[k]; }
  double d1 = -dot3(ab, a), d2 = -dot3(ac, a);
  if (d1 <= 0 && d2 <= 0) { *n = 1; memcpy(v, a, sizeof a); return; }
  double d3 = -dot3(ab, b), d4 = -dot3(ac, b);
  if (d3 >= 0 && d4 <= d3) { memcpy(W[0], b, sizeof b); *n = 1; memcpy(v, b, sizeof b); return; }
  double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) {
    double t = d1 / (d1 - d3);
    *n = 2; /* a, b already in W[0], W[1] */
    for (int k = 0; k < 3; k++) v[k] = a[k] + t * ab[k];
    return;
  }
  double d5 = -dot3(ab, c), d6 = -dot3(ac, c);
  if (d6 >= 0 && d5 <= d6) { memcpy(W[0], c, sizeof c); *n = 1; memcpy(v, c, sizeof c); return; }
  double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) {
    double t = d2 / (d2 - d6);
    memcpy(W[1], c, sizeof c); *n = 2;
    for (int k = 0; k < 3; k++) v[k] = a[k] + t * ac[k];
    return;
  }
  double va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    double t = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    memcpy(W[0], b, sizeof b); memcpy(W[1], c, sizeof c); *n = 2;
    for (int k = 0; k < 3; k++) v[k] = b[k] + t * bc[k];
    return;
  }
  double den = 1.0 / (va + vb + vc), tv = vb * den, tw = vc * den;
  *n = 3;
  for (int k = 0; k < 3; k++) v[k] = a[k] + ab[k] * tv + ac[k] * tw;
}
/* closest point to the origin on a tetrahedron: the best of its four faces (all four are evaluated: with the nearly
 * flat tetrahedra that parallel box faces produce, the inside/outside sign tests alone are not reliable);
 * returns 1 when the origin is strictly inside (the sets intersect) */
static int closest_tetrahedron(double (*W)[3], int* n, double* v) {
  static const int F[4][4] = {{0, 1, 2, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {1, 3, 2, 0}}; /* face (i,j,k), opposite vertex l */
  double best = 1e300, bv[3] = {0, 0, 0}, BW[3][3];
  int bn = 0, outside_any = 0;
  double P[4][3];
  memcpy(P, W, sizeof P);
  for (int f = 0; f < 4; f++) {
    const double *a = P[F[f][0]], *b = P[F[f][1]], *c = P[F[f][2]], *dd = P[F[f][3]];
    double ab[3], ac[3], nrm[3];
    for (int k = 0; k < 3; k++) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; }
    cross3(nrm, ab, ac);
    double so = -dot3(a, nrm);                                        /* origin side */
    double ad[3] = {dd[0] - a[0], dd[1] - a[1], dd[2] - a[2]};
    double sd = dot3(ad, nrm);                                        /* opposite-vertex side */
    if (!(so * sd > 0)) outside_any = 1; /* origin not strictly on the inner side of this face (or the face is degenerate) */
    double T[3][3], tv[3];
    int tn = 3;
    memcpy(T[0], a, 24); memcpy(T[1], b, 24); memcpy(T[2], c, 24);
    closest_triangle(T, &tn, tv);
    double dist = dot3(tv, tv);
    if (dist < best) { best = dist; bn = tn; memcpy(bv, tv, sizeof bv); memcpy(BW, T, sizeof BW); }
  }
  if (!outside_any) return 1;
  *n = bn; memcpy(v, bv, sizeof bv);
  for (int i = 0; i < bn; i++) memcpy(W[i], BW[i], 24);
  return 0;
}
static double box_box_distance(const double* p1, const double* m1, const double* s1, const double* p2, const double* m2,
                               const double* s2, double cutoff) {
  /* separating-axis values (<= true distance); all negative -> penetration depth */
  double d12[3], best = -1e300;
  for (int k = 0; k < 3; k++) d12[k] = p2[k] - p1[k];
  for (int which = 0; which < 2; which++)
    for (int i = 0; i < 3; i++) {
      const double* mm = which ? m2 : m1;
      double ax[3] = {mm[i], mm[3 + i], mm[6 + i]}, ra = 0, rb = 0;
      for (int k = 0; k < 3; k++) {
        double c1[3] = {m1[k], m1[3 + k], m1[6 + k]}, c2[3] = {m2[k], m2[3 + k], m2[6 + k]};
        ra += s1[k] * fabs(dot3(ax, c1)); rb += s2[k] * fabs(dot3(ax, c2));
      }
      double sep = fabs(dot3(ax, d12)) - (ra + rb);
      if (sep > best) best = sep;
    }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double a1[3] = {m1[i], m1[3 + i], m1[6 + i]}, a2[3] = {m2[j], m2[3 + j], m2[6 + j]}, ax[3];
      cross3(ax, a1, a2);
      double len = norm3(ax);
      if (len < 1e-8) continue;
      for (int k = 0; k < 3; k++) ax[k] /= len;
      double ra = 0, rb = 0;
      for (int k = 0; k < 3; k++) {
        double c1[3] = {m1[k], m1[3 + k], m1[6 + k]}, c2[3] = {m2[k], m2[3 + k], m2[6 + k]};
        ra += s1[k] * fabs(dot3(ax, c1)); rb += s2[k] * fabs(dot3(ax, c2));
      }
      double sep = fabs(dot3(ax, d12)) - (ra + rb);
      if (sep > best) best = sep;
    }
  if (best <= 0) return best;
  if (best >= cutoff) return cutoff;
  /* GJK on A - B */
  double W[4][3], v[3] = {-d12[0], -d12[1], -d12[2]}, vv, vv_prev = 1e300;
  int n = 0;
  if (dot3(v, v) < 1e-30) { v[0] = 1; v[1] = v[2] = 0; }
  for (int it = 0; it < 64; it++) {
    double nd[3] = {-v[0], -v[1], -v[2]}, a[3], b[3], w[3];
    box_support(p1, m1, s1, nd, a);
    box_support(p2, m2, s2, v, b);
    for (int k = 0; k < 3; k++) w[k] = a[k] - b[k];
    vv = dot3(v, v);
    if (n > 0 && vv - dot3(v, w) <= 1e-13 * vv) break; /* no progress possible along -v: v is the closest point */
    /* |v| decreases strictly until the optimum; with parallel faces (a whole face of closest points) rounding makes the
     * simplex cycle between equivalent vertices instead: stop at the first iteration that no longer shortens v */
    if (n > 0 && vv >= vv_prev * (1 - 1e-15)) break;
    if (n > 0) vv_prev = vv; /* (the initial v, the centre difference, is not a point of the simplex yet) */
    int dup = 0;
    for (int i = 0; i < n; i++) if (W[i][0] == w[0] && W[i][1] == w[1] && W[i][2] == w[2]) dup = 1;
    if (dup) break;
    memcpy(W[n++], w, sizeof w);
    if (n == 1) memcpy(v, W[0], sizeof w);
    else if (n == 2) closest_segment(W, &n, v);
    else if (n == 3) closest_triangle(W, &n, v);
    else if (closest_tetrahedron(W, &n, v)) return best; /* numerically touching: fall back to the SAT value */
    if (dot3(v, v) < 1e-30) return best;
  }
  double dist = norm3(v);
  return dist < cutoff ? dist : cutoff;
}
/* test hook */
double mjc_box_box_distance(const double* p1, const double* m1, const double* s1, const double* p2, const double* m2,
                            const double* s2, double cutoff) { return box_box_distance(p1, m1, s1, p2, m2, s2, cutoff); }

/* mj_contactParam: combine the two geoms' parameters (equal priority: max friction, solmix-weighted solref/solimp). */
static void contact_param(const mjcModel* m, int g1, int g2, mjcContact* c) {
  int p1 = m->geom_priority[g1], p2 = m->geom_priority[g2];
  c->includemargin = fmax(m->geom_margin[g1], m->geom_margin[g2]) - fmax(m->geom_gap[g1], m->geom_gap[g2]);
  double fri[3];
  if (p1 != p2) {
    int g = p1 > p2 ? g1 : g2;
    c->dim = m->geom_condim[g];
    memcpy(c->solref, m->geom_solref[g], sizeof c->solref);
    memcpy(c->solimp, m->geom_solimp[g], sizeof c->solimp);
    memcpy(fri, m->geom_friction[g], sizeof fri);
  } else {
    c->dim = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
    double s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2], mix;
    if (s1 >= MINVAL && s2 >= MINVAL) mix = s1 / (s1 + s2);
    else if (s1 < MINVAL && s2 < MINVAL) mix = 0.5;
    else mix = s1 < MINVAL ? 0.0 : 1.0;
    if (m->geom_solref[g1][0] > 0 && m->geom_solref[g2][0] > 0)
      for (int k = 0; k < 2; k++) c->solref[k] = mix * m->geom_solref[g1][k] + (1 - mix) * m->geom_solref[g2][k];
    else
      for (int k = 0; k < 2; k++) c->solref[k] = fmin(m->geom_solref[g1][k], m->geom_solref[g2][k]);
    for (int k = 0; k < 5; k++) c->solimp[k] = mix * m->geom_solimp[g1][k] + (1 - mix) * m->geom_solimp[g2][k];
    for (int k = 0; k < 3; k++) fri[k] = fmax(m->geom_friction[g1][k], m->geom_friction[g2][k]);
  }
  c->friction[0] = c->friction[1] = fmax(MINMU, fri[0]);
  c->friction[2] = fmax(MINMU, fri[1]);
  c->friction[3] = c->friction[4] = fmax(MINMU, fri[2]);
}

static void collision(const mjcModel* m, mjcData* d) {
  d->ncon = 0;
  if (m->contact_disabled) return;
  for (int p = 0; p < m->npair; p++) {
    int g1 = m->pair_g1[p], g2 = m->pair_g2[p];
    int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
    if (t1 > t2) { int tmp = g1; g1 = g2; g2 = tmp; tmp = t1; t1 = t2; t2 = tmp; }
    double margin = fmax(m->geom_margin[g1], m->geom_margin[g2]);
    /* bounding-sphere rejection (broad phase): cannot touch if centres are farther than the summed radii */
    {
      double dc[3], r1 = 0, r2 = 0;
      for (int k = 0; k < 3; k++) dc[k] = d->geom_xpos[g2][k] - d->geom_xpos[g1][k];
      if (t1 == MJC_GEOM_BOX) r1 = norm3(m->geom_size[g1]); else if (t1 == MJC_GEOM_SPHERE) r1 = m->geom_size[g1][0];
      else r1 = sqrt(m->geom_size[g1][0] * m->geom_size[g1][0] + m->geom_size[g1][1] * m->geom_size[g1][1]) + m->geom_size[g1][0];
      if (t2 == MJC_GEOM_BOX) r2 = norm3(m->geom_size[g2]); else if (t2 == MJC_GEOM_SPHERE) r2 = m->geom_size[g2][0];
      else r2 = sqrt(m->geom_size[g2][0] * m->geom_size[g2][0] + m->geom_size[g2][1] * m->geom_size[g2][1]) + m->geom_size[g2][0];
      if (norm3(dc) > r1 + r2 + margin) continue;
    }
    RawContact raw[8];
    int n = 0;
    if (t1 == MJC_GEOM_CYLINDER && t2 == MJC_GEOM_CYLINDER)
      n = collide_cylinder_cylinder(d->geom_xpos[g1], d->geom_xmat[g1], m->geom_size[g1], d->geom_xpos[g2], d->geom_xmat[g2], m->geom_size[g2], margin, raw);
    else if (t1 == MJC_GEOM_BOX && t2 == MJC_GEOM_BOX)
      n = collide_box_box(d->geom_xpos[g1], d->geom_xmat[g1], m->geom_size[g1], d->geom_xpos[g2], d->geom_xmat[g2], m->geom_size[g2], margin, raw, 8);
    else if (t1 == MJC_GEOM_SPHERE && t2 == MJC_GEOM_SPHERE)
      n = collide_sphere_sphere(d->geom_xpos[g1], m->geom_size[g1][0], d->geom_xpos[g2], m->geom_size[g2][0], margin, raw);
    else if (t1 == MJC_GEOM_SPHERE && t2 == MJC_GEOM_BOX)
      n = collide_sphere_box(d->geom_xpos[g1], m->geom_size[g1][0], d->geom_xpos[g2], d->geom_xmat[g2], m->geom_size[g2], margin, raw);
    else
      continue; /* pair type not exercised by the BASELINE tasks */
    for (int i = 0; i < n && d->ncon < MJC_MAXCON; i++) {
      mjcContact* c = &d->contact[d->ncon];
      c->dist = raw[i].dist;
      memcpy(c->pos, raw[i].pos, sizeof c->pos);
      memcpy(c->frame, raw[i].normal, 3 * sizeof(double));
      make_frame(c->frame);
      c->geom1 = g1; c->geom2 = g2;
      contact_param(m, g1, g2, c);
      d->ncon++;
    }
  }
}

/* ------------------------------------------------------------------ constraints (mj_makeConstraint etc.) */
static int add_row(mjcData* d, int nv, int type, int id, double pos, double margin, double frictionloss, double diagApprox) {
  int r = d->nefc;
  if (r >= MJC_MAXEFC) return -1;
  memset(d->efc_J[r], 0, sizeof(double) * nv);
  d->efc_type[r] = type; d->efc_id[r] = id; d->efc_pos[r] = pos; d->efc_margin[r] = margin;
  d->efc_frictionloss[r] = frictionloss; d->efc_diagApprox[r] = diagApprox;
  d->nefc++;
  return r;
}

static void make_constraint(const mjcModel* m, mjcData* d) {
  int nv = m->nv;
  d->nefc = 0;
  /* mj_instantiateEquality, joint type: (q1 - q1_0) = poly(q2 - q2_0); always active */
  for (int e = 0; e < m->neq; e++) {
    int j1 = m->eq_j1[e], j2 = m->eq_j2[e];
    const double* c = m->eq_polycoef[e];
    double y = d->qpos[m->jnt_qposadr[j1]] - m->qpos0[m->jnt_qposadr[j1]];
    double x = d->qpos[m->jnt_qposadr[j2]] - m->qpos0[m->jnt_qposadr[j2]];
    double pos = y - (c[0] + x * (c[1] + x * (c[2] + x * (c[3] + x * c[4]))));
    double deriv = c[1] + x * (2 * c[2] + x * (3 * c[3] + x * 4 * c[4]));
    int d1 = m->jnt_dofadr[j1], d2 = m->jnt_dofadr[j2];
    int r = add_row(d, nv, CT_EQUALITY, e, pos, 0, 0, m->dof_invweight0[d1] + m->dof_invweight0[d2]);
    if (r >= 0) { d->efc_J[r][d1] = 1; d->efc_J[r][d2] = -deriv; }
  }
  /* mj_instantiateFriction: one row per dof with frictionloss */
  for (int i = 0; i < nv; i++)
    if (m->dof_frictionloss[i] > 0) {
      int r = add_row(d, nv, CT_FRICTION_DOF, i, 0, 0, m->dof_frictionloss[i], m->dof_invweight0[i]);
      if (r >= 0) d->efc_J[r][i] = 1;
    }
  /* mj_instantiateLimit: slide/hinge joint limits, active when dist < margin */
  for (int j = 0; j < m->njnt; j++) {
    if (!m->jnt_limited[j] || (m->jnt_type[j] != MJC_JNT_SLIDE && m->jnt_type[j] != MJC_JNT_HINGE)) continue;
    double q = d->qpos[m->jnt_qposadr[j]];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[j][side < 0 ? 0 : 1] - q);
      if (dist < m->jnt_margin[j]) {
        int r = add_row(d, nv, CT_LIMIT_JOINT, j, dist, m->jnt_margin[j], 0, m->dof_invweight0[m->jnt_dofadr[j]]);
        if (r >= 0) d->efc_J[r][m->jnt_dofadr[j]] = -side;
      }
    }
  }
  /* mj_instantiateContact */
  double jp1[3 * MJC_MAXNV], jp2[3 * MJC_MAXNV], jd[3 * MJC_MAXNV];
  for (int ci = 0; ci < d->ncon; ci++) {
    mjcContact* c = &d->contact[ci];
    c->efc_address = -1;
    if (c->dist >= c->includemargin) continue;
    int b1 = m->geom_body[c->geom1], b2 = m->geom_body[c->geom2];
    jac_point(m, d, b1, c->pos, jp1, NULL);
    jac_point(m, d, b2, c->pos, jp2, NULL);
    /* relative velocity of body 2 w.r.t. body 1, expressed in the contact frame */
    for (int a = 0; a < 3; a++)
      for (int i = 0; i < nv; i++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += c->frame[3 * a + k] * (jp2[k * nv + i] - jp1[k * nv + i]);
        jd[a * nv + i] = s;
      }
    double tran = m->body_invweight0[b1][0] + m->body_invweight0[b2][0];
    c->efc_address = d->nefc;
    if (c->dim == 1) {
      int r = add_row(d, nv, CT_CONTACT_PYRAMIDAL, ci, c->dist, c->includemargin, 0, tran);
      if (r >= 0) memcpy(d->efc_J[r], jd, sizeof(double) * nv);
    } else if (m->cone == MJC_CONE_ELLIPTIC) {
      for (int a = 0; a < 3; a++) {
        int r = add_row(d, nv, CT_CONTACT_ELLIPTIC, ci, c->dist, c->includemargin, 0, tran);
        if (r >= 0) memcpy(d->efc_J[r], jd + a * nv, sizeof(double) * nv);
      }
    } else {
      for (int a = 1; a < 3; a++)
        for (int sgn = 1; sgn >= -1; sgn -= 2) {
          double mu = c->friction[a - 1];
          int r = add_row(d, nv, CT_CONTACT_PYRAMIDAL, ci, c->dist, c->includemargin, 0, tran + mu * mu * tran);
          if (r >= 0) for (int i = 0; i < nv; i++) d->efc_J[r][i] = jd[i] + sgn * mu * jd[a * nv + i];
        }
    }
  }
  /* efc_vel = J qvel */
  for (int r = 0; r < d->nefc; r++) {
    double s = 0;
    for (int i = 0; i < nv; i++) s += d->efc_J[r][i] * d->qvel[i];
    d->efc_vel[r] = s;
  }
}

/* getimpedance: sigmoid impedance d(r) from solimp = (d0, dwidth, width, midpoint, power) */
static double impedance(const double* solimp, double pos, double margin) {
  double d0 = fmin(fmax(solimp[0], MINIMP), MAXIMP), dw = fmin(fmax(solimp[1], MINIMP), MAXIMP);
  double width = solimp[2], mid = fmin(fmax(solimp[3], MINIMP), MAXIMP), power = fmax(solimp[4], 1.0);
  if (d0 == dw || width <= MINVAL) return 0.5 * (d0 + dw);
  double x = fabs((pos - margin) / width);
  if (x >= 1) return dw;
  if (x <= 0) return d0;
  double y;
  if (power == 1) y = x;
  else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
  else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
  return d0 + y * (dw - d0);
}

/* mj_makeImpedance + mj_referenceConstraint: R, D and aref for every row */
static void make_impedance(const mjcModel* m, mjcData* d) {
  for (int r = 0; r < d->nefc; r++) {
    const double *solref, *solimp;
    int tp = d->efc_type[r], id = d->efc_id[r];
    int friction_row = 0;
    if (tp == CT_EQUALITY) { solref = m->eq_solref[id]; solimp = m->eq_solimp[id]; }
    else if (tp == CT_FRICTION_DOF) { solref = m->dof_solref[id]; solimp = m->dof_solimp[id]; friction_row = 1; }
    else if (tp == CT_LIMIT_JOINT) { solref = m->jnt_solref[id]; solimp = m->jnt_solimp[id]; }
    else {
      solref = d->contact[id].solref; solimp = d->contact[id].solimp;
      if (tp == CT_CONTACT_ELLIPTIC && r > d->contact[id].efc_address) friction_row = 1;
    }
    double ref0 = solref[0], ref1 = solref[1];
    double dmax = fmin(fmax(solimp[1], MINIMP), MAXIMP);
    double imp = impedance(solimp, d->efc_pos[r], d->efc_margin[r]);
    double K, B;
    if (ref0 > 0) { /* standard (timeconst, dampratio); refsafe: timeconst >= 2*timestep */
      if (ref0 < 2 * m->timestep) ref0 = 2 * m->timestep;
      K = 1 / fmax(MINVAL, dmax * dmax * ref0 * ref0 * ref1 * ref1);
      B = 2 / fmax(MINVAL, dmax * ref0);
    } else { K = -ref0 / fmax(MINVAL, dmax * dmax); B = -ref1 / fmax(MINVAL, dmax); }
    if (friction_row) K = 0;
    d->efc_R[r] = fmax(MINVAL, (1 - imp) * d->efc_diagApprox[r] / imp);
    d->efc_aref[r] = -B * d->efc_vel[r] - K * imp * (d->efc_pos[r] - d->efc_margin[r]);
  }
  /* frictional contacts: adjust R in the friction dimensions and set the regularised cone's mu */
  for (int ci = 0; ci < d->ncon; ci++) {
    mjcContact* c = &d->contact[ci];
    int i = c->efc_address;
    if (i < 0 || c->dim < 3) continue;
    double* R = d->efc_R;
    R[i + 1] = R[i] / fmax(MINVAL, m->impratio);
    c->mu = c->friction[0] * sqrt(R[i + 1] / R[i]);
    if (m->cone == MJC_CONE_ELLIPTIC) {
      for (int j = 1; j < c->dim - 1; j++) R[i + j + 1] = R[i + 1] * c->friction[0] * c->friction[0] / (c->friction[j] * c->friction[j]);
    } else {
      double Rpy = 2 * c->mu * c->mu * R[i]; /* = 2 friction^2 R0 / impratio (regularised mu): MuJoCo's common pyramid-edge R */
      for (int j = 0; j < 2 * (c->dim - 1); j++) R[i + j] = Rpy;
    }
  }
  for (int r = 0; r < d->nefc; r++) d->efc_D[r] = 1 / d->efc_R[r];
}

/* ------------------------------------------------------------------ Newton solver (mj_solNewton, primal) */
typedef struct {
  int nv, nefc;
  double Ma[MJC_MAXNV], jar[MJC_MAXEFC], grad[MJC_MAXNV], search[MJC_MAXNV], Mv[MJC_MAXNV], jv[MJC_MAXEFC];
  double H[MJC_MAXNV * MJC_MAXNV], Lh[MJC_MAXNV * MJC_MAXNV];
  double cost, gauss;
} Solver;

/* one constraint (or one elliptic contact) evaluated at x = jar + alpha*jv: adds cost, d/dalpha, d2/dalpha2 */
static void row_eval(const mjcModel* m, const mjcData* d, int r, const double* x, double* cost, double* force, int* state,
                     double* Hc /* 3x3 cone Hessian or NULL */) {
  int tp = d->efc_type[r];
  double D = d->efc_D[r];
  if (tp == CT_EQUALITY) {
    force[0] = -D * x[0]; *cost += 0.5 * D * x[0] * x[0]; state[0] = ST_QUADRATIC;
  } else if (tp == CT_FRICTION_DOF) {
    double f = d->efc_frictionloss[r], R = d->efc_R[r];
    if (x[0] <= -R * f) { force[0] = f; *cost += -0.5 * R * f * f - f * x[0]; state[0] = ST_LINEARNEG; }
    else if (x[0] >= R * f) { force[0] = -f; *cost += -0.5 * R * f * f + f * x[0]; state[0] = ST_LINEARPOS; }
    else { force[0] = -D * x[0]; *cost += 0.5 * D * x[0] * x[0]; state[0] = ST_QUADRATIC; }
  } else if (tp == CT_LIMIT_JOINT || tp == CT_CONTACT_PYRAMIDAL) {
    if (x[0] < 0) { force[0] = -D * x[0]; *cost += 0.5 * D * x[0] * x[0]; state[0] = ST_QUADRATIC; }
    else { force[0] = 0; state[0] = ST_SATISFIED; }
  } else { /* elliptic contact, r is its first row, dim = 3 */
    const mjcContact* c = &d->contact[d->efc_id[r]];
    double mu = c->mu, f1 = c->friction[0], f2 = c->friction[1];
    double U0 = x[0] * mu, U1 = x[1] * f1, U2 = x[2] * f2;
    double N = U0, T = sqrt(U1 * U1 + U2 * U2);
    if (N >= mu * T) { /* top zone: separating */
      force[0] = force[1] = force[2] = 0; state[0] = state[1] = state[2] = ST_SATISFIED;
    } else if (mu * N + T <= 0) { /* bottom zone: fully quadratic */
      for (int k = 0; k < 3; k++) { force[k] = -d->efc_D[r + k] * x[k]; *cost += 0.5 * d->efc_D[r + k] * x[k] * x[k]; state[k] = ST_QUADRATIC; }
    } else { /* middle zone: distance to the dual cone */
      double Dm = D / (mu * mu * (1 + mu * mu)), NmT = N - mu * T;
      *cost += 0.5 * Dm * NmT * NmT;
      force[0] = -Dm * NmT * mu;
      force[1] = T > MINVAL ? -force[0] / T * U1 * f1 : 0;
      force[2] = T > MINVAL ? -force[0] / T * U2 * f2 : 0;
      state[0] = state[1] = state[2] = ST_CONE;
      if (Hc) {
        double S[3] = {mu, f1, f2}, U[3] = {U0, U1, U2}, HU[9];
        double Ti = T > MINVAL ? 1 / T : 0;
        HU[0] = Dm;
        for (int j = 1; j < 3; j++) HU[j] = HU[3 * j] = -Dm * mu * U[j] * Ti;
        for (int j = 1; j < 3; j++)
          for (int k = 1; k < 3; k++)
            HU[3 * j + k] = Dm * mu * mu * U[j] * U[k] * Ti * Ti - Dm * mu * NmT * ((j == k ? Ti : 0) - U[j] * U[k] * Ti * Ti * Ti);
        for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) Hc[3 * j + k] = S[j] * HU[3 * j + k] * S[k];
      }
    }
  }
}

static int row_span(const mjcModel* m, const mjcData* d, int r) {
  (void)m;
  return d->efc_type[r] == CT_CONTACT_ELLIPTIC ? d->contact[d->efc_id[r]].dim : 1;
}

/* cost, forces, states at the current jar; optionally assemble the Newton Hessian */
static void constraint_update(const mjcModel* m, mjcData* d, Solver* s, int want_hessian) {
  int nv = s->nv;
  double cost = 0;
  if (want_hessian) memcpy(s->H, d->M, sizeof(double) * nv * nv);
  for (int r = 0; r < s->nefc;) {
    int span = row_span(m, d, r);
    double Hc[9];
    row_eval(m, d, r, s->jar + r, &cost, d->efc_force + r, d->efc_state + r, want_hessian ? Hc : NULL);
    if (want_hessian) {
      if (d->efc_state[r] == ST_QUADRATIC) {
        for (int k = 0; k < span; k++) {
          if (d->efc_state[r + k] != ST_QUADRATIC) continue;
          const double* J = d->efc_J[r + k];
          double D = d->efc_D[r + k];
          for (int i = 0; i < nv; i++) { if (J[i] == 0) continue; double a = D * J[i]; for (int j = 0; j < nv; j++) s->H[i * nv + j] += a * J[j]; }
        }
      } else if (d->efc_state[r] == ST_CONE) {
        for (int a = 0; a < 3; a++)
          for (int b = 0; b < 3; b++) {
            double h = Hc[3 * a + b];
            if (h == 0) continue;
            const double *Ja = d->efc_J[r + a], *Jb = d->efc_J[r + b];
            for (int i = 0; i < nv; i++) { if (Ja[i] == 0) continue; double t = h * Ja[i]; for (int j = 0; j < nv; j++) s->H[i * nv + j] += t * Jb[j]; }
          }
      }
    }
    r += span;
  }
  /* Gauss term: 0.5 (Ma - qfrc_smooth)^T (qacc - qacc_smooth) */
  double g = 0;
  for (int i = 0; i < nv; i++) g += (s->Ma[i] - d->qfrc_smooth[i]) * (d->qacc[i] - d->qacc_smooth[i]);
  s->gauss = 0.5 * g;
  s->cost = s->gauss + cost;
  for (int i = 0; i < nv; i++) {
    double f = 0;
    for (int r = 0; r < s->nefc; r++) f += d->efc_J[r][i] * d->efc_force[r];
    d->qfrc_constraint[i] = f;
    s->grad[i] = s->Ma[i] - d->qfrc_smooth[i] - f;
  }
}

/* derivative information of the 1-D cost along the search direction at step alpha */
static void ls_eval(const mjcModel* m, const mjcData* d, const Solver* s, double alpha, double g1, double g2,
                    double* cost, double* d1, double* d2) {
  if (g_debug < 0) g_ls_evals++;
  double c = alpha * g1 + 0.5 * alpha * alpha * g2, p1 = g1 + alpha * g2, p2 = g2;
  for (int r = 0; r < s->nefc;) {
    int span = row_span(m, d, r), st[3];
    double x[3], f[3], Hc[9], rc = 0;
    for (int k = 0; k < span; k++) x[k] = s->jar[r + k] + alpha * s->jv[r + k];
    row_eval(m, d, r, x, &rc, f, st, Hc);
    c += rc;
    for (int k = 0; k < span; k++) p1 -= f[k] * s->jv[r + k];
    if (st[0] == ST_CONE) {
      for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) p2 += s->jv[r + a] * Hc[3 * a + b] * s->jv[r + b];
    } else {
      for (int k = 0; k < span; k++) if (st[k] == ST_QUADRATIC) p2 += d->efc_D[r + k] * s->jv[r + k] * s->jv[r + k];
    }
    r += span;
  }
  *cost = c; *d1 = p1; *d2 = p2;
}

/* exact line search: safeguarded 1-D Newton on the convex C1 cost; returns the step */
static double line_search(const mjcModel* m, const mjcData* d, const Solver* s) {
  int nv = s->nv;
  double g1 = 0, g2 = 0, snorm = 0;
  for (int i = 0; i < nv; i++) {
    g1 += s->search[i] * (s->Ma[i] - d->qfrc_smooth[i]);
    g2 += s->search[i] * s->Mv[i];
    snorm += s->search[i] * s->search[i];
  }
  snorm = sqrt(snorm);
  if (snorm < MINVAL) return 0;
  double scale = m->meaninertia * (nv > 1 ? nv : 1);
  double gtol = m->tolerance * m->ls_tolerance * snorm * scale;
  double c0, d1, d2, lo = 0, hi = -1, dlo, dhi = 0, alpha = 0, c;
  ls_eval(m, d, s, 0, g1, g2, &c0, &d1, &d2);
  if (g_debug > 0) fprintf(stderr, "    ls: d1(0) %.6e d2(0) %.6e gtol %.3e\n", d1, d2, gtol);
  if (d1 >= 0 || d2 <= 0) return 0;
  dlo = d1;
  alpha = -d1 / d2;
  double prev_step = 1e300;
  for (int it = 0; it < m->ls_iterations; it++) {
    ls_eval(m, d, s, alpha, g1, g2, &c, &d1, &d2);
    if (g_debug > 1) fprintf(stderr, "      ls it %d alpha %.9e cost-c0 %.6e d1 %.6e d2 %.6e lo %.6e hi %.6e\n", it, alpha, c - c0, d1, d2, lo, hi);
    if (fabs(d1) < gtol) return alpha;
    if (d1 < 0) { lo = alpha; dlo = d1; } else { hi = alpha; dhi = d1; }
    double next = d2 > 0 ? alpha - d1 / d2 : -1;
    if (hi < 0) { /* no upper bracket yet: accept Newton if it moves right, else expand */
      if (!(next > lo)) next = 2 * alpha + MINVAL;
    } else if (!(next > lo && next < hi && fabs(next - alpha) < 0.5 * prev_step)) {
      /* Newton left the bracket or is not contracting (kinks of the piecewise-quadratic cost): bisect */
      next = 0.5 * (lo + hi);
    }
    if (next == alpha) return alpha;
    prev_step = fabs(next - alpha);
    alpha = next;
  }
  (void)dlo; (void)dhi;
  return lo > 0 ? lo : alpha; /* not converged: the last point with negative slope still lowers the convex cost */
}

static void mul_M(const mjcData* d, int nv, const double* x, double* y) {
  for (int i = 0; i < nv; i++) { double s = 0; for (int j = 0; j < nv; j++) s += d->M[i * nv + j] * x[j]; y[i] = s; }
}

static double total_cost_at(const mjcModel* m, mjcData* d, Solver* s, const double* qacc) {
  int nv = s->nv;
  double save[MJC_MAXNV];
  memcpy(save, d->qacc, sizeof(double) * nv);
  memcpy(d->qacc, qacc, sizeof(double) * nv);
  mul_M(d, nv, d->qacc, s->Ma);
  for (int r = 0; r < s->nefc; r++) {
    double v = 0;
    for (int i = 0; i < nv; i++) v += d->efc_J[r][i] * d->qacc[i];
    s->jar[r] = v - d->efc_aref[r];
  }
  constraint_update(m, d, s, 0);
  memcpy(d->qacc, save, sizeof(double) * nv);
  return s->cost;
}

/* mj_fwdConstraint: warm start selection + Newton iterations */
static void fwd_constraint(const mjcModel* m, mjcData* d, Solver* s) {
  int nv = m->nv;
  s->nv = nv; s->nefc = d->nefc;
  d->solver_iter = 0;
  if (d->nefc == 0) {
    memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv);
    memset(d->qfrc_constraint, 0, sizeof(double) * nv);
    return;
  }
  /* warm start: keep qacc_warmstart unless qacc_smooth has lower cost */
  double cw = total_cost_at(m, d, s, d->qacc_warmstart);
  double cs = total_cost_at(m, d, s, d->qacc_smooth);
  memcpy(d->qacc, cw > cs ? d->qacc_smooth : d->qacc_warmstart, sizeof(double) * nv);
  mul_M(d, nv, d->qacc, s->Ma);
  for (int r = 0; r < s->nefc; r++) {
    double v = 0;
    for (int i = 0; i < nv; i++) v += d->efc_J[r][i] * d->qacc[i];
    s->jar[r] = v - d->efc_aref[r];
  }
  constraint_update(m, d, s, 1);
  double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  for (int it = 0; it < m->iterations; it++) {
    double gn = 0;
    for (int i = 0; i < nv; i++) gn += s->grad[i] * s->grad[i];
    d->solver_grad = scale * sqrt(gn);
    d->solver_cost = s->cost;
    if (scale * sqrt(gn) < m->tolerance) break;
    cholesky(s->Lh, s->H, nv);
    for (int i = 0; i < nv; i++) s->search[i] = -s->grad[i];
    chol_solve(s->Lh, s->search, nv);
    mul_M(d, nv, s->search, s->Mv);
    for (int r = 0; r < s->nefc; r++) {
      double v = 0;
      for (int i = 0; i < nv; i++) v += d->efc_J[r][i] * s->search[i];
      s->jv[r] = v;
    }
    double alpha = line_search(m, d, s);
    d->solver_iter = it + 1;
    if (g_debug < 0) g_newton_iters++;
    if (g_debug > 0) {
      double sg = 0; for (int i = 0; i < nv; i++) sg += s->search[i] * s->grad[i];
      fprintf(stderr, "  newton it %d cost %.12e |g| %.3e search.grad %.3e alpha %.6e\n", it, s->cost, scale * sqrt(gn), sg, alpha);
    }
    if (alpha == 0) break;
    double oldcost = s->cost;
    for (int i = 0; i < nv; i++) { d->qacc[i] += alpha * s->search[i]; s->Ma[i] += alpha * s->Mv[i]; }
    for (int r = 0; r < s->nefc; r++) s->jar[r] += alpha * s->jv[r];
    constraint_update(m, d, s, 1);
    if (g_debug > 0) fprintf(stderr, "      -> new cost %.12e improvement %.3e\n", s->cost, scale * (oldcost - s->cost));
    if (scale * (oldcost - s->cost) < m->tolerance) break;
  }
}

/* ------------------------------------------------------------------ mj_forward */
static void sensors(const mjcModel* m, mjcData* d) {
  for (int i = 0; i < m->nsensor; i++) {
    double* out = d->sensordata + m->sens_adr[i];
    int o = m->sens_obj[i];
    switch (m->sens_type[i]) {
      case MJC_SENS_FRAMEPOS: memcpy(out, d->site_xpos[o], 3 * sizeof(double)); break;
      case MJC_SENS_JOINTPOS: out[0] = d->qpos[m->jnt_qposadr[o]]; break;
      case MJC_SENS_FRAMEPOS_BODY: memcpy(out, d->xpos[o], 3 * sizeof(double)); break;
      case MJC_SENS_FRAMEZAXIS_BODY: out[0] = d->xmat[o][2]; out[1] = d->xmat[o][5]; out[2] = d->xmat[o][8]; break;
      case MJC_SENS_DISTANCE: { /* mjSENS_GEOMDIST with body1/body2: min over the geom pairs, initialised to the cutoff */
        double cut = m->sens_cutoff[i], best = cut;
        for (int g1 = 0; g1 < m->ngeom; g1++) {
          if (m->geom_body[g1] != o || m->geom_type[g1] != MJC_GEOM_BOX) continue;
          for (int g2 = 0; g2 < m->ngeom; g2++) {
            if (m->geom_body[g2] != m->sens_obj2[i] || m->geom_type[g2] != MJC_GEOM_BOX) continue;
            double dd = box_box_distance(d->geom_xpos[g1], d->geom_xmat[g1], m->geom_size[g1], d->geom_xpos[g2], d->geom_xmat[g2],
                                         m->geom_size[g2], cut);
            if (dd < best) best = dd;
          }
        }
        out[0] = fmin(fmax(best, -cut), cut); /* sensor post-processing: real-valued data is clipped to +-cutoff */
        break;
      }
    }
  }
}

static void forward(const mjcModel* m, mjcData* d, Solver* s) {
  int nv = m->nv;
  /* mj_fwdPosition */
  kinematics(m, d);
  mass_matrix(m, d);
  collision(m, d);
  make_constraint(m, d);
  make_impedance(m, d);
  sensors(m, d); /* mj_sensorPos: every sensor of the BASELINE tasks is a position-stage sensor */
  /* mj_fwdVelocity: passive forces (joint damping) and bias forces */
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i];
  rne_bias(m, d);
  /* mj_fwdActuation: position servos  force = kp*(clamp(ctrl) - q) - kv*qdot, clamped to forcerange */
  memset(d->qfrc_actuator, 0, sizeof(double) * nv);
  for (int a = 0; a < m->nu; a++) {
    double u = d->ctrl[a];
    if (m->act_ctrllimited[a]) u = fmin(fmax(u, m->act_ctrlrange[a][0]), m->act_ctrlrange[a][1]);
    int dof = m->act_dof[a];
    int j = m->dof_jnt[dof];
    double len = m->act_gear[a] * d->qpos[m->jnt_qposadr[j]], vel = m->act_gear[a] * d->qvel[dof];
    double f = m->act_kp[a] * u - m->act_kp[a] * len - m->act_kv[a] * vel;
    if (m->act_forcelimited[a]) f = fmin(fmax(f, m->act_forcerange[a][0]), m->act_forcerange[a][1]);
    d->actuator_force[a] = f;
    d->qfrc_actuator[dof] += m->act_gear[a] * f;
  }
  /* joint-level clamp of the total actuator force (jnt_actfrcrange, scalar joints) */
  for (int j = 0; j < m->njnt; j++)
    if (m->jnt_actfrclimited[j] && (m->jnt_type[j] == MJC_JNT_SLIDE || m->jnt_type[j] == MJC_JNT_HINGE)) {
      int da = m->jnt_dofadr[j];
      d->qfrc_actuator[da] = fmin(fmax(d->qfrc_actuator[da], m->jnt_actfrcrange[j][0]), m->jnt_actfrcrange[j][1]);
    }
  /* mj_fwdAcceleration */
  for (int i = 0; i < nv; i++) { d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i]; d->qacc_smooth[i] = d->qfrc_smooth[i]; }
  chol_solve(d->L, d->qacc_smooth, nv);
  /* mj_fwdConstraint */
  fwd_constraint(m, d, s);
}

/* mj_Euler (implicit joint damping) / mj_implicit (implicitfast), then mj_advance */
static void integrate(const mjcModel* m, mjcData* d) {
  int nv = m->nv;
  double h = m->timestep, qacc[MJC_MAXNV], A[MJC_MAXNV * MJC_MAXNV], LA[MJC_MAXNV * MJC_MAXNV];
  double diag[MJC_MAXNV];
  int any = 0;
  for (int i = 0; i < nv; i++) { diag[i] = m->dof_damping[i]; }
  if (m->integrator == MJC_INT_IMPLICITFAST) {
    /* qDeriv = d(qfrc_passive + qfrc_actuator)/dqvel (RNE derivative dropped): -damping, -kv through the joint
     * transmission (zero when the actuator force is clamped) */
    for (int a = 0; a < m->nu; a++) {
      double f = d->actuator_force[a];
      int clamped = m->act_forcelimited[a] && (f <= m->act_forcerange[a][0] || f >= m->act_forcerange[a][1]);
      if (!clamped) diag[m->act_dof[a]] += m->act_gear[a] * m->act_gear[a] * m->act_kv[a];
    }
    any = 1;
  } else {
    for (int i = 0; i < nv; i++) if (diag[i] > 0) any = 1;
  }
  if (!any) memcpy(qacc, d->qacc, sizeof(double) * nv);
  else {
    memcpy(A, d->M, sizeof(double) * nv * nv);
    for (int i = 0; i < nv; i++) { A[i * nv + i] += h * diag[i]; qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i]; }
    cholesky(LA, A, nv);
    chol_solve(LA, qacc, nv);
  }
  /* mj_advance: semi-implicit — velocity first, then position with the new velocity */
  for (int i = 0; i < nv; i++) d->qvel[i] += h * qacc[i];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == MJC_JNT_FREE) {
      for (int k = 0; k < 3; k++) d->qpos[qa + k] += h * d->qvel[da + k];
      double w[3] = {d->qvel[da + 3], d->qvel[da + 4], d->qvel[da + 5]};
      double ang = h * normalize3(w), dq[4], nq[4], sn = sin(0.5 * ang);
      dq[0] = cos(0.5 * ang); dq[1] = sn * w[0]; dq[2] = sn * w[1]; dq[3] = sn * w[2];
      quat_mul(nq, d->qpos + qa + 3, dq);
      quat_normalize(nq);
      memcpy(d->qpos + qa + 3, nq, sizeof nq);
    } else d->qpos[qa] += h * d->qvel[da];
  }
  memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * nv);
}

static double* g_stats = 0; /* debug hook: (N,H,4) = iterations, nefc, last checked scaled |grad|, cost */
void mjc_set_stats_buffer(double* p) { g_stats = p; }

int mjc_rollout(const mjcModel* m, const double* x0, int x0_batched, const double* controls, int N, int H,
                double* states, double* sensors_out, int nthread) {
  int nq = m->nq, nv = m->nv, nu = m->nu, ns = m->nsensordata, nx = nq + nv;
#ifdef _OPENMP
  if (nthread <= 0) nthread = omp_get_max_threads();
#else
  nthread = 1;
#endif
  int fail = 0;
#pragma omp parallel num_threads(nthread)
  {
    mjcData* d = (mjcData*)malloc(sizeof(mjcData));
    Solver* s = (Solver*)malloc(sizeof(Solver));
    if (!d || !s) {
#pragma omp atomic write
      fail = 1;
    } else {
#pragma omp for schedule(dynamic, 1)
      for (int n = 0; n < N; n++) {
        const double* x = x0 + (x0_batched ? (size_t)n * nx : 0);
        memcpy(d->qpos, x, sizeof(double) * nq);
        memcpy(d->qvel, x + nq, sizeof(double) * nv);
        memset(d->qacc_warmstart, 0, sizeof(double) * nv); /* mujoco.rollout without initial_warmstart */
        for (int t = 0; t < H; t++) {
          memcpy(d->ctrl, controls + ((size_t)n * H + t) * nu, sizeof(double) * nu);
          forward(m, d, s);
          if (g_debug < 0) g_steps++;
          if (g_stats) {
            double* st = g_stats + ((size_t)n * H + t) * 4;
            double gn = 0;
            for (int i = 0; i < nv; i++) gn += s->grad[i] * s->grad[i];
            st[0] = d->solver_iter; st[1] = d->nefc; st[2] = d->nefc ? sqrt(gn) / (m->meaninertia * nv) : 0; st[3] = d->ncon;
          }
          integrate(m, d);
          double* so = states + ((size_t)n * H + t) * nx;
          memcpy(so, d->qpos, sizeof(double) * nq);
          memcpy(so + nq, d->qvel, sizeof(double) * nv);
          if (sensors_out) memcpy(sensors_out + ((size_t)n * H + t) * ns, d->sensordata, sizeof(double) * ns);
        }
      }
    }
    free(d); free(s);
  }
  return fail;
}

/* Diagnostics for tests / studies: world poses of every geom at qpos (xpos: ngeom x 3, xmat: ngeom x 9). */
void mjc_geom_poses(const mjcModel* m, const double* qpos, double* xpos, double* xmat) {
  mjcData* d = (mjcData*)calloc(1, sizeof(mjcData));
  memcpy(d->qpos, qpos, sizeof(double) * m->nq);
  kinematics(m, d);
  memcpy(xpos, d->geom_xpos, sizeof(double) * 3 * m->ngeom);
  memcpy(xmat, d->geom_xmat, sizeof(double) * 9 * m->ngeom);
  free(d);
}

int mjc_forward_debug(const mjcModel* m, const double* qpos, const double* qvel, const double* ctrl, double* M,
                      double* qfrc_bias, double* qfrc_passive, double* qfrc_actuator, double* qacc_smooth, double* qacc,
                      double* qfrc_constraint, int* ncon_out, double* contact_dist, double* contact_frame,
                      double* contact_pos, int* solver_iter) {
  mjcData* d = (mjcData*)calloc(1, sizeof(mjcData));
  Solver* s = (Solver*)calloc(1, sizeof(Solver));
  int nv = m->nv;
  memcpy(d->qpos, qpos, sizeof(double) * m->nq);
  memcpy(d->qvel, qvel, sizeof(double) * nv);
  memcpy(d->ctrl, ctrl, sizeof(double) * m->nu);
  forward(m, d, s);
  if (M) memcpy(M, d->M, sizeof(double) * nv * nv);
  if (qfrc_bias) memcpy(qfrc_bias, d->qfrc_bias, sizeof(double) * nv);
  if (qfrc_passive) memcpy(qfrc_passive, d->qfrc_passive, sizeof(double) * nv);
  if (qfrc_actuator) memcpy(qfrc_actuator, d->qfrc_actuator, sizeof(double) * nv);
  if (qacc_smooth) memcpy(qacc_smooth, d->qacc_smooth, sizeof(double) * nv);
  if (qacc) memcpy(qacc, d->qacc, sizeof(double) * nv);
  if (qfrc_constraint) memcpy(qfrc_constraint, d->qfrc_constraint, sizeof(double) * nv);
  if (ncon_out) *ncon_out = d->ncon;
  for (int i = 0; i < d->ncon; i++) {
    if (contact_dist) contact_dist[i] = d->contact[i].dist;
    if (contact_frame) memcpy(contact_frame + 9 * i, d->contact[i].frame, 9 * sizeof(double));
    if (contact_pos) memcpy(contact_pos + 3 * i, d->contact[i].pos, 3 * sizeof(double));
  }
  if (solver_iter) *solver_iter = d->solver_iter;
  int nefc = d->nefc;
  free(d); free(s);
  return nefc;
}
