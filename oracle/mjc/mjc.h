/* mjc.h — ORACLE (test infrastructure, not product).
 *
 * Plain-C, double-precision CPU restatement of the MuJoCo 3.5.0 mj_step pipeline for exactly the
 * feature set the BASELINE tasks switch on.  MuJoCo itself is a
 * third-party dependency of the reference (pyproject.toml:33 `mujoco>=3.5.0,<3.6`, call sites
 * judo/utils/mj_rollout_backend.py:36,84) and is absent from /root/reference and from this image,
 * so this file restates its *published* algorithm (MuJoCo documentation, "Computation" chapter)
 * and is anchored on the reference's call sites.  PARITY UNPINNED for the dynamics: the reference
 * ships no golden vectors for rollouts and no MuJoCo binary is available to generate any.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product path (judo_b200/) never does.
 */
#ifndef MJC_H
#define MJC_H

#ifdef __cplusplus
extern "C" {
#endif

#define MJC_MAXBODY 24
#define MJC_MAXJNT 24
#define MJC_MAXNQ 32
#define MJC_MAXNV 24
#define MJC_MAXNU 24
#define MJC_MAXGEOM 80
#define MJC_MAXSITE 8
#define MJC_MAXSENSOR 32
#define MJC_MAXPAIR 2600
#define MJC_MAXEQ 4
#define MJC_MAXCON 96
#define MJC_MAXEFC 400

enum { MJC_JNT_FREE = 0, MJC_JNT_BALL = 1, MJC_JNT_SLIDE = 2, MJC_JNT_HINGE = 3 };
enum { MJC_GEOM_SPHERE = 2, MJC_GEOM_CAPSULE = 3, MJC_GEOM_CYLINDER = 5, MJC_GEOM_BOX = 6 };
enum { MJC_INT_EULER = 0, MJC_INT_IMPLICITFAST = 1 };
enum { MJC_CONE_PYRAMIDAL = 0, MJC_CONE_ELLIPTIC = 1 };
enum { MJC_SENS_FRAMEPOS = 0, MJC_SENS_JOINTPOS = 1, MJC_SENS_FRAMEPOS_BODY = 2, MJC_SENS_FRAMEZAXIS_BODY = 3, MJC_SENS_DISTANCE = 4 };

typedef struct mjcModel mjcModel;

/* Build a model from the flat (int, double) blob written by oracle/mjc.py:serialize_model. */
mjcModel* mjc_model_create(const int* ib, int ni, const double* db, int nd);
void mjc_model_free(mjcModel* m);
int mjc_nq(const mjcModel* m);
int mjc_nv(const mjcModel* m);
int mjc_nu(const mjcModel* m);
int mjc_nsensordata(const mjcModel* m);

/* The reference CPU path restated: MJRolloutBackend.rollout (mj_rollout_backend.py:45-88) ->
 * mujoco.rollout.Rollout.rollout: for each of N rollouts set the state, then for t<H:
 * ctrl = controls[n,t]; mj_step; record (qpos,qvel) and sensordata.
 *   x0        (N, nq+nv) if x0_batched else (nq+nv)
 *   controls  (N, H, nu)
 *   states    (N, H, nq+nv)   state AFTER applying controls[:,t]
 *   sensors   (N, H, nsensordata)   (may be NULL)
 *   nthread   OpenMP threads (<=0: all)
 * Returns 0 on success. */
int mjc_rollout(const mjcModel* m, const double* x0, int x0_batched, const double* controls, int N, int H,
                double* states, double* sensors, int nthread);

/* Diagnostics for tests: one mj_forward at (qpos,qvel,ctrl) with zero warmstart. Any output may be NULL.
 * M is dense nv*nv. Returns the number of constraint rows; ncon_out gets the number of contacts. */
int mjc_forward_debug(const mjcModel* m, const double* qpos, const double* qvel, const double* ctrl,
                      double* M, double* qfrc_bias, double* qfrc_passive, double* qfrc_actuator,
                      double* qacc_smooth, double* qacc, double* qfrc_constraint, int* ncon_out,
                      double* contact_dist, double* contact_frame, double* contact_pos, int* solver_iter);

/* World poses of every geom at qpos (xpos: ngeom x 3, xmat: ngeom x 9 row-major); for tests and studies. */
void mjc_geom_poses(const mjcModel* m, const double* qpos, double* xpos, double* xmat);

/* Signed distance between two boxes (centre p, 3x3 row-major frame m, half-sizes s), clipped above at cutoff: the routine
 * behind the restated mjSENS_GEOMDIST sensors; exported for tests. */
double mjc_box_box_distance(const double* p1, const double* m1, const double* s1, const double* p2, const double* m2,
                            const double* s2, double cutoff);

#ifdef __cplusplus
}
#endif
#endif
