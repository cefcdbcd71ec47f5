/* b200mpc.h — C ABI of the B200 rollout engine (libb200mpc.so).
 *
 * Drop-in boundary for judo's sampling-MPC plan step.  Plain pointers and sizes only; no
 * torch / numpy / MuJoCo types.  All host arrays are C-contiguous float64 unless stated.  Every function
 * returns 0 on success, non-zero on failure; b200mpc_last_error() gives the message.  Not re-entrant per
 * handle (the reference calls the backend under ControllerNode.lock, judo/app/dora/controller.py:131-140);
 * several handles per process are fine (one per GPU).
 *
 * Each entry point cites the reference interface it replaces (paths relative to /root/reference).
 */
#ifndef B200MPC_H
#define B200MPC_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200mpc_handle b200mpc_handle;

enum { B200MPC_TASK_CARTPOLE = 0, B200MPC_TASK_CYLINDER_PUSH = 1, B200MPC_TASK_LEAP_CUBE = 2, B200MPC_TASK_FR3_PICK = 3 };
enum { B200MPC_OPT_MPPI = 0, B200MPC_OPT_CEM = 1, B200MPC_OPT_PS = 2 };

/* Task dimensions as the reference's MjModel reports them (nq, nv, nu, nsensordata). */
typedef struct { int nq, nv, nu, nsensordata, n_cost_params; } b200mpc_dims;

/* ---- lifetime -------------------------------------------------------------------------------------------
 * Replaces: MJRolloutBackend.__init__ (judo/utils/mj_rollout_backend.py:21-36), which deep-copies an MjModel
 * per thread.  Here the model is a flat constant table (task_consts: doubles, layout per task documented in
 * judo_b200/consts.py) baked from the task's MJCF; num_rollouts sizes the device buffers. */
int b200mpc_create(b200mpc_handle** out, int task_id, const double* task_consts, size_t n_consts, int device,
                   int num_rollouts);
void b200mpc_destroy(b200mpc_handle* h);
const char* b200mpc_last_error(const b200mpc_handle* h); /* h may be NULL: error of the last failed create */
int b200mpc_get_dims(const b200mpc_handle* h, b200mpc_dims* out);

/* Replaces: RolloutBackend.update(num_threads) (judo/utils/rollout_backend.py:40-46; called from
 * judo/controller/controller.py:225-226 when num_rollouts changes). */
int b200mpc_update(b200mpc_handle* h, int num_rollouts);
int b200mpc_num_rollouts(const b200mpc_handle* h);

/* ---- contract A: drop-in rollout ---------------------------------------------------------------------------
 * Replaces: RolloutBackend.rollout / MJRolloutBackend.rollout (judo/utils/rollout_backend.py:20-38,
 * judo/utils/mj_rollout_backend.py:45-88).
 *   x0        (nq+nv) if !x0_batched else (N, nq+nv)
 *   controls  (N, H, nu)
 *   states    (N, H, nq+nv)  OUT: state after applying controls[:, t]
 *   sensors   (N, H, nsensordata) OUT, may be NULL
 * N must equal the handle's num_rollouts (the reference asserts the same, mj_rollout_backend.py:82). */
int b200mpc_rollout(b200mpc_handle* h, const double* x0, int x0_batched, const double* controls, int N, int H,
                    double* states, double* sensors);

/* ---- contract B: fused spline -> rollout -> per-step cost -----------------------------------------------------
 * Replaces the Python sequence make_spline + spline eval + rollout + Task.reward
 * (judo/controller/controller.py:261-285; rewards judo/tasks/{cartpole.py:42-78,cylinder_push.py:50-93,
 * leap_cube.py:63-88}).
 *   knots        (N, K, nu)   candidate knots, already clipped/denormalised (controller.py:253-258)
 *   basis        (H, K)       spline basis: controls[n,t,:] = sum_k basis[t,k] * knots[n,k,:] (interp1d is linear
 *                             in the knots; judo_b200/spline.py builds it)
 *   cost_params  task-specific weights (b200mpc_dims.n_cost_params doubles; see judo_b200/consts.py)
 *   cost_NH      (N, H) float32 OUT per-step cost, may be NULL
 *   reward_N     (N) OUT reward as Task.reward returns it (negative total / mean cost) */
int b200mpc_plan_costs(b200mpc_handle* h, const double* x0, const double* knots, int N, int K, const double* basis,
                       int H, const double* cost_params, float* cost_NH, double* reward_N);

/* Reward only, from given trajectories: replaces Task.reward(states, sensors, controls, metadata)
 * (judo/tasks/base.py:55-76) for the built-in tasks when the caller used contract A.
 *   states (N,H,nq+nv), controls (N,H,nu) -> reward_N (N) */
int b200mpc_reward(b200mpc_handle* h, const double* states, const double* controls, int N, int H, const double* cost_params,
                   double* reward_N);
/* Same with the sensor trajectories, for rewards that read them (judo/tasks/fr3_pick.py:248-252: distance sensors, grasp
 * site, end-effector axis).   sensors (N,H,nsensordata); NULL is accepted for the tasks whose reward ignores sensors. */
int b200mpc_reward_sensors(b200mpc_handle* h, const double* states, const double* sensors, const double* controls, int N, int H,
                           const double* cost_params, double* reward_N);

/* ---- optimizer updates ----------------------------------------------------------------------------------------
 * Replace: MPPI.update_nominal_knots (judo/optimizers/mppi.py:61-82),
 *          CrossEntropyMethod.update_nominal_knots (judo/optimizers/cem.py:76-92; ties: higher index first),
 *          PredictiveSampling.update_nominal_knots (judo/optimizers/ps.py:52-65; first maximum).
 *   knots (N,K,nu), rewards (N) -> nominal (K,nu); CEM also writes sigma (K,nu) = clip(std(elites), min, max). */
int b200mpc_update_mppi(b200mpc_handle* h, const double* knots, const double* rewards, int N, int K, double temperature,
                        double* nominal);
int b200mpc_update_cem(b200mpc_handle* h, const double* knots, const double* rewards, int N, int K, int num_elites,
                       double sigma_min, double sigma_max, double* nominal, double* sigma);
int b200mpc_update_ps(b200mpc_handle* h, const double* knots, const double* rewards, int N, int K, double* nominal);

/* ---- fused plan step (one H2D, three kernels, one D2H) ------------------------------------------------------------
 * Replaces one iteration of the while-loop body in Controller.update_action (controller.py:261-288) after
 * sampling/clipping.  opt_params: MPPI {temperature}; CEM {num_elites, sigma_min, sigma_max}; PS {}.
 *   nominal (K,nu) OUT; sigma (K,nu) OUT (CEM only, else may be NULL); reward_N (N) OUT, may be NULL;
 *   elite_idx (n_elite) OUT indices of the best rollouts in descending reward order, may be NULL (n_elite=0).
 * n_elite and CEM's num_elites may be up to 256 (the reference has no limit; up to 8 the whole step is one launch, beyond that the
 * update runs as separate reduction kernels). */
int b200mpc_plan_step(b200mpc_handle* h, const double* x0, const double* knots, int N, int K, const double* basis, int H,
                      const double* cost_params, int optimizer, const double* opt_params, double* nominal, double* sigma,
                      double* reward_N, int* elite_idx, int n_elite);

/* ---- fused plan step with ON-DEVICE sampling (perf mode) -------------------------------------------------------------
 * Replaces sample_control_knots + clip + the loop body (judo/optimizers/{mppi.py:38-59,cem.py:55-74,ps.py:29-50},
 * judo/controller/controller.py:252-288): candidates = clip(nominal + sigma * N(0,1), lo, hi), row 0 = the nominal, drawn inside
 * the rollout kernel with a counter-based Philox4x32-10 keyed by (seed, counter, global rollout index) — same distribution as the
 * reference, NOT NumPy's MT19937 stream (use b200mpc_plan_step with host-sampled knots for seed parity).
 *   nominal_in, sigma_in (K,nu); lo, hi (nu) clip range (+-inf allowed); index_offset: global index of rollout 0 of this call.
 *   OUT: nominal (K,nu), sigma (CEM), reward_N (N, may be NULL), elite_idx / elite_knots (n_elite best candidates, may be NULL),
 *        knots_out (N,K,nu) all generated candidates (may be NULL: they then never leave the GPU). */
int b200mpc_plan_step_sampled(b200mpc_handle* h, const double* x0, const double* nominal_in, const double* sigma_in,
                              const double* lo, const double* hi, int N, int K, const double* basis, int H,
                              const double* cost_params, int optimizer, const double* opt_params, unsigned long long seed,
                              unsigned long long counter, int index_offset, double* nominal, double* sigma, double* reward_N,
                              int* elite_idx, int n_elite, double* elite_knots, double* knots_out);

/* ---- Controller.update_action fast path: ONE call per optimisation iteration ----------------------------------------------------
 * Replaces the body of the while-loop in Controller.update_action (judo/controller/controller.py:250-293) plus update_traces
 * (:323-363) for the built-in optimizers and tasks, WITH the reference's host-side sampling:
 *   candidates = clip(nominal + sigma * np.random.randn(N-1, K, nu), lo, hi), row 0 = the nominal
 *   (judo/optimizers/mppi.py:38-59, cem.py:55-74, ps.py:29-50; clip controller.py:253-258)
 * where the normals come from NumPy's legacy global MT19937 stream, reproduced bit for bit from the generator's own state
 * (mt_key / mt_pos point INTO numpy's mt19937_state: 624 words + position; they are advanced exactly as np.random.randn would).
 * The generator's cached-gaussian flag is not reachable from outside numpy, so the caller keeps it consistent: it draws the first
 * one or two normals through numpy (head[], n_head) so that the flag is known to be clear, this call draws an EVEN number, and when one
 * more is needed the step is split: phase 1 samples, the caller draws that last normal through numpy (tail), phase 2 finishes.
 * Then: spline basis (H, K) for query times time + dt * arange(H) (controller.py:261-262,382-401), one H2D copy, the fused
 * rollout + cost + update kernel (+ the elites' trace sensors from the captured positions), results written by the kernel into
 * pinned host memory, trace segments assembled (controller.py:341-363). */
typedef struct {
  int N, K, H;
  int optimizer;              /* B200MPC_OPT_* */
  int spline_order;           /* 0 zero, 1 linear, 2 cubic */
  int n_elite;                /* best rollouts to list and trace, 0..8 */
  int phase;                  /* 0: whole step; 1: sample only; 2: finish a step started with phase 1 */
  int n_head;                 /* 0..2 normals already drawn by the caller (head[]) */
  int has_tail;               /* phase 2: `tail` is the last normal of the block */
  int n_trace_sensors;        /* framepos trace sensors per elite (0: no traces) */
  int speculate;              /* 1: while the GPU runs this step, draw the NEXT step's block of normals from a COPY of the generator state */
  int use_speculated;         /* 1: b200mpc_controller_speculation() just returned 1 for the (n - n_head) & ~1 normals after the head */
  double time, dt;
  double head[2], tail;
  const double* knot_times;   /* (K) */
  const double* x0;           /* (nq+nv) */
  const double* nominal;      /* (K, nu) nominal knots at knot_times */
  const double* sigma;        /* (K, nu) */
  const double* lo;           /* (nu) */
  const double* hi;           /* (nu) */
  const double* cost_params;
  const double* opt_params;   /* as b200mpc_plan_step */
  const int* trace_cols;      /* (3 * n_trace_sensors) sensordata columns of the trace sensors */
  unsigned int* mt_key;       /* numpy's mt19937_state.key (624 words) */
  int* mt_pos;                /* numpy's mt19937_state.pos */
  /* outputs (caller-owned; any may be NULL except nominal_out) */
  double* nominal_out;        /* (K, nu) */
  double* sigma_out;          /* (K, nu), CEM */
  double* rewards;            /* (N) */
  int* elite_idx;             /* (n_elite) */
  double* traces;             /* (n_elite * n_trace_sensors * (H-1), 2, 3) */
  double* basis_out;          /* (H, K) */
  double* knots_out;          /* (N, K, nu) the candidates (NULL: read them later with b200mpc_last_candidates) */
} b200mpc_step_request;
int b200mpc_controller_step(b200mpc_handle* h, b200mpc_step_request* req);
/* Speculative sampling (hides the sampling time of large blocks behind the GPU's rollout time): with req->speculate the call above draws,
 * while it waits for the GPU, the block of normals the NEXT step will need — from a private copy of the generator state, numpy's own
 * state is not touched.  The noise does not depend on the nominal (judo/optimizers/mppi.py:58), so if nobody drew from or re-seeded the
 * generator in between, the next step would draw exactly these values.  This function says whether that is the case: 1 when a block of
 * n normals is held that was drawn from exactly the state the generator is in NOW (624 key words and the position compare equal).
 * The block covers what the next step would ask THIS library to draw, i.e. everything after its head values: all n normals when the
 * generator's gaussian cache is empty after this step, (n - 1) & ~1 when this step ended with a tail drawn through numpy (the cache
 * then holds that pair's second value, which becomes the next head).  A caller therefore asks twice: before drawing head values with
 * the full block size, and after drawing them with (n - n_head) & ~1.  b200mpc_controller_step with use_speculated = 1 installs the
 * advanced state into the generator and uses the block; otherwise it is dropped and the step samples as usual.  The stream numpy's
 * users see is identical either way: a draw or re-seed in between changes the key / position (a lone cached gaussian being consumed
 * does not, but then the head draw moves the state and the second query fails). */
int b200mpc_controller_speculation(b200mpc_handle* h, const unsigned int* mt_key, const int* mt_pos, size_t n);
/* The candidates of the last b200mpc_controller_step, (N, K, nu), copied out of the pinned staging buffer. */
int b200mpc_last_candidates(b200mpc_handle* h, double* knots_out, int N, int K);
/* Host-only helpers behind the call above, exported for the parity tests (no GPU needed):
 *   b200mpc_legacy_normals : the next n_even legacy normals of the stream (cached-gaussian flag clear);
 *   b200mpc_spline_basis    : interp1d's (H, K) matrix, order 0 zero / 1 linear / 2 cubic. */
int b200mpc_legacy_normals(unsigned int* mt_key, int* mt_pos, double* out, size_t n_even);
int b200mpc_spline_basis(int order, const double* knot_times, int K, const double* query, int H, double* basis_HK);

/* ---- several GPUs behind ONE backend, in one process ------------------------------------------------------------------------------
 * The reference's Controller is one process with one rollout backend (judo/controller/controller.py:72-85); a group keeps that shape:
 * one object whose rollouts are sharded along N over `devices` (rollout 0, the un-noised nominal, on the first; remainders to the lowest
 * ranks), everything issued asynchronously from the calling thread.  Each entry mirrors its single-GPU counterpart above:
 *   b200mpc_group_rollout          = b200mpc_rollout (RolloutBackend.rollout, judo/utils/mj_rollout_backend.py:45-88)
 *   b200mpc_group_plan_step        = b200mpc_plan_step: every device runs the fused rollout+cost kernel on its slice, the slices' rewards
 *                                    go peer-to-peer to the first device, which holds all candidates and runs the optimizer update
 *                                    (one small exchange per plan step) and the elite selection
 *   b200mpc_group_controller_step  = b200mpc_controller_step (same request, same sampling protocol)
 * num_rollouts is the TOTAL over the group.  Errors: b200mpc_group_last_error (g may be NULL: last failed create). */
typedef struct b200mpc_group b200mpc_group;
int b200mpc_group_create(b200mpc_group** out, int task_id, const double* task_consts, size_t n_consts, const int* devices, int n_devices,
                         int num_rollouts);
void b200mpc_group_destroy(b200mpc_group* g);
const char* b200mpc_group_last_error(const b200mpc_group* g);
int b200mpc_group_size(const b200mpc_group* g);
b200mpc_handle* b200mpc_group_handle(b200mpc_group* g, int i); /* the i-th device's handle (owned by the group) */
int b200mpc_group_update(b200mpc_group* g, int num_rollouts);
int b200mpc_group_num_rollouts(const b200mpc_group* g);
int b200mpc_group_rollout(b200mpc_group* g, const double* x0, int x0_batched, const double* controls, int N, int H, double* states,
                          double* sensors);
int b200mpc_group_plan_step(b200mpc_group* g, const double* x0, const double* knots, int N, int K, const double* basis, int H,
                            const double* cost_params, int optimizer, const double* opt_params, double* nominal, double* sigma,
                            double* reward_N, int* elite_idx, int n_elite);
int b200mpc_group_controller_step(b200mpc_group* g, b200mpc_step_request* req);
int b200mpc_group_controller_speculation(b200mpc_group* g, const unsigned int* mt_key, const int* mt_pos, size_t n);
int b200mpc_group_last_candidates(b200mpc_group* g, double* knots_out, int N, int K);
int b200mpc_group_set_trace_capture(b200mpc_group* g, int enable);
int b200mpc_group_elite_traces(b200mpc_group* g, const int* global_rollout_idx, int n, int H, double* traces_out);
long long b200mpc_group_launch_count(const b200mpc_group* g);
long long b200mpc_group_contact_overflows(b200mpc_group* g);

/* ---- resident (device-pointer) API: inputs/outputs already in HBM, asynchronous on `stream` ------------------------
 * Used by bench.py's device-resident measurement and by the multi-GPU sharded plan step (judo_b200/dist.py), where
 * torch owns the allocations and NCCL moves the partials.  All pointers are device pointers; stream is a
 * cudaStream_t passed as void*.  Partials layout (doubles): MPPI  [beta, S, V[K*nu]];
 * top-k  k x [reward, global_index, knots[K*nu]]. */
int b200mpc_plan_costs_dev(b200mpc_handle* h, const double* d_x0, const double* d_knots, int N, int K,
                           const double* d_basis, int H, const double* d_cost_params, float* d_cost_NH,
                           double* d_reward_N, void* stream);
/* One-launch resident plan step: rollout + cost + optimizer update fused (judo_b200/csrc/epilogue.cuh).
 *   finalize=1: writes d_nominal (K*nu), d_sigma (CEM), d_elite (n_elite indices as doubles, best first).
 *   finalize=2: MPPI over the peer exchange (see b200mpc_exchange_*): writes d_nominal, identical on every rank.
 *   finalize=0: writes this rank's partial to d_rank_partial for the all_gather: MPPI [beta, S, V[K*nu]];
 *               CEM num_elites x [reward, global index, knots]; PS 1 x [reward, global index, knots]; indices are
 *               offset by index_offset.  opt_params is a HOST pointer (see b200mpc_plan_step). */
int b200mpc_plan_step_dev(b200mpc_handle* h, const double* d_x0, const double* d_knots, int N, int K, const double* d_basis,
                          int H, const double* d_cost_params, int optimizer, const double* opt_params, int finalize,
                          int index_offset, int n_elite, float* d_cost_NH, double* d_reward_N, double* d_nominal,
                          double* d_sigma, double* d_elite, double* d_rank_partial, void* stream);
int b200mpc_rollout_dev(b200mpc_handle* h, const double* d_x0, int x0_batched, const double* d_controls, int N, int H,
                        double* d_states, double* d_sensors, void* stream);
int b200mpc_mppi_partial_dev(b200mpc_handle* h, const double* d_knots, const double* d_rewards, int N, int KNU,
                             double temperature, double* d_partial, void* stream);
int b200mpc_mppi_combine_dev(b200mpc_handle* h, const double* d_partials, int n_partials, int KNU, double temperature,
                             double* d_nominal, void* stream);
int b200mpc_topk_partial_dev(b200mpc_handle* h, const double* d_knots, const double* d_rewards, int N, int KNU, int k,
                             int index_offset, int prefer_high_index, double* d_partial, void* stream);
int b200mpc_topk_combine_dev(b200mpc_handle* h, const double* d_partials, int n_partials, int KNU, int k,
                             int prefer_high_index, double sigma_min, double sigma_max, double* d_nominal,
                             double* d_sigma, double* d_elite_idx, void* stream);

/* Peer exchange for the multi-GPU fused MPPI step: every rank allocates a small exchange buffer, the 64-byte CUDA IPC handles are
 * exchanged out of band (torch.distributed), and with finalize=2 in b200mpc_plan_step_dev the LAST WARP of the rollout kernel writes
 * the rank's [beta, S, V] partial into every peer's buffer over NVLink, waits for the peers' flags and writes the final nominal
 * knots — the collective of SURVEY.md §8e happens inside the rollout kernel (no NCCL call, no extra launch on the data path). */
int b200mpc_exchange_create(b200mpc_handle* h, int world_size, int rank, unsigned char* ipc_handle_out_64B);
int b200mpc_exchange_open(b200mpc_handle* h, const unsigned char* all_handles_world_x_64B);
/* The same wiring for peers inside ONE process (handles sharing a GPU, or GPUs with peer access enabled): after exchange_create on every
 * handle, pass each one the world_size buffer pointers obtained with b200mpc_exchange_buffer. */
int b200mpc_exchange_buffer(b200mpc_handle* h, void** buffer_out);
int b200mpc_exchange_open_local(b200mpc_handle* h, void* const* peer_buffers_world);
/* Measurement helpers for the scaling bench.  align: a one-warp kernel on `stream` that raises this rank's align flag in every peer buffer
 * and waits for all of them (lines the GPUs up between the L2 flush and the timed step, moving no data).  stamps: %globaltimer (ns) of
 * the last finalize=2 step on this rank: [0] kernel entry, [1] partial published to the peers, [2] all peers' partials seen. */
int b200mpc_exchange_align_dev(b200mpc_handle* h, void* stream);
int b200mpc_exchange_stamps(b200mpc_handle* h, unsigned long long* out3);
/* Two more %globaltimer stamps (ns): [0] exit of this rank's last line-up kernel (with stamps[0] above it brackets the launch gap between the
 * line-up and the timed rollout kernel), [1] the last instruction of the last finalize=2 step's exchange (final nominal written). */
int b200mpc_exchange_align_stamp(b200mpc_handle* h, unsigned long long* out2);

/* Measurement helper for bench.py's issue-bound roofline: DFMA warp instructions per second this GPU sustains with every SM full of
 * independent chains (the path is bound by fp64 issue / dependent latency, not by HBM). */
int b200mpc_fp64_peak(int device, double* dfma_warp_inst_per_s);

/* Number of kernel launches issued through this handle since creation (bench.py's gpu_launches). */
long long b200mpc_launch_count(const b200mpc_handle* h);

/* ---- trace capture (warp-per-rollout tasks: leap_cube, fr3_pick) ----------------------------------------------------------------
 * Replaces the `self.sensors[elite]` lookup of Controller.update_traces (judo/controller/controller.py:323-363) on the fused path,
 * where no (N, H, nsensordata) array exists: with capture enabled the fused kernel keeps the task's "trace" framepos sensors
 * (b200mpc_trace_width doubles per step: leap_cube 5 sites x 3, fr3_pick trace_object + trace_grasp_site) of EVERY rollout of the last
 * plan step in HBM, and b200mpc_elite_traces copies the rows of the given rollouts out: (n, H, trace_width).  Without it the elite
 * rollouts would have to be simulated a second time, which for these serial ms-scale kernels doubles the plan latency. */
int b200mpc_set_trace_capture(b200mpc_handle* h, int enable);
int b200mpc_trace_width(const b200mpc_handle* h);
int b200mpc_elite_traces(b200mpc_handle* h, const int* rollout_idx, int n, int H, double* traces_out);

/* Number of rollout steps run THROUGH THIS HANDLE since its creation in which an articulated-body kernel found more contacts than its
 * per-step buffer holds (leap_cube 30, fr3_pick 48; MuJoCo itself grows its arena) and dropped the surplus.  0 means every rollout so far
 * used the full contact set; callers that need the guarantee check it after planning.  -1 on CUDA errors. */
long long b200mpc_contact_overflows(b200mpc_handle* h);

#ifdef __cplusplus
}
#endif
#endif
