// b200mpc.cu — C-ABI implementation (include/b200mpc.h): handle, device buffers, pinned staging, kernel dispatch.
// Built with: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC  (judo_b200/build.py)
#include "../../include/b200mpc.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"
#ifdef B200MPC_WITH_LEAP
#include "leap.cuh"
#endif
#include "fr3.cuh"

using namespace b2;

namespace b2host {  // host_glue.cpp
void mt19937_normals(uint32_t* key, int* pos, double* out, size_t n);
void assemble_candidates(const double* z, const double* nominal, const double* sigma, const double* lo, const double* hi, int N, int K, int nu,
                         double* knots);
int spline_basis(int order, const double* t, int K, const double* q, int H, double* B);
void trace_segments(const double* sens, int ne, int H, int ns, const int* cols, int nts, double* out);
}  // namespace b2host

static thread_local std::string g_create_error;

struct b200mpc_handle {
  int task = -1, device = 0, N = 0;
  b200mpc_dims dims{};
  std::string err;
  long long launches = 0;
  cudaStream_t stream = nullptr;
  CartpoleConsts cartpole{};
  CylinderPushConsts cyl{};
#ifdef B200MPC_WITH_LEAP
  LeapModel* leap = nullptr;  // device-resident constant table
#endif
  Fr3Model* fr3 = nullptr;
  // device buffers (grown on demand)
  void* d_in = nullptr; size_t d_in_bytes = 0;      // packed inputs [x0 | basis | params | knots/controls]
  void* d_out = nullptr; size_t d_out_bytes = 0;    // packed small outputs [nominal | sigma | elite | reward_N]
  void* d_big = nullptr; size_t d_big_bytes = 0;    // states / sensors / cost matrix
  void* d_part = nullptr; size_t d_part_bytes = 0;  // reduction partials
  void* d_work = nullptr; size_t d_work_bytes = 0;  // per-rollout scratch (leap)
  // trace capture (warp-per-rollout tasks): the fused kernel keeps the trace sensors of every rollout of the LAST plan step
  bool trace_capture = false; void* d_trace = nullptr; size_t d_trace_bytes = 0; int trace_N = 0, trace_H = 0;
  void* h_in = nullptr; size_t h_in_bytes = 0;      // pinned staging
  void* h_out = nullptr; size_t h_out_bytes = 0;
  int zero_copy = 2;           // plan_step: bit0 = kernel READS the pinned staging buffer, bit1 = kernel WRITES results to pinned memory
  bool zero_copy_now = false;  // set for the duration of a zero-copy plan_step
  // peer exchange (multi-GPU fused MPPI): local buffer + peers' buffers opened through CUDA IPC
  void* xchg = nullptr; void* xchg_peer[8] = {nullptr}; int xchg_world = 0, xchg_rank = 0; unsigned long long xchg_epoch = 0, xchg_align_epoch = 0; bool xchg_local = false;
  unsigned long long* d_stamps = nullptr;  // %globaltimer stamps of the last finalize=2 step (b200mpc_exchange_stamps)
  double t_stage = 0, t_launch = 0, t_sync = 0, t_out = 0, t_spec = 0, t_gpu = 0; long long t_calls = 0; bool timing = false;  // B200MPC_TIMING=1
  cudaEvent_t tev0 = nullptr, tev1 = nullptr;  // (timing mode: GPU-side duration of a controller step, H2D to kernel end)
  // b200mpc_controller_step: normals of the current block, captured positions (N, H, nq) for the in-kernel elite traces, and where the
  // last step's candidates sit in the pinned staging buffer
  std::vector<double> zbuf, qtimes; bool step_sampled = false, step_tail_pending = false;
  void* d_traceq = nullptr; size_t d_traceq_bytes = 0;
  size_t cand_off = 0; int cand_N = 0, cand_K = 0;
  bool allow_resident = false;  // set by b200mpc_controller_step around its sampling stage (the device group assembles on the host)
  std::vector<double> trace_tmp;
  // speculative sampling: the next block of normals, the generator state it was drawn from and the state after it
  size_t znext_n = 0; bool znext_valid = false;
  // The speculated block lives in one of two PINNED host slots and is uploaded to the matching device slot by the helper thread as soon
  // as it is drawn.  A step whose whole block was speculated then launches with the normals already resident: the kernel assembles
  // clip(nominal + sigma * z) itself (sampling.cuh, enabled == 2) and neither the host assembly nor the (N, K, nu) upload are on the
  // step's critical path.  Slots alternate, so the block a running kernel reads is never the one being refilled.
  double* h_z[2] = {nullptr, nullptr}; size_t h_z_bytes[2] = {0, 0};
  double* d_z[2] = {nullptr, nullptr}; size_t d_z_bytes[2] = {0, 0};
  cudaStream_t z_stream = nullptr; cudaEvent_t z_ev[2] = {nullptr, nullptr};
  bool z_uploaded[2] = {false, false}; int z_slot = 0 /* slot of znext */, z_cur = -1 /* slot this step's kernel reads, -1: none */;
  bool z_resident = false, z_resident_ok = true;
  // candidates of a device-assembled step are rebuilt on the host only if somebody asks (b200mpc_last_candidates)
  bool cand_lazy = false; const double* cand_z = nullptr; std::vector<double> cand_par;
  uint32_t snap_key[624]; int snap_pos = 0; uint32_t adv_key[624]; int adv_pos = 0;
  // the speculative block is drawn by a helper thread while the GPU runs AND while the caller goes on (copy-out, Python glue, the next
  // call's staging): at C2 drawing 16 K normals takes twice as long as the GPU step.  Readers of znext / adv_* join first (spec_join).
  std::thread spec_thread; std::mutex spec_mu; std::condition_variable spec_cv; bool spec_pending = false, spec_busy = false, spec_quit = false;
  size_t spec_cnt = 0;
};

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                 \
      return 1;                                                                                    \
    }                                                                                              \
  } while (0)

static int fail(b200mpc_handle* h, const std::string& msg) { h->err = msg; return 1; }
// every rank's exchange buffer is mapped (b200mpc_exchange_open succeeded): only then may a step use the in-kernel exchange
static bool exchange_ready(const b200mpc_handle* h) {
  if (h->xchg_world < 1 || !h->xchg) return false;
  for (int g = 0; g < h->xchg_world; g++) if (!h->xchg_peer[g]) return false;
  return true;
}


static int grow(b200mpc_handle* h, void** p, size_t* cur, size_t need, bool pinned) {
  if (need <= *cur) return 0;
  size_t cap = std::max(need, *cur * 2);
  if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); *p = nullptr; *cur = 0; }
  cudaError_t e = pinned ? cudaMallocHost(p, cap) : cudaMalloc(p, cap);
  if (e != cudaSuccess) { h->err = std::string("allocation failed: ") + cudaGetErrorString(e); return 1; }
  *cur = cap;
  return 0;
}

static size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

extern "C" const char* b200mpc_last_error(const b200mpc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int b200mpc_create(b200mpc_handle** out, int task_id, const double* consts, size_t n_consts, int device, int num_rollouts) {
  if (!out) { g_create_error = "out is NULL"; return 1; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e); return 1; }
  if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return 1; }
  if (num_rollouts <= 0) { g_create_error = "num_rollouts must be positive"; return 1; }
  b200mpc_handle* h = new b200mpc_handle();
  h->task = task_id; h->device = device; h->N = num_rollouts;
  auto bad = [&](const std::string& m) { g_create_error = m; delete h; return 1; };
  if (cudaSetDevice(device) != cudaSuccess) return bad("cudaSetDevice failed");
  if (task_id == B200MPC_TASK_CARTPOLE) {
    if (n_consts != sizeof(CartpoleConsts) / sizeof(double)) return bad("cartpole: wrong number of task constants");
    memcpy(&h->cartpole, consts, sizeof(CartpoleConsts));
    h->dims = {CartpoleTask::NQ, CartpoleTask::NV, CartpoleTask::NU, CartpoleTask::NS, CartpoleTask::NCOST};
  } else if (task_id == B200MPC_TASK_CYLINDER_PUSH) {
    if (n_consts != sizeof(CylinderPushConsts) / sizeof(double)) return bad("cylinder_push: wrong number of task constants");
    memcpy(&h->cyl, consts, sizeof(CylinderPushConsts));
    h->dims = {CylinderPushTask::NQ, CylinderPushTask::NV, CylinderPushTask::NU, CylinderPushTask::NS, CylinderPushTask::NCOST};
  } else if (task_id == B200MPC_TASK_LEAP_CUBE) {
#ifdef B200MPC_WITH_LEAP
    std::string msg;
    if (leap_create(&h->leap, consts, n_consts, &msg)) return bad("leap_cube: " + msg);
    h->dims = {LEAP_NQ, LEAP_NV, LEAP_NU, LEAP_NS, LEAP_NCOST};
#else
    return bad("leap_cube kernel not built into this library");
#endif
  } else if (task_id == B200MPC_TASK_FR3_PICK) {
    std::string msg;
    if (fr3_create(&h->fr3, consts, n_consts, &msg)) return bad("fr3_pick: " + msg);
    h->dims = {FR_NQ, FR_NV, FR_NU, FR_NS, FR_NCOST};
  } else return bad("unknown task id");
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return bad("cudaStreamCreate failed");
  if (const char* z = getenv("B200MPC_ZEROCOPY")) h->zero_copy = atoi(z);
  if (const char* z = getenv("B200MPC_Z_RESIDENT")) h->z_resident_ok = atoi(z) != 0;
  h->timing = getenv("B200MPC_TIMING") != nullptr;
  *out = h;
  return 0;
}

static void spec_stop(b200mpc_handle* h);

extern "C" void b200mpc_destroy(b200mpc_handle* h) {
  if (!h) return;
  if (h->timing && h->t_calls)
    fprintf(stderr, "b200mpc host timing over %lld calls (us) [plan_step: stage|launch|wait|copy-out; controller_step: sample|assemble|h2d+launch|wait (+ speculative sampling before the wait)]: %.1f %.1f %.1f %.1f (+ %.1f)\n", h->t_calls,
            h->t_stage / h->t_calls, h->t_launch / h->t_calls, h->t_sync / h->t_calls, h->t_out / h->t_calls, h->t_spec / h->t_calls);
  if (h->timing && h->t_calls && h->t_gpu > 0) fprintf(stderr, "b200mpc controller_step GPU side (H2D .. kernel end, CUDA events): %.1f us\n", h->t_gpu / h->t_calls);
  spec_stop(h);
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (int g = 0; g < 8; g++) if (h->xchg_peer[g] && h->xchg_peer[g] != h->xchg && !h->xchg_local) cudaIpcCloseMemHandle(h->xchg_peer[g]);
  cudaFree(h->xchg); cudaFree(h->d_stamps);
  for (int i = 0; i < 2; i++) { cudaFreeHost(h->h_z[i]); cudaFree(h->d_z[i]); if (h->z_ev[i]) cudaEventDestroy(h->z_ev[i]); }
  if (h->z_stream) cudaStreamDestroy(h->z_stream);
  cudaFree(h->d_in); cudaFree(h->d_out); cudaFree(h->d_big); cudaFree(h->d_part); cudaFree(h->d_work); cudaFree(h->d_trace); cudaFree(h->d_traceq);
  cudaFreeHost(h->h_in); cudaFreeHost(h->h_out);
#ifdef B200MPC_WITH_LEAP
  if (h->leap) { if (getenv("B200MPC_LEAP_PROF")) leap_prof_dump(); leap_destroy(h->leap); }
#endif
  if (h->fr3) fr3_destroy(h->fr3);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

// ---- peer exchange set-up -------------------------------------------------------------------------------------------------
extern "C" int b200mpc_exchange_create(b200mpc_handle* h, int world, int rank, unsigned char* ipc_handle_out /* 64 bytes */) {
  if (!h) return 1;
  if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail(h, "exchange: world must be 1..8 and 0 <= rank < world");
  CK(cudaSetDevice(h->device));
  // a second create on the same handle starts from scratch: mappings of the previous peers are closed, flags and slots are cleared
  // (stale epoch flags >= the restarted epoch would let a step combine partials that were never written)
  for (int g = 0; g < 8; g++) {
    if (h->xchg_peer[g] && h->xchg_peer[g] != h->xchg && !h->xchg_local) cudaIpcCloseMemHandle(h->xchg_peer[g]);
    h->xchg_peer[g] = nullptr;
  }
  if (!h->xchg) CK(cudaMalloc(&h->xchg, EP_XCHG_BYTES));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemset(h->xchg, 0, EP_XCHG_BYTES));
  h->xchg_world = world; h->xchg_rank = rank; h->xchg_epoch = 0; h->xchg_align_epoch = 0; h->xchg_local = false;
  if (!h->d_stamps) { CK(cudaMalloc(&h->d_stamps, 64)); CK(cudaMemset(h->d_stamps, 0, 64)); }
  cudaIpcMemHandle_t ih;
  CK(cudaIpcGetMemHandle(&ih, h->xchg));
  static_assert(sizeof(ih) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(ipc_handle_out, &ih, 64);
  return 0;
}

extern "C" int b200mpc_exchange_open(b200mpc_handle* h, const unsigned char* all_handles /* world x 64 bytes */) {
  if (!h || !h->xchg) return 1;
  CK(cudaSetDevice(h->device));
  for (int g = 0; g < h->xchg_world; g++) {
    if (g == h->xchg_rank) { h->xchg_peer[g] = h->xchg; continue; }
    cudaIpcMemHandle_t ih;
    memcpy(&ih, all_handles + 64 * g, 64);
    cudaError_t e = cudaIpcOpenMemHandle(&h->xchg_peer[g], ih, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {  // leave no half-open exchange behind: callers then stay on the all_gather path
      for (int q = 0; q < 8; q++) { if (h->xchg_peer[q] && h->xchg_peer[q] != h->xchg && !h->xchg_local) cudaIpcCloseMemHandle(h->xchg_peer[q]); h->xchg_peer[q] = nullptr; }
      return fail(h, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
    }
  }
  return 0;
}

// Peers that live in THIS process (handles on the same GPU, or on GPUs with peer access enabled): their exchange buffers are plain device
// pointers, no IPC handle can or needs to be opened.
extern "C" int b200mpc_exchange_buffer(b200mpc_handle* h, void** out) {
  if (!h || !out) return 1;
  if (!h->xchg) return fail(h, "peer exchange not set up (exchange_create)");
  *out = h->xchg;
  return 0;
}
extern "C" int b200mpc_exchange_open_local(b200mpc_handle* h, void* const* peer_buffers /* world pointers */) {
  if (!h || !h->xchg || !peer_buffers) return 1;
  for (int g = 0; g < 8; g++) { if (h->xchg_peer[g] && h->xchg_peer[g] != h->xchg && !h->xchg_local) cudaIpcCloseMemHandle(h->xchg_peer[g]); h->xchg_peer[g] = nullptr; }
  h->xchg_local = true;
  for (int g = 0; g < h->xchg_world; g++) {
    if (g != h->xchg_rank && !peer_buffers[g]) return fail(h, "exchange: NULL peer buffer");
    h->xchg_peer[g] = g == h->xchg_rank ? h->xchg : peer_buffers[g];
  }
  return 0;
}

extern "C" int b200mpc_exchange_align_dev(b200mpc_handle* h, void* stream) {
  if (!h) return 1;
  if (!exchange_ready(h)) return fail(h, "peer exchange not set up (exchange_create/open)");
  CK(cudaSetDevice(h->device));
  PlanEpilogue ep{};
  ep.world = h->xchg_world; ep.rank = h->xchg_rank; ep.epoch = ++h->xchg_align_epoch;
  for (int g = 0; g < h->xchg_world; g++) ep.peer[g] = (double*)h->xchg_peer[g];
  ep.stamps = h->d_stamps;
  exchange_align_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ep);
  CK(cudaGetLastError());
  return 0;
}
extern "C" int b200mpc_exchange_stamps(b200mpc_handle* h, unsigned long long* out3) {
  if (!h || !out3) return 1;
  if (!h->d_stamps) return fail(h, "peer exchange not set up (exchange_create/open)");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpy(out3, h->d_stamps, 24, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int b200mpc_exchange_align_stamp(b200mpc_handle* h, unsigned long long* out2) {
  if (!h || !out2) return 1;
  if (!h->d_stamps) return fail(h, "peer exchange not set up (exchange_create/open)");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpy(out2, h->d_stamps + 3, 16, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int b200mpc_get_dims(const b200mpc_handle* h, b200mpc_dims* out) { if (!h || !out) return 1; *out = h->dims; return 0; }
extern "C" int b200mpc_update(b200mpc_handle* h, int n) { if (!h) return 1; if (n <= 0) return fail(h, "num_rollouts must be positive"); h->N = n; return 0; }
extern "C" int b200mpc_num_rollouts(const b200mpc_handle* h) { return h ? h->N : -1; }
extern "C" long long b200mpc_launch_count(const b200mpc_handle* h) { return h ? h->launches : 0; }

static int trace_width(const b200mpc_handle* h) { return h->task == B200MPC_TASK_LEAP_CUBE ? LEAP_NTRACE : h->task == B200MPC_TASK_FR3_PICK ? FR_NTRACE : 0; }
extern "C" int b200mpc_set_trace_capture(b200mpc_handle* h, int enable) {
  if (!h) return 1;
  if (enable && trace_width(h) == 0) return fail(h, "trace capture exists for the warp-per-rollout tasks only (the others recompute their elite traces in microseconds)");
  h->trace_capture = enable != 0;
  if (!enable) h->trace_N = h->trace_H = 0;
  return 0;
}
extern "C" int b200mpc_trace_width(const b200mpc_handle* h) { return h ? trace_width(h) : 0; }
extern "C" int b200mpc_elite_traces(b200mpc_handle* h, const int* idx, int n, int H, double* out) {
  if (!h) return 1;
  if (!idx || !out || n <= 0) return fail(h, "NULL / empty argument");
  const int nt = trace_width(h);
  if (!h->trace_capture || h->trace_N == 0) return fail(h, "no captured traces: enable b200mpc_set_trace_capture and run a fused plan step first");
  if (H != h->trace_H) return fail(h, "H does not match the captured plan step");
  CK(cudaSetDevice(h->device));
  const size_t row = (size_t)H * nt * 8;
  for (int i = 0; i < n; i++) {
    if (idx[i] < 0 || idx[i] >= h->trace_N) return fail(h, "rollout index out of range");
    CK(cudaMemcpyAsync((char*)out + (size_t)i * row, (const char*)h->d_trace + (size_t)idx[i] * row, row, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
extern "C" long long b200mpc_contact_overflows(b200mpc_handle* h) {
  if (!h) return -1;
  unsigned long long v = 0;
  const void* src = nullptr;  // the counter sits behind the handle's device-resident model table
#ifdef B200MPC_WITH_LEAP
  if (h->leap) src = h->leap + 1;
#endif
  if (h->fr3) src = h->fr3 + 1;
  if (!src) return 0;  // thread-per-rollout tasks: at most one contact, nothing to overflow
  if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess ||
      cudaMemcpy(&v, src, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) { h->err = "b200mpc_contact_overflows: CUDA error"; return -1; }
  return (long long)v;
}

// ------------------------------------------------------------------ launch helpers
static int pick_threads(int N) {
  // latency-bound serial recurrences: spread warps over the 148 SMs first, then fill each SM
  if (N <= 148 * 32) return 32;
  if (N <= 148 * 64 * 4) return 64;
  return 128;
}

template <class Task>
static int launch_rollout(b200mpc_handle* h, const typename Task::Consts& c, const double* d_x0, int batched, const double* d_ctrl,
                          int N, int H, double* d_states, double* d_sensors, cudaStream_t st) {
  int thr = pick_threads(N), grid = (N + thr - 1) / thr;
  PlanEpilogue ep{};
  ep.optimizer = EP_NONE;
  SampleSpec nos{};
  rollout_kernel<Task, false, 1><<<grid, thr, 0, st>>>(c, d_x0, batched, d_ctrl, N, H, 0, nullptr, nullptr, d_states, d_sensors, nullptr, nullptr, ep, nos);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

template <class Task, int MAXK>
static int launch_costs_k(b200mpc_handle* h, const typename Task::Consts& c, const double* d_x0, const double* d_knots, int N, int K,
                          const double* d_basis, int H, const double* d_params, float* d_cost, double* d_reward, const PlanEpilogue& ep,
                          const SampleSpec& smp, cudaStream_t st) {
  int thr = pick_threads(N), grid = (N + thr - 1) / thr;
  size_t smem = rollout_cost_smem<Task>(thr, H, K, d_cost != nullptr);
  auto kern = rollout_kernel<Task, true, MAXK>;
  if (smem > 48 * 1024) {
    if (smem > 227 * 1024) return fail(h, "horizon/knots too large for the shared-memory tile");
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  kern<<<grid, thr, smem, st>>>(c, d_x0, 0, d_knots, N, H, K, d_basis, d_params, nullptr, nullptr, d_cost, d_reward, ep, smp);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}
template <class Task>
static int launch_costs(b200mpc_handle* h, const typename Task::Consts& c, const double* d_x0, const double* d_knots, int N, int K,
                        const double* d_basis, int H, const double* d_params, float* d_cost, double* d_reward, const PlanEpilogue& ep,
                        const SampleSpec& smp, cudaStream_t st) {
  if (K <= 4) return launch_costs_k<Task, 4>(h, c, d_x0, d_knots, N, K, d_basis, H, d_params, d_cost, d_reward, ep, smp, st);
  if (K <= 8) return launch_costs_k<Task, 8>(h, c, d_x0, d_knots, N, K, d_basis, H, d_params, d_cost, d_reward, ep, smp, st);
  if (K <= 12) return launch_costs_k<Task, 12>(h, c, d_x0, d_knots, N, K, d_basis, H, d_params, d_cost, d_reward, ep, smp, st);
  return fail(h, "num_nodes > 12 not supported (reference slider range is 3..12, optimizers/base.py:13)");
}

extern "C" int b200mpc_rollout_dev(b200mpc_handle* h, const double* d_x0, int batched, const double* d_ctrl, int N, int H,
                                   double* d_states, double* d_sensors, void* stream) {
  if (!h) return 1;
  if (N <= 0 || H <= 0) return fail(h, "N and H must be positive");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  switch (h->task) {
    case B200MPC_TASK_CARTPOLE: return launch_rollout<CartpoleTask>(h, h->cartpole, d_x0, batched, d_ctrl, N, H, d_states, d_sensors, st);
    case B200MPC_TASK_CYLINDER_PUSH: return launch_rollout<CylinderPushTask>(h, h->cyl, d_x0, batched, d_ctrl, N, H, d_states, d_sensors, st);
#ifdef B200MPC_WITH_LEAP
    case B200MPC_TASK_LEAP_CUBE: {
      PlanEpilogue none{};
      none.optimizer = EP_NONE;
      if (leap_launch(h->leap, /*cost_mode=*/0, d_x0, batched, d_ctrl, N, H, 0, nullptr, nullptr, d_states, d_sensors, nullptr, nullptr, none, SampleSpec{}, st, &h->err)) return 1;
      h->launches++;
      return 0;
    }
#endif
    case B200MPC_TASK_FR3_PICK: {
      PlanEpilogue none{};
      none.optimizer = EP_NONE;
      if (fr3_launch(h->fr3, /*cost_mode=*/0, d_x0, batched, d_ctrl, N, H, 0, nullptr, nullptr, d_states, d_sensors, nullptr, nullptr, none, SampleSpec{}, st, &h->err)) return 1;
      h->launches++;
      return 0;
    }
  }
  return fail(h, "task not supported");
}

// (N, H, nt) device buffer for the trace capture of this launch, or NULL when capture is off
static int trace_buffer(b200mpc_handle* h, int N, int H, int nt, cudaStream_t st, double** out) {
  *out = nullptr;
  h->trace_N = h->trace_H = 0;
  if (!h->trace_capture) return 0;
  const size_t need = (size_t)N * H * nt * 8;
  if (need > h->d_trace_bytes) { CK(cudaStreamSynchronize(st)); if (grow(h, &h->d_trace, &h->d_trace_bytes, need, false)) return 1; }
  *out = (double*)h->d_trace;
  h->trace_N = N; h->trace_H = H;
  return 0;
}

static int plan_costs_ep(b200mpc_handle* h, const double* d_x0, const double* d_knots, int N, int K, const double* d_basis, int H,
                         const double* d_params, float* d_cost, double* d_reward, const PlanEpilogue& ep, cudaStream_t st,
                         const SampleSpec& smp = SampleSpec{}) {
  if (N <= 0 || H <= 0 || K <= 0) return fail(h, "N, H and K must be positive");
  CK(cudaSetDevice(h->device));
  switch (h->task) {
    case B200MPC_TASK_CARTPOLE: return launch_costs<CartpoleTask>(h, h->cartpole, d_x0, d_knots, N, K, d_basis, H, d_params, d_cost, d_reward, ep, smp, st);
    case B200MPC_TASK_CYLINDER_PUSH: return launch_costs<CylinderPushTask>(h, h->cyl, d_x0, d_knots, N, K, d_basis, H, d_params, d_cost, d_reward, ep, smp, st);
#ifdef B200MPC_WITH_LEAP
    case B200MPC_TASK_LEAP_CUBE: {
      double* d_trace = nullptr;
      if (trace_buffer(h, N, H, LEAP_NTRACE, st, &d_trace)) return 1;
      if (leap_launch(h->leap, /*cost_mode=*/1, d_x0, 0, d_knots, N, H, K, d_basis, d_params, nullptr, nullptr, d_cost, d_reward, ep, smp, st, &h->err, d_trace)) return 1;
      h->launches++;
      return 0;
    }
#endif
    case B200MPC_TASK_FR3_PICK: {
      double* d_trace = nullptr;
      if (trace_buffer(h, N, H, FR_NTRACE, st, &d_trace)) return 1;
      if (fr3_launch(h->fr3, /*cost_mode=*/1, d_x0, 0, d_knots, N, H, K, d_basis, d_params, nullptr, nullptr, d_cost, d_reward, ep, smp, st, &h->err, d_trace)) return 1;
      h->launches++;
      return 0;
    }
  }
  return fail(h, "task not supported");
}

static int run_update(b200mpc_handle* h, int optimizer, const double* opt_params, const double* d_knots, const double* d_rewards, int N,
                      int KNU, double* d_nominal, double* d_sigma, double* d_elite, int n_elite, cudaStream_t st);
static int n_partials_for(int N);

// number of warp partials the fused kernel of this task produces for N rollouts
static int n_warp_partials(const b200mpc_handle* h, int N) {
#ifdef B200MPC_WITH_LEAP
  if (h->task == B200MPC_TASK_LEAP_CUBE) return leap_num_partials(N);
#endif
  if (h->task == B200MPC_TASK_FR3_PICK) return N;
  int thr = pick_threads(N);
  return ((N + thr - 1) / thr) * (thr / 32);
}

// Fill a PlanEpilogue and make sure its scratch (ticket + warp partials) exists.  Output pointers are the caller's.
static int make_epilogue(b200mpc_handle* h, int optimizer, const double* opt_params, int n_elite, int N, int KNU, int finalize,
                         int index_offset, double* d_nominal, double* d_sigma, double* d_elite, double* d_rank_partial,
                         cudaStream_t st, PlanEpilogue* ep) {
  memset(ep, 0, sizeof(*ep));
  ep->optimizer = optimizer;
  int k_cem = optimizer == B200MPC_OPT_CEM ? (int)opt_params[0] : 0;
  if (optimizer == B200MPC_OPT_CEM && k_cem <= 0) return fail(h, "num_elites must be positive");
  ep->k = std::max(n_elite, k_cem);
  ep->k_cem = k_cem;
  ep->n_trace = n_elite;
  if (ep->k > EP_MAXK) return fail(h, "fused epilogue supports at most 8 elites");
  ep->finalize = finalize;
  ep->index_offset = index_offset;
  if (optimizer == B200MPC_OPT_MPPI) ep->temperature = opt_params[0];
  if (optimizer == B200MPC_OPT_CEM) { ep->sigma_min = opt_params[1]; ep->sigma_max = opt_params[2]; }
  const int nw = n_warp_partials(h, N);
  size_t off_m = 16, bytes_m = optimizer == B200MPC_OPT_MPPI ? (size_t)nw * (2 + KNU) * 8 : 0;
  size_t off_t = off_m + bytes_m, bytes_t = (size_t)nw * (ep->k + 1) * 2 * 8;
  size_t need = off_t + bytes_t;
  if (need > h->d_part_bytes) {
    CK(cudaStreamSynchronize(st));
    if (grow(h, &h->d_part, &h->d_part_bytes, need, false)) return 1;
    CK(cudaMemsetAsync(h->d_part, 0, 16, st));  // the ticket
  }
  char* base = (char*)h->d_part;
  ep->ticket = (unsigned int*)base;
  ep->warp_mppi = (double*)(base + off_m);
  ep->warp_topk = (double*)(base + off_t);
  ep->nominal = d_nominal; ep->sigma = d_sigma; ep->elite = d_elite;
  ep->rank_mppi = d_rank_partial; ep->rank_topk = d_rank_partial;
  return 0;
}

extern "C" int b200mpc_plan_costs_dev(b200mpc_handle* h, const double* d_x0, const double* d_knots, int N, int K, const double* d_basis,
                                      int H, const double* d_params, float* d_cost, double* d_reward, void* stream) {
  if (!h) return 1;
  PlanEpilogue ep{};
  ep.optimizer = EP_NONE;
  return plan_costs_ep(h, d_x0, d_knots, N, K, d_basis, H, d_params, d_cost, d_reward, ep, (cudaStream_t)stream);
}

static int plan_step_impl(b200mpc_handle* h, const double* d_x0, const double* d_knots, int N, int K, const double* d_basis,
                          int H, const double* d_params, int optimizer, const double* opt_params, int finalize,
                          int index_offset, int n_elite, float* d_cost, double* d_reward, double* d_nominal, double* d_sigma,
                          double* d_elite, double* d_elite_knots, double* d_rank_partial, const SampleSpec& smp, void* stream,
                          double* d_trace_q = nullptr, double* d_elite_sens = nullptr) {
  if (!h) return 1;
  if (optimizer < 0 || optimizer > 2) return fail(h, "unknown optimizer");
  if (optimizer == B200MPC_OPT_MPPI && !(opt_params && opt_params[0] > 0)) return fail(h, "temperature must be positive");
  if (optimizer == B200MPC_OPT_CEM && !opt_params) return fail(h, "CEM needs {num_elites, sigma_min, sigma_max}");
  if (n_elite < 0 || n_elite > 256) return fail(h, "n_elite must be in 0..256");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int KNU = K * h->dims.nu;
  const int k_need = std::max(n_elite, optimizer == B200MPC_OPT_CEM ? (int)opt_params[0] : 0);
  if (h->task == B200MPC_TASK_LEAP_CUBE || h->task == B200MPC_TASK_FR3_PICK || (k_need > EP_MAXK && finalize != 2)) {
    // warp-per-rollout kernels (ms-scale steps), or more elites than the fused epilogue keeps in registers (8; the reference has no
    // limit on num_elites / max_num_traces): the update runs as separate reduction kernels
    PlanEpilogue none{};
    none.optimizer = EP_NONE;
    none.index_offset = index_offset;
    if (plan_costs_ep(h, d_x0, d_knots, N, K, d_basis, H, d_params, d_cost, d_reward, none, st, smp)) return 1;
    if (smp.enabled) d_knots = smp.knots_out;
    if (finalize) {
      if (run_update(h, optimizer, opt_params, d_knots, d_reward, N, KNU, d_nominal, d_sigma, nullptr, 0, st)) return 1;
      if (n_elite > 0 && d_elite) {
        int nb = n_partials_for(N);
        size_t need = 16 + (size_t)nb * n_elite * (2 + KNU) * 8 + al16((size_t)KNU * 8);  // 16: the ticket slot in front
        if (need > h->d_part_bytes) { CK(cudaStreamSynchronize(st)); if (grow(h, &h->d_part, &h->d_part_bytes, need, false)) return 1; CK(cudaMemsetAsync(h->d_part, 0, 16, st)); }
        double* part = (double*)h->d_part + 2;
        double* dummy = part + (size_t)nb * n_elite * (2 + KNU);
        topk_partial_kernel<<<nb, 256, 0, st>>>(d_knots, d_reward, N, KNU, n_elite, index_offset, 1, part);
        h->launches++;
        CK(cudaGetLastError());
        if (b200mpc_topk_combine_dev(h, part, nb, KNU, n_elite, 1, 0, 0, dummy, nullptr, d_elite, st)) return 1;
        if (d_elite_knots) {
          gather_rows_kernel<<<n_elite, 64, 0, st>>>(d_knots, d_elite, index_offset, KNU, d_elite_knots);
          h->launches++;
          CK(cudaGetLastError());
        }
      }
      return 0;
    }
    if (optimizer == B200MPC_OPT_MPPI) return b200mpc_mppi_partial_dev(h, d_knots, d_reward, N, KNU, opt_params[0], d_rank_partial, st);
    int k = optimizer == B200MPC_OPT_CEM ? (int)opt_params[0] : 1;
    return b200mpc_topk_partial_dev(h, d_knots, d_reward, N, KNU, k, index_offset, optimizer == B200MPC_OPT_CEM, d_rank_partial, st);
  }
  PlanEpilogue ep;
  if (make_epilogue(h, optimizer, opt_params, n_elite, N, KNU, finalize, index_offset, d_nominal, d_sigma, d_elite, d_rank_partial, st, &ep)) return 1;
  ep.elite_knots = d_elite_knots;
  ep.trace_q = d_trace_q; ep.elite_sens = d_elite_sens;
  if (finalize == 2) {
    if (!exchange_ready(h)) return fail(h, "peer exchange not set up (exchange_create/open)");
    const int kout = optimizer == B200MPC_OPT_CEM ? (int)opt_params[0] : 1;
    if ((optimizer == B200MPC_OPT_MPPI ? 2 + KNU : kout * (2 + KNU)) > EP_XCHG_STRIDE) return fail(h, "partial too large for the exchange slot");
    if (n_elite > 0) return fail(h, "elite lists are per rank: pass n_elite = 0 with the peer exchange");
    ep.world = h->xchg_world; ep.rank = h->xchg_rank; ep.epoch = ++h->xchg_epoch;
    ep.stamps = h->d_stamps;
    for (int g = 0; g < h->xchg_world; g++) ep.peer[g] = (double*)h->xchg_peer[g];
  }
  return plan_costs_ep(h, d_x0, d_knots, N, K, d_basis, H, d_params, d_cost, d_reward, ep, st, smp);
}

extern "C" int b200mpc_plan_step_dev(b200mpc_handle* h, const double* d_x0, const double* d_knots, int N, int K, const double* d_basis,
                                     int H, const double* d_params, int optimizer, const double* opt_params, int finalize,
                                     int index_offset, int n_elite, float* d_cost, double* d_reward, double* d_nominal, double* d_sigma,
                                     double* d_elite, double* d_rank_partial, void* stream) {
  return plan_step_impl(h, d_x0, d_knots, N, K, d_basis, H, d_params, optimizer, opt_params, finalize, index_offset, n_elite, d_cost,
                        d_reward, d_nominal, d_sigma, d_elite, nullptr, d_rank_partial, SampleSpec{}, stream);
}

// On-device sampling (judo_b200/csrc/sampling.cuh): candidates = clip(nominal + sigma * N(0,1)) generated inside the rollout kernel.
extern "C" int b200mpc_plan_step_sampled(b200mpc_handle* h, const double* x0, const double* nominal_in, const double* sigma_in,
                                         const double* lo, const double* hi, int N, int K, const double* basis, int H,
                                         const double* params, int optimizer, const double* opt_params, unsigned long long seed,
                                         unsigned long long counter, int index_offset, double* nominal, double* sigma,
                                         double* reward_N, int* elite_idx, int n_elite, double* elite_knots, double* knots_out) {
  if (!h) return 1;
  if (!x0 || !nominal_in || !sigma_in || !lo || !hi || !basis || !params || !nominal) return fail(h, "NULL argument");
  if (N <= 0 || H <= 0 || K <= 0) return fail(h, "N, H and K must be positive");
  if (optimizer < 0 || optimizer > 2) return fail(h, "unknown optimizer");
  if (optimizer == B200MPC_OPT_MPPI && !(opt_params && opt_params[0] > 0)) return fail(h, "temperature must be positive");
  if (optimizer == B200MPC_OPT_CEM && !opt_params) return fail(h, "CEM needs {num_elites, sigma_min, sigma_max}");
  if (n_elite < 0 || n_elite > 256) return fail(h, "n_elite must be in 0..256");
  if (optimizer == B200MPC_OPT_CEM && ((int)opt_params[0] > 256 || (int)opt_params[0] <= 0)) return fail(h, "num_elites must be in 1..256");
  CK(cudaSetDevice(h->device));
  const int nx = h->dims.nq + h->dims.nv, nu = h->dims.nu, np = h->dims.n_cost_params, KNU = K * nu;
  // staged inputs: [x0 | basis | params | nominal | sigma | lo | hi]  (~1-3 KB: the only bytes that cross PCIe on the way in)
  size_t o = 0, ox0 = o; o += al16((size_t)nx * 8);
  size_t ob = o; o += al16((size_t)H * K * 8);
  size_t op = o; o += al16((size_t)np * 8);
  size_t on = o; o += al16((size_t)KNU * 8);
  size_t os = o; o += al16((size_t)KNU * 8);
  size_t ol = o; o += al16((size_t)nu * 8);
  size_t oh = o; o += al16((size_t)nu * 8);
  if (grow(h, &h->h_in, &h->h_in_bytes, o, true) || grow(h, &h->d_in, &h->d_in_bytes, o, false)) return 1;
  if (grow(h, &h->d_big, &h->d_big_bytes, (size_t)N * KNU * 8, false)) return 1;
  char* hp = (char*)h->h_in;
  memcpy(hp + ox0, x0, (size_t)nx * 8); memcpy(hp + ob, basis, (size_t)H * K * 8); memcpy(hp + op, params, (size_t)np * 8);
  memcpy(hp + on, nominal_in, (size_t)KNU * 8); memcpy(hp + os, sigma_in, (size_t)KNU * 8);
  memcpy(hp + ol, lo, (size_t)nu * 8); memcpy(hp + oh, hi, (size_t)nu * 8);
  CK(cudaMemcpyAsync(h->d_in, h->h_in, o, cudaMemcpyHostToDevice, h->stream));
  // outputs (pinned, written by the kernel): [nominal | sigma | elite idx | elite knots | reward]
  // (the fused epilogue lists max(n_elite, num_elites) rollouts: CEM's elites are part of the list even when fewer are asked for)
  const int kslots = std::max(std::max(n_elite, optimizer == B200MPC_OPT_CEM ? (int)opt_params[0] : 0), 1);
  size_t o_nom = 0, o_sig = al16((size_t)KNU * 8), o_el = o_sig + al16((size_t)KNU * 8), o_ek = o_el + al16((size_t)kslots * 8);
  size_t o_rw = o_ek + al16((size_t)kslots * KNU * 8), out_bytes = o_rw + (size_t)N * 8;
  if (grow(h, &h->h_out, &h->h_out_bytes, out_bytes, true)) return 1;
  void* dout_v = nullptr;
  CK(cudaHostGetDevicePointer(&dout_v, h->h_out, 0));
  char* din = (char*)h->d_in; char* dout = (char*)dout_v;
  SampleSpec smp{};
  smp.enabled = 1; smp.seed = seed; smp.counter = counter;
  smp.nominal = (double*)(din + on); smp.sigma = (double*)(din + os); smp.lo = (double*)(din + ol); smp.hi = (double*)(din + oh);
  smp.knots_out = (double*)h->d_big;
  if (plan_step_impl(h, (double*)(din + ox0), nullptr, N, K, (double*)(din + ob), H, (double*)(din + op), optimizer, opt_params, 1, index_offset,
                     n_elite, nullptr, (double*)(dout + o_rw), (double*)(dout + o_nom), (double*)(dout + o_sig), (double*)(dout + o_el),
                     (elite_knots && n_elite > 0) ? (double*)(dout + o_ek) : nullptr, nullptr, smp, h->stream)) return 1;
  if (knots_out) CK(cudaMemcpyAsync(knots_out, h->d_big, (size_t)N * KNU * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const char* ho = (const char*)h->h_out;
  memcpy(nominal, ho + o_nom, (size_t)KNU * 8);
  if (sigma && optimizer == B200MPC_OPT_CEM) memcpy(sigma, ho + o_sig, (size_t)KNU * 8);
  if (elite_idx) for (int i = 0; i < n_elite; i++) elite_idx[i] = (int)((const double*)(ho + o_el))[i];
  if (elite_knots && n_elite > 0) memcpy(elite_knots, ho + o_ek, (size_t)n_elite * KNU * 8);
  if (reward_N) memcpy(reward_N, ho + o_rw, (size_t)N * 8);
  return 0;
}

// ------------------------------------------------------------------ reductions (device-pointer API)
// (64 rollouts per block: at 512 the two blocks of the C4 update walked 128 dependent L2 loads per thread — 65 us for a 1 K x 64 reduction)
static int n_partials_for(int N) { return std::max(1, std::min(128, (N + 63) / 64)); }

extern "C" int b200mpc_mppi_partial_dev(b200mpc_handle* h, const double* d_knots, const double* d_rewards, int N, int KNU,
                                        double temperature, double* d_partial, void* stream) {
  if (!h) return 1;
  if (KNU > 256 * 8) return fail(h, "K*nu too large");
  CK(cudaSetDevice(h->device));
  // a single partial over all N rollouts of this call (grid=1) is what the multi-GPU path gathers
  int chunk = N;
  size_t smem = ((size_t)chunk + 256) * sizeof(double);
  if (smem > 200 * 1024) return fail(h, "mppi_partial_dev: N too large for one partial; split the call");
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(mppi_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mppi_partial_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(d_knots, d_rewards, N, KNU, temperature, d_partial);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

static int mppi_blocks(b200mpc_handle* h, const double* d_knots, const double* d_rewards, int N, int KNU, double temperature,
                       double* d_partial, int nb, cudaStream_t st) {
  int chunk = (N + nb - 1) / nb;
  size_t smem = ((size_t)chunk + 256) * sizeof(double);
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(mppi_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mppi_partial_kernel<<<nb, 256, smem, st>>>(d_knots, d_rewards, N, KNU, temperature, d_partial);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int b200mpc_mppi_combine_dev(b200mpc_handle* h, const double* d_partials, int np, int KNU, double temperature,
                                        double* d_nominal, void* stream) {
  if (!h) return 1;
  CK(cudaSetDevice(h->device));
  mppi_combine_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(d_partials, np, KNU, temperature, d_nominal);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int b200mpc_topk_partial_dev(b200mpc_handle* h, const double* d_knots, const double* d_rewards, int N, int KNU, int k,
                                        int index_offset, int prefer_high, double* d_partial, void* stream) {
  if (!h) return 1;
  if (k <= 0 || k > 256) return fail(h, "num_elites must be in 1..256");
  CK(cudaSetDevice(h->device));
  topk_partial_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_knots, d_rewards, N, KNU, k, index_offset, prefer_high, d_partial);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int b200mpc_topk_combine_dev(b200mpc_handle* h, const double* d_partials, int np, int KNU, int k, int prefer_high,
                                        double sigma_min, double sigma_max, double* d_nominal, double* d_sigma, double* d_elite,
                                        void* stream) {
  if (!h) return 1;
  if (k <= 0 || k > 256) return fail(h, "num_elites must be in 1..256");
  if ((long long)np * k >= (1 << 20)) return fail(h, "too many candidates");
  CK(cudaSetDevice(h->device));
  topk_combine_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_partials, np, KNU, k, prefer_high, sigma_min, sigma_max, d_nominal, d_sigma, d_elite);
  h->launches++;
  CK(cudaGetLastError());
  return 0;
}

// multi-block partials + combine on the handle's own scratch
static int run_update(b200mpc_handle* h, int optimizer, const double* opt_params, const double* d_knots, const double* d_rewards, int N,
                      int KNU, double* d_nominal, double* d_sigma, double* d_elite, int n_elite, cudaStream_t st) {
  int nb = n_partials_for(N);
  // these kernels use their own scratch (d_work) so that the fused path's ticket in d_part is never clobbered
  if (optimizer == B200MPC_OPT_MPPI) {
    if (grow(h, &h->d_work, &h->d_work_bytes, (size_t)nb * (2 + KNU) * 8, false)) return 1;
    if (mppi_blocks(h, d_knots, d_rewards, N, KNU, opt_params[0], (double*)h->d_work, nb, st)) return 1;
    if (b200mpc_mppi_combine_dev(h, (double*)h->d_work, nb, KNU, opt_params[0], d_nominal, st)) return 1;
    return 0;
  }
  int k = optimizer == B200MPC_OPT_CEM ? (int)opt_params[0] : 1;
  int prefer_high = optimizer == B200MPC_OPT_CEM ? 1 : 0;
  if (k <= 0 || k > 256) return fail(h, "num_elites must be in 1..256");
  if (grow(h, &h->d_work, &h->d_work_bytes, (size_t)nb * k * (2 + KNU) * 8, false)) return 1;
  topk_partial_kernel<<<nb, 256, 0, st>>>(d_knots, d_rewards, N, KNU, k, 0, prefer_high, (double*)h->d_work);
  h->launches++;
  CK(cudaGetLastError());
  double smin = optimizer == B200MPC_OPT_CEM ? opt_params[1] : 0, smax = optimizer == B200MPC_OPT_CEM ? opt_params[2] : 0;
  return b200mpc_topk_combine_dev(h, (double*)h->d_work, nb, KNU, k, prefer_high, smin, smax, d_nominal,
                                  optimizer == B200MPC_OPT_CEM ? d_sigma : nullptr, d_elite, st);
}

// ------------------------------------------------------------------ host-buffer API
extern "C" int b200mpc_rollout(b200mpc_handle* h, const double* x0, int batched, const double* controls, int N, int H, double* states,
                               double* sensors) {
  if (!h) return 1;
  if (!x0 || !controls || !states) return fail(h, "NULL argument");
  if (N != h->N) return fail(h, "controls batch size does not match num_rollouts (call update first)");
  if (H <= 0) return fail(h, "H must be positive");
  CK(cudaSetDevice(h->device));
  const int nx = h->dims.nq + h->dims.nv, nu = h->dims.nu, ns = h->dims.nsensordata;
  size_t bx = al16((size_t)(batched ? N : 1) * nx * 8), bc = (size_t)N * H * nu * 8;
  size_t bs = al16((size_t)N * H * nx * 8), be = sensors ? (size_t)N * H * ns * 8 : 0;
  if (grow(h, &h->d_in, &h->d_in_bytes, bx + bc, false) || grow(h, &h->d_big, &h->d_big_bytes, bs + be, false)) return 1;
  char* din = (char*)h->d_in; char* dbig = (char*)h->d_big;
  CK(cudaMemcpyAsync(din, x0, (size_t)(batched ? N : 1) * nx * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(din + bx, controls, bc, cudaMemcpyHostToDevice, h->stream));
  if (b200mpc_rollout_dev(h, (double*)din, batched, (double*)(din + bx), N, H, (double*)dbig, sensors ? (double*)(dbig + bs) : nullptr, h->stream)) return 1;
  CK(cudaMemcpyAsync(states, dbig, (size_t)N * H * nx * 8, cudaMemcpyDeviceToHost, h->stream));
  if (sensors) CK(cudaMemcpyAsync(sensors, dbig + bs, be, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// stage [x0 | basis | params | knots] through pinned memory with ONE host-to-device copy
static int stage_inputs(b200mpc_handle* h, const double* x0, const double* basis, int H, int K, const double* params, const double* knots,
                        int N, size_t* ox0, size_t* obasis, size_t* oparams, size_t* oknots) {
  const int nx = h->dims.nq + h->dims.nv, nu = h->dims.nu, np = h->dims.n_cost_params;
  size_t o = 0;
  *ox0 = o; o += al16((size_t)nx * 8);
  *obasis = o; o += al16((size_t)H * K * 8);
  *oparams = o; o += al16((size_t)np * 8);
  *oknots = o; o += (size_t)N * K * nu * 8;
  if (grow(h, &h->h_in, &h->h_in_bytes, o, true) || grow(h, &h->d_in, &h->d_in_bytes, o, false)) return 1;
  char* hp = (char*)h->h_in;
  memcpy(hp + *ox0, x0, (size_t)nx * 8);
  memcpy(hp + *obasis, basis, (size_t)H * K * 8);
  memcpy(hp + *oparams, params, (size_t)np * 8);
  memcpy(hp + *oknots, knots, (size_t)N * K * nu * 8);
  if (!h->zero_copy_now) CK(cudaMemcpyAsync(h->d_in, h->h_in, o, cudaMemcpyHostToDevice, h->stream));
  return 0;
}

extern "C" int b200mpc_plan_costs(b200mpc_handle* h, const double* x0, const double* knots, int N, int K, const double* basis, int H,
                                  const double* params, float* cost_NH, double* reward_N) {
  if (!h) return 1;
  if (!x0 || !knots || !basis || !params || !reward_N) return fail(h, "NULL argument");
  if (N <= 0 || H <= 0 || K <= 0) return fail(h, "N, H and K must be positive");
  CK(cudaSetDevice(h->device));
  size_t ox0, ob, op, ok;
  if (stage_inputs(h, x0, basis, H, K, params, knots, N, &ox0, &ob, &op, &ok)) return 1;
  size_t br = al16((size_t)N * 8), bc = cost_NH ? (size_t)N * H * 4 : 0;
  if (grow(h, &h->d_big, &h->d_big_bytes, br + bc, false)) return 1;
  char* din = (char*)h->d_in; char* dbig = (char*)h->d_big;
  if (b200mpc_plan_costs_dev(h, (double*)(din + ox0), (double*)(din + ok), N, K, (double*)(din + ob), H, (double*)(din + op),
                             cost_NH ? (float*)(dbig + br) : nullptr, (double*)dbig, h->stream)) return 1;
  CK(cudaMemcpyAsync(reward_N, dbig, (size_t)N * 8, cudaMemcpyDeviceToHost, h->stream));
  if (cost_NH) CK(cudaMemcpyAsync(cost_NH, dbig + br, bc, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int b200mpc_reward(b200mpc_handle* h, const double* states, const double* controls, int N, int H, const double* params,
                              double* reward_N) {
  return b200mpc_reward_sensors(h, states, nullptr, controls, N, H, params, reward_N);
}

extern "C" int b200mpc_reward_sensors(b200mpc_handle* h, const double* states, const double* sensors, const double* controls, int N, int H,
                                      const double* params, double* reward_N) {
  if (!h) return 1;
  if (!states || !controls || !params || !reward_N) return fail(h, "NULL argument");
  if (h->task == B200MPC_TASK_FR3_PICK && !sensors) return fail(h, "fr3_pick's reward reads the sensors (fr3_pick.py:248-252): use b200mpc_reward_sensors");
  if (N <= 0 || H <= 0) return fail(h, "N and H must be positive");
  CK(cudaSetDevice(h->device));
  const int nx = h->dims.nq + h->dims.nv, nu = h->dims.nu, np = h->dims.n_cost_params;
  size_t bs = al16((size_t)N * H * nx * 8), bc = al16((size_t)N * H * nu * 8), bp = al16((size_t)np * 8);
  const size_t be = sensors ? al16((size_t)N * H * h->dims.nsensordata * 8) : 0;
  if (grow(h, &h->d_in, &h->d_in_bytes, bs + bc + bp + be, false) || grow(h, &h->d_out, &h->d_out_bytes, (size_t)N * 8, false)) return 1;
  char* din = (char*)h->d_in;
  if (sensors) CK(cudaMemcpyAsync(din + bs + bc + bp, sensors, (size_t)N * H * h->dims.nsensordata * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(din, states, (size_t)N * H * nx * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(din + bs, controls, (size_t)N * H * nu * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(din + bs + bc, params, (size_t)np * 8, cudaMemcpyHostToDevice, h->stream));
  int thr = 128, grid = (N + thr - 1) / thr;
  switch (h->task) {
    case B200MPC_TASK_CARTPOLE:
      reward_kernel<CartpoleTask><<<grid, thr, 0, h->stream>>>((double*)din, (double*)(din + bs), N, H, (double*)(din + bs + bc), (double*)h->d_out);
      break;
    case B200MPC_TASK_CYLINDER_PUSH:
      reward_kernel<CylinderPushTask><<<grid, thr, 0, h->stream>>>((double*)din, (double*)(din + bs), N, H, (double*)(din + bs + bc), (double*)h->d_out);
      break;
#ifdef B200MPC_WITH_LEAP
    case B200MPC_TASK_LEAP_CUBE:
      if (leap_reward_launch(h->leap, (double*)din, N, H, (double*)(din + bs + bc), (double*)h->d_out, h->stream, &h->err)) return 1;
      break;
#endif
    case B200MPC_TASK_FR3_PICK:
      if (fr3_reward_launch((double*)din, (double*)(din + bs + bc + bp), N, H, (double*)(din + bs + bc), (double*)h->d_out, h->stream, &h->err)) return 1;
      break;
    default: return fail(h, "task not supported");
  }
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(reward_N, h->d_out, (size_t)N * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

static int host_update(b200mpc_handle* h, int optimizer, const double* opt_params, const double* knots, const double* rewards, int N, int K,
                       double* nominal, double* sigma) {
  if (!h) return 1;
  if (!knots || !rewards || !nominal) return fail(h, "NULL argument");
  if (N <= 0 || K <= 0) return fail(h, "N and K must be positive");
  CK(cudaSetDevice(h->device));
  const int KNU = K * h->dims.nu;
  size_t bk = al16((size_t)N * KNU * 8), br = al16((size_t)N * 8), bo = al16((size_t)KNU * 8);
  if (grow(h, &h->d_in, &h->d_in_bytes, bk + br, false) || grow(h, &h->d_out, &h->d_out_bytes, 2 * bo, false)) return 1;
  char* din = (char*)h->d_in; char* dout = (char*)h->d_out;
  CK(cudaMemcpyAsync(din, knots, (size_t)N * KNU * 8, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(din + bk, rewards, (size_t)N * 8, cudaMemcpyHostToDevice, h->stream));
  if (run_update(h, optimizer, opt_params, (double*)din, (double*)(din + bk), N, KNU, (double*)dout, (double*)(dout + bo), nullptr, 0, h->stream)) return 1;
  CK(cudaMemcpyAsync(nominal, dout, (size_t)KNU * 8, cudaMemcpyDeviceToHost, h->stream));
  if (sigma) CK(cudaMemcpyAsync(sigma, dout + bo, (size_t)KNU * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int b200mpc_update_mppi(b200mpc_handle* h, const double* knots, const double* rewards, int N, int K, double temperature, double* nominal) {
  if (h && !(temperature > 0)) return fail(h, "temperature must be positive");
  double p[1] = {temperature};
  return host_update(h, B200MPC_OPT_MPPI, p, knots, rewards, N, K, nominal, nullptr);
}
extern "C" int b200mpc_update_cem(b200mpc_handle* h, const double* knots, const double* rewards, int N, int K, int num_elites, double sigma_min,
                                  double sigma_max, double* nominal, double* sigma) {
  double p[3] = {(double)num_elites, sigma_min, sigma_max};
  return host_update(h, B200MPC_OPT_CEM, p, knots, rewards, N, K, nominal, sigma);
}
extern "C" int b200mpc_update_ps(b200mpc_handle* h, const double* knots, const double* rewards, int N, int K, double* nominal) {
  return host_update(h, B200MPC_OPT_PS, nullptr, knots, rewards, N, K, nominal, nullptr);
}

extern "C" int b200mpc_plan_step(b200mpc_handle* h, const double* x0, const double* knots, int N, int K, const double* basis, int H,
                                 const double* params, int optimizer, const double* opt_params, double* nominal, double* sigma,
                                 double* reward_N, int* elite_idx, int n_elite) {
  if (!h) return 1;
  if (!x0 || !knots || !basis || !params || !nominal) return fail(h, "NULL argument");
  if (N <= 0 || H <= 0 || K <= 0) return fail(h, "N, H and K must be positive");
  if (optimizer == B200MPC_OPT_MPPI && !(opt_params && opt_params[0] > 0)) return fail(h, "temperature must be positive");
  if (optimizer == B200MPC_OPT_CEM && !opt_params) return fail(h, "CEM needs {num_elites, sigma_min, sigma_max}");
  if (n_elite < 0 || n_elite > 256) return fail(h, "n_elite must be in 0..256");
  if (optimizer < 0 || optimizer > 2) return fail(h, "unknown optimizer");
  CK(cudaSetDevice(h->device));
  const int KNU = K * h->dims.nu;
  int k_cem = optimizer == B200MPC_OPT_CEM ? (int)opt_params[0] : 0;
  if (k_cem > 256) return fail(h, "num_elites must be in 1..256");
  // Zero-copy outputs (default): the fused kernel writes nominal/sigma/elite/rewards straight into pinned host memory, so the
  // step is ONE H2D copy + ONE launch.  Zero-copy inputs (bit0) make the kernel pull the candidates over PCIe itself (slower).
  const bool fusable = std::max(n_elite, k_cem) <= EP_MAXK;
  const bool zc_in = (h->zero_copy & 1) && fusable, zc = (h->zero_copy & 2) && fusable;  // zc: outputs
  h->zero_copy_now = zc_in;
  auto T0 = std::chrono::steady_clock::now();
  size_t ox0, ob, op, ok;
  int rc_stage = stage_inputs(h, x0, basis, H, K, params, knots, N, &ox0, &ob, &op, &ok);
  h->zero_copy_now = false;
  if (rc_stage) return 1;
  // packed outputs: [nominal KNU | sigma KNU | elite n_elite | reward N]
  // (the fused epilogue lists max(n_elite, num_elites) rollouts: CEM's elites are part of the list even when fewer are asked for)
  size_t o_nom = 0, o_sig = al16((size_t)KNU * 8), o_el = o_sig + al16((size_t)KNU * 8), o_rw = o_el + al16((size_t)std::max(std::max(n_elite, k_cem), 1) * 8);
  size_t out_bytes = o_rw + (size_t)N * 8;
  if (grow(h, &h->d_out, &h->d_out_bytes, out_bytes, false) || grow(h, &h->h_out, &h->h_out_bytes, out_bytes, true)) return 1;
  char* din = (char*)h->d_in; char* dout = (char*)h->d_out;
  if (zc_in) { void* p = nullptr; CK(cudaHostGetDevicePointer(&p, h->h_in, 0)); din = (char*)p; }
  if (zc) { void* p = nullptr; CK(cudaHostGetDevicePointer(&p, h->h_out, 0)); dout = (char*)p; }
  double* d_reward = (double*)(dout + o_rw);
  auto T1 = std::chrono::steady_clock::now();
  if (std::max(n_elite, k_cem) <= EP_MAXK) {
    // one launch: rollout + cost + optimizer update + elite list
    // multi-GPU handles with an open peer exchange: the MPPI update is GLOBAL (partials cross NVLink inside the kernel)
    const int kout_x = optimizer == B200MPC_OPT_CEM ? (int)opt_params[0] : 1;
    const bool fits_x = (optimizer == B200MPC_OPT_MPPI ? 2 + KNU : kout_x * (2 + KNU)) <= EP_XCHG_STRIDE;
    const int fin = (h->xchg_world > 1 && exchange_ready(h) && fits_x && h->task != B200MPC_TASK_LEAP_CUBE && h->task != B200MPC_TASK_FR3_PICK && n_elite == 0) ? 2 : 1;
    if (b200mpc_plan_step_dev(h, (double*)(din + ox0), (double*)(din + ok), N, K, (double*)(din + ob), H, (double*)(din + op), optimizer,
                              opt_params, fin, 0, n_elite, nullptr, d_reward, (double*)(dout + o_nom), (double*)(dout + o_sig),
                              (double*)(dout + o_el), nullptr, h->stream)) return 1;
  } else {
    // more than 8 elites: separate reduction kernels
    if (b200mpc_plan_costs_dev(h, (double*)(din + ox0), (double*)(din + ok), N, K, (double*)(din + ob), H, (double*)(din + op), nullptr,
                               d_reward, h->stream)) return 1;
    CK(cudaStreamSynchronize(h->stream));  // run_update may re-allocate the partial scratch the fused path uses
    if (run_update(h, optimizer, opt_params, (double*)(din + ok), d_reward, N, KNU, (double*)(dout + o_nom), (double*)(dout + o_sig),
                   nullptr, 0, h->stream)) return 1;
    if (n_elite > 0) {
      int nb = n_partials_for(N);
      size_t scratch = (size_t)nb * n_elite * (2 + KNU) * 8 + al16((size_t)KNU * 8);
      void* tmp = nullptr;
      CK(cudaStreamSynchronize(h->stream));
      CK(cudaMalloc(&tmp, scratch));
      double* part = (double*)tmp;
      double* dummy_nom = part + (size_t)nb * n_elite * (2 + KNU);
      topk_partial_kernel<<<nb, 256, 0, h->stream>>>((double*)(din + ok), d_reward, N, KNU, n_elite, 0, 1, part);
      h->launches++;
      int rc = b200mpc_topk_combine_dev(h, part, nb, KNU, n_elite, 1, 0, 0, dummy_nom, nullptr, (double*)(dout + o_el), h->stream);
      cudaStreamSynchronize(h->stream);
      cudaFree(tmp);
      if (rc) return 1;
    }
  }
  size_t copy_bytes = reward_N ? out_bytes : o_rw;
  if (!zc) CK(cudaMemcpyAsync(h->h_out, h->d_out, copy_bytes, cudaMemcpyDeviceToHost, h->stream));
  auto T2 = std::chrono::steady_clock::now();
  CK(cudaStreamSynchronize(h->stream));
  auto T3 = std::chrono::steady_clock::now();
  const char* ho = (const char*)h->h_out;
  memcpy(nominal, ho + o_nom, (size_t)KNU * 8);
  if (sigma && optimizer == B200MPC_OPT_CEM) memcpy(sigma, ho + o_sig, (size_t)KNU * 8);
  if (elite_idx) for (int i = 0; i < n_elite; i++) elite_idx[i] = (int)((const double*)(ho + o_el))[i];
  if (reward_N) memcpy(reward_N, ho + o_rw, (size_t)N * 8);
  if (h->xchg_world > 1 && exchange_ready(h) && nominal[0] != nominal[0]) {
    // the in-kernel exchange writes NaN when a peer's flag never arrived (bounded spin): surface it instead of poisoning the caller's spline
    bool finite_in = true;
    for (size_t i = 0; i < (size_t)N * KNU && finite_in; i++) finite_in = knots[i] == knots[i];
    if (finite_in) return fail(h, "peer exchange timed out: a rank did not run this plan step (nominal not updated)");
  }
  if (h->timing) {
    auto T4 = std::chrono::steady_clock::now();
    auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    h->t_stage += us(T0, T1); h->t_launch += us(T1, T2); h->t_sync += us(T2, T3); h->t_out += us(T3, T4); h->t_calls++;
  }
  return 0;
}

// ------------------------------------------------------------------ Controller.update_action fast path (include/b200mpc.h)
extern "C" int b200mpc_legacy_normals(unsigned int* mt_key, int* mt_pos, double* out, size_t n_even) {
  if (!mt_key || !mt_pos || (!out && n_even) || (n_even & 1) || *mt_pos < 0 || *mt_pos > 624) return 1;
  b2host::mt19937_normals(mt_key, mt_pos, out, n_even);
  return 0;
}
extern "C" int b200mpc_spline_basis(int order, const double* knot_times, int K, const double* query, int H, double* basis_HK) {
  if (!knot_times || !query || !basis_HK) return 1;
  return b2host::spline_basis(order, knot_times, K, query, H, basis_HK);
}

extern "C" int b200mpc_last_candidates(b200mpc_handle* h, double* knots_out, int N, int K) {
  if (!h) return 1;
  if (!knots_out || !h->h_in || h->cand_N == 0) return fail(h, "no candidates: run b200mpc_controller_step first");
  if (N != h->cand_N || K != h->cand_K) return fail(h, "N / K do not match the last b200mpc_controller_step");
  if (h->cand_lazy) {  // the step assembled its candidates on the device: same arithmetic, here, from the same normals
    const int nu = h->dims.nu, KNU = K * nu;
    const double* par = h->cand_par.data();
    b2host::assemble_candidates(h->cand_z, par, par + KNU, par + 2 * KNU, par + 2 * KNU + nu, N, K, nu, knots_out);
    return 0;
  }
  memcpy(knots_out, (const char*)h->h_in + h->cand_off, (size_t)N * K * h->dims.nu * 8);
  return 0;
}

static void spec_worker(b200mpc_handle* h) {
  std::unique_lock<std::mutex> lk(h->spec_mu);
  for (;;) {
    h->spec_cv.wait(lk, [h] { return h->spec_pending || h->spec_quit; });
    if (h->spec_quit) return;
    h->spec_pending = false;
    const size_t cnt = h->spec_cnt;
    lk.unlock();
    const int slot = h->z_slot;
    b2host::mt19937_normals(h->adv_key, &h->adv_pos, h->h_z[slot], cnt);  // (adv_* / the slot belong to the worker until spec_busy drops)
    bool up = false;
    if (h->z_resident_ok && h->d_z[slot] && h->z_stream && cudaSetDevice(h->device) == cudaSuccess)
      up = cudaMemcpyAsync(h->d_z[slot], h->h_z[slot], cnt * sizeof(double), cudaMemcpyHostToDevice, h->z_stream) == cudaSuccess &&
           cudaEventRecord(h->z_ev[slot], h->z_stream) == cudaSuccess;
    lk.lock();
    h->z_uploaded[slot] = up;
    h->znext_n = cnt;
    h->znext_valid = true;
    h->spec_busy = false;
    h->spec_cv.notify_all();
  }
}
// wait until the helper thread has finished the block it is drawing (no-op when it is idle)
static void spec_join(b200mpc_handle* h) {
  if (!h->spec_thread.joinable()) return;
  std::unique_lock<std::mutex> lk(h->spec_mu);
  h->spec_cv.wait(lk, [h] { return !h->spec_busy; });
}
static void spec_stop(b200mpc_handle* h) {
  if (!h->spec_thread.joinable()) return;
  { std::lock_guard<std::mutex> lk(h->spec_mu); h->spec_quit = true; }
  h->spec_cv.notify_all();
  h->spec_thread.join();
}

extern "C" int b200mpc_controller_speculation(b200mpc_handle* h, const unsigned int* mt_key, const int* mt_pos, size_t n) {
  if (!h || !mt_key || !mt_pos) return 0;
  spec_join(h);
  return h->znext_valid && h->znext_n == n && *mt_pos == h->snap_pos && memcmp(mt_key, h->snap_key, sizeof(h->snap_key)) == 0;
}


// The sampling stage of b200mpc_controller_step (shared with the multi-GPU group): fills h->zbuf with the n normals of this step's
// block.  Returns 0 when the block is complete, 2 after a phase-1 call (the caller still owes the tail), 1 on error.
static int step_sample(b200mpc_handle* h, b200mpc_step_request* rq, size_t n) {
  h->z_resident = false;
  if (rq->phase != 2) {
    h->z_cur = -1;
    h->zbuf.resize(n + 2);
    double* z = h->zbuf.data();
    for (int i = 0; i < rq->n_head; i++) z[i] = rq->head[i];
    const size_t rem = n - (size_t)rq->n_head, gen = rem & ~(size_t)1;
    spec_join(h);
    if (rq->use_speculated) {
      // the `gen` normals that follow the head values were drawn during the previous step's GPU time (step_speculate)
      if (!gen || !b200mpc_controller_speculation(h, rq->mt_key, rq->mt_pos, gen)) return fail(h, "the speculated block does not match the generator state");
      const bool warp_task = h->task == B200MPC_TASK_LEAP_CUBE || h->task == B200MPC_TASK_FR3_PICK;
      if (rq->n_head == 0 && gen == n && rq->phase == 0 && h->z_uploaded[h->z_slot] && h->z_resident_ok && !warp_task && h->allow_resident) {
        h->z_resident = true;   // the whole block is already on the device: no host copy, no host assembly (b200mpc_controller_step)
        h->z_cur = h->z_slot;
      } else {
        memcpy(z + rq->n_head, h->h_z[h->z_slot], gen * sizeof(double));
      }
      memcpy(rq->mt_key, h->adv_key, sizeof(h->adv_key));  // the generator now stands where drawing them would have left it
      *rq->mt_pos = h->adv_pos;
    } else if (gen) {
      if (*rq->mt_pos < 0 || *rq->mt_pos > 624) return fail(h, "generator position out of range");
      b2host::mt19937_normals(rq->mt_key, rq->mt_pos, z + rq->n_head, gen);
    }
    h->znext_valid = false;
    h->step_tail_pending = (rem & 1) != 0;
    h->step_sampled = true;
    if (rq->phase == 1) return 2;
    if (h->step_tail_pending) { h->step_sampled = false; return fail(h, "an odd number of normals is left: sample with phase 1, draw the last one, finish with phase 2"); }
  } else {
    if (!h->step_sampled || h->zbuf.size() != n + 2) return fail(h, "phase 2 without a matching phase 1");
    if (h->step_tail_pending) {
      if (!rq->has_tail) return fail(h, "the block is one normal short: pass it as tail");
      h->zbuf[n - 1] = rq->tail;
    }
  }
  h->step_sampled = false;
  return 0;
}

// While the GPU works: the normals the next step will ask this library for, from a COPY of the generator state
// (b200mpc_controller_speculation).  What the next step asks for follows from the generator's gaussian cache, which is known here:
//   this step ended without a tail -> cache empty -> the next block (n even) is drawn here in full, no head values;
//   this step drew a tail through numpy -> cache occupied -> the next step's head is that cached value (no state change), then
//   (n - 1) & ~1 normals from here, then its own tail.
static void step_speculate(b200mpc_handle* h, const b200mpc_step_request* rq, size_t n) {
  spec_join(h);
  h->znext_valid = false;
  const size_t cnt = !h->step_tail_pending ? ((n & 1) == 0 ? n : 0) : ((n - 1) & ~(size_t)1);
  if (rq->speculate && cnt >= 2 && rq->mt_key && rq->mt_pos && *rq->mt_pos >= 0 && *rq->mt_pos <= 624) {
    memcpy(h->snap_key, rq->mt_key, sizeof(h->snap_key));
    h->snap_pos = *rq->mt_pos;
    memcpy(h->adv_key, h->snap_key, sizeof(h->adv_key));
    h->adv_pos = h->snap_pos;
    const int slot = h->z_cur == 0 ? 1 : 0;  // never the slot the running kernel reads
    bool ok = grow(h, (void**)&h->h_z[slot], &h->h_z_bytes[slot], cnt * sizeof(double), true) == 0;
    if (ok && h->z_resident_ok) {
      if (!h->z_stream) ok = cudaStreamCreateWithFlags(&h->z_stream, cudaStreamNonBlocking) == cudaSuccess;
      for (int i = 0; ok && i < 2; i++) if (!h->z_ev[i]) ok = cudaEventCreateWithFlags(&h->z_ev[i], cudaEventDisableTiming) == cudaSuccess;
      if (ok) ok = grow(h, (void**)&h->d_z[slot], &h->d_z_bytes[slot], cnt * sizeof(double), false) == 0;
    }
    if (!ok) return;  // (no speculation this step: the next one samples in the call)
    h->z_slot = slot; h->z_uploaded[slot] = false;
    if (!h->spec_thread.joinable()) h->spec_thread = std::thread(spec_worker, h);
    { std::lock_guard<std::mutex> lk(h->spec_mu); h->spec_cnt = cnt; h->spec_pending = true; h->spec_busy = true; }
    h->spec_cv.notify_all();  // the block is drawn while the GPU runs and the caller carries on; whoever reads it joins first
  }
}

extern "C" int b200mpc_controller_step(b200mpc_handle* h, b200mpc_step_request* rq) {
  if (!h) return 1;
  if (!rq) return fail(h, "NULL request");
  const int N = rq->N, K = rq->K, H = rq->H, optimizer = rq->optimizer;
  if (!rq->x0 || !rq->nominal || !rq->sigma || !rq->lo || !rq->hi || !rq->cost_params || !rq->knot_times || !rq->nominal_out)
    return fail(h, "NULL argument");
  if (N <= 0 || H <= 0 || K <= 0) return fail(h, "N, H and K must be positive");
  if (optimizer < 0 || optimizer > 2) return fail(h, "unknown optimizer");
  if (optimizer == B200MPC_OPT_MPPI && !(rq->opt_params && rq->opt_params[0] > 0)) return fail(h, "temperature must be positive");
  if (optimizer == B200MPC_OPT_CEM && !rq->opt_params) return fail(h, "CEM needs {num_elites, sigma_min, sigma_max}");
  if (rq->n_elite < 0 || rq->n_elite > EP_MAXK) return fail(h, "n_elite must be in 0..8");
  if (optimizer == B200MPC_OPT_CEM && ((int)rq->opt_params[0] > EP_MAXK || (int)rq->opt_params[0] <= 0)) return fail(h, "num_elites must be in 1..8 on the fast path");
  if (rq->phase < 0 || rq->phase > 2) return fail(h, "phase must be 0, 1 or 2");
  const int nx = h->dims.nq + h->dims.nv, nu = h->dims.nu, np = h->dims.n_cost_params, ns = h->dims.nsensordata, KNU = K * nu;
  const size_t n = (size_t)(N - 1) * KNU;
  if (rq->n_head < 0 || rq->n_head > 2 || (size_t)rq->n_head > n) return fail(h, "n_head must be 0..2 and <= (N-1)*K*nu");
  if (n > (size_t)rq->n_head && (!rq->mt_key || !rq->mt_pos)) return fail(h, "NULL generator state");
  const bool warp_task = h->task == B200MPC_TASK_LEAP_CUBE || h->task == B200MPC_TASK_FR3_PICK;
  const int nts = rq->n_trace_sensors, ne = rq->n_elite;
  if (nts < 0 || (nts > 0 && (!rq->trace_cols || !rq->traces))) return fail(h, "trace sensors requested without trace_cols / traces");
  if (nts > 0 && warp_task && (!h->trace_capture || trace_width(h) != 3 * nts)) return fail(h, "enable b200mpc_set_trace_capture for the traces of this task");
  CK(cudaSetDevice(h->device));
  auto T0 = std::chrono::steady_clock::now();

  // ---- phase 0 / 1: draw the block of normals (head values first, then an even number straight from the generator state)
  {
    h->allow_resident = !rq->knots_out;
    const int rc = step_sample(h, rq, n);
    h->allow_resident = false;
    if (rc == 2) return 0;   // phase 1: sampled, waiting for the caller's tail normal
    if (rc) return 1;
  }
  auto T1 = std::chrono::steady_clock::now();
  const bool resident = h->z_resident;  // this step's normals are already on the device (drawn and uploaded during the previous step)

  // ---- stage [x0 | basis | params | knots] in pinned memory: the basis and the candidates are produced in place
  size_t o = 0;
  const size_t ox0 = o; o += al16((size_t)nx * 8);
  const size_t ob = o; o += al16((size_t)H * K * 8);
  const size_t op = o; o += al16((size_t)np * 8);
  // resident: [nominal | sigma | lo | hi] instead of the candidates
  const size_t ok = o; o += resident ? al16((size_t)(2 * KNU + 2 * nu) * 8) : (size_t)N * KNU * 8;
  if (grow(h, &h->h_in, &h->h_in_bytes, al16(o), true) || grow(h, &h->d_in, &h->d_in_bytes, al16(o), false)) return 1;
  if (resident && grow(h, &h->d_big, &h->d_big_bytes, (size_t)N * KNU * 8, false)) return 1;
  char* hp = (char*)h->h_in;
  memcpy(hp + ox0, rq->x0, (size_t)nx * 8);
  memcpy(hp + op, rq->cost_params, (size_t)np * 8);
  h->qtimes.resize(H);
  for (int i = 0; i < H; i++) h->qtimes[i] = rq->time + rq->dt * (double)i;  // self.time + task.dt * arange(H) (controller.py:261)
  if (b2host::spline_basis(rq->spline_order, rq->knot_times, K, h->qtimes.data(), H, (double*)(hp + ob))) return fail(h, "bad spline request (order / number of knots)");
  if (rq->basis_out) memcpy(rq->basis_out, hp + ob, (size_t)H * K * 8);
  SampleSpec smp{};
  if (resident) {
    double* par = (double*)(hp + ok);
    memcpy(par, rq->nominal, (size_t)KNU * 8); memcpy(par + KNU, rq->sigma, (size_t)KNU * 8);
    memcpy(par + 2 * KNU, rq->lo, (size_t)nu * 8); memcpy(par + 2 * KNU + nu, rq->hi, (size_t)nu * 8);
    h->cand_par.assign(par, par + 2 * KNU + 2 * nu);
    h->cand_lazy = true; h->cand_z = h->h_z[h->z_cur];
    const double* dpar = (const double*)((char*)h->d_in + ok);
    smp.enabled = 2; smp.z = h->d_z[h->z_cur]; smp.nominal = dpar; smp.sigma = dpar + KNU; smp.lo = dpar + 2 * KNU; smp.hi = dpar + 2 * KNU + nu;
    smp.knots_out = (double*)h->d_big;
  } else {
    b2host::assemble_candidates(h->zbuf.data(), rq->nominal, rq->sigma, rq->lo, rq->hi, N, K, nu, (double*)(hp + ok));
    h->cand_lazy = false;
    if (rq->knots_out) memcpy(rq->knots_out, hp + ok, (size_t)N * KNU * 8);
  }
  h->cand_off = ok; h->cand_N = N; h->cand_K = K;
  auto T2 = std::chrono::steady_clock::now();
  if (h->timing) { if (!h->tev0) { cudaEventCreate(&h->tev0); cudaEventCreate(&h->tev1); } cudaEventRecord(h->tev0, h->stream); }
  CK(cudaMemcpyAsync(h->d_in, h->h_in, o, cudaMemcpyHostToDevice, h->stream));
  if (resident) CK(cudaStreamWaitEvent(h->stream, h->z_ev[h->z_cur], 0));  // the upload of the block (finished long ago)

  // ---- outputs, written by the kernels straight into pinned host memory: [nominal | sigma | elite idx | elite sensors | reward]
  const bool kernel_traces = nts > 0 && ne > 0 && !warp_task;
  const int kslots = std::max(std::max(ne, optimizer == B200MPC_OPT_CEM ? (int)rq->opt_params[0] : 0), 1);  // the epilogue lists max(n_elite, num_elites)
  const size_t o_nom = 0, o_sig = al16((size_t)KNU * 8), o_el = o_sig + al16((size_t)KNU * 8), o_es = o_el + al16((size_t)kslots * 8);
  const size_t o_rw = o_es + (kernel_traces ? al16((size_t)ne * H * ns * 8) : 0), out_bytes = o_rw + (size_t)N * 8;
  if (grow(h, &h->h_out, &h->h_out_bytes, out_bytes, true)) return 1;
  if (kernel_traces && grow(h, &h->d_traceq, &h->d_traceq_bytes, (size_t)N * H * h->dims.nq * 8, false)) return 1;
  void* dout_v = nullptr;
  CK(cudaHostGetDevicePointer(&dout_v, h->h_out, 0));
  char* din = (char*)h->d_in; char* dout = (char*)dout_v;
  if (plan_step_impl(h, (double*)(din + ox0), resident ? (double*)h->d_big : (double*)(din + ok), N, K, (double*)(din + ob), H, (double*)(din + op), optimizer,
                     rq->opt_params, /*finalize=*/1, /*index_offset=*/0, ne, nullptr, (double*)(dout + o_rw), (double*)(dout + o_nom), (double*)(dout + o_sig),
                     (double*)(dout + o_el), nullptr, nullptr, smp, h->stream, kernel_traces ? (double*)h->d_traceq : nullptr,
                     kernel_traces ? (double*)(dout + o_es) : nullptr)) return 1;
  if (h->timing) cudaEventRecord(h->tev1, h->stream);
  auto T3 = std::chrono::steady_clock::now();
  step_speculate(h, rq, n);   // while the GPU works: the next step's normals, from a copy of the generator state
  auto T3b = std::chrono::steady_clock::now();
  CK(cudaStreamSynchronize(h->stream));
  if (h->timing) {  // B200MPC_TIMING=1: sample | assemble (basis, candidates) | H2D + launch | wait for the GPU
    auto T4 = std::chrono::steady_clock::now();
    auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    h->t_stage += us(T0, T1); h->t_launch += us(T1, T2); h->t_sync += us(T2, T3); h->t_out += us(T3b, T4); h->t_spec += us(T3, T3b); h->t_calls++;
    float gms = 0; if (cudaEventElapsedTime(&gms, h->tev0, h->tev1) == cudaSuccess) h->t_gpu += gms * 1e3;
  }
  const char* ho = (const char*)h->h_out;
  memcpy(rq->nominal_out, ho + o_nom, (size_t)KNU * 8);
  if (rq->sigma_out && optimizer == B200MPC_OPT_CEM) memcpy(rq->sigma_out, ho + o_sig, (size_t)KNU * 8);
  int el[EP_MAXK];
  for (int i = 0; i < ne; i++) { el[i] = (int)((const double*)(ho + o_el))[i]; if (rq->elite_idx) rq->elite_idx[i] = el[i]; }
  if (rq->rewards) memcpy(rq->rewards, ho + o_rw, (size_t)N * 8);
  if (nts > 0 && ne > 0) {
    if (kernel_traces) {
      b2host::trace_segments((const double*)(ho + o_es), ne, H, ns, rq->trace_cols, nts, rq->traces);
    } else {
      // warp-per-rollout tasks: the fused kernel captured every rollout's trace sensors (N, H, 3 nts); fetch the elites' rows
      const int nt = 3 * nts;
      const size_t row = (size_t)H * nt * 8;
      h->trace_tmp.resize((size_t)ne * H * nt);
      for (int i = 0; i < ne; i++) {
        if (el[i] < 0 || el[i] >= h->trace_N) return fail(h, "elite index out of range");
        CK(cudaMemcpyAsync((char*)h->trace_tmp.data() + (size_t)i * row, (const char*)h->d_trace + (size_t)el[i] * row, row, cudaMemcpyDeviceToHost, h->stream));
      }
      CK(cudaStreamSynchronize(h->stream));
      int cols[3 * 16];
      if (nt > 48) return fail(h, "too many trace sensors");
      for (int i = 0; i < nt; i++) cols[i] = i;
      b2host::trace_segments(h->trace_tmp.data(), ne, H, nt, cols, nts, rq->traces);
    }
  }
  return 0;
}

// ------------------------------------------------------------------ measurement helper: FP64 issue peak of this GPU
// Every resident warp runs 8 independent DFMA chains; enough blocks to fill all SMs.  bench.py divides the fp64 instruction count of the
// rollout kernels by this to state an issue-bound roofline (the path is latency / issue bound, not HBM bound).
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  if (x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 == 12345.678) out[0] = x0;  // keep the chains alive
}

extern "C" int b200mpc_fp64_peak(int device, double* dfma_warp_inst_per_s) {
  if (!dfma_warp_inst_per_s) return 1;
  if (cudaSetDevice(device) != cudaSuccess) return 1;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 1;
  double* d = nullptr;
  if (cudaMalloc(&d, 8) != cudaSuccess) return 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = prop.multiProcessorCount * 8, iters = 4096;
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    fp64_peak_kernel<<<blocks, 256>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return 1; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_inst = (double)blocks * 8 /*warps*/ * iters * 16 * 8;
    if (rep > 0 && ms > 0) best = std::max(best, warp_inst / (ms * 1e-3));
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  *dfma_warp_inst_per_s = best;
  return cudaGetLastError() != cudaSuccess;
}

// ====================================================================================================================================
// Several GPUs behind ONE backend, in one process (include/b200mpc.h: b200mpc_group_*).  The reference's Controller is a single process
// with one rollout backend (judo/controller/controller.py:72-85); the group keeps that shape: one object, rollouts sharded along N over
// its devices (rollout 0, the un-noised nominal, on the first), everything issued asynchronously from the calling thread.
//   contract A: every device rolls out its slice; states / sensors land in the caller's arrays.
//   plan step : every device runs the fused rollout+cost kernel on its slice; the slices' rewards are copied peer-to-peer into device
//               0, which also holds all candidates and runs the optimizer update (reduction kernels over all N) and the elite
//               selection; only nominal / sigma / elites / rewards come back.
struct b200mpc_group {
  std::vector<b200mpc_handle*> hs;
  std::vector<int> lo;            // shard r owns rollouts [lo[r], lo[r+1])
  int N = 0;
  std::string err;
  std::vector<cudaEvent_t> ev;    // per device: "this slice's rollouts are done"
  cudaEvent_t ev_copy = nullptr;  // device 0: all candidates uploaded
  cudaStream_t copy_stream = nullptr;
  void* d_knots_all = nullptr; size_t d_knots_all_bytes = 0;    // device 0
  void* d_reward_all = nullptr; size_t d_reward_all_bytes = 0;  // device 0
  void* h_knots_all = nullptr; size_t h_knots_all_bytes = 0;    // pinned: the whole candidate block
  int cand_N = 0, cand_K = 0;
  std::vector<double> tmp;
};
static thread_local std::string g_group_error;

static int gfail(b200mpc_group* g, const std::string& m) { g->err = m; return 1; }
#define GCK(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) { g->err = std::string(#call) + ": " + cudaGetErrorString(e_); return 1; } \
  } while (0)

static void group_shards(b200mpc_group* g, int N) {
  const int n = (int)g->hs.size(), base = N / n, rem = N % n;
  g->lo.assign(n + 1, 0);
  for (int r = 0; r < n; r++) g->lo[r + 1] = g->lo[r] + base + (r < rem ? 1 : 0);  // remainders go to the lowest ranks (judo_b200/dist.py:shard_range)
  g->N = N;
  for (int r = 0; r < n; r++) g->hs[r]->N = std::max(1, g->lo[r + 1] - g->lo[r]);
}

extern "C" const char* b200mpc_group_last_error(const b200mpc_group* g) { return g ? g->err.c_str() : g_group_error.c_str(); }

extern "C" int b200mpc_group_create(b200mpc_group** out, int task_id, const double* consts, size_t n_consts, const int* devices, int n_devices,
                                    int num_rollouts) {
  if (!out) { g_group_error = "out is NULL"; return 1; }
  *out = nullptr;
  if (!devices || n_devices < 1 || n_devices > 8) { g_group_error = "1..8 devices"; return 1; }
  if (num_rollouts <= 0) { g_group_error = "num_rollouts must be positive"; return 1; }
  for (int i = 0; i < n_devices; i++)
    for (int j = 0; j < i; j++) if (devices[i] == devices[j]) { g_group_error = "duplicate device"; return 1; }
  b200mpc_group* g = new b200mpc_group();
  for (int r = 0; r < n_devices; r++) {
    b200mpc_handle* h = nullptr;
    if (b200mpc_create(&h, task_id, consts, n_consts, devices[r], 1)) {
      g_group_error = std::string("device ") + std::to_string(devices[r]) + ": " + b200mpc_last_error(nullptr);
      for (auto* q : g->hs) b200mpc_destroy(q);
      delete g;
      return 1;
    }
    g->hs.push_back(h);
  }
  g->ev.resize(n_devices);
  bool ok = true;
  for (int r = 0; r < n_devices && ok; r++) ok = cudaSetDevice(devices[r]) == cudaSuccess && cudaEventCreateWithFlags(&g->ev[r], cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaSetDevice(devices[0]) == cudaSuccess && cudaEventCreateWithFlags(&g->ev_copy, cudaEventDisableTiming) == cudaSuccess &&
       cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  if (!ok) { g_group_error = "event / stream creation failed"; for (auto* q : g->hs) b200mpc_destroy(q); delete g; return 1; }
  // direct peer access where the hardware has it (NVLink / NVSwitch); cudaMemcpyPeerAsync stages through the host otherwise
  for (int r = 1; r < n_devices; r++) {
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, devices[0], devices[r]) == cudaSuccess && can) {
      cudaSetDevice(devices[0]);
      cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else cudaGetLastError();
    }
  }
  group_shards(g, num_rollouts);
  *out = g;
  return 0;
}

extern "C" void b200mpc_group_destroy(b200mpc_group* g) {
  if (!g) return;
  if (!g->hs.empty()) cudaSetDevice(g->hs[0]->device);
  if (g->copy_stream) { cudaStreamSynchronize(g->copy_stream); cudaStreamDestroy(g->copy_stream); }
  if (g->ev_copy) cudaEventDestroy(g->ev_copy);
  cudaFree(g->d_knots_all); cudaFree(g->d_reward_all); cudaFreeHost(g->h_knots_all);
  for (size_t r = 0; r < g->hs.size(); r++) { cudaSetDevice(g->hs[r]->device); if (g->ev[r]) cudaEventDestroy(g->ev[r]); }
  for (auto* h : g->hs) b200mpc_destroy(h);
  delete g;
}

extern "C" int b200mpc_group_size(const b200mpc_group* g) { return g ? (int)g->hs.size() : 0; }
extern "C" b200mpc_handle* b200mpc_group_handle(b200mpc_group* g, int i) { return (g && i >= 0 && i < (int)g->hs.size()) ? g->hs[i] : nullptr; }
extern "C" int b200mpc_group_num_rollouts(const b200mpc_group* g) { return g ? g->N : -1; }
extern "C" int b200mpc_group_update(b200mpc_group* g, int num_rollouts) {
  if (!g) return 1;
  if (num_rollouts <= 0) return gfail(g, "num_rollouts must be positive");
  group_shards(g, num_rollouts);
  return 0;
}
extern "C" long long b200mpc_group_launch_count(const b200mpc_group* g) {
  long long n = 0;
  if (g) for (auto* h : g->hs) n += h->launches;
  return n;
}
extern "C" long long b200mpc_group_contact_overflows(b200mpc_group* g) {
  if (!g) return -1;
  long long n = 0;
  for (auto* h : g->hs) { long long v = b200mpc_contact_overflows(h); if (v < 0) { g->err = h->err; return -1; } n += v; }
  return n;
}
extern "C" int b200mpc_group_set_trace_capture(b200mpc_group* g, int enable) {
  if (!g) return 1;
  for (auto* h : g->hs) if (b200mpc_set_trace_capture(h, enable)) return gfail(g, h->err);
  return 0;
}

// contract A over the group: RolloutBackend.rollout with the rollouts split across the devices
extern "C" int b200mpc_group_rollout(b200mpc_group* g, const double* x0, int batched, const double* controls, int N, int H, double* states,
                                     double* sensors) {
  if (!g) return 1;
  if (!x0 || !controls || !states) return gfail(g, "NULL argument");
  if (N != g->N) return gfail(g, "controls batch size does not match num_rollouts (call update first)");
  if (H <= 0) return gfail(g, "H must be positive");
  const int n = (int)g->hs.size();
  b200mpc_handle* h0 = g->hs[0];
  const int nx = h0->dims.nq + h0->dims.nv, nu = h0->dims.nu, ns = h0->dims.nsensordata;
  for (int r = 0; r < n; r++) {  // issue every device's upload + launch first, then the downloads: the devices run concurrently
    b200mpc_handle* h = g->hs[r];
    const int lo = g->lo[r], nr = g->lo[r + 1] - lo;
    if (nr <= 0) continue;
    GCK(cudaSetDevice(h->device));
    const size_t bx = al16((size_t)(batched ? nr : 1) * nx * 8), bc = (size_t)nr * H * nu * 8;
    const size_t bs = al16((size_t)nr * H * nx * 8), be = sensors ? (size_t)nr * H * ns * 8 : 0;
    if (grow(h, &h->d_in, &h->d_in_bytes, bx + bc, false) || grow(h, &h->d_big, &h->d_big_bytes, bs + be, false)) return gfail(g, h->err);
    char* din = (char*)h->d_in; char* dbig = (char*)h->d_big;
    GCK(cudaMemcpyAsync(din, x0 + (batched ? (size_t)lo * nx : 0), (size_t)(batched ? nr : 1) * nx * 8, cudaMemcpyHostToDevice, h->stream));
    GCK(cudaMemcpyAsync(din + bx, controls + (size_t)lo * H * nu, bc, cudaMemcpyHostToDevice, h->stream));
    if (b200mpc_rollout_dev(h, (double*)din, batched, (double*)(din + bx), nr, H, (double*)dbig, sensors ? (double*)(dbig + bs) : nullptr, h->stream))
      return gfail(g, h->err);
  }
  for (int r = 0; r < n; r++) {
    b200mpc_handle* h = g->hs[r];
    const int lo = g->lo[r], nr = g->lo[r + 1] - lo;
    if (nr <= 0) continue;
    GCK(cudaSetDevice(h->device));
    const size_t bs = al16((size_t)nr * H * nx * 8);
    char* dbig = (char*)h->d_big;
    GCK(cudaMemcpyAsync(states + (size_t)lo * H * nx, dbig, (size_t)nr * H * nx * 8, cudaMemcpyDeviceToHost, h->stream));
    if (sensors) GCK(cudaMemcpyAsync(sensors + (size_t)lo * H * ns, dbig + bs, (size_t)nr * H * ns * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  for (int r = 0; r < n; r++) { GCK(cudaSetDevice(g->hs[r]->device)); GCK(cudaStreamSynchronize(g->hs[r]->stream)); }
  return 0;
}

// The sharded plan step over candidates that sit in g->h_knots_all (pinned).  Outputs as b200mpc_plan_step; elite_knots (n_elite, K*nu) may
// be NULL.
static int group_plan_core(b200mpc_group* g, const double* x0, int N, int K, const double* basis, int H, const double* params, int optimizer,
                           const double* opt_params, double* nominal, double* sigma, double* reward_N, int* elite_idx, int n_elite,
                           double* elite_knots) {
  const int n = (int)g->hs.size();
  b200mpc_handle* h0 = g->hs[0];
  const int nx = h0->dims.nq + h0->dims.nv, nu = h0->dims.nu, np = h0->dims.n_cost_params, KNU = K * nu;
  if (optimizer < 0 || optimizer > 2) return gfail(g, "unknown optimizer");
  if (optimizer == B200MPC_OPT_MPPI && !(opt_params && opt_params[0] > 0)) return gfail(g, "temperature must be positive");
  if (optimizer == B200MPC_OPT_CEM && (!opt_params || (int)opt_params[0] <= 0 || (int)opt_params[0] > 256)) return gfail(g, "CEM needs {num_elites in 1..256, sigma_min, sigma_max}");
  if (n_elite < 0 || n_elite > 256) return gfail(g, "n_elite must be in 0..256");
  const double* knots_all = (const double*)g->h_knots_all;
  // device 0: all candidates (the update's weighted sums and the elite rows read them), uploaded on its own stream while the rollouts run
  GCK(cudaSetDevice(h0->device));
  if (grow(h0, &g->d_knots_all, &g->d_knots_all_bytes, (size_t)N * KNU * 8, false) || grow(h0, &g->d_reward_all, &g->d_reward_all_bytes, (size_t)N * 8, false))
    return gfail(g, h0->err);
  GCK(cudaMemcpyAsync(g->d_knots_all, knots_all, (size_t)N * KNU * 8, cudaMemcpyHostToDevice, g->copy_stream));
  GCK(cudaEventRecord(g->ev_copy, g->copy_stream));
  for (int r = 0; r < n; r++) {
    b200mpc_handle* h = g->hs[r];
    const int lo = g->lo[r], nr = g->lo[r + 1] - lo;
    if (nr <= 0) continue;
    GCK(cudaSetDevice(h->device));
    size_t o = 0;
    const size_t ox0 = o; o += al16((size_t)nx * 8);
    const size_t ob = o; o += al16((size_t)H * K * 8);
    const size_t op = o; o += al16((size_t)np * 8);
    const size_t small = o;
    const size_t ok = o; o += (size_t)nr * KNU * 8;
    if (grow(h, &h->h_in, &h->h_in_bytes, small, true) || grow(h, &h->d_in, &h->d_in_bytes, o, false) ||
        grow(h, &h->d_big, &h->d_big_bytes, (size_t)nr * 8, false)) return gfail(g, h->err);
    char* hp = (char*)h->h_in;
    memcpy(hp + ox0, x0, (size_t)nx * 8); memcpy(hp + ob, basis, (size_t)H * K * 8); memcpy(hp + op, params, (size_t)np * 8);
    char* din = (char*)h->d_in;
    GCK(cudaMemcpyAsync(din, hp, small, cudaMemcpyHostToDevice, h->stream));
    GCK(cudaMemcpyAsync(din + ok, knots_all + (size_t)lo * KNU, (size_t)nr * KNU * 8, cudaMemcpyHostToDevice, h->stream));  // straight from the pinned block
    PlanEpilogue none{};
    none.optimizer = EP_NONE;
    none.index_offset = lo;
    if (plan_costs_ep(h, (double*)(din + ox0), (double*)(din + ok), nr, K, (double*)(din + ob), H, (double*)(din + op), nullptr, (double*)h->d_big, none,
                      h->stream)) return gfail(g, h->err);
    GCK(cudaEventRecord(g->ev[r], h->stream));
  }
  // device 0: collect the slices' rewards, then the update over all N
  GCK(cudaSetDevice(h0->device));
  cudaStream_t s0 = h0->stream;
  for (int r = 0; r < n; r++) {
    const int lo = g->lo[r], nr = g->lo[r + 1] - lo;
    if (nr <= 0) continue;
    if (r > 0) GCK(cudaStreamWaitEvent(s0, g->ev[r], 0));
    GCK(cudaMemcpyPeerAsync((double*)g->d_reward_all + lo, h0->device, g->hs[r]->d_big, g->hs[r]->device, (size_t)nr * 8, s0));
  }
  GCK(cudaStreamWaitEvent(s0, g->ev_copy, 0));
  const size_t o_nom = 0, o_sig = al16((size_t)KNU * 8), o_el = o_sig + al16((size_t)KNU * 8), o_ek = o_el + al16((size_t)std::max(n_elite, 1) * 8);
  const size_t o_rw = o_ek + al16((size_t)std::max(n_elite, 1) * KNU * 8), out_bytes = o_rw + (size_t)N * 8;
  if (grow(h0, &h0->h_out, &h0->h_out_bytes, out_bytes, true)) return gfail(g, h0->err);
  void* dout_v = nullptr;
  GCK(cudaHostGetDevicePointer(&dout_v, h0->h_out, 0));
  char* dout = (char*)dout_v;
  const double* dk = (const double*)g->d_knots_all;
  const double* dr = (const double*)g->d_reward_all;
  if (run_update(h0, optimizer, opt_params, dk, dr, N, KNU, (double*)(dout + o_nom), (double*)(dout + o_sig), nullptr, 0, s0)) return gfail(g, h0->err);
  if (n_elite > 0) {
    const int nb = n_partials_for(N);
    const size_t need = 16 + (size_t)nb * n_elite * (2 + KNU) * 8 + al16((size_t)KNU * 8);  // 16: the ticket slot in front
    if (need > h0->d_part_bytes) { GCK(cudaStreamSynchronize(s0)); if (grow(h0, &h0->d_part, &h0->d_part_bytes, need, false)) return gfail(g, h0->err); GCK(cudaMemsetAsync(h0->d_part, 0, 16, s0)); }
    double* part = (double*)h0->d_part + 2;
    double* dummy = part + (size_t)nb * n_elite * (2 + KNU);
    topk_partial_kernel<<<nb, 256, 0, s0>>>(dk, dr, N, KNU, n_elite, 0, 1, part);
    h0->launches++;
    GCK(cudaGetLastError());
    if (b200mpc_topk_combine_dev(h0, part, nb, KNU, n_elite, 1, 0, 0, dummy, nullptr, (double*)(dout + o_el), s0)) return gfail(g, h0->err);
    gather_rows_kernel<<<n_elite, 64, 0, s0>>>(dk, (double*)(dout + o_el), 0, KNU, (double*)(dout + o_ek));
    h0->launches++;
    GCK(cudaGetLastError());
  }
  if (reward_N) GCK(cudaMemcpyAsync((char*)h0->h_out + o_rw, g->d_reward_all, (size_t)N * 8, cudaMemcpyDeviceToHost, s0));
  GCK(cudaStreamSynchronize(s0));
  const char* ho = (const char*)h0->h_out;
  memcpy(nominal, ho + o_nom, (size_t)KNU * 8);
  if (sigma && optimizer == B200MPC_OPT_CEM) memcpy(sigma, ho + o_sig, (size_t)KNU * 8);
  if (elite_idx) for (int i = 0; i < n_elite; i++) elite_idx[i] = (int)((const double*)(ho + o_el))[i];
  if (elite_knots && n_elite > 0) memcpy(elite_knots, ho + o_ek, (size_t)n_elite * KNU * 8);
  if (reward_N) memcpy(reward_N, ho + o_rw, (size_t)N * 8);
  return 0;
}

extern "C" int b200mpc_group_plan_step(b200mpc_group* g, const double* x0, const double* knots, int N, int K, const double* basis, int H,
                                       const double* params, int optimizer, const double* opt_params, double* nominal, double* sigma,
                                       double* reward_N, int* elite_idx, int n_elite) {
  if (!g) return 1;
  if (!x0 || !knots || !basis || !params || !nominal) return gfail(g, "NULL argument");
  if (N <= 0 || H <= 0 || K <= 0) return gfail(g, "N, H and K must be positive");
  if (N != g->N) return gfail(g, "knots batch size does not match num_rollouts (call update first)");
  b200mpc_handle* h0 = g->hs[0];
  const size_t bytes = (size_t)N * K * h0->dims.nu * 8;
  GCK(cudaSetDevice(h0->device));
  if (grow(h0, &g->h_knots_all, &g->h_knots_all_bytes, bytes, true)) return gfail(g, h0->err);
  memcpy(g->h_knots_all, knots, bytes);
  g->cand_N = N; g->cand_K = K;
  return group_plan_core(g, x0, N, K, basis, H, params, optimizer, opt_params, nominal, sigma, reward_N, elite_idx, n_elite, nullptr);
}

// trace sensors (ne, H, trace_width) of the given GLOBAL rollout indices of the last plan step (warp-per-rollout tasks with trace capture)
extern "C" int b200mpc_group_elite_traces(b200mpc_group* g, const int* idx, int ne, int H, double* out) {
  if (!g) return 1;
  if (!idx || !out || ne <= 0) return gfail(g, "NULL / empty argument");
  const int nt = trace_width(g->hs[0]);
  const size_t row = (size_t)H * nt * 8;
  for (int i = 0; i < ne; i++) {
    int r = 0;
    while (r + 1 < (int)g->hs.size() && idx[i] >= g->lo[r + 1]) r++;
    b200mpc_handle* h = g->hs[r];
    const int loc = idx[i] - g->lo[r];
    if (!h->trace_capture || h->trace_N == 0) return gfail(g, "no captured traces: enable trace capture and run a plan step first");
    if (H != h->trace_H || loc < 0 || loc >= h->trace_N) return gfail(g, "rollout index / H do not match the captured plan step");
    GCK(cudaSetDevice(h->device));
    GCK(cudaMemcpyAsync((char*)out + (size_t)i * row, (const char*)h->d_trace + (size_t)loc * row, row, cudaMemcpyDeviceToHost, h->stream));
  }
  for (auto* h : g->hs) { GCK(cudaSetDevice(h->device)); GCK(cudaStreamSynchronize(h->stream)); }
  return 0;
}

extern "C" int b200mpc_group_last_candidates(b200mpc_group* g, double* knots_out, int N, int K) {
  if (!g) return 1;
  if (!knots_out || !g->h_knots_all || g->cand_N == 0) return gfail(g, "no candidates: run a plan step first");
  if (N != g->cand_N || K != g->cand_K) return gfail(g, "N / K do not match the last plan step");
  memcpy(knots_out, g->h_knots_all, (size_t)N * K * g->hs[0]->dims.nu * 8);
  return 0;
}

extern "C" int b200mpc_group_controller_speculation(b200mpc_group* g, const unsigned int* mt_key, const int* mt_pos, size_t n) {
  return g ? b200mpc_controller_speculation(g->hs[0], mt_key, mt_pos, n) : 0;
}

// b200mpc_controller_step over the group: same request, same sampling protocol (the generator bookkeeping lives in the first handle)
extern "C" int b200mpc_group_controller_step(b200mpc_group* g, b200mpc_step_request* rq) {
  if (!g) return 1;
  if (!rq) return gfail(g, "NULL request");
  b200mpc_handle* h0 = g->hs[0];
  const int N = rq->N, K = rq->K, H = rq->H;
  if (!rq->x0 || !rq->nominal || !rq->sigma || !rq->lo || !rq->hi || !rq->cost_params || !rq->knot_times || !rq->nominal_out) return gfail(g, "NULL argument");
  if (N <= 0 || H <= 0 || K <= 0) return gfail(g, "N, H and K must be positive");
  if (N != g->N) return gfail(g, "N does not match num_rollouts (call update first)");
  if (rq->phase < 0 || rq->phase > 2) return gfail(g, "phase must be 0, 1 or 2");
  const int nu = h0->dims.nu, KNU = K * nu, ns = h0->dims.nsensordata;
  const size_t n = (size_t)(N - 1) * KNU;
  if (rq->n_head < 0 || rq->n_head > 2 || (size_t)rq->n_head > n) return gfail(g, "n_head must be 0..2 and <= (N-1)*K*nu");
  if (n > (size_t)rq->n_head && (!rq->mt_key || !rq->mt_pos)) return gfail(g, "NULL generator state");
  const bool warp_task = h0->task == B200MPC_TASK_LEAP_CUBE || h0->task == B200MPC_TASK_FR3_PICK;
  const int nts = rq->n_trace_sensors, ne = rq->n_elite;
  if (nts < 0 || (nts > 0 && (!rq->trace_cols || !rq->traces))) return gfail(g, "trace sensors requested without trace_cols / traces");
  if (nts > 0 && warp_task && (!h0->trace_capture || trace_width(h0) != 3 * nts)) return gfail(g, "enable trace capture for the traces of this task");
  {
    const int rc = step_sample(h0, rq, n);
    if (rc == 2) return 0;
    if (rc) return gfail(g, h0->err);
  }
  GCK(cudaSetDevice(h0->device));
  if (grow(h0, &g->h_knots_all, &g->h_knots_all_bytes, (size_t)N * KNU * 8, true)) return gfail(g, h0->err);
  std::vector<double> basis((size_t)H * K);
  h0->qtimes.resize(H);
  for (int i = 0; i < H; i++) h0->qtimes[i] = rq->time + rq->dt * (double)i;
  if (b2host::spline_basis(rq->spline_order, rq->knot_times, K, h0->qtimes.data(), H, basis.data())) return gfail(g, "bad spline request (order / number of knots)");
  if (rq->basis_out) memcpy(rq->basis_out, basis.data(), (size_t)H * K * 8);
  b2host::assemble_candidates(h0->zbuf.data(), rq->nominal, rq->sigma, rq->lo, rq->hi, N, K, nu, (double*)g->h_knots_all);
  g->cand_N = N; g->cand_K = K;
  if (rq->knots_out) memcpy(rq->knots_out, g->h_knots_all, (size_t)N * KNU * 8);
  // The sharded step blocks until device 0 has the result; the next block of normals is drawn first only when the rollouts are long enough
  // to hide it (the warp-per-rollout tasks), otherwise after.
  std::vector<double> ek((size_t)std::max(ne, 1) * KNU);
  int el[256];
  if (warp_task) step_speculate(h0, rq, n);
  if (group_plan_core(g, rq->x0, N, K, basis.data(), H, rq->cost_params, rq->optimizer, rq->opt_params, rq->nominal_out, rq->sigma_out, rq->rewards,
                      el, ne, ek.data())) return 1;
  if (!warp_task) step_speculate(h0, rq, n);
  if (rq->elite_idx) for (int i = 0; i < ne; i++) rq->elite_idx[i] = el[i];
  if (nts > 0 && ne > 0) {
    if (warp_task) {
      const int nt = 3 * nts;
      if (nt > 48) return gfail(g, "too many trace sensors");
      g->tmp.resize((size_t)ne * H * nt);
      if (b200mpc_group_elite_traces(g, el, ne, H, g->tmp.data())) return 1;
      int cols[48];
      for (int i = 0; i < nt; i++) cols[i] = i;
      b2host::trace_segments(g->tmp.data(), ne, H, nt, cols, nts, rq->traces);
    } else {
      // thread-per-rollout tasks: re-simulate the <= 8 elite candidates on the first device (microseconds) and read their sensors
      std::vector<double> ctrl((size_t)ne * H * nu, 0.0), st((size_t)ne * H * (h0->dims.nq + h0->dims.nv)), se((size_t)ne * H * ns);
      for (int e = 0; e < ne; e++)
        for (int t = 0; t < H; t++)
          for (int k = 0; k < K; k++) {
            const double b = basis[(size_t)t * K + k];
            if (b != 0.0) for (int j = 0; j < nu; j++) ctrl[((size_t)e * H + t) * nu + j] += b * ek[(size_t)e * KNU + k * nu + j];
          }
      const int keep = h0->N;
      h0->N = ne;
      const int rc = b200mpc_rollout(h0, rq->x0, 0, ctrl.data(), ne, H, st.data(), se.data());
      h0->N = keep;
      if (rc) return gfail(g, h0->err);
      b2host::trace_segments(se.data(), ne, H, ns, rq->trace_cols, nts, rq->traces);
    }
  }
  return 0;
}
