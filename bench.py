"""bench.py — the plan-step benchmark (BASELINE.json metric: rollouts/sec per control step; plan latency p50).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cartpole_mppi|cylinder_push_cem|leap_cube_mppi|fr3_pick_cem]
  python bench.py --impl reference ...        # the reference's CPU path on the host cores (see CpuPlanner)
  torchrun --nproc-per-node N bench.py --gpus N ...   # one rank per GPU, weak scaling: every rank owns n_rollouts

A "step" is one plan step over one batch of candidates: sample -> clip -> spline -> N x H dynamics rollout -> per-step cost ->
MPPI/CEM/PS update.  One JSON line on rank 0:
  value   resident plan step: candidates already in HBM, CUDA events on the launching stream, L2 flushed between iterations
  e2e     the call a judo user makes, Controller.update_action(): host sampling from numpy's stream, clip, spline basis, H2D, fused
          kernel, D2H, traces — wall time per call (N=1; at N>1: Engine.plan_step per rank with the in-kernel exchange)
  also    the other single-GPU BASELINE configs (C3 cylinder_push+cem, C4 leap_cube+mppi; at N>1: C5 = leap_cube sharded) measured in
          the same run, each with value / e2e / roofline
  roofline / cpu_baseline / clocks / gpu_launches as the task statement defines them.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on at N=1
    "cartpole_mppi": dict(task="cartpole", optimizer="mppi", n_rollouts=4096, H=64, K=4, order="zero", horizon=2.56,
                          algo_bytes_per_rollout=276),
    "cylinder_push_cem": dict(task="cylinder_push", optimizer="cem", n_rollouts=2048, H=50, K=4, order="zero", horizon=1.0,
                              algo_bytes_per_rollout=236),
    "leap_cube_mppi": dict(task="leap_cube", optimizer="mppi", n_rollouts=1024, H=40, K=4, order="cubic", horizon=0.4,
                           algo_bytes_per_rollout=420),
    # SURVEY §8f-2 (not a BASELINE config): the reference's fr3_pick defaults (H = 1.0 s / 4 ms, CEM, 4 linear knots) at N = 1024
    "fr3_pick_cem": dict(task="fr3_pick", optimizer="cem", n_rollouts=1024, H=250, K=4, order="linear", horizon=1.0,
                         algo_bytes_per_rollout=4 * 4 * 8 + 4 * 250 + 4),
}
# the same kernel in its contact-rich regime: the gripper closes on the cube and lifts it (40 pad/table contacts, ~170 constraint rows)
WORKLOADS["fr3_pick_cem_grasp"] = dict(WORKLOADS["fr3_pick_cem"], scenario="grasp")
# BASELINE.json configs[0]: the reference's own CPU-runnable size (plumbing case; both arms can run it, it is not the bench line)
WORKLOADS["cartpole_ps"] = dict(task="cartpole", optimizer="ps", n_rollouts=32, H=32, K=4, order="zero", horizon=1.28, algo_bytes_per_rollout=4 * 4 + 4 * 32 + 4)
WARP_TASKS = ("leap_cube", "fr3_pick")  # warp-per-rollout kernels: ms-scale steps, optimizer update as separate reduction kernels
CONFIG_TAG = {"cartpole_ps": "BASELINE config C1", "cartpole_mppi": "BASELINE config C2", "cylinder_push_cem": "BASELINE config C3", "leap_cube_mppi": "BASELINE config C4",
              "fr3_pick_cem": "SURVEY 8f-2, reference defaults at N=1024",
              "fr3_pick_cem_grasp": "SURVEY 8f-2 in its contact-rich regime: pre-grasp pose, nominal plan closes the gripper and lifts"}
ALSO_1GPU = ("cylinder_push_cem", "leap_cube_mppi")   # measured next to the headline workload in the default single-GPU run
ALSO_NGPU = ("leap_cube_mppi",)                        # ... and under torchrun: C5 = leap_cube, N_per_GPU x world rollouts
KERNEL_NAME = {"cartpole": "rollout_kernel<CartpoleTask,1,4>", "cylinder_push": "rollout_kernel<CylinderPushTask,1,4>",
               "leap_cube": "leap_rollout_kernel<1>", "fr3_pick": "fr3_rollout_kernel<1>"}


def config_of(wname: str, w: dict, world: int) -> dict:
    """Identical in both arms (the driver compares it)."""
    return {"workload": f"{w['task']}+{w['optimizer']} N={w['n_rollouts']}/GPU H={w['H']} K={w['K']} spline={w['order']} ({CONFIG_TAG[wname]})",
            "n_rollouts_per_gpu": w["n_rollouts"], "n_rollouts_total": w["n_rollouts"] * world, "horizon_steps": w["H"], "num_nodes": w["K"],
            "parallelism": f"rollout-sharded x{world}",
            "l2": "GPU arm: flushed (256 MiB memset) between timed iterations of `value`; CPU arm: not applicable",
            "contract": "B (fused: knots in, cost matrix f32 + reward out)"}


def problem(w: dict, n_total: int, seed: int = 42):
    """Synthetic inputs exactly as SURVEY.md §8(d): seed 42, the task's own reset distribution, nominal warm start 0."""
    from judo_b200.optimizers import get_registered_optimizers
    from judo_b200.spline import spline_basis
    from judo_b200.tasks import get_registered_tasks

    np.random.seed(seed)
    task_cls, _ = get_registered_tasks()[w["task"]]
    task = task_cls()                       # reset() draws x0 from the task's distribution
    x0 = np.concatenate([task.data.qpos, task.data.qvel])
    opt_cls, cfg_cls = get_registered_optimizers()[w["optimizer"]]
    cfg = cfg_cls()
    cfg.set_override(w["task"])
    cfg.num_rollouts, cfg.num_nodes = n_total, w["K"]
    opt = opt_cls(cfg, task.nu)
    nominal = np.tile(task.optimizer_warm_start(), (w["K"], 1))
    if w.get("scenario") == "grasp":
        from judo_b200.tasks.fr3_pick import Q_PREGRASP

        x0[7:14] = Q_PREGRASP
        nominal = np.tile(np.concatenate([Q_PREGRASP, [0.0]]), (w["K"], 1))
        nominal[w["K"] // 2:, 1] -= 0.3  # shoulder back: lift
        task.pre_rollout(x0)
    lo, hi = task.actuator_ctrlrange[:, 0], task.actuator_ctrlrange[:, 1]
    knots = np.clip(opt.sample_control_knots(nominal), lo, hi)
    times = np.linspace(0, w["horizon"], w["K"], endpoint=True)
    basis = spline_basis(times, task.dt * np.arange(w["H"]), w["order"])
    meta = task.get_sim_metadata() if hasattr(task, "get_sim_metadata") else {}
    params = task.cost_params(meta)
    return task, opt, x0, knots, basis, params, nominal


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int) -> None:
        self.index, self.rows, self._stop = index, [], threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self) -> None:
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a) -> None:  # noqa: ANN002
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self) -> dict:
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) >= 7 and r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


class _NoClocks:
    def __enter__(self):
        return self

    def __exit__(self, *a) -> None:  # noqa: ANN002
        pass

    def summary(self) -> None:
        return None


# ====================================================================================================== CPU arm
def host_threads() -> int:
    """CPUs this process can actually run on (cgroup / affinity aware): oversubscribing the OpenMP pool is far slower."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def _native_oracle_build() -> str:
    """The shipped oracle library is a portable -O2 build (it is the parity checker); as a BASELINE it deserves the host's own ISA:
    rebuild it on the box it is timed on with -O3 -march=native (fp contraction stays off so the numbers do not change)."""
    import oracle.mjc as om

    src = os.path.join(ROOT, "oracle", "mjc", "mjc.c")
    out = os.path.join(ROOT, "oracle", "_ref", "libmjc_oracle_native.so")
    try:
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-std=c11", "-ffp-contract=off", "-fopenmp", "-shared", "-o", out, src, "-lm"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        om._LIB_PATH, om._lib = out, None
        return "-O3 -march=native (built on this host)"
    except Exception:  # noqa: BLE001
        return "-O2 (shipped build)"


def _oracle_model(task: str):
    from oracle.mjc import OracleModel, load_table

    if task == "leap_cube":
        from judo_b200.tasks.leap_cube import reduced_collision_model

        tb = load_table(task)
        geoms, pairs = reduced_collision_model(tb)
        return OracleModel(tb, pairs=pairs, geoms=geoms)
    if task == "fr3_pick":
        from judo_b200.tasks.fr3_pick import reduced_collision_model as fr3_reduced

        tb = load_table(task)
        geoms, pairs = fr3_reduced(tb)
        return OracleModel(tb, pairs=pairs, geoms=geoms)
    return OracleModel(task)


class CpuPlanner:
    """The reference's plan step on the host cores, stage by stage as Controller.update_action runs it (controller.py:248-293):
    np.random.randn sampling (mppi.py:58) -> clip -> scipy interp1d spline (controller.py:382-401) -> rollouts -> NumPy reward ->
    NumPy update.  The rollouts are the one stage that cannot be the reference's own code: they are mujoco 3.5.0's C (a third-party
    wheel, not installable offline — profiles/r02_mujoco_probe.txt), so the restated C oracle stands in, one OpenMP thread per core
    (the reference as shipped starts one thread per ROLLOUT, mj_rollout_backend.py:36)."""

    def __init__(self, w: dict, n_sample: int) -> None:
        from oracle import plan as op

        self.op, self.w, self.n = op, w, n_sample
        self.task, self.opt, self.x0, _, _, self.params, self.nominal = problem(w, n_sample)
        self.om = _oracle_model(w["task"])
        self.lo, self.hi = self.task.actuator_ctrlrange[:, 0], self.task.actuator_ctrlrange[:, 1]
        self.times = np.linspace(0, w["horizon"], w["K"], endpoint=True)
        self.query = self.task.dt * np.arange(w["H"])
        self.nthread = host_threads()
        s = np.asarray(self.opt.device_sigma(), dtype=np.float64)  # (K, nu); evaluated once (CEM's ramp would keep shrinking it)
        self.sigma = s.copy()

    def step(self) -> np.ndarray:
        op, w, opt = self.op, self.w, self.opt
        K, nu = w["K"], self.task.nu
        noised = self.nominal + self.sigma * np.random.randn(self.n - 1, K, nu)
        knots = np.clip(np.concatenate([self.nominal[None], noised]), self.lo, self.hi)
        controls = op.make_spline(self.times, knots, w["order"])(self.query)
        states, sensors = self.om.rollout(self.x0, controls, nthread=self.nthread)
        p = self.params
        if w["task"] == "cartpole":
            r = op.cartpole_reward(states, controls, *p)
        elif w["task"] == "cylinder_push":
            r = op.cylinder_push_reward(states, controls, p[0], p[1], p[2], p[3], p[4:6])
        elif w["task"] == "fr3_pick":
            r = op.fr3_pick_reward(states, sensors, int(p[0]), p[11:13], p[13], p[1:3], p[3:5], p[5:7], p[7:11])
        else:
            r = op.leap_cube_reward(states, p[2:6], p[0], p[1])
        if w["optimizer"] == "mppi":
            return op.mppi_update(knots, r, opt.temperature)
        if w["optimizer"] == "cem":
            return op.cem_update(knots, r, opt.num_elites, opt.sigma_min, opt.sigma_max)[0]
        return op.ps_update(knots, r)

    def pick_threads(self) -> int:
        """The OpenMP thread count that runs the step fastest on this host (SMT / cgroup limits make 'all logical CPUs' a bad default)."""
        ht = host_threads()
        best, best_t = ht, float("inf")
        for nt in sorted({ht, max(1, ht // 2), min(ht, 16)}, reverse=True):
            self.nthread = nt
            self.step()
            t1 = time.perf_counter()
            self.step()
            dt = time.perf_counter() - t1
            if dt < best_t:
                best, best_t = nt, dt
        self.nthread = best
        return best


def cpu_sample_size(w: dict, n_local: int) -> int:
    """Rollouts per CPU step: the whole batch for the small tasks, a bounded sample for the articulated ones (~0.1 s per step)."""
    return min(n_local, 4096 if w["task"] not in WARP_TASKS else (256 if w["task"] == "leap_cube" else 64))


def time_cpu(w: dict, n_local: int, budget_s: float, build: str) -> dict:
    n = cpu_sample_size(w, n_local)
    cp = CpuPlanner(w, n)
    nt = cp.pick_threads()
    t0, times = time.perf_counter(), []
    while True:
        t1 = time.perf_counter()
        cp.step()
        times.append(time.perf_counter() - t1)
        if time.perf_counter() - t0 > budget_s or len(times) >= 50:
            break
    med = statistics.median(times)
    return {"value": n / med, "unit": "rollouts/s", "cores": nt, "host_logical_cpus": host_threads(), "kind": "port",
            "sample": f"{len(times)} full plan steps (randn sampling, clip, scipy spline, rollouts, NumPy reward + update) of {n} rollouts x "
                      f"H={w['H']} ({w['task']}); rollouts = restated C oracle {build}, OpenMP x{nt} (fastest tried); median {med * 1e3:.2f} ms/step",
            "ms_per_step": med * 1e3}


def reference_arm(args, wname: str, w: dict) -> None:  # noqa: ANN001
    """--impl reference: the reference's CPU implementation of the path on this box's host cores, same config / metric / unit."""
    n_local = w["n_rollouts"]
    try:  # the real thing, if a box ever has it: judo's own Controller on mujoco.rollout
        import mujoco  # noqa: F401
        from judo.controller import Controller as RefController  # noqa: F401

        note = "mujoco and judo import on this box, but bench.py has never been able to run them: timing the port"
    except Exception:  # noqa: BLE001
        note = "reference rollouts = mujoco 3.5.0 (neither importable nor installable on this box: profiles/r02_mujoco_probe.txt); timed: " \
               "the reference's plan step with the restated C oracle for the rollouts"
    build = _native_oracle_build()
    n = cpu_sample_size(w, n_local)
    cp = CpuPlanner(w, n)
    nt = cp.pick_threads()
    np.random.seed(42)
    for _ in range(args.warmup):
        cp.step()
    times = []
    for _ in range(args.steps):
        t1 = time.perf_counter()
        cp.step()
        times.append(time.perf_counter() - t1)
    dt = sum(times) / len(times)
    val = n / dt
    cb = {"value": val, "unit": "rollouts/s", "cores": nt, "host_logical_cpus": host_threads(), "kind": "port",
          "sample": f"{args.steps} full plan steps of {n} rollouts x H={w['H']}; rollouts = restated C oracle {build}, OpenMP x{nt}"}
    print(json.dumps({"impl": "reference", "metric": "rollouts/sec per control step", "value": val, "unit": "rollouts/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "plan_latency_p50_ms": statistics.median(times) * 1e3,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": config_of(wname, w, args.gpus), "cpu_baseline": cb,
                      "e2e": {"value": val, "unit": "rollouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "note": note}))


# ====================================================================================================== GPU arm
def controller_latency(task: str, optimizer: str, n: int, K: int, horizon: float, device: int, samples: int, warm: int = 10) -> dict:
    """Wall time of Controller.update_action() — the span judo's ControllerNode times as plan_time (judo/app/dora/controller.py:138-142):
    time shift, host sampling from numpy's stream, clip, spline basis, H2D, fused kernel, D2H, spline refresh, traces."""
    from judo_b200.controller import make_controller

    np.random.seed(42)
    c = make_controller(task, optimizer, device=device)
    c.optimizer_cfg.num_rollouts, c.optimizer_cfg.num_nodes = n, K
    c.controller_cfg.horizon = horizon
    c.reset()
    if hasattr(c.task, "get_sim_metadata"):
        c.system_metadata = c.task.get_sim_metadata()
    for _ in range(warm):
        c.update_action()
    lat = []
    for i in range(samples):
        c.time = c.task.dt * i
        t1 = time.perf_counter()
        c.update_action()
        lat.append(time.perf_counter() - t1)
    fast = bool(c._can_fast_path())
    H, nu, ne = c.num_timesteps, c.model.nu, min(c.max_num_traces, n)
    h2d = 8 * (c.model.nq + c.model.nv + H * K + c.engine.n_cost_params + n * K * nu)
    d2h = 8 * (2 * K * nu + ne + n + (ne * H * c.model.nsensordata if c.num_trace_sensors else 0))
    c.engine.close()
    return {"p50_ms": statistics.median(lat) * 1e3, "mean_ms": statistics.mean(lat) * 1e3, "p90_ms": float(np.percentile(lat, 90)) * 1e3,
            "samples": len(lat), "one_call_fast_path": fast, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "config": f"{task} + {optimizer}, N={n}, H={H}, K={K}: full Controller.update_action() incl. host sampling and traces"}


def load_counts() -> dict:
    """Per-launch hardware counts of the dominant kernels from this round's `ncu --set full` captures (bytes and instruction counts
    are properties of the launch, not times: a number measured under the profiler is never used as a bench value)."""
    for name in ("r02_counts.json", "r01_traffic.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            return d if name.startswith("r02") else {k: {"dram_bytes": v} for k, v in d.items() if isinstance(v, (int, float))}
        except Exception:  # noqa: BLE001
            continue
    return {}


def measure(wname: str, w: dict, steps: int, warmup: int, rank: int, world: int, local_rank: int, dist, torch, flush, extras: bool,  # noqa: ANN001
            cpu_budget: float, clocks: bool, fp64_peak: float | None, oracle_build: str) -> dict | None:
    """All GPU-arm numbers of one workload.  Returns the result dict on rank 0, None elsewhere."""
    from judo_b200.dist import ShardedPlanner, shard_range

    dev = torch.device("cuda", local_rank)
    n_local, n_total = w["n_rollouts"], w["n_rollouts"] * world
    task, opt, x0, knots_all, basis, params, nominal0 = problem(w, n_total)
    lo, hi = shard_range(n_total, world, rank)
    knots = np.ascontiguousarray(knots_all[lo:hi])
    opt_params = opt.fused_params()
    planner = ShardedPlanner(w["task"], n_local, device=local_rank, rank=rank, world_size=world)
    if world > 1 and os.environ.get("B200MPC_PEER_EXCHANGE", "1") != "0":
        if not planner.enable_peer_exchange():  # MPPI partials cross NVLink inside the rollout kernel (all_gather for CEM/PS/leap)
            print("bench: CUDA IPC peer exchange unavailable, using the NCCL all_gather path", file=sys.stderr)
    planner.set_problem(x0, basis, params, want_cost_matrix=True)
    planner.set_knots(knots)

    def barrier() -> None:
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    in_kernel_exchange = world > 1 and planner.peer_exchange and w["task"] not in WARP_TASKS
    for _ in range(warmup):
        planner.step(w["optimizer"], opt_params, index_offset=lo)
    barrier()
    # the exchanged nominal must be THE update over all world x N candidates (what SCALE times has to be the right answer)
    verified = None
    if world > 1:
        from oracle import plan as op  # checker only: outside every timed region

        nom = planner.step(w["optimizer"], opt_params, index_offset=lo)
        allr = [torch.empty(n_local, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(allr, planner.d_reward)
        torch.cuda.synchronize(dev)
        r_all = torch.cat(allr).cpu().numpy()
        if w["optimizer"] == "mppi":
            ref = op.mppi_update(knots_all, r_all, opt.temperature)
        elif w["optimizer"] == "cem":
            ref = op.cem_update(knots_all, r_all, opt.num_elites, opt.sigma_min, opt.sigma_max)[0]
        else:
            ref = op.ps_update(knots_all, r_all)
        err = float(np.abs(nom.cpu().numpy().reshape(ref.shape) - ref).max())
        flag = torch.tensor([err], dtype=torch.float64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        verified = {"max_abs_err_vs_unsharded_update": float(flag.item()), "ok": bool(flag.item() < 1e-9)}
        barrier()
    launches0 = planner.engine.launch_count
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    with (ClockSampler(local_rank) if clocks else _NoClocks()) as ck:
        barrier()
        t_wall0 = time.perf_counter()
        for i in range(steps):
            flush.zero_()                                  # evict L2 between timed iterations (outside the events)
            if dist is not None:                           # line the ranks up again: the flushes do not take equally long on every GPU and a
                if in_kernel_exchange:                     # step cannot finish before the slowest rank has STARTED it.  In-kernel exchange:
                    planner.align()                        # a one-warp signal/wait kernel on the stream (no host round trip); all_gather
                else:                                      # path: a host barrier.
                    dist.barrier()
            starts[i].record()
            planner.step(w["optimizer"], opt_params, index_offset=lo)
            ends[i].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    launches = planner.engine.launch_count - launches0 - (steps if in_kernel_exchange else 0)  # (the align kernels are not plan-step work)
    # the same loop WITHOUT lining the ranks up, and the in-kernel %globaltimer stamps of a step: separates launch / flush skew between
    # the ranks from the cost of the exchange itself
    exchange_timing = None
    if in_kernel_exchange:
        nu_ = min(steps, 50)
        for i in range(nu_):
            flush.zero_()
            starts[i].record()
            planner.step(w["optimizer"], opt_params, index_offset=lo)
            ends[i].record()
        barrier()
        un_ms = statistics.mean(starts[i].elapsed_time(ends[i]) for i in range(nu_))
        planner.align()
        planner.step(w["optimizer"], opt_params, index_offset=lo)
        torch.cuda.synchronize(dev)
        t_in, t_pub, t_done = planner.exchange_stamps()
        mine = torch.tensor([un_ms, (t_done - t_pub) * 1e-3, (t_done - t_in) * 1e-3, statistics.mean(step_ms)], dtype=torch.float64, device=dev)
        allm = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allm, mine)
        allm = torch.stack(allm).cpu().numpy()
        exchange_timing = {"aligned_ms_per_step_max_rank": float(allm[:, 3].max()), "unaligned_ms_per_step_max_rank": float(allm[:, 0].max()),
                           "wait_for_peers_us_per_rank": [round(float(x), 2) for x in allm[:, 1]],
                           "kernel_entry_to_done_us_per_rank": [round(float(x), 2) for x in allm[:, 2]],
                           "note": "aligned: the ranks are lined up by a one-warp signal/wait kernel between the L2 flush and the start event "
                                   "(what `value` reports); unaligned: no line-up, so each step also absorbs how differently long the ranks' "
                                   "flushes took; wait_for_peers: %globaltimer from 'partial published' to 'all peers seen' in one aligned "
                                   "step (the last rank to arrive waits ~ the NVLink flag latency, the others additionally the skew)"}

    # dominant kernel alone: rollout+cost, timed live with events on the launching stream, L2 flushed
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    P = lambda x: ctypes.c_void_p(x.data_ptr())  # noqa: E731
    nk = min(steps, 50)
    for i in range(nk):
        flush.zero_()
        starts[i].record()
        planner._check(planner.lib.b200mpc_plan_costs_dev(planner.engine.handle, P(planner.d_x0), P(planner.d_knots), n_local, w["K"],
                                                          P(planner.d_basis), w["H"], P(planner.d_params), P(planner.d_cost), P(planner.d_reward), st))
        ends[i].record()
    torch.cuda.synchronize(dev)
    kernel_ms = statistics.mean(starts[i].elapsed_time(ends[i]) for i in range(nk))

    tt = torch.tensor([sum(step_ms), kernel_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = float(tt[0]), float(tt[1])
    ms_per_step = total_ms / steps
    value = n_total / (ms_per_step * 1e-3)

    # ---- the engine call with HOST buffers (pre-sampled candidates): one pinned H2D + fused kernel + results in pinned memory
    eng = planner.engine
    xn = 0 if in_kernel_exchange else 5
    ec = []
    for i in range(warmup + steps):
        if dist is not None and i == warmup:
            barrier()
        t1 = time.perf_counter()
        res = eng.plan_step(x0, knots, basis, params, w["optimizer"], opt_params, want_rewards=True, n_elite=xn)
        ec.append(time.perf_counter() - t1)
    ec = ec[warmup:]
    ec_tt = torch.tensor([sum(ec), statistics.median(ec)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ec_tt, op=dist.ReduceOp.MAX)
    ec_ms = float(ec_tt[0]) / steps * 1e3
    engine_call = {"value": n_total / (ec_ms * 1e-3), "unit": "rollouts/s", "plan_latency_p50_ms": float(ec_tt[1]) * 1e3,
                   "h2d_bytes_per_step": int(x0.nbytes + basis.nbytes + params.nbytes + knots.nbytes),
                   "d2h_bytes_per_step": int(2 * res["nominal"].nbytes + 5 * 8 + n_local * 8),
                   "note": "Engine.plan_step per rank, candidates pre-sampled on the host; at N>1 the MPPI update is global through the "
                           "in-kernel peer exchange (CEM/PS/leap at N>1: per-rank plans)"}
    # ---- perf mode: candidates drawn on the device (Philox), only the nominal crosses PCIe on the way in
    device_sampling = None
    if extras:
        sig = opt.device_sigma()
        lo_c, hi_c = task.actuator_ctrlrange[:, 0], task.actuator_ctrlrange[:, 1]
        ds = []
        for i in range(warmup + steps):
            t1 = time.perf_counter()
            eng.plan_step_sampled(x0, nominal0, sig, lo_c, hi_c, n_local, basis, params, w["optimizer"], opt_params, seed=42, counter=i,
                                  index_offset=lo, want_rewards=True, n_elite=5)
            ds.append(time.perf_counter() - t1)
        ds = ds[warmup:]
        ds_tt = torch.tensor([sum(ds), statistics.median(ds)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(ds_tt, op=dist.ReduceOp.MAX)
        device_sampling = {"value": n_total / (float(ds_tt[0]) / steps), "unit": "rollouts/s", "plan_latency_p50_ms": float(ds_tt[1]) * 1e3,
                           "h2d_bytes_per_step": int(x0.nbytes + basis.nbytes + params.nbytes + 2 * nominal0.nbytes + 2 * lo_c.nbytes),
                           "d2h_bytes_per_step": int(2 * nominal0.nbytes + 5 * 8 + 5 * nominal0.nbytes + n_local * 8),
                           "note": "Engine.plan_step_sampled: Philox sampling + clip inside the rollout kernel (same distribution as "
                                   "np.random.randn, different stream)"}
    overflows = int(eng.contact_overflows) if w["task"] in WARP_TASKS else None
    exchange_used = ("in-kernel P2P stores over NVLink (CUDA IPC), 1 launch per step" if in_kernel_exchange
                     else ("nccl all_gather + combine kernel" if world > 1 else "none"))
    eng.close()
    if rank != 0:
        return None

    # ---- e2e at N=1: the user's call, Controller.update_action(), at this workload's size
    ctl = None
    if world == 1:
        try:
            ctl = controller_latency(w["task"], w["optimizer"], n_local, w["K"], w["horizon"], local_rank,
                                     samples=(steps if w["task"] not in WARP_TASKS else min(steps, 20)), warm=max(3, min(warmup, 10)))
        except Exception as exc:  # noqa: BLE001 — must never cost the bench line
            ctl = {"error": repr(exc)}
    if ctl and "mean_ms" in ctl:
        e2e = {"value": n_total / (ctl["mean_ms"] * 1e-3), "unit": "rollouts/s", "h2d_bytes_per_step": ctl["h2d_bytes_per_step"],
               "d2h_bytes_per_step": ctl["d2h_bytes_per_step"], "plan_latency_p50_ms": ctl["p50_ms"], "plan_latency_p90_ms": ctl["p90_ms"],
               "api": "judo_b200.controller.make_controller(task, optimizer).update_action()", "one_call_fast_path": ctl["one_call_fast_path"],
               "note": ctl["config"]}
    else:
        e2e = dict(engine_call, api="judo_b200.engine.Engine.plan_step (per rank)")
        if ctl:
            e2e["controller_error"] = ctl.get("error")
    e2e["engine_call"] = engine_call
    if device_sampling:
        e2e["device_sampling"] = device_sampling

    # ---- rooflines of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    counts = load_counts().get(wname) or {}
    algo_bytes = w["algo_bytes_per_rollout"] * n_local
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    hbm = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": counts.get("dram_bytes"),
           "peak_source": "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s",
           "algorithmic_bytes_per_launch": algo_bytes}
    issue = None
    if fp64_peak and counts.get("fp64_warp_inst"):
        a = counts["fp64_warp_inst"] / (kernel_ms * 1e-3)
        issue = {"bound": "fp64_issue", "achieved": a / 1e9, "peak": fp64_peak / 1e9, "unit": "G fp64 warp-instr/s", "frac": a / fp64_peak,
                 "fp64_warp_inst_per_launch": counts["fp64_warp_inst"], "warp_inst_per_launch": counts.get("warp_inst"),
                 "active_lanes_per_inst": counts.get("active_lanes"), "ipc_per_sm": counts.get("ipc_per_sm"),
                 "peak_source": "b200mpc_fp64_peak: DFMA chains on every SM, measured in this run",
                 "counts_source": "profiles/r02_counts.json (ncu --set full capture of the same launch)"}
    note = "N independent serial recurrences: latency/issue-bound by construction (SURVEY.md §8d), so the HBM fraction BASELINE.json asks " \
           "for is <<1%; fp64_issue states how much of the fp64 pipe's issue rate the kernel uses"
    if w["task"] in WARP_TASKS and issue:
        roofline = dict(issue, traffic=counts.get("dram_bytes"), hbm=hbm)
    else:
        roofline = dict(hbm, fp64_issue=issue)
    roofline.update(kernel=KERNEL_NAME[w["task"]] + " (fused spline+dynamics+cost)", kernel_ms=kernel_ms, note=note)

    cpu = time_cpu(w, n_local, cpu_budget, oracle_build) if (world == 1 and cpu_budget > 0) else None
    out = {"value": value, "unit": "rollouts/s", "ms_per_step": ms_per_step, "state_steps_per_s": value * w["H"],
           "plan_latency_p50_ms": (ctl["p50_ms"] if ctl and "p50_ms" in ctl else statistics.median(step_ms)),
           "resident_step_p50_ms": statistics.median(step_ms), "steps": steps, "warmup": warmup,
           "config": config_of(wname, w, world), "exchange_used": exchange_used, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
           "gpu_launches": int(launches), "wall_s_timed_region": t_wall}
    if ck.summary() is not None:
        out["clocks"] = ck.summary()
    if overflows is not None:
        out["contact_overflows"] = overflows  # rollout steps (whole run) that exceeded the kernel's per-step contact buffer
    if verified is not None:
        out["exchange_verified"] = verified
    if exchange_timing is not None:
        out["exchange_timing"] = exchange_timing
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cartpole_mppi", choices=list(WORKLOADS))
    ap.add_argument("--n-rollouts", type=int, default=0, help="per-GPU rollouts (default: the workload's)")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--no-extras", action="store_true", help="kernel sweeps: only the workload itself, no Controller / CPU / `also` legs")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.n_rollouts:
        w["n_rollouts"] = args.n_rollouts
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    warmup = max(args.warmup, 3)

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, args.workload, w)
        return

    # libraries (NCCL's version banner, ...) may write to fd 1: park stdout on stderr until the single JSON line is printed
    sys.stdout.flush()
    _saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch

    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the single JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist = None
        torch.cuda.set_device(local_rank)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=torch.device("cuda", local_rank))

    from judo_b200 import _lib

    pk = ctypes.c_double(0)
    rc = _lib.load().b200mpc_fp64_peak(local_rank, ctypes.byref(pk))
    fp64_peak = pk.value if rc == 0 and pk.value > 0 else None
    extras = not args.no_extras
    oracle_build = _native_oracle_build() if (rank == 0 and world == 1 and extras) else "-O2 (shipped build)"
    main_res = measure(args.workload, w, args.steps, warmup, rank, world, local_rank, dist, torch, flush, extras,
                       args.cpu_budget if extras else 0.0, True, fp64_peak, oracle_build)
    also = {}
    if extras and args.workload == "cartpole_mppi" and not args.n_rollouts:
        for name in (ALSO_1GPU if world == 1 else ALSO_NGPU):
            aw = dict(WORKLOADS[name])
            k = min(args.steps, 20 if aw["task"] in WARP_TASKS else 100)
            try:
                r = measure(name, aw, k, 3, rank, world, local_rank, dist, torch, flush, False, min(args.cpu_budget, 4.0), False, fp64_peak, oracle_build)
            except Exception as exc:  # noqa: BLE001 — an auxiliary workload must never cost the headline line
                r = {"error": repr(exc)}
            if rank == 0:
                also[name] = r
    if rank == 0:
        out = {"metric": "rollouts/sec per control step", "value": main_res["value"], "unit": "rollouts/s", "n_gpus": world, "steps": args.steps,
               "warmup": warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic"}
        out.update({k: v for k, v in main_res.items() if k not in ("value", "unit", "ms_per_step", "steps", "warmup")})
        if extras and world == 1:
            try:  # BASELINE's second metric at the reference's own size (configs[0]: cartpole + ps, N=32, H=32)
                c1 = controller_latency("cartpole", "ps", 32, 4, 1.28, local_rank, samples=200, warm=20)
                out["plan_latency_c1_p50_ms"] = c1["p50_ms"]
                out["plan_latency_c1"] = c1
            except Exception as exc:  # noqa: BLE001
                out["plan_latency_c1"] = {"error": repr(exc)}
        if also:
            out["also"] = also
        sys.stdout.flush()
        os.dup2(_saved_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
