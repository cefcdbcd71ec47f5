"""Controller — mirror of judo/controller/controller.py:34-442 around the B200 engine.

``update_action`` keeps the reference's sequence (controller.py:210-299): time-shift the nominal spline,
pre_optimization, then per iteration  sample -> clip -> (spline -> rollout -> reward -> update)  and finally the
spline / trace refresh.  The bracketed part is ONE fused call (Engine.plan_step: one H2D copy, the fused rollout+cost
kernel, the optimizer-update reduction, one D2H copy) whenever the task has a fused kernel; a user Task with its own
NumPy ``reward`` goes through the drop-in RolloutBackend contract instead (still GPU rollouts).
"""

from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Literal

import numpy as np

from judo_b200.config import OverridableConfig, set_config_overrides
from judo_b200.normalization import IdentityNormalizer, Normalizer, make_normalizer, normalizer_registry
import ctypes

from judo_b200 import _lib
from judo_b200.engine import OPT_IDS, SPLINE_IDS, legacy_stream
from judo_b200.optimizers import Optimizer, OptimizerConfig, fused_optimizer_ok, get_registered_optimizers
from judo_b200.rollout_backend import B200RolloutBackend, RolloutBackend
from judo_b200.spline import spline_basis
from judo_b200.tasks import Task, TaskConfig, fused_task_ok, get_registered_tasks


@dataclass
class ControllerConfig(OverridableConfig):
    """judo/controller/controller.py:34-42."""

    horizon: float = 1.0
    spline_order: Literal["zero", "linear", "cubic"] = "linear"
    control_freq: float = 20.0
    max_opt_iters: int = 1
    max_num_traces: int = 5
    action_normalizer: Literal["none", "min_max", "running"] = "none"


# judo/controller/overrides.py:7-44
set_config_overrides("cylinder_push", ControllerConfig, {"horizon": 1.0, "spline_order": "zero"})
set_config_overrides("cartpole", ControllerConfig, {"horizon": 1.0, "spline_order": "zero"})
set_config_overrides("leap_cube", ControllerConfig, {"horizon": 1.0, "spline_order": "cubic", "max_num_traces": 1})
set_config_overrides("leap_cube_down", ControllerConfig, {"horizon": 1.0, "spline_order": "cubic", "max_num_traces": 1})
# judo/controller/overrides.py:91-102
set_config_overrides("fr3_pick", ControllerConfig, {"horizon": 1.0, "spline_order": "linear", "max_num_traces": 3, "control_freq": 20.0})


class Spline:
    """Callable nominal spline: what make_spline(times, knots, order) returns in the reference (controller.py:382-401)."""

    def __init__(self, times: np.ndarray, knots: np.ndarray, order: str) -> None:
        self.times, self.knots, self.order = np.asarray(times, dtype=np.float64), np.asarray(knots, dtype=np.float64), order

    def __call__(self, query: np.ndarray | float) -> np.ndarray:
        q = np.atleast_1d(np.asarray(query, dtype=np.float64))
        if self.order == "zero" and self.knots.ndim == 2:  # previous-knot hold: a gather (same values as the basis product, 1.0 * knot)
            out = self.knots[np.maximum(np.searchsorted(self.times, q, side="right") - 1, 0)]
            return out[0] if np.ndim(query) == 0 else out
        out = np.einsum("hk,...kj->...hj", spline_basis(self.times, q, self.order), self.knots)
        return out[..., 0, :] if np.ndim(query) == 0 else out


def make_spline(times: np.ndarray, controls: np.ndarray, spline_order: str) -> Spline:
    return Spline(times, controls, spline_order)


class Controller:
    """The controller object (judo/controller/controller.py:45-380)."""

    def __init__(self, controller_config: ControllerConfig, task: Task, optimizer: Optimizer,
                 rollout_backend: Literal["b200"] = "b200", device: int = 0, devices: "list[int] | None" = None) -> None:
        self._controller_cfg = controller_config
        self.task = task
        self.optimizer = optimizer
        self.available_optimizers = get_registered_optimizers()
        self.available_tasks = get_registered_tasks()
        self.model = self.task.model
        # devices=[0, 1, ...]: the rollouts are sharded over several GPUs of this process behind the same backend / Controller surface
        self.rollout_backend: RolloutBackend = B200RolloutBackend(self.task, self.optimizer_cfg.num_rollouts, device=device, devices=devices)
        self.engine = self.rollout_backend.engine
        self.task.engine = self.engine
        self.optimizer.bind(self.engine)
        self._last_policy_output = None
        self.action_normalizer = self._init_action_normalizer()
        self.system_metadata: dict = {}
        N, H = self.optimizer_cfg.num_rollouts, self.num_timesteps
        self.states = np.zeros((N, H, self.model.nq + self.model.nv))
        self.current_state = np.concatenate([self.task.data.qpos, self.task.data.qvel])
        self.sensors = np.zeros((N, H, self.model.nsensordata))
        self.rollout_controls = np.zeros((N, H, self.task.nu))
        self.rewards = np.zeros((N,))
        self.reset()
        self.traces = None
        self.trace_sensors = [i for i, (tp, nm) in enumerate(zip(self.model.sensor_types, self.model.sensor_names)) if tp in ("framepos", "framepos_body") and "trace" in nm]
        self.num_trace_elites = min(self.max_num_traces, len(self.rewards))
        self.num_trace_sensors = len(self.trace_sensors)
        self.sensor_rollout_size = self.num_timesteps - 1
        self.all_traces_rollout_size = self.sensor_rollout_size * self.num_trace_sensors
        self.fused = True  # set False to force the contract-A path (rollout + Task.reward + optimizer update)
        # warp-per-rollout tasks: the fused kernel keeps every rollout's trace sensors so the elites need no second simulation
        self._engine_trace_width = self.engine.trace_width
        self._trace_capture = self.num_trace_sensors > 0 and self._engine_trace_width == 3 * self.num_trace_sensors
        if self._trace_capture:
            self.engine.set_trace_capture(True)
        # "host": np.random.randn exactly as the reference (seed parity).  "device": Philox inside the rollout kernel (perf mode:
        # same distribution, nothing but the nominal crosses PCIe; candidate_knots then holds only the elite candidates).
        self.sampling: Literal["host", "device"] = "host"
        self.device_seed = 0
        self._plan_counter = 0
        # One C call per optimisation iteration (b200mpc_controller_step: sampling from numpy's own stream, clip, spline basis, fused
        # kernel, traces) whenever optimizer, task hooks and normalizer are the built-ins; False forces the NumPy glue below.
        self.fast_path = True
        self._rq = _lib.StepRequest()

    # the candidates of the last plan step; after a fast-path step they are fetched from the engine's staging buffer on first use
    @property
    def candidate_knots(self) -> np.ndarray:
        if self._cand is None:
            self._cand = self.engine.last_candidates(*self._cand_shape)
        return self._cand

    @candidate_knots.setter
    def candidate_knots(self, value: np.ndarray) -> None:
        self._cand = value

    # ---- config views (controller.py:109-208) ----------------------------------------------------------------
    horizon = property(lambda self: self.controller_cfg.horizon)
    nu = property(lambda self: self.task.nu)
    max_num_traces = property(lambda self: self.controller_cfg.max_num_traces)
    max_opt_iters = property(lambda self: self.controller_cfg.max_opt_iters)
    spline_order = property(lambda self: self.controller_cfg.spline_order)
    action_normalizer_type = property(lambda self: self.controller_cfg.action_normalizer)

    @property
    def num_timesteps(self) -> int:
        return int(np.ceil(self.horizon / self.task.dt))

    # (both grids only change with the config; they are rebuilt on a key change instead of on every plan step — at the reference's
    # sizes update_action() is dominated by host-side NumPy call overhead, not by the GPU)
    @property
    def rollout_times(self) -> np.ndarray:
        key = (self.task.dt, self.num_timesteps)
        if getattr(self, "_rt_key", None) != key:
            self._rt_key, self._rt = key, self.task.dt * np.arange(self.num_timesteps)
            self._rt.flags.writeable = False
        return self._rt

    @property
    def spline_timesteps(self) -> np.ndarray:
        key = (self.horizon, self.optimizer_cfg.num_nodes)
        if getattr(self, "_st_key", None) != key:
            self._st_key, self._st = key, np.linspace(0, self.horizon, self.optimizer_cfg.num_nodes, endpoint=True)
            self._st.flags.writeable = False
        return self._st

    @property
    def optimizer_cfg(self) -> OptimizerConfig:
        return self.optimizer.config

    @optimizer_cfg.setter
    def optimizer_cfg(self, cfg: OptimizerConfig) -> None:
        self.optimizer.config = cfg

    @property
    def task_config(self) -> TaskConfig:
        return self.task.config

    @task_config.setter
    def task_config(self, cfg: TaskConfig) -> None:
        self.task.config = cfg

    @property
    def time(self) -> float:
        return self.task.time

    @time.setter
    def time(self, value: float) -> None:
        self.task.time = value

    @property
    def controller_cfg(self) -> ControllerConfig:
        return self._controller_cfg

    @controller_cfg.setter
    def controller_cfg(self, cfg: ControllerConfig) -> None:
        self._controller_cfg = cfg
        self.action_normalizer = self._init_action_normalizer()

    # ---- the plan step (controller.py:210-299) -----------------------------------------------------------------
    def _can_fuse(self) -> bool:
        """The fused kernel restates the BUILT-IN reward / hooks / update: a user subclass that overrides any of them takes the
        contract-A path (GPU rollouts + its own Python code), exactly as the reference would run it."""
        key = (type(self.task), type(self.optimizer))
        if getattr(self, "_fuse_key", None) != key:  # class identities only change when task / optimizer are swapped
            self._fuse_key = key
            self._fuse_ok = self.optimizer.name in OPT_IDS and fused_optimizer_ok(self.optimizer) and fused_task_ok(self.task)
            self._fast_ok = self._fuse_ok and fused_optimizer_ok(self.optimizer, sampling=True)
        return self.fused and self._fuse_ok and self.task.cost_params(self.system_metadata) is not None

    def _can_fast_path(self) -> bool:
        opt = self.optimizer
        k = max(min(self.max_num_traces, opt.num_rollouts), opt.num_elites if opt.name == "cem" else 0)
        return self.fast_path and getattr(self.engine, "supports_controller_step", False) and self.sampling == "host" and \
            type(self.action_normalizer) is IdentityNormalizer and k <= 8 and opt.num_nodes <= 12 and legacy_stream().ok and \
            self._can_fuse() and self._fast_ok and (self.num_trace_sensors == 0 or self._engine_trace_width == 0 or self._trace_capture)

    def _fast_iteration(self, nominal: np.ndarray, new_times: np.ndarray) -> np.ndarray:
        """One iteration of the optimisation loop through b200mpc_controller_step; returns the new nominal knots."""
        opt, rq = self.optimizer, self._rq
        N, K, H, nu = opt.num_rollouts, opt.num_nodes, self.num_timesteps, self.model.nu
        sigma = opt.device_sigma()        # (K, nu) contiguous; CEM: applies the ramp's mutation exactly as sample_control_knots does
        self.task.pre_rollout(self.current_state)
        ne = min(self.max_num_traces, N)
        nts = self.num_trace_sensors
        x0 = np.ascontiguousarray(self.current_state, dtype=np.float64)
        nominal = np.ascontiguousarray(nominal, dtype=np.float64)
        new_times = np.ascontiguousarray(new_times, dtype=np.float64)
        params = np.ascontiguousarray(self.task.cost_params(self.system_metadata), dtype=np.float64)
        fparams = opt.fused_params()
        fparams = np.ascontiguousarray(fparams, dtype=np.float64) if np.size(fparams) else np.zeros(1)
        if getattr(self, "_fast_task", None) is not self.task:  # per-task constants: clip range, trace columns
            self._fast_task = self.task
            rng = self.task.actuator_ctrlrange
            self._fast_lo, self._fast_hi = np.ascontiguousarray(rng[:, 0], dtype=np.float64), np.ascontiguousarray(rng[:, 1], dtype=np.float64)
            self._trace_cols32 = np.array([int(self.model.sensor_adr[s]) + p for s in self.trace_sensors for p in range(3)] or [0], dtype=np.int32)
            rq.lo, rq.hi, rq.trace_cols = self._fast_lo.ctypes.data, self._fast_hi.ctypes.data, self._trace_cols32.ctypes.data
        out_nom = np.empty((K, nu))
        out_sig = np.empty((K, nu)) if opt.name == "cem" else None
        rewards = np.empty(N)
        elite = np.empty(8, dtype=np.int32)
        traces = np.empty((ne * nts * (H - 1), 2, 3)) if nts and ne else None
        basis = np.empty((H, K))
        rq.N, rq.K, rq.H, rq.optimizer, rq.spline_order, rq.n_elite = N, K, H, OPT_IDS[opt.name], SPLINE_IDS[self.spline_order], ne
        rq.n_trace_sensors = nts if traces is not None else 0
        rq.time, rq.dt = self.time, self.task.dt
        rq.knot_times, rq.x0, rq.nominal, rq.sigma = new_times.ctypes.data, x0.ctypes.data, nominal.ctypes.data, sigma.ctypes.data
        rq.cost_params, rq.opt_params = params.ctypes.data, fparams.ctypes.data
        rq.nominal_out, rq.sigma_out, rq.rewards = out_nom.ctypes.data, (out_sig.ctypes.data if out_sig is not None else None), rewards.ctypes.data
        rq.elite_idx, rq.traces, rq.basis_out, rq.knots_out = elite.ctypes.data, (traces.ctypes.data if traces is not None else None), basis.ctypes.data, None
        self.engine.controller_step(rq, legacy_stream(), (N - 1) * K * nu)
        self.rewards, self._elite, self._basis = rewards, elite[:ne], basis
        self._cand, self._cand_shape = None, (N, K)
        self._fast_traces = traces
        self._rollout_cache_valid = False
        return opt.accept_fused({"nominal": out_nom, "sigma": out_sig})

    def update_action(self) -> None:
        assert self.current_state.shape == (self.model.nq + self.model.nv,), "Current state must be of shape (nq + nv,)"
        assert self.optimizer_cfg.num_rollouts > 0, "Need at least one rollout!"
        if self.optimizer_cfg.num_nodes < 4 and self.spline_order == "cubic":
            warnings.warn("Cubic splines require at least 4 nodes. Setting num_nodes=4.", stacklevel=2)
            self.optimizer_cfg.num_nodes = 4

        # time-shift the previous plan onto the new knot times
        new_times = self.time + self.spline_timesteps
        nominal_knots = self.spline(new_times)
        nominal_knots_normalized = self.action_normalizer.normalize(nominal_knots)

        if self.rollout_backend.num_threads != self.optimizer_cfg.num_rollouts:
            self.rollout_backend.update(self.optimizer_cfg.num_rollouts)

        normalizer_cls = normalizer_registry.get(self.action_normalizer_type)
        if normalizer_cls is None:
            warnings.warn(f"Invalid action normalizer type '{self.action_normalizer_type}'. Falling back to 'none'.", stacklevel=2)
            normalizer_cls = IdentityNormalizer
        if not isinstance(self.action_normalizer, normalizer_cls):
            self.action_normalizer = self._init_action_normalizer()

        self.optimizer.pre_optimization(self.times, new_times)

        fast = self._can_fast_path()
        basis = None
        if not fast:
            query = self.time + self.rollout_times
            basis = spline_basis(new_times, query, self.spline_order)  # (H, K): controls = basis @ knots
        if not fast:
            lo = self.action_normalizer.normalize(self.task.actuator_ctrlrange[:, 0])
            hi = self.action_normalizer.normalize(self.task.actuator_ctrlrange[:, 1])
        self._rollout_cache_valid = False
        self._fast_traces = None
        i = 0
        while i < self.max_opt_iters and not self.optimizer.stop_cond():
            if fast:
                nominal_knots_normalized = self._fast_iteration(nominal_knots_normalized, new_times)
                i += 1
                continue
            if self.sampling == "device" and self._can_fuse() and isinstance(self.action_normalizer, IdentityNormalizer):
                self.task.pre_rollout(self.current_state)
                ne = min(self.max_num_traces, self.optimizer_cfg.num_rollouts)
                res = self.engine.plan_step_sampled(self.current_state, nominal_knots_normalized, self.optimizer.device_sigma(), lo, hi,
                                                    self.optimizer_cfg.num_rollouts, basis, self.task.cost_params(self.system_metadata),
                                                    self.optimizer.name, self.optimizer.fused_params(), self.device_seed, self._plan_counter,
                                                    want_rewards=True, n_elite=ne)
                self._plan_counter += 1
                self.rewards = res["rewards"]
                self.candidate_knots = res["elite_knots"]       # only the elite candidates leave the GPU
                self._elite = np.arange(len(res["elite"]))       # ... so they are rows 0..ne-1 of candidate_knots
                self.elite_indices_global = res["elite"]
                nominal_knots_normalized = self.optimizer.accept_fused(res)
                self._basis = basis
                self._rollout_cache_valid = False
                i += 1
                continue
            cand_norm = np.clip(self.optimizer.sample_control_knots(nominal_knots_normalized), lo, hi)
            self.candidate_knots = self.action_normalizer.denormalize(cand_norm)
            self.task.pre_rollout(self.current_state)
            if self._can_fuse():
                # The kernel rolls out (and updates over) the DENORMALISED candidates.  Normalizers are affine per control
                # dimension (normalization.py), and every update is a convex combination of candidates, so mapping the
                # resulting nominal back equals updating in normalised space; CEM's sigma scales by 1/|slope| before its clip.
                identity = isinstance(self.action_normalizer, IdentityNormalizer)
                fparams = self.optimizer.fused_params()
                if not identity and self.optimizer.name == "cem":
                    fparams = np.array([fparams[0], 0.0, np.inf])
                res = self.engine.plan_step(self.current_state, self.candidate_knots, basis, self.task.cost_params(self.system_metadata),
                                            self.optimizer.name, fparams, want_rewards=True,
                                            n_elite=min(self.max_num_traces, self.optimizer_cfg.num_rollouts))
                if not identity:
                    # exact inverse of denormalize (normalize() is NOT it: the reference's eps terms make the pair asymmetric)
                    nu = self.model.nu
                    offset = self.action_normalizer.denormalize(np.zeros(nu))
                    slope = self.action_normalizer.denormalize(np.ones(nu)) - offset
                    res = dict(res)
                    res["nominal"] = (res["nominal"] - offset) / slope
                    if self.optimizer.name == "cem":
                        res["sigma"] = np.clip(res["sigma"] / np.abs(slope), self.optimizer.sigma_min, self.optimizer.sigma_max)
                self.rewards = res["rewards"]
                self._elite = res["elite"]
                nominal_knots_normalized = self.optimizer.accept_fused(res)
                self._basis = basis
                self._rollout_cache_valid = False
            else:
                self.rollout_controls = np.einsum("hk,nkj->nhj", basis, self.candidate_knots)
                sim_controls = self.task.task_to_sim_ctrl(self.rollout_controls)
                self.states, self.sensors, _ = self.rollout_backend.rollout(self.current_state, sim_controls, self._last_policy_output)
                self.task.post_rollout(self.states, self.sensors, self.rollout_controls, self.system_metadata)
                self.rewards = self.task.reward(self.states, self.sensors, self.rollout_controls, self.system_metadata)
                nominal_knots_normalized = self.optimizer.update_nominal_knots(cand_norm, self.rewards)
                self._elite = None
                self._rollout_cache_valid = True
            self.action_normalizer.update(self.candidate_knots)
            i += 1

        self.nominal_knots = self.action_normalizer.denormalize(nominal_knots_normalized)
        self.times = new_times
        self.update_spline(self.times, self.nominal_knots)
        self.update_traces()

    def action(self, time: float) -> np.ndarray:
        return self.spline(time)

    def update_spline(self, times: np.ndarray, controls: np.ndarray) -> None:
        self.spline = make_spline(times, controls, self.spline_order)

    def reset(self) -> None:
        """controller.py:309-321."""
        self.task.reset()
        if self.optimizer_cfg.num_nodes < 4 and self.spline_order == "cubic":
            warnings.warn("Cubic splines require at least 4 nodes. Setting num_nodes=4.", stacklevel=2)
            self.optimizer_cfg.num_nodes = 4
        self.nominal_knots = np.tile(self.task.optimizer_warm_start(), (self.optimizer_cfg.num_nodes, 1))
        self.candidate_knots = np.tile(self.nominal_knots, (self.optimizer_cfg.num_rollouts, 1, 1))
        self.times = self.task.data.time + self.spline_timesteps
        self.update_spline(self.times, self.nominal_knots)

    def update_traces(self) -> None:
        """Elite rollouts' trace sensors as line segments (controller.py:323-363).

        In fused mode the sensors of the <= max_num_traces elite rollouts are recomputed by a tiny contract-A rollout
        of just those candidates instead of materialising (N, H, nsensordata) for everybody."""
        self.sensor_rollout_size = self.num_timesteps - 1
        self.all_traces_rollout_size = self.sensor_rollout_size * self.num_trace_sensors
        self.num_trace_elites = min(self.max_num_traces, self.optimizer_cfg.num_rollouts)
        ne, nts, size = self.num_trace_elites, self.num_trace_sensors, self.sensor_rollout_size
        if ne == 0 or nts == 0:  # max_num_traces = 0 (or a task without trace sensors): no segments, as the reference's empty gather
            self.elite_indices = np.zeros(0, dtype=np.int64)
            self.traces = np.zeros((0, 2, 3))
            return
        if getattr(self, "_fast_traces", None) is not None:  # the fast path's C call already assembled the segments
            self.elite_indices = np.asarray(self._elite[:ne])
            self.traces = self._fast_traces
            return
        if getattr(self, "_rollout_cache_valid", False) or getattr(self, "_elite", None) is None:
            elite = np.argsort(self.rewards)[-ne:][::-1]
            elite_sensors = self.sensors[elite]
        elif self._trace_capture and self.sampling == "host":
            elite = np.asarray(self._elite[:ne])
            self.elite_indices = elite
            self.traces = self._segments(self.engine.elite_traces(elite, self.num_timesteps), ne, nts, size)  # (ne, H, 3 nts), sensor order
            return
        else:
            elite = np.asarray(self._elite[:ne])
            ctrl = np.einsum("hk,nkj->nhj", self._basis, self.candidate_knots[elite])
            N = self.engine.num_rollouts
            self.engine.update(len(elite))
            _, elite_sensors = self.engine.rollout(self.current_state, self.task.task_to_sim_ctrl(ctrl), want_sensors=True)
            self.engine.update(N)
        self.elite_indices = elite
        if getattr(self, "_trace_cols", None) is None:
            self._trace_cols = np.array([int(self.model.sensor_adr[s]) + p for s in self.trace_sensors for p in range(3)], dtype=np.intp)
        self.traces = self._segments(elite_sensors[:, :, self._trace_cols], ne, nts, size)

    @staticmethod
    def _segments(tr: np.ndarray, ne: int, nts: int, size: int) -> np.ndarray:
        """(ne, H, 3 nts) trace-sensor positions -> (ne * nts * (H - 1), 2, 3) line segments in the reference's order
        (controller.py:341-363: consecutive positions of each sensor, elites interleaved with sensors)."""
        out = np.empty((nts * ne, size, 2, 3))
        for s in range(nts):
            blk = tr[:, :, 3 * s:3 * s + 3]
            out[s::nts, :, 0, :] = blk[:, :-1]
            out[s::nts, :, 1, :] = blk[:, 1:]
        return out.reshape(ne * nts * size, 2, 3)

    def update_states(self, state_msg) -> None:  # noqa: ANN001
        """state_msg: any object with qpos, qvel, time, sim_metadata (judo/app/structs.py:30-41)."""
        self.current_state = np.concatenate([state_msg.qpos, state_msg.qvel])
        self.time = state_msg.time
        self.system_metadata = state_msg.sim_metadata

    def _init_action_normalizer(self) -> Normalizer:
        kwargs = {}
        if self.action_normalizer_type == "min_max":
            kwargs = dict(min=self.task.actuator_ctrlrange[:, 0], max=self.task.actuator_ctrlrange[:, 1])
        elif self.action_normalizer_type == "running":
            kwargs = dict(init_std=1.0)
        return make_normalizer(self.action_normalizer_type, self.model.nu, **kwargs)


def make_controller(init_task: str, init_optimizer: str, rollout_backend: Literal["b200"] = "b200", device: int = 0,
                    devices: "list[int] | None" = None) -> Controller:
    """judo/controller/controller.py:404-442."""
    task_entry = get_registered_tasks().get(init_task)
    optimizer_entry = get_registered_optimizers().get(init_optimizer)
    assert task_entry is not None, f"Task {init_task} not found in task registry."
    assert optimizer_entry is not None, f"Optimizer {init_optimizer} not found in optimizer registry."
    task = task_entry[0]()
    optimizer_cls, optimizer_config_cls = optimizer_entry
    optimizer_cfg = optimizer_config_cls()
    optimizer_cfg.set_override(init_task)
    optimizer = optimizer_cls(optimizer_cfg, task.nu)
    controller_cfg = ControllerConfig()
    controller_cfg.set_override(init_task)
    return Controller(controller_config=controller_cfg, task=task, optimizer=optimizer, rollout_backend=rollout_backend, device=device,
                      devices=devices)
