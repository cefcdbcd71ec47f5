"""Task plugin surface — mirror of judo/tasks/base.py:24-204 without MuJoCo objects.

``model`` / ``data`` are light stand-ins exposing the attributes the Controller and user code read
(nq, nv, nu, nsensordata, opt.timestep, actuator_ctrlrange, sensor_adr; qpos, qvel, ctrl, time, mocap_quat).
Built-in tasks evaluate ``reward`` on the GPU (their per-step cost is also fused into the rollout kernel);
a user-defined Task simply overrides ``reward`` with NumPy code and runs through contract A.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Any, Generic, TypeVar

import numpy as np

from judo_b200.consts import load_table


@dataclass
class TaskConfig:
    """judo/tasks/base.py:16-18."""


ConfigT = TypeVar("ConfigT", bound=TaskConfig)


def _model_info(table: dict) -> SimpleNamespace:
    acts = table["actuators"]
    return SimpleNamespace(
        nq=table["nq"], nv=table["nv"], nu=table["nu"], nsensordata=table["nsensordata"], nsensor=len(table["sensors"]),
        nbody=table["nbody"], opt=SimpleNamespace(timestep=table["opt"]["timestep"]),
        actuator_ctrlrange=np.array([a["ctrlrange"] for a in acts], dtype=np.float64).reshape(len(acts), 2),
        actuator_ctrllimited=np.array([a["ctrllimited"] for a in acts], dtype=bool),
        sensor_adr=np.array([s["adr"] for s in table["sensors"]], dtype=np.int64),
        sensor_names=[s["name"] for s in table["sensors"]], sensor_types=[s["type"] for s in table["sensors"]],
        joint_names=[j["name"] for j in table["joints"]], jnt_qposadr=np.array([j["qposadr"] for j in table["joints"]]),
        jnt_dofadr=np.array([j["dofadr"] for j in table["joints"]]), table=table,
    )


class Task(ABC, Generic[ConfigT]):
    """Task definition (judo/tasks/base.py:24-204)."""

    name: str = ""
    config_t: type

    def __init__(self, model_table: str = "") -> None:
        if not model_table:
            raise ValueError("Model path must be provided.")  # same condition as tasks/base.py:31-32
        self.config = self.config_t()
        self.table = load_table(model_table)
        self.model = _model_info(self.table)
        self.data = SimpleNamespace(qpos=np.array(self.table["qpos0"], dtype=np.float64), qvel=np.zeros(self.table["nv"]),
                                    ctrl=np.zeros(self.table["nu"]), time=0.0, mocap_quat=np.tile([1.0, 0, 0, 0], (1, 1)))
        self.engine = None  # bound by the Controller / backend that owns the GPU handle

    # -- reference properties -------------------------------------------------------------------------------
    @property
    def time(self) -> float:
        return self.data.time

    @time.setter
    def time(self, value: float) -> None:
        self.data.time = value

    @property
    def nu(self) -> int:
        return self.model.nu

    @property
    def uses_locomotion_policy(self) -> bool:
        return False

    @property
    def physics_substeps(self) -> int:
        return 1

    @property
    def dt(self) -> float:
        return self.model.opt.timestep * self.physics_substeps

    @property
    def actuator_ctrlrange(self) -> np.ndarray:
        """Actuator limits; unlimited actuators get (-inf, inf) (tasks/base.py:97-103)."""
        limits = self.model.actuator_ctrlrange
        limits[~self.model.actuator_ctrllimited] = np.array([-np.inf, np.inf])
        return limits

    # -- reward ----------------------------------------------------------------------------------------------
    @abstractmethod
    def reward(self, states: np.ndarray, sensors: np.ndarray, controls: np.ndarray,
               system_metadata: dict[str, Any] | None = None) -> np.ndarray:
        """states (N,T,nq+nv), sensors (N,T,ns), controls (N,T,nu) -> rewards (N,)  (tasks/base.py:55-76)."""

    def cost_params(self, system_metadata: dict[str, Any] | None = None) -> np.ndarray | None:
        """Weights vector for the fused kernel; None means "no fused kernel: use rollout + reward()"."""
        return None

    def _gpu_reward(self, states: np.ndarray, controls: np.ndarray, system_metadata: dict[str, Any] | None) -> np.ndarray:
        if self.engine is None:
            from judo_b200.engine import Engine
            self.engine = Engine(self.name, max(1, len(states)))
        return self.engine.reward(states, controls, self.cost_params(system_metadata))

    # -- hooks (no-ops by default, tasks/base.py:121-176) -------------------------------------------------------
    def reset(self) -> None:
        self.data.qpos = np.zeros_like(self.data.qpos)
        self.data.qvel = np.zeros_like(self.data.qvel)

    def pre_rollout(self, curr_state: np.ndarray) -> None: ...
    def post_rollout(self, states: np.ndarray, sensors: np.ndarray, controls: np.ndarray,
                     system_metadata: dict[str, Any] | None = None) -> None: ...
    def pre_sim_step(self) -> None: ...
    def post_sim_step(self) -> None: ...

    def get_sim_metadata(self) -> dict[str, Any]:
        return {}

    def optimizer_warm_start(self) -> np.ndarray:
        return np.zeros(self.nu)

    def task_to_sim_ctrl(self, controls: np.ndarray) -> np.ndarray:
        return controls

    # -- index helpers (tasks/base.py:178-204) -------------------------------------------------------------------
    def get_sensor_start_index(self, sensor_name: str) -> int:
        return int(self.model.sensor_adr[self.model.sensor_names.index(sensor_name)])

    def get_joint_position_start_index(self, joint_name: str) -> int:
        return int(self.model.jnt_qposadr[self.model.joint_names.index(joint_name)])

    def get_joint_velocity_start_index(self, joint_name: str) -> int:
        return int(self.model.nq + self.model.jnt_dofadr[self.model.joint_names.index(joint_name)])
