"""Plant side — mirror of judo/simulation/base.py:13-60 and judo/simulation/mj_simulation.py:13-66.

``B200Simulation.step`` is one ``mj_step`` of the plant, executed by the same CUDA dynamics as the planner (a contract-A
rollout with N = 1, H = 1), so the whole sim -> plan -> act loop runs without MuJoCo for the supported tasks.  The reference may
load a different "sim" MJCF for the plant (e.g. leap_cube_sim.xml); here the plant uses the controller's table.  Each call
starts the constraint solver from a zero warm start (the reference's plant carries qacc_warmstart over; the converged
solution is the same to solver tolerance).
"""

from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from judo_b200.engine import Engine
from judo_b200.structs import MujocoState
from judo_b200.tasks import Task, get_registered_tasks


class Simulation(ABC):
    """judo/simulation/base.py:13-60."""

    def __init__(self, init_task: str = "cylinder_push") -> None:
        self.paused = False
        self.set_task(init_task)

    def set_task(self, task_name: str) -> None:
        entry = get_registered_tasks().get(task_name)
        if entry is None:
            raise ValueError(f"Task {task_name} not found in task registry")
        self.task: Task = entry[0]()
        self.task.reset()

    @abstractmethod
    def step(self, command: np.ndarray) -> None: ...

    def pause(self) -> None:
        self.paused = not self.paused

    @property
    @abstractmethod
    def timestep(self) -> float: ...


class B200Simulation(Simulation):
    """MJSimulation with the GPU dynamics as the plant."""

    def __init__(self, init_task: str = "cylinder_push", device: int = 0) -> None:
        self._device = device
        self._engine: Engine | None = None
        super().__init__(init_task)

    def set_task(self, task_name: str) -> None:
        super().set_task(task_name)
        if self._engine is not None:
            self._engine.close()
        self._engine = Engine(self.task.name, 1, device=self._device)

    def step(self, command: np.ndarray) -> None:
        """ctrl <- task_to_sim_ctrl(command); pre_sim_step; mj_step; post_sim_step (mj_simulation.py:33-46)."""
        if self.paused:
            return
        command = self.task.task_to_sim_ctrl(np.asarray(command, dtype=np.float64))
        self.task.data.ctrl = command[: self.task.model.nu].copy()
        self.task.pre_sim_step()
        x = np.concatenate([self.task.data.qpos, self.task.data.qvel])
        states, _ = self._engine.rollout(x, self.task.data.ctrl.reshape(1, 1, -1), want_sensors=False)
        nq = self.task.model.nq
        self.task.data.qpos, self.task.data.qvel = states[0, 0, :nq].copy(), states[0, 0, nq:].copy()
        self.task.data.time += self.timestep
        self.task.post_sim_step()

    @property
    def sim_state(self) -> MujocoState:
        d = self.task.data
        return MujocoState(time=d.time, qpos=d.qpos, qvel=d.qvel, mocap_quat=d.mocap_quat, sim_metadata=self.task.get_sim_metadata())

    @property
    def timestep(self) -> float:
        return self.task.model.opt.timestep
