"""Rollout backends — the drop-in boundary.

``RolloutBackend`` mirrors judo/utils/rollout_backend.py:10-46.  ``B200RolloutBackend`` has the contract of
MJRolloutBackend (judo/utils/mj_rollout_backend.py:15-101): same arguments, shapes, dtypes, return tuple and
error conditions, but the N x H mj_steps run in one fused CUDA kernel.  Install it in a judo Controller with
``controller.rollout_backend = B200RolloutBackend("cartpole", controller.optimizer_cfg.num_rollouts)``
(the Controller has no injection hook, controller.py:53,72-85).
"""

from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from judo_b200.engine import Engine, MultiEngine


class RolloutBackend(ABC):
    """Abstract rollout backend (judo/utils/rollout_backend.py:10-46)."""

    num_threads: int

    @abstractmethod
    def rollout(self, x0: np.ndarray, controls: np.ndarray, last_policy_output: np.ndarray | None = None
                ) -> tuple[np.ndarray, np.ndarray, np.ndarray | None]:
        """x0 (nq+nv,) or (N, nq+nv); controls (N, H, nu) -> states (N, H, nq+nv), sensors (N, H, ns), None."""

    @abstractmethod
    def update(self, num_threads: int) -> None:
        """Change the number of parallel rollouts."""


class B200RolloutBackend(RolloutBackend):
    """GPU rollouts behind the reference's backend contract."""

    def __init__(self, model: "str | object", num_threads: int, device: int = 0, devices: "list[int] | None" = None) -> None:
        """``model``: a task name ("cartpole", "cylinder_push", "leap_cube", "fr3_pick"), a judo_b200 Task, or an Engine.
        ``devices``: several GPUs of this process behind the one backend (rollouts sharded along N)."""
        if isinstance(model, Engine):
            self.engine = model
            self.engine.update(num_threads)
        else:
            name = model if isinstance(model, str) else getattr(model, "name", None)
            if not isinstance(name, str):
                raise ValueError("B200RolloutBackend needs a task name, a judo_b200 Task or an Engine")
            self.engine = MultiEngine(name, num_threads, devices) if devices is not None and len(devices) > 1 else \
                Engine(name, num_threads, device=(devices[0] if devices else device))
        self.num_threads = num_threads

    def rollout(self, x0: np.ndarray, controls: np.ndarray, last_policy_output: np.ndarray | None = None
                ) -> tuple[np.ndarray, np.ndarray, np.ndarray | None]:
        x0 = np.asarray(x0, dtype=np.float64)
        controls = np.asarray(controls, dtype=np.float64)
        nx = self.engine.nq + self.engine.nv
        # the reference's asserts (mj_rollout_backend.py:78-82) as exceptions
        if x0.ndim not in (1, 2) or x0.shape[-1] != nx:
            raise ValueError(f"x0 must have shape ({nx},) or (N, {nx}); got {x0.shape}")
        if controls.ndim != 3 or controls.shape[-1] != self.engine.nu:
            raise ValueError(f"controls must have shape (N, H, {self.engine.nu}); got {controls.shape}")
        if controls.shape[0] != self.num_threads or (x0.ndim == 2 and x0.shape[0] != controls.shape[0]):
            raise ValueError("controls / x0 batch size must equal num_threads")
        states, sensors = self.engine.rollout(x0, controls, want_sensors=True)
        return states, sensors, None

    def update(self, num_threads: int) -> None:
        self.num_threads = num_threads
        self.engine.update(num_threads)
