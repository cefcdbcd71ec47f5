"""Cartpole — mirror of judo/tasks/cartpole.py:19-84."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np

from judo_b200.tasks.base import Task, TaskConfig


@dataclass
class CartpoleConfig(TaskConfig):
    """judo/tasks/cartpole.py:19-28."""

    w_vertical: float = 10.0
    w_centered: float = 10.0
    w_velocity: float = 0.1
    w_control: float = 0.1
    p_vertical: float = 0.01
    p_centered: float = 0.1


class Cartpole(Task[CartpoleConfig]):
    name = "cartpole"
    config_t = CartpoleConfig

    def __init__(self) -> None:
        super().__init__("cartpole")
        self.reset()

    def cost_params(self, system_metadata: dict[str, Any] | None = None) -> np.ndarray:
        c = self.config
        return np.array([c.w_vertical, c.w_centered, c.w_velocity, c.w_control, c.p_vertical, c.p_centered], dtype=np.float64)

    def reward(self, states: np.ndarray, sensors: np.ndarray, controls: np.ndarray,
               system_metadata: dict[str, Any] | None = None) -> np.ndarray:
        """-(w_v sum sl1(cos th - 1) + w_c sum sl1(x) + w_vel sum quad(vel) + w_u sum quad(u))  (cartpole.py:66-78)."""
        return self._gpu_reward(states, controls, system_metadata)

    def reset(self) -> None:
        """Random state around the hanging pole (cartpole.py:80-84); same RNG draws as the reference."""
        self.data.qpos = np.array([1.0, np.pi]) + np.random.randn(2)
        self.data.qvel = 1e-1 * np.random.randn(2)
