"""Cylinder push — mirror of judo/tasks/cylinder_push.py:17-108."""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import numpy as np

from judo_b200.tasks.base import Task, TaskConfig


@dataclass
class CylinderPushConfig(TaskConfig):
    """judo/tasks/cylinder_push.py:17-35."""

    w_pusher_proximity: float = 0.5
    w_pusher_velocity: float = 0.0
    w_cart_position: float = 0.1
    pusher_goal_offset: float = 0.25
    goal_pos: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0]))


class CylinderPush(Task[CylinderPushConfig]):
    name = "cylinder_push"
    config_t = CylinderPushConfig

    def __init__(self) -> None:
        super().__init__("cylinder_push")
        self.reset()

    def cost_params(self, system_metadata: dict[str, Any] | None = None) -> np.ndarray:
        c = self.config
        return np.array([c.w_pusher_proximity, c.w_pusher_velocity, c.w_cart_position, c.pusher_goal_offset, c.goal_pos[0], c.goal_pos[1]],
                        dtype=np.float64)

    def reward(self, states: np.ndarray, sensors: np.ndarray, controls: np.ndarray,
               system_metadata: dict[str, Any] | None = None) -> np.ndarray:
        """Pusher-behind-the-cart proximity + pusher velocity + cart-to-goal terms, summed over time (cylinder_push.py:69-93)."""
        return self._gpu_reward(states, controls, system_metadata)

    def reset(self) -> None:
        """Pusher on the unit circle, cart on the radius-2 circle (cylinder_push.py:95-108)."""
        theta = 2 * np.pi * np.random.rand(2)
        self.data.qpos = np.array([np.cos(theta[0]), np.sin(theta[0]), 2 * np.cos(theta[1]), 2 * np.sin(theta[1])])
        self.data.qvel = np.zeros(4)
