"""Wire structs either side of the planner — mirror of judo/app/structs.py:30-84."""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import numpy as np

from judo_b200.controller import Spline, make_spline


@dataclass
class MujocoState:
    """State in: what the simulation node sends to the controller (judo/app/structs.py:30-41)."""

    time: float
    qpos: np.ndarray
    qvel: np.ndarray
    xpos: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    xquat: np.ndarray = field(default_factory=lambda: np.zeros((0, 4)))
    mocap_pos: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    mocap_quat: np.ndarray = field(default_factory=lambda: np.zeros((0, 4)))
    sim_metadata: dict[str, Any] = field(default_factory=dict)


@dataclass
class SplineData:
    """Spline out: (times, knots, kind) as the controller publishes it (judo/app/structs.py:58-84)."""

    t: np.ndarray
    x: np.ndarray
    kind: str = "zero"
    extrapolate: bool = True

    def spline(self) -> Spline:
        return make_spline(self.t, self.x, self.kind)
