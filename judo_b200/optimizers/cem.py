"""Cross-entropy method — mirror of judo/optimizers/cem.py:11-92."""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from judo_b200.optimizers.base import Optimizer, OptimizerConfig


@dataclass
class CrossEntropyMethodConfig(OptimizerConfig):
    """judo/optimizers/cem.py:11-17."""

    sigma_min: float = 0.1
    sigma_max: float = 1.0
    num_elites: int = 2


class CrossEntropyMethod(Optimizer[CrossEntropyMethodConfig]):
    """Top-k elites: nominal = mean(elites), sigma = clip(std(elites)).  sigma is stateful (K, nu)."""

    name = "cem"

    def __init__(self, config: CrossEntropyMethodConfig, nu: int) -> None:
        super().__init__(config, nu)
        self.sigma = 0.5 * (self.sigma_min + self.sigma_max) * np.ones((config.num_nodes, nu))

    sigma_min = property(lambda self: self.config.sigma_min)
    sigma_max = property(lambda self: self.config.sigma_max)
    num_elites = property(lambda self: self.config.num_elites)

    def pre_optimization(self, old_times: np.ndarray, new_times: np.ndarray) -> None:
        """Linear re-interpolation (with extrapolation) of sigma when num_nodes changed (cem.py:44-53)."""
        if len(self.sigma) != self.num_nodes:
            t = np.asarray(old_times, dtype=np.float64)
            q = np.asarray(new_times, dtype=np.float64)
            seg = np.clip(np.searchsorted(t, q, side="right") - 1, 0, len(t) - 2)
            w = ((q - t[seg]) / (t[seg + 1] - t[seg]))[:, None]
            self.sigma = (1 - w) * self.sigma[seg] + w * self.sigma[seg + 1]

    def sample_control_knots(self, nominal_knots: np.ndarray) -> np.ndarray:
        if self.use_noise_ramp:
            # the ramp MUTATES sigma on every call (cem.py:70-72)
            self.sigma = np.clip(self.sigma * self._ramp(), self.sigma_min, self.sigma_max)
        return self._noised(nominal_knots, self.sigma[None])

    def device_sigma(self) -> np.ndarray:
        if self.use_noise_ramp:
            self.sigma = np.clip(self.sigma * self._ramp(), self.sigma_min, self.sigma_max)  # same mutation as sample_control_knots
        return self.sigma.copy()

    def update_nominal_knots(self, sampled_knots: np.ndarray, rewards: np.ndarray) -> np.ndarray:
        nominal, self.sigma = self._engine().update_cem(sampled_knots, rewards, self.num_elites, self.sigma_min, self.sigma_max)
        return nominal

    def fused_params(self) -> np.ndarray:
        return np.array([self.num_elites, self.sigma_min, self.sigma_max], dtype=np.float64)

    def accept_fused(self, result: dict) -> np.ndarray:
        self.sigma = result["sigma"]
        return result["nominal"]

