"""MPPI — mirror of judo/optimizers/mppi.py:11-82."""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from judo_b200.optimizers.base import Optimizer, OptimizerConfig


@dataclass
class MPPIConfig(OptimizerConfig):
    """judo/optimizers/mppi.py:13-18."""

    sigma: float = 0.1
    temperature: float = 0.05


class MPPI(Optimizer[MPPIConfig]):
    """Softmin-weighted average of the candidates."""

    name = "mppi"
    sigma = property(lambda self: self.config.sigma)
    temperature = property(lambda self: self.config.temperature)

    def sample_control_knots(self, nominal_knots: np.ndarray) -> np.ndarray:
        sigma = self._ramp() * self.sigma if self.use_noise_ramp else self.sigma
        return self._noised(nominal_knots, sigma)

    def device_sigma(self) -> np.ndarray:
        sigma = self._ramp() * self.sigma if self.use_noise_ramp else np.full((self.num_nodes, 1), self.sigma)
        return np.broadcast_to(sigma, (self.num_nodes, self.nu)).copy()

    def update_nominal_knots(self, sampled_knots: np.ndarray, rewards: np.ndarray) -> np.ndarray:
        """w = softmax(rewards / temperature) (shifted by the best); nominal = sum_n w_n knots_n  — on the GPU."""
        return self._engine().update_mppi(sampled_knots, rewards, self.temperature)

    def fused_params(self) -> np.ndarray:
        return np.array([self.temperature], dtype=np.float64)
