"""Predictive sampling — mirror of judo/optimizers/ps.py:10-65."""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from judo_b200.optimizers.base import Optimizer, OptimizerConfig


@dataclass
class PredictiveSamplingConfig(OptimizerConfig):
    """judo/optimizers/ps.py:10-14."""

    sigma: float = 0.05


class PredictiveSampling(Optimizer[PredictiveSamplingConfig]):
    """Keep the single best candidate."""

    name = "ps"
    sigma = property(lambda self: self.config.sigma)

    def sample_control_knots(self, nominal_knots: np.ndarray) -> np.ndarray:
        sigma = self._ramp() * self.sigma if self.use_noise_ramp else self.sigma
        return self._noised(nominal_knots, sigma)

    def device_sigma(self) -> np.ndarray:
        sigma = self._ramp() * self.sigma if self.use_noise_ramp else np.full((self.num_nodes, 1), self.sigma)
        return np.broadcast_to(sigma, (self.num_nodes, self.nu)).copy()

    def update_nominal_knots(self, sampled_knots: np.ndarray, rewards: np.ndarray) -> np.ndarray:
        return self._engine().update_ps(sampled_knots, rewards)
