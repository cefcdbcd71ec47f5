"""Optimizer plugin surface — mirror of judo/optimizers/base.py:13-96 (same names, argument meaning, shapes).

Sampling stays on the host in "seed-parity" mode: the reference draws from NumPy's global legacy RNG
(np.random.randn, judo/optimizers/mppi.py:58), so identical seeds give identical candidates.  The nominal update is
a reduction over all rollouts and runs on the GPU through the C-ABI (``engine``).
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import TYPE_CHECKING, Generic, TypeVar

import numpy as np

from judo_b200.config import OverridableConfig

if TYPE_CHECKING:
    from judo_b200.engine import Engine


@dataclass
class OptimizerConfig(OverridableConfig):
    """judo/optimizers/base.py:13-21."""

    num_rollouts: int = 16
    num_nodes: int = 4
    use_noise_ramp: bool = False
    noise_ramp: float = 2.5


ConfigT = TypeVar("ConfigT", bound=OptimizerConfig)


class Optimizer(ABC, Generic[ConfigT]):
    """judo/optimizers/base.py:27-96."""

    name: str = ""

    def __init__(self, config: ConfigT, nu: int, override_task_name: str | None = None) -> None:
        self.config = config
        self.nu = nu
        self.engine: "Engine | None" = None
        if override_task_name is not None:
            self.config.set_override(override_task_name)

    def bind(self, engine: "Engine") -> "Optimizer":
        """Attach the GPU engine that executes update_nominal_knots."""
        self.engine = engine
        return self

    def _engine(self) -> "Engine":
        if self.engine is None:
            raise RuntimeError(f"{type(self).__name__}.update_nominal_knots runs on the GPU: call .bind(engine) first "
                               "(judo_b200 has no CPU fallback)")
        return self.engine

    num_rollouts = property(lambda self: self.config.num_rollouts)
    num_nodes = property(lambda self: self.config.num_nodes)
    use_noise_ramp = property(lambda self: self.config.use_noise_ramp)
    noise_ramp = property(lambda self: self.config.noise_ramp)

    def pre_optimization(self, old_times: np.ndarray, new_times: np.ndarray) -> None:
        """Hook before the optimisation loop (judo/optimizers/base.py:57-66)."""

    def stop_cond(self) -> bool:
        """Extra stopping condition (judo/optimizers/base.py:68-74): never, by default."""
        return False

    def _ramp(self) -> np.ndarray:
        key = (self.num_nodes, self.noise_ramp)
        if getattr(self, "_ramp_key", None) != key:  # rebuilt only when the config changes (read-only: callers multiply, never mutate)
            K = self.num_nodes
            self._ramp_key, self._ramp_val = key, self.noise_ramp * np.linspace(1 / K, 1, K, endpoint=True)[:, None]
            self._ramp_val.flags.writeable = False
        return self._ramp_val

    def _noised(self, nominal_knots: np.ndarray, sigma: np.ndarray | float) -> np.ndarray:
        """Row 0 is the un-noised nominal; rows 1.. get sigma * N(0, 1) (mppi.py:58-59, cem.py:73-74, ps.py:49-50)."""
        from judo_b200.engine import legacy_stream

        stream, shape = legacy_stream(), (self.num_rollouts - 1, self.num_nodes, self.nu)
        # numpy's global legacy stream either way (seed parity with the reference); the library's batched sampler is ~3x faster
        noise = stream.randn(shape[0] * shape[1] * shape[2]).reshape(shape) if stream.ok else np.random.randn(*shape)
        out = np.empty((self.num_rollouts, self.num_nodes, self.nu))
        out[0] = nominal_knots
        np.multiply(noise, sigma, out=out[1:])  # same two roundings as nominal + sigma * noise, without the temporaries
        out[1:] += nominal_knots
        return out

    @abstractmethod
    def sample_control_knots(self, nominal_knots: np.ndarray) -> np.ndarray:
        """(num_nodes, nu) -> (num_rollouts, num_nodes, nu)."""

    @abstractmethod
    def update_nominal_knots(self, sampled_knots: np.ndarray, rewards: np.ndarray) -> np.ndarray:
        """(num_rollouts, num_nodes, nu), (num_rollouts,) -> (num_nodes, nu)."""

    def device_sigma(self) -> np.ndarray:
        """(num_nodes, nu) standard deviation of the next draw, for on-device sampling (Engine.plan_step_sampled).  Applies the
        same state changes sample_control_knots would (CEM's ramp mutates sigma)."""
        raise NotImplementedError(f"{type(self).__name__} does not support on-device sampling")

    def fused_params(self) -> np.ndarray:
        """opt_params vector for b200mpc_plan_step."""
        return np.zeros(0)

    def accept_fused(self, result: dict) -> np.ndarray:
        """Consume the output of a fused plan step (state updates such as CEM's sigma) and return the nominal knots."""
        return result["nominal"]
