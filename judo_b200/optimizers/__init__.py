"""Optimizer registry + per-task default overrides — mirror of judo/optimizers/__init__.py:28-46 and the BASELINE
tasks' entries of judo/optimizers/overrides.py:10-109."""

from __future__ import annotations

from typing import Type

from judo_b200.config import set_config_overrides
from judo_b200.optimizers.base import Optimizer, OptimizerConfig
from judo_b200.optimizers.cem import CrossEntropyMethod, CrossEntropyMethodConfig
from judo_b200.optimizers.mppi import MPPI, MPPIConfig
from judo_b200.optimizers.ps import PredictiveSampling, PredictiveSamplingConfig

_COMMON = {"num_nodes": 4, "num_rollouts": 32, "use_noise_ramp": True}
for _task in ("cylinder_push", "cartpole"):  # overrides.py:10-73
    set_config_overrides(_task, PredictiveSamplingConfig, dict(_COMMON))
    set_config_overrides(_task, CrossEntropyMethodConfig, dict(_COMMON, num_elites=2))
    set_config_overrides(_task, MPPIConfig, dict(_COMMON))
# overrides.py:76-109
set_config_overrides("leap_cube", PredictiveSamplingConfig, dict(_COMMON, noise_ramp=4.0, sigma=0.2))
set_config_overrides("leap_cube", CrossEntropyMethodConfig, dict(_COMMON, num_elites=3, noise_ramp=4.0))
set_config_overrides("leap_cube", MPPIConfig, dict(_COMMON, noise_ramp=4.0, sigma=0.2, temperature=0.0025))

# overrides.py:112-146
set_config_overrides("leap_cube_down", PredictiveSamplingConfig, dict(_COMMON, noise_ramp=4.0, sigma=0.2))
set_config_overrides("leap_cube_down", CrossEntropyMethodConfig, dict(_COMMON, num_rollouts=64, num_elites=3, noise_ramp=4.0))
set_config_overrides("leap_cube_down", MPPIConfig, dict(_COMMON, num_rollouts=64, noise_ramp=4.0, sigma=0.2, temperature=0.0025))

# overrides.py:217-254
set_config_overrides("fr3_pick", PredictiveSamplingConfig, dict(num_nodes=8, num_rollouts=64, use_noise_ramp=True, noise_ramp=4.0, sigma=0.2))
set_config_overrides("fr3_pick", CrossEntropyMethodConfig, dict(num_nodes=4, num_rollouts=64, num_elites=3, use_noise_ramp=True, noise_ramp=4.0,
                                                                sigma_min=0.01, sigma_max=0.3))
set_config_overrides("fr3_pick", MPPIConfig, dict(num_nodes=4, num_rollouts=64, use_noise_ramp=True, noise_ramp=4.0, sigma=0.01, temperature=0.002))

_registered_optimizers: dict[str, tuple[Type[Optimizer], Type[OptimizerConfig]]] = {
    "cem": (CrossEntropyMethod, CrossEntropyMethodConfig),
    "mppi": (MPPI, MPPIConfig),
    "ps": (PredictiveSampling, PredictiveSamplingConfig),
}


BUILTIN_OPTIMIZERS: tuple[Type[Optimizer], ...] = (CrossEntropyMethod, MPPI, PredictiveSampling)


def fused_optimizer_ok(opt: Optimizer, sampling: bool = False) -> bool:
    """True when the optimizer's update (and, with ``sampling``, its sampling) is the built-in implementation the fused kernel / the
    C-side sampler restates; a subclass overriding them keeps its own Python code path."""
    base = next((c for c in type(opt).__mro__ if c in BUILTIN_OPTIMIZERS), None)
    if base is None:
        return False
    names = ["update_nominal_knots", "fused_params", "accept_fused"] + (["sample_control_knots", "device_sigma"] if sampling else [])
    return all(getattr(type(opt), m) is getattr(base, m) for m in names)


def get_registered_optimizers() -> dict[str, tuple[Type[Optimizer], Type[OptimizerConfig]]]:
    return _registered_optimizers


def register_optimizer(name: str, controller_type: Type[Optimizer], controller_config_type: Type[OptimizerConfig]) -> None:
    _registered_optimizers[name] = (controller_type, controller_config_type)


__all__ = ["get_registered_optimizers", "register_optimizer", "CrossEntropyMethod", "CrossEntropyMethodConfig", "MPPI", "MPPIConfig",
           "Optimizer", "OptimizerConfig", "PredictiveSampling", "PredictiveSamplingConfig"]
