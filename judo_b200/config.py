"""Per-task overridable configs — mirror of judo/config.py:9-96 (OverridableConfig + override registry)."""

from __future__ import annotations

import warnings
from dataclasses import MISSING, dataclass, fields, is_dataclass
from typing import Any

import numpy as np

_REGISTRY: dict[type, dict[str, dict[str, Any]]] = {}


def _differs(a: Any, b: Any) -> bool:
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return not np.array_equal(a, b)
    return a != b


@dataclass
class OverridableConfig:
    """Dataclass whose fields can be switched to registered per-task values (judo/config.py:12-62)."""

    def __post_init__(self) -> None:
        _REGISTRY.setdefault(self.__class__, {})

    def set_override(self, key: str, reset_to_defaults: bool = True) -> None:
        active = _REGISTRY.get(self.__class__, {}).get(key, {})
        for f in fields(self):
            if active.get(f.name) is not None:
                if _differs(getattr(self, f.name, MISSING), active[f.name]):
                    setattr(self, f.name, active[f.name])
            elif reset_to_defaults:
                if f.default is not MISSING:
                    default = f.default
                elif f.default_factory is not MISSING:
                    default = f.default_factory()
                else:
                    warnings.warn(f"Field '{f.name}' has no default and no override for key '{key}'.", UserWarning, stacklevel=2)
                    continue
                if _differs(getattr(self, f.name, MISSING), default):
                    setattr(self, f.name, default)


def set_config_overrides(override_key: str, cls: type, field_override_values: dict[str, Any]) -> None:
    """Register override values for (cls, key) — judo/config.py:65-96."""
    if not is_dataclass(cls):
        raise TypeError(f"Provided class {cls.__name__} is not a dataclass.")
    slot = _REGISTRY.setdefault(cls, {}).setdefault(override_key, {})
    names = {f.name for f in fields(cls)}
    for name, value in field_override_values.items():
        if name in names:
            slot[name] = value
        else:
            warnings.warn(f"Field '{name}' not found in class '{cls.__name__}'.", UserWarning, stacklevel=2)
