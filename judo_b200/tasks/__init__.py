"""Task registry — mirror of judo/tasks/__init__.py:25-47 for the BASELINE tasks and fr3_pick."""

from __future__ import annotations

from typing import Dict, Tuple, Type

from judo_b200.tasks.base import Task, TaskConfig
from judo_b200.tasks.cartpole import Cartpole, CartpoleConfig
from judo_b200.tasks.cylinder_push import CylinderPush, CylinderPushConfig

_registered_tasks: Dict[str, Tuple[Type[Task], Type[TaskConfig]]] = {
    CylinderPush.name: (CylinderPush, CylinderPushConfig),
    Cartpole.name: (Cartpole, CartpoleConfig),
}

try:  # the leap task registers itself once its kernel is in the library
    from judo_b200.tasks.leap_cube import LeapCube, LeapCubeConfig, LeapCubeDown, LeapCubeDownConfig

    _registered_tasks[LeapCube.name] = (LeapCube, LeapCubeConfig)
    _registered_tasks[LeapCubeDown.name] = (LeapCubeDown, LeapCubeDownConfig)
except ImportError:
    pass

from judo_b200.tasks.fr3_pick import FR3Pick, FR3PickConfig  # noqa: E402

_registered_tasks[FR3Pick.name] = (FR3Pick, FR3PickConfig)


# the classes whose reward / hooks the fused kernels implement (Controller._can_fuse)
BUILTIN_TASKS: Tuple[Type[Task], ...] = tuple(cls for cls, _ in _registered_tasks.values())


def fused_task_ok(task: Task) -> bool:
    """True when ``task`` is a built-in task whose reward and rollout hooks are the built-in implementations the fused kernel restates
    (a subclass that overrides reward / post_rollout / task_to_sim_ctrl must go through rollout() + its own Python code, as in the
    reference: judo/controller/controller.py:267-285)."""
    base = next((c for c in type(task).__mro__ if c in BUILTIN_TASKS), None)
    if base is None:
        return False
    return all(getattr(type(task), m) is getattr(base, m) for m in ("reward", "post_rollout", "task_to_sim_ctrl"))


def get_registered_tasks() -> Dict[str, Tuple[Type[Task], Type[TaskConfig]]]:
    return _registered_tasks


def register_task(name: str, task_type: Type[Task], task_config_type: Type[TaskConfig]) -> None:
    _registered_tasks[name] = (task_type, task_config_type)


__all__ = ["get_registered_tasks", "register_task", "Task", "TaskConfig", "Cartpole", "CartpoleConfig", "CylinderPush", "CylinderPushConfig"]
