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


def get_registered_tasks() -> Dict[str, Tuple[Type[Task], Type[TaskConfig]]]:
    return _registered_tasks


def register_task(name: str, task_type: Type[Task], task_config_type: Type[TaskConfig]) -> None:
    _registered_tasks[name] = (task_type, task_config_type)


__all__ = ["get_registered_tasks", "register_task", "Task", "TaskConfig", "Cartpole", "CartpoleConfig", "CylinderPush", "CylinderPushConfig"]
