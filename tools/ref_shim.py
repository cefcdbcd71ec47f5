"""Import the UNMODIFIED reference Python (/root/reference) in the authoring container.

The reference cannot be imported as-is: judo.gui needs ``viser`` and judo.tasks needs ``mujoco`` (both absent,
SURVEY.md §8c).  This shim registers empty stand-in modules for those two names ONLY so that the reference's
own pure-NumPy code — optimizers, normalizers, quaternion math, the three reward() methods — runs unmodified
and can generate golden vectors (tools/gen_golden.py).  Nothing here runs on the GPU box or in the product.
"""
import sys
import types
from unittest import mock

REFERENCE_ROOT = "/root/reference"


def install() -> None:
    if "viser" not in sys.modules:
        viser = types.ModuleType("viser")
        for name in ("GuiCheckboxHandle", "GuiDropdownHandle", "GuiEvent", "GuiFolderHandle", "GuiInputHandle",
                     "GuiSliderHandle", "MeshHandle", "ViserServer"):
            setattr(viser, name, type(name, (), {}))
        sys.modules["viser"] = viser
    if "mujoco" not in sys.modules:
        mj = mock.MagicMock(name="mujoco")
        mj.__path__ = []
        sys.modules["mujoco"] = mj
        sys.modules["mujoco.rollout"] = mock.MagicMock(name="mujoco.rollout")
    for missing in ("robot_descriptions", "robot_descriptions.loaders", "robot_descriptions.loaders.mujoco"):
        if missing not in sys.modules:
            try:
                __import__(missing)
            except Exception:  # noqa: BLE001
                sys.modules[missing] = mock.MagicMock(name=missing)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
