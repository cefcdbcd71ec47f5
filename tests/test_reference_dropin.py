"""Drop-in on the reference's OWN Controller: tools/ref_dropin_check.py builds the unmodified judo.controller.Controller
from /root/reference, assigns `controller.rollout_backend = B200RolloutBackend(...)` and replays the golden plan steps.  Run in a
subprocess because importing the reference needs stand-in modules (viser, mujoco, omegaconf) that must not leak into this session.

The no-GPU tier runs the kernels' device code on the CPU SIMT emulator; the -m gpu flavour runs libb200mpc.so but needs a box that has
BOTH a GPU and /root/reference (the gpurun box has no reference tree, so it skips there)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.isdir("/root/reference/judo")


def _run(engine: str, tags: list[str]) -> None:
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_dropin_check.py"), "--engine", engine, *tags],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    for t in tags:
        assert f"ok {t}:" in res.stdout


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference tree at /root/reference")
def test_b200_backend_in_the_unmodified_reference_controller_emulated():
    _run("sim", ["cartpole_ps", "cartpole_mppi", "cylinder_push_cem"])


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason="needs a GPU AND the reference tree at /root/reference (absent on the gpurun box)")
def test_b200_backend_in_the_unmodified_reference_controller_gpu():
    _run("gpu", ["cartpole_ps", "cartpole_mppi", "cylinder_push_cem", "leap_cube_mppi", "fr3_pick_cem"])
