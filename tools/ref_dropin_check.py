"""Drop-in proof on the reference's OWN Controller.

Builds the UNMODIFIED judo.controller.Controller from /root/reference (stand-ins only for the absent viser / mujoco / omegaconf modules,
tools/ref_shim.py; the Task's MjModel is a namespace filled from the compiled MJCF table, as in tools/gen_golden.py), installs the B200
backend by attribute assignment — the pattern of the reference's own tests (tests/test_controller/test_controller.py:55,92) and of
INTEGRATION.md §1 — and replays the three golden plan steps.  Everything but the rollouts is then the reference's code: sampling, clip,
scipy spline, Task.reward, update_nominal_knots, update_traces.

    python tools/ref_dropin_check.py [--engine sim|gpu] [tag ...]

--engine gpu uses libb200mpc.so on cuda:0; --engine sim runs the same device code on the CPU SIMT emulator (tests/warpsim), which is what
the no-GPU test tier can do in the authoring container (the GPU box has no /root/reference, this container has no GPU).
Exit code 0 = every comparison passed.
"""
import os
import sys
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import gen_golden as gg  # noqa: E402  (imports the reference through ref_shim)

from judo_b200.rollout_backend import B200RolloutBackend  # noqa: E402


def make_reference_controller(task_name: str, opt_name: str, N: int, horizon: float, seed: int):
    table = gg.load_table(task_name)
    np.random.seed(seed)
    if task_name == "cartpole":
        task = gg.fake_task(gg.Cartpole, gg.CartpoleConfig, table)
    elif task_name == "cylinder_push":
        task = gg.fake_task(gg.CylinderPush, gg.CylinderPushConfig, table)
    elif task_name == "fr3_pick":
        task, _ = gg._fr3_fake(table)
    else:
        task = gg.fake_task(gg.LeapCube, gg.LeapCubeConfig, table, dict(goal_pos=np.array([0.0, 0.03, 0.1]), goal_quat=np.array([1.0, 0, 0, 0]),
                                                                        qpos_home=gg.QPOS_HOME, reset_command=gg.QPOS_HOME[7:].copy()))
    cls, cfg_cls = gg.get_registered_optimizers()[opt_name]
    ocfg = cfg_cls()
    ocfg.set_override(task_name)
    ocfg.num_rollouts = N
    opt = cls(ocfg, table["nu"])
    ccfg = gg.ControllerConfig()
    ccfg.set_override(task_name)
    ccfg.horizon = horizon
    trace_ids = [i for i, s in enumerate(table["sensors"]) if s["type"] in ("framepos", "framepos_body") and "trace" in s["name"]]

    class _Unused(gg.RolloutBackend):   # the ctor wants an MJRolloutBackend(model, num_threads); replaced right below
        def __init__(self, model, num_threads): self.num_threads = num_threads  # noqa: ANN001, E704
        def rollout(self, *a, **k): raise AssertionError("the B200 backend was not installed")  # noqa: ANN002, ANN003, E704
        def update(self, n): self.num_threads = n  # noqa: ANN001, E704

    with mock.patch.object(gg.ref_ctrl, "MJRolloutBackend", _Unused), mock.patch.object(gg.ref_ctrl, "get_trace_sensors", lambda model: trace_ids):
        ctrl = gg.Controller(ccfg, task, opt)
    return ctrl, task, opt


def check(tag: str, engine: str) -> None:
    g = np.load(os.path.join(ROOT, "tests", "golden", f"plan_{tag}.npz"))
    task_name, opt_name, N, horizon, seed, _, _ = g["meta"]
    task_name, opt_name, N, horizon, seed = str(task_name), str(opt_name), int(N), float(horizon), int(seed)
    ctrl, task, opt = make_reference_controller(task_name, opt_name, N, horizon, seed)
    assert type(ctrl).__module__ == "judo.controller.controller", type(ctrl).__module__
    if engine == "sim":
        from tests.sim_engine import SimEngine

        backend = B200RolloutBackend(SimEngine(task_name, N), N)
    else:
        backend = B200RolloutBackend(task_name, N)
    assert isinstance(backend, gg.RolloutBackend) or all(hasattr(backend, a) for a in ("rollout", "update", "num_threads"))
    ctrl.rollout_backend = backend                       # <- the whole integration
    if task_name == "leap_cube":
        ctrl.system_metadata = {"goal_quat": g["goal_quat"]}
    np.testing.assert_array_equal(np.concatenate([task.data.qpos, task.data.qvel]), g["x_init"])
    tol = {"fr3_pick": 1e-6, "leap_cube": 2e-5}.get(task_name, 1e-8)
    launches0 = backend.engine.launch_count
    for p in range(3):
        ctrl.current_state = g[f"p{p}_x0"].copy()
        ctrl.time = float(g[f"p{p}_time"])
        ctrl.update_action()
        np.testing.assert_array_equal(ctrl.candidate_knots, g[f"p{p}_candidate_knots"]) if p == 0 else \
            np.testing.assert_allclose(ctrl.candidate_knots, g[f"p{p}_candidate_knots"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(ctrl.states[..., : ctrl.model.nq], g[f"p{p}_states"][..., : ctrl.model.nq], rtol=0, atol=tol)
        np.testing.assert_allclose(ctrl.sensors, g[f"p{p}_sensors"], rtol=0, atol=tol)
        np.testing.assert_allclose(ctrl.rewards, g[f"p{p}_rewards"], rtol=tol, atol=tol)
        np.testing.assert_allclose(ctrl.nominal_knots, g[f"p{p}_nominal_out"], rtol=0, atol=max(tol, 1e-4 if task_name == "leap_cube" else 0))
        np.testing.assert_allclose(ctrl.traces, g[f"p{p}_traces"], rtol=0, atol=tol)
        assert ctrl.states.dtype == np.float64 and ctrl.states.flags.c_contiguous and ctrl.states.shape == (N, ctrl.num_timesteps, ctrl.model.nq + ctrl.model.nv)
    assert backend.engine.launch_count - launches0 == 3, "one backend launch per reference plan step"
    # Controller.update_action resizes the backend when num_rollouts changes (controller.py:225-226)
    ctrl.optimizer_cfg.num_rollouts = N + 3
    ctrl.update_action()
    assert backend.num_threads == N + 3 and ctrl.states.shape[0] == N + 3
    print(f"ok {tag}: unmodified reference Controller + B200RolloutBackend[{engine}] reproduces 3 golden plan steps (tol {tol:g})")


if __name__ == "__main__":
    args = sys.argv[1:]
    engine = "sim"
    if "--engine" in args:
        i = args.index("--engine")
        engine = args[i + 1]
        del args[i:i + 2]
    for tag in args or ["cartpole_ps", "cartpole_mppi", "cylinder_push_cem"]:
        check(tag, engine)
