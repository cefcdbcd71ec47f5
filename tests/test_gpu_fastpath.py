"""-m gpu: Controller.update_action's one-call fast path (b200mpc_controller_step) against the NumPy-glue path of the same Controller:
identical seeds must give bit-identical candidates, rewards, nominal knots, traces and leave numpy's global generator in the same state."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(task, opt, fast, seed, steps=3, N=None, K=None, horizon=None, iters=1, predraw=0, meta=None):
    from judo_b200.controller import make_controller

    np.random.seed(seed)
    ctrl = make_controller(task, opt)
    if N:
        ctrl.optimizer_cfg.num_rollouts = N
    if K:
        ctrl.optimizer_cfg.num_nodes = K
    if horizon:
        ctrl.controller_cfg.horizon = horizon
    ctrl.controller_cfg.max_opt_iters = iters
    ctrl.fast_path = fast
    np.random.seed(seed)
    ctrl.reset()
    if meta is not None:
        ctrl.system_metadata = meta(ctrl)
    np.random.randn(predraw)   # an odd predraw leaves a cached gaussian in numpy's generator
    assert ctrl._can_fast_path() == fast
    out = []
    for i in range(steps):
        ctrl.time = ctrl.task.dt * 3 * i
        ctrl.update_action()
        out.append(dict(cand=ctrl.candidate_knots.copy(), rewards=ctrl.rewards.copy(), nominal=ctrl.nominal_knots.copy(),
                        traces=None if ctrl.traces is None else ctrl.traces.copy(), elite=np.array(ctrl.elite_indices).copy(),
                        sigma=np.array(getattr(ctrl.optimizer, "sigma", 0.0), dtype=float).copy(), basis=ctrl._basis.copy()))
    after = np.random.randn(4)
    launches = ctrl.engine.launch_count
    ctrl.engine.close()
    return out, after, launches


CASES = [
    ("cartpole", "mppi", dict()),
    ("cartpole", "ps", dict(N=4, K=5)),            # (N-1)*K*nu = 15: odd block -> phase 1 / tail / phase 2
    ("cartpole", "cem", dict(N=257, predraw=1)),   # numpy's gaussian cache is occupied on entry
    ("cylinder_push", "cem", dict(N=64, iters=2)),
    ("cylinder_push", "mppi", dict(N=1, K=4)),     # a single rollout: no noise at all
    ("leap_cube", "mppi", dict(N=12, horizon=0.2, meta=lambda c: c.task.get_sim_metadata())),
    ("fr3_pick", "cem", dict(N=8, horizon=0.1)),
]


@pytest.mark.parametrize("task,opt,kw", CASES)
def test_fast_path_equals_numpy_glue(task, opt, kw):
    a, ra, la = _run(task, opt, True, 11, **kw)
    b, rb, lb = _run(task, opt, False, 11, **kw)
    assert np.array_equal(ra, rb), "numpy's generator must end in the same state"
    # the C and the NumPy cubic basis differ in the last bit (different elimination order), which 20 contact-rich steps amplify;
    # zero / linear bases are identical, and so is everything downstream of them
    exact = task != "leap_cube"
    for i, (x, y) in enumerate(zip(a, b)):
        if exact or i == 0:
            assert np.array_equal(x["cand"], y["cand"])
        if exact:
            assert np.array_equal(x["rewards"], y["rewards"])
            assert np.array_equal(x["nominal"], y["nominal"])
            assert np.array_equal(x["sigma"], y["sigma"])
        else:
            np.testing.assert_allclose(x["cand"], y["cand"], rtol=0, atol=1e-6)
            np.testing.assert_allclose(x["rewards"], y["rewards"], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(x["nominal"], y["nominal"], rtol=0, atol=1e-6)
        assert np.array_equal(x["elite"], y["elite"])
        np.testing.assert_allclose(x["basis"], y["basis"], rtol=0, atol=5e-14)
        assert (x["traces"] is None) == (y["traces"] is None)
        if x["traces"] is not None:
            np.testing.assert_allclose(x["traces"], y["traces"], rtol=0, atol=1e-12 if exact else 1e-6)
    if task in ("cartpole", "cylinder_push"):
        assert la < lb, "the fast path must not relaunch for the elite traces"


def test_fast_path_is_one_launch_per_iteration():
    from judo_b200.controller import make_controller

    ctrl = make_controller("cartpole", "mppi")
    ctrl.update_action()
    n0 = ctrl.engine.launch_count
    for _ in range(5):
        ctrl.update_action()
    assert ctrl.engine.launch_count - n0 == 5
    ctrl.engine.close()


def test_user_subclasses_leave_the_fused_kernel(temp_np_seed):
    """a Task subclass that overrides reward() (or an Optimizer subclass that overrides its update / sampling) must run its
    own Python code, as in the reference (controller.py:267-288), not the built-in fused kernel."""
    from judo_b200.controller import Controller, ControllerConfig
    from judo_b200.optimizers import MPPI, MPPIConfig
    from judo_b200.tasks.cartpole import Cartpole

    calls = {"reward": 0, "sample": 0}

    class MyCartpole(Cartpole):
        def reward(self, states, sensors, controls, system_metadata=None):
            calls["reward"] += 1
            return -np.abs(states[..., 0]).sum(axis=-1)

    class MyMPPI(MPPI):
        def sample_control_knots(self, nominal_knots):
            calls["sample"] += 1
            return super().sample_control_knots(nominal_knots)

    with temp_np_seed(3):
        cfg = ControllerConfig(); cfg.set_override("cartpole")
        ocfg = MPPIConfig(); ocfg.set_override("cartpole")
        c1 = Controller(cfg, MyCartpole(), MPPI(ocfg, 1))
        assert not c1._can_fuse() and not c1._can_fast_path()
        c1.update_action()
        assert calls["reward"] == 1
        np.testing.assert_allclose(c1.rewards, -np.abs(c1.states[..., 0]).sum(axis=-1))
        c2 = Controller(cfg, Cartpole(), MyMPPI(ocfg, 1))
        assert c2._can_fuse() and not c2._can_fast_path()   # custom sampling: NumPy glue, fused kernel for the rest
        c2.update_action()
        assert calls["sample"] == 1
        c1.engine.close(); c2.engine.close()
