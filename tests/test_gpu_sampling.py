"""GPU: on-device (Philox) candidate sampling — the perf-mode replacement of S1-S3 + S4."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import plan as op  # noqa: E402
from oracle.mjc import OracleModel  # noqa: E402


def _setup(task, N, K, H, order="zero"):
    from judo_b200.engine import Engine
    from judo_b200.spline import spline_basis
    from judo_b200.tasks import get_registered_tasks

    cls, cfg = get_registered_tasks()[task]
    t = cls.__new__(cls)
    t.config = cfg()
    eng = Engine(task, N)
    dt = {"cartpole": 0.04, "cylinder_push": 0.02, "leap_cube": 0.01, "fr3_pick": 0.004}[task]
    basis = spline_basis(np.linspace(0, H * dt, K), dt * np.arange(H), order)
    return eng, t, basis


def test_distribution_rows_clip_and_determinism():
    N, K, H = 8192, 4, 8
    eng, t, basis = _setup("cylinder_push", N, K, H)
    rng = np.random.default_rng(0)
    x0 = np.array([1.0, 0.2, 2.0, -0.5, 0, 0, 0, 0.0])
    nominal = rng.normal(size=(K, 2))
    sigma = np.array([[0.3, 0.6], [0.5, 1.0], [0.7, 1.4], [0.9, 1.8]])
    lo, hi = np.array([-10.0, -1.5]), np.array([10.0, np.inf])
    a = eng.plan_step_sampled(x0, nominal, sigma, lo, hi, N, basis, t.cost_params(), "mppi", [0.05], seed=7, counter=3, want_knots=True)
    kn = a["knots"]
    np.testing.assert_array_equal(kn[0], np.clip(nominal, lo, hi))                    # row 0: the un-noised nominal (clipped)
    assert kn[..., 1].min() >= -1.5 and kn[..., 0].max() <= 10.0                      # clip to the actuator range
    free = kn[1:, :, 0]                                                               # dimension 0 is never clipped here
    se = sigma[:, 0] / np.sqrt(N - 1)
    assert np.all(np.abs(free.mean(0) - nominal[:, 0]) < 5 * se)
    assert np.all(np.abs(free.std(0) / sigma[:, 0] - 1) < 0.05)
    z = (free - nominal[:, 0]) / sigma[:, 0]
    assert abs(np.mean(z**3)) < 0.1 and abs(np.mean(z**4) - 3) < 0.25                 # Gaussian skewness / kurtosis
    assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.05 and abs(np.corrcoef(z[:-1, 0], z[1:, 0])[0, 1]) < 0.05
    b = eng.plan_step_sampled(x0, nominal, sigma, lo, hi, N, basis, t.cost_params(), "mppi", [0.05], seed=7, counter=3, want_knots=True)
    np.testing.assert_array_equal(a["knots"], b["knots"])                            # same (seed, counter) -> same bits
    np.testing.assert_array_equal(a["nominal"], b["nominal"])
    c = eng.plan_step_sampled(x0, nominal, sigma, lo, hi, N, basis, t.cost_params(), "mppi", [0.05], seed=7, counter=4, want_knots=True)
    assert not np.array_equal(a["knots"][1:], c["knots"][1:])                        # next plan step: fresh noise
    d = eng.plan_step_sampled(x0, nominal, sigma, lo, hi, N, basis, t.cost_params(), "mppi", [0.05], seed=8, counter=3, want_knots=True)
    assert not np.array_equal(a["knots"][1:], d["knots"][1:])
    eng.close()


@pytest.mark.parametrize("task,nu,optimizer,params", [("cartpole", 1, "mppi", [0.05]), ("cylinder_push", 2, "cem", [3, 0.1, 1.0]),
                                                      ("cartpole", 1, "ps", []), ("leap_cube", 16, "mppi", [0.0025]),
                                                      ("fr3_pick", 8, "cem", [3, 0.01, 0.3])])
def test_sampled_plan_step_is_consistent_with_host_path(task, nu, optimizer, params):
    """The generated candidates, pushed through the seed-parity path (host knots) and through the oracle, give the same
    rewards / nominal / elite list; sharding the launch (index_offset) reproduces the same candidates."""
    N, K, H = 96, 4, 12
    eng, t, basis = _setup(task, N, K, H, {"leap_cube": "cubic", "fr3_pick": "linear"}.get(task, "zero"))
    rng = np.random.default_rng(1)
    sensors_needed = False
    if task == "fr3_pick":
        from judo_b200.tasks.fr3_pick import Phase
        from tests.fr3_cases import Q_GRASP, oracle_model, scenario

        om = oracle_model()
        x0, _ = scenario("grasp", 1, 1)
        nominal = np.tile(np.concatenate([Q_GRASP, [0.0]]), (K, 1))   # closing the gripper on the cube
        lo = np.array([a["ctrlrange"][0] for a in om.table["actuators"]])
        hi = np.array([a["ctrlrange"][1] for a in om.table["actuators"]])
        t.phase, t.arm_pos_slice = Phase.MOVE, slice(7, 16)
        cp = t.cost_params()
        sigma = 0.05 * np.linspace(0.25, 1, K)[:, None] * np.ones((K, nu))
        sensors_needed = True
    elif task == "leap_cube":
        from judo_b200.tasks.leap_cube import QPOS_HOME, reduced_collision_model
        from oracle.mjc import load_table

        tb = load_table(task)
        geoms, pairs = reduced_collision_model(tb)
        om = OracleModel(tb, pairs=pairs, geoms=geoms)
        t.goal_pos = np.array([0.0, 0.03, 0.1])
        x0 = np.concatenate([QPOS_HOME, np.zeros(22)])
        nominal = np.tile(QPOS_HOME[7:], (K, 1))
        lo = np.array([a["ctrlrange"][0] for a in tb["actuators"]])
        hi = np.array([a["ctrlrange"][1] for a in tb["actuators"]])
        cp = t.cost_params({})
        sigma = 0.2 * 4.0 * np.linspace(0.25, 1, K)[:, None] * np.ones((K, nu))
    else:
        om = OracleModel(task)
        x0 = rng.normal(size=om.nq + om.nv) * 0.5 + (np.array([0, 0, 2, 2, 0, 0, 0, 0.0]) if task == "cylinder_push" else 0)
        nominal = 0.3 * rng.normal(size=(K, nu))
        lo, hi = (np.full(nu, -1.8), np.full(nu, 1.8)) if task == "cartpole" else (np.full(nu, -10.0), np.full(nu, 10.0))
        cp = t.cost_params()
        sigma = 0.25 * np.linspace(0.5, 2.0, K)[:, None] * np.ones((K, nu))
    ne = 4
    r = eng.plan_step_sampled(x0, nominal, sigma, lo, hi, N, basis, cp, optimizer, params, seed=11, counter=0, n_elite=ne, want_knots=True)
    kn = r["knots"]
    # (a) the same candidates through the host-knots path
    h = eng.plan_step(x0, kn, basis, cp, optimizer, np.array(params), want_rewards=True, n_elite=ne)
    np.testing.assert_array_equal(r["rewards"], h["rewards"])
    np.testing.assert_array_equal(r["nominal"], h["nominal"])
    np.testing.assert_array_equal(r["elite"], h["elite"])
    np.testing.assert_array_equal(r["elite_knots"], kn[r["elite"]])
    if optimizer == "cem":
        np.testing.assert_array_equal(r["sigma"], h["sigma"])
    # (b) against the oracle
    ctrl = np.einsum("hk,nkj->nhj", basis, kn)
    states, sens = om.rollout(x0, ctrl)
    assert sensors_needed == (task == "fr3_pick")
    ref = {"cartpole": lambda: op.cartpole_reward(states, ctrl), "cylinder_push": lambda: op.cylinder_push_reward(states, ctrl),
           "leap_cube": lambda: op.leap_cube_reward(states), "fr3_pick": lambda: op.fr3_pick_reward(states, sens, 1)}[task]()
    np.testing.assert_allclose(r["rewards"], ref, rtol=1e-7, atol=1e-7)
    # (c) shard invariance: two launches of 48 with index offsets 0 / 48 generate the same 96 candidates
    eng.update(48)
    a = eng.plan_step_sampled(x0, nominal, sigma, lo, hi, 48, basis, cp, optimizer, params, seed=11, counter=0, index_offset=0, want_knots=True)
    b = eng.plan_step_sampled(x0, nominal, sigma, lo, hi, 48, basis, cp, optimizer, params, seed=11, counter=0, index_offset=48, want_knots=True)
    np.testing.assert_array_equal(np.concatenate([a["knots"], b["knots"]]), kn)
    np.testing.assert_array_equal(np.concatenate([a["rewards"], b["rewards"]]), r["rewards"])
    eng.close()


def test_controller_device_sampling_mode(temp_np_seed):
    from judo_b200.controller import make_controller

    with temp_np_seed(2):
        ctrl = make_controller("cartpole", "mppi")
        ctrl.optimizer_cfg.num_rollouts = 512
        ctrl.sampling = "device"
        ctrl.device_seed = 5
        state = np.random.get_state()[1].copy()
        for _ in range(3):
            ctrl.update_action()
        assert np.array_equal(np.random.get_state()[1], state)          # no host RNG draws in device mode
        assert ctrl.rewards.shape == (512,) and np.all(np.isfinite(ctrl.nominal_knots)) and ctrl.traces.shape[1:] == (2, 3)
        assert ctrl.candidate_knots.shape == (5, 4, 1)                    # only the elite candidates came back
        assert np.all(np.abs(ctrl.nominal_knots) <= 1.8)
    # CEM keeps its sigma state in device mode
    with temp_np_seed(2):
        ctrl = make_controller("cylinder_push", "cem")
        ctrl.sampling = "device"
        s0 = ctrl.optimizer.sigma.copy()
        ctrl.update_action()
        assert ctrl.optimizer.sigma.shape == s0.shape and np.all(ctrl.optimizer.sigma >= 0.1) and np.all(ctrl.optimizer.sigma <= 1.0)
