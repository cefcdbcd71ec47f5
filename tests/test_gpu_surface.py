"""GPU: plugin-surface behaviour and edge cases — the analogue of the reference's tests/test_controller/*.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import plan as op  # noqa: E402
from oracle.mjc import OracleModel  # noqa: E402


def test_max_opt_iters_chains_iterations(temp_np_seed):
    """Reference test_max_opt_iters (tests/test_controller/test_controller.py:41-77): with the same seed, the knots the
    optimizer receives in the 2nd iteration of a max_opt_iters=2 run equal the output of a max_opt_iters=1 run."""
    from judo_b200.controller import ControllerConfig, make_controller
    from judo_b200.optimizers import Optimizer, OptimizerConfig

    class Tracker(Optimizer):
        name = "tracker"

        def __init__(self, cfg, nu):
            super().__init__(cfg, nu)
            self.history = []

        def sample_control_knots(self, nominal_knots):
            self.history.append(nominal_knots.copy())
            return nominal_knots + np.random.randn(self.num_rollouts, self.config.num_nodes, self.nu)

        def update_nominal_knots(self, sampled_knots, rewards):
            return sampled_knots[0]

    def run(max_iters):
        with temp_np_seed(42):
            ctrl = make_controller("cylinder_push", "cem")
            ctrl.controller_cfg = ControllerConfig(max_opt_iters=max_iters)
            ctrl.optimizer = Tracker(OptimizerConfig(), ctrl.task.nu)
            ctrl.current_state = np.random.rand(8)
            ctrl.time = 0.0
            ctrl.update_action()
            return ctrl

    c1, c2 = run(1), run(2)
    np.testing.assert_array_equal(c1.optimizer.history[0], c2.optimizer.history[0])
    assert not np.array_equal(c2.optimizer.history[-1], c2.optimizer.history[0])
    np.testing.assert_array_equal(c2.optimizer.history[-1], c1.nominal_knots)


@pytest.mark.parametrize("opt", ["cem", "mppi", "ps"])
@pytest.mark.parametrize("task", ["cylinder_push", "cartpole", "leap_cube", "fr3_pick"])
def test_update_action_shapes(task, opt):
    """Reference test_update_action (:80-112): shapes after one plan step, every optimizer."""
    from judo_b200.controller import make_controller

    np.random.seed(1)
    ctrl = make_controller(task, opt)
    if task == "leap_cube":
        ctrl.controller_cfg.horizon = 0.1
    ctrl.update_action()
    N, K, nu = ctrl.optimizer_cfg.num_rollouts, ctrl.optimizer_cfg.num_nodes, ctrl.task.nu
    assert ctrl.nominal_knots.shape == (K, nu) and ctrl.candidate_knots.shape == (N, K, nu) and ctrl.rewards.shape == (N,)
    ne = min(ctrl.max_num_traces, N)
    assert ctrl.traces.shape == (ne * ctrl.num_trace_sensors * (ctrl.num_timesteps - 1), 2, 3)
    assert np.all(np.isfinite(ctrl.nominal_knots)) and ctrl.action(ctrl.time).shape == (nu,)
    lo, hi = ctrl.task.actuator_ctrlrange[:, 0], ctrl.task.actuator_ctrlrange[:, 1]
    assert np.all(ctrl.candidate_knots >= lo - 1e-12) and np.all(ctrl.candidate_knots <= hi + 1e-12)


@pytest.mark.parametrize("optimizer", ["mppi", "cem", "ps"])
@pytest.mark.parametrize("normalizer", ["min_max", "running"])
def test_action_normalizers(normalizer, optimizer, temp_np_seed):
    """Non-identity normalizers: candidates stay inside ctrlrange, the running statistics equal numpy's over the candidates
    (reference test_action_normalization.py), and the fused step (update over denormalised candidates, affine map applied to
    the result) equals the contract-A step that updates in normalised space like controller.py:253-288."""
    from judo_b200.controller import make_controller

    out = {}
    for fused in (True, False):
        with temp_np_seed(3):
            ctrl = make_controller("cylinder_push", optimizer)
            ctrl.fused = fused
            ctrl.controller_cfg.action_normalizer = normalizer
            ctrl.action_normalizer = ctrl._init_action_normalizer()
            for _ in range(3):
                ctrl.update_action()
            assert np.all(np.abs(ctrl.candidate_knots) <= 10.0 + 1e-12)
            if normalizer == "running":
                assert ctrl.action_normalizer.count > 0 and np.all(np.isfinite(ctrl.action_normalizer.std))
            out[fused] = (ctrl.nominal_knots.copy(), np.array(getattr(ctrl.optimizer, "sigma", 0.0), dtype=float).copy())
    np.testing.assert_allclose(out[True][0], out[False][0], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(out[True][1], out[False][1], rtol=1e-9, atol=1e-11)


def test_running_normalizer_statistics(temp_np_seed):
    from judo_b200.controller import make_controller

    with temp_np_seed(3):
        ctrl = make_controller("cartpole", "mppi")
        ctrl.controller_cfg.action_normalizer = "running"
        ctrl.action_normalizer = ctrl._init_action_normalizer()
        ctrl.update_action()
        flat = ctrl.candidate_knots.reshape(-1, 1)
        np.testing.assert_allclose(ctrl.action_normalizer.mean, flat.mean(0), rtol=1e-12)
        np.testing.assert_allclose(ctrl.action_normalizer.std, np.clip(flat.std(0), 1e-5, 1e3), rtol=1e-9)


def test_user_task_with_numpy_reward_uses_gpu_rollouts(temp_np_seed):
    """A user-defined Task overriding reward() in NumPy keeps working (plugin surface preserved): GPU rollouts, Python reward."""
    from judo_b200.controller import ControllerConfig, Controller
    from judo_b200.optimizers import MPPI, MPPIConfig
    from judo_b200.tasks import Cartpole

    calls = []

    class MyCartpole(Cartpole):
        def cost_params(self, system_metadata=None):
            return None  # no fused kernel for this task

        def reward(self, states, sensors, controls, system_metadata=None):
            calls.append(states.shape)
            return -np.abs(states[..., 0]).sum(-1)

    with temp_np_seed(0):
        task = MyCartpole()
        cfg = MPPIConfig()
        cfg.set_override("cartpole")
        ctrl = Controller(ControllerConfig(horizon=1.0, spline_order="zero"), task, MPPI(cfg, 1))
        ctrl.update_action()
    assert calls == [(32, 25, 4)]
    om = OracleModel("cartpole")
    np.testing.assert_allclose(ctrl.states, om.rollout(ctrl.current_state, ctrl.rollout_controls)[0], atol=1e-9)
    np.testing.assert_allclose(ctrl.nominal_knots, op.mppi_update(ctrl.candidate_knots, ctrl.rewards, 0.05), atol=1e-12)


def test_num_rollouts_change_resizes_backend():
    """controller.py:225-228: the backend is updated when num_rollouts changes between two plan steps."""
    from judo_b200.controller import make_controller

    np.random.seed(5)
    ctrl = make_controller("cartpole", "ps")
    ctrl.update_action()
    ctrl.optimizer_cfg.num_rollouts = 77
    ctrl.optimizer_cfg.num_nodes = 6
    ctrl.controller_cfg.horizon = 1.5
    ctrl.update_action()
    assert ctrl.rollout_backend.num_threads == 77 and ctrl.rewards.shape == (77,) and ctrl.nominal_knots.shape == (6, 1)


@pytest.mark.parametrize("task,nu", [("cartpole", 1), ("cylinder_push", 2), ("leap_cube", 16)])
@pytest.mark.parametrize("N,H,K", [(1, 1, 4), (2, 3, 5), (33, 7, 12), (95, 2, 3)])
def test_ragged_sizes_match_oracle(task, nu, N, H, K):
    """Smallest and odd sizes: one rollout, one step, K up to the reference's slider maximum (12), N not a multiple of 32."""
    from judo_b200.engine import Engine
    from judo_b200.spline import spline_basis
    from judo_b200.tasks import get_registered_tasks

    rng = np.random.default_rng(N * 100 + H)
    eng = Engine(task, N)
    cls, cfg = get_registered_tasks()[task]
    t = cls.__new__(cls)
    t.config = cfg()
    if task == "leap_cube":
        from judo_b200.tasks.leap_cube import QPOS_HOME, reduced_collision_model
        from oracle.mjc import load_table

        tb = load_table(task)
        geoms, pairs = reduced_collision_model(tb)
        om = OracleModel(tb, pairs=pairs, geoms=geoms)
        t.goal_pos = np.array([0.0, 0.03, 0.1])
        x0 = np.concatenate([QPOS_HOME, 0.1 * rng.normal(size=22)])
        knots = QPOS_HOME[7:] + 0.3 * rng.normal(size=(N, K, nu))
        params = t.cost_params({})
    else:
        om = OracleModel(task)
        x0 = rng.normal(size=om.nq + om.nv) * (1.0 if task == "cartpole" else 0.5)
        if task == "cylinder_push":
            x0[2:4] += 2.0
        knots = rng.normal(size=(N, K, nu))
        params = t.cost_params()
    dt = om.table["opt"]["timestep"]
    basis = spline_basis(np.linspace(0, max(H * dt, 4 * dt), K), dt * np.arange(H), "linear")
    reward, cost = eng.plan_costs(x0, knots, basis, params, want_cost_matrix=True)
    ctrl = np.einsum("hk,nkj->nhj", basis, knots)
    states, sensors = om.rollout(x0, ctrl)
    ref = {"cartpole": lambda: op.cartpole_reward(states, ctrl), "cylinder_push": lambda: op.cylinder_push_reward(states, ctrl),
           "leap_cube": lambda: op.leap_cube_reward(states)}[task]()
    np.testing.assert_allclose(reward, ref, rtol=1e-8, atol=1e-8)
    assert cost.shape == (N, H)
    s2, e2 = eng.rollout(x0, ctrl)
    np.testing.assert_allclose(s2, states, rtol=0, atol=1e-8)
    np.testing.assert_allclose(e2, sensors, rtol=0, atol=1e-8)
    for optimizer, pr in (("mppi", [0.05]), ("cem", [2, 0.1, 1.0]), ("ps", [])):
        res = eng.plan_step(x0, knots, basis, params, optimizer, np.array(pr), want_rewards=True, n_elite=min(3, N))
        np.testing.assert_array_equal(res["rewards"], reward)
        refn = {"mppi": lambda: op.mppi_update(knots, reward, 0.05), "cem": lambda: op.cem_update(knots, reward, 2, 0.1, 1.0)[0],
                "ps": lambda: op.ps_update(knots, reward)}[optimizer]()
        np.testing.assert_allclose(res["nominal"], refn, rtol=1e-10, atol=1e-12)
        np.testing.assert_array_equal(res["elite"], np.argsort(reward, kind="stable")[-min(3, N):][::-1])
    eng.close()


def test_engine_argument_errors():
    from judo_b200.engine import Engine

    with pytest.raises(RuntimeError):
        Engine("cartpole", 0)
    eng = Engine("cartpole", 4)
    with pytest.raises(RuntimeError, match="temperature"):
        eng.plan_step(np.zeros(4), np.zeros((4, 4, 1)), np.eye(4), np.ones(6), "mppi", np.array([0.0]))
    with pytest.raises(RuntimeError):
        eng.update(0)
    with pytest.raises(RuntimeError):  # K > 12
        eng.plan_costs(np.zeros(4), np.zeros((4, 13, 1)), np.zeros((3, 13)), np.ones(6))
    eng.close()


def test_closed_loop_sim_plan_act_matches_cpu_loop(temp_np_seed):
    """sim -> plan -> act without MuJoCo: B200Simulation plant + Controller, 15 control steps of cartpole+mppi,
    against the same loop driven by the CPU oracle (oracle plant, oracle rollouts, NumPy reward/update) on the same seed."""
    from judo_b200.controller import make_controller
    from judo_b200.simulation import B200Simulation
    from judo_b200.spline import spline_basis

    n_ctrl, substeps = 15, 2
    with temp_np_seed(11):
        sim = B200Simulation("cartpole")
        ctrl = make_controller("cartpole", "mppi")
        ctrl.task.data.qpos, ctrl.task.data.qvel = sim.task.data.qpos.copy(), sim.task.data.qvel.copy()
        x_init = np.concatenate([sim.task.data.qpos, sim.task.data.qvel])
        rng_state = np.random.get_state()
        gpu_traj = []
        for _ in range(n_ctrl):
            ctrl.update_states(sim.sim_state)
            ctrl.update_action()
            for _ in range(substeps):
                sim.step(ctrl.action(sim.task.data.time))
            gpu_traj.append(np.concatenate([sim.task.data.qpos, sim.task.data.qvel]))
        # CPU replica of the loop
        np.random.set_state(rng_state)
        om = OracleModel("cartpole")
        x, t, dt = x_init.copy(), 0.0, 0.04
        K, N, H = 4, 32, 25
        times = np.linspace(0, 1.0, K)
        nominal = np.zeros((K, 1))
        cpu_traj = []
        for _ in range(n_ctrl):
            new_times = t + np.linspace(0, 1.0, K)
            nominal = op.make_spline(times, nominal, "zero")(new_times)
            cand = np.clip(op.sample_fixed_sigma(nominal, N, 0.1, True, 2.5), -1.8, 1.8)
            u = op.make_spline(new_times, cand, "zero")(t + dt * np.arange(H))
            r = op.cartpole_reward(om.rollout(x, u)[0], u)
            nominal, times = op.mppi_update(cand, r, 0.05), new_times
            for _ in range(substeps):
                a = op.make_spline(times, nominal, "zero")(np.array([t]))
                x = om.rollout(x, a.reshape(1, 1, 1))[0][0, 0]
                t += dt
            cpu_traj.append(x.copy())
    np.testing.assert_allclose(np.array(gpu_traj), np.array(cpu_traj), rtol=0, atol=1e-7)


@pytest.mark.gpu
def test_many_elites_and_traces_are_not_capped(temp_np_seed):
    """the reference accepts any num_elites / max_num_traces; beyond the 8 the fused epilogue keeps in registers the update
    runs as separate reduction kernels instead of raising."""
    from judo_b200.controller import make_controller
    from judo_b200.engine import Engine
    from oracle import plan as op

    rng = np.random.default_rng(0)
    eng = Engine("cartpole", 300)
    knots = rng.normal(size=(300, 4, 1))
    basis = np.eye(4)[np.repeat(np.arange(4), 5)]
    x0, params = np.array([0.1, 3.0, 0.0, 0.0]), np.array([10.0, 10.0, 0.1, 0.1, 0.01, 0.1])
    for k in (9, 20, 100):
        res = eng.plan_step(x0, knots, basis, params, "cem", np.array([k, 0.1, 1.0]), want_rewards=True, n_elite=k)
        nom, sig = op.cem_update(knots, res["rewards"], k, 0.1, 1.0)
        np.testing.assert_allclose(res["nominal"], nom, atol=1e-12)
        np.testing.assert_allclose(res["sigma"], sig, atol=1e-12)
        np.testing.assert_array_equal(res["elite"], np.argsort(res["rewards"], kind="stable")[::-1][:k])
        r2 = eng.plan_step_sampled(x0, np.zeros((4, 1)), np.full((4, 1), 0.3), np.array([-1.8]), np.array([1.8]), 300, basis, params, "cem",
                                   np.array([k, 0.1, 1.0]), seed=1, counter=0, n_elite=k, want_knots=True)
        nom2, sig2 = op.cem_update(r2["knots"], r2["rewards"], k, 0.1, 1.0)
        np.testing.assert_allclose(r2["nominal"], nom2, atol=1e-12)
        np.testing.assert_allclose(r2["elite_knots"], r2["knots"][r2["elite"]], atol=0)
    eng.close()
    with temp_np_seed(1):
        for sampling in ("host", "device"):
            ctrl = make_controller("cylinder_push", "cem")
            ctrl.controller_cfg.max_num_traces = 12
            ctrl.optimizer_cfg.num_elites = 11
            ctrl.sampling = sampling
            ctrl.update_action()
            assert ctrl.traces.shape == (12 * 2 * (ctrl.num_timesteps - 1), 2, 3) and np.isfinite(ctrl.nominal_knots).all()
            ctrl.engine.close()


@pytest.mark.gpu
def test_cem_with_more_elites_than_listed_does_not_clobber_rewards():
    """num_elites > n_elite: the fused epilogue lists max(n_elite, num_elites) rollouts, the host buffers must hold that many."""
    from judo_b200.engine import Engine
    from oracle import plan as op

    rng = np.random.default_rng(1)
    eng = Engine("cartpole", 77)
    knots = rng.normal(size=(77, 4, 1))
    basis = np.eye(4)[np.repeat(np.arange(4), 5)]
    x0, params = np.array([0.1, 3.0, 0.0, 0.0]), np.array([10.0, 10.0, 0.1, 0.1, 0.01, 0.1])
    ref = eng.plan_step(x0, knots, basis, params, "cem", np.array([3, 0.1, 1.0]), n_elite=8)
    for ne in (0, 1, 2):
        res = eng.plan_step(x0, knots, basis, params, "cem", np.array([3, 0.1, 1.0]), n_elite=ne)
        np.testing.assert_array_equal(res["rewards"], ref["rewards"])
        np.testing.assert_array_equal(res["elite"], ref["elite"][:ne])
        np.testing.assert_allclose(res["nominal"], op.cem_update(knots, ref["rewards"], 3, 0.1, 1.0)[0], atol=1e-12)
    r2 = eng.plan_step_sampled(x0, np.zeros((4, 1)), np.full((4, 1), 0.3), np.array([-1.8]), np.array([1.8]), 77, basis, params, "cem",
                               np.array([3, 0.1, 1.0]), seed=1, counter=0, n_elite=1, want_knots=True)
    np.testing.assert_allclose(r2["nominal"], op.cem_update(r2["knots"], r2["rewards"], 3, 0.1, 1.0)[0], atol=1e-12)
    eng.close()
