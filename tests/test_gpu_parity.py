"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs, and against the
golden plan steps produced by the reference Controller.  Tolerances: 1e-4 is north_star's bar for returned costs and
nominal trajectories; the same-algorithm fp64 kernels are held to much tighter bounds here."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import plan as op  # noqa: E402
from oracle.mjc import OracleModel  # noqa: E402


@pytest.fixture(scope="module")
def engines():
    from judo_b200.engine import Engine

    cache = {}

    def get(task, N):
        if task not in cache:
            cache[task] = Engine(task, N)
        cache[task].update(N)
        return cache[task]

    yield get
    for e in cache.values():
        e.close()


def _x0(task, rng):
    if task == "cartpole":
        return np.concatenate([np.array([1.0, np.pi]) + rng.normal(size=2), 0.1 * rng.normal(size=2)])
    th = 2 * np.pi * rng.random(2)
    return np.array([np.cos(th[0]), np.sin(th[0]), 2 * np.cos(th[1]), 2 * np.sin(th[1]), 0, 0, 0, 0.0])


CASES = [("cartpole", 32, 32, 1), ("cartpole", 300, 64, 1), ("cylinder_push", 64, 50, 2), ("cylinder_push", 257, 50, 2)]


@pytest.mark.parametrize("task,N,H,nu", CASES)
def test_rollout_matches_oracle(engines, task, N, H, nu):
    """Contract A: states and sensors of N x H mj_steps."""
    rng = np.random.default_rng(hash((task, N)) % 2**32)
    eng, om = engines(task, N), OracleModel(task)
    x0 = _x0(task, rng)
    scale = 2.5 if task == "cartpole" else 4.0  # beyond ctrlrange/forcerange so both clamps are exercised
    controls = scale * rng.normal(size=(N, H, nu))
    states, sensors = eng.rollout(x0, controls)
    s_ref, e_ref = om.rollout(x0, controls)
    tol = 1e-9 if task == "cartpole" else 1e-6  # contact rows are ill-conditioned (see the contact test)
    np.testing.assert_allclose(states, s_ref, rtol=0, atol=tol)
    np.testing.assert_allclose(sensors, e_ref, rtol=0, atol=tol)
    # batched x0 (mj_rollout_backend.py:64-65 tiles a 1-D x0; a 2-D one is used as is)
    xb = np.stack([_x0(task, rng) for _ in range(N)])
    states_b, _ = eng.rollout(xb, controls, want_sensors=False)
    np.testing.assert_allclose(states_b, om.rollout(xb, controls)[0], rtol=0, atol=tol)


def test_cartpole_joint_limit_and_cylinder_contact_are_exercised(engines):
    """The constraint branches really run in the parity cases: cart driven into the +/-1.8 limit; pusher rams the cart."""
    N, H = 32, 40
    eng, om = engines("cartpole", N), OracleModel("cartpole")
    x0 = np.array([1.7, np.pi, 2.0, 0.0])
    controls = np.full((N, H, 1), 1.8) * np.linspace(0.5, 1.0, N)[:, None, None]
    s, _ = eng.rollout(x0, controls)
    assert s[..., 0].max() > 1.8  # soft limit penetrated -> constraint row active
    np.testing.assert_allclose(s, om.rollout(x0, controls)[0], rtol=0, atol=1e-9)
    eng, om = engines("cylinder_push", N), OracleModel("cylinder_push")
    x0 = np.array([-0.7, 0.05, 0, 0, 0, 0, 0, 0.0])
    controls = np.zeros((N, H, 2))
    controls[:, :, 0] = np.linspace(0.2, 3.0, N)[:, None]
    s, _ = eng.rollout(x0, controls)
    assert np.abs(s[:, -1, 2]).max() > 0.05  # the cart was pushed
    # mu = 1e-5 makes the pyramidal rows nearly parallel with D ~ 1e10 (Newton Hessian cond ~ 1e10): agreement between
    # the two fp64 implementations is ~1e-8 here, four orders inside north_star's 1e-4
    np.testing.assert_allclose(s, om.rollout(x0, controls)[0], rtol=0, atol=1e-6)


@pytest.mark.parametrize("task,N,H,nu,order", [("cartpole", 32, 32, 1, "zero"), ("cartpole", 515, 64, 1, "zero"),
                                               ("cylinder_push", 200, 50, 2, "zero"), ("cylinder_push", 64, 50, 2, "cubic"),
                                               ("cartpole", 64, 25, 1, "linear")])
def test_plan_costs_match_oracle(engines, task, N, H, nu, order):
    """Contract B: fused spline -> rollout -> per-step cost; cost matrix rows sum to -reward."""
    from judo_b200.spline import spline_basis
    from judo_b200.tasks import get_registered_tasks

    rng = np.random.default_rng(7)
    K = 4
    eng, om = engines(task, N), OracleModel(task)
    x0 = _x0(task, rng)
    dt = om.table["opt"]["timestep"]
    times = 0.3 + np.linspace(0, H * dt, K)
    query = 0.3 + dt * np.arange(H)
    knots = rng.normal(size=(N, K, nu)) * (1.0 if task == "cartpole" else 2.0)
    basis = spline_basis(times, query, order)
    tk = get_registered_tasks()[task][0].__new__(get_registered_tasks()[task][0])
    tk.config = get_registered_tasks()[task][1]()
    params = tk.cost_params()
    reward, cost = eng.plan_costs(x0, knots, basis, params, want_cost_matrix=True)
    ctrl = op.make_spline(times, knots, order)(query)
    states, _ = om.rollout(x0, ctrl)
    ref = op.cartpole_reward(states, ctrl) if task == "cartpole" else op.cylinder_push_reward(states, ctrl)
    np.testing.assert_allclose(reward, ref, rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(cost.astype(np.float64).sum(1), -reward, rtol=2e-6)
    reward2, none = eng.plan_costs(x0, knots, basis, params, want_cost_matrix=False)
    assert none is None
    np.testing.assert_array_equal(reward, reward2)
    # reward-only entry point (Task.reward for contract-A callers)
    np.testing.assert_allclose(eng.reward(states, ctrl, params), ref, rtol=1e-12)


def test_rewards_match_reference_golden(engines, golden):
    g = golden("rewards")
    from judo_b200.tasks import Cartpole, CylinderPush

    t = Cartpole()
    t.engine = engines("cartpole", 6)
    np.testing.assert_allclose(t.reward(g["cartpole_states"], None, g["cartpole_controls"]), g["cartpole_rewards"], rtol=1e-12)
    t = CylinderPush()
    t.engine = engines("cylinder_push", 6)
    np.testing.assert_allclose(t.reward(g["cylinder_push_states"], None, g["cylinder_push_controls"]), g["cylinder_push_rewards"], rtol=1e-12)
    t.config.goal_pos = g["cylinder_push_goal"]
    np.testing.assert_allclose(t.reward(g["cylinder_push_states"], None, g["cylinder_push_controls"]), g["cylinder_push_rewards_goal"], rtol=1e-12)


def test_optimizer_updates_match_reference_golden(engines, golden):
    """U1-U3 on the GPU against outputs of the reference's own update_nominal_knots."""
    g = golden("optimizers")
    for ci in range(int(g["ncases"])):
        name, task, nu, N = g[f"c{ci}_meta"][:4]
        eng = engines("cartpole", int(N))  # the update kernels are task-independent (nu comes from the knots shape)
        for it in range(2):
            knots, rewards = g[f"c{ci}_knots{it}"], g[f"c{ci}_rewards{it}"]
            K, nu_ = knots.shape[1:]
            flat = knots.reshape(len(knots), K * nu_, 1)  # engine nu=1: fold (K, nu) into K
            if name == "mppi":
                out = eng.update_mppi(flat, rewards, float(g[f"c{ci}_temperature"]))
            elif name == "ps":
                out = eng.update_ps(flat, rewards)
            else:
                out, sig = eng.update_cem(flat, rewards, int(g[f"c{ci}_num_elites"]), float(g[f"c{ci}_sigma_min"]), float(g[f"c{ci}_sigma_max"]))
                np.testing.assert_allclose(sig.reshape(K, nu_), g[f"c{ci}_sigma_out{it}"], rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(out.reshape(K, nu_), g[f"c{ci}_nominal_out{it}"], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("N,KNU", [(1, 4), (3, 4), (1000, 4), (4096, 4), (8192, 64), (777, 7)])
def test_optimizer_updates_sizes_and_edges(engines, N, KNU):
    rng = np.random.default_rng(N)
    eng = engines("cartpole", N)
    knots = rng.normal(size=(N, KNU, 1))
    rewards = -np.abs(rng.normal(size=N)) * 50
    np.testing.assert_allclose(eng.update_mppi(knots, rewards, 0.05), op.mppi_update(knots, rewards, 0.05), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(eng.update_mppi(knots, rewards, 100.0), op.mppi_update(knots, rewards, 100.0), rtol=1e-11, atol=1e-13)
    np.testing.assert_array_equal(eng.update_ps(knots, rewards), op.ps_update(knots, rewards))
    for k in (1, 2, 3, 5):
        nom, sig = eng.update_cem(knots, rewards, k, 0.1, 1.0)
        rn, rs = op.cem_update(knots, rewards, k, 0.1, 1.0)  # k > N: numpy slicing keeps all N
        np.testing.assert_allclose(nom, rn, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(sig, rs, rtol=1e-12, atol=1e-14)
    # ties: PS takes the first maximum (argmax); documented CEM tie rule = higher index first
    r2 = rewards.copy()
    if N >= 3:
        r2[0] = r2[2] = r2.max() + 1
        np.testing.assert_array_equal(eng.update_ps(knots, r2), knots[0])
        nom, _ = eng.update_cem(knots, r2, 1, 0.1, 1.0)
        np.testing.assert_array_equal(nom, knots[2])


@pytest.mark.parametrize("tag", ["cartpole_ps", "cartpole_mppi", "cylinder_push_cem", "leap_cube_mppi"])
@pytest.mark.parametrize("mode", ["fast", "fused", "contract_a"])
def test_controller_reproduces_reference_plan_steps(golden, temp_np_seed, tag, mode):
    """End to end through the plugin surface: same seed as the reference Controller run -> same candidates (bit exact in the first
    step, where the nominal is the warm start; afterwards they inherit the GPU update's rounding), rewards / nominal knots / traces
    within tolerance, for three consecutive plan steps.  mode: "fast" = update_action's default (one b200mpc_controller_step call per
    iteration: C-side sampling from numpy's stream, clip, basis, fused kernel, traces), "fused" = NumPy glue + Engine.plan_step,
    "contract_a" = the drop-in RolloutBackend path (rollout -> Task.reward -> update_nominal_knots)."""
    from judo_b200.controller import make_controller

    fused = mode != "contract_a"

    g = golden("plan_" + tag)
    task, opt, N, horizon, seed, order, max_traces = g["meta"]
    with temp_np_seed(int(seed)):
        ctrl = make_controller(str(task), str(opt))
        ctrl.optimizer_cfg.num_rollouts = int(N)
        ctrl.controller_cfg.horizon = float(horizon)
        ctrl.fused = fused
        ctrl.fast_path = mode == "fast"
        assert ctrl._can_fast_path() == (mode == "fast")
        # the golden run seeded the RNG and then built its Controller, whose reset() calls Task.reset() once
        np.random.seed(int(seed))
        ctrl.reset()
        tol = 1e-8
        if task == "leap_cube":
            ctrl.system_metadata = {"goal_quat": g["goal_quat"]}
            np.testing.assert_array_equal(ctrl.task.goal_quat, g["goal_quat"])
            tol = 2e-5  # contact-rich 40-step rollouts: see test_leap_rollout_matches_oracle
        np.testing.assert_array_equal(np.concatenate([ctrl.task.data.qpos, ctrl.task.data.qvel]), g["x_init"])
        for p in range(3):
            ctrl.current_state = g[f"p{p}_x0"].copy()
            ctrl.time = float(g[f"p{p}_time"])
            np.testing.assert_allclose(ctrl.nominal_knots, g[f"p{p}_nominal_in"], rtol=0, atol=1e-9)
            ctrl.update_action()
            np.testing.assert_allclose(ctrl.candidate_knots, g[f"p{p}_candidate_knots"], rtol=0, atol=1e-9)
            if p == 0:
                np.testing.assert_array_equal(ctrl.candidate_knots, g["p0_candidate_knots"])
            np.testing.assert_allclose(ctrl.rewards, g[f"p{p}_rewards"], rtol=tol, atol=tol)
            np.testing.assert_allclose(ctrl.nominal_knots, g[f"p{p}_nominal_out"], rtol=0, atol=max(tol, 1e-4 if task == "leap_cube" else 0))
            np.testing.assert_array_equal(ctrl.times, g[f"p{p}_times_out"])
            np.testing.assert_allclose(ctrl.traces, g[f"p{p}_traces"], rtol=0, atol=tol)
            if opt == "cem":
                np.testing.assert_allclose(ctrl.optimizer.sigma, g[f"p{p}_sigma_out"], rtol=1e-9, atol=1e-12)
            if not fused:
                np.testing.assert_allclose(ctrl.states[..., : ctrl.model.nq], g[f"p{p}_states"][..., : ctrl.model.nq], rtol=0, atol=tol)
                np.testing.assert_allclose(ctrl.sensors, g[f"p{p}_sensors"], rtol=0, atol=tol)


def test_backend_contract_and_errors(engines):
    from judo_b200.rollout_backend import B200RolloutBackend, RolloutBackend

    be = B200RolloutBackend("cartpole", 8)
    assert isinstance(be, RolloutBackend) and be.num_threads == 8
    s, e, p = be.rollout(np.zeros(4), np.zeros((8, 5, 1)))
    assert s.shape == (8, 5, 4) and e.shape == (8, 5, 6) and p is None and s.dtype == np.float64 and s.flags.c_contiguous
    with pytest.raises(ValueError):
        be.rollout(np.zeros(4), np.zeros((7, 5, 1)))  # batch != num_threads
    with pytest.raises(ValueError):
        be.rollout(np.zeros(5), np.zeros((8, 5, 1)))
    with pytest.raises(ValueError):
        be.rollout(np.zeros(4), np.zeros((8, 5, 2)))
    be.update(16)
    assert be.num_threads == 16 and be.rollout(np.zeros(4), np.zeros((16, 3, 1)))[0].shape == (16, 3, 4)
    with pytest.raises(RuntimeError):
        be.engine.plan_costs(np.zeros(4), np.zeros((4, 13, 1)), np.zeros((5, 13)), np.zeros(6))  # K > 12
    # determinism: the same call twice gives identical bits
    rng = np.random.default_rng(0)
    c = rng.normal(size=(16, 30, 1))
    a = be.rollout(np.array([0.5, 3.0, 0, 0]), c)[0]
    b = be.rollout(np.array([0.5, 3.0, 0, 0]), c)[0]
    np.testing.assert_array_equal(a, b)


def test_full_size_properties_cartpole_c2(engines):
    """BASELINE config C2 (N=4096, H=64): size-independent properties — row 0 is the nominal rollout, permuting the
    candidates permutes the rewards, MPPI weights sum to one (nominal inside the candidates' hull), chunked == whole."""
    from judo_b200.spline import spline_basis

    rng = np.random.default_rng(42)
    N, H, K = 4096, 64, 4
    eng = engines("cartpole", N)
    x0 = _x0("cartpole", rng)
    nominal = rng.normal(size=(K, 1)) * 0.3
    knots = np.concatenate([nominal[None], nominal + 0.25 * rng.normal(size=(N - 1, K, 1))])
    knots = np.clip(knots, -1.8, 1.8)
    basis = spline_basis(np.linspace(0, 2.56, K), 0.04 * np.arange(H), "zero")
    params = np.array([10, 10, 0.1, 0.1, 0.01, 0.1.__float__()])
    res = eng.plan_step(x0, knots, basis, params, "mppi", np.array([0.05]), want_rewards=True, n_elite=5)
    r = res["rewards"]
    assert np.all(np.isfinite(r)) and np.all(r <= 0)
    perm = rng.permutation(N)
    r_perm, _ = eng.plan_costs(x0, knots[perm], basis, params)
    np.testing.assert_array_equal(r_perm, r[perm])
    eng2 = engines("cartpole", 64)
    r_chunk, _ = eng2.plan_costs(x0, knots[:64], basis, params)
    np.testing.assert_array_equal(r_chunk, r[:64])
    engines("cartpole", N)
    np.testing.assert_allclose(res["nominal"], op.mppi_update(knots, r, 0.05), rtol=1e-10, atol=1e-12)
    assert np.all(res["nominal"] <= knots.max(0) + 1e-12) and np.all(res["nominal"] >= knots.min(0) - 1e-12)
    np.testing.assert_array_equal(res["elite"], np.argsort(r)[-5:][::-1])
    # spot-check 16 of the 4096 rollouts against the oracle
    om = OracleModel("cartpole")
    idx = rng.choice(N, 16, replace=False)
    ctrl = np.einsum("hk,nkj->nhj", basis, knots[idx])
    np.testing.assert_allclose(r[idx], op.cartpole_reward(om.rollout(x0, ctrl)[0], ctrl), rtol=1e-9)


@pytest.mark.parametrize("optimizer,params", [("mppi", [0.05]), ("cem", [3, 0.1, 1.0]), ("ps", [])])
def test_resident_sharded_step_equals_unsharded(optimizer, params):
    """Multi-GPU algebra on one device: two 'ranks' (two handles, disjoint halves of the candidates) leave rank partials;
    combining them equals the single fused launch over all candidates and the oracle update."""
    import ctypes

    import torch

    from judo_b200.dist import ShardedPlanner, shard_range
    from judo_b200.spline import spline_basis

    rng = np.random.default_rng(3)
    N, H, K = 600, 50, 4
    x0 = _x0("cylinder_push", rng)
    knots = rng.normal(size=(N, K, 2)) * 2
    basis = spline_basis(np.linspace(0, 1.0, K), 0.02 * np.arange(H), "zero")
    cp = np.array([0.5, 0.0, 0.1, 0.25, 0.0, 0.0])
    op_params = np.array(params, dtype=np.float64)
    whole = ShardedPlanner("cylinder_push", N)
    whole.set_problem(x0, basis, cp)
    whole.set_knots(knots)
    nominal = whole.step(optimizer, op_params, n_elite=5).cpu().numpy().reshape(K, 2)
    rewards = whole.d_reward.cpu().numpy()
    if optimizer == "mppi":
        ref = op.mppi_update(knots, rewards, 0.05)
    elif optimizer == "cem":
        ref, ref_sigma = op.cem_update(knots, rewards, 3, 0.1, 1.0)
        np.testing.assert_allclose(whole.d_sigma.cpu().numpy().reshape(K, 2), ref_sigma, rtol=1e-12)
    else:
        ref = op.ps_update(knots, rewards)
    np.testing.assert_allclose(nominal, ref, rtol=1e-11, atol=1e-13)
    np.testing.assert_array_equal(whole.d_elite.cpu().numpy()[:5].astype(int), np.argsort(rewards)[-5:][::-1])
    # two ranks
    parts = []
    planners = []
    for rank in range(2):
        lo, hi = shard_range(N, 2, rank)
        pl = ShardedPlanner("cylinder_push", hi - lo, rank=rank, world_size=2)
        pl.set_problem(x0, basis, cp)
        pl.set_knots(knots[lo:hi])
        import judo_b200.dist as D

        captured = {}
        orig = D.gather_partials
        D.gather_partials = lambda local, ws, group=None: (captured.setdefault("p", local.clone()), torch.stack([local, local]))[1]
        try:
            pl.step(optimizer, op_params, index_offset=lo)
        finally:
            D.gather_partials = orig
        parts.append(captured["p"])
        planners.append(pl)
    allp = torch.stack(parts).contiguous()
    pl = planners[0]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda x: ctypes.c_void_p(x.data_ptr())  # noqa: E731
    if optimizer == "mppi":
        pl._check(pl.lib.b200mpc_mppi_combine_dev(pl.engine.handle, P(allp), 2, K * 2, 0.05, P(pl.d_nominal), st))
    else:
        k = 3 if optimizer == "cem" else 1
        pl._check(pl.lib.b200mpc_topk_combine_dev(pl.engine.handle, P(allp), 2, K * 2, k, int(optimizer == "cem"), 0.1, 1.0, P(pl.d_nominal),
                                                  P(pl.d_sigma) if optimizer == "cem" else ctypes.c_void_p(0), P(pl.d_elite), st))
    torch.cuda.synchronize()
    np.testing.assert_allclose(pl.d_nominal.cpu().numpy().reshape(K, 2), ref, rtol=1e-11, atol=1e-13)


# ------------------------------------------------------------------------------------------------ leap_cube (reduced model)
def _leap_oracle():
    from judo_b200.tasks.leap_cube import reduced_collision_model
    from oracle.mjc import load_table

    tb = load_table("leap_cube")
    geoms, pairs = reduced_collision_model(tb)
    return OracleModel(tb, pairs=pairs, geoms=geoms), tb


def _leap_controls(tb, rng, N, H, scale):
    from judo_b200.tasks.leap_cube import QPOS_HOME

    lo = np.array([a["ctrlrange"][0] for a in tb["actuators"]])
    hi = np.array([a["ctrlrange"][1] for a in tb["actuators"]])
    u = QPOS_HOME[7:] + scale * rng.normal(size=(N, 1, 16)) * np.linspace(0.3, 1.0, H)[None, :, None]
    return np.clip(u, lo - 0.2, hi + 0.2)  # slightly outside ctrlrange: the kernel clamps like mj_fwdActuation


@pytest.mark.parametrize("N,H,scale", [(8, 12, 0.0), (24, 40, 0.5), (33, 25, 1.0)])
def test_leap_rollout_matches_oracle(engines, N, H, scale):
    """Contract A on the reduced leap model: 22-dof articulated dynamics, cube-hand contacts with elliptic cones,
    friction loss, joint limits, implicitfast.  Tolerances tiered by horizon (contact dynamics amplify rounding)."""
    from judo_b200.tasks.leap_cube import QPOS_HOME

    om, tb = _leap_oracle()
    rng = np.random.default_rng(N)
    eng = engines("leap_cube", N)
    x0 = np.concatenate([QPOS_HOME, np.zeros(22)])
    controls = _leap_controls(tb, rng, N, H, scale)
    overflows_before = eng.contact_overflows  # process-wide counter
    states, sensors = eng.rollout(x0, controls)
    s_ref, e_ref = om.rollout(x0, controls)
    assert np.all(np.isfinite(states))
    err = np.abs(states - s_ref).max(axis=(0, 2))
    print("leap state error by step:", err[:: max(1, H // 8)])
    assert err[: min(H, 5)].max() < 1e-9          # free fall + first steps: bit-near
    np.testing.assert_allclose(states[..., :23], s_ref[..., :23], rtol=0, atol=1e-5)   # positions over the whole horizon
    np.testing.assert_allclose(sensors, e_ref, rtol=0, atol=1e-5)
    assert eng.contact_overflows == overflows_before  # no contact was dropped (buffer: 24 per step)


def test_leap_plan_costs_and_reward_match_oracle(engines):
    from judo_b200.spline import spline_basis
    from judo_b200.tasks.leap_cube import QPOS_HOME

    om, tb = _leap_oracle()
    rng = np.random.default_rng(5)
    N, H, K = 48, 40, 4
    eng = engines("leap_cube", N)
    x0 = np.concatenate([QPOS_HOME, np.zeros(22)])
    lo = np.array([a["ctrlrange"][0] for a in tb["actuators"]])
    hi = np.array([a["ctrlrange"][1] for a in tb["actuators"]])
    nominal = np.tile(QPOS_HOME[7:], (K, 1))
    knots = np.clip(nominal + 0.2 * 4.0 * np.linspace(0.25, 1, K)[:, None] * rng.normal(size=(N, K, 16)), lo, hi)
    times = np.linspace(0, 0.4, K)
    query = 0.01 * np.arange(H)
    basis = spline_basis(times, query, "cubic")
    gq = rng.normal(size=4)
    gq /= np.linalg.norm(gq)
    params = np.concatenate([[100.0, 0.1], gq, [0.0, 0.03, 0.1]])
    reward, cost = eng.plan_costs(x0, knots, basis, params, want_cost_matrix=True)
    ctrl = op.make_spline(times, knots, "cubic")(query)
    states, _ = om.rollout(x0, ctrl)
    ref = op.leap_cube_reward(states, gq)
    np.testing.assert_allclose(reward, ref, rtol=1e-4, atol=1e-4)          # north_star's bar
    print("leap reward max abs err", np.abs(reward - ref).max(), "median", np.median(np.abs(reward - ref)))
    np.testing.assert_allclose(-cost.astype(np.float64).mean(1), reward, rtol=1e-5)
    np.testing.assert_allclose(eng.reward(states, ctrl, params), ref, rtol=1e-12)
    res = eng.plan_step(x0, knots, basis, params, "mppi", np.array([0.0025]), want_rewards=True, n_elite=3)
    np.testing.assert_array_equal(res["rewards"], reward)
    np.testing.assert_allclose(res["nominal"], op.mppi_update(knots, reward, 0.0025), rtol=1e-10, atol=1e-12)
    np.testing.assert_array_equal(res["elite"], np.argsort(reward)[-3:][::-1])


def test_leap_rewards_match_reference_golden(engines, golden):
    g = golden("rewards")
    from judo_b200.tasks.leap_cube import LeapCube

    t = LeapCube()
    t.engine = engines("leap_cube", 6)
    np.testing.assert_allclose(t.reward(g["leap_states"], None, None, {}), g["leap_rewards_default_goal"], rtol=1e-12)
    np.testing.assert_allclose(t.reward(g["leap_states"], None, None, {"goal_quat": g["leap_goal_quat"]}), g["leap_rewards"], rtol=1e-12)


def test_leap_cube_down_variant_matches_oracle(engines):
    """the palm-down variant is the same kernel with the constant table of leap_cube_palm_down.xml."""
    from judo_b200.tasks.leap_cube import QPOS_HOME_DOWN, LeapCubeDown, reduced_collision_model
    from oracle.mjc import load_table

    tb = load_table("leap_cube_down")
    geoms, pairs = reduced_collision_model(tb)
    om = OracleModel(tb, pairs=pairs, geoms=geoms)
    rng = np.random.default_rng(9)
    N, H = 16, 20
    eng = engines("leap_cube_down", N)
    x0 = np.concatenate([QPOS_HOME_DOWN, np.zeros(22)])
    controls = QPOS_HOME_DOWN[7:] + 0.4 * rng.normal(size=(N, 1, 16)) * np.linspace(0.3, 1, H)[None, :, None]
    s, e = eng.rollout(x0, controls)
    s_ref, e_ref = om.rollout(x0, controls)
    np.testing.assert_allclose(s, s_ref, rtol=0, atol=1e-7)
    np.testing.assert_allclose(e, e_ref, rtol=0, atol=1e-7)
    t = LeapCubeDown()
    assert t.config.w_rot == 0.05 and t.name == "leap_cube_down" and np.allclose(t.goal_pos, [-0.04, -0.035, -0.065])
