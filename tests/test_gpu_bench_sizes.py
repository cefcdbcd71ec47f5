"""-m gpu: parity AT THE BENCHED SIZES.  The inputs are bench.py's own (`bench.problem`: seed 42, the task's reset
distribution, candidates sampled as the reference samples them); the fused kernel runs the full batch and the oracle re-computes a fixed
random sample of >= 256 of its rollouts (plus rollout 0, the un-noised nominal)."""
import numpy as np
import pytest

import bench
from oracle import plan as op

pytestmark = pytest.mark.gpu


def _sample(N: int, n: int, seed: int = 7) -> np.ndarray:
    idx = np.random.default_rng(seed).choice(N, size=min(n, N), replace=False)
    idx[0] = 0
    return np.sort(idx)


def _oracle_reward(w, om, x0, controls, params):
    states, sensors = om.rollout(x0, controls)
    if w["task"] == "cartpole":
        return op.cartpole_reward(states, controls, *params), states, sensors
    if w["task"] == "cylinder_push":
        return op.cylinder_push_reward(states, controls, params[0], params[1], params[2], params[3], params[4:6]), states, sensors
    return op.leap_cube_reward(states, params[2:6], params[0], params[1]), states, sensors


@pytest.mark.parametrize("wname,n_check,rtol,atol", [("cartpole_mppi", 512, 1e-9, 1e-9), ("cylinder_push_cem", 512, 1e-6, 1e-7),
                                                     ("leap_cube_mppi", 256, 1e-6, 1e-8)])
def test_fused_kernel_matches_oracle_at_the_benched_size(wname, n_check, rtol, atol):
    from judo_b200.engine import Engine

    w = bench.WORKLOADS[wname]
    N = w["n_rollouts"]
    task, opt, x0, knots, basis, params, _ = bench.problem(w, N)
    assert knots.shape == (N, w["K"], task.nu) and basis.shape == (w["H"], w["K"])
    eng = Engine(w["task"], N)
    ov0 = eng.contact_overflows
    reward, cost = eng.plan_costs(x0, knots, basis, params, want_cost_matrix=True)
    assert reward.shape == (N,) and cost.shape == (N, w["H"]) and np.isfinite(reward).all()
    # the cost matrix and the reward are the same numbers (sum / mean over time, per task)
    tot = cost.astype(np.float64).sum(axis=1)
    np.testing.assert_allclose(reward, -(tot / w["H"] if w["task"] == "leap_cube" else tot), rtol=2e-6, atol=1e-6)
    # fused update over the full batch == the reference's update of the kernel's own rewards
    res = eng.plan_step(x0, knots, basis, params, w["optimizer"], opt.fused_params(), want_rewards=True, n_elite=5)
    np.testing.assert_array_equal(res["rewards"], reward)
    if w["optimizer"] == "mppi":
        np.testing.assert_allclose(res["nominal"], op.mppi_update(knots, reward, opt.temperature), rtol=0, atol=1e-10)
    else:
        nom, sig = op.cem_update(knots, reward, opt.num_elites, opt.sigma_min, opt.sigma_max)
        np.testing.assert_allclose(res["nominal"], nom, rtol=0, atol=1e-12)
        np.testing.assert_allclose(res["sigma"], sig, rtol=0, atol=1e-12)
    np.testing.assert_array_equal(res["elite"], np.argsort(reward, kind="stable")[::-1][:5])
    # oracle on a fixed sample of the batch
    sel = _sample(N, n_check)
    controls = np.einsum("hk,nkj->nhj", basis, knots[sel])
    om = bench._oracle_model(w["task"])
    r_ref, s_ref, e_ref = _oracle_reward(w, om, x0, controls, params)
    np.testing.assert_allclose(reward[sel], r_ref, rtol=rtol, atol=atol)
    # contract A on the same sample: states and sensors
    eng_a = Engine(w["task"], len(sel))
    s_gpu, e_gpu = eng_a.rollout(x0, controls)
    tol = {"cartpole": 1e-9, "cylinder_push": 1e-6, "leap_cube": 1e-7}[w["task"]]
    np.testing.assert_allclose(s_gpu, s_ref, rtol=0, atol=tol)
    np.testing.assert_allclose(e_gpu, e_ref, rtol=0, atol=tol)
    if w["task"] == "leap_cube":
        assert eng.contact_overflows - ov0 == 0 and eng_a.contact_overflows == 0, "the per-step contact buffer must hold every contact of the benched scenario"
        ncon_active = (np.abs(s_ref[:, 1:, :3] - s_ref[:, :-1, :3]).max() > 0)
        assert ncon_active
    eng.close(); eng_a.close()


def test_contact_overflow_counter_is_per_handle():
    from judo_b200.engine import Engine

    a, b = Engine("leap_cube", 4), Engine("leap_cube", 4)
    assert a.contact_overflows == 0 and b.contact_overflows == 0
    assert Engine("cartpole", 4).contact_overflows == 0
    a.close(); b.close()
