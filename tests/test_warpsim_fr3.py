"""No-GPU tier: the DEVICE CODE of judo_b200/csrc/fr3.cuh on the CPU SIMT emulator (tests/warpsim — test infrastructure, not a
product path) against the C oracle and the reference's golden rewards.  The `-m gpu` tier repeats these through the C ABI."""
import ctypes

import numpy as np
import pytest

from judo_b200.consts import task_consts
from oracle import plan as op
from tests.fr3_cases import oracle_model, scenario

P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731


@pytest.fixture(scope="module")
def sim():
    from tests import warpsim

    return warpsim.lib()


@pytest.fixture(scope="module")
def fr3():
    return np.ascontiguousarray(task_consts("fr3_pick")), oracle_model()


def _rollout(sim, consts, x0, u, wpb=2, reverse=0):
    N, H, _ = u.shape
    s, e = np.zeros((N, H, 31)), np.zeros((N, H, 14))
    sim.sim_fr3_rollout(P(consts), P(x0), int(x0.ndim == 2), P(u), N, H, P(s), P(e), wpb, 3, reverse)
    return s, e


@pytest.mark.parametrize("name,N,H", [("home", 3, 12), ("grasp", 3, 40), ("wild", 3, 30), ("press", 2, 30)])
def test_fr3_rollout_kernel_on_emulator_matches_oracle(sim, fr3, name, N, H):
    consts, om = fr3
    assert sim.sim_fr3_nconsts() == consts.size
    x0, u = scenario(name, N, H)
    s_ref, e_ref = om.rollout(x0, u)
    s, e = _rollout(sim, consts, x0, u)
    np.testing.assert_allclose(s, s_ref, rtol=0, atol=1e-9)
    np.testing.assert_allclose(e, e_ref, rtol=0, atol=1e-9)
    if name == "grasp":
        assert s_ref[:, -1, 2].min() > 0.022, "the cube must actually be lifted (pad-object friction contacts)"
        assert np.abs(s_ref[:, :, 14] - s_ref[:, :, 15]).max() < 5e-3, "finger equality"
        assert (e_ref[:, :, 0].min(1) <= 0).all() and (e_ref[:, :, 1].min(1) <= 0).all(), "both fingers touch the object"
    if name == "press":
        assert (e_ref[:, :, 2:4].min(axis=(1, 2)) <= 0).all(), "the fingertips must reach the table (pad-table contacts)"


def test_fr3_lane_order_and_block_shape_do_not_matter(sim, fr3):
    """Forward vs reverse lane order, 1 vs 3 warps per block: a difference would mean a missing __syncwarp / __syncthreads."""
    consts, _ = fr3
    x0, u = scenario("grasp", 3, 24, seed=3)
    a = _rollout(sim, consts, x0, u, wpb=1, reverse=0)
    b = _rollout(sim, consts, x0, u, wpb=3, reverse=1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("phase", [0, 1, 2, 3])
def test_fr3_fused_cost_kernel_on_emulator_matches_oracle(sim, fr3, phase):
    from judo_b200.spline import spline_basis
    from judo_b200.tasks.fr3_pick import FR3Pick, Phase

    consts, om = fr3
    N, H, K = 3, 20, 4
    x0, u = scenario("grasp", N, K, seed=5)  # K knots per rollout, used as the spline's knots
    knots = np.ascontiguousarray(u)
    times = np.linspace(0, 0.08, K)
    basis = np.ascontiguousarray(spline_basis(times, 0.004 * np.arange(H), "linear"))
    task = FR3Pick.__new__(FR3Pick)
    task.config = FR3Pick.config_t()
    task.phase = Phase(phase)
    task.arm_pos_slice = slice(7, 16)
    params = task.cost_params()
    cost, rew, trace = np.zeros((N, H), dtype=np.float32), np.zeros(N), np.zeros((N, H, 6))
    sim.sim_fr3_plan_costs(P(consts), P(x0), P(knots), N, K, P(basis), H, P(params), P(cost), P(rew), 2, 3, 0, P(trace))
    controls = np.einsum("hk,nkj->nhj", basis, knots)
    states, sensors = om.rollout(x0, controls)
    np.testing.assert_allclose(trace, sensors[..., 8:14], rtol=0, atol=1e-9)  # trace capture: trace_object + trace_grasp_site of every rollout
    ref = op.fr3_pick_reward(states, sensors, phase)
    np.testing.assert_allclose(rew, ref, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(-cost.sum(1), ref, rtol=1e-5, atol=1e-5)  # f32 per-step costs


def test_fr3_reward_kernel_on_emulator_matches_reference_golden(sim, golden):
    """FR3Pick.reward as the unmodified reference computes it (tools/gen_golden.py) for the four phases."""
    from judo_b200.tasks.fr3_pick import FR3Pick, Phase

    g = golden("rewards_fr3")
    s, e = np.ascontiguousarray(g["fr3_states"]), np.ascontiguousarray(g["fr3_sensors"])
    task = FR3Pick.__new__(FR3Pick)
    task.config = FR3Pick.config_t()
    task.arm_pos_slice = slice(7, 16)
    for ph in range(4):
        task.phase = Phase(ph)
        out = np.zeros(len(s))
        sim.sim_fr3_reward(P(s), P(e), s.shape[0], s.shape[1], P(task.cost_params()), P(out))
        np.testing.assert_allclose(out, g[f"fr3_rewards_phase{ph}"], rtol=1e-12)
    gx, gy, ph_, wc = g["fr3_custom"]
    task.config.goal_pos = np.array([gx, gy])
    task.config.pick_height = float(ph_)
    task.config.global_weights.w_coll = float(wc)
    task.phase = Phase.MOVE
    out = np.zeros(len(s))
    sim.sim_fr3_reward(P(s), P(e), s.shape[0], s.shape[1], P(task.cost_params()), P(out))
    np.testing.assert_allclose(out, g["fr3_rewards_custom"], rtol=1e-12)
