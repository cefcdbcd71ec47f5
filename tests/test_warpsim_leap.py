"""No-GPU tier: the DEVICE CODE of judo_b200/csrc/leap.cuh executed on the CPU SIMT emulator (tests/warpsim — test
infrastructure, not a product path) against the C oracle.  This checks the kernel's logic (lane mappings, shuffles,
barriers) here where no GPU exists; the `-m gpu` tier repeats the comparison on the real device through the C ABI.
Running the lanes of each warp in forward and in reverse order must give identical results: a difference means a missing
__syncwarp/__syncthreads."""
import ctypes

import numpy as np
import pytest

from judo_b200.consts import task_consts
from judo_b200.tasks.leap_cube import QPOS_HOME, reduced_collision_model
from oracle import plan as op
from oracle.mjc import OracleModel, load_table

P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731


@pytest.fixture(scope="module")
def sim():
    from tests import warpsim

    return warpsim.lib()


@pytest.fixture(scope="module")
def leap():
    tb = load_table("leap_cube")
    geoms, pairs = reduced_collision_model(tb)
    return np.ascontiguousarray(task_consts("leap_cube")), OracleModel(tb, pairs=pairs, geoms=geoms)


def test_leap_rollout_kernel_on_emulator_matches_oracle(sim, leap):
    consts, om = leap
    assert sim.sim_leap_nconsts() == consts.size
    rng = np.random.default_rng(0)
    N, H = 3, 8
    x0 = np.concatenate([QPOS_HOME, np.zeros(22)])
    x0[2] = 0.07  # the cube starts in contact with the hand
    u = np.ascontiguousarray(QPOS_HOME[7:] + 0.3 * rng.normal(size=(N, H, 16)))
    s_ref, e_ref = om.rollout(x0, u)
    outs = []
    for reverse, wpb in ((0, 1), (1, 2)):
        s, e = np.zeros((N, H, 45)), np.zeros((N, H, 31))
        sim.sim_leap_rollout(P(consts), P(x0), 0, P(u), N, H, P(s), P(e), wpb, 3, reverse)
        np.testing.assert_allclose(s, s_ref, rtol=0, atol=1e-10)
        np.testing.assert_allclose(e, e_ref, rtol=0, atol=1e-10)
        outs.append(s)
    assert np.array_equal(outs[0], outs[1])  # lane order / block shape must not matter


def test_leap_fused_cost_kernel_on_emulator_matches_oracle(sim, leap):
    consts, om = leap
    rng = np.random.default_rng(1)
    N, H, K = 2, 6, 4
    x0 = np.concatenate([QPOS_HOME, np.zeros(22)])
    knots = np.ascontiguousarray(QPOS_HOME[7:] + 0.2 * rng.normal(size=(N, K, 16)))
    basis = np.ascontiguousarray(rng.random((H, K)))
    basis /= basis.sum(1, keepdims=True)
    gq = rng.normal(size=4)
    params = np.concatenate([[100.0, 0.1], gq / np.linalg.norm(gq), [0.0, 0.03, 0.1]])
    cost, rew, trace = np.zeros((N, H), dtype=np.float32), np.zeros(N), np.zeros((N, H, 15))
    sim.sim_leap_plan_costs(P(consts), P(x0), P(knots), N, K, P(basis), H, P(params), P(cost), P(rew), 2, 3, 0, P(trace))
    controls = np.einsum("hk,nkj->nhj", basis, knots)
    np.testing.assert_allclose(trace, om.rollout(x0, controls)[1][..., 16:31], rtol=0, atol=1e-10)  # trace capture: the 5 framepos sensors
    ref = op.leap_cube_reward(om.rollout(x0, controls)[0], params[2:6], 100.0, 0.1)
    np.testing.assert_allclose(rew, ref, rtol=0, atol=1e-10)


def test_leap_hand_hand_contacts_on_emulator_match_oracle(sim, leap):
    """Fingers driven into each other and into the palm (targets drawn over the whole actuator range): finger-finger contacts take the
    dense Newton direction, contacts inside one finger / against the palm stay on the arrow path.  The pair list is MuJoCo's
    (params_and_default.xml:76-101 excludes applied); dropping the hand-hand pairs must change these rollouts."""
    consts, om = leap
    tb = load_table("leap_cube")
    geoms0, pairs0 = reduced_collision_model(tb, hand_hand=False)
    om0 = OracleModel(tb, pairs=pairs0, geoms=geoms0)
    lo = np.array([a["ctrlrange"][0] for a in tb["actuators"]])
    hi = np.array([a["ctrlrange"][1] for a in tb["actuators"]])
    rng = np.random.default_rng(5)
    N, H = 48, 30
    x0 = np.concatenate([QPOS_HOME, np.zeros(22)])
    x0[2] = 0.07
    u = np.ascontiguousarray(np.repeat(lo + (hi - lo) * rng.random((N, 1, 16)), H, axis=1))
    s_ref, e_ref = om.rollout(x0, u)
    matter = np.where(np.abs(s_ref - om0.rollout(x0, u)[0]).max(axis=(1, 2)) > 1e-6)[0]
    assert matter.size >= 6
    pick = np.concatenate([matter[:8], [0, 1]])
    u, s_ref, e_ref = np.ascontiguousarray(u[pick]), s_ref[pick], e_ref[pick]
    sim.sim_leap_set_prof(1)
    outs = []
    for reverse, wpb in ((0, 2), (1, 3)):
        s, e = np.zeros_like(s_ref), np.zeros_like(e_ref)
        sim.sim_leap_rollout(P(consts), P(x0), 0, P(u), len(pick), H, P(s), P(e), wpb, 3, reverse)
        prof = (ctypes.c_ulonglong * 20)()
        sim.sim_leap_read_prof(prof)
        assert prof[16] > 0 and prof[17] > prof[18] > 0  # dense directions ran; finger-finger AND one-finger / palm contacts occurred
        np.testing.assert_allclose(s, s_ref, rtol=0, atol=1e-9)
        np.testing.assert_allclose(e, e_ref, rtol=0, atol=1e-9)
        outs.append(s)
    sim.sim_leap_set_prof(0)
    assert np.array_equal(outs[0], outs[1])


def test_leap_step_with_more_than_32_contacts_on_emulator(sim, leap):
    """Rollout 491 of the C4 bench problem (bench.problem, seed 42) reaches 31 and 36 simultaneous contacts at steps 35 / 39 once the
    hand-hand pairs are in: the line search's low lanes own a second contact there."""
    import bench

    consts, om = leap
    w = bench.WORKLOADS["leap_cube_mppi"]
    _, _, x0, knots, basis, _, _ = bench.problem(w, w["n_rollouts"])
    u = np.ascontiguousarray(np.einsum("hk,nkj->nhj", basis, knots[[491, 0]]))
    s_ref, e_ref = om.rollout(x0, u)
    fwd = om.forward(s_ref[0, 38, :23], s_ref[0, 38, 23:], u[0, 39])
    assert 32 < fwd["ncon"] <= 40
    s, e = np.zeros_like(s_ref), np.zeros_like(e_ref)
    sim.sim_leap_rollout(P(consts), P(x0), 0, P(u), 2, w["H"], P(s), P(e), 2, 3, 1)
    np.testing.assert_allclose(s, s_ref, rtol=0, atol=1e-9)
    np.testing.assert_allclose(e, e_ref, rtol=0, atol=1e-9)
