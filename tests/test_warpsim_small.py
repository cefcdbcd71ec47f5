"""No-GPU tier: the fused thread-per-rollout kernel (cartpole, cylinder_push: spline + closed-form mj_step + cost + the optimizer
update epilogue in ONE launch) executed on the CPU SIMT emulator (tests/warpsim — test infrastructure, not a product path) against
the golden plan steps produced by the unmodified reference Controller and against the C oracle."""
import ctypes

import numpy as np
import pytest

from judo_b200.consts import task_consts
from judo_b200.spline import spline_basis
from oracle import plan as op
from oracle.mjc import OracleModel, load_table

P = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
TASK = {"cartpole": 0, "cylinder_push": 1}
OPT = {"mppi": 0, "cem": 1, "ps": 2}


@pytest.fixture(scope="module")
def sim():
    from tests import warpsim

    return warpsim.lib()


def _plan_step(sim, task, x0, knots, basis, params, opt, opt_params, n_elite=3, threads=32, reverse=0, want_cost=True):
    N, K, nu = knots.shape
    H = basis.shape[0]
    consts = np.ascontiguousarray(task_consts(task))
    knots, basis, x0 = np.ascontiguousarray(knots), np.ascontiguousarray(basis), np.ascontiguousarray(x0)
    params, opt_params = np.ascontiguousarray(params, dtype=np.float64), np.ascontiguousarray(opt_params + [0.0], dtype=np.float64)
    cost = np.zeros((N, H), dtype=np.float32) if want_cost else None
    rew, nom, sig, el = np.zeros(N), np.zeros((K, nu)), np.zeros((K, nu)), np.full(8, -1.0)
    sim.sim_plan_step(TASK[task], P(consts), P(x0), P(knots), N, K, P(basis), H, P(params), OPT[opt], P(opt_params), n_elite, threads, P(cost),
                      P(rew), P(nom), P(sig), P(el), reverse)
    return rew, nom, sig, el[:n_elite].astype(int), cost


@pytest.mark.parametrize("tag", ["cartpole_ps", "cartpole_mppi", "cylinder_push_cem"])
def test_fused_plan_step_on_emulator_reproduces_reference_plan_steps(sim, golden, tag):
    from judo_b200.tasks import get_registered_tasks

    g = golden("plan_" + tag)
    task, opt, N, horizon, seed, order, max_traces = g["meta"]
    task, opt = str(task), str(opt)
    tb = load_table(task)
    dt = tb["opt"]["timestep"]
    t = get_registered_tasks()[task][0]()
    for p in range(3):
        knots = g[f"p{p}_candidate_knots"]
        H = g[f"p{p}_rollout_controls"].shape[1]
        basis = spline_basis(g[f"p{p}_times_out"], float(g[f"p{p}_time"]) + dt * np.arange(H), str(order))
        opt_params = {"mppi": [0.05], "ps": [], "cem": [2, 0.1, 1.0]}[opt]
        rew, nom, sig, elite, cost = _plan_step(sim, task, g[f"p{p}_x0"], knots, basis, t.cost_params(), opt, opt_params, threads=32 if p < 2 else 64,
                                                reverse=p % 2)
        np.testing.assert_allclose(rew, g[f"p{p}_rewards"], rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(nom, g[f"p{p}_nominal_out"], rtol=0, atol=1e-8)
        if opt == "cem":
            np.testing.assert_allclose(sig, g[f"p{p}_sigma_out"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(rew[elite], np.sort(rew)[::-1][:3])  # the trace elites: best three, descending
        np.testing.assert_allclose(-cost.astype(np.float64).sum(1), rew, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("task,nu", [("cartpole", 1), ("cylinder_push", 2)])
def test_contract_a_rollout_on_emulator_matches_oracle(sim, task, nu):
    om = OracleModel(task)
    rng = np.random.default_rng(3)
    N, H = 40, 30
    x0 = (np.concatenate([np.array([1.0, np.pi]) + rng.normal(size=2), 0.1 * rng.normal(size=2)]) if task == "cartpole"
          else np.array([1.0, 0.1, 1.6, 0.2, 0, 0, 0, 0.0]))
    u = np.ascontiguousarray((2.5 if task == "cartpole" else 4.0) * rng.normal(size=(N, H, nu)))  # beyond the clamps; pusher rams the cart
    s, e = np.zeros((N, H, om.nq + om.nv)), np.zeros((N, H, om.nsensordata))
    consts = np.ascontiguousarray(task_consts(task))
    sim.sim_rollout(TASK[task], P(consts), P(x0), 0, P(u), N, H, P(s), P(e), 32)
    s_ref, e_ref = om.rollout(x0, u)
    tol = 1e-9 if task == "cartpole" else 1e-6
    np.testing.assert_allclose(s, s_ref, rtol=0, atol=tol)
    np.testing.assert_allclose(e, e_ref, rtol=0, atol=tol)


@pytest.mark.parametrize("opt,opt_params", [("mppi", [0.05]), ("cem", [3, 0.1, 1.0]), ("ps", [])])
def test_fused_epilogue_on_emulator_matches_reference_update_rules(sim, opt, opt_params):
    """Ragged N (tail warp partly empty, several blocks), ties, K = 5: the in-kernel reduction against oracle.plan (pinned to the
    reference's optimizers by tests/test_oracle_golden.py)."""
    rng = np.random.default_rng(9)
    N, K, H = 77, 5, 6
    knots = rng.normal(size=(N, K, 1))
    knots[13] = knots[40]  # identical candidates -> tied rewards
    basis = spline_basis(np.linspace(0, 0.24, K), 0.04 * np.arange(H), "linear")
    x0 = np.array([0.3, 2.8, 0.0, 0.1])
    params = np.array([10.0, 10.0, 0.1, 0.1, 0.01, 0.1])
    rew, nom, sig, elite, _ = _plan_step(sim, "cartpole", x0, knots, basis, params, opt, opt_params, n_elite=4, threads=32, want_cost=False)
    assert rew[13] == rew[40]
    if opt == "mppi":
        np.testing.assert_allclose(nom, op.mppi_update(knots, rew, 0.05), rtol=1e-12, atol=1e-14)
    elif opt == "cem":
        n2, s2 = op.cem_update(knots, rew, 3, 0.1, 1.0)
        np.testing.assert_allclose(nom, n2, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(sig, s2, rtol=1e-10, atol=1e-14)
    else:
        np.testing.assert_array_equal(nom, op.ps_update(knots, rew))
    np.testing.assert_allclose(np.sort(rew[elite])[::-1], np.sort(rew)[::-1][:4])


@pytest.mark.parametrize("task,N,K", [("cartpole", 70, 4), ("cylinder_push", 45, 6)])
def test_candidates_assembled_in_the_kernel_from_host_normals_are_bit_identical(sim, task, N, K):
    """SampleSpec.enabled == 2 (what b200mpc_controller_step launches when the step's block of normals was drawn and uploaded ahead of it):
    the kernel's clip(nominal + sigma * z) must equal NumPy's, bit for bit (two roundings, np.clip), row 0 the un-noised nominal, and the
    fused step must be the step over those candidates."""
    from judo_b200.tasks import get_registered_tasks

    t = get_registered_tasks()[task][0]()
    nu = t.nu
    rng = np.random.default_rng(3)
    H = 12
    nominal = 0.4 * rng.normal(size=(K, nu))
    sigma = 0.05 + rng.random((K, nu))
    z = rng.normal(size=(N - 1, K, nu))
    lo, hi = np.ascontiguousarray(t.actuator_ctrlrange[:, 0] * 0.2), np.ascontiguousarray(t.actuator_ctrlrange[:, 1] * 0.2)  # tight: clipping happens
    ref = np.concatenate([nominal[None], nominal + sigma * z])
    ref = np.clip(ref, lo, hi)
    assert (ref == lo).any() and (ref == hi).any()
    x0 = np.concatenate([t.data.qpos, t.data.qvel])
    basis = np.ascontiguousarray(rng.random((H, K)))
    basis /= basis.sum(1, keepdims=True)
    params = np.ascontiguousarray(t.cost_params(), dtype=np.float64)
    consts = np.ascontiguousarray(task_consts(task))
    rew, nom, kout = np.zeros(N), np.zeros((K, nu)), np.zeros((N, K, nu))
    sim.sim_plan_step_hostz.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 7 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                                               ctypes.c_double, ctypes.c_int] + [ctypes.c_void_p] * 3
    sim.sim_plan_step_hostz(TASK[task], P(consts), P(np.ascontiguousarray(x0)), P(np.ascontiguousarray(z)), P(np.ascontiguousarray(nominal)),
                            P(np.ascontiguousarray(sigma)), P(lo), P(hi), N, K, P(basis), H, P(params), 0.05, 32, P(rew), P(nom), P(kout))
    assert np.array_equal(kout, ref)
    rew2, nom2, _, _, _ = _plan_step(sim, task, x0, ref, basis, params, "mppi", [0.05], n_elite=0, want_cost=False)
    assert np.array_equal(rew, rew2) and np.array_equal(nom, nom2)
