"""CPU: the oracle restatements against the golden vectors produced by the UNMODIFIED reference (tools/gen_golden.py)."""
import numpy as np
import pytest

from oracle import plan as op
from oracle.mjc import OracleModel, load_table


def test_spline_oracle_matches_reference(golden):
    g = golden("spline")
    for i in range(int(g["ncases"])):
        sp = op.make_spline(g[f"s{i}_times"], g[f"s{i}_knots"], str(g[f"s{i}_kind"]))
        np.testing.assert_allclose(sp(g[f"s{i}_query"]), g[f"s{i}_out"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(sp(g[f"s{i}_query_shift"]), g[f"s{i}_out_shift"], rtol=0, atol=1e-13)


def test_rewards_oracle_matches_reference(golden):
    g = golden("rewards")
    np.testing.assert_allclose(op.cartpole_reward(g["cartpole_states"], g["cartpole_controls"]), g["cartpole_rewards"], rtol=1e-14)
    np.testing.assert_allclose(op.cylinder_push_reward(g["cylinder_push_states"], g["cylinder_push_controls"]), g["cylinder_push_rewards"], rtol=1e-14)
    np.testing.assert_allclose(op.cylinder_push_reward(g["cylinder_push_states"], g["cylinder_push_controls"], goal_pos=g["cylinder_push_goal"]),
                               g["cylinder_push_rewards_goal"], rtol=1e-14)
    np.testing.assert_allclose(op.leap_cube_reward(g["leap_states"]), g["leap_rewards_default_goal"], rtol=1e-13)
    np.testing.assert_allclose(op.leap_cube_reward(g["leap_states"], g["leap_goal_quat"]), g["leap_rewards"], rtol=1e-13)
    np.testing.assert_allclose(op.quat_diff_so3(g["quat_u"], g["quat_v"]), g["quat_diff_so3"], rtol=0, atol=1e-14)


def test_optimizers_oracle_matches_reference(golden, temp_np_seed):
    g = golden("optimizers")
    for ci in range(int(g["ncases"])):
        name, task, nu, N, K, ramp, nr = g[f"c{ci}_meta"]
        N, ramp, nr = int(N), bool(int(ramp)), float(nr)
        with temp_np_seed(7 + ci):
            for it in range(2):
                nominal, rewards = g[f"c{ci}_nominal_in{it}"], g[f"c{ci}_rewards{it}"]
                if name == "cem":
                    knots, sig = op.cem_sample(nominal, g[f"c{ci}_sigma_in{it}"], N, ramp, nr, float(g[f"c{ci}_sigma_min"]), float(g[f"c{ci}_sigma_max"]))
                    nom, sig_out = op.cem_update(knots, rewards, int(g[f"c{ci}_num_elites"]), float(g[f"c{ci}_sigma_min"]), float(g[f"c{ci}_sigma_max"]))
                    np.testing.assert_array_equal(sig_out, g[f"c{ci}_sigma_out{it}"])
                else:
                    knots = op.sample_fixed_sigma(nominal, N, float(g[f"c{ci}_sigma"]), ramp, nr)
                    nom = op.mppi_update(knots, rewards, float(g[f"c{ci}_temperature"])) if name == "mppi" else op.ps_update(knots, rewards)
                np.testing.assert_array_equal(knots, g[f"c{ci}_knots{it}"])
                np.testing.assert_array_equal(nom, g[f"c{ci}_nominal_out{it}"])


def test_fr3_reward_and_phase_oracle_match_reference(golden):
    """FR3Pick.reward (four phases, custom config) and FR3Pick.pre_rollout's phase machine as executed by the reference."""
    g = golden("rewards_fr3")
    s, e = g["fr3_states"], g["fr3_sensors"]
    for ph in range(4):
        np.testing.assert_allclose(op.fr3_pick_reward(s, e, ph), g[f"fr3_rewards_phase{ph}"], rtol=1e-13)
    gx, gy, pick, wc = g["fr3_custom"]
    np.testing.assert_allclose(op.fr3_pick_reward(s, e, 1, goal_pos=(gx, gy), pick_height=pick, w_global=(0.25, wc, 0.005, 2.0)),
                               g["fr3_rewards_custom"], rtol=1e-13)
    assert [op.fr3_pick_phase(x) for x in g["fr3_phase_states"]] == list(g["fr3_phases"])


@pytest.mark.parametrize("tag", ["cartpole_ps", "cartpole_mppi", "cylinder_push_cem", "leap_cube_mppi", "fr3_pick_cem"])
def test_plan_golden_is_self_consistent(golden, tag):
    """The stored reference plan steps: oracle physics reproduces the stored states; oracle.plan reproduces
    controls -> rewards -> nominal -> traces from the stored candidates."""
    g = golden("plan_" + tag)
    task, opt, N, horizon, seed, order, max_traces = g["meta"]
    table = load_table(task)
    if task == "leap_cube":
        from judo_b200.tasks.leap_cube import reduced_collision_model

        geoms, pairs = reduced_collision_model(table)
        om = OracleModel(table, pairs=pairs, geoms=geoms)
    elif task == "fr3_pick":
        from tests.fr3_cases import oracle_model

        om = oracle_model()
    else:
        om = OracleModel(table)
    trace_adrs = [s["adr"] for s in table["sensors"] if s["type"] in ("framepos", "framepos_body") and "trace" in s["name"]]
    dt = table["opt"]["timestep"]
    for p in range(3):
        t = float(g[f"p{p}_time"])
        H = g[f"p{p}_rollout_controls"].shape[1]
        sp = op.make_spline(g[f"p{p}_times_out"], g[f"p{p}_candidate_knots"], order)
        ctrl = sp(t + dt * np.arange(H))
        np.testing.assert_allclose(ctrl, g[f"p{p}_rollout_controls"], rtol=0, atol=1e-14)
        states, sensors = om.rollout(g[f"p{p}_x0"], ctrl)
        np.testing.assert_allclose(states, g[f"p{p}_states"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(sensors, g[f"p{p}_sensors"], rtol=0, atol=1e-12)
        if task == "cartpole":
            rew = op.cartpole_reward(states, ctrl)
        elif task == "cylinder_push":
            rew = op.cylinder_push_reward(states, ctrl)
        elif task == "fr3_pick":
            assert op.fr3_pick_phase(g[f"p{p}_x0"]) == int(g[f"p{p}_phase"])
            rew = op.fr3_pick_reward(states, sensors, int(g[f"p{p}_phase"]))
        else:
            rew = op.leap_cube_reward(states, g["goal_quat"])
        np.testing.assert_allclose(rew, g[f"p{p}_rewards"], rtol=1e-13)
        tr = op.update_traces(g[f"p{p}_sensors"], g[f"p{p}_rewards"], trace_adrs, int(max_traces))
        np.testing.assert_array_equal(tr, g[f"p{p}_traces"])


def test_oracle_mass_matrix_and_bias_against_lagrangian():
    """Independent check of the C oracle's kinematics/CRBA/RNE: M equals the numpy Jacobian form and the bias force
    equals Mdot v - 0.5 d(v'Mv)/dq + dV/dq by finite differences (hinge/slide dofs)."""
    from judo_b200 import mjcf

    rng = np.random.default_rng(0)
    for task in ("cartpole", "leap_cube"):
        tb = load_table(task)
        om = OracleModel(tb, pairs=[])
        q = np.array(tb["qpos0"], dtype=np.float64)
        v = np.zeros(tb["nv"])
        for j in tb["joints"]:
            if j["type"] in ("hinge", "slide"):
                q[j["qposadr"]] = rng.uniform(-0.3, 1.0)
                v[j["dofadr"]] = rng.normal() * 2
        f = om.forward(q, v, np.zeros(tb["nu"]))
        M2 = mjcf.mass_matrix(tb, mjcf.forward_kinematics(tb, q))
        np.testing.assert_allclose(f["M"], M2, rtol=0, atol=1e-15)
        eps = 1e-6
        grav = np.array(tb["opt"]["gravity"])

        def Mq(qq):
            return mjcf.mass_matrix(tb, mjcf.forward_kinematics(tb, qq))

        def V(qq):
            k = mjcf.forward_kinematics(tb, qq)
            return -sum(b["mass"] * grav @ k["xipos"][i] for i, b in enumerate(tb["bodies"]))

        c = np.zeros(tb["nv"])
        dM = {}
        idx = [(j["dofadr"], j["qposadr"]) for j in tb["joints"] if j["type"] in ("hinge", "slide")]
        for d, qa in idx:
            qp, qm = q.copy(), q.copy()
            qp[qa] += eps
            qm[qa] -= eps
            dM[d] = (Mq(qp) - Mq(qm)) / (2 * eps)
            c[d] += -0.5 * v @ dM[d] @ v + (V(qp) - V(qm)) / (2 * eps)
        c += sum(dM[d] * v[d] for d, _ in idx) @ v
        sel = [d for d, _ in idx]
        np.testing.assert_allclose(f["qfrc_bias"][sel], c[sel], rtol=0, atol=5e-9)


def test_oracle_constraint_solution_is_kkt_point():
    """Newton output satisfies the primal optimality condition M(a - a_smooth) = J' f with f from the constraint law."""
    tb = load_table("cylinder_push")
    om = OracleModel(tb)
    f = om.forward(np.array([-0.45, 0.02, 0, 0.0]), np.array([1.0, 0, 0, 0]), np.array([1.0, 0.0]))
    assert f["ncon"] == 1 and f["nefc"] == 4
    np.testing.assert_allclose(f["M"] @ (f["qacc"] - f["qacc_smooth"]), f["qfrc_constraint"], rtol=0, atol=1e-7)
    assert f["qfrc_constraint"][2] > 0 and f["qfrc_constraint"][0] < 0  # pushes the cart away, the pusher back
    tb = load_table("cartpole")
    om = OracleModel(tb)
    f = om.forward(np.array([1.85, 3.0]), np.array([0.5, 0.0]), np.array([1.8]))
    assert f["nefc"] == 1 and f["qfrc_constraint"][0] < 0
    np.testing.assert_allclose(f["M"] @ (f["qacc"] - f["qacc_smooth"]), f["qfrc_constraint"], rtol=0, atol=1e-9)
