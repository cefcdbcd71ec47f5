"""CPU: physical sanity of the restated dynamics (oracle/mjc).  The oracle cannot be pinned to MuJoCo binaries (absent), so
these tests pin it to closed-form physics instead: free fall, energy conservation, Newton's third law in contacts, static
equilibrium on the palm, soft joint limits, actuator clamps."""
import copy

import numpy as np

from oracle.mjc import OracleModel, load_table


def _leap(hand_hand=True):
    from judo_b200.tasks.leap_cube import QPOS_HOME, reduced_collision_model

    tb = load_table("leap_cube")
    geoms, pairs = reduced_collision_model(tb, hand_hand=hand_hand)
    return OracleModel(tb, pairs=pairs, geoms=geoms), tb, QPOS_HOME


def test_free_fall_is_the_semi_implicit_euler_parabola():
    om, tb, home = _leap()
    q = home.copy()
    q[:3] = [0.5, 0.5, 1.0]  # far from the hand: no contacts
    H, h, g = 30, tb["opt"]["timestep"], 9.81
    s, _ = om.rollout(np.concatenate([q, np.zeros(22)]), np.tile(home[7:], (1, H, 1)))
    k = np.arange(1, H + 1)
    np.testing.assert_allclose(s[0, :, 2], 1.0 - g * h * h * k * (k + 1) / 2, rtol=0, atol=1e-12)
    np.testing.assert_allclose(s[0, :, 23 + 2], -g * h * k, rtol=0, atol=1e-12)
    np.testing.assert_allclose(s[0, :, 3:7], np.tile([1, 0, 0, 0], (H, 1)), atol=1e-15)  # no spurious rotation


def test_spinning_free_cube_keeps_angular_momentum():
    om, tb, home = _leap()
    q = home.copy()
    q[:3] = [0.5, 0.5, 1.0]
    v = np.zeros(22)
    v[3:6] = [3.0, -2.0, 1.0]  # isotropic inertia: body angular velocity stays constant
    s, _ = om.rollout(np.concatenate([q, v]), np.tile(home[7:], (1, 50, 1)))
    np.testing.assert_allclose(s[0, :, 23 + 3:23 + 6], np.tile(v[3:6], (50, 1)), atol=1e-12)
    np.testing.assert_allclose(np.linalg.norm(s[0, :, 3:7], axis=1), 1.0, atol=1e-14)
    # rotation angle after 50 steps = |w| * t
    ang = 2 * np.arccos(np.clip(abs(s[0, -1, 3]), 0, 1))
    expected = np.linalg.norm(v[3:6]) * 50 * tb["opt"]["timestep"]
    assert abs(ang - (expected % (2 * np.pi))) < 1e-9 or abs(2 * np.pi - ang - (expected % (2 * np.pi))) < 1e-9


def test_cartpole_energy_is_conserved_without_damping_and_actuation():
    tb = copy.deepcopy(load_table("cartpole"))
    tb["opt"]["timestep"] = 0.001
    for d in tb["dofs"]:
        d["damping"] = 0.0
    tb["joints"][0]["damping"] = 0.0
    tb["actuators"][0]["kp"] = 0.0
    tb["joints"][0]["limited"] = False
    om = OracleModel(tb)
    x0 = np.array([0.0, 2.5, 0.3, -0.5])
    s, _ = om.rollout(x0, np.zeros((1, 2000, 1)))
    mc, mp, l, I = tb["bodies"][1]["mass"], tb["bodies"][2]["mass"], tb["bodies"][2]["ipos"][2], tb["bodies"][2]["inertia"][1]

    def energy(x):
        q0, th, v0, w = x[..., 0], x[..., 1], x[..., 2], x[..., 3]
        T = 0.5 * (mc + mp) * v0**2 + mp * l * np.cos(th) * v0 * w + 0.5 * (I + mp * l * l) * w**2
        return T + mp * 9.81 * l * np.cos(th)

    e = energy(np.vstack([x0, s[0]]))
    assert np.abs(e - e[0]).max() < 2e-3 * abs(e[0])  # first-order integrator, h = 1 ms: small bounded drift


def test_contact_forces_obey_newtons_third_law():
    om = OracleModel("cylinder_push")
    f = om.forward(np.array([-0.45, 0.02, 0.0, 0.0]), np.array([1.0, 0.0, 0.0, 0.0]), np.zeros(2))
    fc = f["qfrc_constraint"]
    assert f["ncon"] == 1 and f["contact_dist"][0] < 0
    np.testing.assert_allclose(fc[0:2] + fc[2:4], 0, atol=1e-9)      # equal and opposite
    n = np.array([0.45, -0.02]) / np.hypot(0.45, 0.02)                 # pusher -> cart
    assert fc[2:4] @ n > 0 and abs(fc[2] * n[1] - fc[3] * n[0]) < 1e-3 * (fc[2:4] @ n)  # repulsive, along the normal (mu = 1e-5)


def test_cube_rests_on_the_hand_in_static_equilibrium():
    om, tb, home = _leap()
    H = 300
    s, _ = om.rollout(np.concatenate([home, np.zeros(22)]), np.tile(home[7:], (1, H, 1)))
    assert np.abs(s[0, -1, 23:29]).max() < 0.05 and 0.03 < s[0, -1, 2] < 0.1   # settled on the palm/fingers
    f = om.forward(s[0, -1, :23], s[0, -1, 23:], home[7:])
    assert f["ncon"] >= 3
    mg = tb["bodies"][2]["mass"] * 9.81
    assert abs(f["qfrc_constraint"][2] - mg) < 0.1 * mg                         # contacts carry the weight
    assert np.all(f["contact_dist"] < 0) and np.all(f["contact_dist"] > -2e-3)  # soft contact: sub-2-mm penetration


def test_joint_limit_and_clamps():
    tb = load_table("leap_cube")
    om, _, home = _leap(hand_hand=False)  # every joint driven to its upper limit: without hand-hand contacts, so that only the limits stop them
    lo = np.array([a["ctrlrange"][0] for a in tb["actuators"]])
    hi = np.array([a["ctrlrange"][1] for a in tb["actuators"]])
    q = home.copy()
    q[:3] = [0.5, 0.5, 1.0]
    x0 = np.concatenate([q, np.zeros(22)])
    far = np.tile(hi + 5.0, (1, 150, 1))      # commands far beyond ctrlrange are clamped (mj_fwdActuation)
    edge = np.tile(hi, (1, 150, 1))
    np.testing.assert_array_equal(om.rollout(x0, far)[0], om.rollout(x0, edge)[0])
    s = om.rollout(x0, edge)[0]
    jr = np.array([j["range"] for j in tb["joints"][1:]])
    assert np.all(s[0, -1, 7:23] < jr[:, 1] + 0.02) and np.all(s[0, -1, 7:23] > lo - 0.3)   # soft limits hold the joints
    # the same command WITH the hand-hand pairs: the fingers run into each other and into the palm long before the limits
    s2 = _leap()[0].rollout(x0, edge)[0]
    assert np.abs(s2[0, -1, 7:23] - s[0, -1, 7:23]).max() > 0.2 and np.all(np.isfinite(s2))


# ------------------------------------------------------------------ fr3_pick (reduced box geometry)
def test_fr3_static_equilibrium_resting_object_and_distance_sensors():
    """The servos hold the home pose against gravity (steady-state sag = gravity torque / kp), the cube rests on the table with a
    penetration of the order m g / k of the soft contact, the distance sensors report the geometric clearances."""
    from tests.fr3_cases import U_HOME, oracle_model
    from judo_b200.tasks.fr3_pick import QPOS_HOME

    om = oracle_model()
    x0 = np.concatenate([QPOS_HOME, np.zeros(15)])
    s, e = om.rollout(x0, np.tile(U_HOME, (1, 500, 1)))
    assert np.abs(s[0, -1, 16:]).max() < 1e-4                      # at rest
    assert np.abs(s[0, -1, 7:14] - QPOS_HOME[7:14]).max() < 0.01   # |sag| <= tau_g / kp ~ 30 Nm / 4500
    assert 0.0199 < s[0, -1, 2] <= 0.02 and np.abs(s[0, -1, :2] - [0.7, 0.0]).max() < 1e-6
    np.testing.assert_allclose(s[0, -1, 3:7], [1, 0, 0, 0], atol=1e-6)
    assert abs(e[0, -1, 4] - (s[0, -2, 2] - 0.02)) < 1e-9          # obj_table = signed clearance (sensors lag one step)
    assert np.all(e[0, :, :4] > 0.1)                               # fingers far from object and table
    # free flight: the cube dropped from 10 cm follows the semi-implicit Euler parabola until it lands
    x1 = x0.copy()
    x1[2] = 0.12
    s, e = om.rollout(x1, np.tile(U_HOME, (1, 20, 1)))
    k = np.arange(1, 21)
    np.testing.assert_allclose(s[0, :, 2], 0.12 - 9.81 * 0.004**2 * k * (k + 1) / 2, rtol=0, atol=1e-12)
    np.testing.assert_allclose(e[0, 1:, 4], s[0, :-1, 2] - 0.02, rtol=0, atol=1e-12)  # GJK distance of a box hovering over the table


def test_fr3_finger_equality_and_force_clamps():
    """Only finger_joint1 is actuated; the joint equality drags finger_joint2 along.  A far-away set point on joint 1 saturates the
    joint-level actuator force limit (87 N m): the first-step acceleration is bounded by it."""
    from tests.fr3_cases import U_HOME, oracle_model
    from judo_b200.tasks.fr3_pick import QPOS_HOME

    om = oracle_model()
    x0 = np.concatenate([QPOS_HOME, np.zeros(15)])
    u = np.tile(U_HOME, (1, 150, 1))
    u[..., 7] = 0.01
    s, _ = om.rollout(x0, u)
    assert abs(s[0, -1, 14] - 0.01) < 2e-3 and abs(s[0, -1, 14] - s[0, -1, 15]) < 1e-3
    u = np.tile(U_HOME, (1, 1, 1))
    u[..., 0] = 2.7
    f = om.forward(QPOS_HOME, np.zeros(15), u[0, 0])
    assert abs(f["qfrc_actuator"][6] - 87.0) < 1e-12               # kp * 2.7 = 12150 clamped to 87
    f2 = om.forward(QPOS_HOME, np.zeros(15), U_HOME)
    assert abs(f2["qfrc_actuator"][6]) < 1e-9
    # friction loss holds a joint against a small torque: zero set-point error on joint 7 -> stays put
    assert abs(f2["qacc"][12]) < 5e-3                              # (soft constraint: regularised, not exactly zero)


def test_fr3_grasp_lifts_the_cube_by_friction():
    from tests.fr3_cases import oracle_model, scenario

    om = oracle_model()
    x0, u = scenario("grasp", 2, 60)
    s, e = om.rollout(x0, u)
    assert s[:, -1, 2].min() > 0.03                                 # lifted well clear of the table
    assert (e[:, -1, 4] > 0.005).all()                             # obj_table distance agrees
    assert np.abs(s[:, -1, 14] - 0.0185).max() < 2e-3              # pads stopped by the 4 cm cube (pad face 1.5 mm inside the finger frame)


def test_box_box_signed_distance_matches_a_bounded_qp():
    """The routine behind the restated distance sensors (SAT depth when penetrating, GJK when separated) against an independent
    solve of  min |x - y|  over the two boxes (L-BFGS-B on the box-bounded QP), on random poses and on the degenerate poses the
    fr3 gripper produces all the time: parallel faces (a whole face of closest points) and axis-aligned boxes."""
    import ctypes

    from scipy.optimize import minimize
    from scipy.spatial.transform import Rotation as Rt

    from oracle.mjc import lib

    L = lib()
    L.mjc_box_box_distance.restype = ctypes.c_double
    dp = ctypes.POINTER(ctypes.c_double)
    P = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(dp)  # noqa: E731
    rng = np.random.default_rng(0)
    nsep = 0
    for it in range(120):
        p1, p2 = 0.1 * rng.normal(size=3), 0.1 * rng.normal(size=3) + np.array([0.15, 0, 0]) * rng.random()
        m1 = Rt.random(random_state=int(rng.integers(1 << 30))).as_matrix()
        m2 = Rt.random(random_state=int(rng.integers(1 << 30))).as_matrix()
        if it % 3 == 0:   # parallel faces: box 2 is box 1's frame turned about a shared axis
            m2 = m1 @ Rt.from_euler("z", rng.uniform(0, 6.28)).as_matrix()
        if it % 7 == 0:
            m1 = m2 = np.eye(3)
        s1, s2 = 0.01 + 0.05 * rng.random(3), 0.01 + 0.05 * rng.random(3)
        m1c, m2c = np.ascontiguousarray(m1), np.ascontiguousarray(m2)
        d = L.mjc_box_box_distance(P(p1), P(m1c), P(s1), P(p2), P(m2c), P(s2), ctypes.c_double(10.0))

        def f(z):
            r = (p1 + m1 @ z[:3]) - (p2 + m2 @ z[3:])
            return r @ r, np.concatenate([2 * m1.T @ r, -2 * m2.T @ r])

        best = min(minimize(f, np.concatenate([rng.uniform(-s1, s1), rng.uniform(-s2, s2)]), jac=True, method="L-BFGS-B",
                            bounds=list(zip(np.concatenate([-s1, -s2]), np.concatenate([s1, s2]))),
                            options=dict(ftol=1e-18, gtol=1e-14, maxiter=5000)).fun for _ in range(3))
        if d > 0:
            nsep += 1
            assert abs(d - np.sqrt(best)) < 1e-7, (it, d, np.sqrt(best))
        else:
            assert best < 1e-10, (it, d, best)   # penetrating (or touching): the QP distance is zero
    assert nsep > 60
