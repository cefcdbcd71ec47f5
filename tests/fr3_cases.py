"""Shared fr3_pick test scenarios (used by the emulator tests on the CPU and the parity tests on the GPU)."""
import numpy as np

from judo_b200.consts import load_table
from judo_b200.tasks.fr3_pick import Q_PREGRASP, QPOS_HOME, reduced_collision_model
from oracle.mjc import OracleModel

# arm configuration whose grasp site sits at (0.7, 0, 0.03) with the hand pointing down and the fingers along world y:
# the open gripper straddles the 4 cm cube at its home position (found by least squares on the compiled kinematics)
Q_GRASP = Q_PREGRASP
U_HOME = np.array([0, -0.7854, 0, -2.3562, 0, 1.5708, 0.7854, 0.04])


def oracle_model() -> OracleModel:
    tb = load_table("fr3_pick")
    geoms, pairs = reduced_collision_model(tb)
    return OracleModel(tb, pairs=pairs, geoms=geoms)


def scenario(name: str, N: int, H: int, seed: int = 0):
    """(x0, controls): 'home' = small noise around the home pose (object resting on the table); 'grasp' = the gripper closes on the
    cube and lifts it (pad-object contacts, friction, the finger equality); 'wild' = uniform controls over 120% of the control
    ranges (ctrl clamps, joint-level force clamps, joint limits); 'press' = the arm pushes the fingertips into the table."""
    rng = np.random.default_rng(seed)
    tb = load_table("fr3_pick")
    x0 = np.concatenate([QPOS_HOME, np.zeros(15)])
    if name == "home":
        u = U_HOME + 0.2 * rng.normal(size=(N, H, 8))
    elif name == "grasp":
        x0[7:14] = Q_GRASP
        u = np.tile(np.concatenate([Q_GRASP, [0.04]]), (N, H, 1)) + 0.02 * rng.normal(size=(N, H, 8))
        u[:, :, 7] = np.linspace(0.04, -0.02, H)[None, :] + 0.005 * rng.normal(size=(N, H))
        u[:, H // 2:, 1] -= 0.3
    elif name == "wild":
        lo = np.array([a["ctrlrange"][0] for a in tb["actuators"]])
        hi = np.array([a["ctrlrange"][1] for a in tb["actuators"]])
        u = lo + (hi - lo) * (1.2 * rng.random((N, H, 8)) - 0.1)
    elif name == "press":
        x0[7:14] = Q_GRASP
        x0[0] = 0.5  # object out of the way
        u = np.tile(np.concatenate([Q_GRASP, [0.02]]), (N, H, 1)) + 0.01 * rng.normal(size=(N, H, 8))
        u[:, :, 1] += 0.15  # shoulder forward/down: the pads hit the table
    else:
        raise ValueError(name)
    return x0, np.ascontiguousarray(u)
