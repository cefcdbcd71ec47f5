"""FR3 pick-and-place — mirror of judo/tasks/fr3_pick.py:14-311, plus the REDUCED collision model the kernel integrates.

Reduced model: the full articulated dynamics of judo/models/xml/fr3_pick.xml (free object + 7 hinge arm
joints + 2 slide finger joints tied by a joint equality, implicitfast, pyramidal cones, impratio 10, armature, friction loss,
joint limits, position servos with kv, joint-level actuator force limits) with this collision geometry: the BOX geoms only —
table, object and the 2 x 5 fingertip pads.  The arm / hand / finger collision meshes (fr3_components/assets.xml:11-58,
`../../meshes/fr3/*.obj`) are not in git and cannot be downloaded here, so mesh geoms are dropped; pad-pad pairs are dropped
too (the pads of the two fingers are >= 3 mm apart over the whole joint range).  The five `distance` sensors are evaluated
over the same box geoms.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from typing import Any

import numpy as np

from judo_b200.tasks.base import Task, TaskConfig

QPOS_HOME = np.array([
    0.7, 0, 0.02, 1, 0, 0, 0,  # object
    0, -0.7854, 0.0, -2.3562, 0.0, 1.5708, 0.7854,  # arm
    0.04, 0.04,  # gripper, equality constrained
])  # judo/tasks/fr3_pick.py:16-22

# Arm configuration whose grasp site sits at (0.7, 0, 0.03) with the hand pointing down and the fingers along world y: the open
# gripper straddles the cube at its home position (least squares on the compiled kinematics; used by tests and bench.py's
# contact-rich workload, not by the task itself)
Q_PREGRASP = np.array([-0.07377, 0.87632, 0.08669, -1.54048, -0.10002, 2.41246, 0.8419])

MINMU = 1e-5
FR3_NCOST = 23


class Phase(Enum):
    """judo/tasks/fr3_pick.py:25-31."""

    LIFT = 0
    MOVE = 1
    PLACE = 2
    HOMING = 3


@dataclass
class LiftConfig:
    w_lift_close: float = 1.0
    w_lift_height: float = 10.0


@dataclass
class MoveConfig:
    w_move_goal: float = 1.0
    w_move_close: float = 10.0


@dataclass
class PlaceConfig:
    w_place_table: float = 1.0
    w_place_goal: float = 1.0


@dataclass
class GlobalConfig:
    w_upright: float = 0.25
    w_coll: float = 0.1
    w_qvel: float = 0.005
    w_open: float = 2.0


@dataclass
class FR3PickConfig(TaskConfig):
    """judo/tasks/fr3_pick.py:80-101."""

    lift_weights: LiftConfig = field(default_factory=LiftConfig)
    move_weights: MoveConfig = field(default_factory=MoveConfig)
    place_weights: PlaceConfig = field(default_factory=PlaceConfig)
    global_weights: GlobalConfig = field(default_factory=GlobalConfig)
    goal_pos: np.ndarray = field(default_factory=lambda: np.array([0.6, 0.4]))
    goal_radius: float = 0.05
    pick_height: float = 0.3


def reduced_collision_model(table: dict) -> tuple[list[dict], list[list[int]]]:
    """(geoms, pairs) of the reduced model.  geoms keeps the table's indexing (mesh geoms stay listed as type "mesh", which
    neither the oracle nor the kernel instantiates); pairs = table x object, table x pads, object x pads, in geom order."""
    geoms = table["geoms"]
    idx = {g["name"]: i for i, g in enumerate(geoms)}
    pads = [i for i, g in enumerate(geoms) if g["type"] == "box" and "finger_collision" in g["name"]]
    assert len(pads) == 10
    pairs = [[idx["table"], idx["box"]]] + [[idx["table"], p] for p in pads] + [[idx["box"], p] for p in pads]
    return geoms, pairs


def _mix_contact(g1: dict, g2: dict) -> tuple[float, list[float], list[float]]:
    """mj_contactParam for equal priorities: max friction, solmix-weighted solref / solimp."""
    assert g1["priority"] == g2["priority"] and g1["condim"] == g2["condim"] == 3 and g1["margin"] == g2["margin"] == 0
    assert g1["gap"] == g2["gap"] == 0 and g1["solmix"] > 0 and g2["solmix"] > 0
    mix = g1["solmix"] / (g1["solmix"] + g2["solmix"])
    assert g1["solref"][0] > 0 and g2["solref"][0] > 0
    solref = [mix * a + (1 - mix) * b for a, b in zip(g1["solref"], g2["solref"])]
    solimp = [mix * a + (1 - mix) * b for a, b in zip(g1["solimp"], g2["solimp"])]
    mu = max(MINMU, g1["friction"][0], g2["friction"][0])
    return mu, solref, solimp


def fr3_consts(table: dict) -> np.ndarray:
    """Flat all-double constant table == struct Fr3Model in judo_b200/csrc/fr3.cuh (same field order).

    Moving bodies are re-indexed: 0 = object, 1..7 = fr3_link1..7, 8 = hand, 9 = left_finger, 10 = right_finger."""
    from judo_b200 import mjcf

    o = table["opt"]
    bodies, joints, dofs = table["bodies"], table["joints"], table["dofs"]
    assert o["integrator"] == "implicitfast" and o["cone"] == "pyramidal" and table["nq"] == 16 and table["nv"] == 15
    bid = {b["name"]: i for i, b in enumerate(bodies)}
    order = ["object"] + [f"fr3_link{k}" for k in range(1, 8)] + ["hand", "left_finger", "right_finger"]
    mb = [bodies[bid[n]] for n in order]
    remap = {bid[n]: k for k, n in enumerate(order)}
    # structure the kernel hard-codes (checked, not assumed)
    assert joints[mb[0]["jntadr"]]["type"] == "free" and np.allclose(mb[0]["ipos"], 0) and np.allclose(mb[0]["iquat"], [1, 0, 0, 0])
    for k in range(1, 8):
        j = joints[mb[k]["jntadr"]]
        assert j["type"] == "hinge" and j["dofadr"] == 5 + k and np.allclose(j["pos"], 0) and table["qpos0"][j["qposadr"]] == 0
        assert (remap.get(mb[k]["parent"], -1) == k - 1) if k > 1 else (mb[k]["parent"] not in remap)
    assert mb[8]["jntnum"] == 0 and remap[mb[8]["parent"]] == 7
    for k, dof in ((9, 13), (10, 14)):
        j = joints[mb[k]["jntadr"]]
        assert j["type"] == "slide" and j["dofadr"] == dof and np.allclose(j["pos"], 0) and remap[mb[k]["parent"]] == 8
        assert table["qpos0"][j["qposadr"]] == 0
    kin = mjcf.forward_kinematics(table, np.array(table["qpos0"], dtype=np.float64))
    root_parent = mb[1]["parent"]
    v: list[float] = [o["timestep"], *o["gravity"], o["impratio"], o["tolerance"], o["ls_tolerance"], table["meaninertia"],
                      float(o["iterations"]), float(o["ls_iterations"])]
    v += kin["xpos"][root_parent].tolist() + kin["xquat"][root_parent].tolist()
    v += [x for b in mb for x in b["pos"]] + [x for b in mb for x in b["quat"]] + [x for b in mb for x in b["ipos"]]
    v += [x for b in mb for x in mjcf.quat_to_mat(np.array(b["iquat"])).ravel()]
    v += [b["mass"] for b in mb] + [x for b in mb for x in b["inertia"]] + [b["invweight0"][0] for b in mb]
    v += [x for b in mb for x in (joints[b["jntadr"]]["axis"] if b["jntnum"] == 1 and joints[b["jntadr"]]["type"] != "free" else [0, 0, 1])]
    v += mb[0]["inertia"]
    v += [d["damping"] for d in dofs] + [d["armature"] for d in dofs] + [d["frictionloss"] for d in dofs] + [d["invweight0"] for d in dofs]
    aj = [joints[mb[k]["jntadr"]] for k in (1, 2, 3, 4, 5, 6, 7, 9, 10)]
    assert all(d["frictionloss"] == 0 and d["damping"] == 0 and d["armature"] == 0 for d in dofs[:6]) and all(d["frictionloss"] > 0 for d in dofs[6:])
    same = lambda key: all(j[key] == aj[0][key] for j in aj)  # noqa: E731
    assert same("solref_friction") and same("solimp_friction") and same("solref_limit") and same("solimp_limit") and same("margin")
    assert all(j["limited"] for j in aj)
    v += aj[0]["solref_friction"] + aj[0]["solimp_friction"]
    v += [j["range"][0] for j in aj] + [j["range"][1] for j in aj] + [aj[0]["margin"]] + aj[0]["solref_limit"] + aj[0]["solimp_limit"]
    (eq,) = table["equalities"]
    assert eq["type"] == "joint" and joints[eq["joint1"]]["dofadr"] == 13 and joints[eq["joint2"]]["dofadr"] == 14
    assert np.allclose(eq["polycoef"], [0, 1, 0, 0, 0])
    v += eq["solref"] + eq["solimp"]
    acts = table["actuators"]
    assert [a["dof"] for a in acts] == [6, 7, 8, 9, 10, 11, 12, 13] and all(a["gear"] == 1 and not a["forcelimited"] and a["ctrllimited"] for a in acts)
    v += [a["kp"] for a in acts] + [a["kv"] for a in acts] + [a["ctrlrange"][0] for a in acts] + [a["ctrlrange"][1] for a in acts]
    v += [float(j["actfrclimited"]) for j in aj] + [j["actfrcrange"][0] for j in aj] + [j["actfrcrange"][1] for j in aj]
    geoms, pairs = reduced_collision_model(table)
    gi = {g["name"]: g for g in geoms}
    tab, obj = gi["table"], gi["box"]
    pads = [geoms[b] for a, b in pairs[1:11]]
    assert bodies[tab["body"]]["jntnum"] == 0 and np.allclose(obj["pos"], 0) and np.allclose(obj["quat"], [1, 0, 0, 0])
    Rt = mjcf.quat_to_mat(kin["xquat"][tab["body"]]) @ mjcf.quat_to_mat(np.array(tab["quat"]))
    pt = kin["xpos"][tab["body"]] + mjcf.quat_to_mat(kin["xquat"][tab["body"]]) @ np.array(tab["pos"])
    v += pt.tolist() + Rt.ravel().tolist() + tab["size"] + obj["size"]
    assert all(np.allclose(p["quat"], [1, 0, 0, 0]) and remap[p["body"]] in (9, 10) for p in pads)
    v += [float(remap[p["body"]]) for p in pads] + [x for p in pads for x in p["pos"]] + [x for p in pads for x in p["size"]]
    # contact classes: 0 table-object, 1 table-pad, 2 object-pad (all pads share their parameters)
    assert all(p["friction"] == pads[0]["friction"] and p["solref"] == pads[0]["solref"] and p["solimp"] == pads[0]["solimp"] for p in pads)
    cls = [_mix_contact(tab, obj), _mix_contact(tab, pads[0]), _mix_contact(obj, pads[0])]
    v += [c[0] for c in cls] + [x for c in cls for x in c[1]] + [x for c in cls for x in c[2]]
    (site,) = table["sites"]
    assert remap[site["body"]] == 8
    v += site["pos"]
    sens = table["sensors"]
    assert [s["type"] for s in sens] == ["distance"] * 5 + ["framezaxis_body", "framepos_body", "framepos"]
    assert [(remap.get(s["obj"]), remap.get(s["obj2"], "table")) for s in sens[:5]] == [(9, 0), (10, 0), (9, "table"), (10, "table"), (0, "table")]
    assert remap[sens[5]["obj"]] == 8 and remap[sens[6]["obj"]] == 0 and all(s["cutoff"] == sens[0]["cutoff"] for s in sens[:5])
    v += [sens[0]["cutoff"]]
    return np.array(v, dtype=np.float64)


class FR3Pick(Task[FR3PickConfig]):
    """judo/tasks/fr3_pick.py:104-311."""

    name = "fr3_pick"
    config_t = FR3PickConfig

    def __init__(self) -> None:
        super().__init__("fr3_pick")
        self.reset_command = np.array([0, 0, 0, -1.57079, 0, 1.57079, -0.7853, 0.0])
        self.obj_pos_adr = self.get_joint_position_start_index("object_joint")
        self.obj_pos_slice = slice(self.obj_pos_adr, self.obj_pos_adr + 3)
        arm_pos_adr = self.get_joint_position_start_index("fr3_joint1")
        self.arm_pos_slice = slice(arm_pos_adr, arm_pos_adr + 9)
        self.left_finger_table_adr = self.get_sensor_start_index("left_finger_table")
        self.right_finger_table_adr = self.get_sensor_start_index("right_finger_table")
        self.grasp_site_adr = self.get_sensor_start_index("trace_grasp_site")
        self.obj_table_adr = self.get_sensor_start_index("obj_table")
        self.ee_z_adr = self.get_sensor_start_index("ee_z")
        self.phase = Phase.LIFT
        self.reset()

    def in_goal_xy(self, curr_state: np.ndarray) -> bool:
        """fr3_pick.py:146-160."""
        obj_pos = curr_state[self.obj_pos_adr:self.obj_pos_adr + 2]
        return bool(np.linalg.norm(obj_pos - self.config.goal_pos) <= self.config.goal_radius)

    def pre_rollout(self, curr_state: np.ndarray) -> None:
        """Phase machine on the current state (fr3_pick.py:191-223; the reference's mj_forward there feeds only commented-out code)."""
        phase = Phase.LIFT
        obj_in_air = curr_state[self.obj_pos_adr + 2] > 0.02 + 1e-3
        if obj_in_air:
            phase = Phase.MOVE
        in_goal_xy = self.in_goal_xy(curr_state)
        if in_goal_xy and obj_in_air:
            phase = Phase.PLACE
        if in_goal_xy and curr_state[self.obj_pos_adr + 2] <= 0.02 + 1e-3:
            phase = Phase.HOMING
        self.phase = phase

    def cost_params(self, system_metadata: dict[str, Any] | None = None) -> np.ndarray:
        """[phase, w_lift_close, w_lift_height, w_move_goal, w_move_close, w_place_table, w_place_goal, w_upright, w_coll, w_qvel,
        w_open, goal_x, goal_y, pick_height, q_home(9)]."""
        c = self.config
        return np.array([float(self.phase.value), c.lift_weights.w_lift_close, c.lift_weights.w_lift_height, c.move_weights.w_move_goal,
                         c.move_weights.w_move_close, c.place_weights.w_place_table, c.place_weights.w_place_goal,
                         c.global_weights.w_upright, c.global_weights.w_coll, c.global_weights.w_qvel, c.global_weights.w_open,
                         *np.asarray(c.goal_pos, dtype=np.float64), c.pick_height, *QPOS_HOME[self.arm_pos_slice]], dtype=np.float64)

    def reward(self, states: np.ndarray, sensors: np.ndarray, controls: np.ndarray,
               system_metadata: dict[str, Any] | None = None) -> np.ndarray:
        """Phase-switched reward (fr3_pick.py:225-311), evaluated on the GPU from states + sensors."""
        if self.engine is None:
            from judo_b200.engine import Engine
            self.engine = Engine(self.name, max(1, len(states)))
        return self.engine.reward(states, controls, self.cost_params(system_metadata), sensors=sensors)

    def reset(self) -> None:
        """fr3_pick.py:313-318."""
        self.data.qpos = QPOS_HOME.copy()
        self.data.qvel = np.zeros(self.model.nv)
        self.data.ctrl = self.reset_command.copy()
