"""LEAP cube rotation — mirror of judo/tasks/leap_cube.py:14-131, plus the collision model the kernel integrates.

Collision model: the full articulated dynamics of judo/models/xml/leap_cube.xml (free cube + 16 hinge
joints, implicitfast, elliptic cones, impratio 100, joint limits, friction loss, position servos with kv) and EVERY geom
pair that survives MuJoCo's static filters (same body, parent-child, the 18 <exclude>s of params_and_default.xml:76-101,
contype/conaffinity): the cube against the 67 hand boxes and the 4 fingertips (71 pairs) and the 1 621 hand-hand pairs
(finger-finger, finger-palm, links of one finger).  The one substitution left: the fingertip convex meshes (tip.obj /
thumb_tip.obj — not in git, assets.xml:8,12) are replaced by spheres at the mesh's nominal centre.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np

from judo_b200.tasks.base import Task, TaskConfig

QPOS_HOME = np.array([
    0.0, 0.03, 0.1, 1.0, 0.0, 0.0, 0.0,  # cube
    0.5, -0.75, 0.75, 0.25,  # index
    0.5, 0.0, 0.75, 0.25,  # middle
    0.5, 0.75, 0.75, 0.25,  # ring
    0.65, 0.9, 0.75, 0.6,  # thumb
])  # judo/tasks/leap_cube.py:16-24

# fingertip substitutes: (local centre in the distal body frame, radius).  The finger pads end at the trace site
# (leap_hand.xml:103,... `trace_*_tip` at y=-0.045 / thumb y=-0.055); the last collision box of the distal link ends at
# y=-0.020 / -0.031, so the pad spans the gap in between.
TIP_SPHERES = {"if_tip": ([0.0, -0.0325, 0.0145], 0.0125), "mf_tip": ([0.0, -0.0325, 0.0145], 0.0125),
               "rf_tip": ([0.0, -0.0325, 0.0145], 0.0125), "th_tip": ([0.0, -0.0430, -0.0150], 0.0125)}


def reduced_collision_model(table: dict, hand_hand: bool = True) -> tuple[list[dict], list[list[int]]]:
    """(geoms, pairs) of the model the kernel integrates: tips -> spheres; pairs = cube x every hand geom (in geom order), then the
    hand-hand pairs that survive MuJoCo's static filters (``table["pairs"]``, judo_b200/mjcf.py:candidate_pairs) in mj_collision's
    order: body pair by body pair, geoms of the first body outermost.  ``hand_hand=False`` gives the round-1 model (cube pairs only)."""
    geoms = []
    for g in table["geoms"]:
        g = dict(g)
        if g["type"] == "mesh":
            centre, radius = TIP_SPHERES[g["name"]]
            g.update(type="sphere", pos=list(centre), quat=[1.0, 0, 0, 0], size=[radius, 0.0, 0.0])
        geoms.append(g)
    cube = next(i for i, g in enumerate(geoms) if g["name"] == "cube")
    pairs = [[min(cube, i), max(cube, i)] for i in range(len(geoms)) if i != cube]
    if hand_hand:
        hh = [(a, b) for a, b in table["pairs"] if cube not in (a, b)]
        hh.sort(key=lambda p: (geoms[p[0]]["body"], geoms[p[1]]["body"], p[0], p[1]))
        pairs += [list(p) for p in hh]
    return geoms, pairs


@dataclass
class LeapCubeConfig(TaskConfig):
    """judo/tasks/leap_cube.py:29-34."""

    w_pos: float = 100.0
    w_rot: float = 0.1


class LeapCube(Task[LeapCubeConfig]):
    name = "leap_cube"
    config_t = LeapCubeConfig

    model_table = "leap_cube"
    home = QPOS_HOME
    default_goal_pos = (0.0, 0.03, 0.1)

    def __init__(self) -> None:
        super().__init__(self.model_table)
        self.goal_pos = np.array(self.default_goal_pos)
        self.goal_quat = np.array([1.0, 0.0, 0.0, 0.0])
        self.qpos_home = self.home
        self.reset_command = self.home[7:].copy()
        self.reset()

    def cost_params(self, system_metadata: dict[str, Any] | None = None) -> np.ndarray:
        """[w_pos, w_rot, goal_quat(4), goal_pos(3)]; goal_quat comes from the simulation (leap_cube.py:71-73)."""
        gq = (system_metadata or {}).get("goal_quat", np.array([1.0, 0.0, 0.0, 0.0]))
        return np.concatenate([[self.config.w_pos, self.config.w_rot], np.asarray(gq, dtype=np.float64), self.goal_pos])

    def reward(self, states: np.ndarray, sensors: np.ndarray, controls: np.ndarray,
               system_metadata: dict[str, Any] | None = None) -> np.ndarray:
        """-(w_pos/2 mean_t |p - goal|^2 + w_rot/2 mean_t |log(q* (x) q_goal)|^2)  (leap_cube.py:63-88)."""
        return self._gpu_reward(states, controls, system_metadata)

    def post_sim_step(self) -> None:
        """Reset on drop; new goal once within 0.4 rad of the current one (leap_cube.py:90-107)."""
        if self.data.qpos[2] < -0.3:
            self.reset()
        u, v = self.data.qpos[3:7] * np.array([1, -1, -1, -1]), self.goal_quat
        w = u[0] * v[0] - u[1:] @ v[1:]
        vec = u[0] * v[1:] + v[0] * u[1:] + np.cross(u[1:], v[1:])
        angle = 2 * np.arctan2(np.linalg.norm(vec), w)
        if angle > np.pi:
            angle -= 2 * np.pi
        if abs(angle) < 0.4:
            self._update_goal_quat()

    def _update_goal_quat(self) -> None:
        """Uniform random unit quaternion (leap_cube.py:109-123); same RNG draws as the reference."""
        uvw = np.random.rand(3)
        goal_quat = np.array([np.sqrt(1 - uvw[0]) * np.sin(2 * np.pi * uvw[1]), np.sqrt(1 - uvw[0]) * np.cos(2 * np.pi * uvw[1]),
                              np.sqrt(uvw[0]) * np.sin(2 * np.pi * uvw[2]), np.sqrt(uvw[0]) * np.cos(2 * np.pi * uvw[2])])
        self.data.mocap_quat[0] = goal_quat
        self.goal_quat = goal_quat

    def reset(self) -> None:
        self.data.qpos = self.qpos_home.copy()
        self.data.qvel = np.zeros(self.model.nv)
        self.data.ctrl = self.reset_command.copy()
        self._update_goal_quat()

    def get_sim_metadata(self) -> dict[str, Any]:
        return {"goal_quat": self.goal_quat}



QPOS_HOME_DOWN = np.array([
    -0.04, -0.035, -0.065, 1.0, 0.0, 0.0, 0.0,  # cube
    1.0, 0.0, 0.8, 0.8,  # index
    1.0, 0.0, 0.8, 0.8,  # middle
    1.0, 0.0, 0.8, 0.8,  # ring
    1.0, 1.0, 0.4, 0.9,  # thumb
])  # judo/tasks/leap_cube_down.py:13-21


@dataclass
class LeapCubeDownConfig(LeapCubeConfig):
    """judo/tasks/leap_cube_down.py:24-30."""

    w_rot: float = 0.05


class LeapCubeDown(LeapCube):
    """LEAP cube rotation with the palm facing down — mirror of judo/tasks/leap_cube_down.py:33-53.  Same kernel and reduced
    collision model, constants from leap_cube_palm_down.xml (hand frame un-tilted), its own home pose / goal position."""

    name = "leap_cube_down"
    config_t = LeapCubeDownConfig
    model_table = "leap_cube_down"
    home = QPOS_HOME_DOWN
    default_goal_pos = (-0.04, -0.035, -0.065)


# ------------------------------------------------------------------------------------------ constant table
LEAP_MAXG = 80
LEAP_MAXHH = 1664  # hand-hand candidate pairs (leap_cube: 1 621)


def leap_consts(table: dict) -> np.ndarray:
    """Flat all-double constant table == struct LeapModel in judo_b200/csrc/leap.cuh (same field order).

    Moving bodies are re-indexed: 0 = cube, 1 + 4f + d = link d of finger f (f: index, middle, ring, thumb).  Static
    bodies (leap_hand, palm) are folded into per-chain base poses and world-frame geoms."""
    from judo_b200 import mjcf

    o = table["opt"]
    bodies, joints, dofs = table["bodies"], table["joints"], table["dofs"]
    assert o["integrator"] == "implicitfast" and o["cone"] == "elliptic" and table["nq"] == 23 and table["nv"] == 22
    kin = mjcf.forward_kinematics(table, np.array(table["qpos0"], dtype=np.float64))
    cube_b = next(i for i, b in enumerate(bodies) if b["name"] == "cube")
    assert joints[bodies[cube_b]["jntadr"]]["type"] == "free" and np.allclose(bodies[cube_b]["ipos"], 0)
    links = [i for i, b in enumerate(bodies) if b["jntnum"] == 1 and joints[b["jntadr"]]["type"] == "hinge"]
    assert len(links) == 16
    moving = [cube_b] + links
    remap = {b: k for k, b in enumerate(moving)}
    # chains: link k (1-based) has parent k-1 unless it is a chain root (its table parent is static)
    for f in range(4):
        root = links[4 * f]
        assert bodies[root]["parent"] not in remap, "finger root must hang off a static body"
        for d in range(1, 4):
            assert bodies[links[4 * f + d]]["parent"] == links[4 * f + d - 1]
            assert joints[bodies[links[4 * f + d]]["jntadr"]]["dofadr"] == 6 + 4 * f + d
    v: list[float] = [o["timestep"], *o["gravity"], o["impratio"], o["tolerance"], o["ls_tolerance"], table["meaninertia"],
                      float(o["iterations"]), float(o["ls_iterations"])]
    base_pos, base_quat = [[0.0, 0.0, 0.0]], [[1.0, 0.0, 0.0, 0.0]]
    for f in range(4):
        p = bodies[links[4 * f]]["parent"]
        base_pos.append(kin["xpos"][p].tolist())
        base_quat.append(kin["xquat"][p].tolist())
    v += [x for r in base_pos for x in r] + [x for r in base_quat for x in r]
    mb = [bodies[b] for b in moving]
    v += [x for b in mb for x in b["pos"]] + [x for b in mb for x in b["quat"]] + [x for b in mb for x in b["ipos"]]
    v += [x for b in mb for x in mjcf.quat_to_mat(np.array(b["iquat"])).ravel()]
    v += [b["mass"] for b in mb] + [x for b in mb for x in b["inertia"]] + [b["invweight0"][0] for b in mb]
    jn = [joints[b["jntadr"]] for b in mb]
    v += [x for j in jn for x in (j["axis"] if j["type"] == "hinge" else [0, 0, 1])]
    v += [x for j in jn for x in (j["pos"] if j["type"] == "hinge" else [0, 0, 0])]
    v += table["qpos0"]
    cb = bodies[cube_b]
    R = mjcf.quat_to_mat(np.array(cb["iquat"]))
    v += (R @ np.diag(cb["inertia"]) @ R.T).ravel().tolist()
    v += [d["damping"] for d in dofs] + [d["frictionloss"] for d in dofs] + [d["invweight0"] for d in dofs]
    fr = [i for i, d in enumerate(dofs) if d["frictionloss"] > 0]
    hj = [joints[bodies[b]["jntadr"]] for b in links]
    assert all(j["solref_friction"] == hj[0]["solref_friction"] and j["solimp_friction"] == hj[0]["solimp_friction"] and
               j["solref_limit"] == hj[0]["solref_limit"] and j["solimp_limit"] == hj[0]["solimp_limit"] and j["margin"] == hj[0]["margin"]
               for j in hj)
    v += [float(len(fr))] + [float(x) for x in fr] + [0.0] * (22 - len(fr)) + hj[0]["solref_friction"] + hj[0]["solimp_friction"]
    v += [float(j["limited"]) for j in hj] + [j["range"][0] for j in hj] + [j["range"][1] for j in hj]
    v += [hj[0]["margin"]] + hj[0]["solref_limit"] + hj[0]["solimp_limit"]
    acts = table["actuators"]
    assert [a["dof"] for a in acts] == list(range(6, 22)) and all(a["gear"] == 1 and not a["forcelimited"] for a in acts)
    v += [a["kp"] for a in acts] + [a["kv"] for a in acts] + [float(a["ctrllimited"]) for a in acts]
    v += [a["ctrlrange"][0] for a in acts] + [a["ctrlrange"][1] for a in acts]
    import os

    geoms, pairs = reduced_collision_model(table, hand_hand=os.environ.get("B200MPC_LEAP_NO_HH", "0") != "1")  # (experiment knob: cube pairs only)
    cube_g = next(g for g in geoms if g["name"] == "cube")
    hand = [g for g in geoms if g["name"] != "cube"]
    assert len(hand) <= LEAP_MAXG and cube_g["type"] == "box" and np.allclose(cube_g["pos"], 0) and np.allclose(cube_g["quat"], [1, 0, 0, 0])
    for g in hand:
        assert g["type"] in ("box", "sphere") and g["condim"] == 3 and g["priority"] == cube_g["priority"] and g["margin"] == 0 and g["gap"] == 0
        assert g["solref"] == cube_g["solref"] and g["solimp"] == cube_g["solimp"] and g["solmix"] == cube_g["solmix"]
    pad = LEAP_MAXG - len(hand)
    gtype, gbody, gpos, gmat, gsize, grb, gmu = [], [], [], [], [], [], []
    for g in hand:
        gtype.append(6.0 if g["type"] == "box" else 2.0)
        Rg = mjcf.quat_to_mat(np.array(g["quat"]))
        pg = np.array(g["pos"])
        if g["body"] in remap:
            gbody.append(float(remap[g["body"]]))
        else:  # static body: fold into the world frame
            gbody.append(-1.0)
            Rb = mjcf.quat_to_mat(kin["xquat"][g["body"]])
            pg = kin["xpos"][g["body"]] + Rb @ pg
            Rg = Rb @ Rg
        gpos.append(pg.tolist())
        gmat.append(Rg.ravel().tolist())
        gsize.append(g["size"])
        grb.append(float(np.linalg.norm(g["size"])) if g["type"] == "box" else g["size"][0])
        gmu.append(max(MINMU_, cube_g["friction"][0], g["friction"][0]))
    v += [float(len(hand))] + gtype + [0.0] * pad + gbody + [0.0] * pad
    v += [x for r in gpos for x in r] + [0.0] * (3 * pad) + [x for r in gmat for x in r] + [0.0] * (9 * pad)
    v += [x for r in gsize for x in r] + [0.0] * (3 * pad) + grb + [0.0] * pad + gmu + [0.0] * pad
    v += cube_g["size"] + [float(np.linalg.norm(cube_g["size"]))] + cube_g["solref"] + cube_g["solimp"]
    sites = table["sites"]
    order = [s["obj"] for s in table["sensors"] if s["type"] == "framepos"]
    assert [s["type"] for s in table["sensors"]] == ["jointpos"] * 16 + ["framepos"] * 5
    v += [float(remap[sites[i]["body"]]) for i in order] + [x for i in order for x in sites[i]["pos"]]
    assert all(d >= 6 for d in fr)  # the cube's free joint has no friction loss (the kernel's per-dof lookup relies on it)
    v += [float(fr.index(d)) if d in fr else -1.0 for d in range(22)]
    # hand-hand pairs: raw sliding friction per hand geom (contact mixing = max of the two) and the pair list as g1 * 256 + g2 over the
    # hand-geom indices, in the order of `pairs` (body-pair major), packed as 16-bit codes
    v += [max(MINMU_, g["friction"][0]) for g in hand] + [0.0] * pad
    hidx = {gi: k for k, gi in enumerate(i for i, g in enumerate(geoms) if g["name"] != "cube")}
    cube_i = next(i for i, g in enumerate(geoms) if g["name"] == "cube")
    hh = [(a, b) for a, b in pairs if cube_i not in (a, b)]
    assert len(hh) <= LEAP_MAXHH
    for a, b in hh:  # same-body and static-static pairs never reach the kernel
        assert geoms[a]["body"] != geoms[b]["body"] and (geoms[a]["body"] in remap or geoms[b]["body"] in remap)
    v += [float(len(hh))]
    codes = np.zeros(LEAP_MAXHH, dtype=np.uint16)  # 16-bit codes, four to a double (raw bits: never touched as floating point)
    codes[:len(hh)] = [hidx[a] * 256 + hidx[b] for a, b in hh]
    return np.concatenate([np.array(v, dtype=np.float64), codes.view(np.float64)])


MINMU_ = 1e-5
