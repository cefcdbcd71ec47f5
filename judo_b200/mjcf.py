"""MJCF-subset compiler: turns a task's MJCF file into the flat constant table the kernels bake in.

This is a HOST TOOL (numpy + xml.etree), not part of the per-step hot path.  It restates the parts
of MuJoCo 3.5.0's model compiler that the three BASELINE tasks exercise (SURVEY.md §8a rows D1-D3,
Appendix A): default classes, body tree, inertia-from-geom for box/capsule/cylinder/sphere,
explicit <inertial>, slide/hinge/free joints, position actuators, framepos/jointpos sensors,
contact excludes and the qpos0 constants ``dof_invweight0`` / ``body_invweight0`` / ``meaninertia``
that MuJoCo's constraint model needs.

The reference loads its models with ``MjSpec.from_file(...).compile()``
(/root/reference/judo/tasks/base.py:35-37); the MJCF files live in
/root/reference/judo/models/xml/.  Because neither MuJoCo nor the reference tree exist on the GPU
box, ``tools/gen_model_tables.py`` runs this compiler here and commits the result as
``judo_b200/models/<task>.json``; the runtime only reads those tables.

Mesh geoms cannot be resolved (the .obj files are not in git, assets.xml:8,12): visual meshes are
dropped (contype=conaffinity=0, density 0) and colliding meshes are recorded with ``type="mesh"``
so the reduced leap model can substitute a primitive and say so.
"""

from __future__ import annotations

import math
import os
import xml.etree.ElementTree as ET
from typing import Any

import numpy as np

MJ_MINVAL = 1e-15

GEOM_TYPES = {"plane": 0, "hfield": 1, "sphere": 2, "capsule": 3, "ellipsoid": 4, "cylinder": 5, "box": 6, "mesh": 7}
JNT_TYPES = {"free": 0, "ball": 1, "slide": 2, "hinge": 3}

_JOINT_DEFAULTS = dict(
    type="hinge", pos="0 0 0", axis="0 0 1", damping="0", frictionloss="0", armature="0", stiffness="0",
    margin="0", solreflimit="0.02 1", solimplimit="0.9 0.95 0.001 0.5 2",
    solreffriction="0.02 1", solimpfriction="0.9 0.95 0.001 0.5 2",
)
_GEOM_DEFAULTS = dict(
    type="sphere", pos="0 0 0", quat="1 0 0 0", friction="1 0.005 0.0001", condim="3", contype="1",
    conaffinity="1", solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2", margin="0", gap="0", priority="0",
    solmix="1", density="1000", group="0",
)


# ----------------------------------------------------------------------------- quaternion helpers
def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    w0, x0, y0, z0 = a
    w1, x1, y1, z1 = b
    return np.array([
        w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1,
        w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
        w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1,
        w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1,
    ])


def quat_to_mat(q: np.ndarray) -> np.ndarray:
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def mat_to_quat(R: np.ndarray) -> np.ndarray:
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return q / np.linalg.norm(q)


def quat_z_to_vec(v: np.ndarray) -> np.ndarray:
    """Quaternion rotating the z axis onto ``v`` (MuJoCo's fromto convention)."""
    v = v / np.linalg.norm(v)
    z = np.array([0.0, 0.0, 1.0])
    axis = np.cross(z, v)
    s = np.linalg.norm(axis)
    c = float(np.dot(z, v))
    if s < 1e-10:
        return np.array([1.0, 0, 0, 0]) if c > 0 else np.array([0.0, 1.0, 0, 0])
    axis = axis / s
    ang = math.atan2(s, c)
    return np.concatenate([[math.cos(ang / 2)], math.sin(ang / 2) * axis])


def _vec(s: str | None, n: int | None = None, fill: list[float] | None = None) -> np.ndarray:
    vals = [float(x) for x in s.split()] if s else []
    if n is not None and len(vals) < n:
        assert fill is not None
        vals = vals + list(fill[len(vals):n])
    return np.array(vals, dtype=np.float64)


# ----------------------------------------------------------------------------- XML loading
def _load_with_includes(path: str) -> ET.Element:
    root = ET.parse(path).getroot()
    base = os.path.dirname(path)

    def expand(elem: ET.Element) -> None:
        i = 0
        while i < len(elem):
            child = elem[i]
            if child.tag == "include":
                inc = _load_with_includes(os.path.join(base, child.attrib["file"]))
                elem.remove(child)
                for k, sub in enumerate(list(inc)):
                    elem.insert(i + k, sub)
                # do not advance: re-examine inserted nodes (already expanded recursively)
                i += len(list(inc))
            else:
                expand(child)
                i += 1

    expand(root)
    return root


class _Defaults:
    """Default-class tree: attribute dicts per (class, element tag), inherited from the parent class."""

    def __init__(self, root: ET.Element) -> None:
        self.table: dict[str, dict[str, dict[str, str]]] = {"main": {}}
        for d in root.findall("default"):
            self._walk(d, "main", top=True)

    def _walk(self, node: ET.Element, parent: str, top: bool = False) -> None:
        name = node.attrib.get("class", "main") if not top else node.attrib.get("class", "main")
        if name not in self.table:
            self.table[name] = {tag: dict(v) for tag, v in self.table[parent].items()}
        for child in node:
            if child.tag == "default":
                self._walk(child, name)
            else:
                cur = self.table[name].setdefault(child.tag, {})
                cur.update(child.attrib)
                # propagate to already-created descendants is unnecessary: children are created after

    def resolve(self, tag: str, elem: ET.Element, childclass: str | None) -> dict[str, str]:
        cls = elem.attrib.get("class", childclass or "main")
        out = dict(self.table.get(cls, self.table["main"]).get(tag, {}))
        out.update({k: v for k, v in elem.attrib.items() if k != "class"})
        return out


# ----------------------------------------------------------------------------- geom inertia
def _geom_mass_inertia(gtype: str, size: np.ndarray, density: float, mass: float | None) -> tuple[float, np.ndarray]:
    """Mass and principal inertia (geom frame) of a primitive — MuJoCo's closed-form solids."""
    if gtype == "sphere":
        r = size[0]
        vol = 4.0 / 3.0 * math.pi * r**3
        m = mass if mass is not None else density * vol
        inertia = np.full(3, 2.0 * m * r * r / 5.0)
    elif gtype == "box":
        a, b, c = size[:3]
        vol = 8 * a * b * c
        m = mass if mass is not None else density * vol
        inertia = m / 3.0 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])
    elif gtype == "cylinder":
        r, h = size[0], 2 * size[1]
        vol = math.pi * r * r * h
        m = mass if mass is not None else density * vol
        ixy = m * (3 * r * r + h * h) / 12.0
        inertia = np.array([ixy, ixy, m * r * r / 2.0])
    elif gtype == "capsule":
        r, h = size[0], 2 * size[1]
        vol = math.pi * r * r * h + 4.0 / 3.0 * math.pi * r**3
        m = mass if mass is not None else density * vol
        m_sph = m * 4 * r / (4 * r + 3 * h)
        m_cyl = m - m_sph
        ixy = m_cyl * (3 * r * r + h * h) / 12.0
        iz = m_cyl * r * r / 2.0
        i_sph = 2 * m_sph * r * r / 5.0
        ixy += i_sph + m_sph * h * (3 * r + 2 * h) / 8.0
        iz += i_sph
        inertia = np.array([ixy, ixy, iz])
    else:
        raise ValueError(f"inertia-from-geom not supported for geom type {gtype!r}")
    return m, inertia


# ----------------------------------------------------------------------------- compiler
def compile_mjcf(path: str) -> dict[str, Any]:
    """Compile an MJCF file into a flat, JSON-serialisable constant table."""
    root = _load_with_includes(path)
    defaults = _Defaults(root)

    comp = {}
    for c in root.findall("compiler"):
        comp.update(c.attrib)
    angle_scale = 1.0 if comp.get("angle", "degree") == "radian" else math.pi / 180.0
    autolimits = comp.get("autolimits", "true") == "true"

    opt = dict(timestep=0.002, integrator="Euler", cone="pyramidal", impratio=1.0, tolerance=1e-8,
               iterations=100, ls_iterations=50, ls_tolerance=0.01, gravity=[0.0, 0.0, -9.81],
               contact_disabled=False)
    for o in root.findall("option"):
        for k, v in o.attrib.items():
            if k in ("timestep", "impratio", "tolerance", "ls_tolerance"):
                opt[k] = float(v)
            elif k in ("iterations", "ls_iterations"):
                opt[k] = int(v)
            elif k == "gravity":
                opt[k] = [float(x) for x in v.split()]
            elif k in ("integrator", "cone"):
                opt[k] = v
        for f in o.findall("flag"):
            if f.attrib.get("contact") == "disable":
                opt["contact_disabled"] = True

    bodies: list[dict] = [dict(name="world", parent=0, pos=[0, 0, 0], quat=[1, 0, 0, 0], mocap=False,
                               ipos=[0, 0, 0], iquat=[1, 0, 0, 0], mass=0.0, inertia=[0, 0, 0],
                               jntadr=-1, jntnum=0, dofadr=-1, dofnum=0)]
    joints: list[dict] = []
    geoms: list[dict] = []
    sites: list[dict] = []
    nq = nv = 0
    qpos0: list[float] = []

    def add_body(elem: ET.Element, parent: int, childclass: str | None, moving: bool = False) -> None:
        nonlocal nq, nv
        moving = moving or any(j.tag in ("joint", "freejoint") for j in elem)
        childclass = elem.attrib.get("childclass", childclass)
        bid = len(bodies)
        pos = _vec(elem.attrib.get("pos", "0 0 0"))
        quat = _vec(elem.attrib.get("quat", "1 0 0 0"))
        quat = quat / np.linalg.norm(quat)
        body = dict(name=elem.attrib.get("name", f"body{bid}"), parent=parent, pos=pos.tolist(),
                    quat=quat.tolist(), mocap=elem.attrib.get("mocap", "false") == "true",
                    jntadr=-1, jntnum=0, dofadr=-1, dofnum=0)
        bodies.append(body)

        # joints
        for j in elem:
            if j.tag == "freejoint":
                a = dict(type="free")
                a.update(j.attrib)
            elif j.tag == "joint":
                a = dict(_JOINT_DEFAULTS)
                a.update(defaults.resolve("joint", j, childclass))
            else:
                continue
            jt = a.get("type", "hinge")
            if body["jntnum"] == 0:
                body["jntadr"] = len(joints)
                body["dofadr"] = nv
            nqj, nvj = {"free": (7, 6), "ball": (4, 3), "slide": (1, 1), "hinge": (1, 1)}[jt]
            rng = _vec(a.get("range", "0 0"))
            if jt == "hinge":
                rng = rng * angle_scale
            lim = a.get("limited", "auto")
            limited = (lim == "true") or (lim == "auto" and autolimits and "range" in a and rng[0] < rng[1])
            axis = _vec(a.get("axis", "0 0 1"))
            if jt in ("slide", "hinge"):
                axis = axis / np.linalg.norm(axis)
            joints.append(dict(
                name=a.get("name", f"joint{len(joints)}"), type=jt, body=bid, qposadr=nq, dofadr=nv,
                pos=_vec(a.get("pos", "0 0 0")).tolist(), axis=axis.tolist(),
                damping=float(a.get("damping", 0)), frictionloss=float(a.get("frictionloss", 0)),
                armature=float(a.get("armature", 0)), limited=bool(limited), range=rng.tolist(),
                margin=float(a.get("margin", 0)),
                actfrclimited=bool(a.get("actuatorfrclimited", "auto") == "true" or
                                   (a.get("actuatorfrclimited", "auto") == "auto" and autolimits and "actuatorfrcrange" in a)),
                actfrcrange=_vec(a.get("actuatorfrcrange", "0 0")).tolist(),
                solref_limit=_vec(a.get("solreflimit", "0.02 1")).tolist(),
                solimp_limit=_vec(a.get("solimplimit"), 5, [0.9, 0.95, 0.001, 0.5, 2]).tolist(),
                solref_friction=_vec(a.get("solreffriction", "0.02 1")).tolist(),
                solimp_friction=_vec(a.get("solimpfriction"), 5, [0.9, 0.95, 0.001, 0.5, 2]).tolist(),
            ))
            if jt == "free":
                qpos0.extend(pos.tolist() + quat.tolist())
            elif jt == "ball":
                qpos0.extend([1, 0, 0, 0])
            else:
                qpos0.append(float(a.get("ref", 0)))
            nq += nqj
            nv += nvj
            body["jntnum"] += 1
            body["dofnum"] += nvj

        # geoms
        body_geoms = []
        for g in elem.findall("geom"):
            a = dict(_GEOM_DEFAULTS)
            a.update(defaults.resolve("geom", g, childclass))
            gtype = a["type"]
            size = _vec(a.get("size", "0"))
            gpos = _vec(a.get("pos", "0 0 0"))
            gquat = _vec(a.get("quat", "1 0 0 0"))
            if "fromto" in a:
                ft = _vec(a["fromto"])
                p0, p1 = ft[:3], ft[3:]
                gpos = 0.5 * (p0 + p1)
                gquat = quat_z_to_vec(p1 - p0)
                size = np.array([size[0], 0.5 * np.linalg.norm(p1 - p0)])
            gquat = gquat / np.linalg.norm(gquat)
            size3 = np.zeros(3)
            size3[: len(size)] = size[:3]
            contype, conaff = int(a["contype"]), int(a["conaffinity"])
            fr = _vec(a["friction"], 3, [1, 0.005, 0.0001])
            rec = dict(
                name=a.get("name", f"geom{len(geoms)}"), type=gtype, body=bid, size=size3.tolist(),
                pos=gpos.tolist(), quat=gquat.tolist(), friction=fr.tolist(), condim=int(a["condim"]),
                contype=contype, conaffinity=conaff, solref=_vec(a["solref"]).tolist(),
                solimp=_vec(a["solimp"], 5, [0.9, 0.95, 0.001, 0.5, 2]).tolist(),
                margin=float(a["margin"]), gap=float(a["gap"]), priority=int(a["priority"]),
                solmix=float(a["solmix"]), mass=(float(a["mass"]) if "mass" in a else None),
                density=float(a["density"]), mesh=a.get("mesh"),
            )
            body_geoms.append(rec)
            if contype == 0 and conaff == 0:
                continue  # visual-only geom: never collides; it still counts for inertia-from-geom above
            geoms.append(rec)

        # sites
        for s in elem.findall("site"):
            a = defaults.resolve("site", s, childclass)
            sq = _vec(a.get("quat", "1 0 0 0"))
            sites.append(dict(name=a.get("name", f"site{len(sites)}"), body=bid,
                              pos=_vec(a.get("pos", "0 0 0")).tolist(), quat=(sq / np.linalg.norm(sq)).tolist()))

        # inertial
        inertial = elem.find("inertial")
        if inertial is not None:
            iq = _vec(inertial.attrib.get("quat", "1 0 0 0"))
            body.update(ipos=_vec(inertial.attrib["pos"]).tolist(), iquat=(iq / np.linalg.norm(iq)).tolist(),
                        mass=float(inertial.attrib["mass"]),
                        inertia=_vec(inertial.attrib["diaginertia"]).tolist())
        else:
            m_tot = 0.0
            com = np.zeros(3)
            parts = []
            for rec in body_geoms:
                if rec["type"] == "mesh":
                    if rec["density"] == 0 or rec["mass"] == 0:
                        continue
                    if not moving:
                        continue  # static body (welded to the world): its inertia never enters the dynamics
                    raise ValueError("mesh inertia requires the mesh asset (unavailable)")
                if rec["mass"] is not None and rec["mass"] == 0:
                    continue
                m, I = _geom_mass_inertia(rec["type"], np.array(rec["size"]), rec["density"], rec["mass"])
                if m <= 0:
                    continue
                parts.append((m, I, np.array(rec["pos"]), quat_to_mat(np.array(rec["quat"]))))
                m_tot += m
                com += m * np.array(rec["pos"])
            if m_tot > 0:
                com /= m_tot
                Ifull = np.zeros((3, 3))
                for m, I, p, R in parts:
                    d = p - com
                    Ifull += R @ np.diag(I) @ R.T + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
                w, V = np.linalg.eigh(Ifull)
                if np.allclose(Ifull, np.diag(np.diag(Ifull)), atol=1e-14 * max(1.0, np.abs(Ifull).max())):
                    w, V = np.diag(Ifull).copy(), np.eye(3)  # already principal: keep the body axes
                if np.linalg.det(V) < 0:
                    V[:, 2] = -V[:, 2]
                body.update(ipos=com.tolist(), iquat=mat_to_quat(V).tolist(), mass=m_tot, inertia=w.tolist())
            else:
                body.update(ipos=[0, 0, 0], iquat=[1, 0, 0, 0], mass=0.0, inertia=[0, 0, 0])

        for child in elem.findall("body"):
            add_body(child, bid, childclass, moving)

    wb = root.find("worldbody")
    assert wb is not None
    # geoms / sites directly in the worldbody belong to body 0
    if wb.findall("geom"):
        raise NotImplementedError("world geoms are not used by the BASELINE tasks")
    for b in wb.findall("body"):
        add_body(b, 0, None)

    # dof table
    dofs: list[dict] = []
    for ji, j in enumerate(joints):
        n = {"free": 6, "ball": 3, "slide": 1, "hinge": 1}[j["type"]]
        for k in range(n):
            dofs.append(dict(body=j["body"], jnt=ji, damping=j["damping"], frictionloss=j["frictionloss"],
                             armature=j["armature"]))
    # dof parent chain (dof_parentid): previous dof of the same body, else last dof of nearest jointed ancestor
    last_dof_of_body: dict[int, int] = {}
    for i, d in enumerate(dofs):
        b = d["body"]
        if b in last_dof_of_body:
            d["parent"] = last_dof_of_body[b]
        else:
            p = bodies[b]["parent"]
            while p != 0 and p not in last_dof_of_body:
                p = bodies[p]["parent"]
            d["parent"] = last_dof_of_body.get(p, -1)
        last_dof_of_body[b] = i

    # weld ids (bodies without joints are welded to their parent)
    for i, b in enumerate(bodies):
        if i == 0:
            b["weldid"] = 0
        else:
            b["weldid"] = i if b["jntnum"] > 0 else bodies[b["parent"]]["weldid"]

    # actuators (position servos only: gain kp, bias [0, -kp, -kv])
    acts = []
    for ablock in root.findall("actuator"):
        for a_el in ablock:
            if a_el.tag != "position":
                raise NotImplementedError(f"actuator type {a_el.tag}")
            a = defaults.resolve("position", a_el, None)
            jname = a["joint"]
            jid = next(i for i, j in enumerate(joints) if j["name"] == jname)
            kp = float(a.get("kp", 1))
            kv = float(a.get("kv", 0))
            cr = _vec(a.get("ctrlrange", "0 0"))
            fr = _vec(a.get("forcerange", "0 0"))
            if joints[jid]["type"] == "hinge":
                cr = cr * 1.0  # ctrlrange of a position servo on a hinge is an angle; leap uses radian
            inherit = float(a.get("inheritrange", 0))
            if inherit > 0:  # position servo: ctrlrange = joint range scaled about its midpoint (MuJoCo user_model: inheritrange)
                assert "ctrlrange" not in a, "inheritrange and ctrlrange are mutually exclusive"
                lo, hi = joints[jid]["range"]
                mid, rad = 0.5 * (lo + hi), 0.5 * (hi - lo) * inherit
                cr = np.array([mid - rad, mid + rad])
                a["ctrlrange"] = "inherited"
            cl = a.get("ctrllimited", "auto")
            fl = a.get("forcelimited", "auto")
            acts.append(dict(
                name=a.get("name", ""), joint=jid, dof=joints[jid]["dofadr"], gear=float(a.get("gear", "1").split()[0]),
                kp=kp, kv=kv,
                ctrllimited=bool(cl == "true" or (cl == "auto" and autolimits and "ctrlrange" in a)),
                ctrlrange=cr.tolist(),
                forcelimited=bool(fl == "true" or (fl == "auto" and autolimits and "forcerange" in a)),
                forcerange=fr.tolist(),
            ))

    # sensors
    sensors = []
    adr = 0
    for sblock in root.findall("sensor"):
        for s in sblock:
            if s.tag == "framepos" and s.attrib.get("objtype") == "site":
                sid = next(i for i, x in enumerate(sites) if x["name"] == s.attrib["objname"])
                sensors.append(dict(name=s.attrib.get("name", ""), type="framepos", obj=sid, adr=adr, dim=3))
                adr += 3
            elif s.tag in ("framepos", "framezaxis"):
                assert s.attrib.get("objtype") == "body"
                bid = next(i for i, x in enumerate(bodies) if x["name"] == s.attrib["objname"])
                sensors.append(dict(name=s.attrib.get("name", ""), type=s.tag + "_body", obj=bid, adr=adr, dim=3))
                adr += 3
            elif s.tag == "distance":
                b1 = next(i for i, x in enumerate(bodies) if x["name"] == s.attrib["body1"])
                b2 = next(i for i, x in enumerate(bodies) if x["name"] == s.attrib["body2"])
                sensors.append(dict(name=s.attrib.get("name", ""), type="distance", obj=b1, obj2=b2,
                                    cutoff=float(s.attrib.get("cutoff", 0)), adr=adr, dim=1))
                adr += 1
            elif s.tag == "jointpos":
                jid = next(i for i, j in enumerate(joints) if j["name"] == s.attrib["joint"])
                sensors.append(dict(name=s.attrib.get("name", ""), type="jointpos", obj=jid, adr=adr, dim=1))
                adr += 1
            else:
                raise NotImplementedError(f"sensor {s.tag}")

    equalities = []
    for eblock in root.findall("equality"):
        for e in eblock:
            if e.tag != "joint":
                raise NotImplementedError(f"equality {e.tag}")
            a = dict(solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2", polycoef="0 1 0 0 0")
            a.update(defaults.resolve("equality", e, None))
            j1 = next(i for i, j in enumerate(joints) if j["name"] == a["joint1"])
            j2 = next(i for i, j in enumerate(joints) if j["name"] == a["joint2"])
            equalities.append(dict(type="joint", joint1=j1, joint2=j2, polycoef=_vec(a["polycoef"], 5, [0, 1, 0, 0, 0]).tolist(),
                                   solref=_vec(a["solref"]).tolist(), solimp=_vec(a["solimp"], 5, [0.9, 0.95, 0.001, 0.5, 2]).tolist()))

    excludes = []
    for cblock in root.findall("contact"):
        for e in cblock.findall("exclude"):
            b1 = next(i for i, b in enumerate(bodies) if b["name"] == e.attrib["body1"])
            b2 = next(i for i, b in enumerate(bodies) if b["name"] == e.attrib["body2"])
            excludes.append([b1, b2])

    model: dict[str, Any] = dict(
        name=root.attrib.get("model", os.path.basename(path)), source=os.path.basename(path), opt=opt,
        nq=nq, nv=nv, nu=len(acts), nbody=len(bodies), njnt=len(joints), ngeom=len(geoms), nsite=len(sites),
        nsensordata=adr, qpos0=[float(x) for x in qpos0], bodies=bodies, joints=joints, dofs=dofs, geoms=geoms,
        sites=sites, actuators=acts, sensors=sensors, excludes=excludes, equalities=equalities,
    )
    model["pairs"] = candidate_pairs(model)
    _set_const(model)
    return model


def candidate_pairs(model: dict) -> list[list[int]]:
    """Geom pairs that survive MuJoCo's static collision filters (mj_collision's body-pair filter):
    same weld body, parent-child weld bodies (unless one is world-welded), <exclude> body pairs and the
    contype/conaffinity bit test.  Mesh geoms are listed too; the task decides what to substitute."""
    bodies = model["bodies"]
    excl = {(min(a, b), max(a, b)) for a, b in model["excludes"]}
    out = []
    for g1 in range(len(model["geoms"])):
        for g2 in range(g1 + 1, len(model["geoms"])):
            a, b = model["geoms"][g1], model["geoms"][g2]
            b1, b2 = a["body"], b["body"]
            if b1 == b2 or (min(b1, b2), max(b1, b2)) in excl:
                continue
            w1, w2 = bodies[b1]["weldid"], bodies[b2]["weldid"]
            if w1 == w2:
                continue
            wp1 = bodies[bodies[w1]["parent"]]["weldid"] if w1 else 0
            wp2 = bodies[bodies[w2]["parent"]]["weldid"] if w2 else 0
            if w1 != 0 and w2 != 0 and (w1 == wp2 or w2 == wp1):
                continue
            if not ((a["contype"] & b["conaffinity"]) or (b["contype"] & a["conaffinity"])):
                continue
            out.append([g1, g2])
    return out


# ----------------------------------------------------------------------------- qpos0 constants
def forward_kinematics(model: dict, qpos: np.ndarray) -> dict[str, np.ndarray]:
    """World poses of body frames, inertial frames, joint anchors/axes at ``qpos`` (numpy restatement)."""
    nb = model["nbody"]
    xpos = np.zeros((nb, 3))
    xquat = np.zeros((nb, 4))
    xquat[0] = [1, 0, 0, 0]
    xanchor = np.zeros((model["njnt"], 3))
    xaxis = np.zeros((model["njnt"], 3))
    for i in range(1, nb):
        b = model["bodies"][i]
        p = b["parent"]
        Rp = quat_to_mat(xquat[p])
        pos = xpos[p] + Rp @ np.array(b["pos"])
        quat = quat_mul(xquat[p], np.array(b["quat"]))
        for k in range(b["jntnum"]):
            j = model["joints"][b["jntadr"] + k]
            qa = j["qposadr"]
            if j["type"] == "free":
                pos = qpos[qa:qa + 3].copy()
                quat = qpos[qa + 3:qa + 7] / np.linalg.norm(qpos[qa + 3:qa + 7])
                xanchor[b["jntadr"] + k] = pos
                xaxis[b["jntadr"] + k] = [0, 0, 1]
                continue
            R = quat_to_mat(quat)
            axis_w = R @ np.array(j["axis"])
            anchor_w = pos + R @ np.array(j["pos"])
            q = qpos[qa] - model["qpos0"][qa]
            if j["type"] == "slide":
                pos = pos + axis_w * q
            elif j["type"] == "hinge":
                dq = np.concatenate([[math.cos(q / 2)], math.sin(q / 2) * np.array(j["axis"])])
                quat = quat_mul(quat, dq)
                R2 = quat_to_mat(quat)
                pos = anchor_w - R2 @ np.array(j["pos"])
            xanchor[b["jntadr"] + k] = anchor_w if j["type"] == "hinge" else pos + R @ np.array(j["pos"])
            xaxis[b["jntadr"] + k] = axis_w
        xpos[i] = pos
        xquat[i] = quat / np.linalg.norm(quat)
    xipos = np.zeros((nb, 3))
    ximat = np.zeros((nb, 3, 3))
    for i in range(nb):
        b = model["bodies"][i]
        R = quat_to_mat(xquat[i])
        xipos[i] = xpos[i] + R @ np.array(b["ipos"])
        ximat[i] = R @ quat_to_mat(np.array(b["iquat"]))
    return dict(xpos=xpos, xquat=xquat, xipos=xipos, ximat=ximat, xanchor=xanchor, xaxis=xaxis)


def body_jacobian(model: dict, kin: dict, body: int, point: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Translational / rotational Jacobians (3 x nv) of ``point`` attached to ``body``."""
    nv = model["nv"]
    jp = np.zeros((3, nv))
    jr = np.zeros((3, nv))
    b = body
    while b != 0:
        bd = model["bodies"][b]
        for k in range(bd["jntnum"]):
            ji = bd["jntadr"] + k
            j = model["joints"][ji]
            d = j["dofadr"]
            if j["type"] == "free":
                jp[:, d:d + 3] = np.eye(3)
                R = quat_to_mat(kin["xquat"][b])
                for a in range(3):
                    ax = R[:, a]
                    jr[:, d + 3 + a] = ax
                    jp[:, d + 3 + a] = np.cross(ax, point - kin["xpos"][b])
            elif j["type"] == "slide":
                jp[:, d] = kin["xaxis"][ji]
            elif j["type"] == "hinge":
                jr[:, d] = kin["xaxis"][ji]
                jp[:, d] = np.cross(kin["xaxis"][ji], point - kin["xanchor"][ji])
        b = bd["parent"]
    return jp, jr


def mass_matrix(model: dict, kin: dict) -> np.ndarray:
    """Joint-space inertia M(q) = sum_b Jp^T m Jp + Jr^T I_w Jr (+ armature)."""
    nv = model["nv"]
    M = np.zeros((nv, nv))
    for i in range(1, model["nbody"]):
        b = model["bodies"][i]
        if b["mass"] <= 0:
            continue
        jp, jr = body_jacobian(model, kin, i, kin["xipos"][i])
        Iw = kin["ximat"][i] @ np.diag(b["inertia"]) @ kin["ximat"][i].T
        M += b["mass"] * jp.T @ jp + jr.T @ Iw @ jr
    for d, dof in enumerate(model["dofs"]):
        M[d, d] += dof["armature"]
    return M


def _set_const(model: dict) -> None:
    """qpos0-dependent constants (MuJoCo's mj_setConst): invweight0 and meaninertia."""
    nv = model["nv"]
    q0 = np.array(model["qpos0"], dtype=np.float64)
    kin = forward_kinematics(model, q0)
    M = mass_matrix(model, kin)
    Minv = np.linalg.inv(M) if nv else np.zeros((0, 0))
    model["meaninertia"] = float(np.mean(np.diag(M))) if nv else 1.0
    body_inv = np.zeros((model["nbody"], 2))
    for i in range(1, model["nbody"]):
        jp, jr = body_jacobian(model, kin, i, kin["xipos"][i])
        A_t = jp @ Minv @ jp.T
        A_r = jr @ Minv @ jr.T
        body_inv[i] = [max(np.trace(A_t) / 3.0, 0.0), max(np.trace(A_r) / 3.0, 0.0)]
        if body_inv[i, 0] < MJ_MINVAL:
            body_inv[i] = 0.0
    for i, b in enumerate(model["bodies"]):
        b["invweight0"] = body_inv[i].tolist()
    dinv = np.diag(Minv).copy() if nv else np.zeros(0)
    for j in model["joints"]:
        d = j["dofadr"]
        if j["type"] == "free":
            dinv[d:d + 3] = dinv[d:d + 3].mean()
            dinv[d + 3:d + 6] = dinv[d + 3:d + 6].mean()
        elif j["type"] == "ball":
            dinv[d:d + 3] = dinv[d:d + 3].mean()
    for d, dof in enumerate(model["dofs"]):
        dof["invweight0"] = float(dinv[d])
