"""ctypes wrapper around the C oracle (oracle/mjc/mjc.c). ORACLE — test infrastructure, not product."""

from __future__ import annotations

import ctypes
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libmjc_oracle.so")
_MODELS = os.path.join(_HERE, "..", "judo_b200", "models")

_JNT = {"free": 0, "ball": 1, "slide": 2, "hinge": 3}
_GEOM = {"sphere": 2, "capsule": 3, "cylinder": 5, "box": 6}
_SENS = {"framepos": 0, "jointpos": 1, "framepos_body": 2, "framezaxis_body": 3, "distance": 4}
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mjc", "mjc.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B" if force else "-s"])
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.mjc_model_create.restype = ctypes.c_void_p
        _lib.mjc_model_create.argtypes = [_ip, ctypes.c_int, _dp, ctypes.c_int]
        _lib.mjc_model_free.argtypes = [ctypes.c_void_p]
        _lib.mjc_rollout.restype = ctypes.c_int
        _lib.mjc_rollout.argtypes = [ctypes.c_void_p, _dp, ctypes.c_int, _dp, ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_int]
        _lib.mjc_forward_debug.restype = ctypes.c_int
        _lib.mjc_forward_debug.argtypes = [ctypes.c_void_p] + [_dp] * 10 + [_ip] + [_dp] * 3 + [_ip]
    return _lib


def load_table(task: str) -> dict:
    with open(os.path.join(_MODELS, f"{task}.json")) as f:
        return json.load(f)


def serialize_model(model: dict, pairs: list | None = None, geoms: list | None = None) -> tuple[np.ndarray, np.ndarray]:
    """Flatten a compiled model table into the (int32, float64) blob mjc_model_create reads."""
    ib: list[int] = []
    db: list[float] = []
    geoms = model["geoms"] if geoms is None else geoms
    pairs = model["pairs"] if pairs is None else pairs
    usable = [i for i, g in enumerate(geoms) if g["type"] in _GEOM]
    remap = {g: k for k, g in enumerate(usable)}
    pairs = [[remap[a], remap[b]] for a, b in pairs if a in remap and b in remap]
    opt = model["opt"]
    ib += [model["nq"], model["nv"], model["nu"], model["nbody"], model["njnt"], len(usable), model["nsite"],
           len(model["sensors"]), model["nsensordata"], len(pairs), len(model.get("equalities", [])),
           {"Euler": 0, "implicitfast": 1}[opt["integrator"]], {"pyramidal": 0, "elliptic": 1}[opt["cone"]],
           int(opt["contact_disabled"]), opt["iterations"], opt["ls_iterations"]]
    db += [opt["timestep"], *opt["gravity"], opt["impratio"], opt["tolerance"], opt["ls_tolerance"], model["meaninertia"]]
    db += model["qpos0"]
    for b in model["bodies"]:
        ib += [b["parent"], b["jntadr"], b["jntnum"]]
        db += [*b["pos"], *b["quat"], *b["ipos"], *b["iquat"], b["mass"], *b["inertia"], *b["invweight0"]]
    for j in model["joints"]:
        ib += [_JNT[j["type"]], j["body"], j["qposadr"], j["dofadr"], int(j.get("limited", False)), int(j.get("actfrclimited", False))]
        db += [*j.get("pos", [0, 0, 0]), *j.get("axis", [0, 0, 1]), *j.get("range", [0, 0]), j.get("margin", 0.0),
               *j.get("solref_limit", [0.02, 1]), *j.get("solimp_limit", [0.9, 0.95, 0.001, 0.5, 2]), *j.get("actfrcrange", [0, 0])]
    for d in model["dofs"]:
        j = model["joints"][d["jnt"]]
        ib += [d["jnt"]]
        db += [d["damping"], d["frictionloss"], d["armature"], d["invweight0"],
               *j.get("solref_friction", [0.02, 1]), *j.get("solimp_friction", [0.9, 0.95, 0.001, 0.5, 2])]
    for gi in usable:
        g = geoms[gi]
        ib += [_GEOM[g["type"]], g["body"], g["condim"], g["priority"]]
        db += [*g["size"], *g["pos"], *g["quat"], *g["friction"], *g["solref"], *g["solimp"], g["margin"], g["gap"], g["solmix"]]
    for a, b in pairs:
        ib += [a, b]
    for s in model["sites"]:
        ib += [s["body"]]
        db += s["pos"]
    for a in model["actuators"]:
        ib += [a["dof"], int(a["ctrllimited"]), int(a["forcelimited"])]
        db += [a["gear"], a["kp"], a["kv"], *a["ctrlrange"], *a["forcerange"]]
    for s in model["sensors"]:
        ib += [_SENS[s["type"]], s["obj"], s.get("obj2", -1), s["adr"]]
        db += [s.get("cutoff", 0.0)]
    for e in model.get("equalities", []):
        ib += [e["joint1"], e["joint2"]]
        db += [*e["polycoef"], *e["solref"], *e["solimp"]]
    return np.array(ib, dtype=np.int32), np.array(db, dtype=np.float64)


class OracleModel:
    """A task model inside the C oracle; ``rollout`` restates MJRolloutBackend.rollout (mj_rollout_backend.py:45-88)."""

    def __init__(self, task_or_table: str | dict, pairs: list | None = None, geoms: list | None = None) -> None:
        self.table = load_table(task_or_table) if isinstance(task_or_table, str) else task_or_table
        ib, db = serialize_model(self.table, pairs, geoms)
        self._h = lib().mjc_model_create(ib.ctypes.data_as(_ip), len(ib), db.ctypes.data_as(_dp), len(db))
        if not self._h:
            raise RuntimeError("mjc_model_create failed (blob/limits mismatch)")
        self.nq, self.nv, self.nu = self.table["nq"], self.table["nv"], self.table["nu"]
        self.nsensordata = self.table["nsensordata"]

    def __del__(self) -> None:
        if getattr(self, "_h", None):
            lib().mjc_model_free(self._h)
            self._h = None

    def rollout(self, x0: np.ndarray, controls: np.ndarray, nthread: int = 0) -> tuple[np.ndarray, np.ndarray]:
        controls = np.ascontiguousarray(controls, dtype=np.float64)
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        N, H, nu = controls.shape
        assert nu == self.nu and x0.shape[-1] == self.nq + self.nv
        states = np.empty((N, H, self.nq + self.nv))
        sensors = np.empty((N, H, self.nsensordata))
        rc = lib().mjc_rollout(self._h, x0.ctypes.data_as(_dp), int(x0.ndim == 2), controls.ctypes.data_as(_dp), N, H,
                               states.ctypes.data_as(_dp), sensors.ctypes.data_as(_dp), nthread)
        if rc:
            raise RuntimeError("mjc_rollout failed")
        return states, sensors

    def forward(self, qpos: np.ndarray, qvel: np.ndarray, ctrl: np.ndarray) -> dict:
        nv = self.nv
        out = {k: np.zeros(nv) for k in ("qfrc_bias", "qfrc_passive", "qfrc_actuator", "qacc_smooth", "qacc", "qfrc_constraint")}
        M = np.zeros((nv, nv))
        ncon = ctypes.c_int(0)
        it = ctypes.c_int(0)
        dist = np.zeros(96)
        frame = np.zeros((96, 9))
        pos = np.zeros((96, 3))
        q = np.ascontiguousarray(qpos, dtype=np.float64).copy()
        v = np.ascontiguousarray(qvel, dtype=np.float64)
        u = np.ascontiguousarray(ctrl, dtype=np.float64)
        P = lambda a: a.ctypes.data_as(_dp)  # noqa: E731
        nefc = lib().mjc_forward_debug(self._h, P(q), P(v), P(u), P(M), P(out["qfrc_bias"]), P(out["qfrc_passive"]),
                                       P(out["qfrc_actuator"]), P(out["qacc_smooth"]), P(out["qacc"]), P(out["qfrc_constraint"]),
                                       ctypes.byref(ncon), P(dist), P(frame), P(pos), ctypes.byref(it))
        out.update(M=M, nefc=nefc, ncon=ncon.value, contact_dist=dist[:ncon.value], contact_frame=frame[:ncon.value],
                   contact_pos=pos[:ncon.value], solver_iter=it.value)
        return out
