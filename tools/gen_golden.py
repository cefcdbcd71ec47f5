"""Generate tests/golden/*.npz by running the UNMODIFIED reference Python (authoring container only).

What is pinned (reference code executed, outputs stored):
  optimizers.npz   judo.optimizers.{MPPI,CrossEntropyMethod,PredictiveSampling}: sample_control_knots under
                   np.random.seed, update_nominal_knots, CEM sigma state, with the per-task overrides.
  spline.npz       judo.controller.controller.make_spline (scipy interp1d) for zero/linear/cubic.
  rewards.npz      Cartpole.reward / CylinderPush.reward / LeapCube.reward (+ math_utils.quat_diff_so3).
  plan_<cfg>.npz   three consecutive reference Controller.update_action() calls where ONLY the rollout backend is
                   replaced (by the C oracle, since MuJoCo is absent): candidate knots, rollout controls, rewards,
                   nominal knots, spline times and traces.  Pins every line of the plan step except the physics.

Run:  python tools/gen_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
import ref_shim  # noqa: E402

ref_shim.install()
from unittest import mock  # noqa: E402

for name in ["omegaconf", "hydra", "hydra.utils", "trimesh", "trimesh.visual", "trimesh.visual.material", "PIL",
             "onnxruntime", "mujoco_extensions", "mujoco_extensions.policy_rollout", "dora_utils", "viser.theme"]:
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:  # noqa: BLE001
            m = mock.MagicMock(name=name)
            m.__path__ = []
            sys.modules[name] = m

import judo.controller.controller as ref_ctrl  # noqa: E402
from judo.controller import Controller, ControllerConfig  # noqa: E402
from judo.optimizers import get_registered_optimizers  # noqa: E402
from judo.tasks.cartpole import Cartpole, CartpoleConfig  # noqa: E402
from judo.tasks.cylinder_push import CylinderPush, CylinderPushConfig  # noqa: E402
from judo.tasks.leap_cube import QPOS_HOME, LeapCube, LeapCubeConfig  # noqa: E402
from judo.utils.math_utils import quat_diff_so3  # noqa: E402
from judo.utils.rollout_backend import RolloutBackend  # noqa: E402

from oracle.mjc import OracleModel, load_table  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)


# ------------------------------------------------------------------ optimizers
def gen_optimizers() -> None:
    out = {}
    reg = get_registered_optimizers()
    cases = [("mppi", "cartpole", 1, 64), ("mppi", "leap_cube", 16, 48), ("ps", "cartpole", 1, 32), ("ps", "leap_cube", 16, 24),
             ("cem", "cylinder_push", 2, 64), ("cem", "leap_cube", 16, 40), ("mppi", "default", 3, 16), ("cem", "default", 3, 16)]
    for ci, (opt_name, task, nu, N) in enumerate(cases):
        cls, cfg_cls = reg[opt_name]
        cfg = cfg_cls()
        if task != "default":
            cfg.set_override(task)
        cfg.num_rollouts = N
        opt = cls(cfg, nu)
        rng = np.random.RandomState(100 + ci)
        nominal = rng.randn(cfg.num_nodes, nu)
        key = f"c{ci}_"
        out[key + "meta"] = np.array([opt_name, task, str(nu), str(N), str(cfg.num_nodes), str(int(cfg.use_noise_ramp)), repr(cfg.noise_ramp)])
        for f in ("sigma", "temperature", "sigma_min", "sigma_max", "num_elites"):
            if hasattr(cfg, f):
                out[key + f] = np.array(getattr(cfg, f))
        np.random.seed(7 + ci)
        for it in range(2):  # two consecutive iterations: exercises CEM's sigma mutation
            if opt_name == "cem":
                out[key + f"sigma_in{it}"] = opt.sigma.copy()
            knots = opt.sample_control_knots(nominal)
            rewards = -np.abs(rng.randn(N)) * 3
            if it == 1:
                rewards[3] = rewards[5]  # a tie
            new_nominal = opt.update_nominal_knots(knots, rewards)
            out[key + f"nominal_in{it}"] = nominal.copy()
            out[key + f"knots{it}"] = knots
            out[key + f"rewards{it}"] = rewards
            out[key + f"nominal_out{it}"] = new_nominal
            if opt_name == "cem":
                out[key + f"sigma_out{it}"] = opt.sigma.copy()
            nominal = new_nominal
    out["ncases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "optimizers.npz"), **out)


# ------------------------------------------------------------------ spline
def gen_spline() -> None:
    out = {}
    rng = np.random.RandomState(3)
    ci = 0
    for kind in ("zero", "linear", "cubic"):
        for K in (4, 7):
            for (t0, horizon, dt, H) in ((0.0, 1.0, 0.02, 50), (3.7, 2.56, 0.04, 64), (0.12, 0.4, 0.01, 40), (1.0, 1.0, 0.25, 8)):
                times = t0 + np.linspace(0, horizon, K, endpoint=True)
                knots = rng.randn(5, K, 3)
                query = t0 + dt * np.arange(H)
                query_shift = t0 + 0.37 * horizon + np.linspace(0, horizon, K, endpoint=True)  # time-shift query (controller.py:220-221)
                sp = ref_ctrl.make_spline(times, knots, kind)
                out[f"s{ci}_kind"] = np.array(kind)
                out[f"s{ci}_times"] = times
                out[f"s{ci}_knots"] = knots
                out[f"s{ci}_query"] = query
                out[f"s{ci}_out"] = sp(query)
                out[f"s{ci}_query_shift"] = query_shift
                out[f"s{ci}_out_shift"] = sp(query_shift)
                ci += 1
    out["ncases"] = np.array(ci)
    np.savez_compressed(os.path.join(OUT, "spline.npz"), **out)


# ------------------------------------------------------------------ rewards
def gen_rewards() -> None:
    rng = np.random.RandomState(11)
    out = {}
    s = rng.randn(6, 9, 4) * 2
    u = rng.randn(6, 9, 1)
    out["cartpole_states"], out["cartpole_controls"] = s, u
    out["cartpole_rewards"] = Cartpole.reward(types.SimpleNamespace(config=CartpoleConfig()), s, None, u)
    s = rng.randn(6, 9, 8)
    u = rng.randn(6, 9, 2)
    out["cylinder_push_states"], out["cylinder_push_controls"] = s, u
    out["cylinder_push_rewards"] = CylinderPush.reward(types.SimpleNamespace(config=CylinderPushConfig()), s, None, u)
    cfg = CylinderPushConfig()
    cfg.goal_pos = np.array([0.3, -0.2])
    out["cylinder_push_goal"] = cfg.goal_pos
    out["cylinder_push_rewards_goal"] = CylinderPush.reward(types.SimpleNamespace(config=cfg), s, None, u)
    s = rng.randn(6, 9, 45)
    s[..., 3:7] /= np.linalg.norm(s[..., 3:7], axis=-1, keepdims=True)
    s[0, 0, 3:7] = [1, 0, 0, 0]          # zero rotation -> safe_normalize_axis branch
    s[1, 2, 3:7] = [-1, 0, 0, 0]         # angle 2*pi -> wrap branch
    s[2, 1, 3:7] = [0, 1, 0, 0]          # exactly pi
    gq = rng.randn(4)
    gq /= np.linalg.norm(gq)
    fake = types.SimpleNamespace(config=LeapCubeConfig(), goal_pos=np.array([0.0, 0.03, 0.1]))
    out["leap_states"], out["leap_goal_quat"] = s, gq
    out["leap_rewards_default_goal"] = LeapCube.reward(fake, s, None, None, {})
    out["leap_rewards"] = LeapCube.reward(fake, s, None, None, {"goal_quat": gq})
    qa = rng.randn(50, 4)
    qa /= np.linalg.norm(qa, axis=-1, keepdims=True)
    out["quat_u"], out["quat_v"] = qa, gq
    out["quat_diff_so3"] = quat_diff_so3(qa, gq)
    np.savez_compressed(os.path.join(OUT, "rewards.npz"), **out)


def _fr3_fake(table: dict, cfg=None):  # noqa: ANN001, ANN202
    """An instance of the reference FR3Pick whose MuJoCo objects are namespaces; indices as its __init__ computes them
    (fr3_pick.py:111-144) from the model's joint / sensor addresses."""
    from judo.tasks.fr3_pick import QPOS_HOME as FR3_HOME, FR3Pick, FR3PickConfig, Phase

    jq = {j["name"]: j["qposadr"] for j in table["joints"]}
    jd = {j["name"]: j["dofadr"] for j in table["joints"]}
    sa = {s_["name"]: s_["adr"] for s_ in table["sensors"]}
    ova = table["nq"] + jd["object_joint"]
    extra = dict(
        reset_command=np.array([0, 0, 0, -1.57079, 0, 1.57079, -0.7853, 0.0]), obj_pos_adr=jq["object_joint"],
        obj_pos_slice=slice(jq["object_joint"], jq["object_joint"] + 3), obj_vel_slice=slice(ova, ova + 3),
        obj_angvel_slice=slice(ova + 3, ova + 6), arm_pos_slice=slice(jq["fr3_joint1"], jq["fr3_joint1"] + 9),
        left_finger_obj_adr=sa["left_finger_obj"], right_finger_obj_adr=sa["right_finger_obj"],
        left_finger_table_adr=sa["left_finger_table"], right_finger_table_adr=sa["right_finger_table"],
        grasp_site_adr=sa["trace_grasp_site"], obj_table_adr=sa["obj_table"], ee_z_adr=sa["ee_z"],
        ee_z_slice=slice(sa["ee_z"], sa["ee_z"] + 3), phase=Phase.LIFT,
        _data=types.SimpleNamespace(qpos=np.zeros(table["nq"]), qvel=np.zeros(table["nv"])),
    )
    t = fake_task(FR3Pick, FR3PickConfig, table, extra)
    if cfg is not None:
        t.config = cfg
    t.data.qpos = FR3_HOME.copy()
    t.data.ctrl = extra["reset_command"].copy()
    return t, Phase


def gen_rewards_fr3() -> None:
    """FR3Pick.reward for the four phases and FR3Pick.pre_rollout's phase machine (fr3_pick.py:191-311), reference code executed."""
    table = load_table("fr3_pick")
    task, Phase = _fr3_fake(table)
    rng = np.random.RandomState(21)
    s = rng.randn(5, 7, 31) * 0.3
    e = rng.randn(5, 7, 14) * 0.2
    e[0, 1, 2] = 0.0      # exactly touching counts as touching (<= 0)
    e[1, :, 3] = -0.01
    out = dict(fr3_states=s, fr3_sensors=e)
    for ph in Phase:
        task.phase = ph
        out[f"fr3_rewards_phase{ph.value}"] = task.reward(s, e, None)
    task.config.goal_pos = np.array([0.5, -0.3])
    task.config.pick_height = 0.2
    task.config.global_weights.w_coll = 0.7
    task.phase = Phase.MOVE
    out["fr3_rewards_custom"] = task.reward(s, e, None)
    out["fr3_custom"] = np.array([0.5, -0.3, 0.2, 0.7])
    # phase machine
    task, Phase = _fr3_fake(table)
    xs, phases = [], []
    for x, y, z in ((0.7, 0.0, 0.02), (0.7, 0.0, 0.0211), (0.62, 0.41, 0.1), (0.6, 0.4, 0.02), (0.6, 0.4, 0.021), (0.6, 0.46, 0.5), (0.0, 0.0, 0.0)):
        st = np.zeros(31)
        st[:3] = [x, y, z]
        st[3] = 1
        task.pre_rollout(st)
        xs.append(st)
        phases.append(task.phase.value)
    out["fr3_phase_states"], out["fr3_phases"] = np.array(xs), np.array(phases)
    np.savez_compressed(os.path.join(OUT, "rewards_fr3.npz"), **out)
    print("fr3 rewards golden: phases", phases)


# ------------------------------------------------------------------ full plan step through the reference Controller
class OracleBackend(RolloutBackend):
    """Stands in for MJRolloutBackend (mujoco absent): same contract, physics from the C oracle."""

    def __init__(self, om: OracleModel, num_threads: int) -> None:
        self.om = om
        self.num_threads = num_threads

    def rollout(self, x0, controls, last_policy_output=None):  # noqa: ANN001
        states, sensors = self.om.rollout(np.asarray(x0), np.asarray(controls))
        return states, sensors, None

    def update(self, num_threads: int) -> None:
        self.num_threads = num_threads


def fake_task(task_cls, cfg_cls, table: dict, extra: dict | None = None):  # noqa: ANN001
    """An instance of the reference Task subclass whose MuJoCo model/data are plain namespaces."""
    t = object.__new__(task_cls)
    t.config = cfg_cls()
    acts = table["actuators"]
    sens_adr = np.array([s["adr"] for s in table["sensors"]])
    t.model = types.SimpleNamespace(
        nq=table["nq"], nv=table["nv"], nu=table["nu"], nsensordata=table["nsensordata"], nsensor=len(table["sensors"]),
        sensor_adr=sens_adr, opt=types.SimpleNamespace(timestep=table["opt"]["timestep"]),
        actuator_ctrlrange=np.array([a["ctrlrange"] for a in acts], dtype=np.float64),
        actuator_ctrllimited=np.array([a["ctrllimited"] for a in acts]),
    )
    t.data = types.SimpleNamespace(qpos=np.zeros(table["nq"]), qvel=np.zeros(table["nv"]), ctrl=np.zeros(table["nu"]), time=0.0,
                                   mocap_quat=np.zeros((1, 4)))
    for k, v in (extra or {}).items():
        setattr(t, k, v)
    return t


def gen_plan(tag: str, task_name: str, opt_name: str, N: int, horizon: float, seed: int, leap_pairs=None) -> None:  # noqa: ANN001
    table = load_table(task_name)
    if task_name == "leap_cube":
        from judo_b200.tasks.leap_cube import reduced_collision_model  # the SAME reduced geometry the product uses
        geoms, pairs = reduced_collision_model(table)
        om = OracleModel(table, pairs=pairs, geoms=geoms)
    elif task_name == "fr3_pick":
        from judo_b200.tasks.fr3_pick import reduced_collision_model as fr3_reduced
        geoms, pairs = fr3_reduced(table)
        om = OracleModel(table, pairs=pairs, geoms=geoms)
    else:
        om = OracleModel(table)
    np.random.seed(seed)
    if task_name == "cartpole":
        task = fake_task(Cartpole, CartpoleConfig, table)
    elif task_name == "cylinder_push":
        task = fake_task(CylinderPush, CylinderPushConfig, table)
    elif task_name == "fr3_pick":
        task, _ = _fr3_fake(table)
    else:
        rc = QPOS_HOME[7:].copy()
        task = fake_task(LeapCube, LeapCubeConfig, table, dict(goal_pos=np.array([0.0, 0.03, 0.1]), goal_quat=np.array([1.0, 0, 0, 0]),
                                                                qpos_home=QPOS_HOME, reset_command=rc))
    cls, cfg_cls = get_registered_optimizers()[opt_name]
    ocfg = cfg_cls()
    ocfg.set_override(task_name)
    ocfg.num_rollouts = N
    opt = cls(ocfg, table["nu"])
    ccfg = ControllerConfig()
    ccfg.set_override(task_name)
    ccfg.horizon = horizon
    trace_ids = [i for i, s in enumerate(table["sensors"]) if s["type"] in ("framepos", "framepos_body") and "trace" in s["name"]]
    with mock.patch.object(ref_ctrl, "MJRolloutBackend", lambda model, num_threads: OracleBackend(om, num_threads)), \
            mock.patch.object(ref_ctrl, "get_trace_sensors", lambda model: trace_ids):
        ctrl = Controller(ccfg, task, opt)
    out = dict(meta=np.array([task_name, opt_name, str(N), repr(horizon), str(seed), ccfg.spline_order, str(ccfg.max_num_traces)]))
    out["x_init"] = np.concatenate([task.data.qpos, task.data.qvel])
    if task_name == "leap_cube":
        ctrl.system_metadata = {"goal_quat": task.goal_quat.copy()}
        out["goal_quat"] = task.goal_quat.copy()
    ctrl.current_state = out["x_init"].copy()
    t = 0.0
    for step in range(3):
        ctrl.time = t
        out[f"p{step}_time"] = np.array(t)
        out[f"p{step}_x0"] = ctrl.current_state.copy()
        out[f"p{step}_times_in"] = ctrl.times.copy()
        out[f"p{step}_nominal_in"] = ctrl.nominal_knots.copy()
        if opt_name == "cem":
            out[f"p{step}_sigma_in"] = opt.sigma.copy()
        ctrl.update_action()
        out[f"p{step}_candidate_knots"] = ctrl.candidate_knots.copy()
        out[f"p{step}_rollout_controls"] = np.array(ctrl.rollout_controls)
        out[f"p{step}_states"] = ctrl.states.copy()
        out[f"p{step}_sensors"] = ctrl.sensors.copy()
        out[f"p{step}_rewards"] = ctrl.rewards.copy()
        out[f"p{step}_nominal_out"] = ctrl.nominal_knots.copy()
        out[f"p{step}_times_out"] = ctrl.times.copy()
        out[f"p{step}_traces"] = ctrl.traces.copy()
        if opt_name == "cem":
            out[f"p{step}_sigma_out"] = opt.sigma.copy()
        if task_name == "fr3_pick":
            out[f"p{step}_phase"] = np.array(task.phase.value)
        # advance the "plant" along the best rollout for a few steps so x0 changes between plans
        best = int(np.argmax(ctrl.rewards))
        k = 2
        ctrl.current_state = ctrl.states[best, k - 1].copy()
        t += k * table["opt"]["timestep"]
    np.savez_compressed(os.path.join(OUT, f"plan_{tag}.npz"), **out)
    print("plan", tag, "rewards[0:3] step0:", out["p0_rewards"][:3], "nominal_out step2:", out["p2_nominal_out"].ravel()[:4])


if __name__ == "__main__":
    which = sys.argv[1:] or ["optimizers", "spline", "rewards", "plans"]
    if "optimizers" in which:
        gen_optimizers()
    if "spline" in which:
        gen_spline()
    if "rewards" in which:
        gen_rewards()
    if "plans" in which:
        gen_plan("cartpole_ps", "cartpole", "ps", 32, 1.28, 42)           # BASELINE config C1
        gen_plan("cartpole_mppi", "cartpole", "mppi", 64, 2.56, 43)       # C2 at a size the fixture can hold
        gen_plan("cylinder_push_cem", "cylinder_push", "cem", 48, 1.0, 44)  # C3, reduced N
    if "fr3" in which:
        gen_rewards_fr3()
        gen_plan("fr3_pick_cem", "fr3_pick", "cem", 12, 0.12, 46)          # §8f-2, reduced N and horizon (H = 30)
    if "leap" in which:
        gen_plan("leap_cube_mppi", "leap_cube", "mppi", 16, 0.4, 45)      # C4, reduced N
    print("golden written to", os.path.abspath(OUT))
